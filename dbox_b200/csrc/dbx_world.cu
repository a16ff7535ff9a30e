// dbx_world.cu — host side of the device world (see dbx_world.h).
#include "dbx_world.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace dbx {

static thread_local std::string g_lastError;
void set_last_error(const std::string& s) { g_lastError = s; }
const char* get_last_error() { return g_lastError.c_str(); }

#define CUDA_OR_FAIL(x, where) do { cudaError_t _e = (x); if (_e != cudaSuccess) return fail(_e, where); } while (0)

int World::fail(cudaError_t e, const char* where) {
  set_last_error(std::string(where) + ": " + cudaGetErrorString(e));
  return DBX_E_CUDA;
}

World::World(float gx, float gy, int device, const dbx_caps* caps) : gx_(gx), gy_(gy) {
  if (caps) caps_ = *caps;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { set_last_error("no CUDA device: dbox_b200 has no CPU fallback"); return; }
  if (device < 0 || device >= ndev) { set_last_error("bad device index"); return; }
  device_ = device;
  if (cudaSetDevice(device) != cudaSuccess) { set_last_error("cudaSetDevice failed"); return; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_last_error("cudaGetDeviceProperties failed"); return; }
  if (!prop.cooperativeLaunch) { set_last_error("device lacks cooperative launch"); return; }
  if (cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking) != cudaSuccess) { set_last_error("stream create failed"); return; }
  L_.sms = prop.multiProcessorCount;
  L_.gridWide = L_.sms * 8;
  L_.coopBlocks = L_.sms;
  L_.coopThreads = 512;
  L_.stream = stream_;
  if (hdr_.reserve(1, false, stream_) != cudaSuccess) { set_last_error("header alloc failed"); return; }
  for (auto& ev : ev_) cudaEventCreate(&ev);
  cudaStreamSynchronize(stream_);
  ok_ = true;
}

World::~World() {
  if (!ok_) return;
  cudaSetDevice(device_);
  cudaStreamSynchronize(stream_);
  if (aux_) { cudaStreamSynchronize(aux_); cudaStreamDestroy(aux_); cudaEventDestroy(evFork_); cudaEventDestroy(evJoin_); }
  if (wm_) cudaFreeHost(wm_);
  if (wmEv_) cudaEventDestroy(wmEv_);
  DevBuf<float4>* f4[] = {&b_xf, &b_xf0, &b_pos, &b_pos0, &b_vel, &b_force, &b_mass, &b_lc, &p_aabb, &p_fat, &bv_box, &c_m0, &c_m1, &c_imp, &c_mat,
                          &s_v0, &s_v1, &s_r0, &s_r1, &s_q0, &s_q1, &s_imp, &s_nm, &s_k, &s_p0, &s_p1, &s_p2, &j_anchor, &j_p0, &j_p1, &j_imp, &j_r, &j_lc, &j_m, &j_k0, &j_k1, &j_k2, &j_k3, &j_p2};
  for (auto* b : f4) b->release();
  j_ids2.release();
  b_toiMin.release(); b_toiOther.release(); b_acc.release(); jp_bits.release(); b_jmask.release();
  rowStage_.release(); rowIds_.release(); t_mass_.release();
  DevBuf<int>* i1[] = {&b_toiEvt, &b_toiFlags, &e_contact, &e_ncand, &e_cand, &bv_pos, &b_wake, &b_root, &b_islAwake, &b_islMinSleep, &b_posNotOk, &b_ovf, &b_world, &f_body, &f_group, &p_key, &moveList, &bv_leaf, &bv_leafAlt,
                       &bv_parent, &bv_visit, &c_toiList, &c_toiCount, &c_colour, &c_free, &c_work, &c_work2, &h_val, &s_contact, &s_hist, &s_pc, &s_root, &j_limit, &j_colour, &j_order, &j_root, &d_levels};
  for (auto* b : i1) b->release();
  b_gs.release(); f_mat.release(); s_p3.release(); b_flags.release(); f_filter.release(); p_flags.release(); c_flags.release();
  b_mask.release(); b_claim.release(); bv_key.release(); bv_keyAlt.release(); jp_keys.release(); c_key.release(); h_key.release();
  d_shapes.release(); p_ids.release(); c_ids.release(); c_fix.release(); j_ids.release(); bv_child.release(); bv_wr.release(); pairs.release(); s_body.release(); c_mk.release();
  for (int i = 0; i < 2; ++i) { inStage_[i].release(); outSnap_[i].release(); }
  ps_a_.release(); ps_b_.release(); ps_key_.release(); ev_a_.release(); ev_b_.release(); qIn_.release(); qOut_.release(); qCount_.release(); qPairs_.release();
  cubTemp.release(); hdr_.release();
  for (auto& ev : ev_) cudaEventDestroy(ev);
  cudaStreamDestroy(stream_);
}

// ------------------------------------------------------------------------------------------------ reference id order
// b2DynamicTree hands out node ids from a LIFO free list (collision/b2dynamictree.d:516-564).  Leaves and the internal
// node created with them always travel through that list in (leaf, parent) blocks, so the id a NEW leaf receives
// depends only on the history of leaf creations/destructions, not on the tree's shape.  Replaying that history keeps
// (min, max) proxy-id order — and with it fixture A/B of every contact — identical to the reference.
int World::allocProxyKey() {
  int id;
  if (!keyFree_.empty()) { id = keyFree_.back(); keyFree_.pop_back(); }
  else { id = keyFresh_; keyFresh_ += (keyLeaves_ >= 1) ? 2 : 1; }
  ++keyLeaves_;
  return id;
}
void World::freeProxyKey(int key) { keyFree_.push_back(key); --keyLeaves_; }

// ------------------------------------------------------------------------------------------------ shapes (host side)
int World::internShape(const DShape& s) {
  std::string k((const char*)&s, sizeof(DShape));
  auto it = shapeIndex_.find(k);
  if (it != shapeIndex_.end()) return it->second;
  int idx = (int)shapes_.size();
  shapes_.push_back(s);
  shapeIndex_.emplace(std::move(k), idx);
  return idx;
}

void World::buildChildShape(const HShape& hs, int child, DShape* out) const {
  std::memset(out, 0, sizeof(DShape));
  const dbx_shape& s = hs.s;
  out->radius = s.radius;
  switch (s.type) {
    case DBX_SHAPE_CIRCLE: out->type = SH_CIRCLE; out->c = V(s.p.x, s.p.y); break;
    case DBX_SHAPE_EDGE:
      out->type = SH_EDGE;
      out->v[0] = V(s.v0.x, s.v0.y); out->v[1] = V(s.v1.x, s.v1.y); out->v[2] = V(s.v2.x, s.v2.y); out->v[3] = V(s.v3.x, s.v3.y);
      out->flags = (s.hasV0 ? SHF_HAS_V0 : 0) | (s.hasV3 ? SHF_HAS_V3 : 0);
      break;
    case DBX_SHAPE_POLYGON:
      out->type = SH_POLYGON; out->count = s.count; out->c = V(s.centroid.x, s.centroid.y);
      for (int i = 0; i < s.count; ++i) { out->v[i] = V(s.vertices[i].x, s.vertices[i].y); out->n[i] = V(s.normals[i].x, s.normals[i].y); }
      break;
    case DBX_SHAPE_CHAIN: {
      // b2ChainShape.GetChildEdge (collision/shapes/b2chainshape.d:162-192)
      const auto& cv = hs.chain;
      const int cnt = (int)cv.size();
      out->type = SH_EDGE;
      out->flags = SHF_CHAIN_CHILD;
      out->v[1] = V(cv[child].x, cv[child].y);
      out->v[2] = V(cv[child + 1].x, cv[child + 1].y);
      if (child > 0) { out->v[0] = V(cv[child - 1].x, cv[child - 1].y); out->flags |= SHF_HAS_V0; }
      else { out->v[0] = V(s.prevVertex.x, s.prevVertex.y); if (s.hasPrev) out->flags |= SHF_HAS_V0; }
      if (child < cnt - 2) { out->v[3] = V(cv[child + 2].x, cv[child + 2].y); out->flags |= SHF_HAS_V3; }
      else { out->v[3] = V(s.nextVertex.x, s.nextVertex.y); if (s.hasNext) out->flags |= SHF_HAS_V3; }
    } break;
  }
}

// Shape.ComputeMass (b2circleshape.d:120-127, b2edgeshape.d:179-186, b2polygonshape.d:376-459, b2chainshape.d:247-254)
void World::computeMass(const HShape& hs, float density, float* mass, dbx_vec2* center, float* I) const {
  const dbx_shape& s = hs.s;
  if (s.type == DBX_SHAPE_CIRCLE) {
    *mass = density * kPi * s.radius * s.radius;
    *center = s.p;
    *I = *mass * (0.5f * s.radius * s.radius + (s.p.x * s.p.x + s.p.y * s.p.y));
  } else if (s.type == DBX_SHAPE_EDGE) {
    *mass = 0.0f; center->x = 0.5f * (s.v1.x + s.v2.x); center->y = 0.5f * (s.v1.y + s.v2.y); *I = 0.0f;
  } else if (s.type == DBX_SHAPE_CHAIN) {
    *mass = 0.0f; center->x = 0.0f; center->y = 0.0f; *I = 0.0f;
  } else {
    const int n = s.count;
    v2 c = V(0.0f, 0.0f), ref = V(0.0f, 0.0f);
    float area = 0.0f, inertia = 0.0f;
    for (int i = 0; i < n; ++i) ref += V(s.vertices[i].x, s.vertices[i].y);
    ref *= 1.0f / n;
    const float k_inv3 = 1.0f / 3.0f;
    for (int i = 0; i < n; ++i) {
      v2 e1 = V(s.vertices[i].x, s.vertices[i].y) - ref;
      v2 e2 = i + 1 < n ? V(s.vertices[i + 1].x, s.vertices[i + 1].y) - ref : V(s.vertices[0].x, s.vertices[0].y) - ref;
      float D = cross(e1, e2);
      float triangleArea = 0.5f * D;
      area += triangleArea;
      c += triangleArea * k_inv3 * (e1 + e2);
      float intx2 = e1.x * e1.x + e2.x * e1.x + e2.x * e2.x;
      float inty2 = e1.y * e1.y + e2.y * e1.y + e2.y * e2.y;
      inertia += (0.25f * k_inv3 * D) * (intx2 + inty2);
    }
    *mass = density * area;
    c *= 1.0f / area;
    v2 ctr = c + ref;
    center->x = ctr.x; center->y = ctr.y;
    *I = density * inertia;
    *I += *mass * (dot(ctr, ctr) - dot(c, c));
  }
}

// b2Body.ResetMassData (dynamics/b2body.d:555-625); fixtures are visited newest first like the reference's list
void World::resetMassData(HBody& hb) {
  dbx_body_state& st = hb.st;
  st.mass = 0.0f; st.invMass = 0.0f; st.I = 0.0f; st.invI = 0.0f;
  st.localCenter = dbx_vec2{0.0f, 0.0f};
  if (st.type == DBX_STATIC_BODY || st.type == DBX_KINEMATIC_BODY) {
    st.c0 = st.p; st.c = st.p; st.a0 = st.a;
    return;
  }
  v2 localCenter = V(0.0f, 0.0f);
  for (auto it = hb.fixtures.rbegin(); it != hb.fixtures.rend(); ++it) {
    const HFixture& f = fixtures_[*it];
    if (f.def.density == 0.0f) continue;
    float m, I; dbx_vec2 c;
    computeMass(f.shape, f.def.density, &m, &c, &I);
    st.mass += m;
    localCenter += m * V(c.x, c.y);
    st.I += I;
  }
  if (st.mass > 0.0f) { st.invMass = 1.0f / st.mass; localCenter *= st.invMass; }
  else { st.mass = 1.0f; st.invMass = 1.0f; }
  if (st.I > 0.0f && (st.flags & DBX_BODY_FIXED_ROTATION) == 0) {
    st.I -= st.mass * dot(localCenter, localCenter);
    st.invI = 1.0f / st.I;
  } else { st.I = 0.0f; st.invI = 0.0f; }
  v2 oldCenter = V(st.c.x, st.c.y);
  st.localCenter = dbx_vec2{localCenter.x, localCenter.y};
  Xf xf; xf.p = V(st.p.x, st.p.y); xf.q = R(st.qs, st.qc);
  v2 c = mul(xf, localCenter);
  st.c0 = st.c = dbx_vec2{c.x, c.y};
  v2 dv = cross(st.w, c - oldCenter);
  st.v.x += dv.x; st.v.y += dv.y;
}

void World::wake(HBody& hb, bool flag) {  // b2Body.SetAwake (b2body.d:827-846)
  dbx_body_state& st = hb.st;
  if (flag) {
    if ((st.flags & DBX_BODY_AWAKE) == 0) { st.flags |= DBX_BODY_AWAKE; st.sleepTime = 0.0f; }
  } else {
    st.flags &= ~DBX_BODY_AWAKE; st.sleepTime = 0.0f;
    st.v = dbx_vec2{0, 0}; st.w = 0.0f; st.force = dbx_vec2{0, 0}; st.torque = 0.0f;
  }
}

// ------------------------------------------------------------------------------------------------ lifecycle
int World::createBody(const dbx_body_def& d) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  HBody hb;
  hb.alive = true;
  dbx_body_state& st = hb.st;
  std::memset(&st, 0, sizeof(st));
  // b2Body ctor (dynamics/b2body.d:1030-1115)
  st.type = d.type;
  st.flags = 0;
  if (d.bullet) st.flags |= DBX_BODY_BULLET;
  if (d.fixedRotation) st.flags |= DBX_BODY_FIXED_ROTATION;
  if (d.allowSleep) st.flags |= DBX_BODY_AUTOSLEEP;
  if (d.awake) st.flags |= DBX_BODY_AWAKE;
  if (d.active) st.flags |= DBX_BODY_ACTIVE;
  st.p = d.position;
  Rot q = rot_from_angle(d.angle);
  st.qs = q.s; st.qc = q.c;
  st.c0 = d.position; st.c = d.position; st.a0 = d.angle; st.a = d.angle; st.alpha0 = 0.0f;
  st.v = d.linearVelocity; st.w = d.angularVelocity;
  st.linearDamping = d.linearDamping; st.angularDamping = d.angularDamping; st.gravityScale = d.gravityScale;
  if (d.type == DBX_DYNAMIC_BODY) { st.mass = 1.0f; st.invMass = 1.0f; }
  hb.xf0 = make_float4(st.p.x, st.p.y, st.qs, st.qc);
  hb.world = 0;
  bodies_.push_back(std::move(hb));
  return (int)bodies_.size() - 1;
}

int World::createFixture(int b, const dbx_fixture_def& d, const dbx_shape& s) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (b < 0 || b >= (int)bodies_.size() || !bodies_[b].alive) return DBX_E_INVALID;
  if ((size_t)b < bodiesSynced_) { int rc = pullBodies(); if (rc < 0) return rc; fullPushBodies_ = true; }
  HBody& hb = bodies_[b];
  HFixture f;
  f.alive = true; f.body = b; f.def = d; f.shape.s = s;
  if (s.type == DBX_SHAPE_CHAIN) {
    if (!s.chainVertices || s.chainCount < 2) return DBX_E_INVALID;
    f.shape.chain.assign(s.chainVertices, s.chainVertices + s.chainCount);
    f.shape.s.chainVertices = nullptr;
  }
  const int fid = (int)fixtures_.size();
  const int childCount = s.type == DBX_SHAPE_CHAIN ? s.chainCount - 1 : 1;
  (void)childCount;
  fixtures_.push_back(std::move(f));
  hb.fixtures.push_back(fid);
  if (hb.st.flags & DBX_BODY_ACTIVE) { int rc = createProxiesFor(fid); if (rc < 0) return rc; }
  if (d.density > 0.0f) resetMassData(bodies_[b]);
  newFixture_ = true;
  return fid;
}

// b2Fixture.CreateProxies (b2fixture.d:450-465) -> b2BroadPhase.CreateProxy (b2broadphase.d:78-84)
int World::createProxiesFor(int fid) {
  HFixture& f = fixtures_[fid];
  const int b = f.body;
  const HBody& hb = bodies_[b];
  const int childCount = f.shape.s.type == DBX_SHAPE_CHAIN ? (int)f.shape.chain.size() - 1 : 1;
  Xf xf; xf.p = V(hb.st.p.x, hb.st.p.y); xf.q = R(hb.st.qs, hb.st.qc);
  for (int i = 0; i < childCount; ++i) {
    DShape ds; buildChildShape(f.shape, i, &ds);
    HProxy p;
    p.alive = true; p.fixture = fid; p.child = i; p.body = b; p.shape = internShape(ds);
    Box box = shape_aabb(&ds, xf);
    p.aabb = pack(box);
    p.fat = make_float4(box.lo.x - kAabbExtension, box.lo.y - kAabbExtension, box.hi.x + kAabbExtension, box.hi.y + kAabbExtension);
    p.key = allocProxyKey();
    p.flags = PF_ALIVE | PF_MOVED;
    int slot;
    if (!proxyFree_.empty()) {
      slot = proxyFree_.back(); proxyFree_.pop_back();
      if ((size_t)slot < proxiesSynced_) { int rc = pullProxies(); if (rc < 0) return rc; fullPushProxies_ = true; }
      proxies_[slot] = p;
    } else { slot = (int)proxies_.size(); proxies_.push_back(p); }
    f.proxies.push_back(slot);
    pendingMoves_.push_back(slot);
  }
  return 0;
}

// b2Body.SetType (dynamics/b2body.d:867-914)
int World::setBodyType(int b, int type) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (b < 0 || b >= (int)bodies_.size() || !bodies_[b].alive || type < DBX_STATIC_BODY || type > DBX_DYNAMIC_BODY) return DBX_E_INVALID;
  if (bodies_[b].st.type == type) return 0;
  int rc = destroyContactsWhere(b, -1, -1, false); if (rc < 0) return rc;     // :896-903 (ends touching contacts, wakes the other bodies)
  HBody* hb = mutBody(b); if (!hb) return DBX_E_INVALID;
  rc = pullProxies(); if (rc < 0) return rc;
  dbx_body_state& st = hb->st;
  st.type = type;
  resetMassData(*hb);
  if (type == DBX_STATIC_BODY) {
    st.v = dbx_vec2{0, 0}; st.w = 0.0f; st.a0 = st.a; st.c0 = st.c;
    Xf xf; xf.p = V(st.p.x, st.p.y); xf.q = R(st.qs, st.qc);
    hb->xf0 = pack(xf);
    for (auto it = hb->fixtures.rbegin(); it != hb->fixtures.rend(); ++it) for (int slot : fixtures_[*it].proxies) {     // SynchronizeFixtures with xf1 == xf2
      HProxy& p = proxies_[slot];
      Box box = shape_aabb(&shapes_[p.shape], xf);
      p.aabb = pack(box);
      if (!contains(BX(p.fat), box)) { p.fat = make_float4(box.lo.x - kAabbExtension, box.lo.y - kAabbExtension, box.hi.x + kAabbExtension, box.hi.y + kAabbExtension); pendingMoves_.push_back(slot); }
    }
  }
  wake(*hb, true);
  st.force = dbx_vec2{0, 0}; st.torque = 0.0f;
  for (auto it = hb->fixtures.rbegin(); it != hb->fixtures.rend(); ++it) for (int slot : fixtures_[*it].proxies) pendingMoves_.push_back(slot);   // TouchProxy (:905-913)
  fullPushProxies_ = true;
  if (!hb->joints.empty()) jointsChanged_ = true;      // joint colouring looks at body types
  return 0;
}

// b2Body.SetMassData (dynamics/b2body.d:502-540)
int World::setMassData(int b, float mass, float cx, float cy, float I) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  HBody* hb = mutBody(b); if (!hb) return DBX_E_INVALID;
  dbx_body_state& st = hb->st;
  if (st.type != DBX_DYNAMIC_BODY) return 0;
  st.invMass = 0.0f; st.I = 0.0f; st.invI = 0.0f;
  st.mass = mass;
  if (st.mass <= 0.0f) st.mass = 1.0f;
  st.invMass = 1.0f / st.mass;
  if (I > 0.0f && (st.flags & DBX_BODY_FIXED_ROTATION) == 0) { st.I = I - st.mass * (cx * cx + cy * cy); st.invI = 1.0f / st.I; }
  const v2 oldCenter = V(st.c.x, st.c.y);
  st.localCenter = dbx_vec2{cx, cy};
  Xf xf; xf.p = V(st.p.x, st.p.y); xf.q = R(st.qs, st.qc);
  const v2 c = mul(xf, V(cx, cy));
  st.c0 = st.c = dbx_vec2{c.x, c.y};
  const v2 dv = cross(st.w, c - oldCenter);
  st.v.x += dv.x; st.v.y += dv.y;
  return 0;
}
int World::resetMass(int b) {                         // b2Body.ResetMassData (b2body.d:555-625)
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  HBody* hb = mutBody(b); if (!hb) return DBX_E_INVALID;
  resetMassData(*hb);
  return 0;
}
int World::setFixedRotation(int b, bool flag) {       // b2body.d:924-945
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  HBody* hb = mutBody(b); if (!hb) return DBX_E_INVALID;
  if (flag == ((hb->st.flags & DBX_BODY_FIXED_ROTATION) != 0)) return 0;
  if (flag) hb->st.flags |= DBX_BODY_FIXED_ROTATION; else hb->st.flags &= ~DBX_BODY_FIXED_ROTATION;
  hb->st.w = 0.0f;
  resetMassData(*hb);
  return 0;
}
int World::setBodyScalars(int b, const float* linearDamping, const float* angularDamping, const float* gravityScale) {   // b2body.d:653-690
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  HBody* hb = mutBody(b); if (!hb) return DBX_E_INVALID;
  if (linearDamping) hb->st.linearDamping = *linearDamping;
  if (angularDamping) hb->st.angularDamping = *angularDamping;
  if (gravityScale) hb->st.gravityScale = *gravityScale;
  return 0;
}
// b2Fixture.SetFilterData + Refilter (dynamics/b2fixture.d:131-178): flag the fixture's contacts, touch its proxies
int World::setFixtureFilter(int f, int category, int mask, int group) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (f < 0 || f >= (int)fixtures_.size() || !fixtures_[f].alive) return DBX_E_INVALID;
  fixtures_[f].def.categoryBits = (uint16_t)category; fixtures_[f].def.maskBits = (uint16_t)mask; fixtures_[f].def.groupIndex = (int16_t)group;
  fullPushFixtures_ = true;
  int rc = destroyContactsWhere(-1, f, -1, true); if (rc < 0) return rc;       // FlagForFiltering (pushes the new filter first)
  for (int slot : fixtures_[f].proxies) pendingMoves_.push_back(slot);          // TouchProxy
  return 0;
}
// b2Fixture.SetSensor (b2fixture.d:108-115); contacts cache the sensor bit, so the fixture's contacts are re-flagged
int World::setFixtureSensor(int f, bool flag) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (f < 0 || f >= (int)fixtures_.size() || !fixtures_[f].alive) return DBX_E_INVALID;
  if (flag == (fixtures_[f].def.isSensor != 0)) return 0;
  HBody* hb = mutBody(fixtures_[f].body); if (!hb) return DBX_E_INVALID;
  wake(*hb, true);
  fixtures_[f].def.isSensor = flag ? 1 : 0;
  fullPushFixtures_ = true;
  int rc = push(); if (rc < 0) return rc;
  CUDA_OR_FAIL(launch_api_resensor(dw_, L_, f), "resensor");
  return 0;
}
// b2Fixture.SetFriction / SetRestitution / SetDensity (b2fixture.d:225-262): existing contacts keep their mixed values,
// the density only counts at the next ResetMassData
int World::setFixtureMaterial(int f, const float* friction, const float* restitution, const float* density) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (f < 0 || f >= (int)fixtures_.size() || !fixtures_[f].alive) return DBX_E_INVALID;
  if (friction) fixtures_[f].def.friction = *friction;
  if (restitution) fixtures_[f].def.restitution = *restitution;
  if (density) fixtures_[f].def.density = *density;
  fullPushFixtures_ = true;
  return 0;
}

// b2Body.SetActive (dynamics/b2body.d:718-775)
int World::setBodyActive(int b, bool flag) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (b < 0 || b >= (int)bodies_.size() || !bodies_[b].alive) return DBX_E_INVALID;
  if (flag == ((bodies_[b].st.flags & DBX_BODY_ACTIVE) != 0)) return 0;
  int rc = 0;
  if (!flag) { rc = destroyContactsWhere(b, -1, -1, false); if (rc < 0) return rc; }
  HBody* hb = mutBody(b); if (!hb) return DBX_E_INVALID;
  rc = pullProxies(); if (rc < 0) return rc;
  if (flag) {
    hb->st.flags |= DBX_BODY_ACTIVE;
    const std::vector<int> fx(hb->fixtures.rbegin(), hb->fixtures.rend());      // m_fixtureList order: newest first
    for (int fid : fx) { rc = createProxiesFor(fid); if (rc < 0) return rc; }
  } else {
    hb->st.flags &= ~DBX_BODY_ACTIVE;
    for (auto it = hb->fixtures.rbegin(); it != hb->fixtures.rend(); ++it) {
      HFixture& f = fixtures_[*it];
      for (int slot : f.proxies) {
        freeProxyKey(proxies_[slot].key);
        proxies_[slot].alive = false; proxies_[slot].flags = 0;
        proxyFree_.push_back(slot);
        pendingMoves_.erase(std::remove(pendingMoves_.begin(), pendingMoves_.end(), slot), pendingMoves_.end());
      }
      f.proxies.clear();
    }
  }
  fullPushBodies_ = true; fullPushProxies_ = true;
  return 0;
}

int World::destroyContactsWhere(int body, int fixture, int otherBody, bool flagOnly) {
  int rc = push(); if (rc < 0) return rc;
  CUDA_OR_FAIL(launch_api_contacts(dw_, L_, body, fixture, otherBody, flagOnly ? 1 : 0), "api_contacts");
  hostBodiesValid_ = false, ++bodyEpoch_;
  return 0;
}

int World::destroyFixture(int fid) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (fid < 0 || fid >= (int)fixtures_.size() || !fixtures_[fid].alive) return DBX_E_INVALID;
  int rc = destroyContactsWhere(-1, fid, -1, false); if (rc < 0) return rc;
  rc = pullBodies(); if (rc < 0) return rc;
  rc = pullProxies(); if (rc < 0) return rc;
  HFixture& f = fixtures_[fid];
  HBody& hb = bodies_[f.body];
  for (int slot : f.proxies) {
    freeProxyKey(proxies_[slot].key);
    proxies_[slot].alive = false; proxies_[slot].flags = 0;
    proxyFree_.push_back(slot);
    pendingMoves_.erase(std::remove(pendingMoves_.begin(), pendingMoves_.end(), slot), pendingMoves_.end());
  }
  f.proxies.clear();
  f.alive = false;
  hb.fixtures.erase(std::remove(hb.fixtures.begin(), hb.fixtures.end(), fid), hb.fixtures.end());
  resetMassData(hb);
  fullPushBodies_ = true; fullPushProxies_ = true;
  return 0;
}

int World::destroyBody(int b) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (b < 0 || b >= (int)bodies_.size() || !bodies_[b].alive) return DBX_E_INVALID;
  std::vector<int> js = bodies_[b].joints;
  for (int j : js) destroyJoint(j);
  int rc = destroyContactsWhere(b, -1, -1, false); if (rc < 0) return rc;
  rc = pullBodies(); if (rc < 0) return rc;
  rc = pullProxies(); if (rc < 0) return rc;
  HBody& hb = bodies_[b];
  for (auto it = hb.fixtures.rbegin(); it != hb.fixtures.rend(); ++it) {   // newest first (b2world.d:148-167)
    HFixture& f = fixtures_[*it];
    for (int slot : f.proxies) {
      freeProxyKey(proxies_[slot].key);
      proxies_[slot].alive = false; proxies_[slot].flags = 0;
      proxyFree_.push_back(slot);
      pendingMoves_.erase(std::remove(pendingMoves_.begin(), pendingMoves_.end(), slot), pendingMoves_.end());
    }
    f.proxies.clear(); f.alive = false;
  }
  hb.fixtures.clear();
  hb.alive = false;
  hb.st.flags = 0;
  fullPushBodies_ = true; fullPushProxies_ = true;
  return 0;
}

int World::createJoint(const dbx_joint_def& d) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (d.type < DBX_JOINT_REVOLUTE || d.type > DBX_JOINT_MOTOR) return DBX_E_INVALID;
  if (d.type == DBX_JOINT_GEAR) return createGearJoint(d);
  if (d.bodyA < 0 || d.bodyB < 0 || d.bodyA >= (int)bodies_.size() || d.bodyB >= (int)bodies_.size() || !bodies_[d.bodyA].alive || !bodies_[d.bodyB].alive || d.bodyA == d.bodyB) return DBX_E_INVALID;
  if (d.type == DBX_JOINT_PULLEY && d.ratio == 0.0f) return DBX_E_INVALID;   // b2pulleyjoint.d:118
  HJoint j; j.alive = true; j.def = d;
  if (d.type == DBX_JOINT_PRISMATIC) {       // b2prismaticjoint.d:173-174: the axis is stored normalised
    v2 ax = V(d.localAxisA.x, d.localAxisA.y); normalize(ax);
    j.def.localAxisA = dbx_vec2{ax.x, ax.y};
  } else if (d.type == DBX_JOINT_MOTOR) {    // b2motorjoint.d:230-231: rA = qA * (-localCenterA), i.e. anchors at the body origins
    j.def.localAnchorA = dbx_vec2{0.0f, 0.0f}; j.def.localAnchorB = dbx_vec2{0.0f, 0.0f};
  } else if (d.type == DBX_JOINT_MOUSE) {    // b2mousejoint.d:70-71: localAnchorB = b2MulT(bodyB.GetTransform(), target)
    int rcp = pullBodies(); if (rcp < 0) return rcp;
    const dbx_body_state& st = bodies_[d.bodyB].st;
    Xf xf; xf.p = V(st.p.x, st.p.y); xf.q = R(st.qs, st.qc);
    const v2 la = mulT(xf, V(d.target.x, d.target.y));
    j.def.localAnchorB = dbx_vec2{la.x, la.y};
  }
  joints_.push_back(j);
  const int jid = (int)joints_.size() - 1;
  bodies_[d.bodyA].joints.push_back(jid);
  bodies_[d.bodyB].joints.push_back(jid);
  jointsChanged_ = true;
  if (!d.collideConnected && bodiesSynced_ > 0 && (size_t)std::max(d.bodyA, d.bodyB) < bodiesSynced_) {
    int rc = destroyContactsWhere(d.bodyA, -1, d.bodyB, true); if (rc < 0) return rc;   // FlagForFiltering (b2world.d:241-256)
  }
  return jid;
}

// b2GearJoint (b2gearjoint.d:84-160): bodies A / B are the second bodies of joint1 / joint2, C / D their first bodies; the
// anchors, axes and reference angles are copied from those joints and the constant is measured on the current pose.
// The record travels in otherwise unused fields of the joint def: groundAnchorA/B = localAnchorC/D, localAxisA = localAxisC,
// linearOffset = localAxisD, referenceAngle / angularOffset = referenceAngleA / B, lengthA = constant.
int World::createGearJoint(const dbx_joint_def& d) {
  const int nJ = (int)joints_.size();
  if (d.joint1 < 0 || d.joint2 < 0 || d.joint1 >= nJ || d.joint2 >= nJ || !joints_[d.joint1].alive || !joints_[d.joint2].alive || d.joint1 == d.joint2) return DBX_E_INVALID;
  const dbx_joint_def& j1 = joints_[d.joint1].def; const dbx_joint_def& j2 = joints_[d.joint2].def;
  auto ok = [](int t) { return t == DBX_JOINT_REVOLUTE || t == DBX_JOINT_PRISMATIC; };
  if (!ok(j1.type) || !ok(j2.type)) { set_last_error("gear: joint1 / joint2 must be revolute or prismatic"); return DBX_E_INVALID; }
  int rc = pullBodies(); if (rc < 0) return rc;
  HJoint j; j.alive = true; j.def = d;
  j.typeA = j1.type; j.typeB = j2.type; j.bodyC = j1.bodyA; j.bodyD = j2.bodyA;
  j.def.bodyA = j1.bodyB; j.def.bodyB = j2.bodyB;
  if (j.def.bodyA == j.def.bodyB) return DBX_E_INVALID;
  auto pose = [&](int b, Xf* xf, float* a) { const dbx_body_state& st = bodies_[b].st; xf->p = V(st.p.x, st.p.y); xf->q = R(st.qs, st.qc); *a = st.a; };
  auto coordinate = [&](const dbx_joint_def& jj, int bFar, int bNear, dbx_vec2* ancFar, dbx_vec2* ancNear, dbx_vec2* axisFar, float* ref) -> float {
    Xf xfN, xfF; float aN, aF;
    pose(bNear, &xfN, &aN); pose(bFar, &xfF, &aF);
    *ancFar = jj.localAnchorA; *ancNear = jj.localAnchorB; *ref = jj.referenceAngle;
    if (jj.type == DBX_JOINT_REVOLUTE) { *axisFar = dbx_vec2{0.0f, 0.0f}; return aN - aF - jj.referenceAngle; }
    *axisFar = jj.localAxisA;      // stored normalised at creation
    const v2 pF = V(jj.localAnchorA.x, jj.localAnchorA.y);
    const v2 pN = mulT(xfF.q, mul(xfN.q, V(jj.localAnchorB.x, jj.localAnchorB.y)) + (xfN.p - xfF.p));
    return dot(pN - pF, V(jj.localAxisA.x, jj.localAxisA.y));
  };
  const float coordinateA = coordinate(j1, j.bodyC, j.def.bodyA, &j.def.groundAnchorA, &j.def.localAnchorA, &j.def.localAxisA, &j.def.referenceAngle);
  const float coordinateB = coordinate(j2, j.bodyD, j.def.bodyB, &j.def.groundAnchorB, &j.def.localAnchorB, &j.def.linearOffset, &j.def.angularOffset);
  j.def.lengthA = coordinateA + d.ratio * coordinateB;
  joints_.push_back(j);
  const int jid = (int)joints_.size() - 1;
  bodies_[j.def.bodyA].joints.push_back(jid);
  bodies_[j.def.bodyB].joints.push_back(jid);
  jointsChanged_ = true;
  if (!d.collideConnected && bodiesSynced_ > 0 && (size_t)std::max(j.def.bodyA, j.def.bodyB) < bodiesSynced_) {
    rc = destroyContactsWhere(j.def.bodyA, -1, j.def.bodyB, true); if (rc < 0) return rc;
  }
  return jid;
}

int World::destroyJoint(int jid) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (jid < 0 || jid >= (int)joints_.size() || !joints_[jid].alive) return DBX_E_INVALID;
  int rc = pullJoints(); if (rc < 0) return rc;
  HJoint& j = joints_[jid];
  const int a = j.def.bodyA, b = j.def.bodyB;
  j.alive = false;
  auto& ja = bodies_[a].joints; ja.erase(std::remove(ja.begin(), ja.end(), jid), ja.end());
  auto& jb = bodies_[b].joints; jb.erase(std::remove(jb.begin(), jb.end(), jid), jb.end());
  jointsChanged_ = true; fullPushJoints_ = true;
  rc = push(); if (rc < 0) return rc;
  CUDA_OR_FAIL(launch_api_wake(dw_, L_, a, b), "api_wake");          // b2world.d:297-298
  hostBodiesValid_ = false, ++bodyEpoch_;
  if (!j.def.collideConnected) { rc = destroyContactsWhere(a, -1, b, true); if (rc < 0) return rc; }
  return 0;
}

int World::setFlags(uint32_t f) {
  if ((f & DBX_WORLD_SUB_STEPPING) && replicated_) { set_last_error("sub-stepping a replicated world is not built (one TOI event per Step would be one over ALL replicas)"); return DBX_E_UNSUPPORTED; }
  if ((flags_ & DBX_WORLD_ALLOW_SLEEP) && !(f & DBX_WORLD_ALLOW_SLEEP)) {
    // b2World.SetAllowSleeping(false) wakes every body (b2world.d:622-640)
    int rc = pullBodies(); if (rc < 0) return rc;
    for (auto& hb : bodies_) if (hb.alive) wake(hb, true);
    fullPushBodies_ = true;
  }
  flags_ = f;
  return 0;
}

// ------------------------------------------------------------------------------------------------ mirroring
static float4 f4(float a, float b, float c, float d) { return make_float4(a, b, c, d); }

int World::pullBodies() {
  if (hostBodiesValid_ || bodiesSynced_ == 0) { hostBodiesValid_ = true; return 0; }
  const size_t n = bodiesSynced_;
  std::vector<float4> xf(n), xf0(n), pos(n), pos0(n), vel(n), frc(n), ms(n), lc(n);
  std::vector<float2> gs(n); std::vector<uint32_t> fl(n);
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(xf.data(), b_xf.p, n * 16, cudaMemcpyDeviceToHost), "pull xf");
  CUDA_OR_FAIL(cudaMemcpy(xf0.data(), b_xf0.p, n * 16, cudaMemcpyDeviceToHost), "pull xf0");
  CUDA_OR_FAIL(cudaMemcpy(pos.data(), b_pos.p, n * 16, cudaMemcpyDeviceToHost), "pull pos");
  CUDA_OR_FAIL(cudaMemcpy(pos0.data(), b_pos0.p, n * 16, cudaMemcpyDeviceToHost), "pull pos0");
  CUDA_OR_FAIL(cudaMemcpy(vel.data(), b_vel.p, n * 16, cudaMemcpyDeviceToHost), "pull vel");
  CUDA_OR_FAIL(cudaMemcpy(frc.data(), b_force.p, n * 16, cudaMemcpyDeviceToHost), "pull force");
  CUDA_OR_FAIL(cudaMemcpy(ms.data(), b_mass.p, n * 16, cudaMemcpyDeviceToHost), "pull mass");
  CUDA_OR_FAIL(cudaMemcpy(lc.data(), b_lc.p, n * 16, cudaMemcpyDeviceToHost), "pull lc");
  CUDA_OR_FAIL(cudaMemcpy(gs.data(), b_gs.p, n * 8, cudaMemcpyDeviceToHost), "pull gs");
  CUDA_OR_FAIL(cudaMemcpy(fl.data(), b_flags.p, n * 4, cudaMemcpyDeviceToHost), "pull flags");
  for (size_t i = 0; i < n; ++i) {
    HBody& hb = bodies_[i];
    if (!hb.alive) continue;
    if (hb.dirty) continue;             // edited on the host since (mutBodyRow): the host row is the newer one
    dbx_body_state& st = hb.st;
    st.p = dbx_vec2{xf[i].x, xf[i].y}; st.qs = xf[i].z; st.qc = xf[i].w;
    hb.xf0 = xf0[i];
    st.c = dbx_vec2{pos[i].x, pos[i].y}; st.a = pos[i].z;
    st.c0 = dbx_vec2{pos0[i].x, pos0[i].y}; st.a0 = pos0[i].z; st.alpha0 = pos0[i].w;
    st.v = dbx_vec2{vel[i].x, vel[i].y}; st.w = vel[i].z;
    st.force = dbx_vec2{frc[i].x, frc[i].y}; st.torque = frc[i].z;
    st.invMass = ms[i].x; st.invI = ms[i].y; st.mass = ms[i].z; st.I = ms[i].w;
    st.localCenter = dbx_vec2{lc[i].x, lc[i].y}; st.linearDamping = lc[i].z; st.angularDamping = lc[i].w;
    st.gravityScale = gs[i].x; st.sleepTime = gs[i].y;
    st.flags = fl[i] & 0xFFFF; st.type = body_type(fl[i]);
  }
  hostBodiesValid_ = true;
  return 0;
}

// one body's row (see bodyEpoch_ in dbx_world.h)
int World::pullBodyRow(int b) {
  HBody& hb = bodies_[b];
  if (hostBodiesValid_ || (size_t)b >= bodiesSynced_ || hb.dirty || hb.validEpoch == bodyEpoch_) return 0;
  // a program that reads many bodies after a step (drawing all of them) is better served by the bulk copy: past 16 single rows
  // since the device last changed, fetch everything once
  if (rowPullEpoch_ != bodyEpoch_) { rowPullEpoch_ = bodyEpoch_; rowPulls_ = 0; }
  if (++rowPulls_ > 16) return pullBodies();
  CUDA_OR_FAIL(rowStage_.reserve(64 * 9, false, stream_), "row stage"); CUDA_OR_FAIL(rowIds_.reserve(64, false, stream_), "row ids");
  float4 r[9];
  CUDA_OR_FAIL(cudaMemcpyAsync(rowIds_.p, &b, 4, cudaMemcpyHostToDevice, stream_), "row id");
  CUDA_OR_FAIL(launch_body_rows(dw_, L_, rowIds_.p, rowStage_.p, 1, false), "row get");
  CUDA_OR_FAIL(cudaMemcpyAsync(r, rowStage_.p, sizeof(r), cudaMemcpyDeviceToHost, stream_), "row d2h");
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  const float4 xf = r[0], xf0 = r[1], pos = r[2], pos0 = r[3], vel = r[4], frc = r[5], ms = r[6], lc = r[7];
  const float2 gs = make_float2(r[8].x, r[8].y);
  uint32_t fl; std::memcpy(&fl, &r[8].z, 4);
  dbx_body_state& st = hb.st;
  st.p = dbx_vec2{xf.x, xf.y}; st.qs = xf.z; st.qc = xf.w;
  hb.xf0 = xf0;
  st.c = dbx_vec2{pos.x, pos.y}; st.a = pos.z;
  st.c0 = dbx_vec2{pos0.x, pos0.y}; st.a0 = pos0.z; st.alpha0 = pos0.w;
  st.v = dbx_vec2{vel.x, vel.y}; st.w = vel.z;
  st.force = dbx_vec2{frc.x, frc.y}; st.torque = frc.z;
  st.invMass = ms.x; st.invI = ms.y; st.mass = ms.z; st.I = ms.w;
  st.localCenter = dbx_vec2{lc.x, lc.y}; st.linearDamping = lc.z; st.angularDamping = lc.w;
  st.gravityScale = gs.x; st.sleepTime = gs.y;
  st.flags = fl & 0xFFFF; st.type = body_type(fl);
  hb.validEpoch = bodyEpoch_;
  return 0;
}
int World::pushBodyRows() {
  const int n = (int)dirtyBodies_.size();
  if (n == 0) return 0;
  CUDA_OR_FAIL(rowStage_.reserve(64 * 9, false, stream_), "row stage"); CUDA_OR_FAIL(rowIds_.reserve(64, false, stream_), "row ids");
  std::vector<float4> rows((size_t)n * 9);
  for (int k = 0; k < n; ++k) {
    const HBody& hb = bodies_[dirtyBodies_[k]];
    const dbx_body_state& st = hb.st;
    float4* r = rows.data() + (size_t)k * 9;
    const uint32_t fl = hb.alive ? (uint32_t)((st.flags & 0xFFFF) | ((uint32_t)st.type << BF_TYPE_SHIFT) | BF_ALIVE) : (uint32_t)0;
    r[0] = f4(st.p.x, st.p.y, st.qs, st.qc); r[1] = hb.xf0; r[2] = f4(st.c.x, st.c.y, st.a, 0.0f); r[3] = f4(st.c0.x, st.c0.y, st.a0, st.alpha0);
    r[4] = f4(st.v.x, st.v.y, st.w, 0.0f); r[5] = f4(st.force.x, st.force.y, st.torque, 0.0f); r[6] = f4(st.invMass, st.invI, st.mass, st.I);
    r[7] = f4(st.localCenter.x, st.localCenter.y, st.linearDamping, st.angularDamping);
    r[8] = f4(st.gravityScale, st.sleepTime, 0.0f, 0.0f); std::memcpy(&r[8].z, &fl, 4);
  }
  // (pageable sources: both copies are complete on return, the vectors may go)
  CUDA_OR_FAIL(cudaMemcpy(rowIds_.p, dirtyBodies_.data(), (size_t)n * 4, cudaMemcpyHostToDevice), "row ids");
  CUDA_OR_FAIL(cudaMemcpy(rowStage_.p, rows.data(), rows.size() * 16, cudaMemcpyHostToDevice), "rows h2d");
  CUDA_OR_FAIL(launch_body_rows(dw_, L_, rowIds_.p, rowStage_.p, n, true), "rows set");
  return 0;
}

int World::pullProxies() {
  if (hostProxiesValid_ || proxiesSynced_ == 0) { hostProxiesValid_ = true; return 0; }
  const size_t n = proxiesSynced_;
  std::vector<float4> aabb(n), fat(n); std::vector<uint32_t> fl(n);
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(aabb.data(), p_aabb.p, n * 16, cudaMemcpyDeviceToHost), "pull aabb");
  CUDA_OR_FAIL(cudaMemcpy(fat.data(), p_fat.p, n * 16, cudaMemcpyDeviceToHost), "pull fat");
  CUDA_OR_FAIL(cudaMemcpy(fl.data(), p_flags.p, n * 4, cudaMemcpyDeviceToHost), "pull pflags");
  for (size_t i = 0; i < n; ++i) { if (!proxies_[i].alive) continue; proxies_[i].aabb = aabb[i]; proxies_[i].fat = fat[i]; proxies_[i].flags = fl[i] & ~PF_MOVED; }
  hostProxiesValid_ = true;
  return 0;
}

int World::pullJoints() {
  const size_t n = jointAt_.size();
  if (hostJointsValid_ || n == 0 || !dw_.hdr) { hostJointsValid_ = true; return 0; }
  std::vector<float4> imp(n); std::vector<int> lim(n);
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(imp.data(), j_imp.p, n * 16, cudaMemcpyDeviceToHost), "pull jimp");
  CUDA_OR_FAIL(cudaMemcpy(lim.data(), j_limit.p, n * 4, cudaMemcpyDeviceToHost), "pull jlim");
  for (size_t k = 0; k < n; ++k) { HJoint& j = joints_[jointAt_[k]]; j.imp[0] = imp[k].x; j.imp[1] = imp[k].y; j.imp[2] = imp[k].z; j.imp[3] = imp[k].w; j.limit = lim[k]; }
  hostJointsValid_ = true;
  return 0;
}

// greedy colouring of the joint graph on the host (the joint set only changes through the API)
// parameter packing of the device joint record (which = 0: j_p0, 1: j_p1); the kernels' side is dbx_solver.cuh / dbx_joints2.cuh
float4 World::jointParams(const dbx_joint_def& d, int which) {
  auto f = [](float a, float b, float c, float e) { return make_float4(a, b, c, e); };
  if (which == 2 && d.type != DBX_JOINT_GEAR) return f(0, 0, 0, 0);
  switch (d.type) {
    case DBX_JOINT_REVOLUTE:  return which == 0 ? f(d.referenceAngle, d.lowerAngle, d.upperAngle, d.maxMotorTorque) : f(d.motorSpeed, 0, 0, 0);
    case DBX_JOINT_DISTANCE:  return which == 0 ? f(d.length, d.frequencyHz, d.dampingRatio, 0) : f(0, 0, 0, 0);
    case DBX_JOINT_PRISMATIC: return which == 0 ? f(d.localAxisA.x, d.localAxisA.y, d.referenceAngle, d.maxMotorForce) : f(d.motorSpeed, d.lowerTranslation, d.upperTranslation, 0);
    case DBX_JOINT_WELD:      return which == 0 ? f(d.referenceAngle, d.frequencyHz, d.dampingRatio, 0) : f(0, 0, 0, 0);
    case DBX_JOINT_WHEEL:     return which == 0 ? f(d.localAxisA.x, d.localAxisA.y, d.maxMotorTorque, d.motorSpeed) : f(d.frequencyHz, d.dampingRatio, 0, 0);
    case DBX_JOINT_ROPE:      return which == 0 ? f(d.maxLength, 0, 0, 0) : f(0, 0, 0, 0);
    case DBX_JOINT_FRICTION:  return which == 0 ? f(d.maxForce, d.maxTorque, 0, 0) : f(0, 0, 0, 0);
    case DBX_JOINT_MOTOR:     return which == 0 ? f(d.linearOffset.x, d.linearOffset.y, d.angularOffset, d.correctionFactor) : f(d.maxForce, d.maxTorque, 0, 0);
    case DBX_JOINT_MOUSE:     return which == 0 ? f(d.target.x, d.target.y, d.maxForce, d.frequencyHz) : f(d.dampingRatio, 0, 0, 0);
    case DBX_JOINT_GEAR:      return which == 0 ? f(d.groundAnchorA.x, d.groundAnchorA.y, d.groundAnchorB.x, d.groundAnchorB.y)
                                   : which == 1 ? f(d.localAxisA.x, d.localAxisA.y, d.linearOffset.x, d.linearOffset.y)
                                                : f(d.referenceAngle, d.angularOffset, d.ratio, d.lengthA);
    case DBX_JOINT_PULLEY:    return which == 0 ? f(d.groundAnchorA.x, d.groundAnchorA.y, d.groundAnchorB.x, d.groundAnchorB.y) : f(d.lengthA, d.lengthB, d.ratio, d.lengthA + d.ratio * d.lengthB);
    default:                  return f(0, 0, 0, 0);
  }
}

// b2MouseJoint.SetTarget (b2mousejoint.d:112-120): wakes bodyB, moves the target
int World::setJointTarget(int jid, float x, float y) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (jid < 0 || jid >= (int)joints_.size() || !joints_[jid].alive || joints_[jid].def.type != DBX_JOINT_MOUSE) return DBX_E_INVALID;
  int rc = pullJoints(); if (rc < 0) return rc;
  joints_[jid].def.target = dbx_vec2{x, y};
  fullPushJoints_ = true;
  rc = push(); if (rc < 0) return rc;
  CUDA_OR_FAIL(launch_api_wake(dw_, L_, joints_[jid].def.bodyB, -1), "api_wake");
  hostBodiesValid_ = false, ++bodyEpoch_;
  return 0;
}

// the joint classes' setters (see include/dbox_b200.h "joint parameters at run time")
int World::setJointParams(int jid, const dbx_joint_def& d, uint32_t mask) {
  if (replicated_) { set_last_error("world is replicated: use dbx_world_set_motor_speeds"); return DBX_E_UNSUPPORTED; }
  if (jid < 0 || jid >= (int)joints_.size() || !joints_[jid].alive || joints_[jid].def.type != d.type) return DBX_E_INVALID;
  int rc = push(); if (rc < 0) return rc;              // the joint is on the device, in its colour slot
  rc = pullJoints(); if (rc < 0) return rc;
  HJoint& j = joints_[jid];
  dbx_joint_def& o = j.def;
  const int t = o.type;
  const bool motorised = t == DBX_JOINT_REVOLUTE || t == DBX_JOINT_PRISMATIC || t == DBX_JOINT_WHEEL;
  bool wakeBoth = false, zeroLimitImpulse = false;
  if ((mask & DBX_JP_MOTOR_SPEED) && motorised) { o.motorSpeed = d.motorSpeed; wakeBoth = true; }
  if ((mask & DBX_JP_MAX_MOTOR) && motorised) { if (t == DBX_JOINT_PRISMATIC) o.maxMotorForce = d.maxMotorForce; else o.maxMotorTorque = d.maxMotorTorque; wakeBoth = true; }
  if ((mask & DBX_JP_ENABLE_MOTOR) && motorised) { o.enableMotor = d.enableMotor ? 1 : 0; wakeBoth = true; }
  if ((mask & DBX_JP_ENABLE_LIMIT) && (t == DBX_JOINT_REVOLUTE || t == DBX_JOINT_PRISMATIC) && (d.enableLimit != 0) != (o.enableLimit != 0)) {
    o.enableLimit = d.enableLimit ? 1 : 0; wakeBoth = true; zeroLimitImpulse = true;
  }
  if (mask & DBX_JP_LIMITS) {
    if (t == DBX_JOINT_REVOLUTE && (d.lowerAngle != o.lowerAngle || d.upperAngle != o.upperAngle)) {
      if (!(d.lowerAngle <= d.upperAngle)) return DBX_E_INVALID;
      o.lowerAngle = d.lowerAngle; o.upperAngle = d.upperAngle; wakeBoth = true; zeroLimitImpulse = true;
    } else if (t == DBX_JOINT_PRISMATIC && (d.lowerTranslation != o.lowerTranslation || d.upperTranslation != o.upperTranslation)) {
      if (!(d.lowerTranslation <= d.upperTranslation)) return DBX_E_INVALID;
      o.lowerTranslation = d.lowerTranslation; o.upperTranslation = d.upperTranslation; wakeBoth = true; zeroLimitImpulse = true;
    }
  }
  if ((mask & DBX_JP_SPRING) && (t == DBX_JOINT_DISTANCE || t == DBX_JOINT_WELD || t == DBX_JOINT_WHEEL || t == DBX_JOINT_MOUSE)) { o.frequencyHz = d.frequencyHz; o.dampingRatio = d.dampingRatio; }
  if (mask & DBX_JP_LENGTH) { if (t == DBX_JOINT_DISTANCE) o.length = d.length; else if (t == DBX_JOINT_ROPE) o.maxLength = d.maxLength; }
  if ((mask & DBX_JP_MAX_FORCE) && (t == DBX_JOINT_FRICTION || t == DBX_JOINT_MOTOR || t == DBX_JOINT_MOUSE)) { o.maxForce = d.maxForce; if (t != DBX_JOINT_MOUSE) o.maxTorque = d.maxTorque; }
  if ((mask & DBX_JP_OFFSETS) && t == DBX_JOINT_MOTOR && (d.linearOffset.x != o.linearOffset.x || d.linearOffset.y != o.linearOffset.y || d.angularOffset != o.angularOffset)) {
    o.linearOffset = d.linearOffset; o.angularOffset = d.angularOffset; wakeBoth = true;
  }
  if ((mask & DBX_JP_CORRECTION) && t == DBX_JOINT_MOTOR) o.correctionFactor = d.correctionFactor;
  if (zeroLimitImpulse) j.imp[2] = 0.0f;               // m_impulse.z = 0 (b2revolutejoint.d:236,252; b2prismaticjoint.d likewise)
  // patch the one record where it lives
  const int slot = jointPos_[jid];
  const int4 ids = make_int4(o.type, o.bodyA, o.bodyB, (o.collideConnected ? 1 : 0) | (o.enableLimit ? 2 : 0) | (o.enableMotor ? 4 : 0) | 8);
  const float4 p0 = jointParams(o, 0), p1 = jointParams(o, 1), imp = make_float4(j.imp[0], j.imp[1], j.imp[2], j.imp[3]);
  CUDA_OR_FAIL(cudaMemcpyAsync(j_ids.p + slot, &ids, 16, cudaMemcpyHostToDevice, stream_), "joint patch");
  CUDA_OR_FAIL(cudaMemcpyAsync(j_p0.p + slot, &p0, 16, cudaMemcpyHostToDevice, stream_), "joint patch");
  CUDA_OR_FAIL(cudaMemcpyAsync(j_p1.p + slot, &p1, 16, cudaMemcpyHostToDevice, stream_), "joint patch");
  if (zeroLimitImpulse) CUDA_OR_FAIL(cudaMemcpyAsync(j_imp.p + slot, &imp, 16, cudaMemcpyHostToDevice, stream_), "joint patch");
  if (wakeBoth) { CUDA_OR_FAIL(launch_api_wake(dw_, L_, o.bodyA, o.bodyB), "api_wake"); hostBodiesValid_ = false, ++bodyEpoch_; }
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");      // the host temporaries go away
  return 0;
}
int World::setMotorSpeeds(const int32_t* joints, const float* speeds, int n) {
  if (n < 0 || (n > 0 && (!joints || !speeds))) return DBX_E_INVALID;
  if (n == 0) return 0;
  int rc = push(); if (rc < 0) return rc;
  // device layout: colour-major; inside a colour replica-major (World::replicate): colour c of the template occupies slots
  // [first[c], first[c] + count[c]) and, replicated, [first[c] * R + r * count[c], ...) for replica r
  const int nId = (int)joints_.size();
  std::vector<int> first(kMaxJointColours + 1, 0), count(kMaxJointColours + 1, 0);
  for (size_t k = 0; k < jointAt_.size(); ++k) { const int c = std::min(std::max(joints_[jointAt_[k]].colour, 0), kMaxJointColours); if (count[c]++ == 0) first[c] = (int)k; }
  std::vector<int> slots((size_t)n);
  for (int k = 0; k < n; ++k) {
    const int g = joints[k];
    if (g < 0 || nId == 0) return DBX_E_INVALID;
    const int r = replicated_ ? g / nId : 0, local = replicated_ ? g % nId : g;
    if (r >= nWorlds_ || local >= nId || !joints_[local].alive) return DBX_E_INVALID;
    const int t = joints_[local].def.type;
    if (t != DBX_JOINT_REVOLUTE && t != DBX_JOINT_PRISMATIC && t != DBX_JOINT_WHEEL) return DBX_E_INVALID;
    const int c = std::min(std::max(joints_[local].colour, 0), kMaxJointColours), sl = jointPos_[local];
    slots[k] = replicated_ ? first[c] * nWorlds_ + r * count[c] + (sl - first[c]) : sl;
    if (r == 0) joints_[local].def.motorSpeed = speeds[k];     // the host definition follows (replica 0 = the template)
  }
  CUDA_OR_FAIL(ioIds_.reserve((size_t)n, false, stream_), "io ids"); CUDA_OR_FAIL(qIn_.reserve(((size_t)n + 3) / 4, false, stream_), "io");
  CUDA_OR_FAIL(cudaMemcpyAsync(ioIds_.p, slots.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream_), "slots h2d");
  CUDA_OR_FAIL(cudaMemcpyAsync(qIn_.p, speeds, (size_t)n * 4, cudaMemcpyHostToDevice, stream_), "speeds h2d");
  CUDA_OR_FAIL(launch_set_motor_speeds(dw_, L_, ioIds_.p, (const float*)qIn_.p, n), "motor_speeds");
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  hostBodiesValid_ = false, ++bodyEpoch_;
  return n;
}

int World::recolourJoints() {
  const int nJ = (int)joints_.size();
  std::vector<unsigned long long> mask(bodies_.size(), 0ull);
  std::vector<int> colour(nJ, -1);
  std::vector<std::vector<int>> byColour(kMaxJointColours);
  std::vector<unsigned long long> jp;
  for (int j = 0; j < nJ; ++j) {
    HJoint& hj = joints_[j];
    if (!hj.alive) continue;
    const int a = hj.def.bodyA, b = hj.def.bodyB;
    const bool dynA = bodies_[a].st.type == DBX_DYNAMIC_BODY, dynB = bodies_[b].st.type == DBX_DYNAMIC_BODY;
    unsigned long long used = (dynA ? mask[a] : 0ull) | (dynB ? mask[b] : 0ull);
    // a gear joint also writes the far bodies of joint1 / joint2
    const bool dynC = hj.bodyC >= 0 && bodies_[hj.bodyC].st.type == DBX_DYNAMIC_BODY, dynD = hj.bodyD >= 0 && bodies_[hj.bodyD].st.type == DBX_DYNAMIC_BODY;
    if (dynC) used |= mask[hj.bodyC];
    if (dynD) used |= mask[hj.bodyD];
    if (!~used) { set_last_error("a body has more than 64 joints"); return DBX_E_CAPACITY; }
    int c = __builtin_ffsll((long long)~used) - 1;
    if (dynA) mask[a] |= 1ull << c;
    if (dynB) mask[b] |= 1ull << c;
    if (dynC) mask[hj.bodyC] |= 1ull << c;
    if (dynD) mask[hj.bodyD] |= 1ull << c;
    colour[j] = c; hj.colour = c;
    byColour[c].push_back(j);
    if (!hj.def.collideConnected) {
      unsigned long long lo = (unsigned)std::min(a, b), hi = (unsigned)std::max(a, b);
      jp.push_back((lo << 32) | hi);
    }
  }
  jointAt_.clear(); jointPos_.assign(nJ, -1);
  int off[kMaxJointColours + 1];
  for (int c = 0; c < kMaxJointColours; ++c) { off[c] = (int)jointAt_.size(); jointAt_.insert(jointAt_.end(), byColour[c].begin(), byColour[c].end()); }
  off[kMaxJointColours] = (int)jointAt_.size();
  for (size_t k = 0; k < jointAt_.size(); ++k) jointPos_[jointAt_[k]] = (int)k;
  size_t maxPerColour = 0;
  nJointColours_ = 0;
  for (int c = 0; c < kMaxJointColours; ++c) { maxPerColour = std::max(maxPerColour, byColour[c].size()); if (!byColour[c].empty()) nJointColours_ = c + 1; }
  jointBlocks_ = (int)std::min<size_t>((maxPerColour + L_.coopThreads - 1) / L_.coopThreads, (size_t)L_.coopBlocks / 4);
  // per body: every colour up to its highest joint colour is closed to its contacts (k_mark_solve / k_colour), so that a
  // body's joints always come before its contacts when joint colour c and contact colour c share a solver phase
  jmaskHost_.assign(bodies_.size(), 0ull);
  for (int j = 0; j < nJ; ++j) {
    const HJoint& hj = joints_[j];
    if (!hj.alive) continue;
    const unsigned long long m = hj.colour >= 63 ? ~0ull : ((1ull << (hj.colour + 1)) - 1ull);
    const int bs[4] = {hj.def.bodyA, hj.def.bodyB, hj.bodyC, hj.bodyD};
    for (int b : bs) if (b >= 0 && bodies_[b].st.type == DBX_DYNAMIC_BODY) jmaskHost_[b] |= m;
  }
  { int rm = uploadJointMasks(); if (rm < 0) return rm; }
  std::sort(jp.begin(), jp.end());
  jp.erase(std::unique(jp.begin(), jp.end()), jp.end());
  nJointPairs_ = (int)jp.size();
  CUDA_OR_FAIL(jp_keys.reserve(std::max<size_t>(1, jp.size()), false, stream_), "jp_keys");
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  if (!jp.empty()) CUDA_OR_FAIL(cudaMemcpy(jp_keys.p, jp.data(), jp.size() * 8, cudaMemcpyHostToDevice), "jp up");
  CUDA_OR_FAIL(cudaMemcpy((char*)hdr_.p + offsetof(Header, jointColourOff), off, sizeof(off), cudaMemcpyHostToDevice), "joff up");
  return uploadJointBits(jp);
}

int World::uploadJointMasks() {
  const size_t nb = bodies_.size() * (size_t)nWorlds_;
  std::vector<unsigned long long> all(std::max<size_t>(nb, 1), 0ull);
  for (size_t r = 0; r < (size_t)nWorlds_; ++r) for (size_t b = 0; b < jmaskHost_.size() && b < bodies_.size(); ++b) all[r * bodies_.size() + b] = jmaskHost_[b];
  jmaskBodies_ = nb;
  CUDA_OR_FAIL(b_jmask.reserve(all.size(), false, stream_), "b_jmask");
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(b_jmask.p, all.data(), all.size() * 8, cudaMemcpyHostToDevice), "jmask up");
  dw_.b_jmask = b_jmask.p;
  return 0;
}

// one bit per body that appears in the sorted joint-pair list: b2Body.ShouldCollide only searches the list for those
int World::uploadJointBits(const std::vector<unsigned long long>& keys) {
  const size_t nb = bodies_.size() * (size_t)nWorlds_;
  if (&keys != &jpHost_) jpHost_ = keys;
  jpBitsBodies_ = nb;
  std::vector<uint32_t> bits((nb + 31) / 32 + 1, 0u);
  for (unsigned long long k : keys) {
    const size_t a = (size_t)(k >> 32), b = (size_t)(k & 0xFFFFFFFFull);
    if (a < nb) bits[a >> 5] |= 1u << (a & 31);
    if (b < nb) bits[b >> 5] |= 1u << (b & 31);
  }
  CUDA_OR_FAIL(jp_bits.reserve(bits.size(), false, stream_), "jp_bits");
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(jp_bits.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice), "jp bits up");
  dw_.jp_bits = jp_bits.p;
  return 0;
}

template <class T, class F> static cudaError_t upload_range(DevBuf<T>& buf, size_t from, size_t to, F&& get) {
  if (to <= from) return cudaSuccess;
  std::vector<T> tmp(to - from);
  for (size_t i = from; i < to; ++i) tmp[i - from] = get(i);
  return cudaMemcpy(buf.p + from, tmp.data(), (to - from) * sizeof(T), cudaMemcpyHostToDevice);
}

// (re)size every device pool for the current object counts (times the replica count)
int World::reserveDevice(bool& rehash) {
  // ---- capacities
  const size_t W_ = (size_t)nWorlds_;
  const size_t nB = bodies_.size() * W_, nF = fixtures_.size() * W_, nP = proxies_.size() * W_, nS = shapes_.size(), nJ = joints_.size() * W_;
  const size_t capB = std::max<size_t>(std::max<size_t>(nB, 1), (size_t)caps_.maxBodies);
  const size_t capP = std::max<size_t>(std::max<size_t>(nP, 1), (size_t)caps_.maxProxies);
  DevBuf<float4>* bf4[] = {&b_xf, &b_xf0, &b_pos, &b_pos0, &b_vel, &b_force, &b_mass, &b_lc};
  for (auto* b : bf4) CUDA_OR_FAIL(b->reserve(capB, true, stream_), "body f4");
  CUDA_OR_FAIL(b_gs.reserve(capB, true, stream_), "b_gs");
  CUDA_OR_FAIL(b_flags.reserve(capB, true, stream_), "b_flags");
  DevBuf<int>* bi[] = {&b_wake, &b_root, &b_islAwake, &b_islMinSleep, &b_ovf, &b_world};
  for (auto* b : bi) CUDA_OR_FAIL(b->reserve(capB, true, stream_), "body int");
  CUDA_OR_FAIL(b_mask.reserve(capB, true, stream_), "b_mask");
  CUDA_OR_FAIL(b_claim.reserve(capB, true, stream_), "b_claim");
  CUDA_OR_FAIL(b_acc.reserve(3 * capB, false, stream_), "b_acc");
  CUDA_OR_FAIL(b_toiMin.reserve(capB, true, stream_), "b_toiMin"); CUDA_OR_FAIL(b_toiOther.reserve(capB, true, stream_), "b_toiOther");
  CUDA_OR_FAIL(b_toiEvt.reserve(capB, true, stream_), "b_toiEvt"); CUDA_OR_FAIL(b_toiFlags.reserve(capB, true, stream_), "b_toiFlags");
  // TOI events handled per pass of k_toi: at least one per resident warp, more for batched worlds (events of different
  // worlds are always independent); the surplus of a pass simply waits for the next one
  const size_t nEv = std::min<size_t>(std::max<size_t>((size_t)L_.coopBlocks * (L_.coopThreads / 32), capB / 16), (size_t)1 << 18);
  CUDA_OR_FAIL(e_contact.reserve(nEv, false, stream_), "e_contact"); CUDA_OR_FAIL(e_ncand.reserve(2 * nEv, false, stream_), "e_ncand");
  CUDA_OR_FAIL(e_cand.reserve(2 * nEv * kToiCand, false, stream_), "e_cand");
  CUDA_OR_FAIL(b_posNotOk.reserve(b_root.cap * (size_t)kMaxPosIters, false, stream_), "b_posNotOk");
  CUDA_OR_FAIL(f_body.reserve(std::max<size_t>(nF, 1), true, stream_), "f_body");
  CUDA_OR_FAIL(f_group.reserve(std::max<size_t>(nF, 1), true, stream_), "f_group");
  CUDA_OR_FAIL(f_mat.reserve(std::max<size_t>(nF, 1), true, stream_), "f_mat");
  CUDA_OR_FAIL(f_filter.reserve(std::max<size_t>(nF, 1), true, stream_), "f_filter");
  CUDA_OR_FAIL(d_shapes.reserve(std::max<size_t>(nS, 1), true, stream_), "shapes");
  CUDA_OR_FAIL(p_ids.reserve(capP, true, stream_), "p_ids");
  CUDA_OR_FAIL(p_key.reserve(capP, true, stream_), "p_key");
  CUDA_OR_FAIL(p_aabb.reserve(capP, true, stream_), "p_aabb");
  CUDA_OR_FAIL(p_fat.reserve(capP, true, stream_), "p_fat");
  CUDA_OR_FAIL(p_flags.reserve(capP, true, stream_), "p_flags");
  const size_t pc = p_ids.cap;
  CUDA_OR_FAIL(moveList.reserve(pc, true, stream_), "moveList");   // keep: pending moves may be waiting (replicate)
  CUDA_OR_FAIL(bv_key.reserve(pc, false, stream_), "bv_key"); CUDA_OR_FAIL(bv_keyAlt.reserve(pc, false, stream_), "bv_keyAlt");
  CUDA_OR_FAIL(bv_leaf.reserve(pc, false, stream_), "bv_leaf"); CUDA_OR_FAIL(bv_leafAlt.reserve(pc, false, stream_), "bv_leafAlt");
  CUDA_OR_FAIL(bv_box.reserve(2 * pc, false, stream_), "bv_box"); CUDA_OR_FAIL(bv_child.reserve(pc, false, stream_), "bv_child"); CUDA_OR_FAIL(bv_wr.reserve(pc, false, stream_), "bv_wr");
  CUDA_OR_FAIL(bv_parent.reserve(2 * pc, false, stream_), "bv_parent"); CUDA_OR_FAIL(bv_visit.reserve(pc, false, stream_), "bv_visit");
  CUDA_OR_FAIL(bv_pos.reserve(pc, false, stream_), "bv_pos");
  {
    size_t need = cub_temp_bytes((int)pc);
    CUDA_OR_FAIL(cubTemp.reserve(need, false, stream_), "cubTemp");
  }
  const size_t capC = std::max<size_t>(std::max<size_t>(std::max<size_t>(1024, 8 * pc), (size_t)caps_.maxContacts), contactFloor_);
  const size_t oldCCap = c_key.cap;
  CUDA_OR_FAIL(c_key.reserve(capC, true, stream_), "c_key");
  DevBuf<float4>* cf4[] = {&c_m0, &c_m1, &c_imp, &c_mat};
  for (auto* b : cf4) CUDA_OR_FAIL(b->reserve(capC, true, stream_), "contact f4");
  CUDA_OR_FAIL(c_ids.reserve(capC, true, stream_), "c_ids"); CUDA_OR_FAIL(c_fix.reserve(capC, true, stream_), "c_fix");
  CUDA_OR_FAIL(c_flags.reserve(capC, true, stream_), "c_flags"); CUDA_OR_FAIL(c_mk.reserve(capC, true, stream_), "c_mk");
  DevBuf<int>* ci[] = {&c_toiCount, &c_colour, &c_free, &c_work, &c_work2};
  CUDA_OR_FAIL(c_toiList.reserve(capC, false, stream_), "c_toiList");
  for (auto* b : ci) CUDA_OR_FAIL(b->reserve(capC, true, stream_), "contact int");
  const size_t cc = c_key.cap;
  size_t hc = 1024;
  while (hc < 2 * cc) hc <<= 1;            // power of two (the probe sequence masks), load factor <= 0.5
  if (h_key.cap < hc) {
    h_key.release(); h_val.release();
    CUDA_OR_FAIL(h_key.reserve(hc, false, stream_), "h_key"); CUDA_OR_FAIL(h_val.reserve(hc, false, stream_), "h_val");
    rehash = true;
  }
  CUDA_OR_FAIL(pairs.reserve(std::max<size_t>(std::max<size_t>(4096, cc), (size_t)caps_.maxPairs), false, stream_), "pairs");
  DevBuf<float4>* sf4[] = {&s_v0, &s_v1, &s_r0, &s_r1, &s_q0, &s_q1, &s_imp, &s_nm, &s_k, &s_p0, &s_p1, &s_p2};
  const size_t sc = std::max<size_t>(cc, nEv * (size_t)kMaxTOIContacts);   // TOI mini-islands borrow solver slots [event * 32, +32)
  for (auto* b : sf4) CUDA_OR_FAIL(b->reserve(sc, false, stream_), "solver f4");
  CUDA_OR_FAIL(s_p3.reserve(sc, false, stream_), "s_p3"); CUDA_OR_FAIL(s_body.reserve(sc, false, stream_), "s_body");
  CUDA_OR_FAIL(s_contact.reserve(sc, false, stream_), "s_contact"); CUDA_OR_FAIL(s_pc.reserve(sc, false, stream_), "s_pc"); CUDA_OR_FAIL(s_root.reserve(sc, false, stream_), "s_root");
  CUDA_OR_FAIL(s_hist.reserve((size_t)kSortBlocks * kMaxColours, false, stream_), "s_hist");
  const size_t capJ = std::max<size_t>(std::max<size_t>(nJ, 1), (size_t)caps_.maxJoints);
  DevBuf<float4>* jf4[] = {&j_anchor, &j_p0, &j_p1, &j_imp, &j_r, &j_lc, &j_m, &j_k0, &j_k1, &j_k2, &j_k3, &j_p2};
  CUDA_OR_FAIL(j_ids2.reserve(capJ, true, stream_), "j_ids2");
  for (auto* b : jf4) CUDA_OR_FAIL(b->reserve(capJ, true, stream_), "joint f4");
  CUDA_OR_FAIL(j_ids.reserve(capJ, true, stream_), "j_ids"); CUDA_OR_FAIL(j_limit.reserve(capJ, true, stream_), "j_limit"); CUDA_OR_FAIL(j_root.reserve(capJ, true, stream_), "j_root");
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  (void)oldCCap;

  return 0;
}

int World::push() {
  cudaSetDevice(device_);
  if (replicated_) return 0;   // nothing on the host can be newer than the device any more
  const size_t nB = bodies_.size(), nF = fixtures_.size(), nP = proxies_.size(), nS = shapes_.size(), nJ = joints_.size();
  const bool anyBody = fullPushBodies_ || nB > bodiesSynced_;
  const bool anyRow = !dirtyBodies_.empty();
  const bool anyFix = fullPushFixtures_ || nF > fixturesSynced_;
  const bool anyProxy = fullPushProxies_ || nP > proxiesSynced_ || !pendingMoves_.empty();
  const size_t proxiesKnown = proxiesSynced_;      // proxies the device had before this push
  const bool anyShape = nS > shapesSynced_;
  const bool anyJoint = fullPushJoints_ || nJ > jointsSynced_ || jointsChanged_;
  if (!anyBody && !anyRow && !anyFix && !anyProxy && !anyShape && !anyJoint && dw_.hdr) return 0;
  if (anyBody) tilesDirty_ = true;       // body set or body types may have changed: the tile solver re-counts and re-sorts
  if (fullPushBodies_ && !hostBodiesValid_) { int rc = pullBodies(); if (rc < 0) return rc; }
  if (fullPushProxies_ && !hostProxiesValid_) { int rc = pullProxies(); if (rc < 0) return rc; }
  if (anyJoint && !hostJointsValid_) { int rc = pullJoints(); if (rc < 0) return rc; }
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");

  bool rehash = false;
  { int rcap = reserveDevice(rehash); if (rcap < 0) return rcap; }

  // ---- bodies
  {
    const size_t from = fullPushBodies_ ? 0 : bodiesSynced_;
    auto S = [&](size_t i) -> const dbx_body_state& { return bodies_[i].st; };
    CUDA_OR_FAIL(upload_range(b_xf, from, nB, [&](size_t i) { return f4(S(i).p.x, S(i).p.y, S(i).qs, S(i).qc); }), "up xf");
    CUDA_OR_FAIL(upload_range(b_xf0, from, nB, [&](size_t i) { return bodies_[i].xf0; }), "up xf0");
    CUDA_OR_FAIL(upload_range(b_pos, from, nB, [&](size_t i) { return f4(S(i).c.x, S(i).c.y, S(i).a, 0.0f); }), "up pos");
    CUDA_OR_FAIL(upload_range(b_pos0, from, nB, [&](size_t i) { return f4(S(i).c0.x, S(i).c0.y, S(i).a0, S(i).alpha0); }), "up pos0");
    CUDA_OR_FAIL(upload_range(b_vel, from, nB, [&](size_t i) { return f4(S(i).v.x, S(i).v.y, S(i).w, 0.0f); }), "up vel");
    CUDA_OR_FAIL(upload_range(b_force, from, nB, [&](size_t i) { return f4(S(i).force.x, S(i).force.y, S(i).torque, 0.0f); }), "up force");
    CUDA_OR_FAIL(upload_range(b_mass, from, nB, [&](size_t i) { return f4(S(i).invMass, S(i).invI, S(i).mass, S(i).I); }), "up mass");
    CUDA_OR_FAIL(upload_range(b_lc, from, nB, [&](size_t i) { return f4(S(i).localCenter.x, S(i).localCenter.y, S(i).linearDamping, S(i).angularDamping); }), "up lc");
    CUDA_OR_FAIL(upload_range(b_gs, from, nB, [&](size_t i) { return make_float2(S(i).gravityScale, S(i).sleepTime); }), "up gs");
    CUDA_OR_FAIL(upload_range(b_flags, from, nB, [&](size_t i) {
      if (!bodies_[i].alive) return (uint32_t)0;
      return (uint32_t)((S(i).flags & 0xFFFF) | ((uint32_t)S(i).type << BF_TYPE_SHIFT) | BF_ALIVE); }), "up flags");
    CUDA_OR_FAIL(upload_range(b_world, from, nB, [&](size_t i) { return bodies_[i].world; }), "up world");
    if (!fullPushBodies_ && !dirtyBodies_.empty()) { refreshView(); int rcr = pushBodyRows(); if (rcr < 0) return rcr; }    // (the pools may just have moved)
    for (int b : dirtyBodies_) bodies_[b].dirty = false;
    dirtyBodies_.clear();
    bodiesSynced_ = nB; fullPushBodies_ = false;
  }
  // ---- fixtures, shapes
  {
    const size_t from = fullPushFixtures_ ? 0 : fixturesSynced_;
    CUDA_OR_FAIL(upload_range(f_body, from, nF, [&](size_t i) { return fixtures_[i].body; }), "up f_body");
    CUDA_OR_FAIL(upload_range(f_mat, from, nF, [&](size_t i) { return make_float2(fixtures_[i].def.friction, fixtures_[i].def.restitution); }), "up f_mat");
    CUDA_OR_FAIL(upload_range(f_filter, from, nF, [&](size_t i) { return (uint32_t)fixtures_[i].def.categoryBits | ((uint32_t)fixtures_[i].def.maskBits << 16); }), "up f_filter");
    CUDA_OR_FAIL(upload_range(f_group, from, nF, [&](size_t i) { return (int)((uint32_t)(uint16_t)fixtures_[i].def.groupIndex | ((fixtures_[i].def.isSensor ? FXF_SENSOR : 0u) << 16)); }), "up f_group");
    fixturesSynced_ = nF; fullPushFixtures_ = false;
    if (nS > shapesSynced_) CUDA_OR_FAIL(cudaMemcpy(d_shapes.p + shapesSynced_, shapes_.data() + shapesSynced_, (nS - shapesSynced_) * sizeof(DShape), cudaMemcpyHostToDevice), "up shapes");
    shapesSynced_ = nS;
  }
  // ---- proxies + move buffer
  {
    const size_t from = fullPushProxies_ ? 0 : proxiesSynced_;
    CUDA_OR_FAIL(upload_range(p_ids, from, nP, [&](size_t i) { const HProxy& p = proxies_[i]; return make_int4(p.fixture, p.child, p.body, p.shape); }), "up p_ids");
    CUDA_OR_FAIL(upload_range(p_key, from, nP, [&](size_t i) { return proxies_[i].key; }), "up p_key");
    CUDA_OR_FAIL(upload_range(p_aabb, from, nP, [&](size_t i) { return proxies_[i].aabb; }), "up p_aabb");
    CUDA_OR_FAIL(upload_range(p_fat, from, nP, [&](size_t i) { return proxies_[i].fat; }), "up p_fat");
    CUDA_OR_FAIL(upload_range(p_flags, from, nP, [&](size_t i) { return proxies_[i].alive ? (proxies_[i].flags | PF_ALIVE) : 0u; }), "up p_flags");
    if (fullPushProxies_ || nP > proxiesSynced_) treeValid_ = false;   // the proxy set (or its boxes) changed under the tree
    proxiesSynced_ = nP; fullPushProxies_ = false;
    if (!pendingMoves_.empty()) {
      // b2BroadPhase.BufferMove (b2broadphase.d:244-257).  The device move list is empty right after a step; several
      // pushes may happen before the next one (every API call that edits contacts pushes), so later moves are appended
      // behind the ones already uploaded, each proxy at most once.
      CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
      int onDevice = 0;      // includes moves the device buffered itself (dbx_world_set_body_states)
      CUDA_OR_FAIL(cudaMemcpy(&onDevice, (char*)hdr_.p + offsetof(Header, nMoved), 4, cudaMemcpyDeviceToHost), "read nMoved");
      std::vector<int> cand, mv;
      for (int slot : pendingMoves_) if (proxies_[slot].alive && !movesUploaded_.count(slot)) cand.push_back(slot);
      std::sort(cand.begin(), cand.end());
      cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
      for (int slot : cand) {      // already in the device list (flagged there)?  only possible for proxies the device knows
        uint32_t fl = 0;
        if ((size_t)slot < proxiesKnown && onDevice > (int)movesOnDevice_.size()) cudaMemcpy(&fl, p_flags.p + slot, 4, cudaMemcpyDeviceToHost);
        if (!(fl & PF_MOVED)) mv.push_back(slot);
      }
      for (int slot : mv) { uint32_t fl = proxies_[slot].flags | PF_ALIVE | PF_MOVED; proxies_[slot].flags = fl & ~PF_MOVED; cudaMemcpy(p_flags.p + slot, &fl, 4, cudaMemcpyHostToDevice); movesUploaded_.insert(slot); }
      // a full proxy push rewrites p_flags from the host copy, which does not carry PF_MOVED: restore it for the earlier moves
      for (int slot : movesOnDevice_) if (proxies_[slot].alive) { uint32_t fl = proxies_[slot].flags | PF_ALIVE | PF_MOVED; cudaMemcpy(p_flags.p + slot, &fl, 4, cudaMemcpyHostToDevice); }
      if (!mv.empty()) CUDA_OR_FAIL(cudaMemcpy(moveList.p + onDevice, mv.data(), mv.size() * 4, cudaMemcpyHostToDevice), "up moves");
      movesOnDevice_.insert(movesOnDevice_.end(), mv.begin(), mv.end());
      int nm = onDevice + (int)mv.size();
      CUDA_OR_FAIL(cudaMemcpy((char*)hdr_.p + offsetof(Header, nMoved), &nm, 4, cudaMemcpyHostToDevice), "up nMoved");
      pendingMoves_.clear();
    }
  }
  // ---- joints: the device arrays hold the alive joints in colour order (slot k = joint jointAt_[k]), so every colour is a
  // contiguous, coalesced range for the solver; any change of the joint set re-colours and re-uploads all of them
  if (anyJoint || !dw_.hdr) {
    int rc = recolourJoints(); if (rc < 0) return rc;
    const size_t nD = jointAt_.size();
    auto J = [&](size_t k) -> const HJoint& { return joints_[jointAt_[k]]; };
    CUDA_OR_FAIL(upload_range(j_ids, 0, nD, [&](size_t k) { const HJoint& j = J(k);
      return make_int4(j.def.type, j.def.bodyA, j.def.bodyB, (j.def.collideConnected ? 1 : 0) | (j.def.enableLimit ? 2 : 0) | (j.def.enableMotor ? 4 : 0) | 8); }), "up j_ids");
    CUDA_OR_FAIL(upload_range(j_anchor, 0, nD, [&](size_t k) { const dbx_joint_def& d = J(k).def; return f4(d.localAnchorA.x, d.localAnchorA.y, d.localAnchorB.x, d.localAnchorB.y); }), "up j_anchor");
    CUDA_OR_FAIL(upload_range(j_p0, 0, nD, [&](size_t k) { return jointParams(J(k).def, 0); }), "up j_p0");
    CUDA_OR_FAIL(upload_range(j_p1, 0, nD, [&](size_t k) { return jointParams(J(k).def, 1); }), "up j_p1");
    CUDA_OR_FAIL(upload_range(j_p2, 0, nD, [&](size_t k) { return jointParams(J(k).def, 2); }), "up j_p2");
    CUDA_OR_FAIL(upload_range(j_ids2, 0, nD, [&](size_t k) { const HJoint& j = J(k); return make_int4(j.bodyC, j.bodyD, j.typeA, j.typeB); }), "up j_ids2");
    CUDA_OR_FAIL(upload_range(j_imp, 0, nD, [&](size_t k) { const HJoint& j = J(k); return f4(j.imp[0], j.imp[1], j.imp[2], j.imp[3]); }), "up j_imp");
    CUDA_OR_FAIL(upload_range(j_limit, 0, nD, [&](size_t k) { return J(k).limit; }), "up j_limit");
    jointsSynced_ = nJ; fullPushJoints_ = false; jointsChanged_ = false;
  } else {
    if (nJointPairs_ > 0 && nB > jpBitsBodies_) { int rc = uploadJointBits(jpHost_); if (rc < 0) return rc; }   // bodies were added: the bit array must cover them
    if (nB > jmaskBodies_) { jmaskHost_.resize(nB, 0ull); int rc = uploadJointMasks(); if (rc < 0) return rc; }
  }
  refreshView();
  if (rehash) CUDA_OR_FAIL(stage_rebuild_hash(dw_, L_), "rehash");
  return 0;
}

void World::refreshView() {
  DevWorld& w = dw_;
  w.hdr = hdr_.p;
  w.nBodies = (int)bodies_.size() * nWorlds_;
  w.b_xf = b_xf.p; w.b_xf0 = b_xf0.p; w.b_pos = b_pos.p; w.b_pos0 = b_pos0.p; w.b_vel = b_vel.p; w.b_force = b_force.p; w.b_mass = b_mass.p; w.b_lc = b_lc.p;
  w.b_gs = b_gs.p; w.b_flags = b_flags.p; w.b_wake = b_wake.p; w.b_root = b_root.p; w.b_islAwake = b_islAwake.p; w.b_islMinSleep = b_islMinSleep.p;
  w.b_acc = b_acc.p; w.b_toiMin = b_toiMin.p; w.b_toiOther = b_toiOther.p; w.b_toiEvt = b_toiEvt.p; w.b_toiFlags = b_toiFlags.p;
  w.e_contact = e_contact.p; w.e_ncand = e_ncand.p; w.e_cand = e_cand.p; w.eventCap = (int)e_contact.cap; w.bv_pos = bv_pos.p;
  w.b_posNotOk = b_posNotOk.p; w.b_mask = b_mask.p; w.b_claim = b_claim.p; w.b_ovf = b_ovf.p; w.b_world = b_world.p;
  w.nFixtures = (int)fixtures_.size() * nWorlds_; w.f_body = f_body.p; w.f_mat = f_mat.p; w.f_filter = f_filter.p; w.f_group = f_group.p;
  w.nShapes = (int)shapes_.size(); w.shapes = d_shapes.p;
  w.nProxies = (int)proxies_.size() * nWorlds_; w.p_ids = p_ids.p; w.p_key = p_key.p; w.p_aabb = p_aabb.p; w.p_fat = p_fat.p; w.p_flags = p_flags.p;
  w.moveList = moveList.p; w.moveCap = (int)moveList.cap;
  w.bv_key = bv_key.p; w.bv_keyAlt = bv_keyAlt.p; w.bv_leaf = bv_leaf.p; w.bv_leafAlt = bv_leafAlt.p; w.bv_box = bv_box.p; w.bv_child = bv_child.p; w.bv_wr = bv_wr.p; w.bv_parent = bv_parent.p; w.bv_visit = bv_visit.p;
  w.pairs = pairs.p; w.pairCap = (int)pairs.cap;
  w.nJointPairs = nJointPairs_; w.jp_keys = jp_keys.p; w.jp_bits = jp_bits.p; w.b_jmask = b_jmask.p;
  w.cCap = (int)c_key.cap; w.c_key = c_key.p; w.c_ids = c_ids.p; w.c_fix = c_fix.p; w.c_flags = c_flags.p; w.c_m0 = c_m0.p; w.c_m1 = c_m1.p; w.c_imp = c_imp.p; w.c_mk = c_mk.p;
  w.c_mat = c_mat.p; w.c_toiList = c_toiList.p; w.c_toiCount = c_toiCount.p; w.c_colour = c_colour.p; w.c_free = c_free.p; w.c_work = c_work.p; w.c_work2 = c_work2.p;
  w.hCap = (int)h_key.cap; w.h_key = h_key.p; w.h_val = h_val.p;
  w.sCap = (int)s_contact.cap; w.s_contact = s_contact.p; w.s_hist = s_hist.p; w.s_body = s_body.p; w.s_v0 = s_v0.p; w.s_v1 = s_v1.p; w.s_r0 = s_r0.p; w.s_r1 = s_r1.p;
  w.s_q0 = s_q0.p; w.s_q1 = s_q1.p; w.s_imp = s_imp.p; w.s_nm = s_nm.p; w.s_k = s_k.p; w.s_pc = s_pc.p; w.s_p0 = s_p0.p; w.s_p1 = s_p1.p; w.s_p2 = s_p2.p; w.s_p3 = s_p3.p; w.s_root = s_root.p;
  w.nJoints = (int)jointAt_.size() * nWorlds_; w.j_ids = j_ids.p; w.j_anchor = j_anchor.p; w.j_p0 = j_p0.p; w.j_p1 = j_p1.p; w.j_imp = j_imp.p; w.j_limit = j_limit.p; w.j_colour = j_colour.p;
  w.j_order = j_order.p; w.j_root = j_root.p; w.j_r = j_r.p; w.j_lc = j_lc.p; w.j_m = j_m.p; w.j_k0 = j_k0.p; w.j_k1 = j_k1.p; w.j_k2 = j_k2.p; w.j_k3 = j_k3.p; w.j_p2 = j_p2.p; w.j_ids2 = j_ids2.p;
  w.nWorlds = nWorlds_; w.keyStride = keyStride_;
  w.jointBlocks = jointBlocks_; w.nJointColours = nJointColours_;
  { const char* e = getenv("DBX_DEBUG"); w.dbgFlags = e ? atoi(e) : 0; }
  w.phaseTimes = phaseBuf_.p; w.phaseCap = phaseBuf_.p ? (int)phaseBuf_.cap : 0;
  L_.cubTemp = cubTemp.p; L_.cubTempBytes = cubTemp.cap;
}

void World::setStepParams(float dt, int vi, int pi) {
  // b2TimeStep (dynamics/b2world.d:380-396)
  dw_.dt = dt;
  dw_.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
  dw_.dtRatio = inv_dt0 * dt;
  dw_.velIters = vi; dw_.posIters = std::min(pi, kMaxPosIters);
  dw_.warmStarting = (flags_ & DBX_WORLD_WARM_STARTING) ? 1 : 0;
  dw_.allowSleep = (flags_ & DBX_WORLD_ALLOW_SLEEP) ? 1 : 0;
  // k_collide lists TOI candidates only for a step whose SolveTOI will run and consume (and empty) the list (b2world.d:414-419)
  dw_.continuous = ((flags_ & DBX_WORLD_CONTINUOUS) && dt > 0.0f) ? 1 : 0;
  dw_.gx = gx_; dw_.gy = gy_;
  dw_.stepIndex = (stepCount_ + 1) & 0xFFFF;
}

int World::checkDeviceError(bool sync) {
  if (sync) CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "step sync");
  int err = 0;
  CUDA_OR_FAIL(cudaMemcpy(&err, (char*)hdr_.p + offsetof(Header, error), 4, cudaMemcpyDeviceToHost), "read error");
  if (err != 0) {
    static const char* const what[] = {"a device pool", "the contact pool (dbx_caps.maxContacts)", "the candidate pair buffer (dbx_caps.maxPairs)",
        "the move buffer (dbx_caps.maxProxies)", "the constraint colours (more than 1024 constraints on one body)", "the solver rows (dbx_caps.maxContacts)",
        "the pair hash (dbx_caps.maxContacts)", "a tree traversal stack (LBVH deeper than the query reserve)",
        "the TOI candidate list (more than 64 non-dynamic / bullet contacts on one body of a TOI event)"};
    const int site = err <= -5 ? ((-err - 5) / 16) : 0;
    set_last_error(std::string("device pool overflow: ") + what[site >= 0 && site < 9 ? site : 0] + "; see dbx_caps");
    err = (err <= -5) ? DBX_E_CAPACITY : err;
    int zero = 0; cudaMemcpy((char*)hdr_.p + offsetof(Header, error), &zero, 4, cudaMemcpyHostToDevice);
    return err;
  }
  return 0;
}

// b2ContactManager.FindNewContacts.  The LBVH is rebuilt when the proxy set changed or every kRebuildPeriod calls; in
// between it is widened for the moved proxies (lbvh_enlarge), which keeps the pair set exact at a fraction of the cost.
int World::findNewContacts(bool deferClear) {
  movesOnDevice_.clear(); movesUploaded_.clear();      // this call consumes the move buffer
  constexpr int kRebuildPeriod = 8;
  const bool rebuild = !treeValid_ || sinceRebuild_ >= kRebuildPeriod;
  CUDA_OR_FAIL(stage_find_new_contacts(dw_, L_, rebuild, deferClear), "find_new_contacts");
  if (rebuild) { treeValid_ = true; sinceRebuild_ = 0; } else ++sinceRebuild_;
  return 0;
}

int World::growContactsIfNeeded() {
  if (!wmPending_ || cudaEventQuery(wmEv_) != cudaSuccess) return 0;
  wmPending_ = false;
  seenMaxColour_ = std::max(seenMaxColour_, wm_[(int)(offsetof(Header, maxColour) / 4)]);   // (monotonic on the device; the tile solver looks at it)
  const size_t high = (size_t)std::max(wm_[0], 0);          // Header::cHigh
  // a small pool is watched more closely and grown by more: a scene that is just being filled (bodies born on top of each
  // other, tumbler.d:78-97) can double its contacts within a few steps, and the memory at stake is nothing
  const size_t factor = c_key.cap <= 65536 ? 4 : 2;
  if (factor * high <= c_key.cap) return 0;
  contactFloor_ = factor * c_key.cap;
  bool rehash = false;
  int rc = reserveDevice(rehash); if (rc < 0) return rc;
  refreshView();
  if (rehash) CUDA_OR_FAIL(stage_rebuild_hash(dw_, L_), "rehash");
  return 0;
}

// Tile solver set-up (dbx_tiles.cu).  Worth it for a single big world: the dynamic bodies are counted (when the body set may
// have changed), cut into at most one tile per CTA of at least 256 bodies, and re-sorted along x every 16 steps -- any
// assignment is VALID (it only decides which constraints are local, boundary or global), a fresh one keeps the tiles compact.
int World::prepareTiles() {
  constexpr int kTileMinBodies = 2048, kTileMinPerTile = 256, kTileMaxPerTile = 4800, kTileSortPeriod = 16;
  if (replicated_ || overrideLevels_ || (dw_.dbgFlags & 64)) return 0;
  if (bodies_.size() >= ((size_t)1 << 27)) return 0;       // (k_solve_tiles keeps a body's exchange flags above its 27-bit id)
  // A body with more than 64 touching contacts (the Tumbler's container) serialises its surplus on overflow colours: hundreds of
  // one-row phases.  k_solve runs those inside one CTA (its tail); the tile solver would make each a grid-wide phase.  The
  // watermark copy of the device header (every 8th step, never waited for) tells when a world has such a hub.
  if (seenMaxColour_ >= kTileColours) return 0;
  if (tilesDirty_) {
    int n = 0, nk = 0;
    for (const HBody& hb : bodies_) if (hb.alive) { if (hb.st.type == DBX_DYNAMIC_BODY) ++n; else if (hb.st.type == DBX_KINEMATIC_BODY) ++nk; }
    nDynamic_ = n; dw_.tileKinematic = nk > 0 ? 1 : 0; tilesDirty_ = false; tilesValid_ = false;
  }
  if (nDynamic_ < kTileMinBodies) return 0;
  const int P = std::max(1, std::min(std::min(L_.coopBlocks, 176), nDynamic_ / kTileMinPerTile));    // (176: what k_tile_scan's histograms in shared memory allow)
  const int T = (nDynamic_ + P - 1) / P;
  if (T > kTileMaxPerTile) return 0;
  const size_t capB = b_root.cap, capC = c_key.cap, capJ = std::max<size_t>(j_ids.cap, 1);
  const size_t nBins = (size_t)2 * P * kTileColours + kMaxColours + 1;
  DevBuf<int>* perBody[] = {&b_tslot_, &t_body_, &b_tclaim_, &b_xflag_, &tValA_, &tValB_};
  for (auto* b : perBody) CUDA_OR_FAIL(b->reserve(capB, false, stream_), "tile bodies");
  CUDA_OR_FAIL(t_mass_.reserve(capB, false, stream_), "tile masses");
  CUDA_OR_FAIL(tKeyA_.reserve(capB, false, stream_), "tile keys"); CUDA_OR_FAIL(tKeyB_.reserve(capB, false, stream_), "tile keys");
  if (tileBodyCap_ != b_tclaim_.cap) {       // fresh buffers: claims at rest, no exchange flags, tiles to be assigned
    CUDA_OR_FAIL(cudaMemsetAsync(b_tclaim_.p, 0x7F, b_tclaim_.cap * 4, stream_), "claims");
    CUDA_OR_FAIL(cudaMemsetAsync(b_xflag_.p, 0, b_xflag_.cap * 4, stream_), "xflags");
    tileBodyCap_ = b_tclaim_.cap; tilesValid_ = false;
  }
  CUDA_OR_FAIL(c_tkey_.reserve(capC, false, stream_), "tile contact keys"); CUDA_OR_FAIL(c_bref_.reserve(capC, false, stream_), "tile contact refs");
  CUDA_OR_FAIL(j_tkey_.reserve(capJ, false, stream_), "tile joint keys"); CUDA_OR_FAIL(j_bref_.reserve(capJ, false, stream_), "tile joint refs");
  CUDA_OR_FAIL(tj_order_.reserve(capJ, false, stream_), "tile joint order");
  CUDA_OR_FAIL(c_tcol_.reserve(capC, false, stream_), "tile contact colours"); CUDA_OR_FAIL(j_tcol_.reserve(capJ, false, stream_), "tile joint colours");
  CUDA_OR_FAIL(t_flag_.reserve((size_t)2 * L_.coopBlocks, false, stream_), "tile flags");
  DevBuf<int>* bins[] = {&t_off_, &t_cur_, &tj_off_, &tj_cur_};
  for (auto* b : bins) CUDA_OR_FAIL(b->reserve(nBins, false, stream_), "tile bins");
  const size_t need = cub_temp_bytes_u32((int)capB);
  if (need > cubTemp.cap) { CUDA_OR_FAIL(cubTemp.reserve(need, false, stream_), "cubTemp"); L_.cubTemp = cubTemp.p; L_.cubTempBytes = cubTemp.cap; }
  DevWorld& w = dw_;
  if (w.nTiles != P || w.tileBodies != T || w.nTileBodies != nDynamic_) tilesValid_ = false;
  w.nTiles = P; w.tileBodies = T; w.nTileBodies = nDynamic_;
  w.b_tslot = b_tslot_.p; w.t_body = t_body_.p; w.b_tclaim = b_tclaim_.p; w.b_xflag = b_xflag_.p; w.t_mass = t_mass_.p;
  w.c_tkey = c_tkey_.p; w.c_bref = c_bref_.p; w.j_tkey = j_tkey_.p; w.j_bref = j_bref_.p; w.c_tcol = c_tcol_.p; w.j_tcol = j_tcol_.p;
  w.t_off = t_off_.p; w.t_cur = t_cur_.p; w.tj_off = tj_off_.p; w.tj_cur = tj_cur_.p; w.tj_order = tj_order_.p; w.t_flag = t_flag_.p;
  if (!tilesValid_ || sinceTileSort_ >= kTileSortPeriod) {
    CUDA_OR_FAIL(stage_tile_assign(w, L_, tKeyA_.p, tKeyB_.p, tValA_.p, tValB_.p), "tile assign");
    tilesValid_ = true; sinceTileSort_ = 0;
  } else ++sinceTileSort_;
  return P;
}

// one b2World.Step (dynamics/b2world.d:367-434) enqueued on the world's stream; no host synchronisation
// `halves`: 1 = up to and including Collide, 2 = everything after it, 3 = the whole step.  The split exists for
// b2ContactListener.PreSolve (b2contact.d:348-355): dbx_world_step_begin / patch_contacts / step_end.
int World::enqueueStep(float dt, int vi, int pi, bool fineEvents, int halves) {
  int rc = 0;
  if (halves & 1) {
    rc = push(); if (rc < 0) return rc;
    if (bodies_.empty()) return 0;
    rc = growContactsIfNeeded(); if (rc < 0) return rc;
    if (newFixture_) { int rf = findNewContacts(); if (rf < 0) return rf; newFixture_ = false; }   // :372-376
    setStepParams(dt, vi, pi);
  } else if (bodies_.empty()) return 0;
  dw_.colourOverride = overrideLevels_ ? 1 : 0;
  dw_.unifiedColours = (!overrideLevels_ && !jointAt_.empty() && !(dw_.dbgFlags & 32)) ? 1 : 0;
  lastUnified_ = dw_.unifiedColours != 0;
  // TOI: the kernel leaves its per-body scratch clean; a world that has not run it yet, or has grown since, resets first.
  // When the scratch is clean the first TOI evaluation is forked onto a second stream right after the solver.
  const bool continuous = (flags_ & DBX_WORLD_CONTINUOUS) && dt > 0.0f;
  const bool toiScratchDirty = !toiClean_ || toiBodies_ != bodies_.size() * (size_t)nWorlds_;
  // (a world big enough to fill the machine gains nothing from the overlap: the two streams just take SMs from each other)
  const bool subStepping = (flags_ & DBX_WORLD_SUB_STEPPING) != 0;
  const bool toiPre = continuous && !toiScratchDirty && stepComplete_ && !subStepping && !overrideLevels_ && !(dw_.dbgFlags & 8) &&
                      bodies_.size() * (size_t)nWorlds_ <= ((size_t)1 << 21);
  auto mark = [&](int i) { if (fineEvents || i == 0 || i == 1 || i == 3 || i == 5 || i == 7 || i == 8 || i == 9) cudaEventRecord(ev_[i], stream_); };
  if (halves & 1) {
    mark(0);
    CUDA_OR_FAIL(stage_collide(dw_, L_), "collide");
    mark(1);
  }
  if (!(halves & 2)) { hostBodiesValid_ = false, ++bodyEpoch_; return 0; }
  if (dw_.psCap > 0) CUDA_OR_FAIL(cudaMemsetAsync((char*)hdr_.p + offsetof(Header, nPostSolve), 0, 4, stream_), "post-solve reset");
  if (stepComplete_ && dt > 0.0f) {
    CUDA_OR_FAIL(stage_islands_and_integrate(dw_, L_), "islands");
    mark(2);
    // batched replicas without joints: solver slots in (replica, colour) order and one CTA per replica (k_solve_worlds)
    // One CTA per replica pays off when a replica has enough constraints to keep a CTA busy between its block barriers (the
    // Pyramid: 212 bodies, ~400 contacts); replicas of a tiny world (a dozen bodies) are better served by the global solver,
    // whose every phase spans all replicas at once.  Measured on 4,096 replicas of a jointed mechanism repeated u times per world:
    // u = 24 (193 bodies, 192 contacts, 96 joints per world) 1.25 ms per step here against 1.67 ms through the global solver;
    // u = 7 (57 bodies) 0.74 against 0.60 ms; 16,384 replicas of u = 1 (9 bodies) 1.74 against 0.44 ms.
    const bool bigEnough = bodies_.size() >= 128 || (dw_.dbgFlags & 1024);
    const bool worldsPath = replicated_ && bigEnough && (jointAt_.empty() || !(dw_.dbgFlags & 512)) && !overrideLevels_ && !(dw_.dbgFlags & 16);
    bool tilesPath = false;
    dw_.tiled = 0;
    if (worldsPath) {
      // Sort exactly the slots in use, with exactly the key bits the colours need.  Both numbers are device-side facts
      // (cHigh: final since the last FindNewContacts; maxColour: monotonic, so a stale read is still an upper bound for
      // what existed then -- it is read AFTER this step's colouring): one small D2H copy and a stream sync per step,
      // ~20 us against a step of tens of milliseconds.
      CUDA_OR_FAIL(launch_mark_and_colour(dw_, L_), "colour");
      if (!wm_) { CUDA_OR_FAIL(cudaMallocHost((void**)&wm_, 256), "watermark"); CUDA_OR_FAIL(cudaEventCreateWithFlags(&wmEv_, cudaEventDisableTiming), "watermark"); }
      CUDA_OR_FAIL(cudaMemcpyAsync(wm_ + 16, hdr_.p, 64 + 32, cudaMemcpyDeviceToHost, stream_), "header peek");
      CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "header peek");
      const size_t n = std::min(c_key.cap, (size_t)std::max(wm_[16 + 0], 1));                       // Header::cHigh
      const int maxColour = std::max(wm_[16 + (int)(offsetof(Header, maxColour) / 4)], 0);
      int colourBits = 1;
      while ((1 << colourBits) <= maxColour && colourBits < 10) ++colourBits;
      CUDA_OR_FAIL(swKeyA_.reserve(n, false, stream_), "world keys"); CUDA_OR_FAIL(swKeyB_.reserve(n, false, stream_), "world keys");
      CUDA_OR_FAIL(swValA_.reserve(n, false, stream_), "world vals"); CUDA_OR_FAIL(swValB_.reserve(n, false, stream_), "world vals");
      CUDA_OR_FAIL(wStart_.reserve((size_t)nWorlds_, false, stream_), "world ranges"); CUDA_OR_FAIL(wEnd_.reserve((size_t)nWorlds_, false, stream_), "world ranges");
      const size_t need = cub_temp_bytes_u32((int)n);
      if (need > cubTemp.cap) { CUDA_OR_FAIL(cubTemp.reserve(need, false, stream_), "cubTemp"); L_.cubTemp = cubTemp.p; L_.cubTempBytes = cubTemp.cap; }
      dw_.w_start = wStart_.p; dw_.w_end = wEnd_.p;
      CUDA_OR_FAIL(stage_colour_and_sort_worlds(dw_, L_, swKeyA_.p, swKeyB_.p, swValA_.p, swValB_.p, (int)n, colourBits), "colour (worlds)");
    } else {
      dw_.s_contact = s_contact.p;
      const int tp = prepareTiles(); if (tp < 0) return tp;
      tilesPath = tp > 0;
      dw_.tiled = tilesPath ? 1 : 0;
      if (tilesPath) CUDA_OR_FAIL(stage_colour_and_sort_tiles(dw_, L_), "colour (tiles)");
      else CUDA_OR_FAIL(stage_colour_and_sort(dw_, L_), "colour");
    }
    lastTiled_ = tilesPath; lastWorldsPath_ = worldsPath;
    mark(3);
    CUDA_OR_FAIL(stage_prepare(dw_, L_), "prepare");
    mark(4);
    if (worldsPath) CUDA_OR_FAIL(stage_solve_worlds(dw_, L_, (int)bodies_.size()), "solve (worlds)");
    else if (tilesPath) CUDA_OR_FAIL(stage_solve_tiles(dw_, L_), "solve (tiles)");
    else CUDA_OR_FAIL(stage_solve(dw_, L_), "solve");
    dw_.tiled = 0;        // the TOI sub-steps that follow build their own rows with plain body ids
    if (dw_.psCap > 0) CUDA_OR_FAIL(launch_post_solve(dw_, L_), "post_solve");   // island.Report (b2island.d:239); before k_toi reuses the rows
    mark(5);
    if (toiPre) {
      if (!aux_) {
        CUDA_OR_FAIL(cudaStreamCreateWithFlags(&aux_, cudaStreamNonBlocking), "aux stream");
        CUDA_OR_FAIL(cudaEventCreateWithFlags(&evFork_, cudaEventDisableTiming), "fork event");
        CUDA_OR_FAIL(cudaEventCreateWithFlags(&evJoin_, cudaEventDisableTiming), "join event");
      }
      CUDA_OR_FAIL(cudaEventRecord(evFork_, stream_), "fork");
      CUDA_OR_FAIL(cudaStreamWaitEvent(aux_, evFork_, 0), "fork wait");
      CUDA_OR_FAIL(stage_toi_pre(dw_, L_, aux_), "toi_pre");
      CUDA_OR_FAIL(cudaEventRecord(evJoin_, aux_), "join");
    }
    CUDA_OR_FAIL(stage_sync_fixtures(dw_, L_), "sync_fixtures");
    mark(6);
    { int rf = findNewContacts(continuous); if (rf < 0) return rf; }
    mark(7);
  } else {
    for (int i = 2; i <= 7; ++i) cudaEventRecord(ev_[i], stream_);
  }
  if (toiPre) CUDA_OR_FAIL(cudaStreamWaitEvent(stream_, evJoin_, 0), "join wait");
  if (continuous) {   // :414-419
    dw_.toiReset = toiScratchDirty ? 1 : 0; dw_.toiPre = toiPre ? 1 : 0; dw_.toiMode = 0;
    // b2World.SetSubStepping (b2world.d:1441-1446): one TOI event per Step; a Step that follows an unfinished one resumes (:1131)
    dw_.subStep = subStepping ? 1 : 0; dw_.toiResume = stepComplete_ ? 0 : 1;
    if (subStepping) CUDA_OR_FAIL(cudaMemsetAsync((char*)hdr_.p + offsetof(Header, toiGlobalMin), 0xFF, 8, stream_), "toi min reset");
    dw_.toiClearMoves = (stepComplete_ && dt > 0.0f) ? 1 : 0;      // this step's FindNewContacts left its move buffer to us
    dw_.toiClearForces = (flags_ & DBX_WORLD_AUTO_CLEAR_FORCES) ? 1 : 0;
    CUDA_OR_FAIL(stage_toi(dw_, L_), "toi");
    toiClean_ = true; toiBodies_ = bodies_.size() * (size_t)nWorlds_;
  }
  mark(8);
  if (continuous && (subStepping || !stepComplete_)) {
    int inc = 0;
    CUDA_OR_FAIL(cudaMemcpyAsync(&inc, (char*)hdr_.p + offsetof(Header, stepIncomplete), 4, cudaMemcpyDeviceToHost, stream_), "step complete?");
    CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
    stepComplete_ = inc == 0;
  }
  if (dt > 0.0f) inv_dt0 = dw_.inv_dt;
  if ((flags_ & DBX_WORLD_AUTO_CLEAR_FORCES) && !continuous) CUDA_OR_FAIL(launch_clear_forces(dw_, L_), "clear_forces");   // else k_toi did it
  mark(9);
  evValid_ = true; evFine_ = fineEvents;
  ++stepCount_;
  if (overrideLevels_) {   // one-shot: hand the colours back to the colouring pass
    overrideLevels_ = false; dw_.colourOverride = 0;
    CUDA_OR_FAIL(cudaMemsetAsync(c_colour.p, 0xFF, c_colour.cap * 4, stream_), "reset colours");
  }
  if ((stepCount_ & 63) == 0) { int rc2 = compactContacts(); if (rc2 < 0) return rc2; }
  if ((stepCount_ & (c_key.cap <= 65536 ? 1 : 7)) == 0 && !wmPending_) {
    if (!wm_) { CUDA_OR_FAIL(cudaMallocHost((void**)&wm_, 256), "watermark"); CUDA_OR_FAIL(cudaEventCreateWithFlags(&wmEv_, cudaEventDisableTiming), "watermark"); }
    CUDA_OR_FAIL(cudaMemcpyAsync(wm_, hdr_.p, 96, cudaMemcpyDeviceToHost, stream_), "watermark");      // cHigh ... maxColour
    CUDA_OR_FAIL(cudaEventRecord(wmEv_, stream_), "watermark");
    wmPending_ = true;
  }
  hostBodiesValid_ = false, ++bodyEpoch_; hostProxiesValid_ = false; hostJointsValid_ = false;
  return 0;
}

// n steps without host round trips in between; one synchronisation at the end
// b2World.Step cut at the point where the reference calls PreSolve: begin = FindNewContacts-if-needed + Collide; the host
// may then read the contacts, poll the begin/end events and patch contacts; end = Solve, SolveTOI, ClearForces
int World::stepBegin(float dt, int vi, int pi) {
  if (!ok_) return DBX_E_NO_DEVICE;
  cudaSetDevice(device_);
  if (midStep_) { set_last_error("step_begin: the previous step was not ended"); return DBX_E_INVALID; }
  stepDt_ = dt; stepVi_ = vi; stepPi_ = pi;
  int rc = enqueueStep(dt, vi, pi, false, 1); if (rc < 0) return rc;
  midStep_ = true;
  return checkDeviceError(true);
}
int World::stepEnd() {
  if (!ok_) return DBX_E_NO_DEVICE;
  cudaSetDevice(device_);
  if (!midStep_) { set_last_error("step_end without step_begin"); return DBX_E_INVALID; }
  midStep_ = false;
  int rc = enqueueStep(stepDt_, stepVi_, stepPi_, false, 2); if (rc < 0) return rc;
  return checkDeviceError(true);
}
// b2Contact.SetEnabled / SetFriction / SetRestitution / SetTangentSpeed (contacts/b2contact.d:137-205) from PreSolve
int World::patchContacts(const dbx_contact_patch* in, int n) {
  if (n < 0 || (n > 0 && !in)) return DBX_E_INVALID;
  if (n == 0 || !dw_.hdr) return 0;
  std::vector<unsigned long long> keys((size_t)n); std::vector<float4> vals((size_t)n); std::vector<int> masks((size_t)n);
  std::unordered_set<unsigned long long> destroyed;
  for (int k = 0; k < n; ++k) {
    const dbx_contact_patch& c = in[k];
    auto proxyKey = [&](int f, int child) -> int {
      if (f < 0 || f >= (int)fixtures_.size() || !fixtures_[f].alive || child < 0 || child >= (int)fixtures_[f].proxies.size()) return -1;
      return proxies_[fixtures_[f].proxies[child]].key;
    };
    const int ka = proxyKey(c.fixtureA, c.childA), kb = proxyKey(c.fixtureB, c.childB);
    if (ka < 0 || kb < 0) return DBX_E_INVALID;
    keys[k] = ((unsigned long long)(unsigned)std::min(ka, kb) << 32) | (unsigned)std::max(ka, kb);
    masks[k] = c.mask | (c.enabled ? 0x100 : 0);
    if ((c.mask & DBX_PATCH_DESTROY) && !destroyed.insert(keys[k]).second) masks[k] = 0;   // the same pair twice: destroy once
    vals[k] = make_float4(c.friction, c.restitution, c.tangentSpeed, 0.0f);
  }
  CUDA_OR_FAIL(patchKeys_.reserve((size_t)n, false, stream_), "patch"); CUDA_OR_FAIL(qIn_.reserve((size_t)n, false, stream_), "patch"); CUDA_OR_FAIL(qCount_.reserve((size_t)n, false, stream_), "patch");
  CUDA_OR_FAIL(cudaMemcpyAsync(patchKeys_.p, keys.data(), (size_t)n * 8, cudaMemcpyHostToDevice, stream_), "patch h2d");
  CUDA_OR_FAIL(cudaMemcpyAsync(qIn_.p, vals.data(), (size_t)n * 16, cudaMemcpyHostToDevice, stream_), "patch h2d");
  CUDA_OR_FAIL(cudaMemcpyAsync(qCount_.p, masks.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream_), "patch h2d");
  CUDA_OR_FAIL(launch_patch_contacts(dw_, L_, patchKeys_.p, qIn_.p, qCount_.p, n), "patch_contacts");
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");     // the host vectors go away
  return n;
}

int World::step(float dt, int vi, int pi, int n) {
  if (!ok_) return DBX_E_NO_DEVICE;
  cudaSetDevice(device_);
  if (midStep_) { set_last_error("step: a split step is open (step_begin without step_end)"); return DBX_E_INVALID; }
  for (int k = 0; k < n; ++k) { int rc = enqueueStep(dt, vi, pi, false); if (rc < 0) return rc; }
  return checkDeviceError(true);
}

// Benchmark helper: n steps, each bracketed by CUDA events ON THE WORLD'S STREAM; optionally a >L2-sized buffer is
// overwritten between steps (outside the timed brackets) so no step starts with the previous step's lines in L2.
// totalMs = sum of the per-step brackets; stageMs[9] = average per step of
// {collide, islands+integrate, colour+sort, prepare, solve, sync_fixtures, find_new_contacts, toi, clear_forces}.
int World::timeSteps(float dt, int vi, int pi, int n, bool flushL2, float* totalMs, float* stageMs) {
  if (!ok_) return DBX_E_NO_DEVICE;
  cudaSetDevice(device_);
  if (midStep_) { set_last_error("time_steps: a split step is open (step_begin without step_end)"); return DBX_E_INVALID; }
  const size_t flushBytes = 256u << 20;
  if (flushL2) CUDA_OR_FAIL(flushBuf_.reserve(flushBytes, false, stream_), "flush buffer");
  double total = 0.0, stage[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < n; ++k) {
    int rc = enqueueStep(dt, vi, pi, stageMs != nullptr); if (rc < 0) return rc;
    if (bodies_.empty()) continue;
    if (flushL2) CUDA_OR_FAIL(cudaMemsetAsync(flushBuf_.p, k & 0xFF, flushBytes, stream_), "l2 flush");
    CUDA_OR_FAIL(cudaEventSynchronize(ev_[9]), "event sync");
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ev_[0], ev_[9]);
    total += ms;
    if (stageMs) for (int i = 0; i < 9; ++i) { float t = 0.0f; cudaEventElapsedTime(&t, ev_[i], ev_[i + 1]); stage[i] += t; }
  }
  if (totalMs) *totalMs = (float)total;
  if (stageMs) for (int i = 0; i < 9; ++i) stageMs[i] = n > 0 ? (float)(stage[i] / n) : 0.0f;
  return checkDeviceError(true);
}

// RL-style per-step I/O: add (fx, fy, torque) to every awake dynamic body (b2Body.ApplyForceToCenter + ApplyTorque with
// wake = false, b2body.d:390-431) from a host buffer, and read all body transforms back into a host buffer.
int World::applyForces(const float* f4, int n) {
  int rc = push(); if (rc < 0) return rc;
  if (n > (int)bodies_.size() * nWorlds_) return DBX_E_INVALID;
  CUDA_OR_FAIL(ioBuf_.reserve(std::max<size_t>(bodies_.size() * (size_t)nWorlds_, 1), false, stream_), "io buffer");
  CUDA_OR_FAIL(cudaMemcpyAsync(ioBuf_.p, f4, (size_t)n * ioRecordBytes(), cudaMemcpyHostToDevice, stream_), "forces h2d");
  if (ioCompact_) CUDA_OR_FAIL(launch_apply_forces3(dw_, L_, (const float*)ioBuf_.p, n), "apply_forces");
  else CUDA_OR_FAIL(launch_apply_forces(dw_, L_, ioBuf_.p, n), "apply_forces");
  hostBodiesValid_ = false, ++bodyEpoch_;
  return n;
}
// bulk SetTransform / SetLinearVelocity / SetAngularVelocity from host arrays (either may be null); ids null = bodies 0..n-1
int World::setBodyStates(const int* ids, const float* pose4, const float* vel4, int n) {
  int rc = push(); if (rc < 0) return rc;
  const size_t nAll = bodies_.size() * (size_t)nWorlds_;
  if (n < 0 || (size_t)n > nAll || (!pose4 && !vel4)) return DBX_E_INVALID;
  if (n == 0) return 0;
  CUDA_OR_FAIL(ioBuf_.reserve(std::max<size_t>(nAll, 1), false, stream_), "io buffer");
  CUDA_OR_FAIL(ioBuf2_.reserve(std::max<size_t>(nAll, 1), false, stream_), "io buffer");
  CUDA_OR_FAIL(ioIds_.reserve(std::max<size_t>(nAll, 1), false, stream_), "io ids");
  if (ids) CUDA_OR_FAIL(cudaMemcpyAsync(ioIds_.p, ids, (size_t)n * 4, cudaMemcpyHostToDevice, stream_), "ids h2d");
  if (pose4) CUDA_OR_FAIL(cudaMemcpyAsync(ioBuf_.p, pose4, (size_t)n * 16, cudaMemcpyHostToDevice, stream_), "pose h2d");
  if (vel4) CUDA_OR_FAIL(cudaMemcpyAsync(ioBuf2_.p, vel4, (size_t)n * 16, cudaMemcpyHostToDevice, stream_), "vel h2d");
  CUDA_OR_FAIL(launch_set_states(dw_, L_, ids ? ioIds_.p : nullptr, pose4 ? ioBuf_.p : nullptr, vel4 ? ioBuf2_.p : nullptr, n), "set_states");
  hostBodiesValid_ = false, ++bodyEpoch_; hostProxiesValid_ = false;
  return n;
}
// world queries (b2world.d:563-587) see the world as it is now: pending host edits are pushed, and the LBVH is rebuilt if
// the proxy set changed under it, otherwise widened for whatever is in the move buffer
int World::refreshTreeForQuery() {
  if (replicated_) { set_last_error("world queries on a replicated world are not built (replicas share coordinates)"); return DBX_E_UNSUPPORTED; }
  int rc = push(); if (rc < 0) return rc;
  if (proxies_.empty()) return 0;
  const bool rebuild = !treeValid_;
  CUDA_OR_FAIL(stage_refresh_tree(dw_, L_, rebuild), "refresh tree");
  if (rebuild) { treeValid_ = true; sinceRebuild_ = 0; }
  return 0;
}
int World::rayCastClosest(const dbx_ray* rays, int n, dbx_ray_hit* out) {
  if (n < 0 || (n > 0 && (!rays || !out))) return DBX_E_INVALID;
  if (n == 0) return 0;
  int rc = refreshTreeForQuery(); if (rc < 0) return rc;
  CUDA_OR_FAIL(qIn_.reserve((size_t)n, false, stream_), "rays"); CUDA_OR_FAIL(qOut_.reserve(2 * (size_t)n, false, stream_), "hits");
  CUDA_OR_FAIL(cudaMemcpyAsync(qIn_.p, rays, (size_t)n * 16, cudaMemcpyHostToDevice, stream_), "rays h2d");
  CUDA_OR_FAIL(launch_raycast(dw_, L_, qIn_.p, n, qOut_.p), "raycast");
  std::vector<float4> h(2 * (size_t)n);
  CUDA_OR_FAIL(cudaMemcpyAsync(h.data(), qOut_.p, h.size() * 16, cudaMemcpyDeviceToHost, stream_), "hits d2h");
  rc = checkDeviceError(true); if (rc < 0) return DBX_E_CAPACITY;
  for (int k = 0; k < n; ++k) {
    const float4 a = h[2 * k], b = h[2 * k + 1];
    dbx_ray_hit& o = out[k];
    std::memcpy(&o.fixture, &a.x, 4); std::memcpy(&o.child, &a.y, 4);
    o.fraction = a.z; o.point = dbx_vec2{a.w, b.x}; o.normal = dbx_vec2{b.y, b.z};
    if (o.fixture < 0) { o.fraction = 1.0f; o.point = dbx_vec2{0, 0}; o.normal = dbx_vec2{0, 0}; o.child = 0; }
  }
  return n;
}
int World::queryAabb(const dbx_aabb* boxes, int n, int capPer, int32_t* counts, int32_t* fixtureChild) {
  if (n < 0 || capPer < 0 || (n > 0 && (!boxes || !counts)) || (n > 0 && capPer > 0 && !fixtureChild)) return DBX_E_INVALID;
  if (n == 0) return 0;
  int rc = refreshTreeForQuery(); if (rc < 0) return rc;
  const size_t np = (size_t)n * (size_t)std::max(capPer, 1);
  CUDA_OR_FAIL(qIn_.reserve((size_t)n, false, stream_), "boxes"); CUDA_OR_FAIL(qCount_.reserve((size_t)n, false, stream_), "counts"); CUDA_OR_FAIL(qPairs_.reserve(np, false, stream_), "pairs");
  CUDA_OR_FAIL(cudaMemcpyAsync(qIn_.p, boxes, (size_t)n * 16, cudaMemcpyHostToDevice, stream_), "boxes h2d");
  CUDA_OR_FAIL(launch_query_aabb(dw_, L_, qIn_.p, n, capPer, qCount_.p, qPairs_.p), "query_aabb");
  std::vector<int2> hp(np);
  CUDA_OR_FAIL(cudaMemcpyAsync(counts, qCount_.p, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_), "counts d2h");
  CUDA_OR_FAIL(cudaMemcpyAsync(hp.data(), qPairs_.p, np * 8, cudaMemcpyDeviceToHost, stream_), "pairs d2h");
  rc = checkDeviceError(true); if (rc < 0) return DBX_E_CAPACITY;
  for (int k = 0; k < n && capPer > 0; ++k) {   // traversal order means nothing: report sorted by (fixture, child)
    const int m = std::min(counts[k], capPer);
    int2* b = hp.data() + (size_t)k * capPer;
    std::sort(b, b + m, [](const int2& x, const int2& y) { return x.x != y.x ? x.x < y.x : x.y < y.y; });
    for (int i = 0; i < m; ++i) { fixtureChild[2 * ((size_t)k * capPer + i)] = b[i].x; fixtureChild[2 * ((size_t)k * capPer + i) + 1] = b[i].y; }
  }
  return n;
}

// b2World.GetTreeHeight / GetTreeBalance / GetTreeQuality for the LBVH (diagnostics; the tree comes down once)
int World::treeStats(int32_t* height, int32_t* maxBalance, float* quality) {
  if (height) *height = 0; if (maxBalance) *maxBalance = 0; if (quality) *quality = 0.0f;
  int rc = refreshTreeForQuery(); if (rc < 0) return rc;
  const int n = (int)proxies_.size() * nWorlds_;
  if (n == 0) return 0;
  std::vector<float4> box((size_t)2 * n - 1); std::vector<int2> child((size_t)std::max(n - 1, 1));
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(box.data(), bv_box.p, box.size() * 16, cudaMemcpyDeviceToHost), "tree d2h");
  if (n > 1) CUDA_OR_FAIL(cudaMemcpy(child.data(), bv_child.p, (size_t)(n - 1) * 8, cudaMemcpyDeviceToHost), "tree d2h");
  auto perimeter = [](const float4& b) { return 2.0f * ((b.z - b.x) + (b.w - b.y)); };
  // internal nodes 0 .. n-2 (root 0), leaves n-1 .. 2n-2; heights by an explicit post-order walk
  std::vector<int> h((size_t)2 * n - 1, 0), stack;
  std::vector<char> seen((size_t)2 * n - 1, 0);
  int balance = 0;
  double total = 0.0;
  const int root = n == 1 ? 0 : 0;
  if (n == 1) { total = perimeter(box[0]); }
  else {
    stack.push_back(root);
    while (!stack.empty()) {
      const int v = stack.back();
      if (v >= n - 1) { h[v] = 0; total += perimeter(box[v]); stack.pop_back(); continue; }
      if (!seen[v]) { seen[v] = 1; stack.push_back(child[v].x); stack.push_back(child[v].y); continue; }
      stack.pop_back();
      const int a = h[child[v].x], b = h[child[v].y];
      h[v] = 1 + std::max(a, b);
      if (h[v] > 1) balance = std::max(balance, std::abs(a - b));
      total += perimeter(box[v]);
    }
  }
  if (height) *height = h[root];
  if (maxBalance) *maxBalance = balance;
  const float rootP = perimeter(box[root]);
  if (quality) *quality = rootP > 0.0f ? (float)(total / rootP) : 0.0f;
  return 0;
}
// b2World.RayCast with the "report everything" callback (see include/dbox_b200.h): all hits per ray, sorted by (fraction, fixture, child)
int World::rayCastAll(const dbx_ray* rays, int n, int capPer, int32_t* counts, dbx_ray_hit* hits) {
  if (n < 0 || capPer < 0 || (n > 0 && (!rays || !counts)) || (n > 0 && capPer > 0 && !hits)) return DBX_E_INVALID;
  if (n == 0) return 0;
  int rc = refreshTreeForQuery(); if (rc < 0) return rc;
  const size_t nh = (size_t)n * (size_t)std::max(capPer, 1);
  CUDA_OR_FAIL(qIn_.reserve((size_t)n, false, stream_), "rays"); CUDA_OR_FAIL(qCount_.reserve((size_t)n, false, stream_), "counts"); CUDA_OR_FAIL(qOut_.reserve(2 * nh, false, stream_), "hits");
  CUDA_OR_FAIL(cudaMemcpyAsync(qIn_.p, rays, (size_t)n * 16, cudaMemcpyHostToDevice, stream_), "rays h2d");
  if (proxies_.empty()) { for (int k = 0; k < n; ++k) counts[k] = 0; return n; }
  CUDA_OR_FAIL(launch_raycast_all(dw_, L_, qIn_.p, n, capPer, qCount_.p, qOut_.p), "raycast_all");
  std::vector<float4> h(2 * nh);
  CUDA_OR_FAIL(cudaMemcpyAsync(counts, qCount_.p, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_), "counts d2h");
  CUDA_OR_FAIL(cudaMemcpyAsync(h.data(), qOut_.p, h.size() * 16, cudaMemcpyDeviceToHost, stream_), "hits d2h");
  rc = checkDeviceError(true); if (rc < 0) return DBX_E_CAPACITY;
  std::vector<dbx_ray_hit> tmp;
  for (int k = 0; k < n && capPer > 0; ++k) {
    const int m = std::min(counts[k], capPer);
    tmp.resize((size_t)m);
    for (int i = 0; i < m; ++i) {
      const float4 a = h[2 * ((size_t)k * capPer + i)], b = h[2 * ((size_t)k * capPer + i) + 1];
      dbx_ray_hit& o = tmp[i];
      std::memcpy(&o.fixture, &a.x, 4); std::memcpy(&o.child, &a.y, 4);
      o.fraction = a.z; o.point = dbx_vec2{a.w, b.x}; o.normal = dbx_vec2{b.y, b.z};
    }
    std::sort(tmp.begin(), tmp.end(), [](const dbx_ray_hit& x, const dbx_ray_hit& y) {
      return x.fraction != y.fraction ? x.fraction < y.fraction : x.fixture != y.fixture ? x.fixture < y.fixture : x.child < y.child; });
    for (int i = 0; i < m; ++i) hits[(size_t)k * capPer + i] = tmp[i];
  }
  return n;
}
// b2Fixture.TestPoint (b2fixture.d:209-212), batched.  The fixture's geometry record is the one its proxy uses; a fixture
// without proxies (inactive body) gets its record interned here.  Chains (and edges) contain no point.
int World::testPoints(const int32_t* fixtures, const dbx_vec2* points, int n, int32_t* inside) {
  if (n < 0 || (n > 0 && (!fixtures || !points || !inside))) return DBX_E_INVALID;
  if (n == 0) return 0;
  if (replicated_) { set_last_error("world queries on a replicated world are not built (replicas share coordinates)"); return DBX_E_UNSUPPORTED; }
  std::vector<float4> q((size_t)n);
  for (int k = 0; k < n; ++k) {
    const int f = fixtures[k];
    if (f < 0 || f >= (int)fixtures_.size() || !fixtures_[f].alive) return DBX_E_INVALID;
    const HFixture& hf = fixtures_[f];
    int shape = -1;
    if (hf.shape.s.type == DBX_SHAPE_CIRCLE || hf.shape.s.type == DBX_SHAPE_POLYGON) {
      if (!hf.proxies.empty()) shape = proxies_[hf.proxies[0]].shape;
      else { DShape ds; buildChildShape(hf.shape, 0, &ds); shape = internShape(ds); }
    }
    float fs, fb; std::memcpy(&fs, &shape, 4); std::memcpy(&fb, &hf.body, 4);
    q[k] = make_float4(points[k].x, points[k].y, fs, fb);
  }
  int rc = push(); if (rc < 0) return rc;
  CUDA_OR_FAIL(qIn_.reserve((size_t)n, false, stream_), "points"); CUDA_OR_FAIL(qCount_.reserve((size_t)n, false, stream_), "inside");
  CUDA_OR_FAIL(cudaMemcpyAsync(qIn_.p, q.data(), (size_t)n * 16, cudaMemcpyHostToDevice, stream_), "points h2d");
  CUDA_OR_FAIL(launch_test_points(dw_, L_, qIn_.p, n, qCount_.p), "test_points");
  CUDA_OR_FAIL(cudaMemcpyAsync(inside, qCount_.p, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_), "inside d2h");
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  return n;
}
// b2World.ShiftOrigin (b2world.d:758-780).  Joint anchors kept in world coordinates live in the host definitions (mouse target
// b2mousejoint.d:174-177, pulley ground anchors b2pulleyjoint.d:227-231) and are re-uploaded with the impulses pulled first.
int World::shiftOrigin(float x, float y) {
  if (midStep_) return DBX_E_INVALID;
  int rc = push(); if (rc < 0) return rc;
  if (bodies_.empty()) return 0;
  bool anyJoint = false;
  for (const HJoint& j : joints_) if (j.alive && (j.def.type == DBX_JOINT_MOUSE || j.def.type == DBX_JOINT_PULLEY)) anyJoint = true;
  if (anyJoint) {
    if (replicated_) { set_last_error("world is replicated: joints with world-space anchors cannot be shifted"); return DBX_E_UNSUPPORTED; }
    rc = pullJoints(); if (rc < 0) return rc;
    for (HJoint& j : joints_) {
      if (!j.alive) continue;
      if (j.def.type == DBX_JOINT_MOUSE) { j.def.target.x -= x; j.def.target.y -= y; }
      else if (j.def.type == DBX_JOINT_PULLEY) { j.def.groundAnchorA.x -= x; j.def.groundAnchorA.y -= y; j.def.groundAnchorB.x -= x; j.def.groundAnchorB.y -= y; }
    }
    fullPushJoints_ = true;
  }
  CUDA_OR_FAIL(launch_shift_origin(dw_, L_, x, y), "shift_origin");
  treeValid_ = false;                       // the LBVH boxes are stale: rebuilt before the next pair query / world query
  hostBodiesValid_ = false, ++bodyEpoch_; hostProxiesValid_ = false;
  if (anyJoint) { rc = push(); if (rc < 0) return rc; }
  return 0;
}
// b2Contact.GetWorldManifold for every contact, in readContacts order (ascending reference pair key)
int World::readWorldManifolds(dbx_world_manifold* out, int cap) {
  if (!dw_.hdr || bodiesSynced_ == 0) return 0;
  int rc = push(); if (rc < 0) return rc;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  int high = 0;
  CUDA_OR_FAIL(cudaMemcpy(&high, (char*)hdr_.p + offsetof(Header, cHigh), 4, cudaMemcpyDeviceToHost), "read cHigh");
  if (high <= 0) return 0;
  const size_t n = (size_t)high;
  CUDA_OR_FAIL(qOut_.reserve(2 * n, false, stream_), "world manifolds");
  CUDA_OR_FAIL(launch_world_manifolds(dw_, L_, high, qOut_.p), "world_manifolds");
  std::vector<float4> wm(2 * n); std::vector<unsigned long long> key(n); std::vector<uint32_t> fl(n); std::vector<uint4> mk(n);
  CUDA_OR_FAIL(cudaMemcpyAsync(wm.data(), qOut_.p, 2 * n * 16, cudaMemcpyDeviceToHost, stream_), "world manifolds d2h");
  CUDA_OR_FAIL(cudaMemcpyAsync(key.data(), c_key.p, n * 8, cudaMemcpyDeviceToHost, stream_), "keys d2h");
  CUDA_OR_FAIL(cudaMemcpyAsync(fl.data(), c_flags.p, n * 4, cudaMemcpyDeviceToHost, stream_), "flags d2h");
  CUDA_OR_FAIL(cudaMemcpyAsync(mk.data(), c_mk.p, n * 16, cudaMemcpyDeviceToHost, stream_), "mk d2h");
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  std::vector<int> order;
  for (size_t i = 0; i < n; ++i) if (fl[i] & CF_ALIVE) order.push_back((int)i);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  int cnt = 0;
  for (int i : order) {
    if (cnt < cap) {
      dbx_world_manifold& o = out[cnt];
      const float4 a = wm[2 * (size_t)i], b = wm[2 * (size_t)i + 1];
      o.normal = dbx_vec2{a.x, a.y}; o.separations[0] = a.z; o.separations[1] = a.w;
      o.points[0] = dbx_vec2{b.x, b.y}; o.points[1] = dbx_vec2{b.z, b.w};
      o.pointCount = (int)mk[i].w; o._pad = 0;
    }
    ++cnt;
  }
  return cnt;
}
// b2ContactListener.PostSolve, deferred (see include/dbox_b200.h): capacity > 0 turns recording on, 0 off
int World::enablePostSolve(int capacity) {
  if (capacity < 0) return DBX_E_INVALID;
  int rc = push(); if (rc < 0) return rc;
  if (capacity > 0) {
    CUDA_OR_FAIL(ps_a_.reserve((size_t)capacity, false, stream_), "post-solve"); CUDA_OR_FAIL(ps_b_.reserve((size_t)capacity, false, stream_), "post-solve");
    CUDA_OR_FAIL(ps_key_.reserve((size_t)capacity, false, stream_), "post-solve");
  }
  dw_.ps_a = ps_a_.p; dw_.ps_b = ps_b_.p; dw_.ps_key = ps_key_.p;
  dw_.psCap = capacity > 0 ? (int)std::min(ps_a_.cap, std::min(ps_b_.cap, ps_key_.cap)) : 0;
  if (dw_.hdr) { CUDA_OR_FAIL(cudaMemsetAsync((char*)hdr_.p + offsetof(Header, nPostSolve), 0, 4, stream_), "post-solve reset"); CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync"); }
  return dw_.psCap;
}
int World::readPostSolve(dbx_post_solve* out, int cap) {
  if (dw_.psCap == 0 || !dw_.hdr) return 0;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  int n = 0;
  CUDA_OR_FAIL(cudaMemcpy(&n, (char*)hdr_.p + offsetof(Header, nPostSolve), 4, cudaMemcpyDeviceToHost), "post-solve count");
  if (n > dw_.psCap) { set_last_error("post-solve buffer overflow: " + std::to_string(n) + " records, capacity " + std::to_string(dw_.psCap)); return DBX_E_CAPACITY; }
  if (!out || cap <= 0 || n == 0) return n;
  std::vector<int4> a((size_t)n); std::vector<float4> b((size_t)n); std::vector<unsigned long long> key((size_t)n);
  CUDA_OR_FAIL(cudaMemcpy(a.data(), ps_a_.p, (size_t)n * 16, cudaMemcpyDeviceToHost), "post-solve d2h");
  CUDA_OR_FAIL(cudaMemcpy(b.data(), ps_b_.p, (size_t)n * 16, cudaMemcpyDeviceToHost), "post-solve d2h");
  CUDA_OR_FAIL(cudaMemcpy(key.data(), ps_key_.p, (size_t)n * 8, cudaMemcpyDeviceToHost), "post-solve d2h");
  std::vector<int> order((size_t)n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int x, int y) {
    const int px = (a[x].w >> 8) & 0xFF, py = (a[y].w >> 8) & 0xFF;
    if (px != py) return px < py;
    // the island solve's records arrive in solver-slot order (meaningless): by key; TOI sub-steps: by key, then arrival
    if (key[x] != key[y]) return key[x] < key[y];
    return x < y;
  });
  for (int k = 0; k < n && k < cap; ++k) {
    const int i = order[k];
    dbx_post_solve& o = out[k];
    o.fixtureA = a[i].x; o.fixtureB = a[i].y; o.childA = a[i].z & 0xFFFF; o.childB = (a[i].z >> 16) & 0xFFFF;
    o.count = a[i].w & 0xFF; o.phase = (a[i].w >> 8) & 0xFF;
    o.normalImpulses[0] = b[i].x; o.tangentImpulses[0] = b[i].y; o.normalImpulses[1] = b[i].z; o.tangentImpulses[1] = b[i].w;
  }
  return n;
}

// user b2ContactFilter, deferred (see include/dbox_b200.h)
int World::setUserFilter(int mode) {
  if (mode < 0 || mode > 3) return DBX_E_INVALID;
  if (replicated_ && mode != 0) { set_last_error("world is replicated: per-contact vetoes address fixtures of the template world only"); return DBX_E_UNSUPPORTED; }
  dw_.userFilter = mode;
  return 0;
}
int World::pollNewContacts(int32_t* out, int cap) {
  if (!dw_.hdr || bodiesSynced_ == 0) return 0;
  int rc = push(); if (rc < 0) return rc;
  if (newFixture_) { int rf = findNewContacts(); if (rf < 0) return rf; newFixture_ = false; }   // pairs of fixtures added since the last step
  const bool peek = !out || cap <= 0;
  int high = 0;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(&high, (char*)hdr_.p + offsetof(Header, cHigh), 4, cudaMemcpyDeviceToHost), "read cHigh");
  if (high <= 0) return 0;
  const size_t room = peek ? 1 : (size_t)cap;
  CUDA_OR_FAIL(qPairs_.reserve(2 * room, false, stream_), "new contacts"); CUDA_OR_FAIL(patchKeys_.reserve(room, false, stream_), "new contacts");
  CUDA_OR_FAIL(cudaMemsetAsync((char*)hdr_.p + offsetof(Header, nNewContacts), 0, 4, stream_), "new contacts reset");
  CUDA_OR_FAIL(launch_list_new_contacts(dw_, L_, (int4*)qPairs_.p, patchKeys_.p, peek ? 0 : cap, peek ? 1 : 0), "list_new_contacts");
  int n = 0;
  CUDA_OR_FAIL(cudaMemcpyAsync(&n, (char*)hdr_.p + offsetof(Header, nNewContacts), 4, cudaMemcpyDeviceToHost, stream_), "new contacts count");
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  if (peek || n == 0) return n;
  const int m = std::min(n, cap);            // the surplus stays tagged for the next poll
  std::vector<int4> rec((size_t)m); std::vector<unsigned long long> key((size_t)m);
  CUDA_OR_FAIL(cudaMemcpy(rec.data(), qPairs_.p, (size_t)m * 16, cudaMemcpyDeviceToHost), "new contacts d2h");
  CUDA_OR_FAIL(cudaMemcpy(key.data(), patchKeys_.p, (size_t)m * 8, cudaMemcpyDeviceToHost), "new contacts d2h");
  std::vector<int> order((size_t)m);
  for (int i = 0; i < m; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  for (int k = 0; k < m; ++k) { const int4 r = rec[order[k]]; out[4 * k] = r.x; out[4 * k + 1] = r.y; out[4 * k + 2] = r.z; out[4 * k + 3] = r.w; }
  return n;
}

// b2World.SetContactListener (dynamics/b2world.d:62-66), deferred form: capacity > 0 turns recording on, 0 off
int World::enableContactEvents(int capacity) {
  if (capacity < 0) return DBX_E_INVALID;
  int rc = push(); if (rc < 0) return rc;
  if (capacity > 0) { CUDA_OR_FAIL(ev_a_.reserve((size_t)capacity, false, stream_), "events"); CUDA_OR_FAIL(ev_b_.reserve((size_t)capacity, false, stream_), "events"); }
  dw_.ev_a = ev_a_.p; dw_.ev_b = ev_b_.p; dw_.evCap = capacity > 0 ? (int)std::min(ev_a_.cap, ev_b_.cap) : 0;
  if (dw_.hdr) { int zero = 0; CUDA_OR_FAIL(cudaMemcpyAsync((char*)hdr_.p + offsetof(Header, nCtEvents), &zero, 4, cudaMemcpyHostToDevice, stream_), "events reset"); CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync"); }
  return dw_.evCap;
}
// events since the last poll in (step, phase, pair key, type) order; the device order is whatever the atomics gave
int World::pollContactEvents(dbx_contact_event* out, int cap) {
  if (dw_.evCap == 0 || !dw_.hdr) return 0;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  int n = 0;
  CUDA_OR_FAIL(cudaMemcpy(&n, (char*)hdr_.p + offsetof(Header, nCtEvents), 4, cudaMemcpyDeviceToHost), "events count");
  if (!out || cap <= 0) return n;
  const int have = std::min(n, dw_.evCap);
  std::vector<int4> a((size_t)have), b((size_t)have);
  if (have > 0) { CUDA_OR_FAIL(cudaMemcpy(a.data(), ev_a_.p, (size_t)have * 16, cudaMemcpyDeviceToHost), "events"); CUDA_OR_FAIL(cudaMemcpy(b.data(), ev_b_.p, (size_t)have * 16, cudaMemcpyDeviceToHost), "events"); }
  int zero = 0; CUDA_OR_FAIL(cudaMemcpy((char*)hdr_.p + offsetof(Header, nCtEvents), &zero, 4, cudaMemcpyHostToDevice), "events reset");
  if (n > dw_.evCap) { set_last_error("contact event buffer overflow: events were lost, raise the capacity of dbx_world_enable_contact_events"); return DBX_E_CAPACITY; }
  // steps are stamped with 16 bits: order them relative to the current step so that a wrap does not reorder
  const int cur = stepCount_ & 0xFFFF;
  std::vector<int> order((size_t)have);
  for (int i = 0; i < have; ++i) order[i] = i;
  auto age = [&](int i) { return (cur - ((a[i].x >> 16) & 0xFFFF)) & 0xFFFF; };
  auto key = [&](int i) { return ((unsigned long long)(unsigned)b[i].w << 32) | (unsigned)b[i].z; };
  std::sort(order.begin(), order.end(), [&](int x, int y) {
    if (age(x) != age(y)) return age(x) > age(y);
    const int px = (a[x].x >> 8) & 0xFF, py = (a[y].x >> 8) & 0xFF;
    if (px != py) return px < py;
    if (key(x) != key(y)) return key(x) < key(y);
    return (a[x].x & 0xFF) < (a[y].x & 0xFF);
  });
  const int m = std::min(have, cap);
  for (int k = 0; k < m; ++k) {
    const int i = order[k];
    dbx_contact_event& e = out[k];
    e.type = a[i].x & 0xFF; e.phase = (a[i].x >> 8) & 0xFF; e.stepsAgo = age(i);
    e.fixtureA = a[i].y; e.fixtureB = a[i].z; e.childA = a[i].w & 0xFFFF; e.childB = (a[i].w >> 16) & 0xFFFF;
    e.bodyA = b[i].x; e.bodyB = b[i].y;
  }
  return have;
}
int World::readTransforms(float* out, int n) {
  int rc = push(); if (rc < 0) return rc;
  if (n > (int)bodies_.size() * nWorlds_) return DBX_E_INVALID;
  if (ioCompact_) {
    CUDA_OR_FAIL(ioBuf2_.reserve(std::max<size_t>(bodies_.size() * (size_t)nWorlds_, 1), false, stream_), "io buffer");
    CUDA_OR_FAIL(launch_pack_poses(dw_, L_, (float*)ioBuf2_.p, n), "pack poses");
    CUDA_OR_FAIL(cudaMemcpyAsync(out, ioBuf2_.p, (size_t)n * 12, cudaMemcpyDeviceToHost, stream_), "poses d2h");
  } else CUDA_OR_FAIL(cudaMemcpyAsync(out, b_xf.p, (size_t)n * 16, cudaMemcpyDeviceToHost, stream_), "xf d2h");
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  return n;
}

// ---- pipelined stepping and bulk I/O (see include/dbox_b200.h) ----
int World::ensureIoStreams() {
  if (h2d_) return 0;
  CUDA_OR_FAIL(cudaStreamCreateWithFlags(&h2d_, cudaStreamNonBlocking), "h2d stream");
  CUDA_OR_FAIL(cudaStreamCreateWithFlags(&d2h_, cudaStreamNonBlocking), "d2h stream");
  for (int i = 0; i < 2; ++i) {
    CUDA_OR_FAIL(cudaEventCreateWithFlags(&inCopied_[i], cudaEventDisableTiming), "io event");
    CUDA_OR_FAIL(cudaEventCreateWithFlags(&inRead_[i], cudaEventDisableTiming), "io event");
    CUDA_OR_FAIL(cudaEventCreateWithFlags(&snapReady_[i], cudaEventDisableTiming), "io event");
    CUDA_OR_FAIL(cudaEventCreateWithFlags(&outDone_[i], cudaEventDisableTiming), "io event");
  }
  return 0;
}
int World::stepAsync(float dt, int vi, int pi) {
  if (!ok_) return DBX_E_NO_DEVICE;
  if (midStep_) return DBX_E_INVALID;
  cudaSetDevice(device_);
  return enqueueStep(dt, vi, pi, false);
}
int World::applyForcesAsync(const float* f4, int n) {
  int rc = push(); if (rc < 0) return rc;
  const size_t nAll = bodies_.size() * (size_t)nWorlds_;
  if (n < 0 || (size_t)n > nAll || (n > 0 && !f4)) return DBX_E_INVALID;
  if (n == 0) return 0;
  rc = ensureIoStreams(); if (rc < 0) return rc;
  const int k = inFlip_; inFlip_ ^= 1;
  if (inStage_[k].cap < nAll) {               // (re)allocation zero-fills on the world's stream: let that land before the copy stream writes
    CUDA_OR_FAIL(inStage_[k].reserve(std::max<size_t>(nAll, 1), false, stream_), "force staging");
    CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
    inReadValid_[k] = false;
  }
  if (inReadValid_[k]) CUDA_OR_FAIL(cudaStreamWaitEvent(h2d_, inRead_[k], 0), "staging reuse");   // its previous consumer has run
  CUDA_OR_FAIL(cudaMemcpyAsync(inStage_[k].p, f4, (size_t)n * ioRecordBytes(), cudaMemcpyHostToDevice, h2d_), "forces h2d");
  CUDA_OR_FAIL(cudaEventRecord(inCopied_[k], h2d_), "io event");
  CUDA_OR_FAIL(cudaStreamWaitEvent(stream_, inCopied_[k], 0), "forces ready");
  if (ioCompact_) CUDA_OR_FAIL(launch_apply_forces3(dw_, L_, (const float*)inStage_[k].p, n), "apply_forces");
  else CUDA_OR_FAIL(launch_apply_forces(dw_, L_, inStage_[k].p, n), "apply_forces");
  CUDA_OR_FAIL(cudaEventRecord(inRead_[k], stream_), "io event");
  inReadValid_[k] = true;
  hostBodiesValid_ = false, ++bodyEpoch_;
  return n;
}
int World::readTransformsAsync(float* out, int n) {
  int rc = push(); if (rc < 0) return rc;
  const size_t nAll = bodies_.size() * (size_t)nWorlds_;
  if (n < 0 || (size_t)n > nAll || (n > 0 && !out)) return DBX_E_INVALID;
  rc = ensureIoStreams(); if (rc < 0) return rc;
  const int ticket = ++ioTicket_;
  const int k = ticket & 1;
  if (n == 0) return ticket;
  if (outSnap_[k].cap < nAll) {
    if (outDoneValid_[k]) CUDA_OR_FAIL(cudaEventSynchronize(outDone_[k]), "io wait");
    CUDA_OR_FAIL(outSnap_[k].reserve(std::max<size_t>(nAll, 1), false, stream_), "transform snapshot");
    outDoneValid_[k] = false;
  }
  if (outDoneValid_[k]) CUDA_OR_FAIL(cudaStreamWaitEvent(stream_, outDone_[k], 0), "snapshot reuse");   // the read two tickets ago has left
  if (ioCompact_) CUDA_OR_FAIL(launch_pack_poses(dw_, L_, (float*)outSnap_[k].p, n), "pose snapshot");
  else CUDA_OR_FAIL(cudaMemcpyAsync(outSnap_[k].p, b_xf.p, (size_t)n * 16, cudaMemcpyDeviceToDevice, stream_), "xf snapshot");
  CUDA_OR_FAIL(cudaEventRecord(snapReady_[k], stream_), "io event");
  CUDA_OR_FAIL(cudaStreamWaitEvent(d2h_, snapReady_[k], 0), "snapshot ready");
  CUDA_OR_FAIL(cudaMemcpyAsync(out, outSnap_[k].p, (size_t)n * ioRecordBytes(), cudaMemcpyDeviceToHost, d2h_), "xf d2h");
  CUDA_OR_FAIL(cudaEventRecord(outDone_[k], d2h_), "io event");
  outDoneValid_[k] = true;
  return ticket;
}
int World::ioWait(int ticket) {
  if (ticket <= 0 || ticket > ioTicket_) return DBX_E_INVALID;
  // a ticket older than the last two shares its slot with a newer read, which was ordered after it
  const int k = ticket & 1;
  if (outDoneValid_[k]) CUDA_OR_FAIL(cudaEventSynchronize(outDone_[k]), "io wait");
  return 0;
}
int World::sync() {
  if (!ok_) return DBX_E_NO_DEVICE;
  if (h2d_) { CUDA_OR_FAIL(cudaStreamSynchronize(h2d_), "sync"); }
  int rc = checkDeviceError(true);
  if (d2h_) { CUDA_OR_FAIL(cudaStreamSynchronize(d2h_), "sync"); }
  return rc;
}

int World::clearForces() {
  int rc = push(); if (rc < 0) return rc;
  if (bodies_.empty()) return 0;
  CUDA_OR_FAIL(launch_clear_forces(dw_, L_), "clear_forces");
  hostBodiesValid_ = false, ++bodyEpoch_;
  return 0;
}

int World::stageFindNewContacts() {
  int rc = push(); if (rc < 0) return rc;
  if (bodies_.empty()) return 0;
  { int rf = findNewContacts(); if (rf < 0) return rf; }
  newFixture_ = false;
  hostBodiesValid_ = false, ++bodyEpoch_; hostProxiesValid_ = false;
  return checkDeviceError(true);
}

int World::stageCollide() {
  int rc = push(); if (rc < 0) return rc;
  if (bodies_.empty()) return 0;
  setStepParams(0.0f, 0, 0);
  CUDA_OR_FAIL(stage_collide(dw_, L_), "collide");
  hostBodiesValid_ = false, ++bodyEpoch_;
  return checkDeviceError(true);
}

// ------------------------------------------------------------------------------------------------ accessors
// body states straight from the device arrays (replicated worlds have no per-replica host mirror)
int World::readBodiesDevice(int from, int count, dbx_body_state* out) {
  const size_t n = (size_t)count;
  std::vector<float4> xf(n), pos(n), pos0(n), vel(n), frc(n), ms(n), lc(n); std::vector<float2> gs(n); std::vector<uint32_t> fl(n);
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  cudaMemcpy(xf.data(), b_xf.p + from, n * 16, cudaMemcpyDeviceToHost); cudaMemcpy(pos.data(), b_pos.p + from, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(pos0.data(), b_pos0.p + from, n * 16, cudaMemcpyDeviceToHost); cudaMemcpy(vel.data(), b_vel.p + from, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(frc.data(), b_force.p + from, n * 16, cudaMemcpyDeviceToHost); cudaMemcpy(ms.data(), b_mass.p + from, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(lc.data(), b_lc.p + from, n * 16, cudaMemcpyDeviceToHost); cudaMemcpy(gs.data(), b_gs.p + from, n * 8, cudaMemcpyDeviceToHost);
  CUDA_OR_FAIL(cudaMemcpy(fl.data(), b_flags.p + from, n * 4, cudaMemcpyDeviceToHost), "read bodies");
  for (size_t i = 0; i < n; ++i) {
    dbx_body_state& st = out[i];
    st.p = dbx_vec2{xf[i].x, xf[i].y}; st.qs = xf[i].z; st.qc = xf[i].w;
    st.c = dbx_vec2{pos[i].x, pos[i].y}; st.a = pos[i].z;
    st.c0 = dbx_vec2{pos0[i].x, pos0[i].y}; st.a0 = pos0[i].z; st.alpha0 = pos0[i].w;
    st.v = dbx_vec2{vel[i].x, vel[i].y}; st.w = vel[i].z;
    st.force = dbx_vec2{frc[i].x, frc[i].y}; st.torque = frc[i].z;
    st.invMass = ms[i].x; st.invI = ms[i].y; st.mass = ms[i].z; st.I = ms[i].w;
    st.localCenter = dbx_vec2{lc[i].x, lc[i].y}; st.linearDamping = lc[i].z; st.angularDamping = lc[i].w;
    st.gravityScale = gs[i].x; st.sleepTime = gs[i].y;
    st.flags = fl[i] & 0xFFFF; st.type = body_type(fl[i]);
  }
  return count;
}

int World::getBody(int b, dbx_body_state* out) {
  if (replicated_) {
    if (b < 0 || b >= (int)bodies_.size() * nWorlds_ || !bodies_[b % (int)bodies_.size()].alive) return DBX_E_INVALID;
    int rc = readBodiesDevice(b, 1, out); return rc < 0 ? rc : 0;
  }
  if (b < 0 || b >= (int)bodies_.size() || !bodies_[b].alive) return DBX_E_INVALID;
  int rc = pullBodyRow(b); if (rc < 0) return rc;         // (one row, not the world: b2Body.GetPosition in a game loop)
  *out = bodies_[b].st;
  return 0;
}

HBody* World::mutBody(int b) {
  if (replicated_) return nullptr;
  if (b < 0 || b >= (int)bodies_.size() || !bodies_[b].alive) return nullptr;
  if (pullBodies() < 0) return nullptr;
  if ((size_t)b < bodiesSynced_) fullPushBodies_ = true;
  return &bodies_[b];
}
// An edit of this body's own row (velocity, force, awake / bullet / autosleep flags): one row comes back from the device and one
// row goes out with the next push -- not every body array of the world both ways (b2Body.ApplyForce once per frame on the 100,000
// body pile used to cost 12 MB each way and a host loop over all bodies per step).  Past 64 edited bodies between two pushes the
// whole-array path is the cheaper one.
HBody* World::mutBodyRow(int b) {
  if (replicated_) return nullptr;
  if (b < 0 || b >= (int)bodies_.size() || !bodies_[b].alive) return nullptr;
  HBody& hb = bodies_[b];
  if ((size_t)b >= bodiesSynced_ || fullPushBodies_ || dirtyBodies_.size() >= 64) return mutBody(b);
  if (pullBodyRow(b) < 0) return nullptr;
  if (!hb.dirty) { hb.dirty = true; dirtyBodies_.push_back(b); }
  return &hb;
}

// b2Body.SetTransform (dynamics/b2body.d:261-285) incl. b2Fixture.Synchronize with xf1 == xf2
int World::setTransform(int b, float x, float y, float angle) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  HBody* hb = mutBody(b);
  if (!hb) return DBX_E_INVALID;
  int rc = pullProxies(); if (rc < 0) return rc;
  dbx_body_state& st = hb->st;
  Rot q = rot_from_angle(angle);
  st.qs = q.s; st.qc = q.c; st.p = dbx_vec2{x, y};
  Xf xf; xf.p = V(x, y); xf.q = q;
  v2 c = mul(xf, V(st.localCenter.x, st.localCenter.y));
  st.c = dbx_vec2{c.x, c.y}; st.a = angle; st.c0 = st.c; st.a0 = angle;
  hb->xf0 = pack(xf);
  for (auto it = hb->fixtures.rbegin(); it != hb->fixtures.rend(); ++it) {
    for (int slot : fixtures_[*it].proxies) {
      HProxy& p = proxies_[slot];
      Box box = shape_aabb(&shapes_[p.shape], xf);
      p.aabb = pack(box);
      if (!contains(BX(p.fat), box)) {   // MoveProxy with zero displacement (b2dynamictree.d:140-184)
        p.fat = make_float4(box.lo.x - kAabbExtension, box.lo.y - kAabbExtension, box.hi.x + kAabbExtension, box.hi.y + kAabbExtension);
        pendingMoves_.push_back(slot);
      }
      if ((size_t)slot < proxiesSynced_) fullPushProxies_ = true;
    }
  }
  return 0;
}

int World::counts(dbx_counts* out) {
  std::memset(out, 0, sizeof(*out));
  for (auto& b : bodies_) if (b.alive) ++out->bodies;
  for (auto& f : fixtures_) if (f.alive) ++out->fixtures;
  for (auto& p : proxies_) if (p.alive) ++out->proxies;
  for (auto& j : joints_) if (j.alive) ++out->joints;
  out->bodies *= nWorlds_; out->fixtures *= nWorlds_; out->proxies *= nWorlds_; out->joints *= nWorlds_;
  out->moves = (int)pendingMoves_.size();
  if (!dw_.hdr || bodiesSynced_ == 0) {
    for (auto& b : bodies_) if (b.alive && (b.st.flags & DBX_BODY_AWAKE) && b.st.type != DBX_STATIC_BODY) ++out->awakeBodies;
    return 0;
  }
  int rc = push(); if (rc < 0) return rc;
  CUDA_OR_FAIL(stage_count(dw_, L_), "count");
  Header h;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(&h, hdr_.p, sizeof(Header), cudaMemcpyDeviceToHost), "read header");
  out->contacts = h.nContacts; out->touching = h.nTouching; out->awakeBodies = h.nAwake; out->colours = h.nColours;
  out->islands = h.nIslands; out->pairs = h.nPairs; out->moves += h.nMoved;
  return 0;
}

int World::profile(dbx_profile* out) {
  std::memset(out, 0, sizeof(*out));
  if (!evValid_) return 0;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  float collide = 0, pre = 0, solve = 0, bp = 0, toi = 0, total = 0;
  cudaEventElapsedTime(&collide, ev_[0], ev_[1]);
  cudaEventElapsedTime(&pre, ev_[1], ev_[3]);
  cudaEventElapsedTime(&solve, ev_[3], ev_[5]);
  cudaEventElapsedTime(&bp, ev_[5], ev_[7]);
  cudaEventElapsedTime(&toi, ev_[7], ev_[8]);
  cudaEventElapsedTime(&total, ev_[0], ev_[9]);
  out->step = total; out->collide = collide; out->solve = pre + solve + bp; out->solveInit = pre; out->solveVelocity = solve;
  out->solvePosition = 0.0f; out->broadphase = bp; out->solveTOI = toi;
  // the island solver is one kernel; it stamps %globaltimer where the reference's three timers sit (b2island.d:147, 166, 237) and
  // its event time is split in those proportions (k_solve and k_solve_tiles; the world-local solver of batched replicas has no
  // single timeline and reports everything as solveVelocity)
  unsigned long long st[4] = {0, 0, 0, 0};
  CUDA_OR_FAIL(cudaMemcpy(st, (char*)hdr_.p + offsetof(Header, solveStamp), sizeof(st), cudaMemcpyDeviceToHost), "read stamps");
  if (!lastWorldsPath_ && st[3] > st[0] && st[1] >= st[0] && st[2] >= st[1] && st[3] >= st[2]) {
    const double span = (double)(st[3] - st[0]);
    const float init = (float)(solve * (double)(st[1] - st[0]) / span), pos = (float)(solve * (double)(st[3] - st[2]) / span);
    out->solveInit = pre + init; out->solvePosition = pos; out->solveVelocity = solve - init - pos;
  }
  return 0;
}

int World::readBodies(dbx_body_state* out, int cap) {
  if (replicated_) {
    const int n = (int)bodies_.size() * nWorlds_;
    if (cap > 0) { int rc = readBodiesDevice(0, std::min(cap, n), out); if (rc < 0) return rc; }
    return n;
  }
  int rc = pullBodies(); if (rc < 0) return rc;
  const int n = (int)bodies_.size();
  for (int i = 0; i < n && i < cap; ++i) { if (bodies_[i].alive) out[i] = bodies_[i].st; else std::memset(out + i, 0, sizeof(*out)); }
  return n;
}

int World::writeBodies(const dbx_body_state* in, int n) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (n > (int)bodies_.size()) return DBX_E_INVALID;
  int rc = pullBodies(); if (rc < 0) return rc;
  for (int i = 0; i < n; ++i) {
    if (!bodies_[i].alive) continue;
    bodies_[i].st = in[i];
    // xf0 is the transform at (c0, a0) (b2body.d:1131-1133)
    Xf x0 = xf_from_sweep(V(in[i].c0.x, in[i].c0.y), in[i].a0, V(in[i].localCenter.x, in[i].localCenter.y));
    if (in[i].a0 == in[i].a && in[i].c0.x == in[i].c.x && in[i].c0.y == in[i].c.y) bodies_[i].xf0 = make_float4(in[i].p.x, in[i].p.y, in[i].qs, in[i].qc);
    else bodies_[i].xf0 = pack(x0);
  }
  fullPushBodies_ = true;
  // the written sweeps may carry an alpha0 from somebody else's unfinished bookkeeping (the reference leaves alpha0 as the last
  // SolveTOI set it and resets it when the next one starts, b2world.d:1131-1137): have the next k_toi start from its reset phase
  toiClean_ = false;
  return n;
}

int World::readProxies(dbx_proxy_rec* out, int cap) {
  int rc = pullProxies(); if (rc < 0) return rc;
  int n = 0;
  for (size_t f = 0; f < fixtures_.size(); ++f) {
    if (!fixtures_[f].alive) continue;
    for (int slot : fixtures_[f].proxies) {
      if (n < cap) {
        const HProxy& p = proxies_[slot];
        dbx_proxy_rec& o = out[n];
        o.fixture = p.fixture; o.child = p.child; o.proxyId = p.key;
        o.aabb.lo = dbx_vec2{p.aabb.x, p.aabb.y}; o.aabb.hi = dbx_vec2{p.aabb.z, p.aabb.w};
        o.fat.lo = dbx_vec2{p.fat.x, p.fat.y}; o.fat.hi = dbx_vec2{p.fat.z, p.fat.w};
      }
      ++n;
    }
  }
  return n;
}

int World::writeProxies(const dbx_proxy_rec* in, int n) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  int rc = pullProxies(); if (rc < 0) return rc;
  for (int i = 0; i < n; ++i) {
    const dbx_proxy_rec& r = in[i];
    if (r.fixture < 0 || r.fixture >= (int)fixtures_.size() || !fixtures_[r.fixture].alive) return DBX_E_INVALID;
    const auto& ps = fixtures_[r.fixture].proxies;
    if (r.child < 0 || r.child >= (int)ps.size()) return DBX_E_INVALID;
    HProxy& p = proxies_[ps[r.child]];
    p.aabb = make_float4(r.aabb.lo.x, r.aabb.lo.y, r.aabb.hi.x, r.aabb.hi.y);
    p.fat = make_float4(r.fat.lo.x, r.fat.lo.y, r.fat.hi.x, r.fat.hi.y);
  }
  fullPushProxies_ = true;
  return n;
}

int World::readJoints(dbx_joint_state* out, int cap) {
  int rc = pullJoints(); if (rc < 0) return rc;
  const int n = (int)joints_.size();
  for (int i = 0; i < n && i < cap; ++i) {
    std::memset(out + i, 0, sizeof(*out));
    if (!joints_[i].alive) continue;
    out[i].type = joints_[i].def.type;
    out[i].impulse[0] = joints_[i].imp[0]; out[i].impulse[1] = joints_[i].imp[1]; out[i].impulse[2] = joints_[i].imp[2];
    out[i].motorImpulse = joints_[i].imp[3]; out[i].limitState = joints_[i].limit;
  }
  return n;
}

int World::writeJoints(const dbx_joint_state* in, int n) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  if (n > (int)joints_.size()) return DBX_E_INVALID;
  int rc = pullJoints(); if (rc < 0) return rc;
  for (int i = 0; i < n; ++i) {
    joints_[i].imp[0] = in[i].impulse[0]; joints_[i].imp[1] = in[i].impulse[1]; joints_[i].imp[2] = in[i].impulse[2];
    joints_[i].imp[3] = in[i].motorImpulse; joints_[i].limit = in[i].limitState;
  }
  fullPushJoints_ = true;
  return n;
}

int World::readMoves(int32_t* out, int cap) {
  // host-buffered moves (the device move list is empty between steps)
  int n = 0;
  std::vector<int> mv = pendingMoves_;
  mv.insert(mv.end(), movesOnDevice_.begin(), movesOnDevice_.end());      // already uploaded by an earlier push, not yet consumed
  std::sort(mv.begin(), mv.end());
  mv.erase(std::unique(mv.begin(), mv.end()), mv.end());
  for (int slot : mv) {
    if (!proxies_[slot].alive) continue;
    if (n < cap) { out[2 * n] = proxies_[slot].fixture; out[2 * n + 1] = proxies_[slot].child; }
    ++n;
  }
  return n;
}

int World::writeMoves(const int32_t* in, int n) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  pendingMoves_.clear();
  for (int i = 0; i < n; ++i) {
    int f = in[2 * i], c = in[2 * i + 1];
    if (f < 0 || f >= (int)fixtures_.size() || !fixtures_[f].alive || c < 0 || c >= (int)fixtures_[f].proxies.size()) return DBX_E_INVALID;
    pendingMoves_.push_back(fixtures_[f].proxies[c]);
  }
  newFixture_ = false;
  return n;
}

int World::readPairs(int32_t* out, int cap) {
  if (!dw_.hdr) return 0;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  int n = 0;
  CUDA_OR_FAIL(cudaMemcpy(&n, (char*)hdr_.p + offsetof(Header, nPairs), 4, cudaMemcpyDeviceToHost), "read nPairs");
  n = std::min(n, (int)pairs.cap);
  std::vector<int2> pr(std::max(n, 1));
  if (n) CUDA_OR_FAIL(cudaMemcpy(pr.data(), pairs.p, (size_t)n * 8, cudaMemcpyDeviceToHost), "read pairs");
  // reference order: sorted by (proxyIdA, proxyIdB) (b2broadphase.d:166, 312-325)
  std::sort(pr.begin(), pr.begin() + n, [&](const int2& a, const int2& b) {
    int ka = proxies_[a.x].key, kb = proxies_[b.x].key;
    if (ka != kb) return ka < kb;
    return proxies_[a.y].key < proxies_[b.y].key; });
  for (int i = 0; i < n && i < cap; ++i) {
    out[4 * i] = proxies_[pr[i].x].fixture; out[4 * i + 1] = proxies_[pr[i].x].child;
    out[4 * i + 2] = proxies_[pr[i].y].fixture; out[4 * i + 3] = proxies_[pr[i].y].child;
  }
  return n;
}

int World::readContacts(dbx_contact_rec* out, int cap) {
  if (!dw_.hdr || bodiesSynced_ == 0) return 0;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  int high = 0;
  CUDA_OR_FAIL(cudaMemcpy(&high, (char*)hdr_.p + offsetof(Header, cHigh), 4, cudaMemcpyDeviceToHost), "read cHigh");
  if (high <= 0) return 0;
  const size_t n = (size_t)high;
  std::vector<unsigned long long> key(n); std::vector<int4> ids(n), fix(n); std::vector<uint32_t> fl(n);
  std::vector<float4> m0(n), m1(n), imp(n), mat(n); std::vector<uint4> mk(n); std::vector<int> tc(n);
  cudaMemcpy(key.data(), c_key.p, n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(ids.data(), c_ids.p, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(fix.data(), c_fix.p, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(fl.data(), c_flags.p, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(m0.data(), c_m0.p, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(m1.data(), c_m1.p, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(imp.data(), c_imp.p, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(mat.data(), c_mat.p, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(mk.data(), c_mk.p, n * 16, cudaMemcpyDeviceToHost);
  CUDA_OR_FAIL(cudaMemcpy(tc.data(), c_toiCount.p, n * 4, cudaMemcpyDeviceToHost), "read contacts");
  std::vector<int> order;
  for (size_t i = 0; i < n; ++i) if (fl[i] & CF_ALIVE) order.push_back((int)i);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });   // deterministic: by reference pair key
  int cnt = 0;
  lastReadSlots_ = order;
  for (int i : order) {
    if (cnt < cap) {
      dbx_contact_rec& o = out[cnt];
      std::memset(&o, 0, sizeof(o));
      o.fixtureA = fix[i].x; o.fixtureB = fix[i].y;
      o.childA = proxies_[ids[i].x % (int)proxies_.size()].child; o.childB = proxies_[ids[i].y % (int)proxies_.size()].child;
      o.flags = fl[i] & 0x3F;
      o.manifold.localNormal = dbx_vec2{m0[i].x, m0[i].y}; o.manifold.localPoint = dbx_vec2{m0[i].z, m0[i].w};
      o.manifold.points[0].localPoint = dbx_vec2{m1[i].x, m1[i].y}; o.manifold.points[1].localPoint = dbx_vec2{m1[i].z, m1[i].w};
      o.manifold.points[0].normalImpulse = imp[i].x; o.manifold.points[0].tangentImpulse = imp[i].y;
      o.manifold.points[1].normalImpulse = imp[i].z; o.manifold.points[1].tangentImpulse = imp[i].w;
      o.manifold.points[0].key = mk[i].x; o.manifold.points[1].key = mk[i].y;
      o.manifold.type = (int)mk[i].z; o.manifold.pointCount = (int)mk[i].w;
      o.friction = mat[i].x; o.restitution = mat[i].y; o.tangentSpeed = mat[i].z; o.toi = mat[i].w; o.toiCount = tc[i];
    }
    ++cnt;
  }
  return cnt;
}

// the persistent colours of the contacts (solver schedule): part of a snapshot, so that a restored world runs the same
// Gauss-Seidel order as the one it was taken from
int World::readContactColours(int32_t* out, int cap) {
  const int n = (int)lastReadSlots_.size();
  if (n == 0 || !out) return n;
  int high = 0;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(&high, (char*)hdr_.p + offsetof(Header, cHigh), 4, cudaMemcpyDeviceToHost), "read cHigh");
  std::vector<int> col((size_t)std::max(high, 1));
  CUDA_OR_FAIL(cudaMemcpy(col.data(), c_colour.p, (size_t)high * 4, cudaMemcpyDeviceToHost), "read colours");
  for (int k = 0; k < n && k < cap; ++k) out[k] = lastReadSlots_[k] < high ? col[lastReadSlots_[k]] : -1;
  return n;
}
int World::writeContactColours(const int32_t* in, int n) {
  if (n <= 0) return 0;
  if ((size_t)n > c_colour.cap) return DBX_E_INVALID;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(c_colour.p, in, (size_t)n * 4, cudaMemcpyHostToDevice), "write colours");   // writeContacts put record k in slot k
  return n;
}
int World::writeContacts(const dbx_contact_rec* in, int n) {
  if (replicated_) { set_last_error("world is replicated: topology and per-body mutation are frozen"); return DBX_E_UNSUPPORTED; }
  int rc = push(); if (rc < 0) return rc;
  if (!dw_.hdr) return DBX_E_INVALID;
  if ((size_t)n > c_key.cap) { set_last_error("contact capacity"); return DBX_E_CAPACITY; }
  const size_t m = (size_t)std::max(n, 1);
  std::vector<unsigned long long> key(m); std::vector<int4> ids(m), fix(m); std::vector<uint32_t> fl(m);
  std::vector<float4> m0(m), m1(m), imp(m), mat(m); std::vector<uint4> mk(m); std::vector<int> tc(m), col(m, -1);
  for (int i = 0; i < n; ++i) {
    const dbx_contact_rec& r = in[i];
    if (r.fixtureA < 0 || r.fixtureB < 0 || r.fixtureA >= (int)fixtures_.size() || r.fixtureB >= (int)fixtures_.size()) return DBX_E_INVALID;
    const HFixture& fa = fixtures_[r.fixtureA]; const HFixture& fb = fixtures_[r.fixtureB];
    if (!fa.alive || !fb.alive || r.childA >= (int)fa.proxies.size() || r.childB >= (int)fb.proxies.size()) return DBX_E_INVALID;
    const int pa = fa.proxies[r.childA], pb = fb.proxies[r.childB];
    const unsigned ka = (unsigned)proxies_[pa].key, kb = (unsigned)proxies_[pb].key;
    key[i] = ((unsigned long long)std::min(ka, kb) << 32) | std::max(ka, kb);
    ids[i] = make_int4(pa, pb, fa.body, fb.body);
    fix[i] = make_int4(r.fixtureA, r.fixtureB, proxies_[pa].shape, proxies_[pb].shape);
    const bool sensor = fa.def.isSensor || fb.def.isSensor;
    fl[i] = (r.flags & 0x3F) | CF_ALIVE | (sensor ? CF_SENSOR : 0);
    m0[i] = f4(r.manifold.localNormal.x, r.manifold.localNormal.y, r.manifold.localPoint.x, r.manifold.localPoint.y);
    m1[i] = f4(r.manifold.points[0].localPoint.x, r.manifold.points[0].localPoint.y, r.manifold.points[1].localPoint.x, r.manifold.points[1].localPoint.y);
    imp[i] = f4(r.manifold.points[0].normalImpulse, r.manifold.points[0].tangentImpulse, r.manifold.points[1].normalImpulse, r.manifold.points[1].tangentImpulse);
    mk[i] = make_uint4(r.manifold.points[0].key, r.manifold.points[1].key, (uint32_t)r.manifold.type, (uint32_t)r.manifold.pointCount);
    mat[i] = f4(r.friction, r.restitution, r.tangentSpeed, r.toi);
    tc[i] = r.toiCount;
  }
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemset(c_flags.p, 0, c_flags.cap * 4), "clear contacts");
  if (n) {
    cudaMemcpy(c_key.p, key.data(), (size_t)n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(c_ids.p, ids.data(), (size_t)n * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(c_fix.p, fix.data(), (size_t)n * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(c_flags.p, fl.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(c_m0.p, m0.data(), (size_t)n * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(c_m1.p, m1.data(), (size_t)n * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(c_imp.p, imp.data(), (size_t)n * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(c_mat.p, mat.data(), (size_t)n * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(c_mk.p, mk.data(), (size_t)n * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(c_toiCount.p, tc.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
    CUDA_OR_FAIL(cudaMemcpy(c_colour.p, col.data(), (size_t)n * 4, cudaMemcpyHostToDevice), "write contacts");
  }
  CUDA_OR_FAIL(launch_insert_contacts(dw_, L_, n), "insert contacts");
  return checkDeviceError(true) < 0 ? DBX_E_CAPACITY : n;
}

// Test hook: replace the colouring of the NEXT step by a caller-supplied level per contact (index = position in the last
// dbx_world_read_contacts result).  With levels derived from the reference's sequential order the coloured solve is a
// topological re-ordering of that order, i.e. arithmetically the same Gauss-Seidel sweep.
int World::setContactLevels(const int32_t* levels, int n) {
  if (n == 0) { overrideLevels_ = false; return 0; }
  if (n != (int)lastReadSlots_.size()) { set_last_error("levels must match the last read_contacts result"); return DBX_E_INVALID; }
  int rc = push(); if (rc < 0) return rc;
  std::vector<int> col(c_colour.cap, -1);
  for (int i = 0; i < n; ++i) {
    if (levels[i] >= kMaxColours) { set_last_error("too many levels"); return DBX_E_CAPACITY; }
    col[lastReadSlots_[i]] = levels[i];
  }
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(c_colour.p, col.data(), col.size() * 4, cudaMemcpyHostToDevice), "levels up");
  overrideLevels_ = true;
  return 0;
}

// Test hook: the schedule the last step ran (see include/dbox_b200.h), as ONE rank space over joints and contacts.
// k_solve: phase = colour, joints of a colour before its contacts (they never share a dynamic body).  k_solve_tiles: class
// (local / boundary / global) above the colour.  Contact colours and bins are persistent device state, joint colours live on
// the host.
int World::readSolveOrder(int32_t* contactRank, int capC, int32_t* jointRank, int capJ, int32_t* info4) {
  const int n = (int)lastReadSlots_.size();
  if (!lastTiled_ && !lastUnified_ && !jointAt_.empty()) {
    set_last_error("read_solve_order: joints solved as phases of their own (level override / DBX_DEBUG 32) have no merged order"); return DBX_E_UNSUPPORTED;
  }
  std::vector<int> col((size_t)std::max(n, 1));
  if (n > 0) { const int rc = readContactColours(col.data(), n); if (rc < 0) return rc; }
  const int P = dw_.nTiles;
  auto classOfBin = [&](int bin) { return bin >= 2 * P * kTileColours ? 2 : bin / (P * kTileColours); };
  std::vector<int> cls((size_t)std::max(n, 1), 0);
  if (lastTiled_ && n > 0) {
    int high = 0;
    CUDA_OR_FAIL(cudaMemcpy(&high, (char*)hdr_.p + offsetof(Header, cHigh), 4, cudaMemcpyDeviceToHost), "read cHigh");
    std::vector<int> key((size_t)std::max(high, 1)), tcol((size_t)std::max(high, 1));
    CUDA_OR_FAIL(cudaMemcpy(key.data(), c_tkey_.p, (size_t)high * 4, cudaMemcpyDeviceToHost), "read tile keys");
    CUDA_OR_FAIL(cudaMemcpy(tcol.data(), c_tcol_.p, (size_t)high * 4, cudaMemcpyDeviceToHost), "read tile colours");
    for (int k = 0; k < n; ++k) if (lastReadSlots_[k] < high && col[k] >= 0) {
      cls[k] = classOfBin(key[lastReadSlots_[k]]);
      if (cls[k] == 1 && tcol[lastReadSlots_[k]] >= 0) col[k] = tcol[lastReadSlots_[k]];     // boundary rows run in their tile's own colouring
    }
  }
  for (int k = 0; k < n && k < capC; ++k) contactRank[k] = col[k] < 0 ? -1 : ((((cls[k] << 12) | col[k]) << 1) | 1);
  if (capJ > 0 && jointRank) {
    std::vector<int> jkey, jtcol;
    if (lastTiled_ && !jointAt_.empty()) {
      jkey.resize(jointAt_.size()); jtcol.resize(jointAt_.size());
      CUDA_OR_FAIL(cudaMemcpy(jkey.data(), j_tkey_.p, jkey.size() * 4, cudaMemcpyDeviceToHost), "read tile joint keys");
      CUDA_OR_FAIL(cudaMemcpy(jtcol.data(), j_tcol_.p, jtcol.size() * 4, cudaMemcpyDeviceToHost), "read tile joint colours");
    }
    for (int j = 0; j < (int)joints_.size() && j < capJ; ++j) {
      jointRank[j] = -1;
      if (!joints_[j].alive) continue;
      int c = 0, colour = joints_[j].colour;
      if (lastTiled_) {
        const int bin = jkey[jointPos_[j]]; if (bin < 0) continue;
        c = classOfBin(bin);
        if (c == 1 && jtcol[jointPos_[j]] >= 0) colour = jtcol[jointPos_[j]];
      }
      jointRank[j] = ((c << 12) | colour) << 1;
    }
  }
  if (info4) {
    int nc = 0;
    if (dw_.hdr) CUDA_OR_FAIL(cudaMemcpy(&nc, (char*)hdr_.p + offsetof(Header, nColours), 4, cudaMemcpyDeviceToHost), "read nColours");
    info4[0] = (lastTiled_ || lastUnified_) ? 1 : 0;     // position passes walk the order backwards
    info4[1] = nc; info4[2] = jointAt_.empty() ? 0 : nJointColours_; info4[3] = lastTiled_ ? P : 0;
  }
  return n;
}

// Test hook ("colour-validity checker"): number of pairs of solver contacts that share a dynamic body
// AND a colour after the last step, i.e. would have raced in the coloured Gauss-Seidel sweep.  Must be 0.
int World::colourConflicts() {
  if (!dw_.hdr || bodiesSynced_ == 0) return 0;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  int high = 0;
  CUDA_OR_FAIL(cudaMemcpy(&high, (char*)hdr_.p + offsetof(Header, cHigh), 4, cudaMemcpyDeviceToHost), "read cHigh");
  if (high <= 0) return 0;
  const size_t n = (size_t)high;
  const size_t nb = replicated_ ? bodies_.size() * (size_t)nWorlds_ : bodiesSynced_;
  std::vector<int4> ids(n); std::vector<uint32_t> fl(n), bfl(nb); std::vector<int> col(n);
  cudaMemcpy(ids.data(), c_ids.p, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(fl.data(), c_flags.p, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(col.data(), c_colour.p, n * 4, cudaMemcpyDeviceToHost);
  CUDA_OR_FAIL(cudaMemcpy(bfl.data(), b_flags.p, nb * 4, cudaMemcpyDeviceToHost), "read flags");
  std::unordered_map<unsigned long long, int> seen;
  int conflicts = 0;
  // unified phases: a joint occupies its colour on its dynamic bodies, and their contacts must sit above it
  std::vector<int> maxJoint(nb, -1);
  if (!replicated_) for (const HJoint& hj : joints_) {
    if (!hj.alive) continue;
    const int bs[4] = {hj.def.bodyA, hj.def.bodyB, hj.bodyC, hj.bodyD};
    for (int b : bs) if (b >= 0 && (size_t)b < nb && body_type(bfl[b]) == BODY_DYNAMIC) maxJoint[b] = std::max(maxJoint[b], hj.colour);
  }
  for (size_t i = 0; i < n; ++i) {
    if ((fl[i] & (CF_ALIVE | CF_SOLVE)) != (CF_ALIVE | CF_SOLVE)) continue;
    if (col[i] < 0) { ++conflicts; continue; }
    const int bs[2] = {ids[i].z, ids[i].w};
    for (int b : bs) {
      if ((size_t)b >= nb || body_type(bfl[b]) != BODY_DYNAMIC) continue;
      unsigned long long k = ((unsigned long long)(unsigned)b << 32) | (unsigned)col[i];
      if (++seen[k] > 1) ++conflicts;
      if (col[i] <= maxJoint[b]) ++conflicts;
    }
  }
  return conflicts;
}

// every 64 steps: re-pack the contact slots in pair-key order and rebuild the pair hash without tombstones (one host
// round trip to learn the slot count; amortised to a few microseconds per step)
int World::compactContacts() {
  CUDA_OR_FAIL(stage_count(dw_, L_), "count");
  Header h;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(&h, hdr_.p, sizeof(Header), cudaMemcpyDeviceToHost), "read header");
  if (h.cHigh <= 0) return 0;
  const size_t n = (size_t)h.cHigh;
  CUDA_OR_FAIL(cmpKeyA_.reserve(n, false, stream_), "cmp"); CUDA_OR_FAIL(cmpKeyB_.reserve(n, false, stream_), "cmp");
  CUDA_OR_FAIL(cmpValA_.reserve(n, false, stream_), "cmp"); CUDA_OR_FAIL(cmpValB_.reserve(n, false, stream_), "cmp");
  size_t need = cub_temp_bytes((int)n);
  if (need > cubTemp.cap) { CUDA_OR_FAIL(cubTemp.reserve(need, false, stream_), "cubTemp"); L_.cubTemp = cubTemp.p; L_.cubTempBytes = cubTemp.cap; }
  // s_v0 (16 B per solver slot, sCap >= cCap) is free between steps and serves as the gather scratch
  CUDA_OR_FAIL(stage_compact_contacts(dw_, L_, h.cHigh, h.nContacts, s_v0.p, cmpKeyA_.p, cmpKeyB_.p, cmpValA_.p, cmpValB_.p), "compact");
  return 0;
}

int World::phaseTimes(unsigned long long* out, int cap) {
  const int kCap = 4096;
  if (!phaseBuf_.p) {
    CUDA_OR_FAIL(phaseBuf_.reserve(kCap, false, stream_), "phase buffer");
    dw_.phaseTimes = phaseBuf_.p; dw_.phaseCap = kCap;
    return 0;   // enabled; stamps are available after the next step
  }
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  const int n = std::min(cap, kCap);
  CUDA_OR_FAIL(cudaMemcpy(out, phaseBuf_.p, (size_t)n * 8, cudaMemcpyDeviceToHost), "phase read");
  CUDA_OR_FAIL(cudaMemset(phaseBuf_.p, 0, (size_t)kCap * 8), "phase clear");
  return n;
}

// Batched independent worlds (BASELINE config 5): the world's current content becomes replica 0 of `copies` disjoint
// replicas living in the SAME device arrays (replica r owns bodies [r*nB, (r+1)*nB) etc.).  Replicas never interact: the
// replica index is part of the Morton key and of the pair test, islands cannot span replicas, and colouring priorities
// use replica-local pair keys, so all replicas evolve bit-identically until something perturbs one of them.
int World::replicate(int copies) {
  if (copies < 1) return DBX_E_INVALID;
  if (replicated_ || stepCount_ > 0) { set_last_error("replicate: call once, before the first step"); return DBX_E_INVALID; }
  if (copies == 1) return 0;
  int rc = push(); if (rc < 0) return rc;
  const int nB = (int)bodies_.size(), nF = (int)fixtures_.size(), nP = (int)proxies_.size(), nJ = (int)jointAt_.size();
  if ((long long)nB * copies > 0x3fffffffLL || (long long)keyFresh_ * copies > 0x7fffffffLL) { set_last_error("replicate: too many replicas"); return DBX_E_CAPACITY; }
  int nMoved = 0;
  CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
  CUDA_OR_FAIL(cudaMemcpy(&nMoved, (char*)hdr_.p + offsetof(Header, nMoved), 4, cudaMemcpyDeviceToHost), "read nMoved");
  nWorlds_ = copies;
  bool rehash = false;
  rc = reserveDevice(rehash); if (rc < 0) return rc;
  refreshView();
  if (rehash) CUDA_OR_FAIL(stage_rebuild_hash(dw_, L_), "rehash");
  keyStride_ = keyFresh_;
  dw_.keyStride = keyStride_;
  CUDA_OR_FAIL(launch_replicate(dw_, L_, nB, nF, nP, nMoved, keyStride_, copies), "replicate");
  // joints: colour-major layout across replicas so that every colour stays one contiguous range
  if (nJ > 0) {
    std::vector<int> off(kMaxJointColours + 1, 0);
    { int c = 0; for (int k = 0; k < nJ; ++k) { while (c < joints_[jointAt_[k]].colour) off[++c] = k; } while (c < kMaxJointColours) off[++c] = nJ; }
    const size_t nD = (size_t)nJ * copies;
    std::vector<int4> ids(nD), id2(nD); std::vector<float4> anc(nD), p0(nD), p1(nD), p2(nD), imp(nD); std::vector<int> lim(nD);
    int newOff[kMaxJointColours + 1];
    for (int c = 0; c <= kMaxJointColours; ++c) newOff[c] = off[c] * copies;
    for (int c = 0; c < kMaxJointColours; ++c) {
      const int cnt = off[c + 1] - off[c];
      for (int r = 0; r < copies; ++r) for (int k = 0; k < cnt; ++k) {
        const HJoint& j = joints_[jointAt_[off[c] + k]];
        const size_t d = (size_t)newOff[c] + (size_t)r * cnt + k;
        const dbx_joint_def& jd = j.def;
        ids[d] = make_int4(jd.type, jd.bodyA + r * nB, jd.bodyB + r * nB, (jd.collideConnected ? 1 : 0) | (jd.enableLimit ? 2 : 0) | (jd.enableMotor ? 4 : 0) | 8);
        anc[d] = make_float4(jd.localAnchorA.x, jd.localAnchorA.y, jd.localAnchorB.x, jd.localAnchorB.y);
        p0[d] = jointParams(jd, 0);
        p1[d] = jointParams(jd, 1);
        p2[d] = jointParams(jd, 2);
        id2[d] = make_int4(j.bodyC >= 0 ? j.bodyC + r * nB : -1, j.bodyD >= 0 ? j.bodyD + r * nB : -1, j.typeA, j.typeB);
        imp[d] = make_float4(j.imp[0], j.imp[1], j.imp[2], j.imp[3]);
        lim[d] = j.limit;
      }
    }
    CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
    cudaMemcpy(j_ids.p, ids.data(), nD * 16, cudaMemcpyHostToDevice); cudaMemcpy(j_anchor.p, anc.data(), nD * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(j_p0.p, p0.data(), nD * 16, cudaMemcpyHostToDevice); cudaMemcpy(j_p1.p, p1.data(), nD * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(j_p2.p, p2.data(), nD * 16, cudaMemcpyHostToDevice); cudaMemcpy(j_ids2.p, id2.data(), nD * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(j_imp.p, imp.data(), nD * 16, cudaMemcpyHostToDevice);
    CUDA_OR_FAIL(cudaMemcpy(j_limit.p, lim.data(), nD * 4, cudaMemcpyHostToDevice), "joints up");
    CUDA_OR_FAIL(cudaMemcpy((char*)hdr_.p + offsetof(Header, jointColourOff), newOff, sizeof(newOff), cudaMemcpyHostToDevice), "joff up");
    // joints that forbid collisions between their bodies, per replica (stays sorted: body ids grow with the replica index)
    if (nJointPairs_ > 0) {
      std::vector<unsigned long long> base(nJointPairs_), all((size_t)nJointPairs_ * copies);
      CUDA_OR_FAIL(cudaMemcpy(base.data(), jp_keys.p, (size_t)nJointPairs_ * 8, cudaMemcpyDeviceToHost), "jp down");
      for (int r = 0; r < copies; ++r) for (int k = 0; k < nJointPairs_; ++k)
        all[(size_t)r * nJointPairs_ + k] = base[k] + (((unsigned long long)(unsigned)(r * nB)) << 32) + (unsigned)(r * nB);
      CUDA_OR_FAIL(jp_keys.reserve(all.size(), false, stream_), "jp_keys");
      CUDA_OR_FAIL(cudaStreamSynchronize(stream_), "sync");
      CUDA_OR_FAIL(cudaMemcpy(jp_keys.p, all.data(), all.size() * 8, cudaMemcpyHostToDevice), "jp up");
      nJointPairs_ *= copies;
      { int rb = uploadJointBits(all); if (rb < 0) return rb; }
    }
    { int rm = uploadJointMasks(); if (rm < 0) return rm; }
    jointBlocks_ = (int)std::min<size_t>(((size_t)jointBlocks_ * L_.coopThreads * copies + L_.coopThreads - 1) / L_.coopThreads, (size_t)L_.coopBlocks / 4);
  }
  replicated_ = true;
  treeValid_ = false;
  refreshView();
  return checkDeviceError(true);
}

}  // namespace dbx
