// dbx_capi.cu — the extern "C" boundary declared in include/dbox_b200.h.  Plain pointers and sizes only.
#include <dlfcn.h>
#include <cstring>
#include <new>
#include "dbx_world.h"

using namespace dbx;

struct dbx_world { World w; dbx_world(float gx, float gy, int dev, const dbx_caps* caps) : w(gx, gy, dev, caps) {} };

static_assert(sizeof(dbx_body_def) == 72 && sizeof(dbx_shape) == 240 && sizeof(dbx_fixture_def) == 32 && sizeof(dbx_joint_def) == 176, "ABI layout");
static_assert(sizeof(dbx_body_state) == 116 && sizeof(dbx_manifold) == 64 && sizeof(dbx_contact_rec) == 104 && sizeof(dbx_proxy_rec) == 44 && sizeof(dbx_contact_event) == 36 && sizeof(dbx_contact_patch) == 36, "ABI layout");

#define W_OR_INVALID(w) do { if (!(w) || !(w)->w.ok()) return DBX_E_INVALID; } while (0)

extern "C" {

int32_t dbx_abi_version(void) { return DBX_ABI_VERSION; }
const char* dbx_last_error(void) { return get_last_error(); }
int32_t dbx_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }

// ---- defaults
void dbx_default_body_def(dbx_body_def* d) {
  std::memset(d, 0, sizeof(*d));
  d->type = DBX_STATIC_BODY; d->allowSleep = 1; d->awake = 1; d->active = 1; d->gravityScale = 1.0f;
}
void dbx_default_fixture_def(dbx_fixture_def* d) {
  std::memset(d, 0, sizeof(*d));
  d->friction = 0.2f; d->categoryBits = 0x0001; d->maskBits = 0xFFFF;
}
void dbx_default_joint_def(dbx_joint_def* d, int32_t type) {
  std::memset(d, 0, sizeof(*d));
  d->type = type;
  // the defaults of the reference's def structs
  if (type == DBX_JOINT_DISTANCE) d->length = 1.0f;                                              // b2distancejoint.d:43-52
  if (type == DBX_JOINT_PRISMATIC || type == DBX_JOINT_WHEEL) d->localAxisA = dbx_vec2{1.0f, 0.0f}; // b2prismaticjoint.d:46, b2wheeljoint.d:46
  if (type == DBX_JOINT_WHEEL) { d->frequencyHz = 2.0f; d->dampingRatio = 0.7f; }                // b2wheeljoint.d:50-52
  if (type == DBX_JOINT_MOUSE) { d->frequencyHz = 5.0f; d->dampingRatio = 0.7f; }                // b2mousejoint.d:43-46
  if (type == DBX_JOINT_MOTOR) { d->maxForce = 1.0f; d->maxTorque = 1.0f; d->correctionFactor = 0.3f; }   // b2motorjoint.d:43-49
  if (type == DBX_JOINT_ROPE) { d->localAnchorA = dbx_vec2{-1.0f, 0.0f}; d->localAnchorB = dbx_vec2{1.0f, 0.0f}; }   // b2ropejoint.d:47-50
  if (type == DBX_JOINT_PULLEY) {                                                                // b2pulleyjoint.d:47-57
    d->groundAnchorA = dbx_vec2{-1.0f, 1.0f}; d->groundAnchorB = dbx_vec2{1.0f, 1.0f};
    d->localAnchorA = dbx_vec2{-1.0f, 0.0f}; d->localAnchorB = dbx_vec2{1.0f, 0.0f}; d->ratio = 1.0f; d->collideConnected = 1;
  }
}

// ---- shape helpers (setup time, host): same arithmetic as the reference's shape classes
void dbx_shape_set_circle(dbx_shape* s, float px, float py, float radius) {
  std::memset(s, 0, sizeof(*s));
  s->type = DBX_SHAPE_CIRCLE; s->radius = radius; s->p.x = px; s->p.y = py;
}
void dbx_shape_set_edge(dbx_shape* s, dbx_vec2 v1, dbx_vec2 v2) {
  std::memset(s, 0, sizeof(*s));
  s->type = DBX_SHAPE_EDGE; s->radius = kPolygonRadius; s->v1 = v1; s->v2 = v2;
}
void dbx_shape_set_box(dbx_shape* s, float hx, float hy) {
  std::memset(s, 0, sizeof(*s));
  s->type = DBX_SHAPE_POLYGON; s->radius = kPolygonRadius; s->count = 4;
  s->vertices[0] = dbx_vec2{-hx, -hy}; s->vertices[1] = dbx_vec2{hx, -hy}; s->vertices[2] = dbx_vec2{hx, hy}; s->vertices[3] = dbx_vec2{-hx, hy};
  s->normals[0] = dbx_vec2{0.0f, -1.0f}; s->normals[1] = dbx_vec2{1.0f, 0.0f}; s->normals[2] = dbx_vec2{0.0f, 1.0f}; s->normals[3] = dbx_vec2{-1.0f, 0.0f};
}
void dbx_shape_set_box_at(dbx_shape* s, float hx, float hy, dbx_vec2 center, float angle) {
  dbx_shape_set_box(s, hx, hy);
  s->centroid = center;
  Xf xf; xf.p = V(center.x, center.y); xf.q = rot_from_angle(angle);
  for (int i = 0; i < 4; ++i) {
    v2 v = mul(xf, V(s->vertices[i].x, s->vertices[i].y)), n = mul(xf.q, V(s->normals[i].x, s->normals[i].y));
    s->vertices[i] = dbx_vec2{v.x, v.y}; s->normals[i] = dbx_vec2{n.x, n.y};
  }
}
// b2PolygonShape.Set: weld near-duplicates, gift-wrap the hull, edge normals, area-weighted centroid
int32_t dbx_shape_set_polygon(dbx_shape* s, const dbx_vec2* pts, int32_t count) {
  if (count < 3) { dbx_shape_set_box(s, 1.0f, 1.0f); return s->count; }
  std::memset(s, 0, sizeof(*s));
  s->type = DBX_SHAPE_POLYGON; s->radius = kPolygonRadius;
  int n = count < kMaxPolygonVertices ? count : kMaxPolygonVertices;
  v2 ps[kMaxPolygonVertices];
  int m = 0;
  for (int i = 0; i < n; ++i) {
    v2 v = V(pts[i].x, pts[i].y);
    bool unique = true;
    for (int j = 0; j < m; ++j) if (dist2(v, ps[j]) < 0.5f * kLinearSlop) { unique = false; break; }
    if (unique) ps[m++] = v;
  }
  n = m;
  if (n < 3) { dbx_shape_set_box(s, 1.0f, 1.0f); return s->count; }
  int i0 = 0; float x0 = ps[0].x;
  for (int i = 1; i < n; ++i) { float x = ps[i].x; if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; } }
  int hull[kMaxPolygonVertices];
  int h = 0, ih = i0;
  for (;;) {
    hull[h] = ih;
    int ie = 0;
    for (int j = 1; j < n; ++j) {
      if (ie == ih) { ie = j; continue; }
      v2 r = ps[ie] - ps[hull[h]], v = ps[j] - ps[hull[h]];
      float c = cross(r, v);
      if (c < 0.0f) ie = j;
      if (c == 0.0f && len2(v) > len2(r)) ie = j;
    }
    ++h; ih = ie;
    if (ie == i0) break;
  }
  if (h < 3) { dbx_shape_set_box(s, 1.0f, 1.0f); return s->count; }
  s->count = h;
  v2 vs[kMaxPolygonVertices];
  for (int i = 0; i < h; ++i) { vs[i] = ps[hull[i]]; s->vertices[i] = dbx_vec2{vs[i].x, vs[i].y}; }
  for (int i = 0; i < h; ++i) {
    v2 e = vs[i + 1 < h ? i + 1 : 0] - vs[i];
    v2 nn = cross(e, 1.0f);
    normalize(nn);
    s->normals[i] = dbx_vec2{nn.x, nn.y};
  }
  v2 c = V(0.0f, 0.0f); float area = 0.0f; const float inv3 = 1.0f / 3.0f;
  for (int i = 0; i < h; ++i) {
    v2 p1 = V(0.0f, 0.0f), p2 = vs[i], p3 = i + 1 < h ? vs[i + 1] : vs[0];
    float D = cross(p2 - p1, p3 - p1);
    float tri = 0.5f * D;
    area += tri;
    c += tri * inv3 * (p1 + p2 + p3);
  }
  c *= 1.0f / area;
  s->centroid = dbx_vec2{c.x, c.y};
  return h;
}
void dbx_shape_set_chain(dbx_shape* s, const dbx_vec2* pts, int32_t n, int32_t loop) {
  std::memset(s, 0, sizeof(*s));
  s->type = DBX_SHAPE_CHAIN; s->radius = kPolygonRadius; s->chainVertices = pts; s->chainCount = n;
  if (loop && n >= 3) { s->prevVertex = pts[n - 2]; s->nextVertex = pts[1]; s->hasPrev = 1; s->hasNext = 1; }
}

// ---- world
dbx_world* dbx_world_create(float gx, float gy, int32_t device, const dbx_caps* caps) {
  dbx_world* w = new (std::nothrow) dbx_world(gx, gy, device, caps);
  if (!w) { set_last_error("out of host memory"); return nullptr; }
  if (!w->w.ok()) { delete w; return nullptr; }   // no device => no world: there is no CPU path
  return w;
}
void dbx_world_destroy(dbx_world* w) { delete w; }
int32_t dbx_world_set_flags(dbx_world* w, uint32_t flags) { W_OR_INVALID(w); return w->w.setFlags(flags); }
uint32_t dbx_world_get_flags(dbx_world* w) { return (w && w->w.ok()) ? w->w.flags() : 0; }
int32_t dbx_world_set_gravity(dbx_world* w, float gx, float gy) { W_OR_INVALID(w); return w->w.setGravity(gx, gy); }

int32_t dbx_body_create(dbx_world* w, const dbx_body_def* def) { W_OR_INVALID(w); if (!def) return DBX_E_INVALID; return w->w.createBody(*def); }
int32_t dbx_body_destroy(dbx_world* w, int32_t body) { W_OR_INVALID(w); return w->w.destroyBody(body); }
int32_t dbx_fixture_create(dbx_world* w, int32_t body, const dbx_fixture_def* def, const dbx_shape* shape) {
  W_OR_INVALID(w); if (!def || !shape) return DBX_E_INVALID; return w->w.createFixture(body, *def, *shape);
}
int32_t dbx_fixture_destroy(dbx_world* w, int32_t fixture) { W_OR_INVALID(w); return w->w.destroyFixture(fixture); }
int32_t dbx_joint_create(dbx_world* w, const dbx_joint_def* def) { W_OR_INVALID(w); if (!def) return DBX_E_INVALID; return w->w.createJoint(*def); }
int32_t dbx_joint_destroy(dbx_world* w, int32_t joint) { W_OR_INVALID(w); return w->w.destroyJoint(joint); }
int32_t dbx_joint_set_target(dbx_world* w, int32_t joint, float x, float y) { W_OR_INVALID(w); return w->w.setJointTarget(joint, x, y); }

int32_t dbx_world_step(dbx_world* w, float dt, int32_t vi, int32_t pi) { W_OR_INVALID(w); return w->w.step(dt, vi, pi, 1); }
int32_t dbx_world_step_n(dbx_world* w, float dt, int32_t vi, int32_t pi, int32_t n) { W_OR_INVALID(w); return w->w.step(dt, vi, pi, n); }
int32_t dbx_world_time_steps(dbx_world* w, float dt, int32_t vi, int32_t pi, int32_t n, int32_t flushL2, float* totalMs, float* stageMs) {
  W_OR_INVALID(w); return w->w.timeSteps(dt, vi, pi, n, flushL2 != 0, totalMs, stageMs);
}
int32_t dbx_world_set_body_states(dbx_world* w, const int32_t* ids, const float* x_y_angle_pad, const float* vx_vy_w_pad, int32_t n) {
  W_OR_INVALID(w); return w->w.setBodyStates(ids, x_y_angle_pad, vx_vy_w_pad, n);
}
int32_t dbx_world_apply_forces(dbx_world* w, const float* fx_fy_torque_pad, int32_t n) { W_OR_INVALID(w); if (!fx_fy_torque_pad && n) return DBX_E_INVALID; return w->w.applyForces(fx_fy_torque_pad, n); }
int32_t dbx_world_read_transforms(dbx_world* w, float* out, int32_t n) { W_OR_INVALID(w); if (!out && n) return DBX_E_INVALID; return w->w.readTransforms(out, n); }
int64_t dbx_world_launch_count(dbx_world* w) { return (w && w->w.ok()) ? (int64_t)w->w.launchCount() : 0; }
int32_t dbx_world_clear_forces(dbx_world* w) { W_OR_INVALID(w); return w->w.clearForces(); }

// ---- body accessors / mutators (dynamics/b2body.d)
int32_t dbx_body_get_state(dbx_world* w, int32_t body, dbx_body_state* out) { W_OR_INVALID(w); if (!out) return DBX_E_INVALID; return w->w.getBody(body, out); }
int32_t dbx_body_set_transform(dbx_world* w, int32_t body, float x, float y, float angle) { W_OR_INVALID(w); return w->w.setTransform(body, x, y, angle); }
int32_t dbx_body_set_linear_velocity(dbx_world* w, int32_t body, float vx, float vy) {
  W_OR_INVALID(w); HBody* b = w->w.mutBodyRow(body); if (!b) return DBX_E_INVALID;
  if (b->st.type == DBX_STATIC_BODY) return 0;
  if (vx * vx + vy * vy > 0.0f) w->w.wake(*b, true);
  b->st.v = dbx_vec2{vx, vy}; return 0;
}
int32_t dbx_body_set_angular_velocity(dbx_world* w, int32_t body, float omega) {
  W_OR_INVALID(w); HBody* b = w->w.mutBodyRow(body); if (!b) return DBX_E_INVALID;
  if (b->st.type == DBX_STATIC_BODY) return 0;
  if (omega * omega > 0.0f) w->w.wake(*b, true);
  b->st.w = omega; return 0;
}
int32_t dbx_body_apply_force(dbx_world* w, int32_t body, float fx, float fy, float px, float py, int32_t wake) {
  W_OR_INVALID(w); HBody* b = w->w.mutBodyRow(body); if (!b) return DBX_E_INVALID;
  if (b->st.type != DBX_DYNAMIC_BODY) return 0;
  if (wake && (b->st.flags & DBX_BODY_AWAKE) == 0) w->w.wake(*b, true);
  if (b->st.flags & DBX_BODY_AWAKE) {
    b->st.force.x += fx; b->st.force.y += fy;
    b->st.torque += cross(V(px, py) - V(b->st.c.x, b->st.c.y), V(fx, fy));
  }
  return 0;
}
int32_t dbx_body_apply_torque(dbx_world* w, int32_t body, float torque, int32_t wake) {
  W_OR_INVALID(w); HBody* b = w->w.mutBodyRow(body); if (!b) return DBX_E_INVALID;
  if (b->st.type != DBX_DYNAMIC_BODY) return 0;
  if (wake && (b->st.flags & DBX_BODY_AWAKE) == 0) w->w.wake(*b, true);
  if (b->st.flags & DBX_BODY_AWAKE) b->st.torque += torque;
  return 0;
}
int32_t dbx_body_apply_linear_impulse(dbx_world* w, int32_t body, float ix, float iy, float px, float py, int32_t wake) {
  W_OR_INVALID(w); HBody* b = w->w.mutBodyRow(body); if (!b) return DBX_E_INVALID;
  if (b->st.type != DBX_DYNAMIC_BODY) return 0;
  if (wake && (b->st.flags & DBX_BODY_AWAKE) == 0) w->w.wake(*b, true);
  if (b->st.flags & DBX_BODY_AWAKE) {
    v2 dv = b->st.invMass * V(ix, iy);
    b->st.v.x += dv.x; b->st.v.y += dv.y;
    b->st.w += b->st.invI * cross(V(px, py) - V(b->st.c.x, b->st.c.y), V(ix, iy));
  }
  return 0;
}
int32_t dbx_body_apply_angular_impulse(dbx_world* w, int32_t body, float impulse, int32_t wake) {
  W_OR_INVALID(w); HBody* b = w->w.mutBodyRow(body); if (!b) return DBX_E_INVALID;
  if (b->st.type != DBX_DYNAMIC_BODY) return 0;
  if (wake && (b->st.flags & DBX_BODY_AWAKE) == 0) w->w.wake(*b, true);
  if (b->st.flags & DBX_BODY_AWAKE) b->st.w += b->st.invI * impulse;
  return 0;
}
int32_t dbx_body_set_awake(dbx_world* w, int32_t body, int32_t flag) {
  W_OR_INVALID(w); HBody* b = w->w.mutBodyRow(body); if (!b) return DBX_E_INVALID; w->w.wake(*b, flag != 0); return 0;
}
int32_t dbx_body_set_bullet(dbx_world* w, int32_t body, int32_t flag) {
  W_OR_INVALID(w); HBody* b = w->w.mutBodyRow(body); if (!b) return DBX_E_INVALID;
  if (flag) b->st.flags |= DBX_BODY_BULLET; else b->st.flags &= ~DBX_BODY_BULLET; return 0;
}
int32_t dbx_body_set_type(dbx_world* w, int32_t body, int32_t type) { W_OR_INVALID(w); return w->w.setBodyType(body, type); }
int32_t dbx_body_set_active(dbx_world* w, int32_t body, int32_t flag) { W_OR_INVALID(w); return w->w.setBodyActive(body, flag != 0); }
int32_t dbx_body_set_mass_data(dbx_world* w, int32_t body, float mass, float cx, float cy, float I) { W_OR_INVALID(w); return w->w.setMassData(body, mass, cx, cy, I); }
int32_t dbx_body_reset_mass_data(dbx_world* w, int32_t body) { W_OR_INVALID(w); return w->w.resetMass(body); }
int32_t dbx_body_set_fixed_rotation(dbx_world* w, int32_t body, int32_t flag) { W_OR_INVALID(w); return w->w.setFixedRotation(body, flag != 0); }
int32_t dbx_body_set_linear_damping(dbx_world* w, int32_t body, float d) { W_OR_INVALID(w); return w->w.setBodyScalars(body, &d, nullptr, nullptr); }
int32_t dbx_body_set_angular_damping(dbx_world* w, int32_t body, float d) { W_OR_INVALID(w); return w->w.setBodyScalars(body, nullptr, &d, nullptr); }
int32_t dbx_body_set_gravity_scale(dbx_world* w, int32_t body, float s) { W_OR_INVALID(w); return w->w.setBodyScalars(body, nullptr, nullptr, &s); }
int32_t dbx_fixture_set_filter(dbx_world* w, int32_t fixture, int32_t categoryBits, int32_t maskBits, int32_t groupIndex) { W_OR_INVALID(w); return w->w.setFixtureFilter(fixture, categoryBits, maskBits, groupIndex); }
int32_t dbx_fixture_set_sensor(dbx_world* w, int32_t fixture, int32_t flag) { W_OR_INVALID(w); return w->w.setFixtureSensor(fixture, flag != 0); }
int32_t dbx_fixture_set_friction(dbx_world* w, int32_t fixture, float v) { W_OR_INVALID(w); return w->w.setFixtureMaterial(fixture, &v, nullptr, nullptr); }
int32_t dbx_fixture_set_restitution(dbx_world* w, int32_t fixture, float v) { W_OR_INVALID(w); return w->w.setFixtureMaterial(fixture, nullptr, &v, nullptr); }
int32_t dbx_fixture_set_density(dbx_world* w, int32_t fixture, float v) { W_OR_INVALID(w); return w->w.setFixtureMaterial(fixture, nullptr, nullptr, &v); }
int32_t dbx_body_set_sleeping_allowed(dbx_world* w, int32_t body, int32_t flag) {
  W_OR_INVALID(w); HBody* b = w->w.mutBodyRow(body); if (!b) return DBX_E_INVALID;
  if (flag) b->st.flags |= DBX_BODY_AUTOSLEEP; else { b->st.flags &= ~DBX_BODY_AUTOSLEEP; w->w.wake(*b, true); }
  return 0;
}

// ---- bulk state
int32_t dbx_world_counts(dbx_world* w, dbx_counts* out) { W_OR_INVALID(w); if (!out) return DBX_E_INVALID; return w->w.counts(out); }
int32_t dbx_world_profile(dbx_world* w, dbx_profile* out) { W_OR_INVALID(w); if (!out) return DBX_E_INVALID; return w->w.profile(out); }
int32_t dbx_world_read_bodies(dbx_world* w, dbx_body_state* out, int32_t cap) { W_OR_INVALID(w); return w->w.readBodies(out, out ? cap : 0); }
int32_t dbx_world_write_bodies(dbx_world* w, const dbx_body_state* in, int32_t n) { W_OR_INVALID(w); if (!in && n) return DBX_E_INVALID; return w->w.writeBodies(in, n); }
int32_t dbx_world_read_contacts(dbx_world* w, dbx_contact_rec* out, int32_t cap) { W_OR_INVALID(w); return w->w.readContacts(out, out ? cap : 0); }
int32_t dbx_world_write_contacts(dbx_world* w, const dbx_contact_rec* in, int32_t n) { W_OR_INVALID(w); if (!in && n) return DBX_E_INVALID; return w->w.writeContacts(in, n); }
int32_t dbx_world_read_proxies(dbx_world* w, dbx_proxy_rec* out, int32_t cap) { W_OR_INVALID(w); return w->w.readProxies(out, out ? cap : 0); }
int32_t dbx_world_write_proxies(dbx_world* w, const dbx_proxy_rec* in, int32_t n) { W_OR_INVALID(w); if (!in && n) return DBX_E_INVALID; return w->w.writeProxies(in, n); }
int32_t dbx_world_read_joints(dbx_world* w, dbx_joint_state* out, int32_t cap) { W_OR_INVALID(w); return w->w.readJoints(out, out ? cap : 0); }
int32_t dbx_world_write_joints(dbx_world* w, const dbx_joint_state* in, int32_t n) { W_OR_INVALID(w); if (!in && n) return DBX_E_INVALID; return w->w.writeJoints(in, n); }
int32_t dbx_world_read_moves(dbx_world* w, int32_t* out, int32_t cap) { W_OR_INVALID(w); return w->w.readMoves(out, out ? cap : 0); }
int32_t dbx_world_write_moves(dbx_world* w, const int32_t* in, int32_t n) { W_OR_INVALID(w); if (!in && n) return DBX_E_INVALID; return w->w.writeMoves(in, n); }
int32_t dbx_world_get_inv_dt0(dbx_world* w, float* out) { W_OR_INVALID(w); if (!out) return DBX_E_INVALID; *out = w->w.inv_dt0; return 0; }
int32_t dbx_world_set_inv_dt0(dbx_world* w, float v) { W_OR_INVALID(w); w->w.inv_dt0 = v; return 0; }

// ---- staged stepping
int32_t dbx_world_stage_find_new_contacts(dbx_world* w) { W_OR_INVALID(w); return w->w.stageFindNewContacts(); }
int32_t dbx_world_stage_collide(dbx_world* w) { W_OR_INVALID(w); return w->w.stageCollide(); }
int32_t dbx_world_read_pairs(dbx_world* w, int32_t* out, int32_t cap) { W_OR_INVALID(w); return w->w.readPairs(out, out ? cap : 0); }
int32_t dbx_world_debug_set_contact_levels(dbx_world* w, const int32_t* levels, int32_t n) { W_OR_INVALID(w); return w->w.setContactLevels(levels, n); }
int32_t dbx_world_debug_read_solve_order(dbx_world* w, int32_t* cc, int32_t capC, int32_t* jc, int32_t capJ, int32_t* info3) { W_OR_INVALID(w); return w->w.readSolveOrder(cc, capC, jc, capJ, info3); }

int32_t dbx_world_debug_colour_conflicts(dbx_world* w) { W_OR_INVALID(w); return w->w.colourConflicts(); }

int32_t dbx_world_debug_phase_times(dbx_world* w, uint64_t* out, int32_t cap) { W_OR_INVALID(w); return w->w.phaseTimes((unsigned long long*)out, cap); }

int32_t dbx_world_debug_header(dbx_world* w, void* out, int32_t bytes) { W_OR_INVALID(w); return w->w.readHeader(out, bytes); }

// ---- batched independent worlds
// final statistics over the ranks: NCCL resolved at run time (no link dependency; the host program owns the communicator)
int32_t dbx_stats_allreduce(void* comm, void* stream, double* sums, int32_t nSums, double* maxima, int32_t nMax) {
  typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  static AllReduceFn allReduce = nullptr;
  if (!comm || nSums < 0 || nMax < 0 || (nSums > 0 && !sums) || (nMax > 0 && !maxima)) { set_last_error("stats_allreduce: bad arguments"); return DBX_E_INVALID; }
  if (!allReduce) {
    void* sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
    if (!sym) { void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); if (lib) sym = dlsym(lib, "ncclAllReduce"); }
    if (!sym) { set_last_error("stats_allreduce: no NCCL in this process (ncclAllReduce not found, libnccl.so.2 not loadable)"); return DBX_E_UNSUPPORTED; }
    allReduce = (AllReduceFn)sym;
  }
  const int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;      // ncclDataType_t / ncclRedOp_t (nccl.h)
  cudaStream_t st = (cudaStream_t)stream;
  double* d = nullptr;
  const size_t n = (size_t)nSums + (size_t)nMax;
  if (n == 0) return 0;
  if (cudaMalloc((void**)&d, n * sizeof(double)) != cudaSuccess) { set_last_error("stats_allreduce: cudaMalloc"); return DBX_E_CUDA; }
  int rc = 0;
  if (nSums) cudaMemcpyAsync(d, sums, (size_t)nSums * 8, cudaMemcpyHostToDevice, st);
  if (nMax) cudaMemcpyAsync(d + nSums, maxima, (size_t)nMax * 8, cudaMemcpyHostToDevice, st);
  if (nSums && allReduce(d, d, (size_t)nSums, kNcclFloat64, kNcclSum, comm, st) != 0) rc = DBX_E_CUDA;
  if (!rc && nMax && allReduce(d + nSums, d + nSums, (size_t)nMax, kNcclFloat64, kNcclMax, comm, st) != 0) rc = DBX_E_CUDA;
  if (!rc) {
    if (nSums) cudaMemcpyAsync(sums, d, (size_t)nSums * 8, cudaMemcpyDeviceToHost, st);
    if (nMax) cudaMemcpyAsync(maxima, d + nSums, (size_t)nMax * 8, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = DBX_E_CUDA;
  } else set_last_error("stats_allreduce: ncclAllReduce failed");
  cudaFree(d);
  return rc;
}
int32_t dbx_world_replicate(dbx_world* w, int32_t copies) { W_OR_INVALID(w); return w->w.replicate(copies); }
int32_t dbx_world_replica_count(dbx_world* w) { W_OR_INVALID(w); return w->w.replicaCount(); }
// ---- snapshot / restore (SURVEY.md 5 "checkpoint / resume"; the reference only has b2World.Dump, dynamics/b2world.d:796-855)
// blob = { magic, version, nb, np, nc, nj, nm, inv_dt0 } + bodies + proxies + contacts + joints + moves + contact colours
namespace { struct SnapHead { uint32_t magic, version; int32_t nb, np, nc, nj, nm; float inv_dt0; }; constexpr uint32_t kSnapMagic = 0x58424431u; }
int64_t dbx_world_export_state(dbx_world* w, void* buf, int64_t cap) {
  W_OR_INVALID(w);
  World& W = w->w;
  W.resetSolverSchedule();      // the tile assignment is re-derived from the state the blob holds, here and in whoever imports it
  SnapHead h{kSnapMagic, 1, 0, 0, 0, 0, 0, 0.0f};
  h.nb = W.readBodies(nullptr, 0); h.np = W.readProxies(nullptr, 0); h.nc = W.readContacts(nullptr, 0); h.nj = W.readJoints(nullptr, 0); h.nm = W.readMoves(nullptr, 0);
  if (h.nb < 0 || h.np < 0 || h.nc < 0 || h.nj < 0 || h.nm < 0) return DBX_E_CUDA;
  h.inv_dt0 = W.inv_dt0;
  const int64_t need = (int64_t)sizeof(SnapHead) + (int64_t)h.nb * sizeof(dbx_body_state) + (int64_t)h.np * sizeof(dbx_proxy_rec) + (int64_t)h.nc * sizeof(dbx_contact_rec) +
                       (int64_t)h.nj * sizeof(dbx_joint_state) + (int64_t)h.nm * 8 + (int64_t)h.nc * 4;
  if (!buf || cap < need) return need;            // size query
  char* p = (char*)buf;
  std::memcpy(p, &h, sizeof(h)); p += sizeof(h);
  if (W.readBodies((dbx_body_state*)p, h.nb) != h.nb) return DBX_E_CUDA; p += (size_t)h.nb * sizeof(dbx_body_state);
  if (W.readProxies((dbx_proxy_rec*)p, h.np) != h.np) return DBX_E_CUDA; p += (size_t)h.np * sizeof(dbx_proxy_rec);
  if (W.readContacts((dbx_contact_rec*)p, h.nc) != h.nc) return DBX_E_CUDA; p += (size_t)h.nc * sizeof(dbx_contact_rec);
  if (W.readJoints((dbx_joint_state*)p, h.nj) != h.nj) return DBX_E_CUDA; p += (size_t)h.nj * sizeof(dbx_joint_state);
  if (W.readMoves((int32_t*)p, h.nm) != h.nm) return DBX_E_CUDA; p += (size_t)h.nm * 8;
  if (W.readContactColours((int32_t*)p, h.nc) != h.nc) return DBX_E_CUDA;      // same order as the contact records above
  return need;
}
int32_t dbx_world_import_state(dbx_world* w, const void* buf, int64_t n) {
  W_OR_INVALID(w);
  World& W = w->w;
  if (!buf || n < (int64_t)sizeof(SnapHead)) return DBX_E_INVALID;
  SnapHead h; std::memcpy(&h, buf, sizeof(h));
  if (h.magic != kSnapMagic || h.version != 1) { set_last_error("import_state: not a dbox_b200 snapshot"); return DBX_E_INVALID; }
  const int64_t need = (int64_t)sizeof(SnapHead) + (int64_t)h.nb * sizeof(dbx_body_state) + (int64_t)h.np * sizeof(dbx_proxy_rec) + (int64_t)h.nc * sizeof(dbx_contact_rec) +
                       (int64_t)h.nj * sizeof(dbx_joint_state) + (int64_t)h.nm * 8 + (int64_t)h.nc * 4;
  // the header is untrusted input: every count must be non-negative, the blob exactly as long as its counts say (anything else
  // is a truncated or foreign buffer), and the counts must be those of THIS world's topology -- the scene is not in the blob, so
  // a snapshot of another scene would otherwise restore half a world
  if (h.nb < 0 || h.np < 0 || h.nc < 0 || h.nj < 0 || h.nm < 0) { set_last_error("import_state: negative record count"); return DBX_E_INVALID; }
  if (n != need) { set_last_error("import_state: blob length does not match its header"); return DBX_E_INVALID; }
  {
    const int nb = W.readBodies(nullptr, 0), np = W.readProxies(nullptr, 0), nj = W.readJoints(nullptr, 0);
    if (nb < 0 || np < 0 || nj < 0) return DBX_E_CUDA;
    if (h.nb != nb || h.np != np || h.nj != nj) {
      set_last_error("import_state: snapshot of a different scene (bodies / proxies / joints do not match this world)"); return DBX_E_INVALID;
    }
    if (h.nm > np) { set_last_error("import_state: more moves than proxies"); return DBX_E_INVALID; }
  }
  const char* p = (const char*)buf + sizeof(SnapHead);
  int rc = W.writeBodies((const dbx_body_state*)p, h.nb); if (rc != h.nb) return rc < 0 ? rc : DBX_E_INVALID; p += (size_t)h.nb * sizeof(dbx_body_state);
  rc = W.writeProxies((const dbx_proxy_rec*)p, h.np); if (rc != h.np) return rc < 0 ? rc : DBX_E_INVALID; p += (size_t)h.np * sizeof(dbx_proxy_rec);
  const dbx_contact_rec* cons = (const dbx_contact_rec*)p; p += (size_t)h.nc * sizeof(dbx_contact_rec);
  if (h.nj > 0) { rc = W.writeJoints((const dbx_joint_state*)p, h.nj); if (rc != h.nj) return rc < 0 ? rc : DBX_E_INVALID; }
  p += (size_t)h.nj * sizeof(dbx_joint_state);
  rc = W.writeContacts(cons, h.nc); if (rc != h.nc) return rc < 0 ? rc : DBX_E_INVALID;
  rc = W.writeMoves((const int32_t*)p, h.nm); if (rc != h.nm) return rc < 0 ? rc : DBX_E_INVALID; p += (size_t)h.nm * 8;
  rc = W.writeContactColours((const int32_t*)p, h.nc); if (rc != h.nc) return rc < 0 ? rc : DBX_E_INVALID;
  W.inv_dt0 = h.inv_dt0;
  W.resetSolverSchedule();
  return 0;
}
int32_t dbx_world_step_begin(dbx_world* w, float dt, int32_t vi, int32_t pi) { W_OR_INVALID(w); return w->w.stepBegin(dt, vi, pi); }
int32_t dbx_world_step_end(dbx_world* w) { W_OR_INVALID(w); return w->w.stepEnd(); }
int32_t dbx_world_patch_contacts(dbx_world* w, const dbx_contact_patch* patches, int32_t n) { W_OR_INVALID(w); return w->w.patchContacts(patches, n); }
int32_t dbx_world_raycast_closest(dbx_world* w, const dbx_ray* rays, int32_t n, dbx_ray_hit* out) { W_OR_INVALID(w); return w->w.rayCastClosest(rays, n, out); }
int32_t dbx_world_set_user_filter(dbx_world* w, int32_t mode) { W_OR_INVALID(w); return w->w.setUserFilter(mode); }
int32_t dbx_world_poll_new_contacts(dbx_world* w, int32_t* out, int32_t cap) { W_OR_INVALID(w); return w->w.pollNewContacts(out, cap); }
int32_t dbx_world_step_async(dbx_world* w, float dt, int32_t vi, int32_t pi) { W_OR_INVALID(w); return w->w.stepAsync(dt, vi, pi); }
int32_t dbx_world_apply_forces_async(dbx_world* w, const float* f, int32_t n) { W_OR_INVALID(w); return w->w.applyForcesAsync(f, n); }
int32_t dbx_world_read_transforms_async(dbx_world* w, float* out, int32_t n) { W_OR_INVALID(w); return w->w.readTransformsAsync(out, n); }
int32_t dbx_world_set_io_format(dbx_world* w, int32_t format) { W_OR_INVALID(w); return w->w.setIoFormat(format); }
int32_t dbx_world_io_wait(dbx_world* w, int32_t ticket) { W_OR_INVALID(w); return w->w.ioWait(ticket); }
int32_t dbx_world_sync(dbx_world* w) { W_OR_INVALID(w); return w->w.sync(); }
int32_t dbx_joint_set_params(dbx_world* w, int32_t joint, const dbx_joint_def* def, uint32_t mask) { W_OR_INVALID(w); if (!def) return DBX_E_INVALID; return w->w.setJointParams(joint, *def, mask); }
int32_t dbx_world_set_motor_speeds(dbx_world* w, const int32_t* joints, const float* speeds, int32_t n) { W_OR_INVALID(w); return w->w.setMotorSpeeds(joints, speeds, n); }
int32_t dbx_world_tree_stats(dbx_world* w, int32_t* height, int32_t* maxBalance, float* quality) { W_OR_INVALID(w); return w->w.treeStats(height, maxBalance, quality); }
int32_t dbx_world_raycast_all(dbx_world* w, const dbx_ray* rays, int32_t n, int32_t capPerRay, int32_t* counts, dbx_ray_hit* hits) { W_OR_INVALID(w); return w->w.rayCastAll(rays, n, capPerRay, counts, hits); }
int32_t dbx_world_test_points(dbx_world* w, const int32_t* fixtures, const dbx_vec2* points, int32_t n, int32_t* inside) { W_OR_INVALID(w); return w->w.testPoints(fixtures, points, n, inside); }
int32_t dbx_world_shift_origin(dbx_world* w, float x, float y) { W_OR_INVALID(w); return w->w.shiftOrigin(x, y); }
int32_t dbx_world_read_world_manifolds(dbx_world* w, dbx_world_manifold* out, int32_t cap) { W_OR_INVALID(w); return w->w.readWorldManifolds(out, cap); }
int32_t dbx_world_enable_post_solve(dbx_world* w, int32_t capacity) { W_OR_INVALID(w); return w->w.enablePostSolve(capacity); }
int32_t dbx_world_read_post_solve(dbx_world* w, dbx_post_solve* out, int32_t cap) { W_OR_INVALID(w); return w->w.readPostSolve(out, cap); }
int32_t dbx_world_query_aabb(dbx_world* w, const dbx_aabb* boxes, int32_t n, int32_t capPerQuery, int32_t* counts, int32_t* fixture_child) {
  W_OR_INVALID(w); return w->w.queryAabb(boxes, n, capPerQuery, counts, fixture_child);
}
int32_t dbx_world_enable_contact_events(dbx_world* w, int32_t capacity) { W_OR_INVALID(w); return w->w.enableContactEvents(capacity); }
int32_t dbx_world_poll_contact_events(dbx_world* w, dbx_contact_event* out, int32_t cap) { W_OR_INVALID(w); return w->w.pollContactEvents(out, cap); }

}  // extern "C"
