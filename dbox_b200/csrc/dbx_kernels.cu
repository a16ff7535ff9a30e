// dbx_kernels.cu — hand-written CUDA kernels (sm_100a) for dbox's per-step world pipeline.
//
// Every kernel is a grid-stride loop over a device-resident count (Header), launched with a fixed grid that is a
// multiple of the SM count, so a whole step is a fixed launch sequence with no host round trip.  The iteration loops
// of the constraint solver run inside ONE persistent cooperative kernel (one CTA per SM) with a global barrier per
// colour.  Nothing here is a dense contraction: the bound is HBM/L2 bandwidth and dependent-launch latency, so the
// levers are coalesced float4 SoA access, L2-resident constraint blocks and as few barriers as the colouring allows.
#include <cub/cub.cuh>
#include "dbx_kernels.cuh"
#include "dbx_solver.cuh"
#include "dbx_colour.cuh"

namespace dbx {

cudaError_t stage_rebuild_hash(const DevWorld& W, const LaunchCfg& L);


// ------------------------------------------------------------------------------------------------ Collide
// b2ContactManager.Collide (dynamics/b2contactmanager.d:251-317) + b2Contact.Update (contacts/b2contact.d:270-356)
// b2ContactListener.BeginContact / EndContact (b2worldcallbacks.d:87-95), deferred: the call sites of the reference
// (b2contact.d:338-346, b2contactmanager.d:60-63) append a record here and the host polls them after the step.
// phase: 1 = Collide, 2 = TOI sub-steps, 3 = destroyed through the API after that step.
enum { EV_BEGIN = 1, EV_END = 2 };
DBX_D void emit_contact_event(const DevWorld& W, int type, int phase, int i, const int4 ids, const int4 fx) {
  if (W.evCap == 0) return;
  const int k = atomicAdd(&W.hdr->nCtEvents, 1);
  if (k >= W.evCap) return;
  const unsigned long long key = W.c_key[i];
  const int childA = W.p_ids[ids.x].y, childB = W.p_ids[ids.y].y;
  W.ev_a[k] = make_int4(type | (phase << 8) | ((W.stepIndex & 0xFFFF) << 16), fx.x, fx.y, (childA & 0xFFFF) | (childB << 16));
  W.ev_b[k] = make_int4(ids.z, ids.w, (int)(unsigned)(key & 0xFFFFFFFFull), (int)(unsigned)(key >> 32));
}
// b2Island.Report (dynamics/b2island.d:438-462), deferred: one record per contact of the island that was just solved, carrying
// the b2ContactImpulse the listener's PostSolve would have been given.  s = solver slot, i = contact slot.
DBX_D void emit_post_solve(const DevWorld& W, int phase, int s, int i) {
  const int k = atomicAdd(&W.hdr->nPostSolve, 1);
  if (k >= W.psCap) return;
  const int4 ids = W.c_ids[i], fx = W.c_fix[i];
  const int childA = W.p_ids[ids.x].y, childB = W.p_ids[ids.y].y;
  const int count = W.s_pc[s] & 0xFF;
  float4 imp = W.s_imp[s];
  if (count < 2) { imp.z = 0.0f; imp.w = 0.0f; }
  W.ps_a[k] = make_int4(fx.x, fx.y, (childA & 0xFFFF) | (childB << 16), count | (phase << 8));
  W.ps_b[k] = imp;
  W.ps_key[k] = W.c_key[i];
}
DBX_D void destroy_contact(const DevWorld& W, int i, uint32_t flags, int bodyA, int bodyB, int pointCount) {
  if (flags & CF_TOUCHING) emit_contact_event(W, EV_END, 1, i, W.c_ids[i], W.c_fix[i]);
  // b2ContactManager.Destroy + b2Contact.Destroy: wake both bodies if the manifold had points and no sensor is involved
  if (pointCount > 0 && !(flags & CF_SENSOR)) { W.b_wake[bodyA] = 1; W.b_wake[bodyB] = 1; }
  hash_remove(W, W.c_key[i]);
  W.c_flags[i] = 0;
  W.c_colour[i] = -1;
  int slot = atomicAdd(&W.hdr->nFree, 1);
  W.c_free[slot] = i;
}

// b2Contact.Update (contacts/b2contact.d:270-356).  `immediateWake`: SetAwake(true) now (TOI loop) instead of deferring to
// the island pass (Collide).  Returns the new flag word (already stored).
DBX_D uint32_t update_contact(const DevWorld& W, int i, uint32_t flags, const int4 ids, const int4 fx, bool immediateWake) {
  uint4 mk = W.c_mk[i];
  flags |= CF_ENABLED;
  if (!immediateWake) flags &= ~CF_PRESOLVE_OFF;      // (Collide: the decision of a step's PreSolve ends with that step)
  const bool wasTouching = (flags & CF_TOUCHING) != 0;
  const Xf xfA = XF(ldcg4(&W.b_xf[ids.z])), xfB = XF(ldcg4(&W.b_xf[ids.w]));
  const DShape* sA = W.shapes + fx.z;
  const DShape* sB = W.shapes + fx.w;
  bool touching;
  if (flags & CF_SENSOR) {
    touching = shapes_overlap(sA, xfA, sB, xfB);
    mk.w = 0;
    W.c_mk[i] = mk;
  } else {
    Manifold m;
    m.type = (int)mk.z; m.localNormal = V(0, 0); m.localPoint = V(0, 0); m.lp[0] = m.lp[1] = V(0, 0); m.key[0] = m.key[1] = 0;
    collide_dispatch(m, sA, xfA, sB, xfB);
    touching = m.pointCount > 0;
    if (touching) {
      const float4 oldImp = W.c_imp[i];
      const int oldCount = (int)mk.w;
      float4 imp = make_float4(0, 0, 0, 0);
      // match new points to old ones by feature key and carry the accumulated impulses (b2contact.d:306-324)
      for (int k = 0; k < m.pointCount; ++k) {
        float ni = 0.0f, ti = 0.0f;
        for (int j = 0; j < oldCount; ++j) {
          uint32_t oldKey = j == 0 ? mk.x : mk.y;
          if (oldKey == m.key[k]) { ni = j == 0 ? oldImp.x : oldImp.z; ti = j == 0 ? oldImp.y : oldImp.w; break; }
        }
        if (k == 0) { imp.x = ni; imp.y = ti; } else { imp.z = ni; imp.w = ti; }
      }
      W.c_m0[i] = make_float4(m.localNormal.x, m.localNormal.y, m.localPoint.x, m.localPoint.y);
      W.c_m1[i] = make_float4(m.lp[0].x, m.lp[0].y, m.lp[1].x, m.lp[1].y);
      W.c_imp[i] = imp;
      W.c_mk[i] = make_uint4(m.key[0], m.key[1], (uint32_t)m.type, (uint32_t)m.pointCount);
    } else {
      mk.w = 0;
      W.c_mk[i] = mk;
    }
    if (touching != wasTouching) {
      if (immediateWake) { wake_body_now(W, ids.z); wake_body_now(W, ids.w); }
      else { W.b_wake[ids.z] = 1; W.b_wake[ids.w] = 1; }
    }
  }
  if (touching != wasTouching) emit_contact_event(W, touching ? EV_BEGIN : EV_END, immediateWake ? 2 : 1, i, ids, fx);
  flags = touching ? (flags | CF_TOUCHING) : (flags & ~CF_TOUCHING);
  // b2ContactListener.PreSolve inside the TOI loop (b2contact.d:348-355 reached from b2world.d:1295,1379): no call-back from the
  // device, so the answer the listener gave after this step's Collide (dbx_world_patch_contacts) is taken to stand
  if (immediateWake && touching && !(flags & CF_SENSOR) && (flags & CF_PRESOLVE_OFF)) flags &= ~CF_ENABLED;
  W.c_flags[i] = flags;
  return flags;
}

__global__ void __launch_bounds__(256, 4) k_collide(const __grid_constant__ DevWorld W) {
  const int n = W.hdr->cHigh;
  GRID_STRIDE(i, n) {
    uint32_t flags = W.c_flags[i];
    if (!(flags & CF_ALIVE)) continue;
    if (flags & CF_FRESH) { flags &= ~CF_FRESH; W.c_flags[i] = flags; }
    const int4 ids = W.c_ids[i];
    const int4 fx = W.c_fix[i];
    const uint32_t flA = W.b_flags[ids.z], flB = W.b_flags[ids.w];
    uint4 mk = W.c_mk[i];
    if (flags & CF_FILTER) {
      if (!body_should_collide(W, ids.w, ids.z, flB, flA) || (!(W.userFilter & 2) && !filter_should_collide(W, fx.x, fx.y))) {
        destroy_contact(W, i, flags, ids.z, ids.w, (int)mk.w);
        continue;
      }
      flags &= ~CF_FILTER;
      W.c_flags[i] = flags;
    }
    bool activeA = (flA & BF_AWAKE) && body_type(flA) != BODY_STATIC;
    bool activeB = (flB & BF_AWAKE) && body_type(flB) != BODY_STATIC;
    // Contacts the TOI pass can ever look at this step (b2world.d:1165-1199: not a sensor, and one side is a bullet or not
    // dynamic) are a few per cent of all contacts: list them here, where every contact and both body flag words are in
    // registers anyway, so that k_toi never has to scan the whole contact pool.
    const bool toiKind = W.continuous && !(flags & CF_SENSOR) &&
                         ((flA & BF_BULLET) || body_type(flA) != BODY_DYNAMIC || (flB & BF_BULLET) || body_type(flB) != BODY_DYNAMIC);
    if (!activeA && !activeB) { if (toiKind) W.c_toiList[atomicAdd(&W.hdr->nToi, 1)] = i; continue; }
    if (!overlap(BX(W.p_fat[ids.x]), BX(W.p_fat[ids.y]))) {
      destroy_contact(W, i, flags, ids.z, ids.w, (int)mk.w);
      continue;
    }
    if (toiKind) W.c_toiList[atomicAdd(&W.hdr->nToi, 1)] = i;
    update_contact(W, i, flags, ids, fx, false);
  }
}

// ------------------------------------------------------------------------------------------------ islands
// Replaces the DFS of b2World.Solve (dynamics/b2world.d:943-1095) with a union-find over constraint edges.
__global__ void __launch_bounds__(256) k_island_init(const __grid_constant__ DevWorld W) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { W.hdr->nSolve = 0; W.hdr->nColours = 0; W.hdr->nIslands = 0; W.hdr->nUncoloured = 0; W.hdr->nUncoloured2 = 0; W.hdr->epoch = W.hdr->epochNext; }
  GRID_STRIDE(b, W.nBodies) {
    W.b_root[b] = b;
    W.b_islAwake[b] = 0;
    W.b_islMinSleep[b] = 0x7f7fffff;  // FLT_MAX bits
    W.b_mask[b] = W.unifiedColours ? W.b_jmask[b] : 0ull;   // colours its joints occupy (and everything below them)
    W.b_ovf[b] = 0;
    if (W.b_wake[b]) { W.b_wake[b] = 0; wake_body_now(W, b); }
  }
  for (int it = 0; it < W.posIters; ++it) { GRID_STRIDE(b, W.nBodies) W.b_posNotOk[it * W.nBodies + b] = 0; }
}

__global__ void __launch_bounds__(256) k_island_union(const __grid_constant__ DevWorld W) {
  const int n = W.hdr->cHigh;
  GRID_STRIDE(i, n) {
    uint32_t flags = W.c_flags[i];
    // contacts qualify iff enabled, touching and non-sensor (b2world.d:1016-1030)
    if ((flags & (CF_ALIVE | CF_TOUCHING | CF_ENABLED | CF_SENSOR)) != (CF_ALIVE | CF_TOUCHING | CF_ENABLED)) continue;
    int4 ids = W.c_ids[i];
    if (body_type(W.b_flags[ids.z]) == BODY_STATIC || body_type(W.b_flags[ids.w]) == BODY_STATIC) continue;  // statics end the search (:998-1003)
    if (W.nWorlds == 1) uf_unite(W.b_root, ids.z, ids.w, true); else uf_unite_small(W.b_root, ids.z, ids.w);
  }
  GRID_STRIDE(j, W.nJoints) {
    int4 ids = W.j_ids[j];
    if (!(ids.w & 8)) continue;
    uint32_t fa = W.b_flags[ids.y], fb = W.b_flags[ids.z];
    if (!(fa & BF_ACTIVE) || !(fb & BF_ACTIVE)) continue;                                  // other body must be active (:1058-1062)
    if (body_type(fa) == BODY_STATIC || body_type(fb) == BODY_STATIC) continue;
    if (W.nWorlds == 1) uf_unite(W.b_root, ids.y, ids.z, true); else uf_unite_small(W.b_root, ids.y, ids.z);
  }
}

__global__ void __launch_bounds__(256) k_island_flatten(const __grid_constant__ DevWorld W) {
  GRID_STRIDE(b, W.nBodies) {
    uint32_t f = W.b_flags[b];
    if (!(f & BF_ALIVE)) continue;
    // read-only walk: a compressing find here could let another thread's late path-halving store overwrite this body's
    // final root with an intermediate ancestor (observed as a body dropping out of its island)
    int r = b;
    for (;;) { int p = __ldcg(&W.b_root[r]); if (p == r) break; r = p; }
    __stcg(&W.b_root[b], r);
    // seeds: awake, active, non-static (b2world.d:963-979)
    if ((f & BF_AWAKE) && (f & BF_ACTIVE) && body_type(f) != BODY_STATIC) W.b_islAwake[r] = 1;
  }
}

// wake every body of an awake island (b2world.d:996), flag island membership, integrate velocities (b2island.d:82-116)
__global__ void __launch_bounds__(256) k_island_wake_integrate(const __grid_constant__ DevWorld W) {
  const float h = W.dt;
  GRID_STRIDE(b, W.nBodies) {
    uint32_t f = W.b_flags[b];
    if (!(f & BF_ALIVE)) continue;
    int type = body_type(f);
    bool in = type != BODY_STATIC && (f & BF_ACTIVE) && W.b_islAwake[W.b_root[b]];
    if (!in) { if (f & BF_ISLAND) W.b_flags[b] = f & ~BF_ISLAND; continue; }
    if (!(f & BF_AWAKE)) { W.b_gs[b].y = 0.0f; }
    f |= BF_AWAKE | BF_ISLAND;
    W.b_flags[b] = f;
    if (W.b_root[b] == b) atomicAdd(&W.hdr->nIslands, 1);
    float4 pos = W.b_pos[b];
    float4 pos0 = W.b_pos0[b];
    pos0.x = pos.x; pos0.y = pos.y; pos0.z = pos.z;   // c0 = c, a0 = a
    W.b_pos0[b] = pos0;
    W.b_xf0[b] = W.b_xf[b];
    if (type == BODY_DYNAMIC) {
      float4 vel = W.b_vel[b];
      const float4 frc = W.b_force[b];
      const float4 ms = W.b_mass[b];
      const float4 lc = W.b_lc[b];
      const float gs = W.b_gs[b].x;
      v2 v = V(vel.x, vel.y);
      float w = vel.z;
      v += h * (gs * V(W.gx, W.gy) + ms.x * V(frc.x, frc.y));
      w += h * ms.y * frc.z;
      v *= 1.0f / (1.0f + h * lc.z);
      w *= 1.0f / (1.0f + h * lc.w);
      W.b_vel[b] = make_float4(v.x, v.y, w, 0.0f);
    }
  }
}

__global__ void __launch_bounds__(256) k_mark_solve(const __grid_constant__ DevWorld W) { mark_solve_body(W); }
__global__ void __launch_bounds__(512) k_colour(const __grid_constant__ DevWorld W) { colour_body(W); }

// ------------------------------------------------------------------------------------------------ colour counting sort
// One radix digit (the colour) over the contact slots: per-CTA histograms, one scan, scatter.  Output: s_contact in
// colour order and colourOff[]; the solver then walks one contiguous range per colour.
// s_hist is colour-major: s_hist[c * kSortBlocks + block]
__global__ void __launch_bounds__(256) k_sort_hist(const __grid_constant__ DevWorld W) {
  __shared__ int hist[kMaxColours];
  for (int c = threadIdx.x; c < kMaxColours; c += blockDim.x) hist[c] = 0;
  __syncthreads();
  const int n = W.hdr->cHigh;
  const int chunk = (n + gridDim.x - 1) / gridDim.x;
  const int beg = blockIdx.x * chunk, end = min(n, beg + chunk);
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
    if (W.c_flags[i] & CF_SOLVE) atomicAdd(&hist[W.c_colour[i]], 1);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kMaxColours; c += blockDim.x) W.s_hist[c * kSortBlocks + blockIdx.x] = hist[c];
}
// one warp per colour: exclusive prefix over the per-CTA counts (coalesced, shuffle scan); total -> colourOff[c] (temporarily)
__global__ void __launch_bounds__(256) k_sort_scan_blocks(const __grid_constant__ DevWorld W) {
  const int lane = threadIdx.x & 31;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= kMaxColours) return;
  int* row = W.s_hist + c * kSortBlocks;
  int carry = 0;
  for (int b0 = 0; b0 < kSortBlocks; b0 += 32) {
    const int b = b0 + lane;
    const int v = b < kSortBlocks ? row[b] : 0;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (b < kSortBlocks) row[b] = carry + x - v;
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) W.hdr->colourOff[c] = carry;
}
// one CTA: exclusive prefix over the colour totals -> colourOff[], nSolve, nColours
__global__ void __launch_bounds__(kMaxColours) k_sort_scan_colours(const __grid_constant__ DevWorld W) {
  __shared__ int warpSum[32];
  __shared__ int lastColour;
  const int c = threadIdx.x, lane = c & 31, wid = c >> 5;
  if (c == 0) lastColour = 0;
  const int v = W.hdr->colourOff[c];
  int x = v;
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) warpSum[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int w = warpSum[lane];
    int y = w;
    for (int o = 1; o < 32; o <<= 1) { int z = __shfl_up_sync(0xffffffffu, y, o); if (lane >= o) y += z; }
    warpSum[lane] = y - w;
  }
  __syncthreads();
  const int excl = warpSum[wid] + x - v;
  if (v > 0) atomicMax(&lastColour, c + 1);
  __shared__ int tot[kMaxColours];
  tot[c] = v;
  __syncthreads();
  if (c == 0) {
    // tail = the longest suffix of colours holding at most kTailContacts constraints in total
    int t = lastColour, acc = 0;
    while (t > 0 && acc + tot[t - 1] <= kTailContacts) { acc += tot[t - 1]; --t; }
    if (lastColour - t < 2 && t > 0) t = lastColour;   // a single small colour gains nothing from the CTA-local path
    W.hdr->tailStart = W.colourOverride ? lastColour : t;
  }
  W.hdr->colourOff[c] = excl;
  if (c == kMaxColours - 1) {
    const int total = excl + v;
    W.hdr->colourOff[kMaxColours] = total;
    W.hdr->nSolve = total;
    if (total > W.sCap) W.hdr->error = E_SOLVER_ROWS;
  }
  if (c == 0) W.hdr->nColours = lastColour;
}
__global__ void __launch_bounds__(256) k_sort_scatter(const __grid_constant__ DevWorld W) {
  __shared__ int cursor[kMaxColours];
  for (int c = threadIdx.x; c < kMaxColours; c += blockDim.x) cursor[c] = W.s_hist[c * kSortBlocks + blockIdx.x] + W.hdr->colourOff[c];
  __syncthreads();
  const int n = W.hdr->cHigh;
  const int chunk = (n + gridDim.x - 1) / gridDim.x;
  const int beg = blockIdx.x * chunk, end = min(n, beg + chunk);
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
    if (W.c_flags[i] & CF_SOLVE) {
      int pos = atomicAdd(&cursor[W.c_colour[i]], 1);
      if (pos < W.sCap) W.s_contact[pos] = i;
    }
  }
}

__global__ void __launch_bounds__(256) k_apply_forces(const __grid_constant__ DevWorld W, const float4* forces, int n) {
  GRID_STRIDE(b, n) {
    const uint32_t f = W.b_flags[b];
    if (!(f & BF_ALIVE) || body_type(f) != BODY_DYNAMIC || !(f & BF_AWAKE)) continue;
    const float4 a = forces[b];
    float4 cur = W.b_force[b];
    cur.x += a.x; cur.y += a.y; cur.z += a.z;
    W.b_force[b] = cur;
  }
}
// DBX_IO_COMPACT: 12-byte records (fx, fy, torque) in, (p.x, p.y, angle) out
__global__ void __launch_bounds__(256) k_apply_forces3(const __grid_constant__ DevWorld W, const float* forces, int n) {
  GRID_STRIDE(b, n) {
    const uint32_t f = W.b_flags[b];
    if (!(f & BF_ALIVE) || body_type(f) != BODY_DYNAMIC || !(f & BF_AWAKE)) continue;
    float4 cur = W.b_force[b];
    cur.x += forces[3 * b]; cur.y += forces[3 * b + 1]; cur.z += forces[3 * b + 2];
    W.b_force[b] = cur;
  }
}
__global__ void __launch_bounds__(256) k_pack_poses(const __grid_constant__ DevWorld W, float* out, int n) {
  GRID_STRIDE(b, n) {
    const float4 xf = W.b_xf[b];
    out[3 * b] = xf.x; out[3 * b + 1] = xf.y; out[3 * b + 2] = W.b_pos[b].z;
  }
}
__global__ void __launch_bounds__(256) k_clear_forces(const __grid_constant__ DevWorld W) {
  GRID_STRIDE(b, W.nBodies) W.b_force[b] = make_float4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------ SynchronizeFixtures
// b2Body.SynchronizeFixtures (b2body.d:1129-1141) -> b2Fixture.Synchronize (b2fixture.d:480-502) -> b2DynamicTree.MoveProxy
// (collision/b2dynamictree.d:140-184): swept tight AABB, fat-box containment test, predictive fattening, move buffer.
DBX_D void sync_proxy(const DevWorld& W, int p, uint32_t pf, int body) {
  {
    const int4 ids = W.p_ids[p];
    const Xf xf1 = XF(ldcg4(&W.b_xf0[body])), xf2 = XF(ldcg4(&W.b_xf[body]));
    const DShape* s = W.shapes + ids.w;
    Box aabb = combine(shape_aabb(s, xf1), shape_aabb(s, xf2));
    W.p_aabb[p] = pack(aabb);
    Box fat = BX(W.p_fat[p]);
    if (contains(fat, aabb)) return;
    v2 displacement = xf2.p - xf1.p;
    Box b = aabb;
    v2 r = V(kAabbExtension, kAabbExtension);
    b.lo = b.lo - r;
    b.hi = b.hi + r;
    v2 d = kAabbMultiplier * displacement;
    if (d.x < 0.0f) b.lo.x += d.x; else b.hi.x += d.x;
    if (d.y < 0.0f) b.lo.y += d.y; else b.hi.y += d.y;
    W.p_fat[p] = pack(b);
    if (!(pf & PF_MOVED)) {
      W.p_flags[p] = pf | PF_MOVED;
      int slot = atomicAdd(&W.hdr->nMoved, 1);
      if (slot < W.moveCap) W.moveList[slot] = p; else W.hdr->error = E_MOVES;
    }
  }
}
__global__ void __launch_bounds__(256) k_sync_fixtures(const __grid_constant__ DevWorld W) {
  GRID_STRIDE(p, W.nProxies) {
    uint32_t pf = W.p_flags[p];
    if (!(pf & PF_ALIVE)) continue;
    const int body = W.p_ids[p].z;
    const uint32_t bf = W.b_flags[body];
    if (!(bf & BF_ISLAND) || body_type(bf) == BODY_STATIC) continue;   // b2world.d:1103-1118
    sync_proxy(W, p, pf, body);
  }
}

// ------------------------------------------------------------------------------------------------ LBVH broadphase
// Replaces the incremental dynamic tree (collision/b2dynamictree.d:566-915) by a linear BVH rebuilt over the same
// persistent fat AABBs: Morton keys -> radix sort -> Karras hierarchy -> bottom-up refit.  The pair SET it reports is
// the one b2BroadPhase.UpdatePairs computes (collision/b2broadphase.d:139-195): for every proxy in the move buffer,
// all proxies whose fat AABB overlaps its fat AABB, each unordered pair once.
DBX_D unsigned f2ord(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
DBX_D float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_bounds_init(const __grid_constant__ DevWorld W) {
  unsigned* b = (unsigned*)W.hdr->bounds;
  b[0] = b[1] = 0xFFFFFFFFu; b[2] = b[3] = 0u;
  W.hdr->nPairs = 0; W.hdr->nFresh = 0;
}
__global__ void __launch_bounds__(256) k_bounds(const __grid_constant__ DevWorld W) {
  float lx = FLT_MAX, ly = FLT_MAX, hx = -FLT_MAX, hy = -FLT_MAX;
  GRID_STRIDE(p, W.nProxies) {
    if (!(W.p_flags[p] & PF_ALIVE)) continue;
    float4 f = W.p_fat[p];
    float cx = 0.5f * (f.x + f.z), cy = 0.5f * (f.y + f.w);
    lx = fminf(lx, cx); ly = fminf(ly, cy); hx = fmaxf(hx, cx); hy = fmaxf(hy, cy);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o));
    hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o));
  }
  if ((threadIdx.x & 31) == 0 && lx <= hx) {
    unsigned* b = (unsigned*)W.hdr->bounds;
    atomicMin(&b[0], f2ord(lx)); atomicMin(&b[1], f2ord(ly)); atomicMax(&b[2], f2ord(hx)); atomicMax(&b[3], f2ord(hy));
  }
}
DBX_D unsigned expand_bits15(unsigned v) {  // 15 bits -> every other bit
  v &= 0x7FFF;
  v = (v | (v << 8)) & 0x00FF00FF;
  v = (v | (v << 4)) & 0x0F0F0F0F;
  v = (v | (v << 2)) & 0x33333333;
  v = (v | (v << 1)) & 0x55555555;
  return v;
}
__global__ void __launch_bounds__(256) k_morton(const __grid_constant__ DevWorld W) {
  const unsigned* b = (const unsigned*)W.hdr->bounds;
  const float lx = ord2f(b[0]), ly = ord2f(b[1]), hx = ord2f(b[2]), hy = ord2f(b[3]);
  const float sx = hx > lx ? 32767.0f / (hx - lx) : 0.0f, sy = hy > ly ? 32767.0f / (hy - ly) : 0.0f;
  GRID_STRIDE(p, W.nProxies) {
    unsigned long long key;
    if (W.p_flags[p] & PF_ALIVE) {
      float4 f = W.p_fat[p];
      float cx = 0.5f * (f.x + f.z), cy = 0.5f * (f.y + f.w);
      unsigned ix = (unsigned)fminf(fmaxf((cx - lx) * sx, 0.0f), 32767.0f);
      unsigned iy = (unsigned)fminf(fmaxf((cy - ly) * sy, 0.0f), 32767.0f);
      unsigned m = expand_bits15(ix) | (expand_bits15(iy) << 1);
      key = ((unsigned long long)(unsigned)W.b_world[W.p_ids[p].z] << 30) | m;
    } else {
      key = (unsigned long long)(unsigned)W.nWorlds << 30;  // dead slots sort last (past every replica) and get an empty box
    }
    W.bv_key[p] = key;
    W.bv_leaf[p] = p;
  }
}
// Karras 2012: each internal node finds its key range from common-prefix lengths
DBX_D int lbvh_delta(const unsigned long long* keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll((long long)(a ^ b));
}
__global__ void __launch_bounds__(256) k_lbvh_hierarchy(const __grid_constant__ DevWorld W, const unsigned long long* keys) {
  const int n = W.nProxies;
  GRID_STRIDE(i, n - 1) {
    int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1) if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
      t = (t + 1) >> 1;
      if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int left = (min(i, j) == gamma) ? (n - 1 + gamma) : gamma;            // leaves live at [n-1, 2n-1)
    int right = (max(i, j) == gamma + 1) ? (n - 1 + gamma + 1) : gamma + 1;
    W.bv_child[i] = make_int2(left, right);
    W.bv_wr[i] = make_int2((int)(keys[min(i, j)] >> 30), (int)(keys[max(i, j)] >> 30));   // replica range under this node
    W.bv_parent[left] = i;
    W.bv_parent[right] = i;
    W.bv_visit[i] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) W.bv_parent[0] = -1;
}
__global__ void __launch_bounds__(256) k_lbvh_refit(const __grid_constant__ DevWorld W, const int* leaves) {
  const int n = W.nProxies;
  GRID_STRIDE(k, n) {
    int p = leaves[k];
    float4 box = (W.p_flags[p] & PF_ALIVE) ? W.p_fat[p] : make_float4(FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX);
    int node = n - 1 + k;
    W.bv_box[node] = box;
    W.bv_pos[p] = k;
    if (n == 1) { W.bv_parent[0] = -1; continue; }
    int parent = W.bv_parent[node];
    while (parent >= 0) {
      __threadfence();
      if (atomicAdd(&W.bv_visit[parent], 1) == 0) break;   // the second child to arrive continues upward
      int2 ch = W.bv_child[parent];
      float4 a = __ldcg(&W.bv_box[ch.x]), b = __ldcg(&W.bv_box[ch.y]);
      float4 u = make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
      __stcg(&W.bv_box[parent], u);
      parent = W.bv_parent[parent];
    }
  }
}

// warp-cooperative query: one warp per moved proxy walks the tree with a shared frontier; each lane tests one node
// The frontier is a LIFO in shared memory popped 32 nodes at a time (32 interleaved depth-first walks: it holds about
// 32 x depth nodes for a box that overlaps everything, e.g. a wall of a rotating container).  When it gets within
// kQueryReserve of its capacity the walk degrades to one node at a time, a plain depth-first walk whose stack grows by
// at most one per level, so any tree up to kQueryReserve deep is traversed whatever the query box is.
constexpr int kQueryStack = 1024;      // k_query: 8 warps x 4 KB
constexpr int kQueryStackToi = 256;    // k_toi: 16 warps x 1 KB
constexpr int kQueryReserve = 96;
DBX_D void query_proxy(const DevWorld& W, const int* leaves, int* stack, int cap, int lane, int p) {
  const int n = W.nProxies;
  if (!(W.p_flags[p] & PF_ALIVE)) return;
  const Box fat = BX(__ldcg(&W.p_fat[p]));
  const int keyP = W.p_key[p];
  const int worldP = W.b_world[W.p_ids[p].z];
  int top = 1;
  if (lane == 0) stack[0] = (n == 1) ? (n - 1) : 0;
  __syncwarp();
  while (top > 0) {
    int take = (top + 32 <= cap - kQueryReserve) ? min(top, 32) : 1;   // pop up to 32 nodes; one when nearly full
    int node = lane < take ? stack[top - take + lane] : -1;
    top -= take;
    __syncwarp();
    bool hit = false;
    bool isLeaf = node >= n - 1;
    // the children are fetched together with the box, not after the overlap test: one L2 round trip per level, not two
    int2 ch = make_int2(-1, -1);
    if (node >= 0 && !isLeaf) ch = W.bv_child[node];
    if (node >= 0) {
      hit = overlap(fat, BX(__ldcg(&W.bv_box[node])));
      // replicas share coordinates: without this test every query would descend into every replica's subtree
      if (hit && !isLeaf && W.nWorlds > 1) { const int2 wr = W.bv_wr[node]; hit = worldP >= wr.x && worldP <= wr.y; }
    }
    if (hit && isLeaf) {
      int q = leaves[node - (n - 1)];
      // each unordered pair once: from the lower-key proxy when both moved (UpdatePairs sorts and dedups the same set)
      if (q != p && (W.p_flags[q] & PF_ALIVE) && W.b_world[W.p_ids[q].z] == worldP) {
        int keyQ = W.p_key[q];
        bool qMoved = (W.p_flags[q] & PF_MOVED) != 0;
        if (!qMoved || keyP < keyQ) {
          int slot = atomicAdd(&W.hdr->nPairs, 1);
          if (slot < W.pairCap) W.pairs[slot] = keyP < keyQ ? make_int2(p, q) : make_int2(q, p);
          else W.hdr->error = E_PAIRS;
        }
      }
    }
    bool push = hit && !isLeaf;
    unsigned ballot = __ballot_sync(0xffffffffu, push);
    int offset = __popc(ballot & ((1u << lane) - 1));
    int total = __popc(ballot);
    if (top + 2 * total > cap) { if (lane == 0) W.hdr->error = E_QUERY_STACK; break; }   // tree deeper than kQueryReserve: report, never corrupt
    if (push) {
      int base = top + 2 * offset;
      stack[base] = ch.x; stack[base + 1] = ch.y;
    }
    top += 2 * total;
    __syncwarp();
  }
  __syncwarp();
}
__global__ void __launch_bounds__(256) k_query(const __grid_constant__ DevWorld W, const int* leaves) {
  __shared__ int stacks[8][kQueryStack];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nMoved = min(W.hdr->nMoved, W.moveCap);
  for (int mIdx = warp; mIdx < nMoved; mIdx += nwarps) query_proxy(W, leaves, stacks[wib], kQueryStack, lane, W.moveList[mIdx]);
}

// ------------------------------------------------------------------------------------------------ world queries
// b2World.RayCast / QueryAABB on the LBVH (dynamics/b2world.d:563-587), batched: one thread per ray / per box.
// b2Shape.RayCast: circle b2circleshape.d:67-94, edge (and chain children) b2edgeshape.d:96-150, polygon b2polygonshape.d:279-332
DBX_D bool shape_raycast(const DShape* S, Xf xf, v2 P1, v2 P2, float maxFraction, float* fraction, v2* normalOut) {
  if (S->type == SH_CIRCLE) {
    const v2 position = xf.p + mul(xf.q, S->c);
    const v2 s = P1 - position;
    const float b = dot(s, s) - S->radius * S->radius;
    const v2 r = P2 - P1;
    const float c = dot(s, r);
    const float rr = dot(r, r);
    const float sigma = c * c - rr * b;
    if (sigma < 0.0f || rr < kEpsilon) return false;
    float a = -(c + sqrtf(sigma));
    if (0.0f <= a && a <= maxFraction * rr) {
      a /= rr;
      *fraction = a;
      v2 n = s + a * r;
      normalize(n);
      *normalOut = n;
      return true;
    }
    return false;
  }
  const v2 p1 = mulT(xf.q, P1 - xf.p), p2 = mulT(xf.q, P2 - xf.p);
  const v2 d = p2 - p1;
  if (S->type == SH_EDGE) {
    const v2 v1 = S->v[1], v2_ = S->v[2];
    const v2 e = v2_ - v1;
    v2 normal = V(e.y, -e.x);
    normalize(normal);
    const float numerator = dot(normal, v1 - p1);
    const float denominator = dot(normal, d);
    if (denominator == 0.0f) return false;
    const float t = numerator / denominator;
    if (t < 0.0f || maxFraction < t) return false;
    const v2 q = p1 + t * d;
    const v2 r = v2_ - v1;
    const float rr = dot(r, r);
    if (rr == 0.0f) return false;
    const float sc = dot(q - v1, r) / rr;
    if (sc < 0.0f || 1.0f < sc) return false;
    *fraction = t;
    *normalOut = numerator > 0.0f ? -mul(xf.q, normal) : mul(xf.q, normal);
    return true;
  }
  float lower = 0.0f, upper = maxFraction;
  int index = -1;
  for (int i = 0; i < S->count; ++i) {
    const float numerator = dot(S->n[i], S->v[i] - p1);
    const float denominator = dot(S->n[i], d);
    if (denominator == 0.0f) {
      if (numerator < 0.0f) return false;
    } else {
      if (denominator < 0.0f && numerator < lower * denominator) { lower = numerator / denominator; index = i; }
      else if (denominator > 0.0f && numerator < upper * denominator) upper = numerator / denominator;
    }
    if (upper < lower) return false;
  }
  if (index >= 0) { *fraction = lower; *normalOut = mul(xf.q, S->n[index]); return true; }
  return false;
}

constexpr int kRayStack = 96;
// out: 8 floats per ray = (fixture, child, fraction, point.x, point.y, normal.x, normal.y, -) with ints stored bitwise
__global__ void __launch_bounds__(128) k_raycast(const __grid_constant__ DevWorld W, const int* leaves, const float4* rays, int nRays, float4* out) {
  const int n = W.nProxies;
  GRID_STRIDE(k, nRays) {
    const float4 ry = rays[k];
    const v2 p1 = V(ry.x, ry.y), p2 = V(ry.z, ry.w);
    int bestFixture = -1, bestChild = 0, bestKey = 0x7fffffff;
    float maxFraction = 1.0f;
    v2 bestNormal = V(0.0f, 0.0f);
    v2 r = p2 - p1;
    if (n > 0 && dot(r, r) > 0.0f) {
      normalize(r);
      const v2 v = cross(1.0f, r), abs_v = V(fabsr(v.x), fabsr(v.y));
      int stack[kRayStack];
      int top = 0;
      stack[top++] = (n == 1) ? (n - 1) : 0;
      while (top > 0) {
        const int node = stack[--top];
        const float4 bx = __ldcg(&W.bv_box[node]);
        const v2 t = p1 + maxFraction * (p2 - p1);
        const v2 slo = vmin(p1, t), shi = vmax(p1, t);
        if (bx.z < slo.x || bx.w < slo.y || shi.x < bx.x || shi.y < bx.y) continue;       // b2TestOverlap(node.aabb, segmentAABB)
        const v2 c = 0.5f * V(bx.x + bx.z, bx.y + bx.w), h = 0.5f * V(bx.z - bx.x, bx.w - bx.y);
        const float separation = fabsr(dot(v, p1 - c)) - dot(abs_v, h);
        if (separation > 0.0f) continue;
        if (node >= n - 1) {
          const int q = leaves[node - (n - 1)];
          if (!(W.p_flags[q] & PF_ALIVE)) continue;
          const int4 ids = W.p_ids[q];   // fixture child body shape
          float fraction; v2 normal;
          if (!shape_raycast(W.shapes + ids.w, XF(W.b_xf[ids.z]), p1, p2, maxFraction, &fraction, &normal)) continue;
          const int key = W.p_key[q];
          if (fraction < maxFraction || bestFixture < 0 || key < bestKey) {
            bestFixture = ids.x; bestChild = ids.y; bestKey = key; bestNormal = normal; maxFraction = fraction;
          }
        } else {
          if (top + 2 > kRayStack) { W.hdr->error = E_QUERY_STACK; break; }
          const int2 ch = W.bv_child[node];
          stack[top++] = ch.x; stack[top++] = ch.y;
        }
      }
    }
    const v2 point = (1.0f - maxFraction) * p1 + maxFraction * p2;
    out[2 * k] = make_float4(__int_as_float(bestFixture), __int_as_float(bestChild), maxFraction, point.x);
    out[2 * k + 1] = make_float4(point.y, bestNormal.x, bestNormal.y, 0.0f);
  }
}
// counts[k] = number of proxies whose fat box overlaps boxes[k]; the first capPer of them as (fixture, child) in out
__global__ void __launch_bounds__(128) k_query_aabb(const __grid_constant__ DevWorld W, const int* leaves, const float4* boxes, int nBoxes, int capPer, int* counts, int2* out) {
  const int n = W.nProxies;
  GRID_STRIDE(k, nBoxes) {
    const Box box = BX(boxes[k]);
    int count = 0;
    if (n > 0) {
      int stack[kRayStack];
      int top = 0;
      stack[top++] = (n == 1) ? (n - 1) : 0;
      while (top > 0) {
        const int node = stack[--top];
        if (!overlap(BX(__ldcg(&W.bv_box[node])), box)) continue;
        if (node >= n - 1) {
          const int q = leaves[node - (n - 1)];
          if (!(W.p_flags[q] & PF_ALIVE) || !overlap(BX(W.p_fat[q]), box)) continue;   // the leaf box may be wider than the fat box
          const int4 ids = W.p_ids[q];
          if (count < capPer) out[(size_t)k * capPer + count] = make_int2(ids.x, ids.y);
          ++count;
        } else {
          if (top + 2 > kRayStack) { W.hdr->error = E_QUERY_STACK; break; }
          const int2 ch = W.bv_child[node];
          stack[top++] = ch.x; stack[top++] = ch.y;
        }
      }
    }
    counts[k] = count;
  }
}

// b2World.RayCast with a callback that returns 1 (every fixture on the ray, no clipping): the ray keeps its full length
// (b2dynamictree.d:303-316 with value == maxFraction == 1), each hit is one record of two float4 as in k_raycast
__global__ void __launch_bounds__(128) k_raycast_all(const __grid_constant__ DevWorld W, const int* leaves, const float4* rays, int nRays, int capPer, int* counts, float4* out) {
  const int n = W.nProxies;
  GRID_STRIDE(k, nRays) {
    const float4 ry = rays[k];
    const v2 p1 = V(ry.x, ry.y), p2 = V(ry.z, ry.w);
    int count = 0;
    v2 r = p2 - p1;
    if (n > 0 && dot(r, r) > 0.0f) {
      normalize(r);
      const v2 v = cross(1.0f, r), abs_v = V(fabsr(v.x), fabsr(v.y));
      const v2 slo = vmin(p1, p2), shi = vmax(p1, p2);
      int stack[kRayStack];
      int top = 0;
      stack[top++] = (n == 1) ? (n - 1) : 0;
      while (top > 0) {
        const int node = stack[--top];
        const float4 bx = __ldcg(&W.bv_box[node]);
        if (bx.z < slo.x || bx.w < slo.y || shi.x < bx.x || shi.y < bx.y) continue;
        const v2 c = 0.5f * V(bx.x + bx.z, bx.y + bx.w), h = 0.5f * V(bx.z - bx.x, bx.w - bx.y);
        const float separation = fabsr(dot(v, p1 - c)) - dot(abs_v, h);
        if (separation > 0.0f) continue;
        if (node >= n - 1) {
          const int q = leaves[node - (n - 1)];
          if (!(W.p_flags[q] & PF_ALIVE)) continue;
          const int4 ids = W.p_ids[q];   // fixture child body shape
          float fraction; v2 normal;
          if (!shape_raycast(W.shapes + ids.w, XF(W.b_xf[ids.z]), p1, p2, 1.0f, &fraction, &normal)) continue;
          if (count < capPer) {
            const v2 point = (1.0f - fraction) * p1 + fraction * p2;
            float4* o = out + 2 * ((size_t)k * capPer + count);
            o[0] = make_float4(__int_as_float(ids.x), __int_as_float(ids.y), fraction, point.x);
            o[1] = make_float4(point.y, normal.x, normal.y, 0.0f);
          }
          ++count;
        } else {
          if (top + 2 > kRayStack) { W.hdr->error = E_QUERY_STACK; break; }
          const int2 ch = W.bv_child[node];
          stack[top++] = ch.x; stack[top++] = ch.y;
        }
      }
    }
    counts[k] = count;
  }
}
// b2Shape.TestPoint: polygon b2polygonshape.d:265-279, circle b2circleshape.d:60-65; an edge (and so a chain child) contains
// no point (b2edgeshape.d:84-87, b2chainshape.d:196-199)
DBX_D bool shape_test_point(const DShape* S, Xf xf, v2 p) {
  if (S->type == SH_CIRCLE) {
    const v2 center = xf.p + mul(xf.q, S->c);
    const v2 d = p - center;
    return dot(d, d) <= S->radius * S->radius;
  }
  if (S->type != SH_POLYGON) return false;
  const v2 pLocal = mulT(xf.q, p - xf.p);
  for (int i = 0; i < S->count; ++i) {
    const float d = dot(S->n[i], pLocal - S->v[i]);
    if (d > 0.0f) return false;
  }
  return true;
}
// b2Fixture.TestPoint (b2fixture.d:209-212), batched: q = (point.x, point.y, shape index | -1, body index), one thread each
__global__ void __launch_bounds__(128) k_test_points(const __grid_constant__ DevWorld W, const float4* q, int n, int* inside) {
  GRID_STRIDE(k, n) {
    const float4 t = q[k];
    const int shape = __float_as_int(t.z), body = __float_as_int(t.w);
    inside[k] = (shape >= 0 && shape < W.nShapes && body >= 0 && body < W.nBodies) ? (shape_test_point(W.shapes + shape, XF(W.b_xf[body]), V(t.x, t.y)) ? 1 : 0) : 0;
  }
}
// b2World.ShiftOrigin (dynamics/b2world.d:758-780): bodies (m_xf.p, m_sweep.c0, m_sweep.c) and the tree's boxes
// (b2dynamictree.d:503-511: the persistent fat AABBs; b2FixtureProxy.aabb is left as it is there, and here).  xf0 is this
// library's cached transform at (c0, a0) and moves with c0.  The LBVH is rebuilt by the host before its next use.
__global__ void __launch_bounds__(256) k_shift_origin(const __grid_constant__ DevWorld W, float ox, float oy) {
  GRID_STRIDE(b, W.nBodies) {
    if (!(W.b_flags[b] & BF_ALIVE)) continue;
    float4 xf = W.b_xf[b], xf0 = W.b_xf0[b], pos = W.b_pos[b], pos0 = W.b_pos0[b];
    xf.x -= ox; xf.y -= oy; xf0.x -= ox; xf0.y -= oy; pos.x -= ox; pos.y -= oy; pos0.x -= ox; pos0.y -= oy;
    W.b_xf[b] = xf; W.b_xf0[b] = xf0; W.b_pos[b] = pos; W.b_pos0[b] = pos0;
  }
  GRID_STRIDE(p, W.nProxies) {
    if (!(W.p_flags[p] & PF_ALIVE)) continue;
    float4 fat = W.p_fat[p];
    fat.x -= ox; fat.y -= oy; fat.z -= ox; fat.w -= oy;
    W.p_fat[p] = fat;
  }
}
// b2Contact.GetWorldManifold (contacts/b2contact.d:77-91) -> b2WorldManifold.Initialize (collision/b2collision.d:123-191) for the
// contact slots [0, high): out[2i] = (normal.xy, separations), out[2i+1] = (points[0], points[1]); untouched for dead slots
__global__ void __launch_bounds__(256) k_world_manifolds(const __grid_constant__ DevWorld W, int high, float4* out) {
  GRID_STRIDE(i, high) {
    float4 o0 = make_float4(0, 0, 0, 0), o1 = make_float4(0, 0, 0, 0);
    const uint32_t fl = W.c_flags[i];
    const uint4 mk = W.c_mk[i];
    const int pointCount = (int)mk.w, type = (int)mk.z;
    if ((fl & CF_ALIVE) && pointCount > 0) {
      const int4 ids = W.c_ids[i], fx = W.c_fix[i];
      const Xf xfA = XF(W.b_xf[ids.z]), xfB = XF(W.b_xf[ids.w]);
      const float radiusA = W.shapes[fx.z].radius, radiusB = W.shapes[fx.w].radius;
      const float4 m0 = W.c_m0[i], m1 = W.c_m1[i];
      const v2 localNormal = V(m0.x, m0.y), localPoint = V(m0.z, m0.w);
      const v2 lp[2] = {V(m1.x, m1.y), V(m1.z, m1.w)};
      v2 normal = V(0.0f, 0.0f), wp[2] = {V(0.0f, 0.0f), V(0.0f, 0.0f)};
      float sep[2] = {0.0f, 0.0f};
      if (type == MAN_CIRCLES) {
        normal = V(1.0f, 0.0f);
        const v2 pointA = mul(xfA, localPoint), pointB = mul(xfB, lp[0]);
        if (dist2(pointA, pointB) > kEpsilon * kEpsilon) { normal = pointB - pointA; normalize(normal); }
        const v2 cA = pointA + radiusA * normal, cB = pointB - radiusB * normal;
        wp[0] = 0.5f * (cA + cB);
        sep[0] = dot(cB - cA, normal);
      } else if (type == MAN_FACE_A) {
        normal = mul(xfA.q, localNormal);
        const v2 planePoint = mul(xfA, localPoint);
        for (int k = 0; k < pointCount && k < 2; ++k) {
          const v2 clipPoint = mul(xfB, lp[k]);
          const v2 cA = clipPoint + (radiusA - dot(clipPoint - planePoint, normal)) * normal;
          const v2 cB = clipPoint - radiusB * normal;
          wp[k] = 0.5f * (cA + cB);
          sep[k] = dot(cB - cA, normal);
        }
      } else if (type == MAN_FACE_B) {
        normal = mul(xfB.q, localNormal);
        const v2 planePoint = mul(xfB, localPoint);
        for (int k = 0; k < pointCount && k < 2; ++k) {
          const v2 clipPoint = mul(xfA, lp[k]);
          const v2 cB = clipPoint + (radiusB - dot(clipPoint - planePoint, normal)) * normal;
          const v2 cA = clipPoint - radiusA * normal;
          wp[k] = 0.5f * (cA + cB);
          sep[k] = dot(cA - cB, normal);
        }
        normal = -normal;
      }
      o0 = make_float4(normal.x, normal.y, sep[0], sep[1]);
      o1 = make_float4(wp[0].x, wp[0].y, wp[1].x, wp[1].y);
    }
    out[2 * (size_t)i] = o0; out[2 * (size_t)i + 1] = o1;
  }
}
// PostSolve records of the island solve that just ran (b2island.d:239): one per solver contact
__global__ void __launch_bounds__(256) k_post_solve(const __grid_constant__ DevWorld W) {
  const int n = min(W.hdr->nSolve, W.sCap);
  GRID_STRIDE(s, n) emit_post_solve(W, 1, s, W.s_contact[s]);
}

// b2ContactManager.AddPair (dynamics/b2contactmanager.d:52-176) + b2Contact.Create (contacts/b2contact.d:375-400)
DBX_D void add_pair(const DevWorld& W, int2 pr) {
  const int4 pa = W.p_ids[pr.x], pb = W.p_ids[pr.y];   // fixture child body shape
  int bodyA = pa.z, bodyB = pb.z;
  if (bodyA == bodyB) return;
  const unsigned long long key = ((unsigned long long)(unsigned)W.p_key[pr.x] << 32) | (unsigned)W.p_key[pr.y];
  if (hash_find(W, key) >= 0) return;                   // a contact for this (fixture, child) pair already exists (:75-100)
  const uint32_t flA = W.b_flags[bodyA], flB = W.b_flags[bodyB];
  if (!body_should_collide(W, bodyB, bodyA, flB, flA)) return;
  if (!(W.userFilter & 2) && !filter_should_collide(W, pa.x, pb.x)) return;   // bit 1: the user's filter replaces the default one
  // type registry (b2contact.d:425-437): A must be the primary type
  const int t1 = W.shapes[pa.w].type, t2 = W.shapes[pb.w].type;
  bool has, primary;
  if (t1 == SH_EDGE && t2 == SH_EDGE) { has = false; primary = false; }
  else if (t1 == SH_CIRCLE) { has = true; primary = (t2 == SH_CIRCLE); }
  else if (t1 == SH_EDGE) { has = true; primary = true; }
  else { has = true; primary = (t2 != SH_EDGE); }        // polygon vs circle/polygon primary; vs edge swapped
  if (!has) return;
  int proxyA = pr.x, proxyB = pr.y;
  int4 ia = pa, ib = pb;
  if (!primary) { int t = proxyA; proxyA = proxyB; proxyB = t; int4 tt = ia; ia = ib; ib = tt; }
  int slot;
  int f = atomicSub(&W.hdr->nFree, 1);
  if (f > 0) slot = W.c_free[f - 1];
  else { atomicAdd(&W.hdr->nFree, 1); slot = atomicAdd(&W.hdr->cHigh, 1); }
  if (slot >= W.cCap) { W.hdr->error = E_CONTACTS; atomicSub(&W.hdr->cHigh, 1); return; }
  const bool sensor = ((W.f_group[ia.x] >> 16) & FXF_SENSOR) || ((W.f_group[ib.x] >> 16) & FXF_SENSOR);
  const float2 mA = W.f_mat[ia.x], mB = W.f_mat[ib.x];
  W.c_key[slot] = key;
  W.c_ids[slot] = make_int4(proxyA, proxyB, ia.z, ib.z);
  W.c_fix[slot] = make_int4(ia.x, ib.x, ia.w, ib.w);
  W.c_flags[slot] = CF_ALIVE | CF_ENABLED | CF_FRESH | (sensor ? CF_SENSOR : 0) | ((W.userFilter & 1) ? CF_NEW : 0);
  { const int k = atomicAdd(&W.hdr->nFresh, 1); if (k < W.cCap) W.c_work[k] = slot; }   // c_work is idle between the colouring pass and the next step
  W.c_m0[slot] = make_float4(0, 0, 0, 0);
  W.c_m1[slot] = make_float4(0, 0, 0, 0);
  W.c_imp[slot] = make_float4(0, 0, 0, 0);
  W.c_mk[slot] = make_uint4(0, 0, 0, 0);
  // b2MixFriction / b2MixRestitution (b2contact.d:32-42)
  W.c_mat[slot] = make_float4(sqrtf(mA.x * mB.x), mA.y > mB.y ? mA.y : mB.y, 0.0f, 1.0f);
  W.c_toiCount[slot] = 0;
  W.c_colour[slot] = -1;
  if (!hash_insert(W, key, slot)) W.hdr->error = E_HASH;
  if (!sensor) { wake_body_now(W, ia.z); wake_body_now(W, ib.z); }   // :168-173
}
__global__ void __launch_bounds__(256) k_add_pairs(const __grid_constant__ DevWorld W) {
  const int n = min(W.hdr->nPairs, W.pairCap);
  GRID_STRIDE(k, n) add_pair(W, W.pairs[k]);
}
__global__ void __launch_bounds__(256) k_clear_moves(const __grid_constant__ DevWorld W) {
  const int nMoved = min(W.hdr->nMoved, W.moveCap);
  GRID_STRIDE(k, nMoved) { int p = W.moveList[k]; W.p_flags[p] &= ~PF_MOVED; }
}
__global__ void k_reset_moves(const __grid_constant__ DevWorld W) { W.hdr->nMoved = 0; }

// ------------------------------------------------------------------------------------------------ maintenance
__global__ void __launch_bounds__(256) k_hash_clear(const __grid_constant__ DevWorld W) {
  GRID_STRIDE(i, W.hCap) W.h_key[i] = kHashEmpty;
  if (blockIdx.x == 0 && threadIdx.x == 0) { W.hdr->nTomb = 0; }
}
__global__ void __launch_bounds__(256) k_hash_fill(const __grid_constant__ DevWorld W) {
  const int n = W.hdr->cHigh;
  GRID_STRIDE(i, n) if (W.c_flags[i] & CF_ALIVE) { if (!hash_insert(W, W.c_key[i], i)) W.hdr->error = E_HASH; }
}
__global__ void __launch_bounds__(256) k_count(const __grid_constant__ DevWorld W) {
  const int n = W.hdr->cHigh;
  int alive = 0, touching = 0, awake = 0;
  GRID_STRIDE(i, n) { uint32_t f = W.c_flags[i]; if (f & CF_ALIVE) { ++alive; if (f & CF_TOUCHING) ++touching; } }
  GRID_STRIDE(b, W.nBodies) { uint32_t f = W.b_flags[b]; if ((f & BF_ALIVE) && (f & BF_AWAKE) && body_type(f) != BODY_STATIC) ++awake; }
  for (int o = 16; o > 0; o >>= 1) {
    alive += __shfl_xor_sync(0xffffffffu, alive, o); touching += __shfl_xor_sync(0xffffffffu, touching, o); awake += __shfl_xor_sync(0xffffffffu, awake, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (alive) atomicAdd(&W.hdr->nContacts, alive);
    if (touching) atomicAdd(&W.hdr->nTouching, touching);
    if (awake) atomicAdd(&W.hdr->nAwake, awake);
  }
}
__global__ void k_count_reset(const __grid_constant__ DevWorld W) { W.hdr->nContacts = 0; W.hdr->nTouching = 0; W.hdr->nAwake = 0; }
__global__ void __launch_bounds__(256) k_set_levels(const __grid_constant__ DevWorld W, const int* levels, int n) {
  GRID_STRIDE(i, n) W.c_colour[i] = levels[i];
}
// after a state import: contact slots [0, n) are all alive, nothing is free
__global__ void k_import_reset(const __grid_constant__ DevWorld W, int n) { W.hdr->cHigh = n; W.hdr->nFree = 0; }


// ------------------------------------------------------------------------------------------------ contact compaction
// Contact slots are handed out by atomics, so after a while neighbours in space are strangers in memory.  Every few
// dozen steps the alive contacts are re-packed in reference pair-key order (= proxy creation order = spatial order for
// scenes built in a sweep), which restores coalescing for Collide, the island pass and constraint setup.
__global__ void __launch_bounds__(256) k_compact_keys(const __grid_constant__ DevWorld W, int n, unsigned long long* keys, int* vals) {
  GRID_STRIDE(i, n) { keys[i] = (W.c_flags[i] & CF_ALIVE) ? W.c_key[i] : ~0ull; vals[i] = i; }
}
template <class T> __global__ void __launch_bounds__(256) k_gather(const T* __restrict__ src, T* __restrict__ dst, const int* __restrict__ perm, int n) {
  GRID_STRIDE(i, n) dst[i] = src[perm[i]];
}
__global__ void k_compact_finish(const __grid_constant__ DevWorld W, int nAlive, int oldHigh) {
  GRID_STRIDE(i, oldHigh) if (i >= nAlive) { W.c_flags[i] = 0; W.c_colour[i] = -1; }
  if (blockIdx.x == 0 && threadIdx.x == 0) { W.hdr->cHigh = nAlive; W.hdr->nFree = 0; }
}
template <class T> static cudaError_t permute_array(const LaunchCfg& L, T* arr, void* scratch, const int* perm, int n) {
  ++L.launches; k_gather<T><<<L.gridWide, 256, 0, L.stream>>>(arr, (T*)scratch, perm, n);
  return cudaMemcpyAsync(arr, scratch, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, L.stream);
}
// `high` = current hdr->cHigh, `nAlive` = alive contacts (both read back by the caller); scratch >= 16 * high bytes
cudaError_t stage_compact_contacts(DevWorld& W, const LaunchCfg& L, int high, int nAlive, void* scratch, unsigned long long* keyA, unsigned long long* keyB, int* valA, int* valB) {
  if (high <= 0) return cudaSuccess;
  ++L.launches; k_compact_keys<<<L.gridWide, 256, 0, L.stream>>>(W, high, keyA, valA);
  cub::DoubleBuffer<unsigned long long> keys(keyA, keyB);
  cub::DoubleBuffer<int> vals(valA, valB);
  size_t bytes = L.cubTempBytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(L.cubTemp, bytes, keys, vals, high, 0, 64, L.stream);
  if (e != cudaSuccess) return e;
  const int* perm = vals.Current();
  if ((e = permute_array(L, W.c_key, scratch, perm, high)) != cudaSuccess) return e;
  if ((e = permute_array(L, W.c_ids, scratch, perm, high)) != cudaSuccess) return e;
  if ((e = permute_array(L, W.c_fix, scratch, perm, high)) != cudaSuccess) return e;
  if ((e = permute_array(L, W.c_flags, scratch, perm, high)) != cudaSuccess) return e;
  if ((e = permute_array(L, W.c_m0, scratch, perm, high)) != cudaSuccess) return e;
  if ((e = permute_array(L, W.c_m1, scratch, perm, high)) != cudaSuccess) return e;
  if ((e = permute_array(L, W.c_imp, scratch, perm, high)) != cudaSuccess) return e;
  if ((e = permute_array(L, W.c_mk, scratch, perm, high)) != cudaSuccess) return e;
  if ((e = permute_array(L, W.c_mat, scratch, perm, high)) != cudaSuccess) return e;
  if ((e = permute_array(L, W.c_toiCount, scratch, perm, high)) != cudaSuccess) return e;
  if ((e = permute_array(L, W.c_colour, scratch, perm, high)) != cudaSuccess) return e;
  ++L.launches; k_compact_finish<<<L.gridWide, 256, 0, L.stream>>>(W, nAlive, high);
  return stage_rebuild_hash(W, L);
}

// ------------------------------------------------------------------------------------------------ replicas
// dbx_world_replicate: replica r > 0 of every body / fixture / proxy is a copy of replica 0 with its indices shifted.
__global__ void __launch_bounds__(256) k_replicate(const __grid_constant__ DevWorld W, int nB, int nF, int nP, int nMoved, int keyStride, int copies) {
  GRID_STRIDE(idx, nB * copies) {
    const int r = idx / nB, b = idx - r * nB;
    if (r > 0) {
      W.b_xf[idx] = W.b_xf[b]; W.b_xf0[idx] = W.b_xf0[b]; W.b_pos[idx] = W.b_pos[b]; W.b_pos0[idx] = W.b_pos0[b]; W.b_vel[idx] = W.b_vel[b];
      W.b_force[idx] = W.b_force[b]; W.b_mass[idx] = W.b_mass[b]; W.b_lc[idx] = W.b_lc[b]; W.b_gs[idx] = W.b_gs[b]; W.b_flags[idx] = W.b_flags[b];
    }
    W.b_world[idx] = r;
  }
  GRID_STRIDE(idx, nF * copies) {
    const int r = idx / nF, f = idx - r * nF;
    if (r > 0) { W.f_body[idx] = W.f_body[f] + r * nB; W.f_mat[idx] = W.f_mat[f]; W.f_filter[idx] = W.f_filter[f]; W.f_group[idx] = W.f_group[f]; }
  }
  GRID_STRIDE(idx, nP * copies) {
    const int r = idx / nP, p = idx - r * nP;
    if (r > 0) {
      int4 ids = W.p_ids[p];
      ids.x += r * nF; ids.z += r * nB;
      W.p_ids[idx] = ids;
      W.p_key[idx] = W.p_key[p] + r * keyStride;
      W.p_aabb[idx] = W.p_aabb[p]; W.p_fat[idx] = W.p_fat[p]; W.p_flags[idx] = W.p_flags[p];
    }
  }
  GRID_STRIDE(idx, nMoved * copies) {
    const int r = idx / nMoved, k = idx - r * nMoved;
    if (r > 0) W.moveList[idx] = W.moveList[k] + r * nP;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { W.hdr->nMoved = nMoved * copies; if (nMoved * copies > W.moveCap) W.hdr->error = E_MOVES; }
}

// ------------------------------------------------------------------------------------------------ API-time edits
// DestroyFixture / DestroyBody / SetActive(false) (b2body.d:211-227, b2world.d:136-145) destroy the matching contacts;
// CreateJoint / DestroyJoint with collideConnected == false flag them for re-filtering (b2world.d:241-256, 344-359).
__global__ void __launch_bounds__(256) k_api_contacts(const __grid_constant__ DevWorld W, int body, int fixture, int otherBody, int flagOnly) {
  const int n = W.hdr->cHigh;
  GRID_STRIDE(i, n) {
    uint32_t flags = W.c_flags[i];
    if (!(flags & CF_ALIVE)) continue;
    const int4 ids = W.c_ids[i];
    const int4 fx = W.c_fix[i];
    bool match;
    if (fixture >= 0) match = fx.x == fixture || fx.y == fixture;
    else if (otherBody >= 0) match = (ids.z == body && ids.w == otherBody) || (ids.z == otherBody && ids.w == body);
    else match = ids.z == body || ids.w == body;
    if (!match) continue;
    if (flagOnly) { W.c_flags[i] = flags | CF_FILTER; continue; }
    const int pointCount = (int)W.c_mk[i].w;
    if (pointCount > 0 && !(flags & CF_SENSOR)) { wake_body_now(W, ids.z); wake_body_now(W, ids.w); }
    if (flags & CF_TOUCHING) emit_contact_event(W, EV_END, 3, i, ids, fx);
    hash_remove(W, W.c_key[i]);
    W.c_flags[i] = 0;
    W.c_colour[i] = -1;
    int slot = atomicAdd(&W.hdr->nFree, 1);
    W.c_free[slot] = i;
  }
}
// mask bits: 1 enabled, 2 friction, 4 restitution, 8 tangentSpeed; bit 8 carries the `enabled` value
__global__ void __launch_bounds__(256) k_patch_contacts(const __grid_constant__ DevWorld W, const unsigned long long* keys, const float4* vals, const int* masks, int n) {
  GRID_STRIDE(k, n) {
    const int i = hash_find(W, keys[k]);
    if (i < 0) continue;
    const int m = masks[k];
    if (m & 16) {   // vetoed by the user's b2ContactFilter: b2ContactManager.Destroy (b2contactmanager.d:183-246)
      const uint32_t flags = W.c_flags[i];
      const int4 ids = W.c_ids[i];
      if ((int)W.c_mk[i].w > 0 && !(flags & CF_SENSOR)) { wake_body_now(W, ids.z); wake_body_now(W, ids.w); }
      if (flags & CF_TOUCHING) emit_contact_event(W, EV_END, 3, i, ids, W.c_fix[i]);
      hash_remove(W, W.c_key[i]);
      W.c_flags[i] = 0;
      W.c_colour[i] = -1;
      const int slot = atomicAdd(&W.hdr->nFree, 1);
      W.c_free[slot] = i;
      continue;
    }
    if (m & 1) { uint32_t f = W.c_flags[i]; W.c_flags[i] = (m & 0x100) ? ((f | CF_ENABLED) & ~CF_PRESOLVE_OFF) : ((f & ~CF_ENABLED) | CF_PRESOLVE_OFF); }
    if (m & 14) {
      float4 mat = W.c_mat[i];
      const float4 v = vals[k];
      if (m & 2) mat.x = v.x;
      if (m & 4) mat.y = v.y;
      if (m & 8) mat.z = v.z;
      W.c_mat[i] = mat;
    }
  }
}
// user contact filter, deferred: list (and untag unless `peek`) the contacts created since the last poll
__global__ void __launch_bounds__(256) k_list_new_contacts(const __grid_constant__ DevWorld W, int4* out, unsigned long long* keys, int cap, int peek) {
  const int n = W.hdr->cHigh;
  GRID_STRIDE(i, n) {
    const uint32_t flags = W.c_flags[i];
    if ((flags & (CF_ALIVE | CF_NEW)) != (CF_ALIVE | CF_NEW)) continue;
    const int k = atomicAdd(&W.hdr->nNewContacts, 1);
    if (peek || k >= cap) continue;
    W.c_flags[i] = flags & ~CF_NEW;
    const int4 ids = W.c_ids[i], fx = W.c_fix[i];
    out[k] = make_int4(fx.x, W.p_ids[ids.x].y, fx.y, W.p_ids[ids.y].y);
    keys[k] = W.c_key[i];
  }
}
// b2Fixture.SetSensor: the contacts of that fixture re-derive their cached sensor bit from the two fixtures
__global__ void __launch_bounds__(256) k_api_resensor(const __grid_constant__ DevWorld W, int fixture) {
  const int n = W.hdr->cHigh;
  GRID_STRIDE(i, n) {
    const uint32_t flags = W.c_flags[i];
    if (!(flags & CF_ALIVE)) continue;
    const int4 fx = W.c_fix[i];
    if (fx.x != fixture && fx.y != fixture) continue;
    const bool sensor = ((W.f_group[fx.x] >> 16) & FXF_SENSOR) || ((W.f_group[fx.y] >> 16) & FXF_SENSOR);
    W.c_flags[i] = sensor ? (flags | CF_SENSOR) : (flags & ~CF_SENSOR);
  }
}
// bulk b2RevoluteJoint / b2PrismaticJoint / b2WheelJoint.SetMotorSpeed: slots index the (colour-sorted) device joint arrays
__global__ void __launch_bounds__(256) k_api_motor_speeds(const __grid_constant__ DevWorld W, const int* slots, const float* speeds, int n) {
  GRID_STRIDE(k, n) {
    const int j = slots[k];
    if (j < 0 || j >= W.nJoints) continue;
    const int4 ids = W.j_ids[j];
    if (ids.x == JT_REVOLUTE || ids.x == 2 /* prismatic */) { float4 p = W.j_p1[j]; p.x = speeds[k]; W.j_p1[j] = p; }
    else if (ids.x == 7 /* wheel */) { float4 p = W.j_p0[j]; p.w = speeds[k]; W.j_p0[j] = p; }
    else continue;
    wake_body_now(W, ids.y); wake_body_now(W, ids.z);          // m_bodyA.SetAwake(true); m_bodyB.SetAwake(true)
  }
}
__global__ void k_api_wake(const __grid_constant__ DevWorld W, int a, int b) {
  if (a >= 0) wake_body_now(W, a);
  if (b >= 0) wake_body_now(W, b);
}

// Bulk b2Body.SetTransform + SetLinearVelocity + SetAngularVelocity (dynamics/b2body.d:261-285, 296-326) for n bodies of a
// (possibly replicated) world: the reset call of a batched-worlds loop.  Pass 1 writes the body state and tags the body,
// pass 2 runs b2Fixture.Synchronize(xf, xf) (zero displacement) for the proxies of tagged bodies, pass 3 removes the tags.
__global__ void __launch_bounds__(256) k_api_set_states(const __grid_constant__ DevWorld W, const int* ids, const float4* pose, const float4* vel, int n) {
  GRID_STRIDE(k, n) {
    const int b = ids ? ids[k] : k;
    if (b < 0 || b >= W.nBodies) continue;
    const uint32_t f = W.b_flags[b];
    if (!(f & BF_ALIVE)) continue;
    if (pose) {
      const float4 ps = pose[k];
      Xf xf; xf.p = V(ps.x, ps.y); xf.q = rot_from_angle(ps.z);
      const float4 lc = W.b_lc[b];
      const v2 c = mul(xf, V(lc.x, lc.y));
      W.b_xf[b] = pack(xf); W.b_xf0[b] = pack(xf);
      W.b_pos[b] = make_float4(c.x, c.y, ps.z, 0.0f);
      W.b_pos0[b] = make_float4(c.x, c.y, ps.z, W.b_pos0[b].w);
      W.b_toiFlags[b] = TF_SYNC;
    }
    if (vel && body_type(f) != BODY_STATIC) {
      const float4 v = vel[k];
      if (v.x * v.x + v.y * v.y > 0.0f || v.z * v.z > 0.0f) wake_body_now(W, b);
      W.b_vel[b] = make_float4(v.x, v.y, v.z, 0.0f);
    }
  }
}
__global__ void __launch_bounds__(256) k_api_sync_tagged(const __grid_constant__ DevWorld W) {
  GRID_STRIDE(p, W.nProxies) {
    const uint32_t pf = W.p_flags[p];
    if (!(pf & PF_ALIVE)) continue;
    const int body = W.p_ids[p].z;
    if (!(W.b_toiFlags[body] & TF_SYNC)) continue;
    sync_proxy(W, p, pf, body);
  }
}
__global__ void __launch_bounds__(256) k_api_untag(const __grid_constant__ DevWorld W, const int* ids, int n) {
  GRID_STRIDE(k, n) { const int b = ids ? ids[k] : k; if (b >= 0 && b < W.nBodies) W.b_toiFlags[b] = 0; }
}

// ------------------------------------------------------------------------------------------------ time of impact
// b2World.SolveTOI (dynamics/b2world.d:1127-1452) + b2Island.SolveTOI (b2island.d:282-416).
// The reference handles one event at a time in ascending alpha.  Events whose mini-islands share no movable body
// commute, so each pass handles, in parallel, every candidate event that holds the minimum (alpha, slot) on all the
// movable bodies it would touch; the others wait for the next pass.  One persistent cooperative kernel runs the whole
// loop (TOI evaluation -> arbitration -> mini-island solve -> SynchronizeFixtures -> FindNewContacts) on the device.
DBX_D float atomic_min_f(float* addr, float v) {
  return v >= 0.0f ? __int_as_float(atomicMin((int*)addr, __float_as_int(v))) : __uint_as_float(atomicMax((unsigned*)addr, __float_as_uint(v)));
}
DBX_D float atomic_max_f(float* addr, float v) {
  return v >= 0.0f ? __int_as_float(atomicMax((int*)addr, __float_as_int(v))) : __uint_as_float(atomicMin((unsigned*)addr, __float_as_uint(v)));
}
DBX_D Sweep load_sweep(const DevWorld& W, int b) {
  const float4 lc = W.b_lc[b], p0 = ldcg4(&W.b_pos0[b]), p = ldcg4(&W.b_pos[b]);
  Sweep s; s.localCenter = V(lc.x, lc.y); s.c0 = V(p0.x, p0.y); s.c = V(p.x, p.y); s.a0 = p0.z; s.a = p.z; s.alpha0 = p0.w;
  return s;
}
struct BodyBackup { float4 pos0, pos, xf, xf0; };
DBX_D BodyBackup backup_body(const DevWorld& W, int b) {
  BodyBackup k; k.pos0 = ldcg4(&W.b_pos0[b]); k.pos = ldcg4(&W.b_pos[b]); k.xf = ldcg4(&W.b_xf[b]); k.xf0 = ldcg4(&W.b_xf0[b]); return k;
}
DBX_D void restore_body(const DevWorld& W, int b, const BodyBackup& k) {   // m_sweep = backup; SynchronizeTransform()
  stcg4(&W.b_pos0[b], k.pos0); stcg4(&W.b_pos[b], k.pos); stcg4(&W.b_xf[b], k.xf); stcg4(&W.b_xf0[b], k.xf0);
}
// b2Body.Advance (b2body.d:1172-1180); statics are left alone (only their alpha0 would change, which nothing reads here)
DBX_D void advance_body(const DevWorld& W, int b, float alpha) {
  if (body_type(W.b_flags[b]) == BODY_STATIC) return;
  Sweep s = load_sweep(W, b);
  sweep_advance(s, alpha);
  s.c = s.c0; s.a = s.a0;
  Xf xf = xf_from_sweep(s.c, s.a, s.localCenter);
  stcg4(&W.b_pos0[b], make_float4(s.c0.x, s.c0.y, s.a0, s.alpha0));
  stcg4(&W.b_pos[b], make_float4(s.c.x, s.c.y, s.a, 0.0f));
  stcg4(&W.b_xf[b], pack(xf));
  stcg4(&W.b_xf0[b], pack(xf));
}

// Between rebuilds the LBVH stays a valid acceleration structure if every moved proxy's new fat box is merged into its
// leaf and all ancestors (boxes only ever grow until the next rebuild); the reported pair set does not depend on it.
DBX_D void lbvh_enlarge(const DevWorld& W, int p) {
  const int n = W.nProxies;
  const float4 f = __ldcg(&W.p_fat[p]);
  int node = n - 1 + W.bv_pos[p];
  __stcg(&W.bv_box[node], f);
  node = W.bv_parent[node];
  while (node >= 0) {
    // boxes only grow and a parent contains its children, so the first ancestor that already contains f ends the walk
    const float4 cur = __ldcg(&W.bv_box[node]);
    if (cur.x <= f.x && cur.y <= f.y && cur.z >= f.z && cur.w >= f.w) break;
    float* bx = (float*)&W.bv_box[node];
    atomic_min_f(bx + 0, f.x); atomic_min_f(bx + 1, f.y); atomic_max_f(bx + 2, f.z); atomic_max_f(bx + 3, f.w);
    node = W.bv_parent[node];
  }
}
__global__ void __launch_bounds__(256) k_lbvh_enlarge(const __grid_constant__ DevWorld W) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { W.hdr->nPairs = 0; W.hdr->nFresh = 0; }   // first kernel of a FindNewContacts without rebuild
  const int nMoved = min(W.hdr->nMoved, W.moveCap);
  GRID_STRIDE(k, nMoved) lbvh_enlarge(W, W.moveList[k]);
}

// event priority: earlier alpha first; exact ties on a shared body are broken by a hash of the replica-local pair key, so
// the choice does not depend on slot numbers (which atomics hand out in a run-dependent order)
DBX_D unsigned long long toi_prio(const DevWorld& W, float alpha, int contact, int bodyOfContact) {
  return ((unsigned long long)__float_as_uint(alpha) << 32) | (unsigned)(mix64(local_key(W, W.c_key[contact], bodyOfContact)) >> 32);
}

// (a) evaluate b2TimeOfImpact for every eligible contact that has no cached value (b2world.d:1155-1265), in two halves so
// that a warp can gather the few contacts that need the long computation and run them on full lanes:
// toi_classify = the cheap filters (returns true when b2TimeOfImpact must run), toi_compute = the computation.
DBX_D void toi_publish(const DevWorld& W, int i, const int4 ids, uint32_t fa, uint32_t fb, float alpha) {
  if (1.0f - 10.0f * kEpsilon < alpha) return;   // never becomes an event (:1267-1272)
  const unsigned long long prio = toi_prio(W, alpha, i, ids.z);
  if (body_type(fa) != BODY_STATIC) atomicMin(&W.b_toiMin[ids.z], prio);
  if (body_type(fb) != BODY_STATIC) atomicMin(&W.b_toiMin[ids.w], prio);
  if (W.subStep) atomicMin(&W.hdr->toiGlobalMin, prio);
}
// `first`: this is the step's first look at the contact, which doubles as the reset of b2world.d:1131-1146 (skipped when a
// sub-stepped world resumes an unfinished SolveTOI, W.toiResume) -- forget the cached TOI, the island flag and the sub-step count
DBX_D bool toi_classify(const DevWorld& W, int i, bool first) {
  uint32_t flags = W.c_flags[i];
  if (!(flags & CF_ALIVE)) return false;
  if (first) {
    if ((flags & CF_FRESH) && W.toiMode == 1) return false;   // the overlapped launch must not look at a contact still being built
    if (flags & (CF_TOI | CF_ISLAND)) { flags &= ~(CF_TOI | CF_ISLAND); W.c_flags[i] = flags; }
    if (W.c_toiCount[i] != 0) W.c_toiCount[i] = 0;
  }
  const int4 ids = W.c_ids[i];
  if ((flags & CF_TOI) && ((__ldcg(&W.b_toiFlags[ids.z]) | __ldcg(&W.b_toiFlags[ids.w])) & TF_INVAL)) { flags &= ~(CF_TOI | CF_ISLAND); W.c_flags[i] = flags; }
  if (!(flags & CF_ENABLED)) return false;
  if (W.c_toiCount[i] > kMaxSubSteps) return false;
  const uint32_t fa = W.b_flags[ids.z], fb = W.b_flags[ids.w];
  if (flags & CF_TOI) { toi_publish(W, i, ids, fa, fb, W.c_mat[i].w); return false; }
  if (flags & CF_SENSOR) return false;
  const int typeA = body_type(fa), typeB = body_type(fb);
  const bool activeA = (fa & BF_AWAKE) && typeA != BODY_STATIC, activeB = (fb & BF_AWAKE) && typeB != BODY_STATIC;
  if (!activeA && !activeB) return false;
  const bool collideA = (fa & BF_BULLET) || typeA != BODY_DYNAMIC, collideB = (fb & BF_BULLET) || typeB != BODY_DYNAMIC;
  return collideA || collideB;
}
DBX_D void toi_compute(const DevWorld& W, int i) {
  const int4 ids = W.c_ids[i];
  const uint32_t fa = W.b_flags[ids.z], fb = W.b_flags[ids.w];
  const int typeA = body_type(fa), typeB = body_type(fb);
  Sweep sA = load_sweep(W, ids.z), sB = load_sweep(W, ids.w);
  // bring both sweeps to the later alpha0 (:1214-1225); done on local copies, see DESIGN.md
  float alpha0 = sA.alpha0;
  if (typeA == BODY_STATIC) { alpha0 = sB.alpha0; sA.alpha0 = alpha0; }
  else if (typeB == BODY_STATIC) { alpha0 = sA.alpha0; sB.alpha0 = alpha0; }
  else if (sA.alpha0 < sB.alpha0) { alpha0 = sB.alpha0; sweep_advance(sA, alpha0); }
  else if (sB.alpha0 < sA.alpha0) { alpha0 = sA.alpha0; sweep_advance(sB, alpha0); }
  const int4 fx = W.c_fix[i];
  DProxy pA = make_proxy(W.shapes + fx.z), pB = make_proxy(W.shapes + fx.w);
  float beta;
  float alpha = 1.0f;
  int state = time_of_impact(&beta, pA, sA, pB, sB, 1.0f);
  if (state == TOI_TOUCHING) alpha = fminr(alpha0 + (1.0f - alpha0) * beta, 1.0f);
  float4 mat = W.c_mat[i]; mat.w = alpha; W.c_mat[i] = mat;
  W.c_flags[i] |= CF_TOI;
  toi_publish(W, i, ids, fa, fb, alpha);
}
// one warp: scan contacts [.., n) in strides of the whole grid, queue the ones that need b2TimeOfImpact in `q` (>= 64 ints of
// shared memory private to the warp) and run them 32 at a time
// the k-th contact the TOI pass looks at: k_collide's list first, then the contacts created since (k_add_pairs' list)
DBX_D int toi_slot(const DevWorld& W, int k, int nList) { return k < nList ? W.c_toiList[k] : W.c_work[k - nList]; }
DBX_D void toi_evaluate_all(const DevWorld& W, int n, int nList, int warp, int nwarps, int lane, int* q, bool first) {
  int qn = 0;
  for (int base = warp * 32; base < n; base += nwarps * 32) {
    const int k = base + lane;
    const int i = k < n ? toi_slot(W, k, nList) : 0;
    const bool need = k < n && toi_classify(W, i, first);
    const unsigned m = __ballot_sync(0xffffffffu, need);
    if (need) q[qn + __popc(m & ((1u << lane) - 1u))] = i;
    qn += __popc(m);
    __syncwarp();
    if (qn >= 32) { qn -= 32; toi_compute(W, q[qn + lane]); __syncwarp(); }
  }
  if (lane < qn) toi_compute(W, q[lane]);
  __syncwarp();
}

// (d) one event, start to finish, by one thread (b2world.d:1274-1440)
DBX_D void toi_process_event(const DevWorld& W, int e, float dtStep) {
  const int i0 = W.e_contact[e];
  const int4 ids0 = W.c_ids[i0];
  const int bA = ids0.z, bB = ids0.w;
  const float minAlpha = W.c_mat[i0].w;
  const unsigned long long prio = toi_prio(W, minAlpha, i0, bA);
  const uint32_t fA = W.b_flags[bA], fB = W.b_flags[bB];
  // arbitration over the movable bodies this event would pull in besides bA/bB
  for (int side = 0; side < 2; ++side) {
    const int nc = min(W.e_ncand[2 * e + side], kToiCand);
    const int body = side == 0 ? bA : bB;
    for (int k = 0; k < nc; ++k) {
      const int4 ids = W.c_ids[W.e_cand[(2 * e + side) * kToiCand + k]];
      const int other = ids.z == body ? ids.w : ids.z;
      if (other == bA || other == bB || body_type(W.b_flags[other]) == BODY_STATIC) continue;
      if (__ldcg(&W.b_toiOther[other]) != prio) return;                          // a better event wants that body: wait
      const int oe = __ldcg(&W.b_toiEvt[other]);
      if (oe >= 0 && oe != e) {
        const int oc = W.e_contact[oe];
        const unsigned long long op = toi_prio(W, W.c_mat[oc].w, oc, W.c_ids[oc].z);
        if (op < prio) return;
      }
    }
  }
  atomicAdd(&W.hdr->toiEvents, 1);
  const BodyBackup backup1 = backup_body(W, bA), backup2 = backup_body(W, bB);
  advance_body(W, bA, minAlpha);
  advance_body(W, bB, minAlpha);
  uint32_t flags0 = update_contact(W, i0, W.c_flags[i0], ids0, W.c_fix[i0], true);
  flags0 &= ~CF_TOI;
  W.c_toiCount[i0] += 1;
  if (!(flags0 & CF_ENABLED) || !(flags0 & CF_TOUCHING)) {
    W.c_flags[i0] = flags0 & ~CF_ENABLED;
    restore_body(W, bA, backup1);
    restore_body(W, bB, backup2);
    return;
  }
  W.c_flags[i0] = flags0;
  atomicAdd(&W.hdr->toiSolved, 1);
  wake_body_now(W, bA);
  wake_body_now(W, bB);
  int bodies[2 * kMaxTOIContacts], contacts[kMaxTOIContacts];
  int nb = 0, nc = 0;
  bodies[nb++] = bA; bodies[nb++] = bB; contacts[nc++] = i0;
  for (int side = 0; side < 2; ++side) {
    const int body = side == 0 ? bA : bB;
    const uint32_t fbody = side == 0 ? fA : fB;
    if (body_type(fbody) != BODY_DYNAMIC) continue;
    const int ncand = min(W.e_ncand[2 * e + side], kToiCand);
    int* cand = W.e_cand + (2 * e + side) * kToiCand;
    // newest first, like the body's contact list: approximated by descending (replica-local) pair key
    for (int a = 1; a < ncand; ++a) {
      const int v = cand[a]; const unsigned long long kv = local_key(W, W.c_key[v], body);
      int b = a - 1;
      while (b >= 0 && local_key(W, W.c_key[cand[b]], body) < kv) { cand[b + 1] = cand[b]; --b; }
      cand[b + 1] = v;
    }
    for (int k = 0; k < ncand; ++k) {
      if (nb == 2 * kMaxTOIContacts) break;
      if (nc == kMaxTOIContacts) break;
      const int ci = cand[k];
      bool already = false;
      for (int t = 0; t < nc; ++t) if (contacts[t] == ci) { already = true; break; }
      if (already) continue;
      const int4 ids = W.c_ids[ci];
      const int other = ids.z == body ? ids.w : ids.z;
      bool otherIn = false;
      for (int t = 0; t < nb; ++t) if (bodies[t] == other) { otherIn = true; break; }
      const BodyBackup backup = backup_body(W, other);
      if (!otherIn) advance_body(W, other, minAlpha);
      const uint32_t fl = update_contact(W, ci, W.c_flags[ci], ids, W.c_fix[ci], true);
      if (!(fl & CF_ENABLED) || !(fl & CF_TOUCHING)) { restore_body(W, other, backup); continue; }
      contacts[nc++] = ci;
      if (otherIn) continue;
      if (body_type(W.b_flags[other]) != BODY_STATIC) wake_body_now(W, other);
      bodies[nb++] = other;
    }
  }
  // ---- b2Island.SolveTOI with subStep {dt = (1 - alpha) dt, 20 position iterations, no warm starting}
  const int sBase = e * kMaxTOIContacts;
  // A small mini-island (nearly all are: a body and what it rests on) keeps its bodies in thread-local arrays while the two
  // solver loops run, the rows referring to them by index (BodyView mode 1): the loops are one thread's chain of dependent
  // steps, and a trip to L2 and back for the bodies inside every row was a third of it.
  constexpr int kToiLocalBodies = 8;
  float4 lpos[kToiLocalBodies], lvel[kToiLocalBodies];
  const bool local = nb <= kToiLocalBodies;
  BodyView view; view.vel = lvel; view.pos = lpos; view.off = 0; view.mode = 1;
  auto rows_to_local = [&]() {
    for (int k = 0; k < nc; ++k) {
      int2 bd = W.s_body[sBase + k];
      int x = 0, y = 0;
      for (int t = 0; t < nb; ++t) { if (bodies[t] == bd.x) x = t; if (bodies[t] == bd.y) y = t; }
      W.s_body[sBase + k] = make_int2(x, y);
    }
  };
  for (int k = 0; k < nc; ++k) prepare_contact(W, sBase + k, contacts[k], -1.0f);
  if (local) {
    for (int t = 0; t < nb; ++t) lpos[t] = ldcg4(&W.b_pos[bodies[t]]);
    rows_to_local();
    for (int it = 0; it < 20; ++it) {
      float minSeparation = 0.0f;
      for (int k = 0; k < nc; ++k) minSeparation = fminr(minSeparation, contact_solve_position(W, sBase + k, 0, 1, view));   // (bodies[0] = bA, bodies[1] = bB)
      if (minSeparation >= -1.5f * kLinearSlop) break;
    }
    for (int t = 0; t < nb; ++t) if (body_type(W.b_flags[bodies[t]]) == BODY_DYNAMIC) stcg4(&W.b_pos[bodies[t]], lpos[t]);
  } else for (int it = 0; it < 20; ++it) {
    float minSeparation = 0.0f;
    for (int k = 0; k < nc; ++k) minSeparation = fminr(minSeparation, contact_solve_position(W, sBase + k, bA, bB));
    if (minSeparation >= -1.5f * kLinearSlop) break;
  }
  // leap of faith: the TOI bodies' c0/a0 become the solved pose (b2island.d:352-355); refresh transforms for the velocity pass
  for (int t = 0; t < nb; ++t) {
    const int b = bodies[t];
    if (body_type(W.b_flags[b]) == BODY_STATIC) continue;
    const float4 pos = ldcg4(&W.b_pos[b]); const float4 lc = W.b_lc[b];
    const Xf xf = xf_from_sweep(V(pos.x, pos.y), pos.z, V(lc.x, lc.y));
    stcg4(&W.b_xf[b], pack(xf));
    if (b == bA || b == bB) {
      float4 p0 = ldcg4(&W.b_pos0[b]); p0.x = pos.x; p0.y = pos.y; p0.z = pos.z;
      stcg4(&W.b_pos0[b], p0);
      stcg4(&W.b_xf0[b], pack(xf));
    }
  }
  for (int k = 0; k < nc; ++k) prepare_contact(W, sBase + k, contacts[k], -1.0f);
  if (local) {
    for (int t = 0; t < nb; ++t) lvel[t] = ldcg4(&W.b_vel[bodies[t]]);
    rows_to_local();
    for (int it = 0; it < W.velIters; ++it) for (int k = 0; k < nc; ++k) contact_solve_velocity(W, sBase + k, view);
    for (int t = 0; t < nb; ++t) if (body_type(W.b_flags[bodies[t]]) == BODY_DYNAMIC) stcg4(&W.b_vel[bodies[t]], lvel[t]);
  } else for (int it = 0; it < W.velIters; ++it) for (int k = 0; k < nc; ++k) contact_solve_velocity(W, sBase + k);
  if (W.psCap > 0) for (int k = 0; k < nc; ++k) emit_post_solve(W, 2, sBase + k, contacts[k]);   // island.Report (b2island.d:414)
  const float h = (1.0f - minAlpha) * dtStep;
  for (int t = 0; t < nb; ++t) {
    const int b = bodies[t];
    const int type = body_type(W.b_flags[b]);
    if (type == BODY_STATIC) continue;
    float4 pos = ldcg4(&W.b_pos[b]), vel = ldcg4(&W.b_vel[b]);
    v2 c = V(pos.x, pos.y), v = V(vel.x, vel.y);
    float a = pos.z, w = vel.z;
    v2 translation = h * v;
    if (dot(translation, translation) > kMaxTranslationSquared) { float ratio = kMaxTranslation / len(translation); v *= ratio; }
    float rotation = h * w;
    if (rotation * rotation > kMaxRotationSquared) { float ratio = kMaxRotation / fabsr(rotation); w *= ratio; }
    c += h * v;
    a += h * w;
    stcg4(&W.b_pos[b], make_float4(c.x, c.y, a, 0.0f));
    stcg4(&W.b_vel[b], make_float4(v.x, v.y, w, 0.0f));
    const float4 lc = W.b_lc[b];
    stcg4(&W.b_xf[b], pack(xf_from_sweep(c, a, V(lc.x, lc.y))));
    if (type == BODY_DYNAMIC) atomicOr(&W.b_toiFlags[b], TF_INVAL | TF_SYNC);   // :1423-1440
  }
}

__global__ void __launch_bounds__(512) k_toi(const __grid_constant__ DevWorld W) {
  __shared__ int stacks[16][kQueryStackToi];
  Header* H = W.hdr;
  const unsigned nb = gridDim.x;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warp = tid >> 5, nwarps = nth >> 5;
  const int eventCap = W.eventCap;
  int tmark = 0;
#define TMARK() do { if (W.phaseTimes && tid == 0 && tmark < 64) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); W.phaseTimes[3000 + tmark++] = t_; } } while (0)
  TMARK();
  // The per-body TOI scratch is left clean by the previous launch (see the end of this kernel) and the per-contact part
  // of the reset (:1131-1146) rides on the first classification; only a world whose scratch may be dirty (first step,
  // bodies added since) pays for a reset phase.
  if (W.toiReset) {
    for (int b = tid; b < W.nBodies; b += nth) {
      if (!W.toiResume) { float4 p0 = W.b_pos0[b]; if (p0.w != 0.0f) { p0.w = 0.0f; W.b_pos0[b] = p0; } }
      W.b_toiMin[b] = ~0ull; W.b_toiOther[b] = ~0ull; W.b_toiEvt[b] = -1; W.b_toiFlags[b] = W.toiResume ? (W.b_toiFlags[b] & TF_INVAL) : 0;
    }
    if (tid == 0) H->nEvents = 0;
    grid_barrier(&H->barrier, nb); TMARK();
  }
  if (W.toiClearMoves && W.toiMode == 0) {   // tail of the step's FindNewContacts (b2broadphase.d:190-191): forget the move buffer
    const int nMoved = min(H->nMoved, W.moveCap);
    for (int k = tid; k < nMoved; k += nth) { const int p = W.moveList[k]; W.p_flags[p] &= ~PF_MOVED; }
  }
  bool incomplete = false;
  for (int pass = 0; pass < 1024; ++pass) {
    // (a) TOI evaluation + per-body minima
    {
      // few contacts per thread (one big world): evaluate in place, every chain on its own warp; many (batched worlds):
      // gather the eligible ones so that b2TimeOfImpact runs on full warps
      const bool first = pass == 0 && (!W.toiPre || W.toiMode == 1) && !W.toiResume;
      const int nFresh = min(*((volatile int*)&H->nFresh), W.cCap);
      const int nList = min(*((volatile int*)&H->nToi), W.cCap);
      const int n = nList + (W.toiMode == 1 ? 0 : nFresh);     // the overlapped first evaluation leaves the fresh ones alone
      if (pass == 0 && W.toiPre && W.toiMode == 0 && nFresh <= W.cCap) {
        // k_toi_pre has evaluated and published everything except the contacts FindNewContacts created meanwhile
        // (one per warp while they last: a b2TimeOfImpact chain is long and divergent)
        const int stride = nFresh * 32 <= nth ? 32 : 1;
        if (tid % stride == 0) for (int k = tid / stride; k < nFresh; k += nth / stride) { const int i = W.c_work[k]; if (toi_classify(W, i, false)) toi_compute(W, i); }
      } else if (n <= 8 * nth) { for (int k = tid; k < n; k += nth) { const int i = toi_slot(W, k, nList); if (toi_classify(W, i, first)) toi_compute(W, i); } }
      else toi_evaluate_all(W, n, nList, warp, nwarps, lane, stacks[wib], first);
    }
    // The overlapped first evaluation is this same kernel (same code, warm instruction caches) launched as a plain grid
    // on the second stream: it stops here, before any grid barrier.  Contacts FindNewContacts creates meanwhile carry
    // CF_FRESH and are left to the real launch.
    if (W.toiMode == 1) return;
    grid_barrier(&H->barrier, nb); TMARK();
    if (pass == 0 && W.toiClearMoves && tid == 0) H->nMoved = 0;   // every CTA has read it; next used three barriers on
    // (b) winners: the minimum on every movable body they touch
    {
      const int nList = min(*((volatile int*)&H->nToi), W.cCap);
      const int n = nList + min(*((volatile int*)&H->nFresh), W.cCap);
      for (int k = tid; k < n; k += nth) {
        const int i = toi_slot(W, k, nList);
        const uint32_t flags = W.c_flags[i];
        if ((flags & (CF_ALIVE | CF_ENABLED | CF_TOI)) != (CF_ALIVE | CF_ENABLED | CF_TOI)) continue;
        if (W.c_toiCount[i] > kMaxSubSteps) continue;
        const float alpha = W.c_mat[i].w;
        if (1.0f - 10.0f * kEpsilon < alpha) continue;
        const int4 ids = W.c_ids[i];
        const unsigned long long prio = toi_prio(W, alpha, i, ids.z);
        const bool movA = body_type(W.b_flags[ids.z]) != BODY_STATIC, movB = body_type(W.b_flags[ids.w]) != BODY_STATIC;
        if ((movA && __ldcg(&W.b_toiMin[ids.z]) != prio) || (movB && __ldcg(&W.b_toiMin[ids.w]) != prio)) continue;
        if (W.subStep && __ldcg(&H->toiGlobalMin) != prio) continue;     // one event per Step: the earliest of all
        const int e = atomicAdd(&H->nEvents, 1);
        if (e >= eventCap) continue;     // stays a candidate for the next pass
        W.e_contact[e] = i;
        W.e_ncand[2 * e] = 0; W.e_ncand[2 * e + 1] = 0;
        if (movA) W.b_toiEvt[ids.z] = e;
        if (movB) W.b_toiEvt[ids.w] = e;
      }
      for (int b = tid; b < W.nBodies; b += nth) if (W.b_toiFlags[b] & TF_INVAL) W.b_toiFlags[b] &= ~TF_INVAL;
    }
    grid_barrier(&H->barrier, nb); TMARK();
    const int nEvents = min(*((volatile int*)&H->nEvents), eventCap);
    if (nEvents == 0) break;
    // (c) contacts of the event bodies that may join their mini-islands (b2world.d:1319-1411)
    {
      const int nList = min(*((volatile int*)&H->nToi), W.cCap);
      const int n = nList + min(*((volatile int*)&H->nFresh), W.cCap);
      for (int k = tid; k < n; k += nth) {
        const int i = toi_slot(W, k, nList);
        const uint32_t flags = W.c_flags[i];
        if (!(flags & CF_ALIVE) || (flags & CF_SENSOR)) continue;
        const int4 ids = W.c_ids[i];
        for (int side = 0; side < 2; ++side) {
          const int body = side == 0 ? ids.z : ids.w, other = side == 0 ? ids.w : ids.z;
          const int e = __ldcg(&W.b_toiEvt[body]);
          if (e < 0 || e >= nEvents) continue;
          const int ec = W.e_contact[e];
          if (ec == i) continue;
          const uint32_t fbody = W.b_flags[body], fother = W.b_flags[other];
          if (body_type(fbody) != BODY_DYNAMIC) continue;
          if (body_type(fother) == BODY_DYNAMIC && !(fbody & BF_BULLET) && !(fother & BF_BULLET)) continue;
          const int evSide = W.c_ids[ec].z == body ? 0 : 1;
          const int slot = atomicAdd(&W.e_ncand[2 * e + evSide], 1);
          if (slot < kToiCand) W.e_cand[(2 * e + evSide) * kToiCand + slot] = i; else H->error = E_TOI_CANDIDATES;
          if (body_type(fother) != BODY_STATIC) {
            const unsigned long long prio = toi_prio(W, W.c_mat[ec].w, ec, W.c_ids[ec].z);
            atomicMin(&W.b_toiOther[other], prio);
          }
        }
      }
    }
    grid_barrier(&H->barrier, nb); TMARK();
    // (d) events
    // one thread per event: a pass with few events (one big world) gives each its own warp, a pass with many (batched
    // worlds) fills the lanes
    for (int e = lane * nwarps + warp; e < nEvents; e += nth) toi_process_event(W, e, W.dt);
    grid_barrier(&H->barrier, nb); TMARK();
    // (e) SynchronizeFixtures of the island's dynamic bodies (:1433), then FindNewContacts (:1444)
    for (int p = tid; p < W.nProxies; p += nth) {
      const uint32_t pf = W.p_flags[p];
      if (!(pf & PF_ALIVE)) continue;
      const int body = W.p_ids[p].z;
      if (!(__ldcg(&W.b_toiFlags[body]) & TF_SYNC)) continue;
      sync_proxy(W, p, pf, body);
    }
    if (tid == 0) H->nPairs = 0;
    grid_barrier(&H->barrier, nb); TMARK();
    // (a TOI sub-step rarely carries a proxy out of its fat box: without moved proxies FindNewContacts has nothing to do, and
    // its two grid barriers are skipped.  The count is final here and every CTA reads the same value.)
    const int nMovedNow = min(*((volatile int*)&H->nMoved), W.moveCap);
    {
      // the step's LBVH is still valid for every proxy that did not move; widen it for the ones that did
      for (int k = tid; k < nMovedNow; k += nth) lbvh_enlarge(W, W.moveList[k]);
      for (int b = tid; b < W.nBodies; b += nth) {
        W.b_toiMin[b] = ~0ull; W.b_toiOther[b] = ~0ull; W.b_toiEvt[b] = -1;
        if (W.b_toiFlags[b] & TF_SYNC) W.b_toiFlags[b] &= ~TF_SYNC;
      }
      if (tid == 0) { H->nEvents = 0; H->toiGlobalMin = ~0ull; }
    }
    grid_barrier(&H->barrier, nb); TMARK();
    if (nMovedNow > 0) {
      for (int k = warp; k < nMovedNow; k += nwarps) query_proxy(W, W.bv_sorted, stacks[wib], kQueryStackToi, lane, W.moveList[k]);
      grid_barrier(&H->barrier, nb); TMARK();
      const int nPairs = min(*((volatile int*)&H->nPairs), W.pairCap);
      for (int k = tid; k < nPairs; k += nth) add_pair(W, W.pairs[k]);
      for (int k = tid; k < nMovedNow; k += nth) { const int p = W.moveList[k]; W.p_flags[p] &= ~PF_MOVED; }
      grid_barrier(&H->barrier, nb); TMARK();
      if (tid == 0) H->nMoved = 0;      // (next touched by the SynchronizeFixtures of the next pass, four barriers on)
    }
    const bool stopHere = W.subStep && *((volatile int*)&H->toiSolved) > 0;   // b2world.d:1441-1446: m_stepComplete = false; break
    if (W.subStep) { grid_barrier(&H->barrier, nb); TMARK(); }
    if (stopHere) { incomplete = true; break; }
  }
  // leave the per-body scratch clean for the next step (every CTA is past the last arbitration phase here); a sub-stepped
  // world that stopped after one event keeps what the resumed SolveTOI needs: the sweeps' alpha0 and the pending invalidations
  for (int b = tid; b < W.nBodies; b += nth) {
    if (W.b_toiMin[b] != ~0ull) W.b_toiMin[b] = ~0ull;
    if (W.b_toiOther[b] != ~0ull) W.b_toiOther[b] = ~0ull;
    if (W.b_toiEvt[b] != -1) W.b_toiEvt[b] = -1;
    if (incomplete) { const int f = W.b_toiFlags[b]; if (f & ~TF_INVAL) W.b_toiFlags[b] = f & TF_INVAL; }
    else {
      if (W.b_toiFlags[b] != 0) W.b_toiFlags[b] = 0;
      float4 p0 = W.b_pos0[b]; if (p0.w != 0.0f) { p0.w = 0.0f; W.b_pos0[b] = p0; }
    }
    if (W.toiClearForces) W.b_force[b] = make_float4(0, 0, 0, 0);
  }
  if (tid == 0) { H->nEvents = 0; H->nToi = 0; H->nFresh = 0; H->stepIncomplete = incomplete ? 1 : 0; H->toiSolved = 0; H->toiGlobalMin = ~0ull; }
}

#undef TMARK
// ------------------------------------------------------------------------------------------------ host launchers
#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

// The barrier is a ticket counter: a barrier is complete when the counter reaches the next multiple of the CTA count, so
// it needs no reset between launches as long as every launch uses the same grid and leaves it on a multiple (all do).
// It is zeroed every 4096 launches, long before 2^32 tickets.
// The barrier word only ever counts up inside a launch and is a multiple of the CTA count between launches.  Worst case per
// launch is kMaxColours phases x 11 passes x 148 CTAs = 1.7 M tickets (k_toi: 1024 passes x ~10 barriers x 148 = 1.5 M), so a
// reset every 1024 cooperative launches keeps it below 2^31 with room to spare; grid_barrier compares wrap-safely as well.
static cudaError_t launch_coop(const void* fn, const DevWorld& W, const LaunchCfg& L) {
  if ((L.coopLaunches++ & 1023) == 0) CK(cudaMemsetAsync(&W.hdr->barrier, 0, sizeof(unsigned), L.stream));
  void* args[] = {(void*)&W};
  ++L.launches;
  return cudaLaunchCooperativeKernel(fn, dim3(L.coopBlocks), dim3(L.coopThreads), args, 0, L.stream);
}

cudaError_t stage_collide(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_collide<<<L.gridWide, 256, 0, L.stream>>>(W);   // appends to the TOI list; k_toi empties it on its way out
  return cudaGetLastError();
}

cudaError_t stage_islands_and_integrate(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_island_init<<<L.gridWide, 256, 0, L.stream>>>(W);
  ++L.launches; k_island_union<<<L.gridWide, 256, 0, L.stream>>>(W);
  ++L.launches; k_island_flatten<<<L.gridWide, 256, 0, L.stream>>>(W);
  ++L.launches; k_island_wake_integrate<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}

// ---- world-major solver order for the world-local solver (dbx_solve.cu): key = replica << 10 | colour over every contact
// slot (slots not in the solver sort to the end), CUB radix sort, then the slot range of every replica
// The sort key keeps only `colourBits` of the colour when the host knows (from the last header it saw) that no colour
// needs more: fewer radix passes.  A colour that does not fit raises the sticky error instead of corrupting the order.
__global__ void __launch_bounds__(256) k_world_keys(const __grid_constant__ DevWorld W, unsigned* keys, int* vals, int n, int colourBits) {
  const int high = W.hdr->cHigh;
  int count = 0, maxc = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0 && high > n) W.hdr->error = E_SOLVER_ROWS;
  GRID_STRIDE(i, n) {
    unsigned key = (unsigned)W.nWorlds << colourBits;
    if (i < high && (W.c_flags[i] & CF_SOLVE)) {
      const int4 ids = W.c_ids[i];
      const int c = W.c_colour[i];
      if (c >= (1 << colourBits)) W.hdr->error = E_COLOURS;
      key = ((unsigned)W.b_world[ids.z] << colourBits) | (unsigned)(c & ((1 << colourBits) - 1));
      ++count; maxc = max(maxc, c + 1);
    }
    keys[i] = key; vals[i] = i;
  }
  for (int o = 16; o > 0; o >>= 1) { count += __shfl_xor_sync(0xffffffffu, count, o); maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o)); }
  if ((threadIdx.x & 31) == 0 && count > 0) { atomicAdd(&W.hdr->nSolve, count); atomicMax(&W.hdr->nColours, maxc); }
}
__global__ void __launch_bounds__(256) k_world_ranges(const __grid_constant__ DevWorld W, const unsigned* keys, int colourBits) {
  const int n = min(W.hdr->nSolve, W.sCap);
  if (blockIdx.x == 0 && threadIdx.x == 0 && W.hdr->nSolve > W.sCap) W.hdr->error = E_SOLVER_ROWS;
  GRID_STRIDE(s, n) {
    const int w = (int)(keys[s] >> colourBits);
    if (s == 0 || (int)(keys[s - 1] >> colourBits) != w) W.w_start[w] = s;
    if (s == n - 1 || (int)(keys[s + 1] >> colourBits) != w) W.w_end[w] = s + 1;
  }
}
cudaError_t launch_mark_and_colour(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_mark_solve<<<L.gridWide, 256, 0, L.stream>>>(W);
  CK(cudaGetLastError());
  if (!W.colourOverride) CK(launch_coop((const void*)k_colour, W, L));
  return cudaGetLastError();
}
// keysA/B, valsA/B: n entries each; on return W.s_contact / W.sw_key point at the sorted buffers
cudaError_t stage_colour_and_sort_worlds(DevWorld& W, const LaunchCfg& L, unsigned* keysA, unsigned* keysB, int* valsA, int* valsB, int n, int colourBits) {
  CK(cudaMemsetAsync(W.w_start, 0, (size_t)W.nWorlds * 4, L.stream));
  CK(cudaMemsetAsync(W.w_end, 0, (size_t)W.nWorlds * 4, L.stream));
  ++L.launches; k_world_keys<<<L.gridWide, 256, 0, L.stream>>>(W, keysA, valsA, n, colourBits);
  cub::DoubleBuffer<unsigned> keys(keysA, keysB);
  cub::DoubleBuffer<int> vals(valsA, valsB);
  int worldBits = 1;
  while ((1 << worldBits) < W.nWorlds + 1) ++worldBits;
  size_t bytes = L.cubTempBytes;
  CK(cub::DeviceRadixSort::SortPairs(L.cubTemp, bytes, keys, vals, n, 0, colourBits + worldBits, L.stream));
  W.s_contact = vals.Current();
  W.sw_key = keys.Current();
  W.swColourBits = colourBits;
  ++L.launches; k_world_ranges<<<L.gridWide, 256, 0, L.stream>>>(W, keys.Current(), colourBits);
  return cudaGetLastError();
}
size_t cub_temp_bytes_u32(int n) {
  size_t bytes = 0;
  cub::DoubleBuffer<unsigned> k(nullptr, nullptr);
  cub::DoubleBuffer<int> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, n, 0, 32);
  return bytes + 256;
}

cudaError_t stage_colour_and_sort(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_mark_solve<<<L.gridWide, 256, 0, L.stream>>>(W);
  CK(cudaGetLastError());
  if (!W.colourOverride) CK(launch_coop((const void*)k_colour, W, L));
  ++L.launches; k_sort_hist<<<kSortBlocks, 256, 0, L.stream>>>(W);
  ++L.launches; k_sort_scan_blocks<<<kMaxColours / 8, 256, 0, L.stream>>>(W);
  ++L.launches; k_sort_scan_colours<<<1, kMaxColours, 0, L.stream>>>(W);
  ++L.launches; k_sort_scatter<<<kSortBlocks, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}


cudaError_t stage_sync_fixtures(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_sync_fixtures<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}

size_t cub_temp_bytes(int maxProxies) {
  size_t bytes = 0;
  cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
  cub::DoubleBuffer<int> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, maxProxies, 0, 64);
  return bytes + 256;
}

cudaError_t stage_toi(DevWorld& W, const LaunchCfg& L) {
  return launch_coop((const void*)k_toi, W, L);
}
cudaError_t stage_toi_pre(const DevWorld& W, const LaunchCfg& L, cudaStream_t aux) {
  DevWorld P = W;
  P.toiMode = 1; P.toiReset = 0; P.toiPre = 0; P.phaseTimes = nullptr;
  // three quarters of the SMs: at 125 registers x 512 threads a CTA owns a whole register file, and the broadphase kernels this
  // overlaps with need somewhere to run
  ++L.launches; k_toi<<<(3 * L.coopBlocks + 3) / 4, L.coopThreads, 0, aux>>>(P);
  return cudaGetLastError();
}

// (re)build the LBVH alone, or widen it for the proxies in the move buffer: what a world query needs before it can run
cudaError_t stage_refresh_tree(DevWorld& W, const LaunchCfg& L, bool rebuild) {
  const int n = W.nProxies;
  if (n == 0) return cudaSuccess;
  if (!rebuild) { ++L.launches; k_lbvh_enlarge<<<L.gridWide, 256, 0, L.stream>>>(W); return cudaGetLastError(); }
  ++L.launches; k_bounds_init<<<1, 1, 0, L.stream>>>(W);
  ++L.launches; k_bounds<<<L.gridWide, 256, 0, L.stream>>>(W);
  ++L.launches; k_morton<<<L.gridWide, 256, 0, L.stream>>>(W);
  cub::DoubleBuffer<unsigned long long> keys(W.bv_key, W.bv_keyAlt);
  cub::DoubleBuffer<int> vals(W.bv_leaf, W.bv_leafAlt);
  int worldBits = 1;
  while ((1 << worldBits) < W.nWorlds + 1) ++worldBits;
  size_t bytes = L.cubTempBytes;
  CK(cub::DeviceRadixSort::SortPairs(L.cubTemp, bytes, keys, vals, n, 0, 30 + worldBits + 1, L.stream));
  W.bv_sorted = vals.Current();
  ++L.launches; k_lbvh_hierarchy<<<L.gridWide, 256, 0, L.stream>>>(W, keys.Current());
  ++L.launches; k_lbvh_refit<<<L.gridWide, 256, 0, L.stream>>>(W, W.bv_sorted);
  return cudaGetLastError();
}
cudaError_t launch_raycast(const DevWorld& W, const LaunchCfg& L, const float4* rays, int n, float4* out) {
  ++L.launches; k_raycast<<<(n + 127) / 128, 128, 0, L.stream>>>(W, W.bv_sorted, rays, n, out);
  return cudaGetLastError();
}
cudaError_t launch_query_aabb(const DevWorld& W, const LaunchCfg& L, const float4* boxes, int n, int capPer, int* counts, int2* out) {
  ++L.launches; k_query_aabb<<<(n + 127) / 128, 128, 0, L.stream>>>(W, W.bv_sorted, boxes, n, capPer, counts, out);
  return cudaGetLastError();
}

cudaError_t launch_raycast_all(const DevWorld& W, const LaunchCfg& L, const float4* rays, int n, int capPer, int* counts, float4* out) {
  ++L.launches; k_raycast_all<<<(n + 127) / 128, 128, 0, L.stream>>>(W, W.bv_sorted, rays, n, capPer, counts, out);
  return cudaGetLastError();
}
cudaError_t launch_test_points(const DevWorld& W, const LaunchCfg& L, const float4* q, int n, int* inside) {
  ++L.launches; k_test_points<<<(n + 127) / 128, 128, 0, L.stream>>>(W, q, n, inside);
  return cudaGetLastError();
}
cudaError_t launch_shift_origin(const DevWorld& W, const LaunchCfg& L, float ox, float oy) {
  ++L.launches; k_shift_origin<<<L.gridWide, 256, 0, L.stream>>>(W, ox, oy);
  return cudaGetLastError();
}
cudaError_t launch_world_manifolds(const DevWorld& W, const LaunchCfg& L, int high, float4* out) {
  ++L.launches; k_world_manifolds<<<L.gridWide, 256, 0, L.stream>>>(W, high, out);
  return cudaGetLastError();
}
cudaError_t launch_post_solve(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_post_solve<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}

cudaError_t stage_find_new_contacts(DevWorld& W, const LaunchCfg& L, bool rebuild, bool deferClear) {
  const int n = W.nProxies;
  if (n == 0 || rebuild) { ++L.launches; k_bounds_init<<<1, 1, 0, L.stream>>>(W); }
  if (n > 0 && !rebuild) {
    ++L.launches; k_lbvh_enlarge<<<L.gridWide, 256, 0, L.stream>>>(W);
    ++L.launches; k_query<<<L.gridWide, 256, 0, L.stream>>>(W, W.bv_sorted);
    ++L.launches; k_add_pairs<<<L.gridWide, 256, 0, L.stream>>>(W);
  } else if (n > 0) {
    ++L.launches; k_bounds<<<L.gridWide, 256, 0, L.stream>>>(W);
    ++L.launches; k_morton<<<L.gridWide, 256, 0, L.stream>>>(W);
    cub::DoubleBuffer<unsigned long long> keys(W.bv_key, W.bv_keyAlt);
    cub::DoubleBuffer<int> vals(W.bv_leaf, W.bv_leafAlt);
    int worldBits = 1;
    while ((1 << worldBits) < W.nWorlds + 1) ++worldBits;
    size_t bytes = L.cubTempBytes;
    CK(cub::DeviceRadixSort::SortPairs(L.cubTemp, bytes, keys, vals, n, 0, 30 + worldBits + 1, L.stream));
    const unsigned long long* sk = keys.Current();
    const int* sl = vals.Current();
    W.bv_sorted = sl;
    if (n > 1) ++L.launches; k_lbvh_hierarchy<<<L.gridWide, 256, 0, L.stream>>>(W, sk);
    ++L.launches; k_lbvh_refit<<<L.gridWide, 256, 0, L.stream>>>(W, sl);
    ++L.launches; k_query<<<L.gridWide, 256, 0, L.stream>>>(W, sl);
    ++L.launches; k_add_pairs<<<L.gridWide, 256, 0, L.stream>>>(W);
  }
  if (!deferClear) {   // otherwise k_toi, which follows, empties the move buffer
    ++L.launches; k_clear_moves<<<L.gridWide, 256, 0, L.stream>>>(W);
    ++L.launches; k_reset_moves<<<1, 1, 0, L.stream>>>(W);
  }
  return cudaGetLastError();
}

cudaError_t stage_rebuild_hash(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_hash_clear<<<L.gridWide, 256, 0, L.stream>>>(W);
  ++L.launches; k_hash_fill<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}

cudaError_t stage_count(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_count_reset<<<1, 1, 0, L.stream>>>(W);
  ++L.launches; k_count<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}

cudaError_t launch_insert_contacts(const DevWorld& W, const LaunchCfg& L, int n) {
  ++L.launches; k_import_reset<<<1, 1, 0, L.stream>>>(W, n);
  return stage_rebuild_hash(W, L);
}

cudaError_t launch_api_contacts(const DevWorld& W, const LaunchCfg& L, int body, int fixture, int otherBody, int flagOnly) {
  ++L.launches; k_api_contacts<<<L.gridWide, 256, 0, L.stream>>>(W, body, fixture, otherBody, flagOnly);
  return cudaGetLastError();
}
cudaError_t launch_patch_contacts(const DevWorld& W, const LaunchCfg& L, const unsigned long long* keys, const float4* vals, const int* masks, int n) {
  ++L.launches; k_patch_contacts<<<(n + 255) / 256, 256, 0, L.stream>>>(W, keys, vals, masks, n);
  return cudaGetLastError();
}
cudaError_t launch_list_new_contacts(const DevWorld& W, const LaunchCfg& L, int4* out, unsigned long long* keys, int cap, int peek) {
  ++L.launches; k_list_new_contacts<<<L.gridWide, 256, 0, L.stream>>>(W, out, keys, cap, peek);
  return cudaGetLastError();
}
cudaError_t launch_api_resensor(const DevWorld& W, const LaunchCfg& L, int fixture) {
  ++L.launches; k_api_resensor<<<L.gridWide, 256, 0, L.stream>>>(W, fixture);
  return cudaGetLastError();
}
cudaError_t launch_set_motor_speeds(const DevWorld& W, const LaunchCfg& L, const int* slots, const float* speeds, int n) {
  ++L.launches; k_api_motor_speeds<<<(n + 255) / 256, 256, 0, L.stream>>>(W, slots, speeds, n);
  return cudaGetLastError();
}
// Row-granular body sync (World::pullBodyRow / push of edited rows): n bodies' rows as nine float4 each -- xf, xf0, pos, pos0, vel,
// force, mass, lc, (gravityScale, sleepTime, flags bits, -) -- so that a per-body edit costs one small copy each way, not ten
constexpr int kBodyRowQuads = 9;
__global__ void __launch_bounds__(64) k_body_rows_get(const __grid_constant__ DevWorld W, const int* ids, float4* rows, int n) {
  GRID_STRIDE(k, n) {
    const int b = ids[k];
    float4* r = rows + (size_t)k * kBodyRowQuads;
    const float2 gs = W.b_gs[b];
    r[0] = W.b_xf[b]; r[1] = W.b_xf0[b]; r[2] = W.b_pos[b]; r[3] = W.b_pos0[b]; r[4] = W.b_vel[b]; r[5] = W.b_force[b]; r[6] = W.b_mass[b]; r[7] = W.b_lc[b];
    r[8] = make_float4(gs.x, gs.y, __uint_as_float(W.b_flags[b]), 0.0f);
  }
}
__global__ void __launch_bounds__(64) k_body_rows_set(const __grid_constant__ DevWorld W, const int* ids, const float4* rows, int n) {
  GRID_STRIDE(k, n) {
    const int b = ids[k];
    const float4* r = rows + (size_t)k * kBodyRowQuads;
    W.b_xf[b] = r[0]; W.b_xf0[b] = r[1]; W.b_pos[b] = r[2]; W.b_pos0[b] = r[3]; W.b_vel[b] = r[4]; W.b_force[b] = r[5]; W.b_mass[b] = r[6]; W.b_lc[b] = r[7];
    W.b_gs[b] = make_float2(r[8].x, r[8].y); W.b_flags[b] = __float_as_uint(r[8].z);
  }
}
cudaError_t launch_body_rows(const DevWorld& W, const LaunchCfg& L, const int* ids, float4* rows, int n, bool set) {
  ++L.launches;
  if (set) k_body_rows_set<<<(n + 63) / 64, 64, 0, L.stream>>>(W, ids, rows, n);
  else k_body_rows_get<<<(n + 63) / 64, 64, 0, L.stream>>>(W, ids, rows, n);
  return cudaGetLastError();
}
cudaError_t launch_api_wake(const DevWorld& W, const LaunchCfg& L, int a, int b) {
  ++L.launches; k_api_wake<<<1, 1, 0, L.stream>>>(W, a, b);
  return cudaGetLastError();
}
cudaError_t launch_set_states(const DevWorld& W, const LaunchCfg& L, const int* ids, const float4* pose, const float4* vel, int n) {
  ++L.launches; k_api_set_states<<<L.gridWide, 256, 0, L.stream>>>(W, ids, pose, vel, n);
  if (pose) {
    ++L.launches; k_api_sync_tagged<<<L.gridWide, 256, 0, L.stream>>>(W);
    ++L.launches; k_api_untag<<<L.gridWide, 256, 0, L.stream>>>(W, ids, n);
  }
  return cudaGetLastError();
}
cudaError_t launch_apply_forces3(const DevWorld& W, const LaunchCfg& L, const float* forces, int n) {
  ++L.launches; k_apply_forces3<<<L.gridWide, 256, 0, L.stream>>>(W, forces, n);
  return cudaGetLastError();
}
cudaError_t launch_pack_poses(const DevWorld& W, const LaunchCfg& L, float* out, int n) {
  ++L.launches; k_pack_poses<<<L.gridWide, 256, 0, L.stream>>>(W, out, n);
  return cudaGetLastError();
}
cudaError_t launch_apply_forces(const DevWorld& W, const LaunchCfg& L, const float4* forces, int n) {
  ++L.launches; k_apply_forces<<<L.gridWide, 256, 0, L.stream>>>(W, forces, n);
  return cudaGetLastError();
}
cudaError_t launch_clear_forces(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_clear_forces<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}

cudaError_t launch_replicate(const DevWorld& W, const LaunchCfg& L, int nB, int nF, int nP, int nMoved, int keyStride, int copies) {
  ++L.launches; k_replicate<<<L.gridWide, 256, 0, L.stream>>>(W, nB, nF, nP, nMoved, keyStride, copies);
  return cudaGetLastError();
}
cudaError_t launch_set_levels(const DevWorld& W, const LaunchCfg& L, const int* d_levels, int n) {
  ++L.launches; k_set_levels<<<L.gridWide, 256, 0, L.stream>>>(W, d_levels, n);
  return cudaGetLastError();
}

}  // namespace dbx
