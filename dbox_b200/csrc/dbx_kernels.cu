// dbx_kernels.cu — hand-written CUDA kernels (sm_100a) for dbox's per-step world pipeline.
//
// Every kernel is a grid-stride loop over a device-resident count (Header), launched with a fixed grid that is a
// multiple of the SM count, so a whole step is a fixed launch sequence with no host round trip.  The iteration loops
// of the constraint solver run inside ONE persistent cooperative kernel (one CTA per SM) with a global barrier per
// colour.  Nothing here is a dense contraction: the bound is HBM/L2 bandwidth and dependent-launch latency, so the
// levers are coalesced float4 SoA access, L2-resident constraint blocks and as few barriers as the colouring allows.
#include <cub/cub.cuh>
#include "dbx_kernels.cuh"

namespace dbx {

cudaError_t stage_rebuild_hash(const DevWorld& W, const LaunchCfg& L);

#define GRID_STRIDE(i, n) for (int i = blockIdx.x * blockDim.x + threadIdx.x, _gs = gridDim.x * blockDim.x; i < (n); i += _gs)

// ------------------------------------------------------------------------------------------------ small device utilities
DBX_D float4 ldcg4(const float4* p) { return __ldcg(p); }
DBX_D void stcg4(float4* p, float4 v) { __stcg(p, v); }

// 64-bit mix (bijective) for the pair hash and the colouring priorities
DBX_HD unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
DBX_D int hash_find(const DevWorld& W, unsigned long long key) {
  unsigned mask = (unsigned)W.hCap - 1;
  unsigned h = (unsigned)mix64(key) & mask;
  for (int probe = 0; probe < W.hCap; ++probe) {
    unsigned long long k = W.h_key[h];
    if (k == key) return W.h_val[h];
    if (k == kHashEmpty) return -1;
    h = (h + 1) & mask;
  }
  return -1;
}
// keys are unique per insertion batch, so a CAS on the key word is enough
DBX_D bool hash_insert(const DevWorld& W, unsigned long long key, int val) {
  unsigned mask = (unsigned)W.hCap - 1;
  unsigned h = (unsigned)mix64(key) & mask;
  for (int probe = 0; probe < W.hCap; ++probe) {
    unsigned long long old = atomicCAS(&W.h_key[h], kHashEmpty, key);
    if (old == kHashEmpty) { W.h_val[h] = val; return true; }
    h = (h + 1) & mask;
  }
  return false;
}
DBX_D void hash_remove(const DevWorld& W, unsigned long long key) {
  unsigned mask = (unsigned)W.hCap - 1;
  unsigned h = (unsigned)mix64(key) & mask;
  for (int probe = 0; probe < W.hCap; ++probe) {
    unsigned long long k = W.h_key[h];
    if (k == key) { W.h_key[h] = kHashTomb; atomicAdd(&W.hdr->nTomb, 1); return; }
    if (k == kHashEmpty) return;
    h = (h + 1) & mask;
  }
}

// lock-free union-find; the larger index is always hooked under the smaller, so a component's root is its minimum id
DBX_D int uf_find(int* parent, int x) {
  for (;;) {
    int p = parent[x];
    if (p == x) return x;
    int gp = parent[p];
    if (gp != p) parent[x] = gp;  // path halving (benign race)
    x = p;
  }
}
DBX_D void uf_unite(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    if (atomicCAS(&parent[a], a, b) == a) return;
  }
}

// global barrier for the persistent kernels: all CTAs are co-resident (cooperative launch, one per SM)
DBX_D void grid_barrier(unsigned* counter, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned ticket = atomicAdd(counter, 1u);
    unsigned target = (ticket / nblocks + 1u) * nblocks;
    while (*((volatile unsigned*)counter) < target) { }
    __threadfence();
  }
  __syncthreads();
}

// b2ContactFilter.ShouldCollide (dynamics/b2worldcallbacks.d:52-64)
DBX_D bool filter_should_collide(const DevWorld& W, int fA, int fB) {
  short gA = (short)(W.f_group[fA] & 0xFFFF), gB = (short)(W.f_group[fB] & 0xFFFF);
  if (gA == gB && gA != 0) return gA > 0;
  uint32_t a = W.f_filter[fA], b = W.f_filter[fB];
  uint32_t catA = a & 0xFFFF, maskA = a >> 16, catB = b & 0xFFFF, maskB = b >> 16;
  return (maskA & catB) != 0 && (catA & maskB) != 0;
}
// b2Body.ShouldCollide (dynamics/b2body.d:1149-1170): at least one dynamic body, no joint that forbids it
DBX_D bool body_should_collide(const DevWorld& W, int bA, int bB, uint32_t flA, uint32_t flB) {
  if (body_type(flA) != BODY_DYNAMIC && body_type(flB) != BODY_DYNAMIC) return false;
  if (W.nJointPairs > 0) {
    unsigned long long lo = (unsigned)min(bA, bB), hi = (unsigned)max(bA, bB);
    unsigned long long k = (lo << 32) | hi;
    int l = 0, r = W.nJointPairs;
    while (l < r) { int m = (l + r) >> 1; if (W.jp_keys[m] < k) l = m + 1; else r = m; }
    if (l < W.nJointPairs && W.jp_keys[l] == k) return false;
  }
  return true;
}
DBX_D void wake_body_now(const DevWorld& W, int b) {  // b2Body.SetAwake(true) (b2body.d:829-835)
  uint32_t old = atomicOr(&W.b_flags[b], BF_AWAKE);
  if (!(old & BF_AWAKE)) W.b_gs[b].y = 0.0f;
}

// ------------------------------------------------------------------------------------------------ Collide
// b2ContactManager.Collide (dynamics/b2contactmanager.d:251-317) + b2Contact.Update (contacts/b2contact.d:270-356)
DBX_D void destroy_contact(const DevWorld& W, int i, uint32_t flags, int bodyA, int bodyB, int pointCount) {
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
  flags = touching ? (flags | CF_TOUCHING) : (flags & ~CF_TOUCHING);
  W.c_flags[i] = flags;
  return flags;
}

__global__ void __launch_bounds__(256) k_collide(const __grid_constant__ DevWorld W) {
  const int n = W.hdr->cHigh;
  GRID_STRIDE(i, n) {
    uint32_t flags = W.c_flags[i];
    if (!(flags & CF_ALIVE)) continue;
    const int4 ids = W.c_ids[i];
    const int4 fx = W.c_fix[i];
    const uint32_t flA = W.b_flags[ids.z], flB = W.b_flags[ids.w];
    uint4 mk = W.c_mk[i];
    if (flags & CF_FILTER) {
      if (!body_should_collide(W, ids.w, ids.z, flB, flA) || !filter_should_collide(W, fx.x, fx.y)) {
        destroy_contact(W, i, flags, ids.z, ids.w, (int)mk.w);
        continue;
      }
      flags &= ~CF_FILTER;
      W.c_flags[i] = flags;
    }
    bool activeA = (flA & BF_AWAKE) && body_type(flA) != BODY_STATIC;
    bool activeB = (flB & BF_AWAKE) && body_type(flB) != BODY_STATIC;
    if (!activeA && !activeB) continue;
    if (!overlap(BX(W.p_fat[ids.x]), BX(W.p_fat[ids.y]))) {
      destroy_contact(W, i, flags, ids.z, ids.w, (int)mk.w);
      continue;
    }
    update_contact(W, i, flags, ids, fx, false);
  }
}

// ------------------------------------------------------------------------------------------------ islands
// Replaces the DFS of b2World.Solve (dynamics/b2world.d:943-1095) with a union-find over constraint edges.
__global__ void __launch_bounds__(256) k_island_init(const __grid_constant__ DevWorld W) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { W.hdr->nSolve = 0; W.hdr->nIslands = 0; W.hdr->nUncoloured = 0; W.hdr->nUncoloured2 = 0; }
  GRID_STRIDE(b, W.nBodies) {
    W.b_root[b] = b;
    W.b_islAwake[b] = 0;
    W.b_islMinSleep[b] = 0x7f7fffff;  // FLT_MAX bits
    W.b_mask[b] = 0ull;
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
    uf_unite(W.b_root, ids.z, ids.w);
  }
  GRID_STRIDE(j, W.nJoints) {
    int4 ids = W.j_ids[j];
    if (!(ids.w & 8)) continue;
    uint32_t fa = W.b_flags[ids.y], fb = W.b_flags[ids.z];
    if (!(fa & BF_ACTIVE) || !(fb & BF_ACTIVE)) continue;                                  // other body must be active (:1058-1062)
    if (body_type(fa) == BODY_STATIC || body_type(fb) == BODY_STATIC) continue;
    uf_unite(W.b_root, ids.y, ids.z);
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

// mark the contacts the solver takes this step and rebuild the per-body colour masks from the persistent colours
__global__ void __launch_bounds__(256) k_mark_solve(const __grid_constant__ DevWorld W) {
  const int n = W.hdr->cHigh;
  GRID_STRIDE(i, n) {
    uint32_t flags = W.c_flags[i];
    if (!(flags & CF_ALIVE)) continue;
    bool solve = false;
    int4 ids;
    if ((flags & (CF_TOUCHING | CF_ENABLED | CF_SENSOR)) == (CF_TOUCHING | CF_ENABLED)) {
      ids = W.c_ids[i];
      uint32_t fa = W.b_flags[ids.z], fb = W.b_flags[ids.w];
      // the contact is in an island iff one of its non-static bodies is (b2world.d:1006-1046)
      solve = ((fa & BF_ISLAND) && body_type(fa) != BODY_STATIC) || ((fb & BF_ISLAND) && body_type(fb) != BODY_STATIC);
    }
    if (!solve) {
      if (flags & CF_SOLVE) W.c_flags[i] = flags & ~CF_SOLVE;
      if (!W.colourOverride) W.c_colour[i] = -1;   // a colour is held only while the contact is in the solver (and so in the masks)
      continue;
    }
    if (!(flags & CF_SOLVE)) W.c_flags[i] = flags | CF_SOLVE;
    if (W.colourOverride) { if (W.c_colour[i] < 0) W.c_colour[i] = kMaxColours - 1; continue; }   // test hook: caller-supplied schedule
    int col = W.c_colour[i];
    uint32_t fa = W.b_flags[ids.z], fb = W.b_flags[ids.w];
    if (col >= 0 && col < kMaskColours) {
      if (body_type(fa) == BODY_DYNAMIC) atomicOr(&W.b_mask[ids.z], 1ull << col);
      if (body_type(fb) == BODY_DYNAMIC) atomicOr(&W.b_mask[ids.w], 1ull << col);
    } else {
      W.c_colour[i] = -1;
      int slot = atomicAdd(&W.hdr->nUncoloured, 1);
      W.c_work[slot] = i;
    }
  }
}

// ------------------------------------------------------------------------------------------------ graph colouring
// New touching contacts take the lowest colour free on both dynamic bodies.  Conflicts between contacts coloured in the
// same round are arbitrated Jones-Plassmann style: per body the contact with the highest (key-derived) priority wins,
// so the outcome is independent of thread scheduling.  Static/kinematic bodies are never written by the solver and do
// not constrain colours.  Runs as one persistent cooperative kernel; rounds loop on the device.
// pair key with the replica offset removed, so that every replica of a batched world arbitrates (and hence colours) alike
DBX_D unsigned long long local_key(const DevWorld& W, unsigned long long key, int body) {
  if (W.keyStride == 0) return key;
  const unsigned long long o = (unsigned long long)(unsigned)(W.b_world[body] * W.keyStride);
  return key - (o << 32) - o;
}
__global__ void __launch_bounds__(512) k_colour(const __grid_constant__ DevWorld W) {
  Header* H = W.hdr;
  const unsigned nb = gridDim.x;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  int* cur = W.c_work; int* nxt = W.c_work2;
  int n = H->nUncoloured;
  unsigned epoch = H->epoch;
  if (epoch > 0xF0000u) {   // the round stamp is 20 bits wide: recycle it long before it wraps
    for (int b = tid; b < W.nBodies; b += nth) W.b_claim[b] = 0ull;
    epoch = 0;
    grid_barrier(&H->barrier, nb);
  }
  int guard = 0;
  while (n > 0 && guard++ < 4096) {
    ++epoch;
    // phase 1: claim both bodies
    for (int k = tid; k < n; k += nth) {
      int i = cur[k];
      int4 ids = W.c_ids[i];
      unsigned long long pr = ((unsigned long long)(epoch & 0xFFFFF) << 44) | (mix64(local_key(W, W.c_key[i], ids.z)) >> 20);
      if (body_type(W.b_flags[ids.z]) == BODY_DYNAMIC) atomicMax(&W.b_claim[ids.z], pr);
      if (body_type(W.b_flags[ids.w]) == BODY_DYNAMIC) atomicMax(&W.b_claim[ids.w], pr);
    }
    if (tid == 0) H->nUncoloured2 = 0;
    grid_barrier(&H->barrier, nb);
    // phase 2: winners take a colour
    for (int k = tid; k < n; k += nth) {
      int i = cur[k];
      int4 ids = W.c_ids[i];
      unsigned long long pr = ((unsigned long long)(epoch & 0xFFFFF) << 44) | (mix64(local_key(W, W.c_key[i], ids.z)) >> 20);
      bool dynA = body_type(W.b_flags[ids.z]) == BODY_DYNAMIC, dynB = body_type(W.b_flags[ids.w]) == BODY_DYNAMIC;
      bool win = (!dynA || __ldcg(&W.b_claim[ids.z]) == pr) && (!dynB || __ldcg(&W.b_claim[ids.w]) == pr);
      if (win) {
        unsigned long long used = (dynA ? __ldcg(&W.b_mask[ids.z]) : 0ull) | (dynB ? __ldcg(&W.b_mask[ids.w]) : 0ull);
        int col;
        if (~used) {
          col = __ffsll((long long)~used) - 1;
          if (dynA) __stcg(&W.b_mask[ids.z], __ldcg(&W.b_mask[ids.z]) | (1ull << col));
          if (dynB) __stcg(&W.b_mask[ids.w], __ldcg(&W.b_mask[ids.w]) | (1ull << col));
        } else {
          // more than 64 touching contacts on one body: serialise the surplus on private overflow lanes of that body
          int oa = dynA ? W.b_ovf[ids.z] : 0, ob = dynB ? W.b_ovf[ids.w] : 0;
          int o = max(oa, ob);
          if (dynA) W.b_ovf[ids.z] = o + 1;
          if (dynB) W.b_ovf[ids.w] = o + 1;
          col = kMaskColours + o;
          if (col >= kMaxColours) { col = kMaxColours - 1; H->error = -5; }
        }
        W.c_colour[i] = col;
      } else {
        int slot = atomicAdd(&H->nUncoloured2, 1);
        nxt[slot] = i;
      }
    }
    grid_barrier(&H->barrier, nb);
    n = *((volatile int*)&H->nUncoloured2);
    int* t = cur; cur = nxt; nxt = t;
    grid_barrier(&H->barrier, nb);
  }
  if (tid == 0) { H->epoch = epoch; H->nUncoloured = 0; }
}

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
    if (total > W.sCap) W.hdr->error = -5;
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

// ------------------------------------------------------------------------------------------------ contact constraints
// b2ContactSolver ctor + InitializeVelocityConstraints (contacts/b2contactsolver.d:244-450): one thread per solver contact.
// `s` = solver slot to fill, `i` = contact slot; warmScale < 0 means "no warm starting" (TOI sub-steps, b2world.d:1419)
DBX_D void prepare_contact(const DevWorld& W, int s, int i, float warmScale) {
  {
    const int4 ids = W.c_ids[i];
    const int4 fx = W.c_fix[i];
    const int bA = ids.z, bB = ids.w;
    const float4 msA = W.b_mass[bA], msB = W.b_mass[bB];
    const float4 lcA4 = W.b_lc[bA], lcB4 = W.b_lc[bB];
    const float4 posA = ldcg4(&W.b_pos[bA]), posB = ldcg4(&W.b_pos[bB]);
    const float4 velA = ldcg4(&W.b_vel[bA]), velB = ldcg4(&W.b_vel[bB]);
    const float4 xA = ldcg4(&W.b_xf[bA]), xB = ldcg4(&W.b_xf[bB]);
    const float4 m0 = W.c_m0[i], m1 = W.c_m1[i], cimp = W.c_imp[i], mat = W.c_mat[i];
    const uint4 mk = W.c_mk[i];
    const float radiusA = W.shapes[fx.z].radius, radiusB = W.shapes[fx.w].radius;
    const float mA = msA.x, iA = msA.y, mB = msB.x, iB = msB.y;
    const v2 localCenterA = V(lcA4.x, lcA4.y), localCenterB = V(lcB4.x, lcB4.y);
    const v2 cA = V(posA.x, posA.y), cB = V(posB.x, posB.y);
    const v2 vA = V(velA.x, velA.y), vB = V(velB.x, velB.y);
    const float wA = velA.z, wB = velB.z;
    const int pointCount = (int)mk.w, type = (int)mk.z;
    // xf from (c, a): q is the body's stored rotation (always sin/cos of a), p = c - q * localCenter (:371-375)
    Xf xfA, xfB;
    xfA.q = R(xA.z, xA.w); xfB.q = R(xB.z, xB.w);
    xfA.p = cA - mul(xfA.q, localCenterA);
    xfB.p = cB - mul(xfB.q, localCenterB);
    const v2 localNormal = V(m0.x, m0.y), localPoint = V(m0.z, m0.w);
    const v2 lp0 = V(m1.x, m1.y), lp1 = V(m1.z, m1.w);
    // b2WorldManifold.Initialize (collision/b2collision.d:123-191)
    v2 normal, wp[2];
    if (type == MAN_CIRCLES) {
      normal = V(1.0f, 0.0f);
      v2 pointA = mul(xfA, localPoint), pointB = mul(xfB, lp0);
      if (dist2(pointA, pointB) > kEpsilon * kEpsilon) { normal = pointB - pointA; normalize(normal); }
      v2 ca = pointA + radiusA * normal, cb = pointB - radiusB * normal;
      wp[0] = 0.5f * (ca + cb); wp[1] = wp[0];
    } else if (type == MAN_FACE_A) {
      normal = mul(xfA.q, localNormal);
      v2 planePoint = mul(xfA, localPoint);
      for (int k = 0; k < pointCount; ++k) {
        v2 clipPoint = mul(xfB, k == 0 ? lp0 : lp1);
        v2 ca = clipPoint + (radiusA - dot(clipPoint - planePoint, normal)) * normal;
        v2 cb = clipPoint - radiusB * normal;
        wp[k] = 0.5f * (ca + cb);
      }
    } else {
      normal = mul(xfB.q, localNormal);
      v2 planePoint = mul(xfB, localPoint);
      for (int k = 0; k < pointCount; ++k) {
        v2 clipPoint = mul(xfA, k == 0 ? lp0 : lp1);
        v2 cb = clipPoint + (radiusB - dot(clipPoint - planePoint, normal)) * normal;
        v2 ca = clipPoint - radiusA * normal;
        wp[k] = 0.5f * (ca + cb);
      }
      normal = -normal;
    }
    const float friction = mat.x, restitution = mat.y, tangentSpeed = mat.z;
    const v2 tangent = cross(normal, 1.0f);
    float4 r[2], q[2];
    r[1] = make_float4(0, 0, 0, 0); q[1] = make_float4(0, 0, 0, 0);
    for (int k = 0; k < pointCount; ++k) {
      v2 rA = wp[k] - cA, rB = wp[k] - cB;
      float rnA = cross(rA, normal), rnB = cross(rB, normal);
      float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
      float normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
      float rtA = cross(rA, tangent), rtB = cross(rB, tangent);
      float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
      float tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
      float velocityBias = 0.0f;
      float vRel = dot(normal, vB + cross(wB, rB) - vA - cross(wA, rA));
      if (vRel < -kVelocityThreshold) velocityBias = -restitution * vRel;
      r[k] = make_float4(rA.x, rA.y, rB.x, rB.y);
      q[k] = make_float4(normalMass, tangentMass, velocityBias, 0.0f);
    }
    int vcCount = pointCount;
    float4 nm = make_float4(0, 0, 0, 0), K = make_float4(0, 0, 0, 0);
    if (pointCount == 2) {
      // block-solver matrix with the reference's conditioning test (:418-448)
      float rn1A = cross(V(r[0].x, r[0].y), normal), rn1B = cross(V(r[0].z, r[0].w), normal);
      float rn2A = cross(V(r[1].x, r[1].y), normal), rn2B = cross(V(r[1].z, r[1].w), normal);
      float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
      float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
      float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
      const float k_maxConditionNumber = 1000.0f;
      if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
        M22 Km; Km.ex = V(k11, k12); Km.ey = V(k12, k22);
        M22 inv = inverse(Km);
        K = make_float4(k11, k12, k12, k22);
        nm = make_float4(inv.ex.x, inv.ex.y, inv.ey.x, inv.ey.y);
      } else {
        vcCount = 1;
      }
    }
    // warm-start impulses scaled by dtRatio (:309-313)
    float4 imp = warmScale >= 0.0f ? make_float4(warmScale * cimp.x, warmScale * cimp.y, warmScale * cimp.z, warmScale * cimp.w)
                                   : make_float4(0, 0, 0, 0);
    if (pointCount < 2) { imp.z = 0.0f; imp.w = 0.0f; }
    W.s_body[s] = make_int2(bA, bB);
    W.s_v0[s] = make_float4(normal.x, normal.y, friction, tangentSpeed);
    W.s_v1[s] = make_float4(mA, iA, mB, iB);
    W.s_r0[s] = r[0]; W.s_r1[s] = r[1];
    W.s_q0[s] = q[0]; W.s_q1[s] = q[1];
    W.s_imp[s] = imp;
    W.s_nm[s] = nm; W.s_k[s] = K;
    W.s_pc[s] = vcCount | (type << 8) | (pointCount << 16);
    W.s_p0[s] = m1;
    W.s_p1[s] = m0;
    W.s_p2[s] = make_float4(localCenterA.x, localCenterA.y, localCenterB.x, localCenterB.y);
    W.s_p3[s] = make_float2(radiusA, radiusB);
    uint32_t fa = W.b_flags[bA];
    W.s_root[s] = body_type(fa) != BODY_STATIC ? W.b_root[bA] : W.b_root[bB];
  }
}
__global__ void __launch_bounds__(256) k_prepare(const __grid_constant__ DevWorld W) {
  const int n = min(W.hdr->nSolve, W.sCap);
  const float warmScale = W.warmStarting ? W.dtRatio : -1.0f;
  GRID_STRIDE(s, n) prepare_contact(W, s, W.s_contact[s], warmScale);
}

// velocities of one body pair; only dynamic bodies are ever written (statics/kinematics have zero inverse mass)
struct BodyVel { v2 vA, vB; float wA, wB; };
DBX_D BodyVel load_vel(const DevWorld& W, int2 bd) {
  float4 a = ldcg4(&W.b_vel[bd.x]), b = ldcg4(&W.b_vel[bd.y]);
  BodyVel r; r.vA = V(a.x, a.y); r.wA = a.z; r.vB = V(b.x, b.y); r.wB = b.z; return r;
}
DBX_D void store_vel(const DevWorld& W, int2 bd, const BodyVel& r, float mA, float iA, float mB, float iB) {
  if (mA != 0.0f || iA != 0.0f) stcg4(&W.b_vel[bd.x], make_float4(r.vA.x, r.vA.y, r.wA, 0.0f));
  if (mB != 0.0f || iB != 0.0f) stcg4(&W.b_vel[bd.y], make_float4(r.vB.x, r.vB.y, r.wB, 0.0f));
}

// b2ContactSolver.WarmStart (:452-490)
DBX_D void contact_warm_start(const DevWorld& W, int s) {
  const int2 bd = W.s_body[s];
  const float4 v0 = W.s_v0[s], v1 = W.s_v1[s], imp = W.s_imp[s];
  const int pointCount = W.s_pc[s] & 0xFF;
  const float mA = v1.x, iA = v1.y, mB = v1.z, iB = v1.w;
  BodyVel bv = load_vel(W, bd);
  const v2 normal = V(v0.x, v0.y), tangent = cross(normal, 1.0f);
  for (int k = 0; k < pointCount; ++k) {
    const float4 r = k == 0 ? W.s_r0[s] : W.s_r1[s];
    const float ni = k == 0 ? imp.x : imp.z, ti = k == 0 ? imp.y : imp.w;
    v2 P = ni * normal + ti * tangent;
    bv.wA -= iA * cross(V(r.x, r.y), P);
    bv.vA -= mA * P;
    bv.wB += iB * cross(V(r.z, r.w), P);
    bv.vB += mB * P;
  }
  store_vel(W, bd, bv, mA, iA, mB, iB);
}

// b2ContactSolver.SolveVelocityConstraints (:492-772): friction rows, then 1-point clamp or the 2-point block solver
// constraint block of one solver contact, loadable ahead of the barrier that precedes its colour (only s_imp ever changes,
// and only through the thread that owns the slot)
struct VC { int2 bd; int pc; float4 v0, v1, r0, r1, q0, q1, imp, nm, K; };
DBX_D void vc_load(const DevWorld& W, int s, VC& c) {
  c.bd = W.s_body[s]; c.pc = W.s_pc[s];
  c.v0 = W.s_v0[s]; c.v1 = W.s_v1[s]; c.r0 = W.s_r0[s]; c.q0 = W.s_q0[s]; c.imp = W.s_imp[s];
  c.r1 = W.s_r1[s]; c.q1 = W.s_q1[s]; c.nm = W.s_nm[s]; c.K = W.s_k[s];
}
DBX_D void contact_solve_velocity(const DevWorld& W, int s, const VC& c) {
  const int2 bd = c.bd;
  const float4 v0 = c.v0, v1 = c.v1;
  float4 imp = c.imp;
  const int pointCount = c.pc & 0xFF;
  const float mA = v1.x, iA = v1.y, mB = v1.z, iB = v1.w;
  const float4 r0 = c.r0, q0 = c.q0;
  const float4 r1 = c.r1, q1 = c.q1;
  BodyVel bv = load_vel(W, bd);
  v2 vA = bv.vA, vB = bv.vB; float wA = bv.wA, wB = bv.wB;
  const v2 normal = V(v0.x, v0.y), tangent = cross(normal, 1.0f);
  const float friction = v0.z, tangentSpeed = v0.w;
  for (int k = 0; k < pointCount; ++k) {
    const float4 r = k == 0 ? r0 : r1;
    const float tangentMass = k == 0 ? q0.y : q1.y;
    const v2 rA = V(r.x, r.y), rB = V(r.z, r.w);
    float ni = k == 0 ? imp.x : imp.z, ti = k == 0 ? imp.y : imp.w;
    v2 dv = vB + cross(wB, rB) - vA - cross(wA, rA);
    float vt = dot(dv, tangent) - tangentSpeed;
    float lambda = tangentMass * (-vt);
    float maxFriction = friction * ni;
    float newImpulse = fclampr(ti + lambda, -maxFriction, maxFriction);
    lambda = newImpulse - ti;
    if (k == 0) imp.y = newImpulse; else imp.w = newImpulse;
    v2 P = lambda * tangent;
    vA -= mA * P; wA -= iA * cross(rA, P);
    vB += mB * P; wB += iB * cross(rB, P);
  }
  if (pointCount == 1) {
    const v2 rA = V(r0.x, r0.y), rB = V(r0.z, r0.w);
    v2 dv = vB + cross(wB, rB) - vA - cross(wA, rA);
    float vn = dot(dv, normal);
    float lambda = -q0.x * (vn - q0.z);
    float newImpulse = fmaxr(imp.x + lambda, 0.0f);
    lambda = newImpulse - imp.x;
    imp.x = newImpulse;
    v2 P = lambda * normal;
    vA -= mA * P; wA -= iA * cross(rA, P);
    vB += mB * P; wB += iB * cross(rB, P);
  } else {
    const float4 nm = c.nm, K = c.K;
    const v2 rA1 = V(r0.x, r0.y), rB1 = V(r0.z, r0.w), rA2 = V(r1.x, r1.y), rB2 = V(r1.z, r1.w);
    v2 a = V(imp.x, imp.z);
    v2 dv1 = vB + cross(wB, rB1) - vA - cross(wA, rA1);
    v2 dv2 = vB + cross(wB, rB2) - vA - cross(wA, rA2);
    float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
    v2 b;
    b.x = vn1 - q0.z;
    b.y = vn2 - q1.z;
    M22 Km; Km.ex = V(K.x, K.y); Km.ey = V(K.z, K.w);
    M22 NM; NM.ex = V(nm.x, nm.y); NM.ey = V(nm.z, nm.w);
    b -= mul(Km, a);
    v2 x;
    bool found = false;
    // the four LCP cases in the reference's order (:633-765)
    x = -mul(NM, b);
    if (x.x >= 0.0f && x.y >= 0.0f) found = true;
    if (!found) {
      x.x = -q0.x * b.x; x.y = 0.0f;
      vn2 = Km.ex.y * x.x + b.y;
      if (x.x >= 0.0f && vn2 >= 0.0f) found = true;
    }
    if (!found) {
      x.x = 0.0f; x.y = -q1.x * b.y;
      vn1 = Km.ey.x * x.y + b.x;
      if (x.y >= 0.0f && vn1 >= 0.0f) found = true;
    }
    if (!found) {
      x.x = 0.0f; x.y = 0.0f;
      vn1 = b.x; vn2 = b.y;
      if (vn1 >= 0.0f && vn2 >= 0.0f) found = true;
    }
    if (found) {
      v2 d = x - a;
      v2 P1 = d.x * normal, P2 = d.y * normal;
      vA -= mA * (P1 + P2);
      wA -= iA * (cross(rA1, P1) + cross(rA2, P2));
      vB += mB * (P1 + P2);
      wB += iB * (cross(rB1, P1) + cross(rB2, P2));
      imp.x = x.x; imp.z = x.y;
    }
  }
  W.s_imp[s] = imp;
  bv.vA = vA; bv.vB = vB; bv.wA = wA; bv.wB = wB;
  store_vel(W, bd, bv, mA, iA, mB, iB);
}

DBX_D void contact_solve_velocity(const DevWorld& W, int s) { VC c; vc_load(W, s, c); contact_solve_velocity(W, s, c); }

// b2ContactSolver.SolvePositionConstraints (:73-149) + b2PositionSolverManifold (:816-868); returns min separation
// toiA/toiB >= 0 selects SolveTOIPositionConstraints (:152-242): only those two bodies keep their mass, Baumgarte 0.75
DBX_D float contact_solve_position(const DevWorld& W, int s, int toiA = -1, int toiB = -1) {
  const int2 bd = W.s_body[s];
  const float4 v1 = W.s_v1[s], p0 = W.s_p0[s], p1 = W.s_p1[s], p2 = W.s_p2[s];
  const float2 p3 = W.s_p3[s];
  const int pc = W.s_pc[s];
  const int type = (pc >> 8) & 0xFF, pointCount = pc >> 16;
  float mA = v1.x, iA = v1.y, mB = v1.z, iB = v1.w;
  const bool toi = toiA >= 0;
  if (toi) {
    if (bd.x != toiA && bd.x != toiB) { mA = 0.0f; iA = 0.0f; }
    if (bd.y != toiA && bd.y != toiB) { mB = 0.0f; iB = 0.0f; }
  }
  const float baumgarte = toi ? kToiBaumgarte : kBaumgarte;
  const v2 localCenterA = V(p2.x, p2.y), localCenterB = V(p2.z, p2.w);
  const v2 localNormal = V(p1.x, p1.y), localPoint = V(p1.z, p1.w);
  float4 pa = ldcg4(&W.b_pos[bd.x]), pb = ldcg4(&W.b_pos[bd.y]);
  v2 cA = V(pa.x, pa.y), cB = V(pb.x, pb.y);
  float aA = pa.z, aB = pb.z;
  float minSeparation = 0.0f;
  for (int j = 0; j < pointCount; ++j) {
    Xf xfA, xfB;
    xfA.q = rot_from_angle(aA); xfB.q = rot_from_angle(aB);
    xfA.p = cA - mul(xfA.q, localCenterA);
    xfB.p = cB - mul(xfB.q, localCenterB);
    v2 normal, point; float separation;
    if (type == MAN_CIRCLES) {
      v2 pointA = mul(xfA, localPoint), pointB = mul(xfB, V(p0.x, p0.y));
      normal = pointB - pointA;
      normalize(normal);
      point = 0.5f * (pointA + pointB);
      separation = dot(pointB - pointA, normal) - p3.x - p3.y;
    } else if (type == MAN_FACE_A) {
      normal = mul(xfA.q, localNormal);
      v2 planePoint = mul(xfA, localPoint);
      v2 clipPoint = mul(xfB, j == 0 ? V(p0.x, p0.y) : V(p0.z, p0.w));
      separation = dot(clipPoint - planePoint, normal) - p3.x - p3.y;
      point = clipPoint;
    } else {
      normal = mul(xfB.q, localNormal);
      v2 planePoint = mul(xfB, localPoint);
      v2 clipPoint = mul(xfA, j == 0 ? V(p0.x, p0.y) : V(p0.z, p0.w));
      separation = dot(clipPoint - planePoint, normal) - p3.x - p3.y;
      point = clipPoint;
      normal = -normal;
    }
    v2 rA = point - cA, rB = point - cB;
    minSeparation = fminr(minSeparation, separation);
    float C = fclampr(baumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
    float rnA = cross(rA, normal), rnB = cross(rB, normal);
    float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
    float impulse = K > 0.0f ? -C / K : 0.0f;
    v2 P = impulse * normal;
    cA -= mA * P; aA -= iA * cross(rA, P);
    cB += mB * P; aB += iB * cross(rB, P);
  }
  if (mA != 0.0f || iA != 0.0f) stcg4(&W.b_pos[bd.x], make_float4(cA.x, cA.y, aA, 0.0f));
  if (mB != 0.0f || iB != 0.0f) stcg4(&W.b_pos[bd.y], make_float4(cB.x, cB.y, aB, 0.0f));
  return minSeparation;
}

// ------------------------------------------------------------------------------------------------ joints
// revolute: dynamics/joints/b2revolutejoint.d:319-636; distance: b2distancejoint.d:211-373
DBX_D bool joint_active(const DevWorld& W, int j);
DBX_D void joint_init(const DevWorld& W, int j) {
  if (!joint_active(W, j)) { W.j_root[j] = -1; return; }   // later phases only look at j_root
  const int4 ids = W.j_ids[j];
  const int bA = ids.y, bB = ids.z;
  const float4 msA = W.b_mass[bA], msB = W.b_mass[bB];
  const float4 lcA4 = W.b_lc[bA], lcB4 = W.b_lc[bB];
  const float4 anc = W.j_anchor[j];
  const float mA = msA.x, iA = msA.y, mB = msB.x, iB = msB.y;
  const v2 localCenterA = V(lcA4.x, lcA4.y), localCenterB = V(lcB4.x, lcB4.y);
  const float4 posA = W.b_pos[bA], posB = W.b_pos[bB];
  const float4 xA = W.b_xf[bA], xB = W.b_xf[bB];
  float4 velA = ldcg4(&W.b_vel[bA]), velB = ldcg4(&W.b_vel[bB]);
  v2 vA = V(velA.x, velA.y), vB = V(velB.x, velB.y); float wA = velA.z, wB = velB.z;
  const float aA = posA.z, aB = posB.z;
  const Rot qA = R(xA.z, xA.w), qB = R(xB.z, xB.w);  // = b2Rot(aA), b2Rot(aB): positions are not integrated yet
  const v2 rA = mul(qA, V(anc.x, anc.y) - localCenterA);
  const v2 rB = mul(qB, V(anc.z, anc.w) - localCenterB);
  W.j_r[j] = make_float4(rA.x, rA.y, rB.x, rB.y);
  W.j_lc[j] = make_float4(localCenterA.x, localCenterA.y, localCenterB.x, localCenterB.y);
  W.j_m[j] = make_float4(mA, iA, mB, iB);
  W.j_root[j] = body_type(W.b_flags[bA]) != BODY_STATIC ? W.b_root[bA] : W.b_root[bB];
  float4 imp = W.j_imp[j];
  if (ids.x == JT_REVOLUTE) {
    const float4 p0 = W.j_p0[j];
    const bool enableLimit = (ids.w & 2) != 0, enableMotor = (ids.w & 4) != 0;
    const bool fixedRotation = (iA + iB == 0.0f);
    v3 ex, ey, ez;
    ex.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
    ey.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
    ez.x = -rA.y * iA - rB.y * iB;
    ex.y = ey.x;
    ey.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
    ez.y = rA.x * iA + rB.x * iB;
    ex.z = ez.x;
    ey.z = ez.y;
    ez.z = iA + iB;
    float motorMass = iA + iB;
    if (motorMass > 0.0f) motorMass = 1.0f / motorMass;
    if (!enableMotor || fixedRotation) imp.w = 0.0f;
    int limitState = W.j_limit[j];
    if (enableLimit && !fixedRotation) {
      float jointAngle = aB - aA - p0.x;
      if (fabsr(p0.z - p0.y) < 2.0f * kAngularSlop) limitState = LIM_EQUAL;
      else if (jointAngle <= p0.y) { if (limitState != LIM_LOWER) imp.z = 0.0f; limitState = LIM_LOWER; }
      else if (jointAngle >= p0.z) { if (limitState != LIM_UPPER) imp.z = 0.0f; limitState = LIM_UPPER; }
      else { limitState = LIM_INACTIVE; imp.z = 0.0f; }
    } else {
      limitState = LIM_INACTIVE;
    }
    W.j_limit[j] = limitState;
    if (W.warmStarting) {
      imp.x *= W.dtRatio; imp.y *= W.dtRatio; imp.z *= W.dtRatio;
      imp.w *= W.dtRatio;
      v2 P = V(imp.x, imp.y);
      vA -= mA * P;
      wA -= iA * (cross(rA, P) + imp.w + imp.z);
      vB += mB * P;
      wB += iB * (cross(rB, P) + imp.w + imp.z);
    } else {
      imp = make_float4(0, 0, 0, 0);
    }
    W.j_k0[j] = make_float4(ex.x, ex.y, ex.z, motorMass);
    W.j_k1[j] = make_float4(ey.x, ey.y, ey.z, 0.0f);
    W.j_k2[j] = make_float4(ez.x, ez.y, ez.z, 0.0f);
  } else {  // JT_DISTANCE
    const float4 p0 = W.j_p0[j];
    const v2 cA = V(posA.x, posA.y), cB = V(posB.x, posB.y);
    v2 u = cB + rB - cA - rA;
    float length = len(u);
    if (length > kLinearSlop) u *= 1.0f / length; else u = V(0.0f, 0.0f);
    float crAu = cross(rA, u), crBu = cross(rB, u);
    float invMass = mA + iA * crAu * crAu + mB + iB * crBu * crBu;
    float mass = invMass != 0.0f ? 1.0f / invMass : 0.0f;
    float gamma = 0.0f, bias = 0.0f;
    if (p0.y > 0.0f) {
      float C = length - p0.x;
      float omega = 2.0f * kPi * p0.y;
      float d = 2.0f * mass * p0.z * omega;
      float k = mass * omega * omega;
      float h = W.dt;
      gamma = h * (d + h * k);
      gamma = gamma != 0.0f ? 1.0f / gamma : 0.0f;
      bias = C * h * k * gamma;
      invMass += gamma;
      mass = invMass != 0.0f ? 1.0f / invMass : 0.0f;
    }
    if (W.warmStarting) {
      imp.x *= W.dtRatio;
      v2 P = imp.x * u;
      vA -= mA * P; wA -= iA * cross(rA, P);
      vB += mB * P; wB += iB * cross(rB, P);
    } else {
      imp.x = 0.0f;
    }
    W.j_k0[j] = make_float4(u.x, u.y, mass, gamma);
    W.j_k1[j] = make_float4(bias, 0.0f, 0.0f, 0.0f);
  }
  W.j_imp[j] = imp;
  if (mA != 0.0f || iA != 0.0f) stcg4(&W.b_vel[bA], make_float4(vA.x, vA.y, wA, 0.0f));
  if (mB != 0.0f || iB != 0.0f) stcg4(&W.b_vel[bB], make_float4(vB.x, vB.y, wB, 0.0f));
}

DBX_D void joint_solve_velocity(const DevWorld& W, int j) {
  const int4 ids = W.j_ids[j];
  const int bA = ids.y, bB = ids.z;
  const float4 m = W.j_m[j], r = W.j_r[j];
  const float mA = m.x, iA = m.y, mB = m.z, iB = m.w;
  const v2 rA = V(r.x, r.y), rB = V(r.z, r.w);
  float4 velA = ldcg4(&W.b_vel[bA]), velB = ldcg4(&W.b_vel[bB]);
  v2 vA = V(velA.x, velA.y), vB = V(velB.x, velB.y); float wA = velA.z, wB = velB.z;
  float4 imp = W.j_imp[j];
  if (ids.x == JT_REVOLUTE) {
    const float4 p0 = W.j_p0[j], p1 = W.j_p1[j];
    const float4 k0 = W.j_k0[j], k1 = W.j_k1[j], k2 = W.j_k2[j];
    const bool enableLimit = (ids.w & 2) != 0, enableMotor = (ids.w & 4) != 0;
    const int limitState = W.j_limit[j];
    const bool fixedRotation = (iA + iB == 0.0f);
    if (enableMotor && limitState != LIM_EQUAL && !fixedRotation) {
      float Cdot = wB - wA - p1.x;
      float impulse = -k0.w * Cdot;
      float oldImpulse = imp.w;
      float maxImpulse = W.dt * p0.w;
      imp.w = fclampr(imp.w + impulse, -maxImpulse, maxImpulse);
      impulse = imp.w - oldImpulse;
      wA -= iA * impulse;
      wB += iB * impulse;
    }
    const v3 ex = V3(k0.x, k0.y, k0.z), ey = V3(k1.x, k1.y, k1.z), ez = V3(k2.x, k2.y, k2.z);
    if (enableLimit && limitState != LIM_INACTIVE && !fixedRotation) {
      v2 Cdot1 = vB + cross(wB, rB) - vA - cross(wA, rA);
      float Cdot2 = wB - wA;
      v3 s = solve33(ex, ey, ez, V3(Cdot1.x, Cdot1.y, Cdot2));
      v3 impulse = V3(-s.x, -s.y, -s.z);
      if (limitState == LIM_EQUAL) {
        imp.x += impulse.x; imp.y += impulse.y; imp.z += impulse.z;
      } else if (limitState == LIM_LOWER) {
        float newImpulse = imp.z + impulse.z;
        if (newImpulse < 0.0f) {
          v2 rhs = -Cdot1 + imp.z * V(ez.x, ez.y);
          v2 reduced = solve22(ex.x, ey.x, ex.y, ey.y, rhs);
          impulse.x = reduced.x; impulse.y = reduced.y; impulse.z = -imp.z;
          imp.x += reduced.x; imp.y += reduced.y; imp.z = 0.0f;
        } else { imp.x += impulse.x; imp.y += impulse.y; imp.z += impulse.z; }
      } else if (limitState == LIM_UPPER) {
        float newImpulse = imp.z + impulse.z;
        if (newImpulse > 0.0f) {
          v2 rhs = -Cdot1 + imp.z * V(ez.x, ez.y);
          v2 reduced = solve22(ex.x, ey.x, ex.y, ey.y, rhs);
          impulse.x = reduced.x; impulse.y = reduced.y; impulse.z = -imp.z;
          imp.x += reduced.x; imp.y += reduced.y; imp.z = 0.0f;
        } else { imp.x += impulse.x; imp.y += impulse.y; imp.z += impulse.z; }
      }
      v2 P = V(impulse.x, impulse.y);
      vA -= mA * P;
      wA -= iA * (cross(rA, P) + impulse.z);
      vB += mB * P;
      wB += iB * (cross(rB, P) + impulse.z);
    } else {
      v2 Cdot = vB + cross(wB, rB) - vA - cross(wA, rA);
      v2 impulse = solve22(ex.x, ey.x, ex.y, ey.y, -Cdot);
      imp.x += impulse.x; imp.y += impulse.y;
      vA -= mA * impulse; wA -= iA * cross(rA, impulse);
      vB += mB * impulse; wB += iB * cross(rB, impulse);
    }
  } else {
    const float4 k0 = W.j_k0[j], k1 = W.j_k1[j];
    const v2 u = V(k0.x, k0.y);
    v2 vpA = vA + cross(wA, rA), vpB = vB + cross(wB, rB);
    float Cdot = dot(u, vpB - vpA);
    float impulse = -k0.z * (Cdot + k1.x + k0.w * imp.x);
    imp.x += impulse;
    v2 P = impulse * u;
    vA -= mA * P; wA -= iA * cross(rA, P);
    vB += mB * P; wB += iB * cross(rB, P);
  }
  W.j_imp[j] = imp;
  if (mA != 0.0f || iA != 0.0f) stcg4(&W.b_vel[bA], make_float4(vA.x, vA.y, wA, 0.0f));
  if (mB != 0.0f || iB != 0.0f) stcg4(&W.b_vel[bB], make_float4(vB.x, vB.y, wB, 0.0f));
}

// returns true when the joint is within tolerance
DBX_D bool joint_solve_position(const DevWorld& W, int j) {
  const int4 ids = W.j_ids[j];
  const int bA = ids.y, bB = ids.z;
  const float4 m = W.j_m[j], lc = W.j_lc[j], anc = W.j_anchor[j];
  const float mA = m.x, iA = m.y, mB = m.z, iB = m.w;
  float4 pa = ldcg4(&W.b_pos[bA]), pb = ldcg4(&W.b_pos[bB]);
  v2 cA = V(pa.x, pa.y), cB = V(pb.x, pb.y); float aA = pa.z, aB = pb.z;
  bool ok;
  if (ids.x == JT_REVOLUTE) {
    const float4 p0 = W.j_p0[j];
    const float motorMass = W.j_k0[j].w;
    const bool enableLimit = (ids.w & 2) != 0;
    const int limitState = W.j_limit[j];
    float angularError = 0.0f, positionError = 0.0f;
    const bool fixedRotation = (iA + iB == 0.0f);
    if (enableLimit && limitState != LIM_INACTIVE && !fixedRotation) {
      float angle = aB - aA - p0.x;
      float limitImpulse = 0.0f;
      if (limitState == LIM_EQUAL) {
        float C = fclampr(angle - p0.y, -kMaxAngularCorrection, kMaxAngularCorrection);
        limitImpulse = -motorMass * C;
        angularError = fabsr(C);
      } else if (limitState == LIM_LOWER) {
        float C = angle - p0.y;
        angularError = -C;
        C = fclampr(C + kAngularSlop, -kMaxAngularCorrection, 0.0f);
        limitImpulse = -motorMass * C;
      } else if (limitState == LIM_UPPER) {
        float C = angle - p0.z;
        angularError = C;
        C = fclampr(C - kAngularSlop, 0.0f, kMaxAngularCorrection);
        limitImpulse = -motorMass * C;
      }
      aA -= iA * limitImpulse;
      aB += iB * limitImpulse;
    }
    {
      Rot qA = rot_from_angle(aA), qB = rot_from_angle(aB);
      v2 rA = mul(qA, V(anc.x, anc.y) - V(lc.x, lc.y));
      v2 rB = mul(qB, V(anc.z, anc.w) - V(lc.z, lc.w));
      v2 C = cB + rB - cA - rA;
      positionError = len(C);
      float k11 = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
      float k12 = -iA * rA.x * rA.y - iB * rB.x * rB.y;
      float k22 = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
      v2 impulse = -solve22(k11, k12, k12, k22, C);
      cA -= mA * impulse; aA -= iA * cross(rA, impulse);
      cB += mB * impulse; aB += iB * cross(rB, impulse);
    }
    ok = positionError <= kLinearSlop && angularError <= kAngularSlop;
  } else {
    const float4 p0 = W.j_p0[j];
    if (p0.y > 0.0f) return true;  // soft joints have no position constraint (b2distancejoint.d:337-341)
    const float mass = W.j_k0[j].z;
    Rot qA = rot_from_angle(aA), qB = rot_from_angle(aB);
    v2 rA = mul(qA, V(anc.x, anc.y) - V(lc.x, lc.y));
    v2 rB = mul(qB, V(anc.z, anc.w) - V(lc.z, lc.w));
    v2 u = cB + rB - cA - rA;
    float length = normalize(u);
    float C = length - p0.x;
    C = fclampr(C, -kMaxLinearCorrection, kMaxLinearCorrection);
    float impulse = -mass * C;
    v2 P = impulse * u;
    cA -= mA * P; aA -= iA * cross(rA, P);
    cB += mB * P; aB += iB * cross(rB, P);
    ok = fabsr(C) < kLinearSlop;
  }
  if (mA != 0.0f || iA != 0.0f) stcg4(&W.b_pos[bA], make_float4(cA.x, cA.y, aA, 0.0f));
  if (mB != 0.0f || iB != 0.0f) stcg4(&W.b_pos[bB], make_float4(cB.x, cB.y, aB, 0.0f));
  return ok;
}

DBX_D bool joint_active(const DevWorld& W, int j) {
  const int4 ids = W.j_ids[j];
  if (!(ids.w & 8)) return false;
  uint32_t fa = W.b_flags[ids.y], fb = W.b_flags[ids.z];
  if (!(fa & BF_ACTIVE) || !(fb & BF_ACTIVE)) return false;
  return ((fa & BF_ISLAND) && body_type(fa) != BODY_STATIC) || ((fb & BF_ISLAND) && body_type(fb) != BODY_STATIC);
}

// ------------------------------------------------------------------------------------------------ the persistent solver
// b2Island.Solve (dynamics/b2island.d:118-279) for ALL awake islands at once.  Colour c of an iteration is one
// barrier-delimited phase; within a colour no two constraints touch the same dynamic body, so every read-modify-write
// of a body is exclusive and the result equals sequential Gauss-Seidel in colour order.
__global__ void __launch_bounds__(512) k_solve(const __grid_constant__ DevWorld W) {
  Header* H = W.hdr;
  const unsigned nb = gridDim.x;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int nColours = H->nColours;
  const int nJointColours = W.nJoints > 0 ? min(W.nJointColours, kMaxJointColours) : 0;
  // colour offsets live in shared memory: every barrier invalidates L1, and a phase must not start with an L2 round trip
  // (let alone 64 of them over empty joint colours) just to learn its own range
  __shared__ int coff[kMaxColours + 1];
  __shared__ int joff[kMaxJointColours + 1];
  for (int c = threadIdx.x; c <= nColours; c += blockDim.x) coff[c] = H->colourOff[c];
  for (int c = threadIdx.x; c <= kMaxJointColours; c += blockDim.x) joff[c] = H->jointColourOff[c];
  __syncthreads();
  // Role split: the first JB CTAs only ever run joint code, the others only contact code, so the two large code paths never
  // evict each other from an SM's instruction cache (measured: +4.4 us on every joint->contact switch otherwise).
  const int JB = (W.nJoints > 0 && (int)nb >= 8) ? min(max(W.jointBlocks, 1), (int)nb / 2) : 0;
  const bool jointRole = JB == 0 || (int)blockIdx.x < JB;
  const bool contactRole = JB == 0 || (int)blockIdx.x >= JB;
  const int jtid = tid, jnth = JB == 0 ? nth : JB * blockDim.x;
  const int ctid = JB == 0 ? tid : tid - JB * blockDim.x, cnth = JB == 0 ? nth : nth - JB * blockDim.x;
  // Tail colours (together at most kTailContacts constraints) run inside ONE CTA with CTA-scope barriers: a colour with a
  // few hundred constraints is not worth a 1.7 us global barrier per pass.
  const int T = min(H->tailStart, nColours);
  const bool tailBlock = blockIdx.x == nb - 1;
  int phaseIdx = 0;
#define PHASE_MARK() do { if (W.phaseTimes && tid == 0 && phaseIdx < W.phaseCap) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); W.phaseTimes[phaseIdx++] = t_; } } while (0)
  PHASE_MARK();
  // debug window (DBX_DEBUG bit 1): per-CTA arrival / release stamps of 8 consecutive barriers starting at phase (flags >> 8)
  const bool dbgWin = (W.dbgFlags & 2) && W.phaseTimes != nullptr;
  const int dbgP0 = W.dbgFlags >> 8;
  int barIdx = 0;
#define GB() do { \
    if (dbgWin && threadIdx.x == 0 && barIdx >= dbgP0 && barIdx < dbgP0 + 8) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); W.phaseTimes[512 + ((barIdx - dbgP0) * nb + blockIdx.x) * 2] = t_; } \
    grid_barrier(&H->barrier, nb); \
    if (dbgWin && threadIdx.x == 0 && barIdx >= dbgP0 && barIdx < dbgP0 + 8) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); W.phaseTimes[512 + ((barIdx - dbgP0) * nb + blockIdx.x) * 2 + 1] = t_; } \
    ++barIdx; PHASE_MARK(); } while (0)
  const bool haveTail = T < nColours && coff[T] < coff[nColours];

  // contacts warm start (b2island.d:138-141), colour by colour
  if (W.warmStarting) {
    for (int c = 0; c < T; ++c) {
      int beg = coff[c], end = coff[c + 1];
      if (beg == end) continue;
      if (contactRole) for (int s = beg + ctid; s < end; s += cnth) contact_warm_start(W, s);
      GB();
    }
    if (haveTail) {
      if (tailBlock) for (int c = T; c < nColours; ++c) {
        for (int s = coff[c] + threadIdx.x; s < coff[c + 1]; s += blockDim.x) contact_warm_start(W, s);
        __syncthreads();
      }
      GB();
    }
  }
  // joints: InitVelocityConstraints incl. their warm start (:143-146)
  for (int c = 0; c < nJointColours; ++c) {
    int beg = joff[c], end = joff[c + 1];
    if (beg == end) continue;
    if (jointRole) for (int k = beg + jtid; k < end; k += jnth) joint_init(W, k);
    GB();
  }
  // velocity iterations: all joints, then all contacts (:153-161)
  VC pre; int preS = -1;
  for (int it = 0; it < W.velIters; ++it) {
    for (int c = 0; c < nJointColours; ++c) {
      int beg = joff[c], end = joff[c + 1];
      if (beg == end) continue;
      if (jointRole && !(W.dbgFlags & 1)) for (int k = beg + jtid; k < end; k += jnth) if (W.j_root[k] >= 0) joint_solve_velocity(W, k);
      GB();
    }
    for (int c = 0; c < T; ++c) {
      int beg = coff[c], end = coff[c + 1];
      if (beg == end) continue;
      if (contactRole) {
        // the first item of this colour was fetched before the previous barrier; fetch the next colour's before this one
        int s = beg + ctid;
        if (s < end) {
          if (preS != s) vc_load(W, s, pre);
          contact_solve_velocity(W, s, pre);
          for (s += cnth; s < end; s += cnth) contact_solve_velocity(W, s);
        }
        int cn = c + 1;
        while (cn < T && coff[cn] == coff[cn + 1]) ++cn;
        if (cn >= T) { cn = 0; while (cn < T && coff[cn] == coff[cn + 1]) ++cn; }
        preS = -1;
        if (cn < T && (cn > c || it + 1 < W.velIters)) { int sn = coff[cn] + ctid; if (sn < coff[cn + 1]) { vc_load(W, sn, pre); preS = sn; } }
      }
      GB();
    }
    if (haveTail) {
      if (tailBlock) for (int c = T; c < nColours; ++c) {
        for (int s = coff[c] + threadIdx.x; s < coff[c + 1]; s += blockDim.x) contact_solve_velocity(W, s);
        __syncthreads();
      }
      GB();
    }
  }
  // StoreImpulses (:164) + integrate positions (:168-200)
  {
    const int n = min(H->nSolve, W.sCap);
    for (int s = tid; s < n; s += nth) {
      const int i = W.s_contact[s];
      const int vcCount = W.s_pc[s] & 0xFF;
      float4 imp = W.s_imp[s];
      float4 old = W.c_imp[i];
      old.x = imp.x; old.y = imp.y;
      if (vcCount == 2) { old.z = imp.z; old.w = imp.w; }
      W.c_imp[i] = old;
    }
    const float h = W.dt;
    for (int b = tid; b < W.nBodies; b += nth) {
      uint32_t f = W.b_flags[b];
      if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
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
    }
  }
  GB();
  // position iterations: contacts then joints, each island stops once all of its constraints are within tolerance (:206-224)
  for (int it = 0; it < W.posIters; ++it) {
    int* notOk = W.b_posNotOk + it * W.nBodies;
    const int* prev = it > 0 ? W.b_posNotOk + (it - 1) * W.nBodies : nullptr;
    for (int c = 0; c < T; ++c) {
      int beg = coff[c], end = coff[c + 1];
      if (beg == end) continue;
      if (contactRole) for (int s = beg + ctid; s < end; s += cnth) {
        int root = W.s_root[s];
        if (prev && __ldcg(&prev[root]) == 0) continue;
        float minSep = contact_solve_position(W, s);
        if (!(minSep >= -3.0f * kLinearSlop)) notOk[root] = 1;
      }
      GB();
    }
    if (haveTail) {
      if (tailBlock) for (int c = T; c < nColours; ++c) {
        for (int s = coff[c] + threadIdx.x; s < coff[c + 1]; s += blockDim.x) {
          int root = W.s_root[s];
          if (prev && __ldcg(&prev[root]) == 0) continue;
          float minSep = contact_solve_position(W, s);
          if (!(minSep >= -3.0f * kLinearSlop)) notOk[root] = 1;
        }
        __syncthreads();
      }
      GB();
    }
    for (int c = 0; c < nJointColours; ++c) {
      int beg = joff[c], end = joff[c + 1];
      if (beg == end) continue;
      if (jointRole) for (int k = beg + jtid; k < end; k += jnth) {
        const int j = k;
        int root = W.j_root[j];
        if (root < 0) continue;
        if (prev && __ldcg(&prev[root]) == 0) continue;
        if (!joint_solve_position(W, j)) notOk[root] = 1;
      }
      GB();
    }
  }
  // write back + SynchronizeTransform (:227-235), sleep bookkeeping (:241-269), ClearForces (b2world.d:443-450)
  {
    const float h = W.dt;
    const float linTolSqr = kLinearSleepTolerance * kLinearSleepTolerance;
    const float angTolSqr = kAngularSleepTolerance * kAngularSleepTolerance;
    for (int b = tid; b < W.nBodies; b += nth) {
      uint32_t f = W.b_flags[b];
      if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
      float4 pos = ldcg4(&W.b_pos[b]);
      float4 lc = W.b_lc[b];
      Xf xf = xf_from_sweep(V(pos.x, pos.y), pos.z, V(lc.x, lc.y));
      W.b_xf[b] = pack(xf);
      if (W.allowSleep) {
        float4 vel = ldcg4(&W.b_vel[b]);
        float2 gs = W.b_gs[b];
        if (!(f & BF_AUTOSLEEP) || vel.z * vel.z > angTolSqr || dot(V(vel.x, vel.y), V(vel.x, vel.y)) > linTolSqr) gs.y = 0.0f;
        else gs.y += h;
        W.b_gs[b] = gs;
        atomicMin(&W.b_islMinSleep[W.b_root[b]], __float_as_int(gs.y));
      }
    }
  }
  GB();
  if (W.allowSleep) {
    const int* last = W.posIters > 0 ? W.b_posNotOk + (W.posIters - 1) * W.nBodies : nullptr;
    for (int b = tid; b < W.nBodies; b += nth) {
      uint32_t f = W.b_flags[b];
      if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
      int root = W.b_root[b];
      bool positionSolved = last && __ldcg(&last[root]) == 0;
      float minSleep = __int_as_float(__ldcg(&W.b_islMinSleep[root]));
      if (minSleep >= kTimeToSleep && positionSolved) {
        // b2Body.SetAwake(false) (b2body.d:837-845)
        W.b_flags[b] = f & ~BF_AWAKE;
        W.b_gs[b].y = 0.0f;
        W.b_vel[b] = make_float4(0, 0, 0, 0);
        W.b_force[b] = make_float4(0, 0, 0, 0);
      }
    }
  }
}

#undef GB
#undef PHASE_MARK
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
      if (slot < W.moveCap) W.moveList[slot] = p; else W.hdr->error = -5;
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
  W.hdr->nPairs = 0;
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
constexpr int kQueryStack = 192;
DBX_D void query_proxy(const DevWorld& W, const int* leaves, int* stack, int lane, int p) {
  const int n = W.nProxies;
  if (!(W.p_flags[p] & PF_ALIVE)) return;
  const Box fat = BX(__ldcg(&W.p_fat[p]));
  const int keyP = W.p_key[p];
  const int worldP = W.b_world[W.p_ids[p].z];
  int top = 1;
  if (lane == 0) stack[0] = (n == 1) ? (n - 1) : 0;
  __syncwarp();
  while (top > 0) {
    int take = min(top, 32);   // pop up to 32 nodes
    int node = lane < take ? stack[top - take + lane] : -1;
    top -= take;
    __syncwarp();
    bool hit = false;
    bool isLeaf = node >= n - 1;
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
          else W.hdr->error = -5;
        }
      }
    }
    bool push = hit && !isLeaf;
    unsigned ballot = __ballot_sync(0xffffffffu, push);
    int offset = __popc(ballot & ((1u << lane) - 1));
    int total = __popc(ballot);
    if (top + 2 * total > kQueryStack) { if (lane == 0) W.hdr->error = -5; break; }   // frontier overflow: report, never corrupt
    if (push) {
      int2 ch = W.bv_child[node];
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
  for (int mIdx = warp; mIdx < nMoved; mIdx += nwarps) query_proxy(W, leaves, stacks[wib], lane, W.moveList[mIdx]);
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
  if (!filter_should_collide(W, pa.x, pb.x)) return;
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
  if (slot >= W.cCap) { W.hdr->error = -5; atomicSub(&W.hdr->cHigh, 1); return; }
  const bool sensor = ((W.f_group[ia.x] >> 16) & FXF_SENSOR) || ((W.f_group[ib.x] >> 16) & FXF_SENSOR);
  const float2 mA = W.f_mat[ia.x], mB = W.f_mat[ib.x];
  W.c_key[slot] = key;
  W.c_ids[slot] = make_int4(proxyA, proxyB, ia.z, ib.z);
  W.c_fix[slot] = make_int4(ia.x, ib.x, ia.w, ib.w);
  W.c_flags[slot] = CF_ALIVE | CF_ENABLED | (sensor ? CF_SENSOR : 0);
  W.c_m0[slot] = make_float4(0, 0, 0, 0);
  W.c_m1[slot] = make_float4(0, 0, 0, 0);
  W.c_imp[slot] = make_float4(0, 0, 0, 0);
  W.c_mk[slot] = make_uint4(0, 0, 0, 0);
  // b2MixFriction / b2MixRestitution (b2contact.d:32-42)
  W.c_mat[slot] = make_float4(sqrtf(mA.x * mB.x), mA.y > mB.y ? mA.y : mB.y, 0.0f, 1.0f);
  W.c_toiCount[slot] = 0;
  W.c_colour[slot] = -1;
  if (!hash_insert(W, key, slot)) W.hdr->error = -5;
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
  GRID_STRIDE(i, n) if (W.c_flags[i] & CF_ALIVE) { if (!hash_insert(W, W.c_key[i], i)) W.hdr->error = -5; }
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
  if (blockIdx.x == 0 && threadIdx.x == 0) { W.hdr->nMoved = nMoved * copies; if (nMoved * copies > W.moveCap) W.hdr->error = -5; }
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
    hash_remove(W, W.c_key[i]);
    W.c_flags[i] = 0;
    W.c_colour[i] = -1;
    int slot = atomicAdd(&W.hdr->nFree, 1);
    W.c_free[slot] = i;
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
    float* bx = (float*)&W.bv_box[node];
    atomic_min_f(bx + 0, f.x); atomic_min_f(bx + 1, f.y); atomic_max_f(bx + 2, f.z); atomic_max_f(bx + 3, f.w);
    node = W.bv_parent[node];
  }
}
__global__ void __launch_bounds__(256) k_lbvh_enlarge(const __grid_constant__ DevWorld W) {
  const int nMoved = min(W.hdr->nMoved, W.moveCap);
  GRID_STRIDE(k, nMoved) lbvh_enlarge(W, W.moveList[k]);
}

// event priority: earlier alpha first; exact ties on a shared body are broken by a hash of the replica-local pair key, so
// the choice does not depend on slot numbers (which atomics hand out in a run-dependent order)
DBX_D unsigned long long toi_prio(const DevWorld& W, float alpha, int contact, int bodyOfContact) {
  return ((unsigned long long)__float_as_uint(alpha) << 32) | (unsigned)(mix64(local_key(W, W.c_key[contact], bodyOfContact)) >> 32);
}

// (a) evaluate b2TimeOfImpact for every eligible contact that has no cached value (b2world.d:1155-1265)
DBX_D void toi_evaluate(const DevWorld& W, int i) {
  uint32_t flags = W.c_flags[i];
  if (!(flags & CF_ALIVE)) return;
  const int4 ids = W.c_ids[i];
  if ((flags & CF_TOI) && ((__ldcg(&W.b_toiFlags[ids.z]) | __ldcg(&W.b_toiFlags[ids.w])) & TF_INVAL)) { flags &= ~(CF_TOI | CF_ISLAND); W.c_flags[i] = flags; }
  if (!(flags & CF_ENABLED)) return;
  if (W.c_toiCount[i] > kMaxSubSteps) return;
  float alpha = 1.0f;
  const uint32_t fa = W.b_flags[ids.z], fb = W.b_flags[ids.w];
  if (flags & CF_TOI) {
    alpha = W.c_mat[i].w;
  } else {
    if (flags & CF_SENSOR) return;
    const int typeA = body_type(fa), typeB = body_type(fb);
    const bool activeA = (fa & BF_AWAKE) && typeA != BODY_STATIC, activeB = (fb & BF_AWAKE) && typeB != BODY_STATIC;
    if (!activeA && !activeB) return;
    const bool collideA = (fa & BF_BULLET) || typeA != BODY_DYNAMIC, collideB = (fb & BF_BULLET) || typeB != BODY_DYNAMIC;
    if (!collideA && !collideB) return;
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
    int state = time_of_impact(&beta, pA, sA, pB, sB, 1.0f);
    if (state == TOI_TOUCHING) alpha = fminr(alpha0 + (1.0f - alpha0) * beta, 1.0f);
    else alpha = 1.0f;
    float4 mat = W.c_mat[i]; mat.w = alpha; W.c_mat[i] = mat;
    flags |= CF_TOI;
    W.c_flags[i] = flags;
  }
  if (1.0f - 10.0f * kEpsilon < alpha) return;   // never becomes an event (:1267-1272)
  const unsigned long long prio = toi_prio(W, alpha, i, ids.z);
  if (body_type(fa) != BODY_STATIC) atomicMin(&W.b_toiMin[ids.z], prio);
  if (body_type(fb) != BODY_STATIC) atomicMin(&W.b_toiMin[ids.w], prio);
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
  for (int k = 0; k < nc; ++k) prepare_contact(W, sBase + k, contacts[k], -1.0f);
  for (int it = 0; it < 20; ++it) {
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
  for (int it = 0; it < W.velIters; ++it) for (int k = 0; k < nc; ++k) contact_solve_velocity(W, sBase + k);
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
  __shared__ int stacks[16][kQueryStack];
  Header* H = W.hdr;
  const unsigned nb = gridDim.x;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warp = tid >> 5, nwarps = nth >> 5;
  const int eventCap = W.eventCap;
  int tmark = 0;
#define TMARK() do { if (W.phaseTimes && tid == 0 && tmark < 64) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); W.phaseTimes[3000 + tmark++] = t_; } } while (0)
  TMARK();
  // reset (m_stepComplete is always true here: sub-stepping is not supported) (:1131-1146)
  {
    const int n = H->cHigh;
    for (int i = tid; i < n; i += nth) {
      uint32_t f = W.c_flags[i];
      if (!(f & CF_ALIVE)) continue;
      if (f & (CF_TOI | CF_ISLAND)) W.c_flags[i] = f & ~(CF_TOI | CF_ISLAND);
      W.c_toiCount[i] = 0;
      float4 mat = W.c_mat[i]; if (mat.w != 1.0f) { mat.w = 1.0f; W.c_mat[i] = mat; }
    }
    for (int b = tid; b < W.nBodies; b += nth) {
      float4 p0 = W.b_pos0[b]; if (p0.w != 0.0f) { p0.w = 0.0f; W.b_pos0[b] = p0; }
      W.b_toiMin[b] = ~0ull; W.b_toiOther[b] = ~0ull; W.b_toiEvt[b] = -1; W.b_toiFlags[b] = 0;
    }
    if (tid == 0) H->nEvents = 0;
  }
  grid_barrier(&H->barrier, nb); TMARK();
  for (int pass = 0; pass < 1024; ++pass) {
    // (a) TOI evaluation + per-body minima
    {
      const int n = *((volatile int*)&H->cHigh);
      for (int i = tid; i < n; i += nth) toi_evaluate(W, i);
    }
    grid_barrier(&H->barrier, nb); TMARK();
    // (b) winners: the minimum on every movable body they touch
    {
      const int n = *((volatile int*)&H->cHigh);
      for (int i = tid; i < n; i += nth) {
        const uint32_t flags = W.c_flags[i];
        if ((flags & (CF_ALIVE | CF_ENABLED | CF_TOI)) != (CF_ALIVE | CF_ENABLED | CF_TOI)) continue;
        if (W.c_toiCount[i] > kMaxSubSteps) continue;
        const float alpha = W.c_mat[i].w;
        if (1.0f - 10.0f * kEpsilon < alpha) continue;
        const int4 ids = W.c_ids[i];
        const unsigned long long prio = toi_prio(W, alpha, i, ids.z);
        const bool movA = body_type(W.b_flags[ids.z]) != BODY_STATIC, movB = body_type(W.b_flags[ids.w]) != BODY_STATIC;
        if ((movA && __ldcg(&W.b_toiMin[ids.z]) != prio) || (movB && __ldcg(&W.b_toiMin[ids.w]) != prio)) continue;
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
      const int n = *((volatile int*)&H->cHigh);
      for (int i = tid; i < n; i += nth) {
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
          if (slot < kToiCand) W.e_cand[(2 * e + evSide) * kToiCand + slot] = i; else H->error = -5;
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
    {
      // the step's LBVH is still valid for every proxy that did not move; widen it for the ones that did
      const int nMoved = min(*((volatile int*)&H->nMoved), W.moveCap);
      for (int k = tid; k < nMoved; k += nth) lbvh_enlarge(W, W.moveList[k]);
      for (int b = tid; b < W.nBodies; b += nth) {
        W.b_toiMin[b] = ~0ull; W.b_toiOther[b] = ~0ull; W.b_toiEvt[b] = -1;
        if (W.b_toiFlags[b] & TF_SYNC) W.b_toiFlags[b] &= ~TF_SYNC;
      }
      if (tid == 0) H->nEvents = 0;
    }
    grid_barrier(&H->barrier, nb); TMARK();
    {
      const int nMoved = min(*((volatile int*)&H->nMoved), W.moveCap);
      for (int k = warp; k < nMoved; k += nwarps) query_proxy(W, W.bv_sorted, stacks[wib], lane, W.moveList[k]);
    }
    grid_barrier(&H->barrier, nb); TMARK();
    {
      const int nPairs = min(*((volatile int*)&H->nPairs), W.pairCap);
      for (int k = tid; k < nPairs; k += nth) add_pair(W, W.pairs[k]);
      const int nMoved = min(*((volatile int*)&H->nMoved), W.moveCap);
      for (int k = tid; k < nMoved; k += nth) { const int p = W.moveList[k]; W.p_flags[p] &= ~PF_MOVED; }
    }
    grid_barrier(&H->barrier, nb); TMARK();
    if (tid == 0) H->nMoved = 0;
    grid_barrier(&H->barrier, nb); TMARK();
  }
}

#undef TMARK
// ------------------------------------------------------------------------------------------------ host launchers
#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

static cudaError_t launch_coop(const void* fn, const DevWorld& W, const LaunchCfg& L) {
  CK(cudaMemsetAsync(&W.hdr->barrier, 0, sizeof(unsigned), L.stream));
  void* args[] = {(void*)&W};
  ++L.launches;
  return cudaLaunchCooperativeKernel(fn, dim3(L.coopBlocks), dim3(L.coopThreads), args, 0, L.stream);
}

cudaError_t stage_collide(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_collide<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}

cudaError_t stage_islands_and_integrate(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_island_init<<<L.gridWide, 256, 0, L.stream>>>(W);
  ++L.launches; k_island_union<<<L.gridWide, 256, 0, L.stream>>>(W);
  ++L.launches; k_island_flatten<<<L.gridWide, 256, 0, L.stream>>>(W);
  ++L.launches; k_island_wake_integrate<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
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

cudaError_t stage_prepare(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_prepare<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}
cudaError_t stage_solve(const DevWorld& W, const LaunchCfg& L) {
  return launch_coop((const void*)k_solve, W, L);
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

cudaError_t stage_find_new_contacts(DevWorld& W, const LaunchCfg& L, bool rebuild) {
  const int n = W.nProxies;
  ++L.launches; k_bounds_init<<<1, 1, 0, L.stream>>>(W);
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
  ++L.launches; k_clear_moves<<<L.gridWide, 256, 0, L.stream>>>(W);
  ++L.launches; k_reset_moves<<<1, 1, 0, L.stream>>>(W);
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
