// dbx_util.cuh — small device utilities shared by the kernel translation units (pair hash, union-find, grid barrier, filters).
#pragma once
#include "dbx_kernels.cuh"

namespace dbx {

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
// Keys are unique per insertion batch and the caller has already looked the key up (it is not in the table), so a CAS on
// the key word is enough and the first tombstone on the probe path can be recycled: without that a scene that keeps
// creating and destroying contacts fills the table with tombstones between two rebuilds.
DBX_D bool hash_insert(const DevWorld& W, unsigned long long key, int val) {
  unsigned mask = (unsigned)W.hCap - 1;
  unsigned h = (unsigned)mix64(key) & mask;
  for (int probe = 0; probe < W.hCap; ++probe) {
    unsigned long long cur = W.h_key[h];
    if (cur == kHashEmpty || cur == kHashTomb) {
      unsigned long long old = atomicCAS(&W.h_key[h], cur, key);
      if (old == cur) { W.h_val[h] = val; if (cur == kHashTomb) atomicSub(&W.hdr->nTomb, 1); return true; }
    }
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

// lock-free union-find.  Roots are hooked by a bijective hash of the body id (smaller hash wins), not by the id itself:
// a stack of consecutively numbered bodies would otherwise hook into one chain as deep as the stack, and the dependent
// loads of walking it are what the island pass costs.  A component's root is still a pure function of its member set.
DBX_D unsigned uf_rank(int x, bool hashed) { return hashed ? (unsigned)x * 0x9E3779B1u : (unsigned)x; }
DBX_D int uf_find(int* parent, int x) {
  for (;;) {
    int p = parent[x];
    if (p == x) return x;
    int gp = parent[p];
    if (gp != p) parent[x] = gp;  // path halving (benign race)
    x = p;
  }
}
// plain variant (finds with path halving, one after the other): better for the thousands of small islands of batched worlds
DBX_D void uf_unite_small(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (uf_rank(a, true) < uf_rank(b, true)) { int t = a; a = b; b = t; }
    if (atomicCAS(&parent[a], a, b) == a) return;
  }
}
DBX_D void uf_unite(int* parent, int a, int b, bool hashed) {
  const int a0 = a, b0 = b;
  for (;;) {
    // both root walks in lock-step: two independent loads in flight per hop instead of one
    for (;;) {
      const int pa = parent[a], pb = parent[b];
      if (pa == a && pb == b) break;
      a = pa; b = pb;
    }
    if (a == b) break;
    if (uf_rank(a, hashed) < uf_rank(b, hashed)) { int t = a; a = b; b = t; }
    if (atomicCAS(&parent[a], a, b) == a) { a = b; break; }
  }
  // shortcut the two starting points to the root just found (always one of their ancestors)
  if (a0 != a && parent[a0] != a0) parent[a0] = a;
  if (b0 != a && parent[b0] != b0) parent[b0] = a;
}

// b2Profile's solveInit / solveVelocity / solvePosition split (b2timestep.d:37-47): the island solvers stamp their phases
DBX_D void solve_stamp(const DevWorld& W, int k) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); W.hdr->solveStamp[k] = t; }
}
// global barrier for the persistent kernels: all CTAs are co-resident (cooperative launch, one per SM)
DBX_D void grid_barrier(unsigned* counter, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned ticket = atomicAdd(counter, 1u);
    unsigned target = (ticket / nblocks + 1u) * nblocks;
    while ((int)(*((volatile unsigned*)counter) - target) < 0) { }   // wrap-safe: the distance to the target is always < 2^31
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
  if (W.nJointPairs > 0 && ((W.jp_bits[bA >> 5] >> (bA & 31)) & (W.jp_bits[bB >> 5] >> (bB & 31)) & 1u)) {
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

}  // namespace dbx
