// dbx_colour.cuh — solver membership + graph colouring of the contacts, as device functions shared by the stand-alone kernels
// (dbx_kernels.cu: k_mark_solve, k_colour) and by the tile solver's fused colour-and-sort kernel (dbx_tiles.cu).
#pragma once
#include "dbx_util.cuh"
#include "dbx_tilekey.cuh"

namespace dbx {

// mark the contacts the solver takes this step and rebuild the per-body colour masks from the persistent colours
DBX_D void mark_solve_body(const DevWorld& W) {
  if (W.tiled) {
    __shared__ int sjoff[kMaxJointColours + 1];
    for (int c = threadIdx.x; c <= kMaxJointColours; c += blockDim.x) sjoff[c] = W.hdr->jointColourOff[c];
    __syncthreads();
    tile_key_joints(W, sjoff);
    // the masses in tile order, for the solver's prologue (a body's mass may change between any two steps)
    // and "takes part in this step", so that the solver need not gather the body flags
    GRID_STRIDE(p, W.nTileBodies) {
      const int b = W.t_body[p];
      const float4 ms = W.b_mass[b];
      W.t_mass[p] = make_float2(ms.x, ms.y);
      if ((W.b_flags[b] & (BF_ALIVE | BF_ISLAND)) == (BF_ALIVE | BF_ISLAND)) atomicOr(&W.b_xflag[p], XF_ISLAND);
    }
  }
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
    // a colour kept from an earlier step is void if a joint has since claimed it (or a higher one) on either body
    if (W.unifiedColours && col >= 0 && col < kMaskColours &&
        ((((body_type(fa) == BODY_DYNAMIC ? W.b_jmask[ids.z] : 0ull) | (body_type(fb) == BODY_DYNAMIC ? W.b_jmask[ids.w] : 0ull)) >> col) & 1ull)) col = -1;
    if (col >= 0 && col < kMaskColours) {
      if (body_type(fa) == BODY_DYNAMIC) atomicOr(&W.b_mask[ids.z], 1ull << col);
      if (body_type(fb) == BODY_DYNAMIC) atomicOr(&W.b_mask[ids.w], 1ull << col);
      if (W.tiled) tile_key_contact(W, i, ids.z, ids.w, col);
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
constexpr int kColourSoloMax = 8192;
// pair key with the replica offset removed, so that every replica of a batched world arbitrates (and hence colours) alike
DBX_D unsigned long long local_key(const DevWorld& W, unsigned long long key, int body) {
  if (W.keyStride == 0) return key;
  const unsigned long long o = (unsigned long long)(unsigned)(W.b_world[body] * W.keyStride);
  return key - (o << 32) - o;
}
// (every CTA of a cooperative launch calls this; it synchronises the grid through Header::barrier)
// A short worklist (the few hundred contacts a settled scene gains per step) is coloured by CTA 0 alone: its rounds then meet at
// block barriers instead of grid barriers (three per round, ~3 us each).  Every CTA takes the same decision: the two header words
// it depends on are not written while the kernel runs (k_island_init resets / advances them at the start of the next step).
DBX_D void colour_body(const DevWorld& W) {
  Header* H = W.hdr;
  const unsigned nb = gridDim.x;
  int* cur = W.c_work; int* nxt = W.c_work2;
  int n = H->nUncoloured;
  unsigned epoch = H->epoch;
  const bool solo = n <= kColourSoloMax && epoch <= 0xF0000u;
  if (solo && blockIdx.x != 0) return;
  const int tid = solo ? (int)threadIdx.x : (int)(blockIdx.x * blockDim.x + threadIdx.x), nth = solo ? (int)blockDim.x : (int)(gridDim.x * blockDim.x);
#define COLOUR_SYNC() do { if (solo) __syncthreads(); else grid_barrier(&H->barrier, nb); } while (0)
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
    COLOUR_SYNC();
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
          if (col >= kMaxColours) { col = kMaxColours - 1; H->error = E_COLOURS; }
        }
        W.c_colour[i] = col;
        if (W.tiled) tile_key_contact(W, i, ids.z, ids.w, col);
        if (col > *((volatile int*)&H->maxColour)) atomicMax(&H->maxColour, col);
      } else {
        int slot = atomicAdd(&H->nUncoloured2, 1);
        nxt[slot] = i;
      }
    }
    COLOUR_SYNC();
    n = *((volatile int*)&H->nUncoloured2);
    int* t = cur; cur = nxt; nxt = t;
    COLOUR_SYNC();
  }
  if (tid == 0) H->epochNext = epoch;
#undef COLOUR_SYNC
}

}  // namespace dbx
