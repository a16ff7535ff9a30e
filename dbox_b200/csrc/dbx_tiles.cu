// dbx_tiles.cu — the tile solver: b2Island.Solve (dynamics/b2island.d:118-279) for one big world with the bodies of a spatial
// tile resident in ONE CTA's shared memory.
//
// k_solve (dbx_solve.cu) runs every colour of every Gauss-Seidel pass as a grid-wide phase: ~104 dependent phases of ~4 us on
// the 100,000-body pile, each a grid barrier plus an L2 round trip for two bodies.  Here the dynamic bodies are sorted along x
// and cut into P tiles of T bodies (P <= one CTA per SM).  A constraint whose dynamic bodies sit in one tile is LOCAL (class L):
// its CTA walks the local colours with __syncthreads() between them, velocities and positions in shared memory.  A constraint
// between tiles s and s + 1 whose two bodies are both claimed by boundary s is a BOUNDARY constraint (class B), solved by CTA s
// with its own body in shared memory and the neighbour's body in the global arrays; everything else (a body reaching over two
// tiles, a gear joint, an overflow colour of a hub body) is GLOBAL (class G) and runs as grid-wide colour phases like k_solve's.
// One pass = L colours (block barriers) -> publish the exchange bodies -> grid barrier -> B colours (block barriers)
// [-> publish -> grid barrier -> G colours, a grid barrier each] -> grid barrier -> read the exchange bodies back:
// two grid barriers per pass instead of one per colour.  Position passes run the same schedule backwards, so that a body
// still meets its contacts before its joints (b2island.d:206-216), as in k_solve's unified phases.
//
// The order is a Gauss-Seidel sweep like any other: within a phase no two constraints share a dynamic body (the global
// colouring is proper, and L / B / G phases never overlap in time on a body: boundary claims are exclusive).  The schedule
// is reported by dbx_world_debug_read_solve_order, and tests hand it to the sequential oracle.
#include <cub/cub.cuh>
#include "dbx_solver.cuh"
#include "dbx_kernels.cuh"

namespace dbx {

#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

DBX_D unsigned ordered_bits(float x) { const unsigned u = __float_as_uint(x); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }

// ------------------------------------------------------------------------------------------------ tile assignment (every few steps)
// key = x of the body's centre for dynamic bodies (monotonic bit pattern), all ones otherwise: after the sort the dynamic bodies
// are the first nTileBodies entries, left to right
__global__ void __launch_bounds__(256) k_tile_body_keys(const __grid_constant__ DevWorld W, unsigned* keys, int* vals) {
  GRID_STRIDE(b, W.nBodies) {
    const uint32_t f = W.b_flags[b];
    const bool dyn = (f & BF_ALIVE) && body_type(f) == BODY_DYNAMIC;
    keys[b] = dyn ? min(ordered_bits(W.b_pos[b].x), 0xFFFFFFFEu) : 0xFFFFFFFFu;
    vals[b] = b;
  }
}
__global__ void __launch_bounds__(256) k_tile_slots(const __grid_constant__ DevWorld W, const unsigned* keys, const int* vals) {
  GRID_STRIDE(p, W.nBodies) {
    const int b = vals[p];
    if (keys[p] != 0xFFFFFFFFu && p < W.nTileBodies) { W.b_tslot[b] = p; W.t_body[p] = b; }
    else W.b_tslot[b] = -1;
  }
}

// ------------------------------------------------------------------------------------------------ classification (every step)
struct TileEnds { int sA, sB, tA, tB; };
DBX_D TileEnds tile_ends(const DevWorld& W, int bA, int bB) {
  TileEnds e; e.sA = W.b_tslot[bA]; e.sB = W.b_tslot[bB];
  e.tA = e.sA >= 0 ? e.sA / W.tileBodies : -1; e.tB = e.sB >= 0 ? e.sB / W.tileBodies : -1;
  return e;
}
// pass 1: every body learns the lowest boundary any of its tile-crossing constraints straddles; the bins are emptied
__global__ void __launch_bounds__(256) k_tile_claim(const __grid_constant__ DevWorld W) {
  const int nBins = 2 * W.nTiles * kTileColours + kMaxColours;
  GRID_STRIDE(k, nBins) { W.t_cur[k] = 0; W.tj_cur[k] = 0; }
  const int n = W.hdr->cHigh;
  GRID_STRIDE(i, n) {
    if (!(W.c_flags[i] & CF_SOLVE)) continue;
    const int4 ids = W.c_ids[i];
    const TileEnds e = tile_ends(W, ids.z, ids.w);
    if (e.tA < 0 || e.tB < 0 || e.tA == e.tB) continue;
    const int lo = min(e.tA, e.tB), hi = max(e.tA, e.tB);
    if (hi - lo == 1) { atomicMin(&W.b_tclaim[ids.z], lo); atomicMin(&W.b_tclaim[ids.w], lo); }
  }
  GRID_STRIDE(j, W.nJoints) {
    if (!joint_active(W, j)) continue;
    const int4 ids = W.j_ids[j];
    if (ids.x == JT_GEAR) continue;
    const TileEnds e = tile_ends(W, ids.y, ids.z);
    if (e.tA < 0 || e.tB < 0 || e.tA == e.tB) continue;
    const int lo = min(e.tA, e.tB), hi = max(e.tA, e.tB);
    if (hi - lo == 1) { atomicMin(&W.b_tclaim[ids.y], lo); atomicMin(&W.b_tclaim[ids.z], lo); }
  }
}
// class, owner tile and the two body references of a constraint between bodies bA, bB with colour `col`; returns the bin
DBX_D int tile_classify(const DevWorld& W, int bA, int bB, int col, bool forceGlobal, int2* bref) {
  const int P = W.nTiles;
  const TileEnds e = tile_ends(W, bA, bB);
  int cls, owner = 0;
  if (forceGlobal || col >= kTileColours || (e.tA < 0 && e.tB < 0)) cls = 2;
  else if (e.tA < 0 || e.tB < 0 || e.tA == e.tB) { cls = 0; owner = e.tA >= 0 ? e.tA : e.tB; }
  else {
    const int lo = min(e.tA, e.tB), hi = max(e.tA, e.tB);
    if (hi - lo == 1 && W.b_tclaim[bA] == lo && W.b_tclaim[bB] == lo) { cls = 1; owner = lo; } else cls = 2;
  }
  if (cls == 2) {
    if (e.sA >= 0) atomicOr(&W.b_xflag[bA], XF_G);
    if (e.sB >= 0) atomicOr(&W.b_xflag[bB], XF_G);
    *bref = make_int2(bA | kRefGlobal, bB | kRefGlobal);
    return 2 * P * kTileColours + min(col, kMaxColours - 1);
  }
  if (cls == 1) {
    atomicOr(&W.b_xflag[bA], e.tA == owner ? XF_OWNB : XF_FOREIGN);
    atomicOr(&W.b_xflag[bB], e.tB == owner ? XF_OWNB : XF_FOREIGN);
  }
  *bref = make_int2(e.tA == owner ? e.sA : (bA | kRefGlobal), e.tB == owner ? e.sB : (bB | kRefGlobal));
  return (cls * P + owner) * kTileColours + col;
}
// pass 2: bin and body references per solver contact / active joint; histogram of the bins
__global__ void __launch_bounds__(256) k_tile_key(const __grid_constant__ DevWorld W) {
  // joint slots are colour-major (World::recolourJoints): the colour of slot j is the range of jointColourOff it falls into
  __shared__ int sjoff[kMaxJointColours + 1];
  for (int c = threadIdx.x; c <= kMaxJointColours; c += blockDim.x) sjoff[c] = W.hdr->jointColourOff[c];
  __syncthreads();
  const int n = W.hdr->cHigh;
  GRID_STRIDE(i, n) {
    if (!(W.c_flags[i] & CF_SOLVE)) continue;
    const int4 ids = W.c_ids[i];
    int2 br;
    const int bin = tile_classify(W, ids.z, ids.w, W.c_colour[i], false, &br);
    W.c_tkey[i] = bin; W.c_bref[i] = br; W.c_tcol[i] = -1;
    atomicAdd(&W.t_cur[bin], 1);
  }
  GRID_STRIDE(j, W.nJoints) {
    if (!joint_active(W, j)) { W.j_tkey[j] = -1; W.j_root[j] = -1; continue; }
    const int4 ids = W.j_ids[j];
    int2 br;
    int lo = 0, hi = kMaxJointColours;          // largest c with sjoff[c] <= j
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (sjoff[mid] <= j) lo = mid; else hi = mid - 1; }
    const int bin = tile_classify(W, ids.y, ids.z, lo, ids.x == JT_GEAR, &br);
    if (ids.x == JT_GEAR) {   // the far bodies of joint1 / joint2 are written too (b2gearjoint.d:352-386)
      const int4 id2 = W.j_ids2[j];
      if (W.b_tslot[id2.x] >= 0) atomicOr(&W.b_xflag[id2.x], XF_G);
      if (W.b_tslot[id2.y] >= 0) atomicOr(&W.b_xflag[id2.y], XF_G);
    }
    W.j_tkey[j] = bin; W.j_bref[j] = br; W.j_tcol[j] = -1;
    atomicAdd(&W.tj_cur[bin], 1);
  }
}
// one CTA: exclusive scan of both histograms -> offsets; totals and class sizes into the header; cursors back to zero
__global__ void __launch_bounds__(1024) k_tile_scan(const __grid_constant__ DevWorld W) {
  __shared__ int part[2][1024];
  __shared__ int stats[4];     // B constraints, G constraints, highest colour in use + 1
  const int P = W.nTiles, nBins = 2 * P * kTileColours + kMaxColours;
  const int t = threadIdx.x, per = (nBins + 1023) / 1024;
  const int beg = min(t * per, nBins), end = min(beg + per, nBins);
  if (t < 4) stats[t] = 0;
  int sc = 0, sj = 0, nb = 0, ng = 0, maxc = 0;
  for (int k = beg; k < end; ++k) {
    const int c = W.t_cur[k], j = W.tj_cur[k];
    sc += c; sj += j;
    if (c + j > 0) {
      const int col = k < 2 * P * kTileColours ? k % kTileColours : k - 2 * P * kTileColours;
      maxc = max(maxc, col + 1);
      if (k >= 2 * P * kTileColours) ng += c + j; else if (k >= P * kTileColours) nb += c + j;
    }
  }
  part[0][t] = sc; part[1][t] = sj;
  __syncthreads();
  if (nb) atomicAdd(&stats[0], nb);
  if (ng) atomicAdd(&stats[1], ng);
  if (maxc) atomicMax(&stats[2], maxc);
  // Hillis-Steele over the 1024 partial sums (two arrays at once)
  for (int o = 1; o < 1024; o <<= 1) {
    const int a = t >= o ? part[0][t - o] : 0, b = t >= o ? part[1][t - o] : 0;
    __syncthreads();
    part[0][t] += a; part[1][t] += b;
    __syncthreads();
  }
  int oc = part[0][t] - sc, oj = part[1][t] - sj;
  for (int k = beg; k < end; ++k) {
    const int c = W.t_cur[k], j = W.tj_cur[k];
    W.t_off[k] = oc; W.tj_off[k] = oj; oc += c; oj += j;
    W.t_cur[k] = 0; W.tj_cur[k] = 0;
  }
  if (t == 1023) {
    const int total = part[0][1023];
    W.t_off[nBins] = total; W.tj_off[nBins] = part[1][1023];
    W.hdr->nSolve = total;
    if (total > W.sCap) W.hdr->error = E_SOLVER_ROWS;
  }
  __syncthreads();
  if (t == 0) { W.hdr->nTileB = stats[0]; W.hdr->nTileG = stats[1]; W.hdr->nColours = stats[2]; W.hdr->tailStart = stats[2]; }
}
__global__ void __launch_bounds__(256) k_tile_scatter(const __grid_constant__ DevWorld W) {
  const int n = W.hdr->cHigh;
  GRID_STRIDE(i, n) {
    if (!(W.c_flags[i] & CF_SOLVE)) continue;
    const int bin = W.c_tkey[i];
    const int s = W.t_off[bin] + atomicAdd(&W.t_cur[bin], 1);
    if (s < W.sCap) W.s_contact[s] = i;
  }
  GRID_STRIDE(j, W.nJoints) {
    const int bin = W.j_tkey[j];
    if (bin < 0) continue;
    W.tj_order[W.tj_off[bin] + atomicAdd(&W.tj_cur[bin], 1)] = j;
  }
}

// ------------------------------------------------------------------------------------------------ the solver
enum { TM_INIT = 0, TM_VEL = 1, TM_POS = 2 };
constexpr int kTileBMax = 1024;      // boundary constraints one CTA re-colours locally (more: it walks them by global colour)
constexpr int kTileBColours = 32;

// one constraint of a phase: item >= 0 is a contact row (solver slot), item < 0 a joint (~joint slot).  Deliberately not
// inlined: the kernel below reaches it from its local, boundary and global loops and should hold ONE copy of the row code.
__device__ __noinline__ void tile_item(const DevWorld& W, BodyView view, int mode, int item, int* notOk, const int* prev) {
  if (item >= 0) {
    if (mode == TM_VEL) contact_solve_velocity(W, item, view);
    else if (mode == TM_POS) {
      const int root = W.s_root[item];
      if (prev && __ldcg(&prev[root]) == 0) return;
      const float minSep = contact_solve_position(W, item, -1, -1, view);
      if (!(minSep >= -3.0f * kLinearSlop)) __stcg(&notOk[root], 1);
    }
  } else {
    const int j = ~item;
    if (mode == TM_INIT) joint_init(W, j, view);
    else if (mode == TM_VEL) { if (W.j_root[j] >= 0) joint_solve_velocity(W, j, view); }
    else {
      const int root = W.j_root[j];
      if (root < 0) return;
      if (prev && __ldcg(&prev[root]) == 0) return;
      if (!joint_solve_position(W, j, view)) __stcg(&notOk[root], 1);
    }
  }
}
// pull the row a thread will need in its NEXT phase from L2 into L1 while it works on the current one (rows are read-only
// during the solve except for s_imp, which only the owning thread writes)
DBX_D void pf(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
DBX_D void tile_prefetch_row(const DevWorld& W, int mode, int s) {
  pf(&W.s_body[s]); pf(&W.s_pc[s]); pf(&W.s_v1[s]);
  if (mode == TM_VEL) {
    pf(&W.s_v0[s]); pf(&W.s_r0[s]); pf(&W.s_r1[s]); pf(&W.s_q0[s]); pf(&W.s_q1[s]); pf(&W.s_imp[s]); pf(&W.s_nm[s]); pf(&W.s_k[s]);
  } else {
    pf(&W.s_p0[s]); pf(&W.s_p1[s]); pf(&W.s_p2[s]); pf(&W.s_p3[s]); pf(&W.s_root[s]);
  }
}

__global__ void __launch_bounds__(512) k_solve_tiles(const __grid_constant__ DevWorld W) {
  extern __shared__ float4 sm4[];                 // [T] velocities, [T] positions, then [T] body ids, [T] exchange flags
  __shared__ int offL[kTileColours + 1], joffL[kTileColours + 1], offB[kTileColours + 1], joffB[kTileColours + 1];
  __shared__ int offG[kMaxColours + 1], joffG[kTileColours + 1];
  __shared__ int phL[kTileColours], nPhL;          // the local colours that hold anything, in order
  __shared__ int sBItem[kTileBMax], sBOff[kTileBColours + 1], nPhB, bDirect;
  Header* H = W.hdr;
  const unsigned nb = gridDim.x;
  const int lt = threadIdx.x, ln = blockDim.x;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
  const int P = W.nTiles, T = W.tileBodies;
  const int tile = blockIdx.x;
  const int s0 = tile * T, n = tile < P ? max(0, min(T, W.nTileBodies - s0)) : 0;
  float4* sVel = sm4; float4* sPos = sm4 + T;
  int* sBody = (int*)(sm4 + 2 * T); int* sFlag = sBody + T;
  BodyView view; view.vel = sVel; view.pos = sPos; view.off = s0;
  const BodyView noView;
  {
    const int baseG = 2 * P * kTileColours;
    for (int c = lt; c <= kTileColours; c += ln) {
      if (tile < P) {
        offL[c] = W.t_off[tile * kTileColours + c]; joffL[c] = W.tj_off[tile * kTileColours + c];
        offB[c] = W.t_off[(P + tile) * kTileColours + c]; joffB[c] = W.tj_off[(P + tile) * kTileColours + c];
      } else { offL[c] = offB[c] = joffL[c] = joffB[c] = 0; }
      joffG[c] = W.tj_off[baseG + c];
    }
    for (int c = lt; c <= kMaxColours; c += ln) offG[c] = W.t_off[baseG + c];
  }
  const int nColours = H->nColours;
  const int nG = H->nTileG, nCross = H->nTileB + nG;      // uniform over the grid
  const int nLoc = min(nColours, kTileColours);
#define GB() grid_barrier(&H->barrier, nb)
  // debug (dbx_world_debug_phase_times): %globaltimer stamps of CTA 0 at [0 ..) and of the middle CTA at [1024 ..)
  int markIdx = 0;
  const bool marking = W.phaseTimes != nullptr && lt == 0 && (blockIdx.x == 0 || blockIdx.x == nb / 2);
  const int markBase = blockIdx.x == 0 ? 0 : 1024;
#define MARK() do { if (marking && markIdx < 1000) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); W.phaseTimes[markBase + markIdx++] = t_; } } while (0)
  MARK();
  __syncthreads();
  if (lt == 0) {
    int k = 0;
    for (int c = 0; c < nLoc; ++c) if (offL[c] != offL[c + 1] || joffL[c] != joffL[c + 1]) phL[k++] = c;
    nPhL = k;
  }
  // Boundary constraints: the global colouring spreads a boundary's ~150 rows over every colour in use (ten phases of a
  // dozen rows each); among themselves they need three or four.  Re-colour them greedily, joints first and in ascending global
  // colour (so a body still meets its joints before its contacts), and walk them through an index list in that order.
  {
    const int nBJ = joffB[kTileColours] - joffB[0], nBC = offB[kTileColours] - offB[0], nB = nBJ + nBC;
    int2* sRef = (int2*)sm4;                          // scratch in the body area (the bodies come in afterwards)
    unsigned* sMask = (unsigned*)(sRef + kTileBMax);  // [2T] local colours in use per own / neighbour body
    const bool fits = nB <= kTileBMax;              // (tile_smem_bytes leaves room for this scratch whatever T is)
    if (lt == 0) { bDirect = fits ? 0 : 1; nPhB = 0; }
    if (fits && nB > 0) {
      for (int k = lt; k < 2 * T; k += ln) sMask[k] = 0u;
      for (int k = lt; k < nB; k += ln) {
        int item; int2 br;
        if (k < nBJ) { const int j = W.tj_order[joffB[0] + k]; item = ~j; br = W.j_bref[j]; }
        else { const int s = offB[0] + (k - nBJ); item = s; br = W.c_bref[W.s_contact[s]]; }
        sBItem[k] = item;
        // index of each body in the mask array: own tile [0, T), the right-hand neighbour's [T, 2T), -1 for a body nobody moves
        int ia = -1, ib = -1;
        if (br.x >= 0) ia = br.x - s0; else { const int sl = W.b_tslot[br.x & 0x7fffffff]; if (sl >= 0) ia = sl - s0; }
        if (br.y >= 0) ib = br.y - s0; else { const int sl = W.b_tslot[br.y & 0x7fffffff]; if (sl >= 0) ib = sl - s0; }
        sRef[k] = make_int2(ia, ib);
      }
      __syncthreads();
      if (lt == 0) {
        int count[kTileBColours];
        for (int c = 0; c < kTileBColours; ++c) count[c] = 0;
        bool ok = true;
        for (int k = 0; k < nB && ok; ++k) {
          const int2 r = sRef[k];
          const unsigned used = (r.x >= 0 ? sMask[r.x] : 0u) | (r.y >= 0 ? sMask[r.y] : 0u);
          if (!~used) { ok = false; break; }
          const int lc = __ffs((int)~used) - 1;
          if (r.x >= 0) sMask[r.x] |= 1u << lc;
          if (r.y >= 0) sMask[r.y] |= 1u << lc;
          sRef[k].x = lc;                             // (the mask indices are not needed any more)
          ++count[lc];
        }
        if (!ok) bDirect = 1;
        else {
          int acc = 0, np = 0;
          for (int c = 0; c < kTileBColours; ++c) { sBOff[c] = acc; acc += count[c]; if (count[c]) np = c + 1; count[c] = sBOff[c]; }
          for (int c = np; c <= kTileBColours; ++c) sBOff[c] = acc;
          nPhB = np;
          // stable counting sort of the items by local colour, in place via the scratch: sRef[k].y <- destination
          for (int k = 0; k < nB; ++k) sRef[k].y = count[sRef[k].x]++;
        }
      }
      __syncthreads();
      if (!bDirect) {
        int item = 0, dst = -1, lc = 0;
        // (every thread keeps at most two items: nB <= 1024 = 2 x 512)
        int item2 = 0, dst2 = -1, lc2 = 0;
        if (lt < nB) { item = sBItem[lt]; dst = sRef[lt].y; lc = sRef[lt].x; }
        if (lt + ln < nB) { item2 = sBItem[lt + ln]; dst2 = sRef[lt + ln].y; lc2 = sRef[lt + ln].x; }
        __syncthreads();
        if (dst >= 0) { sBItem[dst] = item; if (item >= 0) W.c_tcol[W.s_contact[item]] = lc; else W.j_tcol[~item] = lc; }
        if (dst2 >= 0) { sBItem[dst2] = item2; if (item2 >= 0) W.c_tcol[W.s_contact[item2]] = lc2; else W.j_tcol[~item2] = lc2; }
      }
    }
    __syncthreads();
  }
  const bool bLocal = bDirect == 0;

  // bodies in: velocities with the contacts' warm start folded in (see k_solve), positions, ids, exchange flags
  {
    const float k = 1.0f / 4294967296.0f;
    for (int i = lt; i < n; i += ln) {
      const int b = W.t_body[s0 + i];
      float4 vel = ldcg4(&W.b_vel[b]);
      if (W.warmStarting) {
        const long long ax = (long long)__ldcg(&W.b_acc[3 * b]), ay = (long long)__ldcg(&W.b_acc[3 * b + 1]), aw = (long long)__ldcg(&W.b_acc[3 * b + 2]);
        if ((ax | ay | aw) != 0) {
          vel.x += (float)ax * k; vel.y += (float)ay * k; vel.z += (float)aw * k;
          __stcg(&W.b_acc[3 * b], 0ull); __stcg(&W.b_acc[3 * b + 1], 0ull); __stcg(&W.b_acc[3 * b + 2], 0ull);
        }
      }
      sVel[i] = vel; sPos[i] = ldcg4(&W.b_pos[b]);
      sBody[i] = b; sFlag[i] = W.b_xflag[b];
    }
    __syncthreads();
  }
  MARK();

  int sweepNo = 0;
  // ---- the phases of one class
  // local colour `c` of this tile: joints, then rows; the thread's row of the NEXT non-empty colour is prefetched meanwhile
  auto local_phase = [&](int mode, int k, int kNext, int* notOk, const int* prev) {
    const int c = phL[k];
    const int jb = joffL[c], nj = joffL[c + 1] - jb, beg = offL[c], total = nj + (offL[c + 1] - beg);
    if (mode == TM_INIT && nj == 0) return;
    if (kNext >= 0 && mode != TM_INIT) {
      const int cn = phL[kNext];
      const int sn = offL[cn] + lt - (joffL[cn + 1] - joffL[cn]);
      if (sn >= offL[cn] && sn < offL[cn + 1]) tile_prefetch_row(W, mode, sn);
    }
    const bool fine = marking && sweepNo == 4 && k < 16;       // debug: clock stamps of one velocity pass at [2048 + 4 k ..)
    long long c0 = 0, c1 = 0;
    if (fine) c0 = clock64();
    for (int q = lt; q < (mode == TM_INIT ? nj : total); q += ln) tile_item(W, view, mode, q < nj ? ~W.tj_order[jb + q] : beg + (q - nj), notOk, prev);
    if (fine) c1 = clock64();
    __syncthreads();
    if (fine) { const long long c2 = clock64(); unsigned long long* o = W.phaseTimes + 2048 + (blockIdx.x == 0 ? 0 : 128) + 4 * k; o[0] = (unsigned long long)(c1 - c0); o[1] = (unsigned long long)(c2 - c1); o[2] = (unsigned long long)total; o[3] = (unsigned long long)c; }
  };
  auto boundary_phases = [&](int mode, bool backwards, int* notOk, const int* prev) {
    if (bLocal) {
      for (int k = 0; k < nPhB; ++k) {
        const int c = backwards ? nPhB - 1 - k : k;
        for (int q = sBOff[c] + lt; q < sBOff[c + 1]; q += ln) { const int item = sBItem[q]; if (mode != TM_INIT || item < 0) tile_item(W, view, mode, item, notOk, prev); }
        __syncthreads();
      }
    } else {
      for (int k = 0; k < nLoc; ++k) {
        const int c = backwards ? nLoc - 1 - k : k;
        const int jb = joffB[c], nj = joffB[c + 1] - jb, beg = offB[c], total = nj + (offB[c + 1] - beg);
        if (total == 0 || (mode == TM_INIT && nj == 0)) continue;
        for (int q = lt; q < (mode == TM_INIT ? nj : total); q += ln) tile_item(W, view, mode, q < nj ? ~W.tj_order[jb + q] : beg + (q - nj), notOk, prev);
        __syncthreads();
      }
    }
  };
  auto global_phases = [&](int mode, bool backwards, int* notOk, const int* prev) {
    for (int k = 0; k < nColours; ++k) {
      const int c = backwards ? nColours - 1 - k : k;
      const int jb = c < kTileColours ? joffG[c] : 0, nj = c < kTileColours ? joffG[c + 1] - jb : 0, beg = offG[c], total = nj + (offG[c + 1] - beg);
      if (total == 0 || (mode == TM_INIT && nj == 0)) continue;
      for (int q = gt; q < (mode == TM_INIT ? nj : total); q += gn) tile_item(W, noView, mode, q < nj ? ~W.tj_order[jb + q] : beg + (q - nj), notOk, prev);
      GB();
    }
  };
  // exchange bodies between shared memory and the global arrays, selected by exchange flags
  auto publish = [&](bool positions, int need, int both) {
    for (int i = lt; i < n; i += ln) {
      const int f = sFlag[i];
      if (!(f & need) || (f & both) != both) continue;
      if (positions) stcg4(&W.b_pos[sBody[i]], sPos[i]); else stcg4(&W.b_vel[sBody[i]], sVel[i]);
    }
  };
  // one Gauss-Seidel pass.  forward: L, B, G with the colours upwards; backward (position passes): G, B, L downwards
  auto sweep = [&](int mode, int* notOk, const int* prev) {
    const bool pos = mode == TM_POS;
    ++sweepNo;
    if (!pos) {
      for (int k = 0; k < nPhL; ++k) local_phase(mode, k, k + 1 < nPhL ? k + 1 : -1, notOk, prev);
      MARK();
      if (nCross == 0) return;
      publish(false, XF_FOREIGN | XF_G, 0);
      GB();
      MARK();
      boundary_phases(mode, false, notOk, prev);
      MARK();
      if (nG > 0) {
        publish(false, XF_G, XF_OWNB | XF_G);
        GB();
        global_phases(mode, false, notOk, prev);
      } else GB();
      for (int i = lt; i < n; i += ln) if (sFlag[i] & (XF_FOREIGN | XF_G)) sVel[i] = ldcg4(&W.b_vel[sBody[i]]);
      __syncthreads();
      MARK();
    } else {
      if (nCross > 0) {
        publish(true, XF_FOREIGN | XF_G, 0);
        GB();
        if (nG > 0) {
          global_phases(mode, true, notOk, prev);
          for (int i = lt; i < n; i += ln) { const int f = sFlag[i]; if ((f & (XF_OWNB | XF_G)) == (XF_OWNB | XF_G)) sPos[i] = ldcg4(&W.b_pos[sBody[i]]); }
          __syncthreads();
        }
        boundary_phases(mode, true, notOk, prev);
        GB();
        // the neighbour's boundary pass and the global phases wrote the global copy; a body this tile's own boundary rows
        // moved after the global phases is newest in shared memory
        for (int i = lt; i < n; i += ln) {
          const int f = sFlag[i];
          if ((f & XF_FOREIGN) || ((f & XF_G) && !(f & XF_OWNB))) sPos[i] = ldcg4(&W.b_pos[sBody[i]]);
        }
        __syncthreads();
      }
      for (int k = nPhL - 1; k >= 0; --k) local_phase(mode, k, k > 0 ? k - 1 : -1, notOk, prev);
    }
  };

  if (W.nJoints > 0) sweep(TM_INIT, nullptr, nullptr);                        // joints: InitVelocityConstraints + warm start (:143-146)
  for (int it = 0; it < W.velIters; ++it) sweep(TM_VEL, nullptr, nullptr);    // :153-161
  // StoreImpulses (:164)
  {
    const int ns = min(H->nSolve, W.sCap);
    for (int s = gt; s < ns; s += gn) {
      const int i = W.s_contact[s];
      const int vcCount = W.s_pc[s] & 0xFF;
      const float4 imp = W.s_imp[s];
      float4 old = W.c_imp[i];
      old.x = imp.x; old.y = imp.y;
      if (vcCount == 2) { old.z = imp.z; old.w = imp.w; }
      W.c_imp[i] = old;
    }
  }
  // integrate positions (:168-200): tile bodies in shared memory, the island's other bodies (kinematic) in the global arrays
  const float h = W.dt;
  auto integrate = [&](float4& pos, float4& vel) {
    v2 c = V(pos.x, pos.y), v = V(vel.x, vel.y);
    float a = pos.z, w = vel.z;
    v2 translation = h * v;
    if (dot(translation, translation) > kMaxTranslationSquared) { float ratio = kMaxTranslation / len(translation); v *= ratio; }
    float rotation = h * w;
    if (rotation * rotation > kMaxRotationSquared) { float ratio = kMaxRotation / fabsr(rotation); w *= ratio; }
    c += h * v;
    a += h * w;
    pos = make_float4(c.x, c.y, a, 0.0f); vel = make_float4(v.x, v.y, w, 0.0f);
  };
  for (int i = lt; i < n; i += ln) {
    const uint32_t f = W.b_flags[sBody[i]];
    if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
    float4 pos = sPos[i], vel = sVel[i];
    integrate(pos, vel);
    sPos[i] = pos; sVel[i] = vel;
  }
  for (int b = gt; b < W.nBodies; b += gn) {
    if (W.b_tslot[b] >= 0) continue;
    const uint32_t f = W.b_flags[b];
    if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
    float4 pos = ldcg4(&W.b_pos[b]), vel = ldcg4(&W.b_vel[b]);
    integrate(pos, vel);
    stcg4(&W.b_pos[b], pos); stcg4(&W.b_vel[b], vel);
  }
  GB();
  MARK();
  // position iterations (:206-224) with the per-island early-out flags of k_solve
  for (int it = 0; it < W.posIters; ++it) {
    int* notOk = W.b_posNotOk + it * W.nBodies;
    const int* prev = it > 0 ? W.b_posNotOk + (it - 1) * W.nBodies : nullptr;
    sweep(TM_POS, notOk, prev);     // (a pass starts with publish + grid barrier when islands can span tiles: the flags of the pass before are in)
    MARK();
  }
  // write back + SynchronizeTransform (:227-235), sleep bookkeeping (:241-269); the claim / exchange scratch goes back to rest
  {
    const float linTolSqr = kLinearSleepTolerance * kLinearSleepTolerance;
    const float angTolSqr = kAngularSleepTolerance * kAngularSleepTolerance;
    auto finish = [&](int b, uint32_t f, float4 pos, float4 vel) {
      const float4 lc = W.b_lc[b];
      W.b_xf[b] = pack(xf_from_sweep(V(pos.x, pos.y), pos.z, V(lc.x, lc.y)));
      if (W.allowSleep) {
        float2 gs = W.b_gs[b];
        if (!(f & BF_AUTOSLEEP) || vel.z * vel.z > angTolSqr || dot(V(vel.x, vel.y), V(vel.x, vel.y)) > linTolSqr) gs.y = 0.0f;
        else gs.y += h;
        W.b_gs[b] = gs;
        atomicMin(&W.b_islMinSleep[W.b_root[b]], __float_as_int(gs.y));
      }
    };
    for (int i = lt; i < n; i += ln) {
      const int b = sBody[i];
      if (sFlag[i]) W.b_xflag[b] = 0;
      W.b_tclaim[b] = 0x7fffffff;
      const uint32_t f = W.b_flags[b];
      if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
      const float4 pos = sPos[i], vel = sVel[i];
      stcg4(&W.b_pos[b], pos); stcg4(&W.b_vel[b], vel);
      finish(b, f, pos, vel);
    }
    for (int b = gt; b < W.nBodies; b += gn) {
      if (W.b_tslot[b] >= 0) continue;
      const uint32_t f = W.b_flags[b];
      if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
      finish(b, f, ldcg4(&W.b_pos[b]), ldcg4(&W.b_vel[b]));
    }
  }
  if (W.allowSleep) {
    GB();
    const int* last = W.posIters > 0 ? W.b_posNotOk + (W.posIters - 1) * W.nBodies : nullptr;
    for (int b = gt; b < W.nBodies; b += gn) {
      const uint32_t f = W.b_flags[b];
      if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
      const int root = W.b_root[b];
      const bool positionSolved = last && __ldcg(&last[root]) == 0;
      const float minSleep = __int_as_float(__ldcg(&W.b_islMinSleep[root]));
      if (minSleep >= kTimeToSleep && positionSolved) {
        W.b_flags[b] = f & ~BF_AWAKE;        // b2Body.SetAwake(false) (b2body.d:837-845)
        W.b_gs[b].y = 0.0f;
        W.b_vel[b] = make_float4(0, 0, 0, 0);
        W.b_force[b] = make_float4(0, 0, 0, 0);
      }
    }
  }
  MARK();
#undef GB
#undef MARK
}

// ------------------------------------------------------------------------------------------------ host side
size_t tile_smem_bytes(int tileBodies) { return (size_t)tileBodies * 40 + (size_t)kTileBMax * 8; }   // bodies; the re-colouring scratch needs 8 kTileBMax + 8 T

cudaError_t stage_tile_assign(const DevWorld& W, const LaunchCfg& L, unsigned* keysA, unsigned* keysB, int* valsA, int* valsB) {
  ++L.launches; k_tile_body_keys<<<L.gridWide, 256, 0, L.stream>>>(W, keysA, valsA);
  cub::DoubleBuffer<unsigned> keys(keysA, keysB);
  cub::DoubleBuffer<int> vals(valsA, valsB);
  size_t bytes = L.cubTempBytes;
  CK(cub::DeviceRadixSort::SortPairs(L.cubTemp, bytes, keys, vals, W.nBodies, 0, 32, L.stream));
  ++L.launches; k_tile_slots<<<L.gridWide, 256, 0, L.stream>>>(W, keys.Current(), vals.Current());
  return cudaGetLastError();
}
// colouring as in stage_colour_and_sort, then the constraints sorted by (class, tile, colour) instead of by colour
cudaError_t stage_colour_and_sort_tiles(const DevWorld& W, const LaunchCfg& L) {
  CK(launch_mark_and_colour(W, L));
  ++L.launches; k_tile_claim<<<L.gridWide, 256, 0, L.stream>>>(W);
  ++L.launches; k_tile_key<<<L.gridWide, 256, 0, L.stream>>>(W);
  ++L.launches; k_tile_scan<<<1, 1024, 0, L.stream>>>(W);
  ++L.launches; k_tile_scatter<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}
cudaError_t stage_solve_tiles(const DevWorld& W, const LaunchCfg& L) {
  static bool attr = false;
  if (!attr) { CK(cudaFuncSetAttribute((const void*)k_solve_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
  if ((L.coopLaunches++ & 1023) == 0) CK(cudaMemsetAsync(&W.hdr->barrier, 0, sizeof(unsigned), L.stream));   // see launch_coop
  void* args[] = {(void*)&W};
  ++L.launches;
  return cudaLaunchCooperativeKernel((const void*)k_solve_tiles, dim3(L.coopBlocks), dim3(L.coopThreads), args, tile_smem_bytes(W.tileBodies), L.stream);
}

}  // namespace dbx
