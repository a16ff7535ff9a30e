// dbx_tiles.cu — the tile solver: b2Island.Solve (dynamics/b2island.d:118-279) for one big world with the bodies of a spatial
// tile resident in ONE CTA's shared memory.
//
// k_solve (dbx_solve.cu) runs every colour of every Gauss-Seidel pass as a grid-wide phase: ~104 dependent phases of ~4 us on
// the 100,000-body pile, each a grid barrier plus an L2 round trip for two bodies.  Here the dynamic bodies are sorted along x
// and cut into P tiles of T bodies (P <= one CTA per SM).  A constraint whose dynamic bodies sit in one tile is LOCAL (class L):
// its CTA walks the local colours with __syncthreads() between them, bodies and rows in shared memory.  A constraint between
// tiles s and s + 1 whose bodies sit in the right half of s and the left half of s + 1 is a BOUNDARY constraint (class B),
// solved by CTA s after its local ones; the halves make that claim exclusive without arbitration.  Everything else (a reach
// over more than one boundary or from the wrong half, a gear joint, an overflow colour of a hub body) is GLOBAL (class G) and
// runs as grid-wide colour phases like k_solve's.
//
// What lives in a tile's shared memory for the whole solve: its bodies (velocity and position slots whose fourth components
// carry the inverse mass and inertia; id + exchange flags), its local rows (six float4 arrays, staged with TMA bulk copies),
// its local and boundary joints (revolute / distance), its boundary rows, and slots for the bodies of the right-hand neighbour
// those boundary constraints touch (loaded at the start of a boundary pass, written back at its end).  Whatever does not fit
// stays in the global arrays and is reached through L2 (every such branch is forced by a DBX_DEBUG bit in the tests).
//
// One pass without global constraints = L colours (block barriers) -> publish the exchange bodies -> wait for the right-hand
// neighbour's publication (a flag per tile) -> B colours, re-coloured per tile (block barriers) -> signal -> wait for the left-hand
// neighbour's boundary pass -> read the exchange bodies back: no grid barrier.  With global constraints the two waits become grid
// barriers and the G colours follow, a grid barrier each.  Position passes run the same schedule backwards, so that a body
// still meets its contacts before its joints (b2island.d:206-216), as in k_solve's unified phases.
//
// The order is a Gauss-Seidel sweep like any other: within a phase no two constraints share a dynamic body (the global
// colouring is proper, the boundary re-colouring is proper among a tile's boundary constraints, and L / B / G phases never
// overlap in time on a body).  The schedule is reported by dbx_world_debug_read_solve_order, and tests hand it to the
// sequential oracle.  Measurements and what bounds the kernel: DESIGN.md section 7.0.
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

// (classification of the constraints: dbx_tilekey.cuh, called from k_mark_solve / k_colour)
// one CTA: exclusive scan of both histograms -> offsets, 1024 bins per round (coalesced, warp shuffles); totals and class sizes
// into the header; cursors back to zero for the scatter; handshake flags of k_solve_tiles back to zero
// Both histograms (20,000 bins each on 148 tiles) go through shared memory: in with coalesced loads that are all under way at
// once, every thread scans its own chunk of consecutive bins there (chunks an odd number of words apart: no bank conflicts), one
// block-wide scan of the 1,024 chunk totals, out with coalesced stores.  (Round by round -- load 1,024 bins, scan, store, carry
// -- the same work took 29 us: twenty rounds of dependent L2 round trips and block barriers on a single SM.)
__global__ void __launch_bounds__(1024) k_tile_scan(const __grid_constant__ DevWorld W) {
  extern __shared__ int sbin[];
  __shared__ int wsum[2][32];
  __shared__ int chunkBase[2][1024];
  __shared__ int stats[4];
  const int P = W.nTiles, nBins = 2 * P * kTileColours + kMaxColours;
  const int per = (nBins + 1023) >> 10, stride = per | 1;
  int* A = sbin; int* B = sbin + 1024 * stride;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  if (t < 4) stats[t] = 0;
  __syncthreads();
  int nb = 0, ng = 0, maxc = 0;
  for (int base = t; base < nBins; base += 8 * 1024) {
    int cs[8], js[8];                                  // eight rounds of loads under way before the first is used
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int k = base + u * 1024; cs[u] = k < nBins ? W.t_cur[k] : 0; js[u] = k < nBins ? W.tj_cur[k] : 0; }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = base + u * 1024;
      if (k >= nBins) break;
      const int c = cs[u], j = js[u];
      const int ch = k / per, idx = ch * stride + (k - ch * per);
      A[idx] = c; B[idx] = j;
      if (c + j > 0) {
        const int col = k < 2 * P * kTileColours ? k % kTileColours : k - 2 * P * kTileColours;
        maxc = max(maxc, col + 1);
        if (k >= 2 * P * kTileColours) ng += c + j; else if (k >= P * kTileColours) nb += c + j;
      }
    }
  }
  __syncthreads();
  int sa = 0, sb = 0;
  for (int i = 0; i < per; ++i) {
    if (t * per + i >= nBins) break;
    const int a = A[t * stride + i], b = B[t * stride + i];
    A[t * stride + i] = sa; B[t * stride + i] = sb;
    sa += a; sb += b;
  }
  int x = sa, y = sb;
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, x, o), b = __shfl_up_sync(0xffffffffu, y, o);
    if (lane >= o) { x += a; y += b; }
  }
  if (lane == 31) { wsum[0][wid] = x; wsum[1][wid] = y; }
  __syncthreads();
  if (wid == 0) {
    const int v0 = wsum[0][lane], v1 = wsum[1][lane];
    int p = v0, q = v1;
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, p, o), b = __shfl_up_sync(0xffffffffu, q, o);
      if (lane >= o) { p += a; q += b; }
    }
    wsum[0][lane] = p - v0; wsum[1][lane] = q - v1;          // exclusive prefix of the warp totals
  }
  __syncthreads();
  chunkBase[0][t] = wsum[0][wid] + x - sa; chunkBase[1][t] = wsum[1][wid] + y - sb;
  if (nb) atomicAdd(&stats[0], nb);
  if (ng) atomicAdd(&stats[1], ng);
  if (maxc) atomicMax(&stats[2], maxc);
  for (int k = t; k < 2 * P; k += 1024) W.t_flag[k] = 0;
  __syncthreads();
  for (int k = t; k < nBins; k += 1024) {
    const int ch = k / per, idx = ch * stride + (k - ch * per);
    W.t_off[k] = chunkBase[0][ch] + A[idx]; W.tj_off[k] = chunkBase[1][ch] + B[idx];
    W.t_cur[k] = 0; W.tj_cur[k] = 0;
  }
  if (t == 1023) {
    const int total = chunkBase[0][t] + sa;
    W.t_off[nBins] = total; W.tj_off[nBins] = chunkBase[1][t] + sb;
    W.hdr->nSolve = total;
    if (total > W.sCap) W.hdr->error = E_SOLVER_ROWS;
    W.hdr->nTileB = stats[0]; W.hdr->nTileG = stats[1]; W.hdr->nColours = stats[2]; W.hdr->tailStart = stats[2];
  }
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
constexpr int kTileThreads = 384;    // k_solve_tiles: a colour of a tile holds ~300 rows at most; fewer threads leave each more registers (measured: -7 us)
constexpr int kTileNbrMax = 128;                 // bodies of the right-hand neighbour a CTA's boundary constraints may stage in shared memory
constexpr int kTileNoItem = (int)0x80000000;     // an empty position of a tile's boundary order (padding between a colour's joints and rows)
constexpr int kTileBMax = 2 * kTileThreads;      // boundary constraints one CTA re-colours locally (more: it walks them by global colour)
constexpr int kTileBColours = 32;
constexpr int kTileJointsMax = 512;  // local joints of a tile that may live in shared memory

// A constraint that works on rows in the GLOBAL arrays: item >= 0 is a contact row (solver slot), item < 0 a joint
// (~joint slot).  Boundary and global phases, all joints, and the local rows of a tile too big for shared memory come here;
// deliberately not inlined, so the kernel holds one copy of this code.
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
// pull a row the thread will need in its NEXT phase from L2 into L1 while it works on the current one (rows are read-only
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
// ---- TMA: contiguous pieces of the row arrays go to shared memory as bulk asynchronous copies that complete on an mbarrier
DBX_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
DBX_D void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DBX_D void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DBX_D void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
DBX_D void bulk_g2s(void* dstSmem, const void* srcGlobal, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kTileThreads) k_solve_tiles(const __grid_constant__ DevWorld W) {
  // dynamic shared memory: body slots [TX = T + kTileNbrMax] (own bodies, then staged neighbour bodies), row slots [R] (local rows,
  // then staged boundary rows), joints [nJS]
  //   float4 vel[TX] pos[TX] | a0..a5[R] (velocity: v0 r0 r1 q0 q1 imp; position: p0 p1 p2, p3 + island + flag) | int2 bd[R] | int body[T] | int pc[R] | joints 192 B each
  extern __shared__ float4 sm4[];
  __shared__ int offL[kTileColours + 1], joffL[kTileColours + 1], offB[kTileColours + 1], joffB[kTileColours + 1];
  __shared__ int joffG[kTileColours + 1];
  __shared__ int sNbrBody[kTileNbrMax], nNbr;       // the neighbour's bodies the boundary constraints reach: ids, count
  __shared__ int phL[kTileColours], nPhL;          // the local colours that hold anything, in order
  __shared__ int sBItem[kTileBMax], sBOff[kTileBColours + 1], sBCnt[kTileBColours], sBCntJ[kTileBColours], nPhB, bDirect, nBLen;
  __shared__ unsigned long long rowBar;             // mbarrier of the row staging copies
  __shared__ DevWorld sW;                           // W with the joint arrays pointing at the tile's shared-memory copies (local joints)
  __shared__ int jointsBad;
  Header* H = W.hdr;
  const unsigned nb = gridDim.x;
  const int lt = threadIdx.x, ln = blockDim.x;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
  const int P = W.nTiles, T = W.tileBodies;
  const int tile = blockIdx.x;
  const int s0 = tile * T, n = tile < P ? max(0, min(T, W.nTileBodies - s0)) : 0;
  unsigned dynBytes; asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dynBytes));
  const int nbrMax = (W.dbgFlags & 32768) ? 8 : kTileNbrMax;       // (debug: a tiny limit, so that tests see the overflow branch)
  const int TX = T + kTileNbrMax;                   // body slots: the tile's own [0, T), staged neighbour bodies [T, T + nNbr)
  float4* sVel = sm4; float4* sPos = sm4 + TX;
  const int* offG = W.t_off + 2 * P * kTileColours; // (global rows are the rare case: their offsets stay in L2)
  BodyView view; view.vel = sVel; view.pos = sPos; view.off = s0; view.mode = 2;
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
    if (lt == 0) { mbar_init(&rowBar, 1); nNbr = 0; }
  }
  const int nColours = H->nColours;
  const int nG = H->nTileG, nCross = H->nTileB + nG;      // uniform over the grid
  for (int k = gt; k < 2 * P * kTileColours + kMaxColours; k += gn) { W.t_cur[k] = 0; W.tj_cur[k] = 0; }   // the next step's histogram starts from zero
  const int nLoc = min(nColours, kTileColours);
#define GB() grid_barrier(&H->barrier, nb)
  // debug (dbx_world_debug_phase_times): %globaltimer stamps of CTA 0 at [0 ..) and of the middle CTA at [1024 ..)
  int markIdx = 0;
  const bool marking = W.phaseTimes != nullptr && lt == 0 && (blockIdx.x == 0 || blockIdx.x == nb / 2);
  const int markBase = blockIdx.x == 0 ? 0 : 1024;
#define MARK() do { if (marking && markIdx < 1000) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); W.phaseTimes[markBase + markIdx++] = t_; } } while (0)
  MARK();
  if (W.phaseTimes != nullptr && lt == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); W.phaseTimes[3300 + blockIdx.x] = t_; }   // debug: start skew of the CTAs
  __syncthreads();
  MARK();
  const int rs0 = offL[0], nRows = offL[kTileColours] - rs0;        // the tile's local rows: one contiguous piece of the row arrays
  // the tile's local joints (revolute / distance) move into shared memory too when they fit beside the rows: 192 B each
  const int jl0 = joffL[0], nJL = tile < P ? joffL[kTileColours] - jl0 : 0;
  const int jb0 = joffB[0], nBJ = tile < P ? joffB[kTileColours] - jb0 : 0, nBC = tile < P ? offB[kTileColours] - offB[0] : 0, nB = nBJ + nBC;
  const long long bodyBytes = 36ll * T + 32ll * kTileNbrMax;
  auto rows_beside = [&](int joints) { return (int)(((long long)dynBytes - bodyBytes - 192ll * joints - 32) / 108) & ~1; };
  const int R0 = (int)(((long long)dynBytes - bodyBytes) / 108) & ~1;
  // (the boundary's joints join the local ones in shared memory when there is room for them too)
  const int nJS = (nJL + nBJ <= kTileJointsMax && nRows <= rows_beside(nJL + nBJ)) ? nJL + nBJ : nJL;
  const int Rj = rows_beside(nJS);
  const bool jointRoom = nJS > 0 && nJS <= kTileJointsMax && Rj > 0 && nRows <= Rj && !(W.dbgFlags & 2048);
  const int R = jointRoom ? Rj : R0;
  float4* ra0 = sm4 + 2 * TX; float4* ra1 = ra0 + R; float4* ra2 = ra1 + R; float4* ra3 = ra2 + R; float4* ra4 = ra3 + R; float4* ra5 = ra4 + R;
  int2* rbd = (int2*)(ra5 + R);
  int* sBody = (int*)(rbd + R); int* rpc = sBody + T;      // sBody: body id | exchange flags << 27
  // (a body's inverse mass rides in the fourth component of its velocity slot, its inverse inertia in that of its position)
  auto body_id = [&](int i) { return sBody[i] & 0x07FFFFFF; };
  auto body_xf = [&](int i) { return (int)((unsigned)sBody[i] >> 27); };
  auto set3 = [](float4* p, float4 v) { p->x = v.x; p->y = v.y; p->z = v.z; };
  auto xyz0 = [](float4 v) { return make_float4(v.x, v.y, v.z, 0.0f); };
  float4* sj = (float4*)(((size_t)(rpc + R) + 15) & ~(size_t)15);  // joints: 11 float4 arrays [nJL], then bref int2, limit, root
  const bool rowsLocal = nRows > 0 && nRows <= R;                   // they fit: they live in shared memory for the whole solve
  if (lt == 0) {
    // the tile's local rows into shared memory: the six float4 arrays as TMA bulk copies, under way while the boundary rows
    // are re-coloured and the bodies come in (the two narrow arrays follow by hand)
    if (rowsLocal) {
      const unsigned bytes = (unsigned)nRows * 16u;
      mbar_expect_tx(&rowBar, 6u * bytes);
      bulk_g2s(ra0, W.s_v0 + rs0, bytes, &rowBar); bulk_g2s(ra1, W.s_r0 + rs0, bytes, &rowBar); bulk_g2s(ra2, W.s_r1 + rs0, bytes, &rowBar);
      bulk_g2s(ra3, W.s_q0 + rs0, bytes, &rowBar); bulk_g2s(ra4, W.s_q1 + rs0, bytes, &rowBar); bulk_g2s(ra5, W.s_imp + rs0, bytes, &rowBar);
    }
    int k = 0;
    for (int c = 0; c < nLoc; ++c) if (offL[c] != offL[c + 1] || joffL[c] != joffL[c + 1]) phL[k++] = c;
    nPhL = k;
  }
  // Boundary constraints: the global colouring spreads a boundary's ~150 rows over every colour in use (ten phases of a dozen
  // rows each); among themselves they need three or four.  Re-colour them here: Jones-Plassmann rounds in shared memory, an
  // item takes the lowest colour free on both bodies once it holds the smallest priority on both (priority = joints before
  // contacts, then a hash of the item -- so a body still meets its joints first and the outcome does not depend on scheduling).
  {
    unsigned* sMask = (unsigned*)rbd;                 // scratch in the tail of the dynamic area (references, ids: filled afterwards)
    unsigned* sClaim = sMask + 2 * T;                 // [2T] each: own tile's bodies [0, T), the right-hand neighbour's [T, 2T)
    const bool fits = nB <= kTileBMax && R0 > 0 && (size_t)16 * T <= (size_t)dynBytes - (size_t)(32 * TX + 96 * R);
    if (lt == 0) { bDirect = fits ? 0 : 1; nPhB = 0; nBLen = 0; }
    if (lt < kTileBColours) { sBCnt[lt] = 0; sBCntJ[lt] = 0; }
    if (fits && nB > 0) {
      for (int k = lt; k < 2 * T; k += ln) { sMask[k] = 0u; sClaim[k] = 0xFFFFFFFFu; }
      // every thread keeps at most two items in registers (nB <= kTileBMax = 2 x kTileThreads)
      int item[2], ia[2], ib[2], lc[2], jg[2]; unsigned pr[2]; int2 brs[2];
      for (int u = 0; u < 2; ++u) {
        const int k = lt + u * ln;
        item[u] = 0; ia[u] = ib[u] = -1; lc[u] = k < nB ? -1 : 0; pr[u] = 0xFFFFFFFFu; brs[u] = make_int2(-1, -1);
        if (k >= nB) continue;
        int2 br;
        jg[u] = -1;
        if (k < nBJ) { const int j = W.tj_order[joffB[0] + k]; item[u] = ~k; jg[u] = j; br = W.j_bref[j];     /* a joint item: ~index in the boundary's joint list */ pr[u] = ((unsigned)(mix64((unsigned long long)j + 1ull) >> 33)) & 0x7FFFFC00u; }
        else {
          // (the hash is of the pair key, not of the slot: slots inside a bin come out of an atomic scatter in any order)
          const int s = offB[0] + (k - nBJ), i = W.s_contact[s]; item[u] = s; br = W.c_bref[i];
          pr[u] = 0x80000000u | (((unsigned)(mix64(W.c_key[i]) >> 33)) & 0x7FFFFC00u);
        }
        pr[u] |= (unsigned)k;                        // unique (two pair keys with equal hash bits: practically never)
        // index of each body in the scratch: own tile [0, T), the neighbour's [T, 2T), -1 for a body nobody moves
        if (br.x >= 0) ia[u] = br.x - s0; else { const int sl = W.b_tslot[br.x & 0x7fffffff]; if (sl >= 0) ia[u] = sl - s0; }
        if (br.y >= 0) ib[u] = br.y - s0; else { const int sl = W.b_tslot[br.y & 0x7fffffff]; if (sl >= 0) ib[u] = sl - s0; }
        brs[u] = br;
      }
      __syncthreads();
      for (int round = 0; round < 4096; ++round) {
        int open = 0;
        for (int u = 0; u < 2; ++u) if (lc[u] < 0) { ++open; if (ia[u] >= 0) atomicMin(&sClaim[ia[u]], pr[u]); if (ib[u] >= 0) atomicMin(&sClaim[ib[u]], pr[u]); }
        if (__syncthreads_count(open) == 0) break;
        bool won[2] = {false, false};
        for (int u = 0; u < 2; ++u) if (lc[u] < 0 && (ia[u] < 0 || sClaim[ia[u]] == pr[u]) && (ib[u] < 0 || sClaim[ib[u]] == pr[u])) {
          const unsigned used = (ia[u] >= 0 ? sMask[ia[u]] : 0u) | (ib[u] >= 0 ? sMask[ib[u]] : 0u);
          won[u] = true;
          if (!~used) { bDirect = 1; lc[u] = 0; continue; }       // more than 32 boundary rows on one body: walk them by global colour
          lc[u] = __ffs((int)~used) - 1;
          if (ia[u] >= 0) sMask[ia[u]] |= 1u << lc[u];
          if (ib[u] >= 0) sMask[ib[u]] |= 1u << lc[u];
        }
        __syncthreads();
        for (int u = 0; u < 2; ++u) if (won[u]) { if (ia[u] >= 0) sClaim[ia[u]] = 0xFFFFFFFFu; if (ib[u] >= 0) sClaim[ib[u]] = 0xFFFFFFFFu; }
        __syncthreads();
      }
      // counting sort of the items by local colour (the order inside a colour is free).  Inside a colour the joints come first
      // and the rows start on a warp of their own: a warp that held both would run the two code paths one after the other.
      for (int u = 0; u < 2; ++u) if (lt + u * ln < nB) atomicAdd(item[u] < 0 ? &sBCntJ[lc[u]] : &sBCnt[lc[u]], 1);
      __syncthreads();
      if (lt == 0) {
        int padded = 0;
        for (int c = 0; c < kTileBColours; ++c) padded += ((sBCntJ[c] + 31) & ~31) + sBCnt[c];
        const bool pad = padded <= kTileBMax;
        int acc = 0, np = 0;
        for (int c = 0; c < kTileBColours; ++c) {
          const int nj = sBCntJ[c], nc = sBCnt[c], rowsAt = acc + (pad ? (nj + 31) & ~31 : nj);
          sBOff[c] = acc; sBCntJ[c] = acc; sBCnt[c] = rowsAt;
          for (int q = acc + nj; q < rowsAt; ++q) sBItem[q] = kTileNoItem;
          acc = rowsAt + nc;
          if (nj + nc) np = c + 1;
        }
        sBOff[kTileBColours] = acc;
        nPhB = np; nBLen = acc;
      }
      __syncthreads();
      if (!bDirect) for (int u = 0; u < 2; ++u) if (lt + u * ln < nB) {
        sBItem[atomicAdd(item[u] < 0 ? &sBCntJ[lc[u]] : &sBCnt[lc[u]], 1)] = item[u];
        if (item[u] >= 0) W.c_tcol[W.s_contact[item[u]]] = lc[u]; else W.j_tcol[jg[u]] = lc[u];    // for dbx_world_debug_read_solve_order
      }
      // The neighbour's bodies these constraints reach (a hundred or so) get slots of their own behind the tile's bodies: a
      // boundary pass loads them once, works in shared memory, and writes them back once -- instead of a trip to L2 and back
      // inside every row.  sClaim is all-ones again by now; its neighbour half serves as the map body -> slot.
      if (!bDirect && !(W.dbgFlags & 4096)) {
        for (int u = 0; u < 2; ++u) {
          const int x[2] = {ia[u], ib[u]}, id[2] = {brs[u].x, brs[u].y};
          for (int e = 0; e < 2; ++e) if (x[e] >= T && atomicCAS(&sClaim[x[e]], 0xFFFFFFFFu, 0xFFFFFFFEu) == 0xFFFFFFFFu) {
            const int k = atomicAdd(&nNbr, 1);
            if (k < nbrMax) { sNbrBody[k] = id[e] & 0x7fffffff; sClaim[x[e]] = (unsigned)k; }
          }
        }
        __syncthreads();
        if (nNbr <= nbrMax) {
          for (int u = 0; u < 2; ++u) if (lt + u * ln < nB) {
            int2 br = brs[u];
            if (ia[u] >= T) br.x = s0 + T + (int)sClaim[ia[u]];
            if (ib[u] >= T) br.y = s0 + T + (int)sClaim[ib[u]];
            if (item[u] >= 0) W.s_body[item[u]] = br; else W.j_bref[jg[u]] = br;      // (both rewritten from scratch every step)
          }
        }
        __syncthreads();
        if (lt == 0 && nNbr > nbrMax) nNbr = 0;
      }
    }
    __syncthreads();
  }
  const bool bLocal = bDirect == 0;
  const int nNb = nNbr, nBL = nBLen;                // nBL: length of the boundary's phase order (items + padding)
  MARK();
  long long pc0 = 0; if (marking) pc0 = clock64();
#define PSTAMP(i) do { if (marking) { W.phaseTimes[3900 + (blockIdx.x == 0 ? 0 : 16) + (i)] = (unsigned long long)(clock64() - pc0); } } while (0)

  if (rowsLocal) for (int k = lt; k < nRows; k += ln) { rbd[k] = W.s_body[rs0 + k]; rpc[k] = W.s_pc[rs0 + k]; }
  // Boundary rows too, as many as the spare row slots take, from the END of the boundary's phase order (its later colours are
  // the small ones: a colour that sits in shared memory whole runs at the local rows' pace).  Position q of the order gets
  // slot nRows + q - qS0; a joint's position stays empty.
  const int nBS = (bLocal && rowsLocal && nNb > 0 && !(W.dbgFlags & 8192)) ? min(nBL, R - nRows) : 0, qS0 = nBL - nBS;
  for (int q = qS0 + lt; q < nBL; q += ln) {
    const int s = sBItem[q], x = nRows + q - qS0;
    if (s < 0) continue;
    ra0[x] = W.s_v0[s]; ra1[x] = W.s_r0[s]; ra2[x] = W.s_r1[s]; ra3[x] = W.s_q0[s]; ra4[x] = W.s_q1[s]; ra5[x] = W.s_imp[s];
    rbd[x] = W.s_body[s]; rpc[x] = W.s_pc[s];
  }
  PSTAMP(0);
  // local joints in: definition, accumulated impulses, limit state, body references; the per-step temporaries are written by
  // their own joint_init.  sW is W with the joint arrays redirected, indexed by the joint's position in the tile's list.
  float4* sjIds = sj; float4* sjAnchor = sj + nJS; float4* sjP0 = sj + 2 * nJS; float4* sjP1 = sj + 3 * nJS; float4* sjImp = sj + 4 * nJS;
  int2* sjBref = (int2*)(sj + 11 * nJS); int* sjLimit = (int*)(sjBref + nJS); int* sjRoot = sjLimit + nJS;
  const int nJSe = bLocal ? nJS : nJL;              // (boundary constraints walked by global colour stay in the global arrays)
  auto staged_joint = [&](int x) { return x < nJL ? W.tj_order[jl0 + x] : W.tj_order[jb0 + (x - nJL)]; };
  if (lt == 0) jointsBad = 0;
  __syncthreads();
  if (jointRoom) {
    static_assert(sizeof(DevWorld) % 4 == 0, "DevWorld is copied word by word");
    for (int k = lt; k < (int)(sizeof(DevWorld) / 4); k += ln) ((int*)&sW)[k] = ((const int*)&W)[k];
    for (int x = lt; x < nJSe; x += ln) {
      const int j = staged_joint(x);
      const int4 ids = W.j_ids[j];
      if (ids.x != JT_REVOLUTE && ids.x != JT_DISTANCE) jointsBad = 1;        // (other joint types read more arrays: they stay in L2)
      ((int4*)sjIds)[x] = ids; sjAnchor[x] = W.j_anchor[j]; sjP0[x] = W.j_p0[j]; sjP1[x] = W.j_p1[j]; sjImp[x] = W.j_imp[j];
      sjBref[x] = W.j_bref[j]; sjLimit[x] = W.j_limit[j]; sjRoot[x] = -1;
    }
  }
  __syncthreads();
  PSTAMP(1);
  const bool jointsLocal = jointRoom && jointsBad == 0;
  if (jointsLocal && lt == 0) {
    sW.j_ids = (int4*)sjIds; sW.j_anchor = sjAnchor; sW.j_p0 = sjP0; sW.j_p1 = sjP1; sW.j_imp = sjImp;
    sW.j_r = sj + 5 * nJS; sW.j_lc = sj + 6 * nJS; sW.j_m = sj + 7 * nJS; sW.j_k0 = sj + 8 * nJS; sW.j_k1 = sj + 9 * nJS; sW.j_k2 = sj + 10 * nJS;
    sW.j_bref = sjBref; sW.j_limit = sjLimit; sW.j_root = sjRoot;
  }
  __syncthreads();
  // bodies in: velocities with the contacts' warm start folded in (see k_solve), positions, inverse masses, ids, exchange flags
  PSTAMP(2);
  {
    const float k = 1.0f / 4294967296.0f;
    // (two bodies per trip, every load of both under way before the first is used; what is kept in tile order -- ids, masses,
    // exchange flags, warm-start accumulators -- comes in coalesced, only velocity and position are gathered by body id: the
    // L1 works through a gathered load one 128-byte line at a time, and that queue is what this block waits for)
    for (int i = lt; i < n; i += 2 * ln) {
      const int i1 = i + ln;
      const bool two = i1 < n;
      const int e0 = s0 + i, e1 = two ? s0 + i1 : e0;
      const int b0 = W.t_body[e0], b1 = W.t_body[e1];
      float4 vel[2] = {ldcg4(&W.b_vel[b0]), ldcg4(&W.b_vel[b1])};
      const float4 pos[2] = {ldcg4(&W.b_pos[b0]), ldcg4(&W.b_pos[b1])};
      long long acc[2][3] = {{0, 0, 0}, {0, 0, 0}};
      if (W.warmStarting) {
        for (int c = 0; c < 3; ++c) { acc[0][c] = (long long)__ldcg(&W.b_acc[3 * e0 + c]); acc[1][c] = (long long)__ldcg(&W.b_acc[3 * e1 + c]); }
      }
      const float2 ms[2] = {W.t_mass[e0], W.t_mass[e1]};
      const int fl[2] = {W.b_xflag[e0], W.b_xflag[e1]};
      for (int u = 0; u < (two ? 2 : 1); ++u) {
        const int e = u ? e1 : e0, iu = u ? i1 : i;
        if ((acc[u][0] | acc[u][1] | acc[u][2]) != 0) {
          vel[u].x += (float)acc[u][0] * k; vel[u].y += (float)acc[u][1] * k; vel[u].z += (float)acc[u][2] * k;
          __stcg(&W.b_acc[3 * e], 0ull); __stcg(&W.b_acc[3 * e + 1], 0ull); __stcg(&W.b_acc[3 * e + 2], 0ull);
        }
        sVel[iu] = make_float4(vel[u].x, vel[u].y, vel[u].z, ms[u].x); sPos[iu] = make_float4(pos[u].x, pos[u].y, pos[u].z, ms[u].y);
        sBody[iu] = (u ? b1 : b0) | (fl[u] << 27);
      }
    }
    PSTAMP(3);
    for (int k = lt; k < nNb; k += ln) {
      const int b = sNbrBody[k];
      const float4 ms = W.b_mass[b];
      const float4 pos = ldcg4(&W.b_pos[b]);          // (boundary joints read positions when they initialise)
      sPos[T + k] = make_float4(pos.x, pos.y, pos.z, ms.y);
      sVel[T + k] = make_float4(0.0f, 0.0f, 0.0f, ms.x);
    }
    __syncthreads();
    PSTAMP(4);
  }
  MARK();
  if (rowsLocal) mbar_wait(&rowBar, 0);
  MARK();

  // ---- the phases of one class
  int sweepNo = 0;
  // one row that lives in shared memory (slot x; sg: its slot in the global row arrays)
  auto smem_row = [&](int mode, int x, int sg, int* notOk, const int* prev) {
    const int2 bd = rbd[x];
    const float2 mA = bd.x >= 0 ? make_float2(sVel[bd.x - s0].w, sPos[bd.x - s0].w) : make_float2(0.0f, 0.0f);
    const float2 mB = bd.y >= 0 ? make_float2(sVel[bd.y - s0].w, sPos[bd.y - s0].w) : make_float2(0.0f, 0.0f);
    if (mode == TM_VEL) {
      VC v; v.bd = bd; v.pc = rpc[x]; v.v0 = ra0[x]; v.v1 = make_float4(mA.x, mA.y, mB.x, mB.y); v.r0 = ra1[x]; v.r1 = ra2[x]; v.q0 = ra3[x]; v.q1 = ra4[x]; v.imp = ra5[x];
      if ((v.pc & 0xFF) == 2) { v.nm = W.s_nm[sg]; v.K = W.s_k[sg]; }      // (the block solver's matrices stay in L2: no room; fetching them a phase ahead into registers was measured slower)
      ra5[x] = contact_velocity_row(W, v, view);
    } else {
      const float4 p3 = ra3[x];
      const int root = __float_as_int(p3.z);                // (staged with the position rows)
      if (prev && __float_as_int(p3.w) == 0) return;
      PCn p; p.bd = bd; p.pc = rpc[x]; p.v1 = make_float4(mA.x, mA.y, mB.x, mB.y); p.p0 = ra0[x]; p.p1 = ra1[x]; p.p2 = ra2[x]; p.p3 = make_float2(p3.x, p3.y);
      const float minSep = contact_position_row(W, p, -1, -1, view);
      if (!(minSep >= -3.0f * kLinearSlop)) __stcg(&notOk[root], 1);
    }
  };
  // "did my island still move in the pass before", for every row in shared memory: one round of loads at the start of a position
  // pass (after the barrier that completes the flags) instead of a trip to L2 inside every row
  auto fetch_prev_flags = [&](const int* prev) {
    if (!rowsLocal || !prev) return;
    for (int k = lt; k < nRows + nBS; k += ln) ra3[k].w = __int_as_float(__ldcg(&prev[__float_as_int(ra3[k].z)]));
    __syncthreads();
  };
  // local colours of this tile, upwards or downwards: joints first, then rows -- from shared memory when they fit
  auto local_phases = [&](int mode, bool backwards, int* notOk, const int* prev) {
    for (int kk = 0; kk < nPhL; ++kk) {
      const int k = backwards ? nPhL - 1 - kk : kk;
      const int c = phL[k];
      const int jb = joffL[c], nj = joffL[c + 1] - jb, beg = offL[c], nr = offL[c + 1] - beg;
      if (mode == TM_INIT && nj == 0) continue;
      const int njPad = (nj + 31) & ~31;
      const bool fine = marking && (sweepNo == 4 || sweepNo == 11) && kk < 16;       // debug: clock stamps of one velocity pass at [2048 + 4 kk ..), of one position pass at [2304 + 4 kk ..)
      long long c0 = 0, c1 = 0;
      if (fine) c0 = clock64();
      if (jointsLocal) { for (int q = lt; q < nj; q += ln) tile_item(sW, view, mode, ~(jb - jl0 + q), notOk, prev); }
      else for (int q = lt; q < nj; q += ln) tile_item(W, view, mode, ~W.tj_order[jb + q], notOk, prev);
      if (mode != TM_INIT) {
        if (rowsLocal) {
          // (the rows start on the first warp after the joints': a warp holding both would run the two code paths in turn)
          for (int q = lt - njPad; q < nr; q += ln) if (q >= 0) smem_row(mode, beg - rs0 + q, beg + q, notOk, prev);
        } else {
          const int kn = backwards ? k - 1 : k + 1;
          if (kn >= 0 && kn < nPhL) {
            const int cn = phL[kn];
            const int sn = offL[cn] + lt - (joffL[cn + 1] - joffL[cn]);
            if (sn >= offL[cn] && sn < offL[cn + 1]) tile_prefetch_row(W, mode, sn);
          }
          for (int q = lt - nj; q < nr; q += ln) if (q >= 0) tile_item(W, view, mode, beg + q, notOk, prev);
        }
      }
      if (fine) c1 = clock64();
      __syncthreads();
      if (fine) { const long long c2 = clock64(); unsigned long long* o = W.phaseTimes + 2048 + (sweepNo == 11 ? 256 : 0) + (blockIdx.x == 0 ? 0 : 128) + 4 * kk; o[0] = (unsigned long long)(c1 - c0); o[1] = (unsigned long long)(c2 - c1); o[2] = (unsigned long long)(nj + nr); o[3] = (unsigned long long)c; }
    }
  };
  const bool bJointsLocal = jointsLocal && nJSe > nJL;
  auto boundary_phases = [&](int mode, bool backwards, int* notOk, const int* prev) {
    if (bLocal) {
      if (nNb > 0) {
        for (int k = lt; k < nNb; k += ln) { if (mode == TM_POS) set3(&sPos[T + k], ldcg4(&W.b_pos[sNbrBody[k]])); else set3(&sVel[T + k], ldcg4(&W.b_vel[sNbrBody[k]])); }
        __syncthreads();
      }
      for (int k = 0; k < nPhB; ++k) {
        const int c = backwards ? nPhB - 1 - k : k;
        const bool fine = marking && (sweepNo == 4 || sweepNo == 11) && k < 8;        // debug: [3700 ..)
        long long c0 = 0, c1 = 0; int nj = 0;
        if (fine) { c0 = clock64(); for (int q = sBOff[c]; q < sBOff[c + 1]; ++q) nj += sBItem[q] < 0; }
        for (int q = sBOff[c] + lt; q < sBOff[c + 1]; q += ln) {
          const int item = sBItem[q];
          if (item == kTileNoItem) continue;
          if (item < 0) {                                     // a joint: ~index in the boundary's joint list
            if (bJointsLocal) tile_item(sW, view, mode, ~(nJL + ~item), notOk, prev);
            else tile_item(W, view, mode, ~W.tj_order[jb0 + ~item], notOk, prev);
          } else if (mode != TM_INIT) {
            if (q >= qS0) smem_row(mode, nRows + q - qS0, item, notOk, prev); else tile_item(W, view, mode, item, notOk, prev);
          }
        }
        if (fine) c1 = clock64();
        __syncthreads();
        if (fine) { const long long c2 = clock64(); unsigned long long* o = W.phaseTimes + 3700 + (sweepNo == 11 ? 64 : 0) + (blockIdx.x == 0 ? 0 : 32) + 4 * k; o[0] = (unsigned long long)(c1 - c0); o[1] = (unsigned long long)(c2 - c1); o[2] = (unsigned long long)(sBOff[c + 1] - sBOff[c]); o[3] = (unsigned long long)nj; }
      }
      for (int k = lt; k < nNb; k += ln) { if (mode == TM_POS) stcg4(&W.b_pos[sNbrBody[k]], xyz0(sPos[T + k])); else stcg4(&W.b_vel[sNbrBody[k]], xyz0(sVel[T + k])); }
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
      const int f = body_xf(i);
      if (!(f & need) || (f & both) != both) continue;
      if (positions) stcg4(&W.b_pos[body_id(i)], xyz0(sPos[i])); else stcg4(&W.b_vel[body_id(i)], xyz0(sVel[i]));
    }
  };
  // Neighbour handshakes instead of grid barriers (when there are no global rows): tile s only ever exchanges bodies with
  // tiles s - 1 and s + 1, so "my right neighbour has published pass k" and "my left neighbour's boundary rows of pass k are
  // done" are all it has to wait for -- a slow tile holds up its neighbours, not the machine.  Flags count passes upwards.
  const bool handshake = nG == 0 && !(W.dbgFlags & 128);
  int* flagL = W.t_flag; int* flagB = W.t_flag + P;
  auto signal = [&](int* flag) {
    __syncthreads();
    if (lt == 0 && tile < P) { __threadfence(); atomicExch(flag + tile, sweepNo); }
  };
  auto await = [&](int* flag, int other) {
    if (lt == 0 && tile < P && other >= 0 && other < P) {
      while (*((volatile int*)(flag + other)) < sweepNo) { }
      __threadfence();
    }
    __syncthreads();
  };
  // one Gauss-Seidel pass.  forward: L, B, G with the colours upwards; backward (position passes): G, B, L downwards
  auto sweep = [&](int mode, int* notOk, const int* prev) {
    ++sweepNo;
    if (mode != TM_POS) {
      local_phases(mode, false, notOk, prev);
      MARK();
      if (nCross == 0) return;
      publish(false, XF_FOREIGN | XF_G, 0);
      if (handshake) {
        signal(flagL); await(flagL, tile + 1);
        MARK();
        boundary_phases(mode, false, notOk, prev);
        MARK();
        signal(flagB); await(flagB, tile - 1);
        for (int i = lt; i < n; i += ln) if (body_xf(i) & XF_FOREIGN) set3(&sVel[i], ldcg4(&W.b_vel[body_id(i)]));
        __syncthreads();
        MARK();
        return;
      }
      GB();
      MARK();
      boundary_phases(mode, false, notOk, prev);
      MARK();
      if (nG > 0) {
        publish(false, XF_G, XF_OWNB | XF_G);
        GB();
        global_phases(mode, false, notOk, prev);
      } else GB();
      for (int i = lt; i < n; i += ln) if (body_xf(i) & (XF_FOREIGN | XF_G)) set3(&sVel[i], ldcg4(&W.b_vel[body_id(i)]));
      __syncthreads();
      MARK();
    } else {
      if (nCross > 0 && handshake) {
        if (prev && W.phaseTimes != nullptr && lt == 0 && sweepNo == W.velIters + (W.nJoints > 0 ? 1 : 0) + 2) {   // debug: arrival at the first grid barrier since the start
          unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
          W.phaseTimes[3500 + blockIdx.x] = t_;
          { unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); W.phaseTimes[3650 + blockIdx.x] = sm_ + 1; }
          W.phaseTimes[3940 + blockIdx.x] = (unsigned long long)nPhL * 100000000ull + (unsigned long long)nRows * 10000ull + (unsigned long long)(nPhB * 1000 + nB)
                                        + ((unsigned long long)nNb << 32) + ((unsigned long long)nJS << 42) + ((unsigned long long)(jointsLocal ? 1 : 0) << 52) + ((unsigned long long)nBS << 53);
        }
        if (prev) GB();            // the per-island flags of the pass before must be in for every tile alike
        fetch_prev_flags(prev);
        publish(true, XF_FOREIGN, 0);
        signal(flagL); await(flagL, tile + 1);
        boundary_phases(mode, true, notOk, prev);
        signal(flagB); await(flagB, tile - 1);
        for (int i = lt; i < n; i += ln) if (body_xf(i) & XF_FOREIGN) set3(&sPos[i], ldcg4(&W.b_pos[body_id(i)]));
        __syncthreads();
      } else if (nCross > 0) {
        publish(true, XF_FOREIGN | XF_G, 0);
        GB();
        fetch_prev_flags(prev);
        if (nG > 0) {
          global_phases(mode, true, notOk, prev);
          for (int i = lt; i < n; i += ln) { const int f = body_xf(i); if ((f & (XF_OWNB | XF_G)) == (XF_OWNB | XF_G)) set3(&sPos[i], ldcg4(&W.b_pos[body_id(i)])); }
          __syncthreads();
        }
        boundary_phases(mode, true, notOk, prev);
        GB();
        // the neighbour's boundary pass and the global phases wrote the global copy; a body this tile's own boundary rows
        // moved after the global phases is newest in shared memory
        for (int i = lt; i < n; i += ln) {
          const int f = body_xf(i);
          if ((f & XF_FOREIGN) || ((f & XF_G) && !(f & XF_OWNB))) set3(&sPos[i], ldcg4(&W.b_pos[body_id(i)]));
        }
        __syncthreads();
      } else fetch_prev_flags(prev);
      local_phases(mode, true, notOk, prev);
    }
  };

  solve_stamp(W, 0);
  if (W.nJoints > 0) sweep(TM_INIT, nullptr, nullptr);                        // joints: InitVelocityConstraints + warm start (:143-146)
  solve_stamp(W, 1);
  for (int it = 0; it < W.velIters; ++it) sweep(TM_VEL, nullptr, nullptr);    // :153-161
  // StoreImpulses (:164).  A tile's local rows hold their working impulses in shared memory: they go back to the row array
  // (PostSolve records read them there) and into the persistent manifolds from here; the position rows take their place.
  auto store_impulse = [&](int s, float4 imp, int pc) {
    const int i = W.s_contact[s];
    float4 old = W.c_imp[i];
    old.x = imp.x; old.y = imp.y;
    if ((pc & 0xFF) == 2) { old.z = imp.z; old.w = imp.w; }
    W.c_imp[i] = old;
  };
  if (rowsLocal) {
    for (int k = lt; k < nRows; k += ln) { const float4 imp = ra5[k]; W.s_imp[rs0 + k] = imp; store_impulse(rs0 + k, imp, rpc[k]); }
    for (int q = qS0 + lt; q < nBL; q += ln) { const int s = sBItem[q]; if (s >= 0) stcg4(&W.s_imp[s], ra5[nRows + q - qS0]); }     // (into the manifolds with the other boundary rows below)
    __syncthreads();
    if (W.posIters > 0) {
      if (lt == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const unsigned bytes = (unsigned)nRows * 16u;
        mbar_expect_tx(&rowBar, 3u * bytes);
        bulk_g2s(ra0, W.s_p0 + rs0, bytes, &rowBar); bulk_g2s(ra1, W.s_p1 + rs0, bytes, &rowBar); bulk_g2s(ra2, W.s_p2 + rs0, bytes, &rowBar);
      }
      for (int k = lt; k < nRows; k += ln) { const float2 r = W.s_p3[rs0 + k]; ra3[k] = make_float4(r.x, r.y, __int_as_float(W.s_root[rs0 + k]), 0.0f); }   // .z: the row's island, .w: its flag of the pass before
      for (int q = qS0 + lt; q < nBL; q += ln) {
        const int s = sBItem[q], x = nRows + q - qS0;
        if (s < 0) { ra3[x] = make_float4(0.0f, 0.0f, __int_as_float(0), 0.0f); continue; }
        const float2 r = W.s_p3[s];
        ra0[x] = W.s_p0[s]; ra1[x] = W.s_p1[s]; ra2[x] = W.s_p2[s]; ra3[x] = make_float4(r.x, r.y, __int_as_float(W.s_root[s]), 0.0f);
      }
    }
  } else if (tile < P) {
    for (int s = offL[0] + lt; s < offL[kTileColours]; s += ln) store_impulse(s, W.s_imp[s], W.s_pc[s]);
  }
  if (tile < P) for (int s = offB[0] + lt; s < offB[kTileColours]; s += ln) store_impulse(s, ldcg4(&W.s_imp[s]), W.s_pc[s]);    // this CTA solved them
  if (nG > 0) for (int s = offG[0] + gt; s < offG[kMaxColours]; s += gn) store_impulse(s, ldcg4(&W.s_imp[s]), W.s_pc[s]);       // (grid barrier after the last global phase)
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
    if (!(body_xf(i) & XF_ISLAND)) continue;
    float4 pos = sPos[i], vel = sVel[i];
    integrate(pos, vel);
    set3(&sPos[i], pos); set3(&sVel[i], vel);
  }
  for (int b = gt; b < W.nBodies; b += gn) {
    if (W.b_tslot[b] >= 0) continue;
    const uint32_t f = W.b_flags[b];
    if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
    float4 pos = ldcg4(&W.b_pos[b]), vel = ldcg4(&W.b_vel[b]);
    integrate(pos, vel);
    stcg4(&W.b_pos[b], pos); stcg4(&W.b_vel[b], vel);
  }
  if (W.tileKinematic || nG > 0) GB();      // (kinematic bodies were integrated in the global arrays: the position rows of every tile read them)
  __syncthreads();                          // integrated bodies and hand-staged position rows: written by one thread, read by any (a tile
                                            // without boundary or global constraints reaches its first position colour with no other barrier)
  if (rowsLocal && W.posIters > 0) mbar_wait(&rowBar, 1);
  solve_stamp(W, 2);
  MARK();
  // position iterations (:206-224) with the per-island early-out flags of k_solve
  for (int it = 0; it < W.posIters; ++it) {
    int* notOk = W.b_posNotOk + it * W.nBodies;
    const int* prev = it > 0 ? W.b_posNotOk + (it - 1) * W.nBodies : nullptr;
    sweep(TM_POS, notOk, prev);     // (a pass starts with publish + grid barrier when islands can span tiles: the flags of the pass before are in)
    MARK();
  }
  if (jointsLocal) for (int x = lt; x < nJSe; x += ln) { const int j = staged_joint(x); W.j_imp[j] = sjImp[x]; W.j_limit[j] = sjLimit[x]; W.j_root[j] = sjRoot[x]; }
  solve_stamp(W, 3);
  // write back + SynchronizeTransform (:227-235), sleep bookkeeping (:241-269); the exchange flags go back to rest
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
      const int b = body_id(i);
      const int xf = body_xf(i);
      if (xf) W.b_xflag[s0 + i] = 0;
      if (!(xf & XF_ISLAND)) continue;
      const float4 pos = xyz0(sPos[i]), vel = xyz0(sVel[i]);
      stcg4(&W.b_pos[b], pos); stcg4(&W.b_vel[b], vel);
      finish(b, W.allowSleep ? W.b_flags[b] : 0u, pos, vel);
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
// dynamic shared memory of k_solve_tiles: everything the SM offers (bodies 36 B each + neighbour slots, then rows of 108 B, joints of 192 B)
size_t tile_smem_bytes(int tileBodies) {
  static size_t avail = 0;
  if (!avail) {
    int dev = 0, optin = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncAttributes fa{}; cudaFuncGetAttributes(&fa, (const void*)k_solve_tiles);
    avail = (size_t)optin > fa.sharedSizeBytes ? (size_t)optin - fa.sharedSizeBytes : 0;
  }
  return std::max(avail, (size_t)tileBodies * 36 + 32 * kTileNbrMax + 4096) > avail ? 0 : avail;       // 0: the tile does not fit
}

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
  CK(launch_mark_and_colour(W, L));          // with W.tiled set, k_mark_solve / k_colour also bin every constraint (dbx_tilekey.cuh)
  {
    const int nBins = 2 * W.nTiles * kTileColours + kMaxColours;
    const size_t smem = (size_t)2 * 1024 * ((((size_t)nBins + 1023) >> 10) | 1) * sizeof(int);
    // (function attributes are per device: a process may hold worlds on several)
    static size_t allowed[64] = {};
    int dev = 0; CK(cudaGetDevice(&dev)); dev &= 63;
    if (smem > allowed[dev]) { CK(cudaFuncSetAttribute((const void*)k_tile_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); allowed[dev] = smem; }
    ++L.launches; k_tile_scan<<<1, 1024, smem, L.stream>>>(W);
  }
  ++L.launches; k_tile_scatter<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}
cudaError_t stage_solve_tiles(const DevWorld& W, const LaunchCfg& L) {
  const size_t smem = tile_smem_bytes(W.tileBodies);
  if (smem == 0) return cudaErrorInvalidConfiguration;
  static bool attr[64] = {};
  int dev = 0; CK(cudaGetDevice(&dev)); dev &= 63;
  if (!attr[dev]) { CK(cudaFuncSetAttribute((const void*)k_solve_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr[dev] = true; }
  if ((L.coopLaunches++ & 1023) == 0) CK(cudaMemsetAsync(&W.hdr->barrier, 0, sizeof(unsigned), L.stream));   // see launch_coop
  void* args[] = {(void*)&W};
  ++L.launches;
  return cudaLaunchCooperativeKernel((const void*)k_solve_tiles, dim3(L.coopBlocks), dim3(kTileThreads), args, smem, L.stream);
}

}  // namespace dbx
