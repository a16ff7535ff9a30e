// dbx_solve.cu — constraint setup and the persistent coloured Gauss-Seidel island solver (sm_100a).
// Separate translation unit so that fused multiply-add can be enabled for it alone (csrc/Makefile, SOLVER_FMAD): nothing
// in here has to be bit-identical to the reference (feature keys, manifolds and the pair set come from dbx_kernels.cu).
// Default is un-contracted: FMA bought 1.5 % and cost the bit-identity between the solver's own code paths.
#include "dbx_solver.cuh"

namespace dbx {

__global__ void __launch_bounds__(256) k_prepare(const __grid_constant__ DevWorld W) {
  const int n = min(W.hdr->nSolve, W.sCap);
  const float warmScale = W.warmStarting ? W.dtRatio : -1.0f;
  GRID_STRIDE(s, n) {
    const int i = W.s_contact[s];
    prepare_contact(W, s, i, warmScale);
    if (W.tiled) W.s_body[s] = W.c_bref[i];     // tile solver: how the row reaches its two bodies (dbx_solver.cuh, BodyView)
  }
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
  solve_stamp(W, 0);

  // contacts warm start (b2island.d:138-141): fold the per-body accumulators k_prepare filled into the velocities
  if (W.warmStarting) {
    const float k = 1.0f / 4294967296.0f;
    for (int b = tid; b < W.nBodies; b += nth) {
      const long long ax = (long long)__ldcg(&W.b_acc[3 * b]), ay = (long long)__ldcg(&W.b_acc[3 * b + 1]), aw = (long long)__ldcg(&W.b_acc[3 * b + 2]);
      if ((ax | ay | aw) == 0) continue;
      float4 vel = ldcg4(&W.b_vel[b]);
      vel.x += (float)ax * k; vel.y += (float)ay * k; vel.z += (float)aw * k;
      stcg4(&W.b_vel[b], vel);
      __stcg(&W.b_acc[3 * b], 0ull); __stcg(&W.b_acc[3 * b + 1], 0ull); __stcg(&W.b_acc[3 * b + 2], 0ull);
    }
    GB();
  }
  // joints: InitVelocityConstraints incl. their warm start (:143-146)
  for (int c = 0; c < nJointColours; ++c) {
    int beg = joff[c], end = joff[c + 1];
    if (beg == end) continue;
    if (jointRole) for (int k = beg + jtid; k < end; k += jnth) joint_init(W, k);
    GB();
  }
  // velocity iterations: all joints, then all contacts (:153-161).  With unified colours (W.unifiedColours: contacts on a
  // jointed body only take colours above that body's joint colours, k_mark_solve / k_colour) joint colour c and contact
  // colour c share one phase -- every body still sees its joints before its contacts -- which saves the joint colours'
  // barriers in every pass; otherwise (colour override hook) joint colours run as phases of their own first.
  const bool unified = W.unifiedColours != 0;
  const int nPhases = unified ? max(T, nJointColours) : nJointColours + T;
  VC pre; int preS = -1;
  solve_stamp(W, 1);
  for (int it = 0; it < W.velIters; ++it) {
    for (int p = 0; p < nPhases; ++p) {
      const int jc = unified ? (p < nJointColours ? p : -1) : (p < nJointColours ? p : -1);
      const int c = unified ? (p < T ? p : -1) : (p >= nJointColours ? p - nJointColours : -1);
      bool any = false;
      if (jc >= 0 && joff[jc] != joff[jc + 1]) {
        any = true;
        if (jointRole && !(W.dbgFlags & 1)) for (int k = joff[jc] + jtid; k < joff[jc + 1]; k += jnth) if (W.j_root[k] >= 0) joint_solve_velocity(W, k);
      }
      if (c >= 0 && coff[c] != coff[c + 1]) {
        any = true;
        const int beg = coff[c], end = coff[c + 1];
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
      }
      if (any) GB();
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
  solve_stamp(W, 2);
  // position iterations: contacts then joints, each island stops once all of its constraints are within tolerance
  // (:206-224).  Unified colours run from the highest colour down, so that every body again sees its contacts (higher
  // colours) before its joints; slot q of the pass is the tail colours, a contact colour, a joint colour or both.
  for (int it = 0; it < W.posIters; ++it) {
    int* notOk = W.b_posNotOk + it * W.nBodies;
    const int* prev = it > 0 ? W.b_posNotOk + (it - 1) * W.nBodies : nullptr;
    for (int q = 0; q <= nPhases; ++q) {
      bool tail; int c, jc;
      if (unified) { tail = q == 0; const int p = nPhases - q; c = (!tail && p < T) ? p : -1; jc = (!tail && p < nJointColours) ? p : -1; }
      else { tail = q == T; c = q < T ? q : -1; jc = q > T ? q - T - 1 : -1; }
      if (tail) {
        if (haveTail) {
          if (tailBlock) for (int k = 0; k < nColours - T; ++k) {
            const int tc = unified ? nColours - 1 - k : T + k;
            for (int s = coff[tc] + threadIdx.x; s < coff[tc + 1]; s += blockDim.x) {
              int root = W.s_root[s];
              if (prev && __ldcg(&prev[root]) == 0) continue;
              float minSep = contact_solve_position(W, s);
              if (!(minSep >= -3.0f * kLinearSlop)) notOk[root] = 1;
            }
            __syncthreads();
          }
          GB();
        }
        continue;
      }
      bool any = false;
      if (c >= 0 && coff[c] != coff[c + 1]) {
        any = true;
        if (contactRole) for (int s = coff[c] + ctid; s < coff[c + 1]; s += cnth) {
          int root = W.s_root[s];
          if (prev && __ldcg(&prev[root]) == 0) continue;
          float minSep = contact_solve_position(W, s);
          if (!(minSep >= -3.0f * kLinearSlop)) notOk[root] = 1;
        }
      }
      if (jc >= 0 && joff[jc] != joff[jc + 1]) {
        any = true;
        if (jointRole) for (int k = joff[jc] + jtid; k < joff[jc + 1]; k += jnth) {
          int root = W.j_root[k];
          if (root < 0) continue;
          if (prev && __ldcg(&prev[root]) == 0) continue;
          if (!joint_solve_position(W, k)) notOk[root] = 1;
        }
      }
      if (any) GB();
    }
  }
  solve_stamp(W, 3);
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

// ------------------------------------------------------------------------------------------------ world-local solver
// Batched replicas (dbx_world_replicate) are thousands of small, independent constraint graphs.  Running them through
// k_solve means one grid barrier per colour for ALL of them and every row streamed from HBM on every iteration.  Here one
// CTA takes one replica at a time and runs the whole of b2Island.Solve for it -- warm-start fold, velocity iterations,
// integration, position iterations with the per-island early-out, write-back and sleep -- with __syncthreads() between
// colours: a replica's rows (tens of kB) stay in L1/L2 for all eleven passes and no CTA ever waits for another.
// Same row functions, same colour order, same arithmetic as k_solve: results are bit-identical (tests compare a replica
// with the same world stepped alone through k_solve).  Solver slots arrive sorted by (replica, colour); a replica's joints
// are solved here too, from the global body arrays (see nJC below).
constexpr int kWorldMaxColours = 256;
// kLocal: the replica's body velocities and positions live in shared memory for the duration of its solve (loaded once,
// written back once); otherwise (a replica too big for shared memory) they stay in the global arrays.
template <bool kLocal> __global__ void __launch_bounds__(128) k_solve_worlds(const __grid_constant__ DevWorld W, int bodiesPerWorld) {
  extern __shared__ float4 sBodies[];              // kLocal: [bodiesPerWorld] velocities, then [bodiesPerWorld] positions
  __shared__ int cstart[kWorldMaxColours + 1];
  __shared__ int ncol;
  __shared__ int joff[kMaxJointColours + 1];
  const int t = threadIdx.x, B = blockDim.x;
  // Joints (replicas of a jointed template, e.g. articulated robots): the device joint arrays are colour-major and, inside a
  // colour, replica-major (World::replicate), so replica w's joints of colour c are [joff[c] + w * cnt, + cnt) with
  // cnt = (joff[c + 1] - joff[c]) / nWorlds.  Order as in b2Island.Solve: velocity passes joints then contacts (:153-161), position
  // passes contacts then joints (:206-216); the position colours run downwards like k_solve's unified phases, so that a replica
  // stays bit-identical to the same world stepped alone.  Joint code reads and writes the global body arrays (kLocal is off).
  const int nJC = (!kLocal && W.nJoints > 0) ? min(W.nJointColours, kMaxJointColours) : 0;
  for (int c = t; c <= kMaxJointColours; c += B) joff[c] = W.hdr->jointColourOff[c];
  __syncthreads();
  for (int w = blockIdx.x; w < W.nWorlds; w += gridDim.x) {
    const int beg = W.w_start[w], end = W.w_end[w];
    // (a replica with no solver contact -- asleep, or in free fall -- has beg == end: only the body loops do anything)
    const int b0 = w * bodiesPerWorld, b1 = b0 + bodiesPerWorld;
    BodyView bvw; bvw.vel = sBodies; bvw.pos = sBodies + bodiesPerWorld; bvw.off = b0; bvw.mode = 1;
    const BodyView view = kLocal ? bvw : BodyView();
    // colour boundaries of this replica's slot range (sorted by colour): the slots where the colour changes, in order
    if (t == 0) ncol = 0;
    __syncthreads();
    for (int s = beg + t; s < end; s += B) {
      const unsigned cm = (1u << W.swColourBits) - 1u;
      const int c = (int)(W.sw_key[s] & cm);
      if (s == beg || c != (int)(W.sw_key[s - 1] & cm)) { const int k = atomicAdd(&ncol, 1); if (k < kWorldMaxColours) cstart[k] = s; }
    }
    __syncthreads();
    const int nc = min(ncol, kWorldMaxColours);
    if (ncol > kWorldMaxColours && t == 0) W.hdr->error = E_COLOURS;
    if (t == 0) {                                  // a handful of entries: insertion sort
      for (int a = 1; a < nc; ++a) { const int v = cstart[a]; int b = a - 1; while (b >= 0 && cstart[b] > v) { cstart[b + 1] = cstart[b]; --b; } cstart[b + 1] = v; }
      cstart[nc] = end;
    }
    __syncthreads();
    // bodies in (kLocal); the contact warm start is folded in on the way (see k_solve)
    {
      const float k = 1.0f / 4294967296.0f;
      for (int b = b0 + t; b < b1; b += B) {
        float4 vel = ldcg4(&W.b_vel[b]);
        if (W.warmStarting) {
          const long long ax = (long long)__ldcg(&W.b_acc[3 * b]), ay = (long long)__ldcg(&W.b_acc[3 * b + 1]), aw = (long long)__ldcg(&W.b_acc[3 * b + 2]);
          if ((ax | ay | aw) != 0) {
            vel.x += (float)ax * k; vel.y += (float)ay * k; vel.z += (float)aw * k;
            __stcg(&W.b_acc[3 * b], 0ull); __stcg(&W.b_acc[3 * b + 1], 0ull); __stcg(&W.b_acc[3 * b + 2], 0ull);
            if (!kLocal) stcg4(&W.b_vel[b], vel);
          }
        }
        if (kLocal) { bvw.vel[b - b0] = vel; bvw.pos[b - b0] = ldcg4(&W.b_pos[b]); }
      }
      __syncthreads();
    }
    // joints: InitVelocityConstraints incl. their warm start (b2island.d:143-146), colour by colour
    for (int c = 0; c < nJC; ++c) {
      const int cnt = (joff[c + 1] - joff[c]) / W.nWorlds;
      if (cnt == 0) continue;
      for (int k = joff[c] + w * cnt + t; k < joff[c] + (w + 1) * cnt; k += B) joint_init(W, k);
      __syncthreads();
    }
    for (int it = 0; it < W.velIters; ++it) {
      for (int c = 0; c < nJC; ++c) {
        const int cnt = (joff[c + 1] - joff[c]) / W.nWorlds;
        if (cnt == 0) continue;
        for (int k = joff[c] + w * cnt + t; k < joff[c] + (w + 1) * cnt; k += B) if (W.j_root[k] >= 0) joint_solve_velocity(W, k);
        __syncthreads();
      }
      for (int c = 0; c < nc; ++c) {
        for (int s = cstart[c] + t; s < cstart[c + 1]; s += B) contact_solve_velocity(W, s, view);
        __syncthreads();
      }
    }
    // StoreImpulses + integrate positions
    for (int s = beg + t; s < end; s += B) {
      const int i = W.s_contact[s];
      const int vcCount = W.s_pc[s] & 0xFF;
      const float4 imp = W.s_imp[s];
      float4 old = W.c_imp[i];
      old.x = imp.x; old.y = imp.y;
      if (vcCount == 2) { old.z = imp.z; old.w = imp.w; }
      W.c_imp[i] = old;
    }
    const float h = W.dt;
    for (int b = b0 + t; b < b1; b += B) {
      const uint32_t f = W.b_flags[b];
      if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
      float4 pos = kLocal ? bvw.pos[b - b0] : ldcg4(&W.b_pos[b]), vel = kLocal ? bvw.vel[b - b0] : ldcg4(&W.b_vel[b]);
      v2 c = V(pos.x, pos.y), v = V(vel.x, vel.y);
      float a = pos.z, wv = vel.z;
      v2 translation = h * v;
      if (dot(translation, translation) > kMaxTranslationSquared) { float ratio = kMaxTranslation / len(translation); v *= ratio; }
      float rotation = h * wv;
      if (rotation * rotation > kMaxRotationSquared) { float ratio = kMaxRotation / fabsr(rotation); wv *= ratio; }
      c += h * v;
      a += h * wv;
      if (kLocal) { bvw.pos[b - b0] = make_float4(c.x, c.y, a, 0.0f); bvw.vel[b - b0] = make_float4(v.x, v.y, wv, 0.0f); }
      else { stcg4(&W.b_pos[b], make_float4(c.x, c.y, a, 0.0f)); stcg4(&W.b_vel[b], make_float4(v.x, v.y, wv, 0.0f)); }
    }
    __syncthreads();
    for (int it = 0; it < W.posIters; ++it) {
      int* notOk = W.b_posNotOk + it * W.nBodies;
      const int* prev = it > 0 ? W.b_posNotOk + (it - 1) * W.nBodies : nullptr;
      for (int q = 0; q < nc; ++q) {
        const int c = nJC > 0 ? nc - 1 - q : q;        // jointed worlds: downwards, like k_solve's unified phases
        for (int s = cstart[c] + t; s < cstart[c + 1]; s += B) {
          const int root = W.s_root[s];
          if (prev && __ldcg(&prev[root]) == 0) continue;
          const float minSep = contact_solve_position(W, s, -1, -1, view);
          if (!(minSep >= -3.0f * kLinearSlop)) __stcg(&notOk[root], 1);
        }
        __syncthreads();
      }
      for (int c = nJC - 1; c >= 0; --c) {
        const int cnt = (joff[c + 1] - joff[c]) / W.nWorlds;
        if (cnt == 0) continue;
        for (int k = joff[c] + w * cnt + t; k < joff[c] + (w + 1) * cnt; k += B) {
          const int root = W.j_root[k];
          if (root < 0) continue;
          if (prev && __ldcg(&prev[root]) == 0) continue;
          if (!joint_solve_position(W, k)) __stcg(&notOk[root], 1);
        }
        __syncthreads();
      }
    }
    // write back + SynchronizeTransform, sleep bookkeeping
    {
      const float linTolSqr = kLinearSleepTolerance * kLinearSleepTolerance;
      const float angTolSqr = kAngularSleepTolerance * kAngularSleepTolerance;
      for (int b = b0 + t; b < b1; b += B) {
        const uint32_t f = W.b_flags[b];
        if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
        const float4 pos = kLocal ? bvw.pos[b - b0] : ldcg4(&W.b_pos[b]);
        const float4 vel = kLocal ? bvw.vel[b - b0] : ldcg4(&W.b_vel[b]);
        if (kLocal) { stcg4(&W.b_pos[b], pos); stcg4(&W.b_vel[b], vel); }
        const float4 lc = W.b_lc[b];
        W.b_xf[b] = pack(xf_from_sweep(V(pos.x, pos.y), pos.z, V(lc.x, lc.y)));
        if (W.allowSleep) {
          float2 gs = W.b_gs[b];
          if (!(f & BF_AUTOSLEEP) || vel.z * vel.z > angTolSqr || dot(V(vel.x, vel.y), V(vel.x, vel.y)) > linTolSqr) gs.y = 0.0f;
          else gs.y += h;
          W.b_gs[b] = gs;
          atomicMin(&W.b_islMinSleep[W.b_root[b]], __float_as_int(gs.y));
        }
      }
    }
    __syncthreads();
    if (W.allowSleep) {
      const int* last = W.posIters > 0 ? W.b_posNotOk + (W.posIters - 1) * W.nBodies : nullptr;
      for (int b = b0 + t; b < b1; b += B) {
        const uint32_t f = W.b_flags[b];
        if ((f & (BF_ALIVE | BF_ISLAND)) != (BF_ALIVE | BF_ISLAND)) continue;
        const int root = W.b_root[b];
        const bool positionSolved = last && __ldcg(&last[root]) == 0;
        const float minSleep = __int_as_float(__ldcg(&W.b_islMinSleep[root]));
        if (minSleep >= kTimeToSleep && positionSolved) {
          W.b_flags[b] = f & ~BF_AWAKE;
          W.b_gs[b].y = 0.0f;
          W.b_vel[b] = make_float4(0, 0, 0, 0);
          W.b_force[b] = make_float4(0, 0, 0, 0);
        }
      }
    }
    __syncthreads();
  }
}
#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

cudaError_t stage_solve_worlds(const DevWorld& W, const LaunchCfg& L, int bodiesPerWorld) {
  const int grid = W.nWorlds < L.coopBlocks * 8 ? W.nWorlds : L.coopBlocks * 8;
  const size_t smem = (size_t)bodiesPerWorld * 32;
  ++L.launches;
  if (smem <= 40 * 1024 && W.nJoints == 0) k_solve_worlds<true><<<grid, 128, smem, L.stream>>>(W, bodiesPerWorld);
  else k_solve_worlds<false><<<grid, 128, 0, L.stream>>>(W, bodiesPerWorld);
  return cudaGetLastError();
}
cudaError_t stage_prepare(const DevWorld& W, const LaunchCfg& L) {
  ++L.launches; k_prepare<<<L.gridWide, 256, 0, L.stream>>>(W);
  return cudaGetLastError();
}
cudaError_t stage_solve(const DevWorld& W, const LaunchCfg& L) {
  if ((L.coopLaunches++ & 1023) == 0) CK(cudaMemsetAsync(&W.hdr->barrier, 0, sizeof(unsigned), L.stream));   // see launch_coop
  void* args[] = {(void*)&W};
  ++L.launches;
  return cudaLaunchCooperativeKernel((const void*)k_solve, dim3(L.coopBlocks), dim3(L.coopThreads), args, 0, L.stream);
}

}  // namespace dbx
