// dbx_solver.cuh — per-constraint device functions of the sequential-impulse solver (contacts, revolute and distance joints).
// Included by dbx_solve.cu (the persistent island solver) and by dbx_kernels.cu (TOI mini-islands).
#pragma once
#include "dbx_util.cuh"
#include "dbx_joints2.cuh"

namespace dbx {

// ------------------------------------------------------------------------------------------------ contact constraints
// b2ContactSolver ctor + InitializeVelocityConstraints (contacts/b2contactsolver.d:244-450): one thread per solver contact.
// `s` = solver slot to fill, `i` = contact slot; warmScale < 0 means "no warm starting" (TOI sub-steps, b2world.d:1419)
// (tile solver: the accumulators are indexed by tile slot, so that a tile reads and clears its own as one contiguous piece;
// every body with mass has a slot.  Either way they are all zero again once the solver has started.)
DBX_D void acc_add(const DevWorld& W, int b, float x, float y, float w) {
  const float k = 4294967296.0f;   // 2^32: scaling a float by a power of two is exact
  const int e = W.tiled ? W.b_tslot[b] : b;
  atomicAdd(&W.b_acc[3 * e + 0], (unsigned long long)__float2ll_rn(x * k));
  atomicAdd(&W.b_acc[3 * e + 1], (unsigned long long)__float2ll_rn(y * k));
  atomicAdd(&W.b_acc[3 * e + 2], (unsigned long long)__float2ll_rn(w * k));
}
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
    // b2ContactSolver.WarmStart (:452-489), order-free: the reference adds each contact's impulse to its two bodies one
    // contact after the other; here every contact adds its velocity delta to a per-body 32.32 fixed-point accumulator
    // (integer atomics commute, so the sum does not depend on the schedule) and k_solve folds the accumulators into the
    // velocities in one pass.  That replaces one barrier-delimited phase per colour by a single one.
    if (warmScale >= 0.0f && (imp.x != 0.0f || imp.y != 0.0f || imp.z != 0.0f || imp.w != 0.0f)) {
      v2 P = V(0.0f, 0.0f); float LA = 0.0f, LB = 0.0f;
      // over the VELOCITY constraint's points: when the conditioning test above drops the block solver, the reference's
      // vc.pointCount becomes 1 (:444-446) and WarmStart (:459) no longer applies the second point's impulse
      for (int k = 0; k < vcCount; ++k) {
        const v2 Pk = (k == 0 ? imp.x : imp.z) * normal + (k == 0 ? imp.y : imp.w) * tangent;
        P += Pk;
        LA += cross(V(r[k].x, r[k].y), Pk);
        LB += cross(V(r[k].z, r[k].w), Pk);
      }
      if (mA != 0.0f || iA != 0.0f) acc_add(W, bA, -(mA * P.x), -(mA * P.y), -(iA * LA));
      if (mB != 0.0f || iB != 0.0f) acc_add(W, bB, mB * P.x, mB * P.y, iB * LB);
    }
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

// velocities of one body pair; only dynamic bodies are ever written (statics/kinematics have zero inverse mass)
struct BodyVel { v2 vA, vB; float wA, wB; };
// Optional CTA-local copy of bodies: vel / pos point into shared memory.  A body reference `ref` in a solver row or joint is
//   * a plain body id (>= 0) when there is no view (k_solve, TOI mini-islands: the global arrays through L2), or with the
//     world-local view of k_solve_worlds (the replica's bodies [off, off + n) live in shared memory);
//   * with the tile view of k_solve_tiles: ref >= 0 is the body's position in tile order, owned by THIS CTA (shared memory at
//     ref - off); ref < 0 carries a body id in its low 31 bits and means "not mine": a static / kinematic body, or a body of
//     another tile that is currently published in the global arrays.
// passed by value.  mode (a literal at every construction site, so the branches below fold away after inlining):
//   0 no view   1 every reference is a plain body id inside [off, off + n) (k_solve_worlds)   2 tile references (k_solve_tiles)
struct BodyView { float4* vel = nullptr; float4* pos = nullptr; int off = 0; int mode = 0; };
constexpr int kRefGlobal = (int)0x80000000;
DBX_D float4 ld_vel(const DevWorld& W, BodyView view, int ref) {
  if (view.mode == 1 || (view.mode == 2 && ref >= 0)) return view.vel[ref - view.off];
  return ldcg4(&W.b_vel[ref & 0x7fffffff]);
}
// (tile solver: the fourth component of a body's slot holds its inverse mass (velocity) / inverse inertia (position); a store
// leaves it alone)
DBX_D void st_vel(const DevWorld& W, BodyView view, int ref, float4 v) {
  if (view.mode == 1) view.vel[ref - view.off] = v;
  else if (view.mode == 2 && ref >= 0) { float4* p = &view.vel[ref - view.off]; p->x = v.x; p->y = v.y; p->z = v.z; }
  else stcg4(&W.b_vel[ref & 0x7fffffff], v);
}
DBX_D float4 ld_pos(const DevWorld& W, BodyView view, int ref) {
  if (view.mode == 1 || (view.mode == 2 && ref >= 0)) return view.pos[ref - view.off];
  return ldcg4(&W.b_pos[ref & 0x7fffffff]);
}
DBX_D void st_pos(const DevWorld& W, BodyView view, int ref, float4 v) {
  if (view.mode == 1) view.pos[ref - view.off] = v;
  else if (view.mode == 2 && ref >= 0) { float4* p = &view.pos[ref - view.off]; p->x = v.x; p->y = v.y; p->z = v.z; }
  else stcg4(&W.b_pos[ref & 0x7fffffff], v);
}
DBX_D BodyVel load_vel(const DevWorld& W, int2 bd, BodyView view = BodyView()) {
  const float4 a = ld_vel(W, view, bd.x), b = ld_vel(W, view, bd.y);
  BodyVel r; r.vA = V(a.x, a.y); r.wA = a.z; r.vB = V(b.x, b.y); r.wB = b.z; return r;
}
DBX_D void store_vel(const DevWorld& W, int2 bd, const BodyVel& r, float mA, float iA, float mB, float iB, BodyView view = BodyView()) {
  if (mA != 0.0f || iA != 0.0f) st_vel(W, view, bd.x, make_float4(r.vA.x, r.vA.y, r.wA, 0.0f));
  if (mB != 0.0f || iB != 0.0f) st_vel(W, view, bd.y, make_float4(r.vB.x, r.vB.y, r.wB, 0.0f));
}

// b2ContactSolver.WarmStart (:452-490)
// b2ContactSolver.SolveVelocityConstraints (:492-772): friction rows, then 1-point clamp or the 2-point block solver
// constraint block of one solver contact, loadable ahead of the barrier that precedes its colour (only s_imp ever changes,
// and only through the thread that owns the slot)
struct VC { int2 bd; int pc; float4 v0, v1, r0, r1, q0, q1, imp, nm, K; };
DBX_D void vc_load(const DevWorld& W, int s, VC& c) {
  c.bd = W.s_body[s]; c.pc = W.s_pc[s];
  c.v0 = W.s_v0[s]; c.v1 = W.s_v1[s]; c.r0 = W.s_r0[s]; c.q0 = W.s_q0[s]; c.imp = W.s_imp[s];
  c.r1 = W.s_r1[s]; c.q1 = W.s_q1[s]; c.nm = W.s_nm[s]; c.K = W.s_k[s];
}
// the row itself: reads and writes the two bodies' velocities, returns the row's new accumulated impulses (n0 t0 n1 t1)
DBX_D float4 contact_velocity_row(const DevWorld& W, const VC& c, BodyView view = BodyView()) {
  const int2 bd = c.bd;
  const float4 v0 = c.v0, v1 = c.v1;
  float4 imp = c.imp;
  const int pointCount = c.pc & 0xFF;
  const float mA = v1.x, iA = v1.y, mB = v1.z, iB = v1.w;
  const float4 r0 = c.r0, q0 = c.q0;
  const float4 r1 = c.r1, q1 = c.q1;
  BodyVel bv = load_vel(W, bd, view);
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
  bv.vA = vA; bv.vB = vB; bv.wA = wA; bv.wB = wB;
  store_vel(W, bd, bv, mA, iA, mB, iB, view);
  return imp;
}
DBX_D void contact_solve_velocity(const DevWorld& W, int s, const VC& c, BodyView view = BodyView()) { W.s_imp[s] = contact_velocity_row(W, c, view); }
DBX_D void contact_solve_velocity(const DevWorld& W, int s, BodyView view = BodyView()) { VC c; vc_load(W, s, c); contact_solve_velocity(W, s, c, view); }

// b2ContactSolver.SolvePositionConstraints (:73-149) + b2PositionSolverManifold (:816-868); returns min separation
// toiA/toiB >= 0 selects SolveTOIPositionConstraints (:152-242): only those two bodies keep their mass, Baumgarte 0.75
// position constraint block of one solver contact (b2ContactPositionConstraint, :801-814)
struct PCn { int2 bd; int pc; float4 v1, p0, p1, p2; float2 p3; };
DBX_D void pcn_load(const DevWorld& W, int s, PCn& c) {
  c.bd = W.s_body[s]; c.pc = W.s_pc[s];
  c.v1 = W.s_v1[s]; c.p0 = W.s_p0[s]; c.p1 = W.s_p1[s]; c.p2 = W.s_p2[s]; c.p3 = W.s_p3[s];
}
DBX_D float contact_position_row(const DevWorld& W, const PCn& c, int toiA = -1, int toiB = -1, BodyView view = BodyView()) {
  const int2 bd = c.bd;
  const float4 v1 = c.v1, p0 = c.p0, p1 = c.p1, p2 = c.p2;
  const float2 p3 = c.p3;
  const int pc = c.pc;
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
  const float4 pa = ld_pos(W, view, bd.x), pb = ld_pos(W, view, bd.y);
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
  if (mA != 0.0f || iA != 0.0f) st_pos(W, view, bd.x, make_float4(cA.x, cA.y, aA, 0.0f));
  if (mB != 0.0f || iB != 0.0f) st_pos(W, view, bd.y, make_float4(cB.x, cB.y, aB, 0.0f));
  return minSeparation;
}
DBX_D float contact_solve_position(const DevWorld& W, int s, int toiA = -1, int toiB = -1, BodyView view = BodyView()) {
  PCn c; pcn_load(W, s, c);
  return contact_position_row(W, c, toiA, toiB, view);
}

// ------------------------------------------------------------------------------------------------ joints
// revolute: dynamics/joints/b2revolutejoint.d:319-636; distance: b2distancejoint.d:211-373
DBX_D bool joint_active(const DevWorld& W, int j);
DBX_D void joint_init(const DevWorld& W, int j, BodyView view = BodyView()) {
  if (!joint_active(W, j)) { W.j_root[j] = -1; return; }   // later phases only look at j_root
  const int4 ids = W.j_ids[j];
  const int bA = ids.y, bB = ids.z;
  const int2 jb = view.mode == 2 ? W.j_bref[j] : make_int2(bA, bB);   // how this joint's two bodies are reached (see BodyView)
  const float4 msA = W.b_mass[bA], msB = W.b_mass[bB];
  const float4 lcA4 = W.b_lc[bA], lcB4 = W.b_lc[bB];
  const float4 anc = W.j_anchor[j];
  const float mA = msA.x, iA = msA.y, mB = msB.x, iB = msB.y;
  const v2 localCenterA = V(lcA4.x, lcA4.y), localCenterB = V(lcB4.x, lcB4.y);
  const float4 posA = W.b_pos[bA], posB = W.b_pos[bB];
  const float4 xA = W.b_xf[bA], xB = W.b_xf[bB];
  float4 velA = ld_vel(W, view, jb.x), velB = ld_vel(W, view, jb.y);
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
  if (ids.x != JT_REVOLUTE && ids.x != JT_DISTANCE) {
    JCtx c; c.mA = mA; c.iA = iA; c.mB = mB; c.iB = iB; c.rA = rA; c.rB = rB; c.cA = V(posA.x, posA.y); c.cB = V(posB.x, posB.y);
    c.aA = aA; c.aB = aB; c.qA = qA; c.qB = qB; c.vA = vA; c.vB = vB; c.wA = wA; c.wB = wB; c.imp = imp;
    joint2_init(W, j, ids.x, ids.w, bB, c);
    vA = c.vA; vB = c.vB; wA = c.wA; wB = c.wB; imp = c.imp;
  } else if (ids.x == JT_REVOLUTE) {
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
  if (mA != 0.0f || iA != 0.0f) st_vel(W, view, jb.x, make_float4(vA.x, vA.y, wA, 0.0f));
  if (mB != 0.0f || iB != 0.0f) st_vel(W, view, jb.y, make_float4(vB.x, vB.y, wB, 0.0f));
}

DBX_D void joint_solve_velocity(const DevWorld& W, int j, BodyView view = BodyView()) {
  const int4 ids = W.j_ids[j];
  const int bA = ids.y, bB = ids.z;
  const int2 jb = view.mode == 2 ? W.j_bref[j] : make_int2(bA, bB);   // how this joint's two bodies are reached (see BodyView)
  const float4 m = W.j_m[j], r = W.j_r[j];
  const float mA = m.x, iA = m.y, mB = m.z, iB = m.w;
  const v2 rA = V(r.x, r.y), rB = V(r.z, r.w);
  float4 velA = ld_vel(W, view, jb.x), velB = ld_vel(W, view, jb.y);
  v2 vA = V(velA.x, velA.y), vB = V(velB.x, velB.y); float wA = velA.z, wB = velB.z;
  float4 imp = W.j_imp[j];
  if (ids.x != JT_REVOLUTE && ids.x != JT_DISTANCE) {
    JCtx c; c.mA = mA; c.iA = iA; c.mB = mB; c.iB = iB; c.rA = rA; c.rB = rB; c.vA = vA; c.vB = vB; c.wA = wA; c.wB = wB; c.imp = imp;
    joint2_solve_velocity(W, j, ids.x, ids.w, c);
    vA = c.vA; vB = c.vB; wA = c.wA; wB = c.wB; imp = c.imp;
  } else if (ids.x == JT_REVOLUTE) {
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
  if (mA != 0.0f || iA != 0.0f) st_vel(W, view, jb.x, make_float4(vA.x, vA.y, wA, 0.0f));
  if (mB != 0.0f || iB != 0.0f) st_vel(W, view, jb.y, make_float4(vB.x, vB.y, wB, 0.0f));
}

// returns true when the joint is within tolerance
DBX_D bool joint_solve_position(const DevWorld& W, int j, BodyView view = BodyView()) {
  const int4 ids = W.j_ids[j];
  const int bA = ids.y, bB = ids.z;
  const int2 jb = view.mode == 2 ? W.j_bref[j] : make_int2(bA, bB);   // how this joint's two bodies are reached (see BodyView)
  const float4 m = W.j_m[j], lc = W.j_lc[j], anc = W.j_anchor[j];
  const float mA = m.x, iA = m.y, mB = m.z, iB = m.w;
  float4 pa = ld_pos(W, view, jb.x), pb = ld_pos(W, view, jb.y);
  v2 cA = V(pa.x, pa.y), cB = V(pb.x, pb.y); float aA = pa.z, aB = pb.z;
  bool ok;
  if (ids.x != JT_REVOLUTE && ids.x != JT_DISTANCE) {
    const Rot qA = rot_from_angle(aA), qB = rot_from_angle(aB);
    const v2 rA = mul(qA, V(anc.x, anc.y) - V(lc.x, lc.y)), rB = mul(qB, V(anc.z, anc.w) - V(lc.z, lc.w));
    ok = joint2_solve_position(W, j, ids.x, ids.w, mA, iA, mB, iB, rA, rB, qA, cA, aA, cB, aB);
  } else if (ids.x == JT_REVOLUTE) {
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
  if (mA != 0.0f || iA != 0.0f) st_pos(W, view, jb.x, make_float4(cA.x, cA.y, aA, 0.0f));
  if (mB != 0.0f || iB != 0.0f) st_pos(W, view, jb.y, make_float4(cB.x, cB.y, aB, 0.0f));
  return ok;
}

DBX_D bool joint_active(const DevWorld& W, int j) {
  const int4 ids = W.j_ids[j];
  if (!(ids.w & 8)) return false;
  uint32_t fa = W.b_flags[ids.y], fb = W.b_flags[ids.z];
  if (!(fa & BF_ACTIVE) || !(fb & BF_ACTIVE)) return false;
  return ((fa & BF_ISLAND) && body_type(fa) != BODY_STATIC) || ((fb & BF_ISLAND) && body_type(fb) != BODY_STATIC);
}

}  // namespace dbx
