// dbx_joints2.cuh — the second-wave joints: prismatic, weld, wheel, rope, friction, motor, mouse,
// pulley.  Same three hooks as every reference joint (InitVelocityConstraints / SolveVelocityConstraints /
// SolvePositionConstraints), run by k_solve inside the joint colours.  Per-joint device record (dbx_device.cuh):
//   j_p0 / j_p1   parameters, packed per type by World::push (table in dbx_world.cu)
//   j_imp         accumulated impulses (warm start); j_limit  limit state
//   j_k0 .. j_k3  per-step temporaries written by init, read by the velocity iterations
#pragma once
#include "dbx_util.cuh"

namespace dbx {

enum { JT_REVOLUTE_ = 1, JT_PRISMATIC = 2, JT_PULLEY = 4, JT_MOUSE = 5, JT_GEAR = 6, JT_WHEEL = 7, JT_WELD = 8, JT_FRICTION = 9, JT_ROPE = 10, JT_MOTOR = 11 };

struct JCtx {
  float mA, iA, mB, iB;
  v2 rA, rB, cA, cB; float aA, aB; Rot qA, qB;
  v2 vA, vB; float wA, wB;
  float4 imp;
};
DBX_D void japply(JCtx& c, v2 P, float LA, float LB) { c.vA -= c.mA * P; c.wA -= c.iA * LA; c.vB += c.mB * P; c.wB += c.iB * LB; }

// point-to-point effective mass shared by friction and motor (b2frictionjoint.d:205-213)
DBX_D float4 point_mass_inverse(const JCtx& c) {
  M22 K;
  K.ex.x = c.mA + c.mB + c.iA * c.rA.y * c.rA.y + c.iB * c.rB.y * c.rB.y;
  K.ex.y = -c.iA * c.rA.x * c.rA.y - c.iB * c.rB.x * c.rB.y;
  K.ey.x = K.ex.y;
  K.ey.y = c.mA + c.mB + c.iA * c.rA.x * c.rA.x + c.iB * c.rB.x * c.rB.x;
  M22 inv = inverse(K);
  return make_float4(inv.ex.x, inv.ex.y, inv.ey.x, inv.ey.y);
}
DBX_D v2 mul22(float4 m, v2 v) { return V(m.x * v.x + m.z * v.y, m.y * v.x + m.w * v.y); }   // (ex.x ex.y ey.x ey.y) * v
// b2Mat33 helpers on three float4 columns (common/b2math.d:419-466)
DBX_D void sym_inverse33(v3 ex, v3 ey, v3 ez, v3& ox, v3& oy, v3& oz) {
  float det = dot(ex, cross(ey, ez));
  if (det != 0.0f) det = 1.0f / det;
  const float a11 = ex.x, a12 = ey.x, a13 = ez.x, a22 = ey.y, a23 = ez.y, a33 = ez.z;
  ox.x = det * (a22 * a33 - a23 * a23);
  ox.y = det * (a13 * a23 - a12 * a33);
  ox.z = det * (a12 * a23 - a13 * a22);
  oy.x = ox.y;
  oy.y = det * (a11 * a33 - a13 * a13);
  oy.z = det * (a13 * a12 - a11 * a23);
  oz.x = ox.z;
  oz.y = oy.z;
  oz.z = det * (a11 * a22 - a12 * a12);
}
DBX_D void inverse22_of33(v3 ex, v3 ey, v3& ox, v3& oy, v3& oz) {
  const float a = ex.x, b = ey.x, c = ex.y, d = ey.y;
  float det = a * d - b * c;
  if (det != 0.0f) det = 1.0f / det;
  ox = V3(det * d, -det * c, 0.0f);
  oy = V3(-det * b, det * a, 0.0f);
  oz = V3(0.0f, 0.0f, 0.0f);
}
DBX_D void weld_k(const float mA, const float iA, const float mB, const float iB, v2 rA, v2 rB, v3& ex, v3& ey, v3& ez) {
  ex.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
  ey.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
  ez.x = -rA.y * iA - rB.y * iB;
  ex.y = ey.x;
  ey.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
  ez.y = rA.x * iA + rB.x * iB;
  ex.z = ez.x;
  ey.z = ez.y;
  ez.z = iA + iB;
}

// ------------------------------------------------------------------------------------------------ gear (b2gearjoint.d:245-500)
// bodies A, B (the joint's own pair) move through JCtx; C, D (the other ends of joint1 / joint2) are handled here.
// p0 = (localAnchorC, localAnchorD), p1 = (localAxisC, localAxisD), p2 = (referenceAngleA, referenceAngleB, ratio, constant)
// k0 = (JvAC, JvBD), k1 = (JwA, JwB, JwC, JwD), k2.x = mass, k3 = (mC, iC, mD, iD)
struct GearJ { v2 JvAC, JvBD; float JwA, JwB, JwC, JwD, mass; };
DBX_D GearJ gear_jacobian(int typeA, int typeB, float ratio, float4 p0, float4 p1, Rot qC, Rot qD, v2 rA, v2 rB, v2 lcC, v2 lcD,
                          float mA, float iA, float mB, float iB, float mC, float iC, float mD, float iD) {
  GearJ g; g.mass = 0.0f;
  if (typeA == JT_REVOLUTE_) { g.JvAC = V(0.0f, 0.0f); g.JwA = 1.0f; g.JwC = 1.0f; g.mass += iA + iC; }
  else {
    const v2 u = mul(qC, V(p1.x, p1.y));
    const v2 rC = mul(qC, V(p0.x, p0.y) - lcC);
    g.JvAC = u; g.JwC = cross(rC, u); g.JwA = cross(rA, u);
    g.mass += mC + mA + iC * g.JwC * g.JwC + iA * g.JwA * g.JwA;
  }
  if (typeB == JT_REVOLUTE_) { g.JvBD = V(0.0f, 0.0f); g.JwB = ratio; g.JwD = ratio; g.mass += ratio * ratio * (iB + iD); }
  else {
    const v2 u = mul(qD, V(p1.z, p1.w));
    const v2 rD = mul(qD, V(p0.z, p0.w) - lcD);
    g.JvBD = ratio * u; g.JwD = ratio * cross(rD, u); g.JwB = ratio * cross(rB, u);
    g.mass += ratio * ratio * (mD + mB) + iD * g.JwD * g.JwD + iB * g.JwB * g.JwB;
  }
  return g;
}
DBX_D void gear_init(const DevWorld& W, int j, JCtx& c) {
  const int4 id2 = W.j_ids2[j];
  const int bC = id2.x, bD = id2.y;
  const float4 p0 = W.j_p0[j], p1 = W.j_p1[j], p2 = W.j_p2[j];
  const float4 msC = W.b_mass[bC], msD = W.b_mass[bD], xC = W.b_xf[bC], xD = W.b_xf[bD], lcC4 = W.b_lc[bC], lcD4 = W.b_lc[bD];
  float4 velC = ldcg4(&W.b_vel[bC]), velD = ldcg4(&W.b_vel[bD]);
  const float mC = msC.x, iC = msC.y, mD = msD.x, iD = msD.y;
  GearJ g = gear_jacobian(id2.z, id2.w, p2.z, p0, p1, R(xC.z, xC.w), R(xD.z, xD.w), c.rA, c.rB, V(lcC4.x, lcC4.y), V(lcD4.x, lcD4.y),
                          c.mA, c.iA, c.mB, c.iB, mC, iC, mD, iD);
  g.mass = g.mass > 0.0f ? 1.0f / g.mass : 0.0f;
  W.j_k0[j] = make_float4(g.JvAC.x, g.JvAC.y, g.JvBD.x, g.JvBD.y);
  W.j_k1[j] = make_float4(g.JwA, g.JwB, g.JwC, g.JwD);
  W.j_k2[j] = make_float4(g.mass, 0.0f, 0.0f, 0.0f);
  W.j_k3[j] = make_float4(mC, iC, mD, iD);
  if (W.warmStarting) {
    const float imp = c.imp.x;     // the reference does not scale the gear impulse by dtRatio (b2gearjoint.d:329-339)
    c.vA += (c.mA * imp) * g.JvAC; c.wA += c.iA * imp * g.JwA;
    c.vB += (c.mB * imp) * g.JvBD; c.wB += c.iB * imp * g.JwB;
    if (mC != 0.0f || iC != 0.0f) { v2 vC = V(velC.x, velC.y) - (mC * imp) * g.JvAC; stcg4(&W.b_vel[bC], make_float4(vC.x, vC.y, velC.z - iC * imp * g.JwC, 0.0f)); }
    if (mD != 0.0f || iD != 0.0f) { v2 vD = V(velD.x, velD.y) - (mD * imp) * g.JvBD; stcg4(&W.b_vel[bD], make_float4(vD.x, vD.y, velD.z - iD * imp * g.JwD, 0.0f)); }
  } else c.imp.x = 0.0f;
}
DBX_D void gear_solve_velocity(const DevWorld& W, int j, JCtx& c) {
  const int4 id2 = W.j_ids2[j];
  const int bC = id2.x, bD = id2.y;
  const float4 k0 = W.j_k0[j], k1 = W.j_k1[j], k3 = W.j_k3[j];
  const float mass = W.j_k2[j].x;
  const v2 JvAC = V(k0.x, k0.y), JvBD = V(k0.z, k0.w);
  const float4 velC = ldcg4(&W.b_vel[bC]), velD = ldcg4(&W.b_vel[bD]);
  v2 vC = V(velC.x, velC.y), vD = V(velD.x, velD.y); float wC = velC.z, wD = velD.z;
  float Cdot = dot(JvAC, c.vA - vC) + dot(JvBD, c.vB - vD);
  Cdot += (k1.x * c.wA - k1.z * wC) + (k1.y * c.wB - k1.w * wD);
  const float impulse = -mass * Cdot;
  c.imp.x += impulse;
  c.vA += (c.mA * impulse) * JvAC; c.wA += c.iA * impulse * k1.x;
  c.vB += (c.mB * impulse) * JvBD; c.wB += c.iB * impulse * k1.y;
  vC -= (k3.x * impulse) * JvAC; wC -= k3.y * impulse * k1.z;
  vD -= (k3.z * impulse) * JvBD; wD -= k3.w * impulse * k1.w;
  if (k3.x != 0.0f || k3.y != 0.0f) stcg4(&W.b_vel[bC], make_float4(vC.x, vC.y, wC, 0.0f));
  if (k3.z != 0.0f || k3.w != 0.0f) stcg4(&W.b_vel[bD], make_float4(vD.x, vD.y, wD, 0.0f));
}
DBX_D bool gear_solve_position(const DevWorld& W, int j, float mA, float iA, float mB, float iB, v2 rA, v2 rB, v2& cA, float& aA, v2& cB, float& aB) {
  const int4 id2 = W.j_ids2[j];
  const int bC = id2.x, bD = id2.y;
  const float4 p0 = W.j_p0[j], p1 = W.j_p1[j], p2 = W.j_p2[j], k3 = W.j_k3[j];
  const float4 lcC4 = W.b_lc[bC], lcD4 = W.b_lc[bD];
  const v2 lcC = V(lcC4.x, lcC4.y), lcD = V(lcD4.x, lcD4.y);
  float4 pc = ldcg4(&W.b_pos[bC]), pd = ldcg4(&W.b_pos[bD]);
  v2 cC = V(pc.x, pc.y), cD = V(pd.x, pd.y); float aC = pc.z, aD = pd.z;
  const Rot qC = rot_from_angle(aC), qD = rot_from_angle(aD);
  const float ratio = p2.z;
  const GearJ g = gear_jacobian(id2.z, id2.w, ratio, p0, p1, qC, qD, rA, rB, lcC, lcD, mA, iA, mB, iB, k3.x, k3.y, k3.z, k3.w);
  float coordinateA, coordinateB;
  if (id2.z == JT_REVOLUTE_) coordinateA = aA - aC - p2.x;
  else { const v2 pC = V(p0.x, p0.y) - lcC; const v2 pA = mulT(qC, rA + (cA - cC)); coordinateA = dot(pA - pC, V(p1.x, p1.y)); }
  if (id2.w == JT_REVOLUTE_) coordinateB = aB - aD - p2.y;
  else { const v2 pD = V(p0.z, p0.w) - lcD; const v2 pB = mulT(qD, rB + (cB - cD)); coordinateB = dot(pB - pD, V(p1.z, p1.w)); }
  const float C = (coordinateA + ratio * coordinateB) - p2.w;
  float impulse = 0.0f;
  if (g.mass > 0.0f) impulse = -C / g.mass;
  cA += mA * impulse * g.JvAC; aA += iA * impulse * g.JwA;
  cB += mB * impulse * g.JvBD; aB += iB * impulse * g.JwB;
  cC -= k3.x * impulse * g.JvAC; aC -= k3.y * impulse * g.JwC;
  cD -= k3.z * impulse * g.JvBD; aD -= k3.w * impulse * g.JwD;
  if (k3.x != 0.0f || k3.y != 0.0f) stcg4(&W.b_pos[bC], make_float4(cC.x, cC.y, aC, 0.0f));
  if (k3.z != 0.0f || k3.w != 0.0f) stcg4(&W.b_pos[bD], make_float4(cD.x, cD.y, aD, 0.0f));
  return true;     // linearError stays 0 in the reference (b2gearjoint.d:393, 499)
}

// ------------------------------------------------------------------------------------------------ InitVelocityConstraints
DBX_D void joint2_init(const DevWorld& W, int j, int type, int flags, int bB, JCtx& c) {
  const float4 p0 = W.j_p0[j], p1 = W.j_p1[j];
  const bool enableLimit = (flags & 2) != 0, enableMotor = (flags & 4) != 0;
  const float h = W.dt;
  float4& imp = c.imp;
  if (type == JT_ROPE) {                     // b2ropejoint.d:165-237
    v2 u = c.cB + c.rB - c.cA - c.rA;
    const float length = len(u);
    const float C = length - p0.x;
    W.j_limit[j] = C > 0.0f ? LIM_UPPER : LIM_INACTIVE;
    if (length > kLinearSlop) {
      u *= 1.0f / length;
      const float crA = cross(c.rA, u), crB = cross(c.rB, u);
      const float invMass = c.mA + c.iA * crA * crA + c.mB + c.iB * crB * crB;
      const float mass = invMass != 0.0f ? 1.0f / invMass : 0.0f;
      W.j_k0[j] = make_float4(u.x, u.y, length, mass);
      if (W.warmStarting) { imp.x *= W.dtRatio; const v2 P = imp.x * u; japply(c, P, cross(c.rA, P), cross(c.rB, P)); }
      else imp.x = 0.0f;
    } else {
      W.j_k0[j] = make_float4(0.0f, 0.0f, length, 0.0f);
      imp.x = 0.0f;
    }
  } else if (type == JT_WELD) {              // b2weldjoint.d:196-285; p0 = (referenceAngle, frequencyHz, dampingRatio, -)
    v3 ex, ey, ez, mx, my, mz;
    weld_k(c.mA, c.iA, c.mB, c.iB, c.rA, c.rB, ex, ey, ez);
    float gamma = 0.0f, bias = 0.0f;
    if (p0.y > 0.0f) {
      inverse22_of33(ex, ey, mx, my, mz);
      float invM = c.iA + c.iB;
      const float m = invM > 0.0f ? 1.0f / invM : 0.0f;
      const float C = c.aB - c.aA - p0.x;
      const float omega = 2.0f * kPi * p0.y;
      const float d = 2.0f * m * p0.z * omega;
      const float k = m * omega * omega;
      gamma = h * (d + h * k);
      gamma = gamma != 0.0f ? 1.0f / gamma : 0.0f;
      bias = C * h * k * gamma;
      invM += gamma;
      mz.z = invM != 0.0f ? 1.0f / invM : 0.0f;
    } else if (ez.z == 0.0f) {
      inverse22_of33(ex, ey, mx, my, mz);
    } else {
      sym_inverse33(ex, ey, ez, mx, my, mz);
    }
    W.j_k0[j] = make_float4(mx.x, mx.y, mx.z, gamma);
    W.j_k1[j] = make_float4(my.x, my.y, my.z, bias);
    W.j_k2[j] = make_float4(mz.x, mz.y, mz.z, 0.0f);
    if (W.warmStarting) {
      imp.x *= W.dtRatio; imp.y *= W.dtRatio; imp.z *= W.dtRatio;
      const v2 P = V(imp.x, imp.y);
      japply(c, P, cross(c.rA, P) + imp.z, cross(c.rB, P) + imp.z);
    } else imp = make_float4(0, 0, 0, 0);
  } else if (type == JT_FRICTION || type == JT_MOTOR) {   // b2frictionjoint.d:178-232, b2motorjoint.d:223-283
    W.j_k0[j] = point_mass_inverse(c);
    float angularMass = c.iA + c.iB;
    if (angularMass > 0.0f) angularMass = 1.0f / angularMass;
    if (type == JT_MOTOR) {                  // p0 = (linearOffset.x, .y, angularOffset, correctionFactor)
      const v2 linearError = c.cB + c.rB - c.cA - c.rA - mul(c.qA, V(p0.x, p0.y));
      W.j_k1[j] = make_float4(angularMass, linearError.x, linearError.y, c.aB - c.aA - p0.z);
    } else {
      W.j_k1[j] = make_float4(angularMass, 0.0f, 0.0f, 0.0f);
    }
    if (W.warmStarting) {
      imp.x *= W.dtRatio; imp.y *= W.dtRatio; imp.z *= W.dtRatio;
      const v2 P = V(imp.x, imp.y);
      japply(c, P, cross(c.rA, P) + imp.z, cross(c.rB, P) + imp.z);
    } else imp = make_float4(0, 0, 0, 0);
  } else if (type == JT_MOUSE) {             // b2mousejoint.d:190-262; p0 = (target.x, .y, maxForce, frequencyHz), p1.x = dampingRatio
    const float mass = W.b_mass[bB].z;
    const float omega = 2.0f * kPi * p0.w;
    const float d = 2.0f * mass * p1.x * omega;
    const float k = mass * (omega * omega);
    float gamma = h * (d + h * k);
    if (gamma != 0.0f) gamma = 1.0f / gamma;
    const float beta = h * k * gamma;
    M22 K;
    K.ex.x = c.mB + c.iB * c.rB.y * c.rB.y + gamma;
    K.ex.y = -c.iB * c.rB.x * c.rB.y;
    K.ey.x = K.ex.y;
    K.ey.y = c.mB + c.iB * c.rB.x * c.rB.x + gamma;
    const M22 inv = inverse(K);
    v2 C = c.cB + c.rB - V(p0.x, p0.y);
    C *= beta;
    W.j_k0[j] = make_float4(inv.ex.x, inv.ex.y, inv.ey.x, inv.ey.y);
    W.j_k1[j] = make_float4(C.x, C.y, gamma, 0.0f);
    c.wB *= 0.98f;
    if (W.warmStarting) {
      imp.x *= W.dtRatio; imp.y *= W.dtRatio;
      const v2 P = V(imp.x, imp.y);
      c.vB += c.mB * P; c.wB += c.iB * cross(c.rB, P);
    } else imp = make_float4(0, 0, 0, 0);
  } else if (type == JT_PRISMATIC) {         // b2prismaticjoint.d:391-533; p0 = (axis.x, axis.y, referenceAngle, maxMotorForce), p1 = (motorSpeed, lower, upper, -)
    const v2 d = (c.cB - c.cA) + c.rB - c.rA;
    const v2 localX = V(p0.x, p0.y), localY = cross(1.0f, localX);
    const v2 axis = mul(c.qA, localX);
    const float a1 = cross(d + c.rA, axis), a2 = cross(c.rB, axis);
    float motorMass = c.mA + c.mB + c.iA * a1 * a1 + c.iB * a2 * a2;
    if (motorMass > 0.0f) motorMass = 1.0f / motorMass;
    const v2 perp = mul(c.qA, localY);
    const float s1 = cross(d + c.rA, perp), s2 = cross(c.rB, perp);
    const float k11 = c.mA + c.mB + c.iA * s1 * s1 + c.iB * s2 * s2;
    const float k12 = c.iA * s1 + c.iB * s2;
    const float k13 = c.iA * s1 * a1 + c.iB * s2 * a2;
    float k22 = c.iA + c.iB;
    if (k22 == 0.0f) k22 = 1.0f;
    const float k23 = c.iA * a1 + c.iB * a2;
    const float k33 = c.mA + c.mB + c.iA * a1 * a1 + c.iB * a2 * a2;
    int limitState = W.j_limit[j];
    if (enableLimit) {
      const float jointTranslation = dot(axis, d);
      if (fabsr(p1.z - p1.y) < 2.0f * kLinearSlop) limitState = LIM_EQUAL;
      else if (jointTranslation <= p1.y) { if (limitState != LIM_LOWER) { limitState = LIM_LOWER; imp.z = 0.0f; } }
      else if (jointTranslation >= p1.z) { if (limitState != LIM_UPPER) { limitState = LIM_UPPER; imp.z = 0.0f; } }
      else { limitState = LIM_INACTIVE; imp.z = 0.0f; }
    } else { limitState = LIM_INACTIVE; imp.z = 0.0f; }
    W.j_limit[j] = limitState;
    if (!enableMotor) imp.w = 0.0f;
    W.j_k0[j] = make_float4(axis.x, axis.y, perp.x, perp.y);
    W.j_k1[j] = make_float4(s1, s2, a1, a2);
    W.j_k2[j] = make_float4(k11, k12, k13, k22);
    W.j_k3[j] = make_float4(k23, k33, motorMass, 0.0f);
    if (W.warmStarting) {
      imp.x *= W.dtRatio; imp.y *= W.dtRatio; imp.z *= W.dtRatio; imp.w *= W.dtRatio;
      const v2 P = imp.x * perp + (imp.w + imp.z) * axis;
      const float LA = imp.x * s1 + imp.y + (imp.w + imp.z) * a1;
      const float LB = imp.x * s2 + imp.y + (imp.w + imp.z) * a2;
      japply(c, P, LA, LB);
    } else imp = make_float4(0, 0, 0, 0);
  } else if (type == JT_WHEEL) {             // b2wheeljoint.d:302-441; p0 = (axis.x, axis.y, maxMotorTorque, motorSpeed), p1 = (frequencyHz, dampingRatio, -, -)
    // imp = (impulse, springImpulse, -, motorImpulse)
    const v2 d = c.cB + c.rB - c.cA - c.rA;
    const v2 localX = V(p0.x, p0.y), localY = cross(1.0f, localX);
    const v2 ay = mul(c.qA, localY);
    const float sAy = cross(d + c.rA, ay), sBy = cross(c.rB, ay);
    float mass = c.mA + c.mB + c.iA * sAy * sAy + c.iB * sBy * sBy;
    if (mass > 0.0f) mass = 1.0f / mass;
    float springMass = 0.0f, bias = 0.0f, gamma = 0.0f;
    // m_ax / m_sAx / m_sBx are only refreshed while the spring is on (b2wheeljoint.d:341-345); otherwise the previous
    // values (zero for a joint that never had one) stay in the record, as they stay in the reference object
    float4 k0 = W.j_k0[j], k1 = W.j_k1[j];
    if (p1.x > 0.0f) {
      const v2 ax = mul(c.qA, localX);
      const float sAx = cross(d + c.rA, ax), sBx = cross(c.rB, ax);
      k0.x = ax.x; k0.y = ax.y; k1.x = sAx; k1.y = sBx;
      const float invMass = c.mA + c.mB + c.iA * sAx * sAx + c.iB * sBx * sBx;
      if (invMass > 0.0f) {
        springMass = 1.0f / invMass;
        const float C = dot(d, ax);
        const float omega = 2.0f * kPi * p1.x;
        const float dd = 2.0f * springMass * p1.y * omega;
        const float k = springMass * omega * omega;
        gamma = h * (dd + h * k);
        if (gamma > 0.0f) gamma = 1.0f / gamma;
        bias = C * h * k * gamma;
        springMass = invMass + gamma;
        if (springMass > 0.0f) springMass = 1.0f / springMass;
      }
    } else imp.y = 0.0f;
    float motorMass = 0.0f;
    if (enableMotor) { motorMass = c.iA + c.iB; if (motorMass > 0.0f) motorMass = 1.0f / motorMass; }
    else imp.w = 0.0f;
    k0.z = ay.x; k0.w = ay.y; k1.z = sAy; k1.w = sBy;
    W.j_k0[j] = k0; W.j_k1[j] = k1;
    W.j_k2[j] = make_float4(mass, motorMass, springMass, bias);
    W.j_k3[j] = make_float4(gamma, 0.0f, 0.0f, 0.0f);
    if (W.warmStarting) {
      imp.x *= W.dtRatio; imp.y *= W.dtRatio; imp.w *= W.dtRatio;
      const v2 ax = V(k0.x, k0.y);
      const v2 P = imp.x * ay + imp.y * ax;
      const float LA = imp.x * sAy + imp.y * k1.x + imp.w;
      const float LB = imp.x * sBy + imp.y * k1.y + imp.w;
      japply(c, P, LA, LB);
    } else imp = make_float4(0, 0, 0, 0);
  } else if (type == JT_GEAR) {
    gear_init(W, j, c);
  } else if (type == JT_PULLEY) {            // b2pulleyjoint.d:238-326; p0 = (groundA, groundB), p1 = (lengthA, lengthB, ratio, constant)
    v2 uA = c.cA + c.rA - V(p0.x, p0.y), uB = c.cB + c.rB - V(p0.z, p0.w);
    const float lengthA = len(uA), lengthB = len(uB);
    if (lengthA > 10.0f * kLinearSlop) uA *= 1.0f / lengthA; else uA = V(0.0f, 0.0f);
    if (lengthB > 10.0f * kLinearSlop) uB *= 1.0f / lengthB; else uB = V(0.0f, 0.0f);
    const float ruA = cross(c.rA, uA), ruB = cross(c.rB, uB);
    const float mA = c.mA + c.iA * ruA * ruA, mB = c.mB + c.iB * ruB * ruB;
    float mass = mA + p1.z * p1.z * mB;
    if (mass > 0.0f) mass = 1.0f / mass;
    W.j_k0[j] = make_float4(uA.x, uA.y, uB.x, uB.y);
    W.j_k1[j] = make_float4(mass, 0.0f, 0.0f, 0.0f);
    if (W.warmStarting) {
      imp.x *= W.dtRatio;
      const v2 PA = -(imp.x) * uA, PB = (-p1.z * imp.x) * uB;
      c.vA += c.mA * PA; c.wA += c.iA * cross(c.rA, PA);
      c.vB += c.mB * PB; c.wB += c.iB * cross(c.rB, PB);
    } else imp.x = 0.0f;
  }
}

// ------------------------------------------------------------------------------------------------ SolveVelocityConstraints
DBX_D void joint2_solve_velocity(const DevWorld& W, int j, int type, int flags, JCtx& c) {
  const float4 p0 = W.j_p0[j], p1 = W.j_p1[j];
  const float4 k0 = W.j_k0[j], k1 = W.j_k1[j];
  const bool enableLimit = (flags & 2) != 0, enableMotor = (flags & 4) != 0;
  const float h = W.dt, inv_h = W.inv_dt;
  float4& imp = c.imp;
  if (type == JT_ROPE) {                     // b2ropejoint.d:239-281
    const v2 u = V(k0.x, k0.y);
    const v2 vpA = c.vA + cross(c.wA, c.rA), vpB = c.vB + cross(c.wB, c.rB);
    const float C = k0.z - p0.x;
    float Cdot = dot(u, vpB - vpA);
    if (C < 0.0f) Cdot += inv_h * C;
    float impulse = -k0.w * Cdot;
    const float oldImpulse = imp.x;
    imp.x = fminr(0.0f, imp.x + impulse);
    impulse = imp.x - oldImpulse;
    const v2 P = impulse * u;
    japply(c, P, cross(c.rA, P), cross(c.rB, P));
  } else if (type == JT_WELD) {              // b2weldjoint.d:287-353
    const float4 k2 = W.j_k2[j];
    const v3 mx = V3(k0.x, k0.y, k0.z), my = V3(k1.x, k1.y, k1.z), mz = V3(k2.x, k2.y, k2.z);
    if (p0.y > 0.0f) {
      const float Cdot2 = c.wB - c.wA;
      const float impulse2 = -mz.z * (Cdot2 + k1.w + k0.w * imp.z);
      imp.z += impulse2;
      c.wA -= c.iA * impulse2; c.wB += c.iB * impulse2;
      const v2 Cdot1 = c.vB + cross(c.wB, c.rB) - c.vA - cross(c.wA, c.rA);
      const v2 impulse1 = -V(mx.x * Cdot1.x + my.x * Cdot1.y, mx.y * Cdot1.x + my.y * Cdot1.y);
      imp.x += impulse1.x; imp.y += impulse1.y;
      japply(c, impulse1, cross(c.rA, impulse1), cross(c.rB, impulse1));
    } else {
      const v2 Cdot1 = c.vB + cross(c.wB, c.rB) - c.vA - cross(c.wA, c.rA);
      const float Cdot2 = c.wB - c.wA;
      // -b2Mul(m_mass, Cdot) with b2Mul(A, v) = v.x * A.ex + v.y * A.ey + v.z * A.ez (b2math.d:679-682)
      const v3 mv = V3(Cdot1.x * mx.x + Cdot1.y * my.x + Cdot2 * mz.x, Cdot1.x * mx.y + Cdot1.y * my.y + Cdot2 * mz.y, Cdot1.x * mx.z + Cdot1.y * my.z + Cdot2 * mz.z);
      const v3 impulse = V3(-mv.x, -mv.y, -mv.z);
      imp.x += impulse.x; imp.y += impulse.y; imp.z += impulse.z;
      const v2 P = V(impulse.x, impulse.y);
      japply(c, P, cross(c.rA, P) + impulse.z, cross(c.rB, P) + impulse.z);
    }
  } else if (type == JT_FRICTION || type == JT_MOTOR) {   // b2frictionjoint.d:234-300, b2motorjoint.d:285-357
    const bool motor = type == JT_MOTOR;
    const float maxForce = motor ? p1.x : p0.x, maxTorque = motor ? p1.y : p0.y, corr = motor ? p0.w : 0.0f;
    {
      float Cdot = c.wB - c.wA;
      if (motor) Cdot = c.wB - c.wA + inv_h * corr * k1.w;
      float impulse = -k1.x * Cdot;
      const float oldImpulse = imp.z;
      const float maxImpulse = h * maxTorque;
      imp.z = fclampr(imp.z + impulse, -maxImpulse, maxImpulse);
      impulse = imp.z - oldImpulse;
      c.wA -= c.iA * impulse; c.wB += c.iB * impulse;
    }
    {
      v2 Cdot = c.vB + cross(c.wB, c.rB) - c.vA - cross(c.wA, c.rA);
      if (motor) Cdot = c.vB + cross(c.wB, c.rB) - c.vA - cross(c.wA, c.rA) + inv_h * corr * V(k1.y, k1.z);
      v2 impulse = -mul22(k0, Cdot);
      const v2 oldImpulse = V(imp.x, imp.y);
      v2 lin = oldImpulse + impulse;
      const float maxImpulse = h * maxForce;
      if (dot(lin, lin) > maxImpulse * maxImpulse) { normalize(lin); lin *= maxImpulse; }
      imp.x = lin.x; imp.y = lin.y;
      impulse = lin - oldImpulse;
      japply(c, impulse, cross(c.rA, impulse), cross(c.rB, impulse));
    }
  } else if (type == JT_MOUSE) {             // b2mousejoint.d:263-287
    const v2 Cdot = c.vB + cross(c.wB, c.rB);
    const v2 old = V(imp.x, imp.y);
    v2 impulse = mul22(k0, -(Cdot + V(k1.x, k1.y) + k1.z * old));
    v2 acc = old + impulse;
    const float maxImpulse = h * p0.z;
    if (dot(acc, acc) > maxImpulse * maxImpulse) acc *= maxImpulse / len(acc);
    imp.x = acc.x; imp.y = acc.y;
    impulse = acc - old;
    c.vB += c.mB * impulse; c.wB += c.iB * cross(c.rB, impulse);
  } else if (type == JT_PRISMATIC) {         // b2prismaticjoint.d:535-630
    const float4 k2 = W.j_k2[j], k3 = W.j_k3[j];
    const v2 axis = V(k0.x, k0.y), perp = V(k0.z, k0.w);
    const float s1 = k1.x, s2 = k1.y, a1 = k1.z, a2 = k1.w;
    const int limitState = W.j_limit[j];
    if (enableMotor && limitState != LIM_EQUAL) {
      const float Cdot = dot(axis, c.vB - c.vA) + a2 * c.wB - a1 * c.wA;
      float impulse = k3.z * (p1.x - Cdot);
      const float oldImpulse = imp.w;
      const float maxImpulse = h * p0.w;
      imp.w = fclampr(imp.w + impulse, -maxImpulse, maxImpulse);
      impulse = imp.w - oldImpulse;
      japply(c, impulse * axis, impulse * a1, impulse * a2);
    }
    v2 Cdot1;
    Cdot1.x = dot(perp, c.vB - c.vA) + s2 * c.wB - s1 * c.wA;
    Cdot1.y = c.wB - c.wA;
    const v3 ex = V3(k2.x, k2.y, k2.z), ey = V3(k2.y, k2.w, k3.x), ez = V3(k2.z, k3.x, k3.y);
    if (enableLimit && limitState != LIM_INACTIVE) {
      const float Cdot2 = dot(axis, c.vB - c.vA) + a2 * c.wB - a1 * c.wA;
      const v3 f1 = V3(imp.x, imp.y, imp.z);
      v3 df = solve33(ex, ey, ez, V3(-Cdot1.x, -Cdot1.y, -Cdot2));
      imp.x += df.x; imp.y += df.y; imp.z += df.z;
      if (limitState == LIM_LOWER) imp.z = fmaxr(imp.z, 0.0f);
      else if (limitState == LIM_UPPER) imp.z = fminr(imp.z, 0.0f);
      const v2 b = -Cdot1 - (imp.z - f1.z) * V(ez.x, ez.y);
      const v2 f2r = solve22(ex.x, ey.x, ex.y, ey.y, b) + V(f1.x, f1.y);
      imp.x = f2r.x; imp.y = f2r.y;
      df = V3(imp.x - f1.x, imp.y - f1.y, imp.z - f1.z);
      const v2 P = df.x * perp + df.z * axis;
      japply(c, P, df.x * s1 + df.y + df.z * a1, df.x * s2 + df.y + df.z * a2);
    } else {
      const v2 df = solve22(ex.x, ey.x, ex.y, ey.y, -Cdot1);
      imp.x += df.x; imp.y += df.y;
      japply(c, df.x * perp, df.x * s1 + df.y, df.x * s2 + df.y);
    }
  } else if (type == JT_WHEEL) {             // b2wheeljoint.d:445-507
    const float4 k2 = W.j_k2[j], k3 = W.j_k3[j];
    const v2 ax = V(k0.x, k0.y), ay = V(k0.z, k0.w);
    const float sAx = k1.x, sBx = k1.y, sAy = k1.z, sBy = k1.w;
    {
      const float Cdot = dot(ax, c.vB - c.vA) + sBx * c.wB - sAx * c.wA;
      const float impulse = -k2.z * (Cdot + k2.w + k3.x * imp.y);
      imp.y += impulse;
      japply(c, impulse * ax, impulse * sAx, impulse * sBx);
    }
    {
      const float Cdot = c.wB - c.wA - p0.w;
      float impulse = -k2.y * Cdot;
      const float oldImpulse = imp.w;
      const float maxImpulse = h * p0.z;
      imp.w = fclampr(imp.w + impulse, -maxImpulse, maxImpulse);
      impulse = imp.w - oldImpulse;
      c.wA -= c.iA * impulse; c.wB += c.iB * impulse;
    }
    {
      const float Cdot = dot(ay, c.vB - c.vA) + sBy * c.wB - sAy * c.wA;
      const float impulse = -k2.x * Cdot;
      imp.x += impulse;
      japply(c, impulse * ay, impulse * sAy, impulse * sBy);
    }
  } else if (type == JT_GEAR) {
    gear_solve_velocity(W, j, c);
  } else if (type == JT_PULLEY) {            // b2pulleyjoint.d:329-354
    const v2 uA = V(k0.x, k0.y), uB = V(k0.z, k0.w);
    const v2 vpA = c.vA + cross(c.wA, c.rA), vpB = c.vB + cross(c.wB, c.rB);
    const float Cdot = -dot(uA, vpA) - p1.z * dot(uB, vpB);
    const float impulse = -k1.x * Cdot;
    imp.x += impulse;
    const v2 PA = -impulse * uA, PB = -p1.z * impulse * uB;
    c.vA += c.mA * PA; c.wA += c.iA * cross(c.rA, PA);
    c.vB += c.mB * PB; c.wB += c.iB * cross(c.rB, PB);
  }
}

// ------------------------------------------------------------------------------------------------ SolvePositionConstraints
// positions come in and go out through cA/aA/cB/aB; rA/rB are the anchors rotated by the CURRENT angles
DBX_D bool joint2_solve_position(const DevWorld& W, int j, int type, int flags, float mA, float iA, float mB, float iB,
                                 v2 rA, v2 rB, Rot qA, v2& cA, float& aA, v2& cB, float& aB) {
  const float4 p0 = W.j_p0[j], p1 = W.j_p1[j];
  const bool enableLimit = (flags & 2) != 0;
  if (type == JT_ROPE) {                     // b2ropejoint.d:283-310
    v2 u = cB + rB - cA - rA;
    const float length = normalize(u);
    float C = length - p0.x;
    C = fclampr(C, 0.0f, kMaxLinearCorrection);
    const float impulse = -W.j_k0[j].w * C;
    const v2 P = impulse * u;
    cA -= mA * P; aA -= iA * cross(rA, P);
    cB += mB * P; aB += iB * cross(rB, P);
    return length - p0.x < kLinearSlop;
  }
  if (type == JT_WELD) {                     // b2weldjoint.d:355-448
    v3 ex, ey, ez;
    weld_k(mA, iA, mB, iB, rA, rB, ex, ey, ez);
    float positionError, angularError;
    if (p0.y > 0.0f) {
      const v2 C1 = cB + rB - cA - rA;
      positionError = len(C1); angularError = 0.0f;
      const v2 P = -solve22(ex.x, ey.x, ex.y, ey.y, C1);
      cA -= mA * P; aA -= iA * cross(rA, P);
      cB += mB * P; aB += iB * cross(rB, P);
    } else {
      const v2 C1 = cB + rB - cA - rA;
      const float C2 = aB - aA - p0.x;
      positionError = len(C1); angularError = fabsr(C2);
      v3 impulse;
      if (ez.z > 0.0f) { const v3 s = solve33(ex, ey, ez, V3(C1.x, C1.y, C2)); impulse = V3(-s.x, -s.y, -s.z); }
      else { const v2 s = solve22(ex.x, ey.x, ex.y, ey.y, C1); impulse = V3(-s.x, -s.y, 0.0f); }
      const v2 P = V(impulse.x, impulse.y);
      cA -= mA * P; aA -= iA * (cross(rA, P) + impulse.z);
      cB += mB * P; aB += iB * (cross(rB, P) + impulse.z);
    }
    return positionError <= kLinearSlop && angularError <= kAngularSlop;
  }
  if (type == JT_PRISMATIC) {                // b2prismaticjoint.d:632-745
    const v2 d = cB + rB - cA - rA;
    const v2 localX = V(p0.x, p0.y), localY = cross(1.0f, localX);
    const v2 axis = mul(qA, localX);
    const float a1 = cross(d + rA, axis), a2 = cross(rB, axis);
    const v2 perp = mul(qA, localY);
    const float s1 = cross(d + rA, perp), s2 = cross(rB, perp);
    v3 impulse;
    v2 C1;
    C1.x = dot(perp, d);
    C1.y = aB - aA - p0.z;
    float linearError = fabsr(C1.x);
    const float angularError = fabsr(C1.y);
    bool active = false;
    float C2 = 0.0f;
    if (enableLimit) {
      const float translation = dot(axis, d);
      if (fabsr(p1.z - p1.y) < 2.0f * kLinearSlop) {
        C2 = fclampr(translation, -kMaxLinearCorrection, kMaxLinearCorrection);
        linearError = fmaxr(linearError, fabsr(translation));
        active = true;
      } else if (translation <= p1.y) {
        C2 = fclampr(translation - p1.y + kLinearSlop, -kMaxLinearCorrection, 0.0f);
        linearError = fmaxr(linearError, p1.y - translation);
        active = true;
      } else if (translation >= p1.z) {
        C2 = fclampr(translation - p1.z - kLinearSlop, 0.0f, kMaxLinearCorrection);
        linearError = fmaxr(linearError, translation - p1.z);
        active = true;
      }
    }
    const float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
    const float k12 = iA * s1 + iB * s2;
    float k22 = iA + iB;
    if (k22 == 0.0f) k22 = 1.0f;
    if (active) {
      const float k13 = iA * s1 * a1 + iB * s2 * a2;
      const float k23 = iA * a1 + iB * a2;
      const float k33 = mA + mB + iA * a1 * a1 + iB * a2 * a2;
      impulse = solve33(V3(k11, k12, k13), V3(k12, k22, k23), V3(k13, k23, k33), V3(-C1.x, -C1.y, -C2));
    } else {
      const v2 impulse1 = solve22(k11, k12, k12, k22, -C1);
      impulse = V3(impulse1.x, impulse1.y, 0.0f);
    }
    const v2 P = impulse.x * perp + impulse.z * axis;
    const float LA = impulse.x * s1 + impulse.y + impulse.z * a1;
    const float LB = impulse.x * s2 + impulse.y + impulse.z * a2;
    cA -= mA * P; aA -= iA * LA;
    cB += mB * P; aB += iB * LB;
    return linearError <= kLinearSlop && angularError <= kAngularSlop;
  }
  if (type == JT_WHEEL) {                    // b2wheeljoint.d:509-560
    const v2 d = (cB - cA) + rB - rA;
    const v2 localY = cross(1.0f, V(p0.x, p0.y));
    const v2 ay = mul(qA, localY);
    const float sAy = cross(d + rA, ay), sBy = cross(rB, ay);
    const float C = dot(d, ay);
    const float4 k1 = W.j_k1[j];             // the reference divides by the velocity phase's m_sAy / m_sBy (:527)
    const float k = mA + mB + iA * k1.z * k1.z + iB * k1.w * k1.w;
    const float impulse = k != 0.0f ? -C / k : 0.0f;
    const v2 P = impulse * ay;
    cA -= mA * P; aA -= iA * (impulse * sAy);
    cB += mB * P; aB += iB * (impulse * sBy);
    return fabsr(C) <= kLinearSlop;
  }
  if (type == JT_PULLEY) {                   // b2pulleyjoint.d:356-440
    v2 uA = cA + rA - V(p0.x, p0.y), uB = cB + rB - V(p0.z, p0.w);
    const float lengthA = len(uA), lengthB = len(uB);
    if (lengthA > 10.0f * kLinearSlop) uA *= 1.0f / lengthA; else uA = V(0.0f, 0.0f);
    if (lengthB > 10.0f * kLinearSlop) uB *= 1.0f / lengthB; else uB = V(0.0f, 0.0f);
    const float ruA = cross(rA, uA), ruB = cross(rB, uB);
    const float ma = mA + iA * ruA * ruA, mb = mB + iB * ruB * ruB;
    float mass = ma + p1.z * p1.z * mb;
    if (mass > 0.0f) mass = 1.0f / mass;
    const float C = p1.w - lengthA - p1.z * lengthB;
    const float linearError = fabsr(C);
    const float impulse = -mass * C;
    const v2 PA = -impulse * uA, PB = -p1.z * impulse * uB;
    cA += mA * PA; aA += iA * cross(rA, PA);
    cB += mB * PB; aB += iB * cross(rB, PB);
    return linearError < kLinearSlop;
  }
  if (type == JT_GEAR) return gear_solve_position(W, j, mA, iA, mB, iB, rA, rB, cA, aA, cB, aB);
  return true;                               // friction, motor, mouse: no position correction
}

}  // namespace dbx
