// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_world.h).  PARITY UNPINNED.
// Second-wave joints of SURVEY.md 8(a) row a25, each hook restated from the reference file named next to it.
#include "orc_world.h"

namespace orc {

// the block every reference InitVelocityConstraints starts with (e.g. b2weldjoint.d:198-205)
void JointCommon::loadBodies() {
  indexA = bodyA->islandIndex; indexB = bodyB->islandIndex;
  localCenterA = bodyA->sweep.localCenter; localCenterB = bodyB->sweep.localCenter;
  invMassA = bodyA->invMass; invMassB = bodyB->invMass; invIA = bodyA->invI; invIB = bodyB->invI;
}
#define LOAD_POS() V2 cA = data.positions[indexA].c; float aA = data.positions[indexA].a; V2 cB = data.positions[indexB].c; float aB = data.positions[indexB].a
#define LOAD_VEL() V2 vA = data.velocities[indexA].v; float wA = data.velocities[indexA].w; V2 vB = data.velocities[indexB].v; float wB = data.velocities[indexB].w
#define STORE_VEL() data.velocities[indexA].v = vA; data.velocities[indexA].w = wA; data.velocities[indexB].v = vB; data.velocities[indexB].w = wB
#define STORE_POS() data.positions[indexA].c = cA; data.positions[indexA].a = aA; data.positions[indexB].c = cB; data.positions[indexB].a = aB

// ------------------------------------------------------------------ rope (b2ropejoint.d:165-310)
void RopeJoint::initVelocityConstraints(const SolverData& data) {
  loadBodies();
  LOAD_POS(); LOAD_VEL();
  Rot qA(aA), qB(aB);
  rA = mul(qA, localAnchorA - localCenterA);
  rB = mul(qB, localAnchorB - localCenterB);
  u = cB + rB - cA - rA;
  length = u.len();
  float C = length - maxLength;
  state = C > 0.0f ? kAtUpperLimit : kInactiveLimit;
  if (length > kLinearSlop) u *= 1.0f / length;
  else { u = V2(0, 0); mass = 0.0f; impulse = 0.0f; return; }
  float crA = cross(rA, u), crB = cross(rB, u);
  float invMass = invMassA + invIA * crA * crA + invMassB + invIB * crB * crB;
  mass = invMass != 0.0f ? 1.0f / invMass : 0.0f;
  if (data.step.warmStarting) {
    impulse *= data.step.dtRatio;
    V2 P = impulse * u;
    vA -= invMassA * P; wA -= invIA * cross(rA, P);
    vB += invMassB * P; wB += invIB * cross(rB, P);
  } else impulse = 0.0f;
  STORE_VEL();
}
void RopeJoint::solveVelocityConstraints(const SolverData& data) {
  LOAD_VEL();
  V2 vpA = vA + cross(wA, rA), vpB = vB + cross(wB, rB);
  float C = length - maxLength;
  float Cdot = dot(u, vpB - vpA);
  if (C < 0.0f) Cdot += data.step.inv_dt * C;   // predictive constraint
  float imp = -mass * Cdot;
  float oldImpulse = impulse;
  impulse = minT(0.0f, impulse + imp);
  imp = impulse - oldImpulse;
  V2 P = imp * u;
  vA -= invMassA * P; wA -= invIA * cross(rA, P);
  vB += invMassB * P; wB += invIB * cross(rB, P);
  STORE_VEL();
}
bool RopeJoint::solvePositionConstraints(const SolverData& data) {
  LOAD_POS();
  Rot qA(aA), qB(aB);
  V2 rA_ = mul(qA, localAnchorA - localCenterA), rB_ = mul(qB, localAnchorB - localCenterB);
  V2 u_ = cB + rB_ - cA - rA_;
  float len = u_.normalize();
  float C = len - maxLength;
  C = clampT(C, 0.0f, kMaxLinearCorrection);
  float imp = -mass * C;
  V2 P = imp * u_;
  cA -= invMassA * P; aA -= invIA * cross(rA_, P);
  cB += invMassB * P; aB += invIB * cross(rB_, P);
  STORE_POS();
  return len - maxLength < kLinearSlop;
}

// ------------------------------------------------------------------ weld (b2weldjoint.d:196-448)
static void weldK(M33& K, V2 rA, V2 rB, float mA, float mB, float iA, float iB) {
  K.ex.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
  K.ey.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
  K.ez.x = -rA.y * iA - rB.y * iB;
  K.ex.y = K.ey.x;
  K.ey.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
  K.ez.y = rA.x * iA + rB.x * iB;
  K.ex.z = K.ez.x;
  K.ey.z = K.ez.y;
  K.ez.z = iA + iB;
}
void WeldJoint::initVelocityConstraints(const SolverData& data) {
  loadBodies();
  float aA = data.positions[indexA].a, aB = data.positions[indexB].a;
  LOAD_VEL();
  Rot qA(aA), qB(aB);
  rA = mul(qA, localAnchorA - localCenterA);
  rB = mul(qB, localAnchorB - localCenterB);
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  M33 K; weldK(K, rA, rB, mA, mB, iA, iB);
  if (frequencyHz > 0.0f) {
    K.inverse22(&mass);
    float invM = iA + iB;
    float m = invM > 0.0f ? 1.0f / invM : 0.0f;
    float C = aB - aA - referenceAngle;
    float omega = 2.0f * kPi * frequencyHz;
    float d = 2.0f * m * dampingRatio * omega;
    float k = m * omega * omega;
    float h = data.step.dt;
    gamma = h * (d + h * k);
    gamma = gamma != 0.0f ? 1.0f / gamma : 0.0f;
    bias = C * h * k * gamma;
    invM += gamma;
    mass.ez.z = invM != 0.0f ? 1.0f / invM : 0.0f;
  } else if (K.ez.z == 0.0f) {
    K.inverse22(&mass); gamma = 0.0f; bias = 0.0f;
  } else {
    K.symInverse33(&mass); gamma = 0.0f; bias = 0.0f;
  }
  if (data.step.warmStarting) {
    impulse *= data.step.dtRatio;
    V2 P(impulse.x, impulse.y);
    vA -= mA * P; wA -= iA * (cross(rA, P) + impulse.z);
    vB += mB * P; wB += iB * (cross(rB, P) + impulse.z);
  } else impulse = V3(0, 0, 0);
  STORE_VEL();
}
void WeldJoint::solveVelocityConstraints(const SolverData& data) {
  LOAD_VEL();
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  if (frequencyHz > 0.0f) {
    float Cdot2 = wB - wA;
    float impulse2 = -mass.ez.z * (Cdot2 + bias + gamma * impulse.z);
    impulse.z += impulse2;
    wA -= iA * impulse2; wB += iB * impulse2;
    V2 Cdot1 = vB + cross(wB, rB) - vA - cross(wA, rA);
    V2 impulse1 = -mul22(mass, Cdot1);
    impulse.x += impulse1.x; impulse.y += impulse1.y;
    V2 P = impulse1;
    vA -= mA * P; wA -= iA * cross(rA, P);
    vB += mB * P; wB += iB * cross(rB, P);
  } else {
    V2 Cdot1 = vB + cross(wB, rB) - vA - cross(wA, rA);
    float Cdot2 = wB - wA;
    V3 Cdot(Cdot1.x, Cdot1.y, Cdot2);
    V3 imp = -mul(mass, Cdot);
    impulse += imp;
    V2 P(imp.x, imp.y);
    vA -= mA * P; wA -= iA * (cross(rA, P) + imp.z);
    vB += mB * P; wB += iB * (cross(rB, P) + imp.z);
  }
  STORE_VEL();
}
bool WeldJoint::solvePositionConstraints(const SolverData& data) {
  LOAD_POS();
  Rot qA(aA), qB(aB);
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  V2 rA_ = mul(qA, localAnchorA - localCenterA), rB_ = mul(qB, localAnchorB - localCenterB);
  float positionError, angularError;
  M33 K; weldK(K, rA_, rB_, mA, mB, iA, iB);
  if (frequencyHz > 0.0f) {
    V2 C1 = cB + rB_ - cA - rA_;
    positionError = C1.len(); angularError = 0.0f;
    V2 P = -K.solve22(C1);
    cA -= mA * P; aA -= iA * cross(rA_, P);
    cB += mB * P; aB += iB * cross(rB_, P);
  } else {
    V2 C1 = cB + rB_ - cA - rA_;
    float C2 = aB - aA - referenceAngle;
    positionError = C1.len(); angularError = absT(C2);
    V3 C(C1.x, C1.y, C2);
    V3 imp;
    if (K.ez.z > 0.0f) imp = -K.solve33(C);
    else { V2 i2 = -K.solve22(C1); imp = V3(i2.x, i2.y, 0.0f); }
    V2 P(imp.x, imp.y);
    cA -= mA * P; aA -= iA * (cross(rA_, P) + imp.z);
    cB += mB * P; aB += iB * (cross(rB_, P) + imp.z);
  }
  STORE_POS();
  return positionError <= kLinearSlop && angularError <= kAngularSlop;
}

// ------------------------------------------------------------------ friction (b2frictionjoint.d:178-318)
static M22 pointK(V2 rA, V2 rB, float mA, float mB, float iA, float iB) {
  M22 K;
  K.ex.x = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
  K.ex.y = -iA * rA.x * rA.y - iB * rB.x * rB.y;
  K.ey.x = K.ex.y;
  K.ey.y = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
  return K;
}
void FrictionJoint::initVelocityConstraints(const SolverData& data) {
  loadBodies();
  float aA = data.positions[indexA].a, aB = data.positions[indexB].a;
  LOAD_VEL();
  Rot qA(aA), qB(aB);
  rA = mul(qA, localAnchorA - localCenterA);
  rB = mul(qB, localAnchorB - localCenterB);
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  linearMass = pointK(rA, rB, mA, mB, iA, iB).inverse();
  angularMass = iA + iB;
  if (angularMass > 0.0f) angularMass = 1.0f / angularMass;
  if (data.step.warmStarting) {
    linearImpulse *= data.step.dtRatio; angularImpulse *= data.step.dtRatio;
    V2 P(linearImpulse.x, linearImpulse.y);
    vA -= mA * P; wA -= iA * (cross(rA, P) + angularImpulse);
    vB += mB * P; wB += iB * (cross(rB, P) + angularImpulse);
  } else { linearImpulse = V2(0, 0); angularImpulse = 0.0f; }
  STORE_VEL();
}
void FrictionJoint::solveVelocityConstraints(const SolverData& data) {
  LOAD_VEL();
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  float h = data.step.dt;
  {
    float Cdot = wB - wA;
    float imp = -angularMass * Cdot;
    float oldImpulse = angularImpulse;
    float maxImpulse = h * maxTorque;
    angularImpulse = clampT(angularImpulse + imp, -maxImpulse, maxImpulse);
    imp = angularImpulse - oldImpulse;
    wA -= iA * imp; wB += iB * imp;
  }
  {
    V2 Cdot = vB + cross(wB, rB) - vA - cross(wA, rA);
    V2 imp = -mul(linearMass, Cdot);
    V2 oldImpulse = linearImpulse;
    linearImpulse += imp;
    float maxImpulse = h * maxForce;
    if (linearImpulse.len2() > maxImpulse * maxImpulse) { linearImpulse.normalize(); linearImpulse *= maxImpulse; }
    imp = linearImpulse - oldImpulse;
    vA -= mA * imp; wA -= iA * cross(rA, imp);
    vB += mB * imp; wB += iB * cross(rB, imp);
  }
  STORE_VEL();
}
bool FrictionJoint::solvePositionConstraints(const SolverData&) { return true; }

// ------------------------------------------------------------------ motor (b2motorjoint.d:223-385)
void MotorJoint::initVelocityConstraints(const SolverData& data) {
  loadBodies();
  LOAD_POS(); LOAD_VEL();
  Rot qA(aA), qB(aB);
  rA = mul(qA, -localCenterA);
  rB = mul(qB, -localCenterB);
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  linearMass = pointK(rA, rB, mA, mB, iA, iB).inverse();
  angularMass = iA + iB;
  if (angularMass > 0.0f) angularMass = 1.0f / angularMass;
  linearError = cB + rB - cA - rA - mul(qA, linearOffset);
  angularError = aB - aA - angularOffset;
  if (data.step.warmStarting) {
    linearImpulse *= data.step.dtRatio; angularImpulse *= data.step.dtRatio;
    V2 P(linearImpulse.x, linearImpulse.y);
    vA -= mA * P; wA -= iA * (cross(rA, P) + angularImpulse);
    vB += mB * P; wB += iB * (cross(rB, P) + angularImpulse);
  } else { linearImpulse = V2(0, 0); angularImpulse = 0.0f; }
  STORE_VEL();
}
void MotorJoint::solveVelocityConstraints(const SolverData& data) {
  LOAD_VEL();
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  float h = data.step.dt, inv_h = data.step.inv_dt;
  {
    float Cdot = wB - wA + inv_h * correctionFactor * angularError;
    float imp = -angularMass * Cdot;
    float oldImpulse = angularImpulse;
    float maxImpulse = h * maxTorque;
    angularImpulse = clampT(angularImpulse + imp, -maxImpulse, maxImpulse);
    imp = angularImpulse - oldImpulse;
    wA -= iA * imp; wB += iB * imp;
  }
  {
    V2 Cdot = vB + cross(wB, rB) - vA - cross(wA, rA) + inv_h * correctionFactor * linearError;
    V2 imp = -mul(linearMass, Cdot);
    V2 oldImpulse = linearImpulse;
    linearImpulse += imp;
    float maxImpulse = h * maxForce;
    if (linearImpulse.len2() > maxImpulse * maxImpulse) { linearImpulse.normalize(); linearImpulse *= maxImpulse; }
    imp = linearImpulse - oldImpulse;
    vA -= mA * imp; wA -= iA * cross(rA, imp);
    vB += mB * imp; wB += iB * cross(rB, imp);
  }
  STORE_VEL();
}
bool MotorJoint::solvePositionConstraints(const SolverData&) { return true; }

// ------------------------------------------------------------------ mouse (b2mousejoint.d:190-300)
void MouseJoint::initVelocityConstraints(const SolverData& data) {
  indexB = bodyB->islandIndex; localCenterB = bodyB->sweep.localCenter; invMassB = bodyB->invMass; invIB = bodyB->invI;
  V2 cB = data.positions[indexB].c; float aB = data.positions[indexB].a;
  V2 vB = data.velocities[indexB].v; float wB = data.velocities[indexB].w;
  Rot qB(aB);
  float m = bodyB->mass;
  float omega = 2.0f * kPi * frequencyHz;
  float d = 2.0f * m * dampingRatio * omega;
  float k = m * (omega * omega);
  float h = data.step.dt;
  gamma = h * (d + h * k);
  if (gamma != 0.0f) gamma = 1.0f / gamma;
  beta = h * k * gamma;
  rB = mul(qB, localAnchorB - localCenterB);
  M22 K;
  K.ex.x = invMassB + invIB * rB.y * rB.y + gamma;
  K.ex.y = -invIB * rB.x * rB.y;
  K.ey.x = K.ex.y;
  K.ey.y = invMassB + invIB * rB.x * rB.x + gamma;
  mass = K.inverse();
  C = cB + rB - targetA;
  C *= beta;
  wB *= 0.98f;   // cheat with some damping
  if (data.step.warmStarting) {
    impulse *= data.step.dtRatio;
    vB += invMassB * impulse;
    wB += invIB * cross(rB, impulse);
  } else impulse = V2(0, 0);
  data.velocities[indexB].v = vB; data.velocities[indexB].w = wB;
}
void MouseJoint::solveVelocityConstraints(const SolverData& data) {
  V2 vB = data.velocities[indexB].v; float wB = data.velocities[indexB].w;
  V2 Cdot = vB + cross(wB, rB);
  V2 imp = mul(mass, -(Cdot + C + gamma * impulse));
  V2 oldImpulse = impulse;
  impulse += imp;
  float maxImpulse = data.step.dt * maxForce;
  if (impulse.len2() > maxImpulse * maxImpulse) impulse *= maxImpulse / impulse.len();
  imp = impulse - oldImpulse;
  vB += invMassB * imp;
  wB += invIB * cross(rB, imp);
  data.velocities[indexB].v = vB; data.velocities[indexB].w = wB;
}
bool MouseJoint::solvePositionConstraints(const SolverData&) { return true; }

// ------------------------------------------------------------------ prismatic (b2prismaticjoint.d:391-745)
void PrismaticJoint::initVelocityConstraints(const SolverData& data) {
  loadBodies();
  LOAD_POS(); LOAD_VEL();
  Rot qA(aA), qB(aB);
  V2 rA_ = mul(qA, localAnchorA - localCenterA), rB_ = mul(qB, localAnchorB - localCenterB);
  V2 d = (cB - cA) + rB_ - rA_;
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  {
    axis = mul(qA, localXAxisA);
    a1 = cross(d + rA_, axis);
    a2 = cross(rB_, axis);
    motorMass = mA + mB + iA * a1 * a1 + iB * a2 * a2;
    if (motorMass > 0.0f) motorMass = 1.0f / motorMass;
  }
  {
    perp = mul(qA, localYAxisA);
    s1 = cross(d + rA_, perp);
    s2 = cross(rB_, perp);
    float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
    float k12 = iA * s1 + iB * s2;
    float k13 = iA * s1 * a1 + iB * s2 * a2;
    float k22 = iA + iB;
    if (k22 == 0.0f) k22 = 1.0f;   // bodies with fixed rotation
    float k23 = iA * a1 + iB * a2;
    float k33 = mA + mB + iA * a1 * a1 + iB * a2 * a2;
    K.ex = V3(k11, k12, k13); K.ey = V3(k12, k22, k23); K.ez = V3(k13, k23, k33);
  }
  if (enableLimit) {
    float jointTranslation = dot(axis, d);
    if (absT(upperTranslation - lowerTranslation) < 2.0f * kLinearSlop) limitState = kEqualLimits;
    else if (jointTranslation <= lowerTranslation) { if (limitState != kAtLowerLimit) { limitState = kAtLowerLimit; impulse.z = 0.0f; } }
    else if (jointTranslation >= upperTranslation) { if (limitState != kAtUpperLimit) { limitState = kAtUpperLimit; impulse.z = 0.0f; } }
    else { limitState = kInactiveLimit; impulse.z = 0.0f; }
  } else { limitState = kInactiveLimit; impulse.z = 0.0f; }
  if (enableMotor == false) motorImpulse = 0.0f;
  if (data.step.warmStarting) {
    impulse *= data.step.dtRatio; motorImpulse *= data.step.dtRatio;
    V2 P = impulse.x * perp + (motorImpulse + impulse.z) * axis;
    float LA = impulse.x * s1 + impulse.y + (motorImpulse + impulse.z) * a1;
    float LB = impulse.x * s2 + impulse.y + (motorImpulse + impulse.z) * a2;
    vA -= mA * P; wA -= iA * LA;
    vB += mB * P; wB += iB * LB;
  } else { impulse = V3(0, 0, 0); motorImpulse = 0.0f; }
  STORE_VEL();
}
void PrismaticJoint::solveVelocityConstraints(const SolverData& data) {
  LOAD_VEL();
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  if (enableMotor && limitState != kEqualLimits) {
    float Cdot = dot(axis, vB - vA) + a2 * wB - a1 * wA;
    float imp = motorMass * (motorSpeed - Cdot);
    float oldImpulse = motorImpulse;
    float maxImpulse = data.step.dt * maxMotorForce;
    motorImpulse = clampT(motorImpulse + imp, -maxImpulse, maxImpulse);
    imp = motorImpulse - oldImpulse;
    V2 P = imp * axis;
    float LA = imp * a1, LB = imp * a2;
    vA -= mA * P; wA -= iA * LA;
    vB += mB * P; wB += iB * LB;
  }
  V2 Cdot1;
  Cdot1.x = dot(perp, vB - vA) + s2 * wB - s1 * wA;
  Cdot1.y = wB - wA;
  if (enableLimit && limitState != kInactiveLimit) {
    float Cdot2 = dot(axis, vB - vA) + a2 * wB - a1 * wA;
    V3 Cdot(Cdot1.x, Cdot1.y, Cdot2);
    V3 f1 = impulse;
    V3 df = K.solve33(-Cdot);
    impulse += df;
    if (limitState == kAtLowerLimit) impulse.z = maxT(impulse.z, 0.0f);
    else if (limitState == kAtUpperLimit) impulse.z = minT(impulse.z, 0.0f);
    V2 b = -Cdot1 - (impulse.z - f1.z) * V2(K.ez.x, K.ez.y);
    V2 f2r = K.solve22(b) + V2(f1.x, f1.y);
    impulse.x = f2r.x; impulse.y = f2r.y;
    df = impulse - f1;
    V2 P = df.x * perp + df.z * axis;
    float LA = df.x * s1 + df.y + df.z * a1;
    float LB = df.x * s2 + df.y + df.z * a2;
    vA -= mA * P; wA -= iA * LA;
    vB += mB * P; wB += iB * LB;
  } else {
    V2 df = K.solve22(-Cdot1);
    impulse.x += df.x; impulse.y += df.y;
    V2 P = df.x * perp;
    float LA = df.x * s1 + df.y, LB = df.x * s2 + df.y;
    vA -= mA * P; wA -= iA * LA;
    vB += mB * P; wB += iB * LB;
  }
  STORE_VEL();
}
bool PrismaticJoint::solvePositionConstraints(const SolverData& data) {
  LOAD_POS();
  Rot qA(aA), qB(aB);
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  V2 rA_ = mul(qA, localAnchorA - localCenterA), rB_ = mul(qB, localAnchorB - localCenterB);
  V2 d = cB + rB_ - cA - rA_;
  V2 ax = mul(qA, localXAxisA);
  float a1_ = cross(d + rA_, ax), a2_ = cross(rB_, ax);
  V2 pp = mul(qA, localYAxisA);
  float s1_ = cross(d + rA_, pp), s2_ = cross(rB_, pp);
  V3 imp;
  V2 C1;
  C1.x = dot(pp, d);
  C1.y = aB - aA - referenceAngle;
  float linearError = absT(C1.x), angularError = absT(C1.y);
  bool active = false;
  float C2 = 0.0f;
  if (enableLimit) {
    float translation = dot(ax, d);
    if (absT(upperTranslation - lowerTranslation) < 2.0f * kLinearSlop) {
      C2 = clampT(translation, -kMaxLinearCorrection, kMaxLinearCorrection);
      linearError = maxT(linearError, absT(translation));
      active = true;
    } else if (translation <= lowerTranslation) {
      C2 = clampT(translation - lowerTranslation + kLinearSlop, -kMaxLinearCorrection, 0.0f);
      linearError = maxT(linearError, lowerTranslation - translation);
      active = true;
    } else if (translation >= upperTranslation) {
      C2 = clampT(translation - upperTranslation - kLinearSlop, 0.0f, kMaxLinearCorrection);
      linearError = maxT(linearError, translation - upperTranslation);
      active = true;
    }
  }
  if (active) {
    float k11 = mA + mB + iA * s1_ * s1_ + iB * s2_ * s2_;
    float k12 = iA * s1_ + iB * s2_;
    float k13 = iA * s1_ * a1_ + iB * s2_ * a2_;
    float k22 = iA + iB;
    if (k22 == 0.0f) k22 = 1.0f;
    float k23 = iA * a1_ + iB * a2_;
    float k33 = mA + mB + iA * a1_ * a1_ + iB * a2_ * a2_;
    M33 Km; Km.ex = V3(k11, k12, k13); Km.ey = V3(k12, k22, k23); Km.ez = V3(k13, k23, k33);
    V3 C(C1.x, C1.y, C2);
    imp = Km.solve33(-C);
  } else {
    float k11 = mA + mB + iA * s1_ * s1_ + iB * s2_ * s2_;
    float k12 = iA * s1_ + iB * s2_;
    float k22 = iA + iB;
    if (k22 == 0.0f) k22 = 1.0f;
    M22 Km(V2(k11, k12), V2(k12, k22));
    V2 impulse1 = Km.solve(-C1);
    imp = V3(impulse1.x, impulse1.y, 0.0f);
  }
  V2 P = imp.x * pp + imp.z * ax;
  float LA = imp.x * s1_ + imp.y + imp.z * a1_;
  float LB = imp.x * s2_ + imp.y + imp.z * a2_;
  cA -= mA * P; aA -= iA * LA;
  cB += mB * P; aB += iB * LB;
  STORE_POS();
  return linearError <= kLinearSlop && angularError <= kAngularSlop;
}

// ------------------------------------------------------------------ wheel (b2wheeljoint.d:302-560)
void WheelJoint::initVelocityConstraints(const SolverData& data) {
  loadBodies();
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  LOAD_POS(); LOAD_VEL();
  Rot qA(aA), qB(aB);
  V2 rA_ = mul(qA, localAnchorA - localCenterA), rB_ = mul(qB, localAnchorB - localCenterB);
  V2 d = cB + rB_ - cA - rA_;
  {
    ay = mul(qA, localYAxisA);
    sAy = cross(d + rA_, ay);
    sBy = cross(rB_, ay);
    mass = mA + mB + iA * sAy * sAy + iB * sBy * sBy;
    if (mass > 0.0f) mass = 1.0f / mass;
  }
  springMass = 0.0f; bias = 0.0f; gamma = 0.0f;
  if (frequencyHz > 0.0f) {
    ax = mul(qA, localXAxisA);
    sAx = cross(d + rA_, ax);
    sBx = cross(rB_, ax);
    float invMass = mA + mB + iA * sAx * sAx + iB * sBx * sBx;
    if (invMass > 0.0f) {
      springMass = 1.0f / invMass;
      float C = dot(d, ax);
      float omega = 2.0f * kPi * frequencyHz;
      float dd = 2.0f * springMass * dampingRatio * omega;
      float k = springMass * omega * omega;
      float h = data.step.dt;
      gamma = h * (dd + h * k);
      if (gamma > 0.0f) gamma = 1.0f / gamma;
      bias = C * h * k * gamma;
      springMass = invMass + gamma;
      if (springMass > 0.0f) springMass = 1.0f / springMass;
    }
  } else springImpulse = 0.0f;
  if (enableMotor) {
    motorMass = iA + iB;
    if (motorMass > 0.0f) motorMass = 1.0f / motorMass;
  } else { motorMass = 0.0f; motorImpulse = 0.0f; }
  if (data.step.warmStarting) {
    impulse *= data.step.dtRatio; springImpulse *= data.step.dtRatio; motorImpulse *= data.step.dtRatio;
    V2 P = impulse * ay + springImpulse * ax;
    float LA = impulse * sAy + springImpulse * sAx + motorImpulse;
    float LB = impulse * sBy + springImpulse * sBx + motorImpulse;
    vA -= invMassA * P; wA -= invIA * LA;
    vB += invMassB * P; wB += invIB * LB;
  } else { impulse = 0.0f; springImpulse = 0.0f; motorImpulse = 0.0f; }
  STORE_VEL();
}
void WheelJoint::solveVelocityConstraints(const SolverData& data) {
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  LOAD_VEL();
  {
    float Cdot = dot(ax, vB - vA) + sBx * wB - sAx * wA;
    float imp = -springMass * (Cdot + bias + gamma * springImpulse);
    springImpulse += imp;
    V2 P = imp * ax;
    float LA = imp * sAx, LB = imp * sBx;
    vA -= mA * P; wA -= iA * LA;
    vB += mB * P; wB += iB * LB;
  }
  {
    float Cdot = wB - wA - motorSpeed;
    float imp = -motorMass * Cdot;
    float oldImpulse = motorImpulse;
    float maxImpulse = data.step.dt * maxMotorTorque;
    motorImpulse = clampT(motorImpulse + imp, -maxImpulse, maxImpulse);
    imp = motorImpulse - oldImpulse;
    wA -= iA * imp; wB += iB * imp;
  }
  {
    float Cdot = dot(ay, vB - vA) + sBy * wB - sAy * wA;
    float imp = -mass * Cdot;
    impulse += imp;
    V2 P = imp * ay;
    float LA = imp * sAy, LB = imp * sBy;
    vA -= mA * P; wA -= iA * LA;
    vB += mB * P; wB += iB * LB;
  }
  STORE_VEL();
}
bool WheelJoint::solvePositionConstraints(const SolverData& data) {
  LOAD_POS();
  Rot qA(aA), qB(aB);
  V2 rA_ = mul(qA, localAnchorA - localCenterA), rB_ = mul(qB, localAnchorB - localCenterB);
  V2 d = (cB - cA) + rB_ - rA_;
  V2 ay_ = mul(qA, localYAxisA);
  float sAy_ = cross(d + rA_, ay_), sBy_ = cross(rB_, ay_);
  float C = dot(d, ay_);
  // the reference uses the velocity phase's m_sAy / m_sBy here (b2wheeljoint.d:527), not the fresh values
  float k = invMassA + invMassB + invIA * sAy * sAy + invIB * sBy * sBy;
  float imp = k != 0.0f ? -C / k : 0.0f;
  V2 P = imp * ay_;
  float LA = imp * sAy_, LB = imp * sBy_;
  cA -= invMassA * P; aA -= invIA * LA;
  cB += invMassB * P; aB += invIB * LB;
  STORE_POS();
  return absT(C) <= kLinearSlop;
}

// ------------------------------------------------------------------ pulley (b2pulleyjoint.d:238-440)
void PulleyJoint::initVelocityConstraints(const SolverData& data) {
  loadBodies();
  LOAD_POS(); LOAD_VEL();
  Rot qA(aA), qB(aB);
  rA = mul(qA, localAnchorA - localCenterA);
  rB = mul(qB, localAnchorB - localCenterB);
  uA = cA + rA - groundAnchorA;
  uB = cB + rB - groundAnchorB;
  float lenA = uA.len(), lenB = uB.len();
  if (lenA > 10.0f * kLinearSlop) uA *= 1.0f / lenA; else uA = V2(0, 0);
  if (lenB > 10.0f * kLinearSlop) uB *= 1.0f / lenB; else uB = V2(0, 0);
  float ruA = cross(rA, uA), ruB = cross(rB, uB);
  float mA = invMassA + invIA * ruA * ruA;
  float mB = invMassB + invIB * ruB * ruB;
  mass = mA + ratio * ratio * mB;
  if (mass > 0.0f) mass = 1.0f / mass;
  if (data.step.warmStarting) {
    impulse *= data.step.dtRatio;
    V2 PA = -(impulse) * uA;
    V2 PB = (-ratio * impulse) * uB;
    vA += invMassA * PA; wA += invIA * cross(rA, PA);
    vB += invMassB * PB; wB += invIB * cross(rB, PB);
  } else impulse = 0.0f;
  STORE_VEL();
}
void PulleyJoint::solveVelocityConstraints(const SolverData& data) {
  LOAD_VEL();
  V2 vpA = vA + cross(wA, rA), vpB = vB + cross(wB, rB);
  float Cdot = -dot(uA, vpA) - ratio * dot(uB, vpB);
  float imp = -mass * Cdot;
  impulse += imp;
  V2 PA = -imp * uA;
  V2 PB = -ratio * imp * uB;
  vA += invMassA * PA; wA += invIA * cross(rA, PA);
  vB += invMassB * PB; wB += invIB * cross(rB, PB);
  STORE_VEL();
}
bool PulleyJoint::solvePositionConstraints(const SolverData& data) {
  LOAD_POS();
  Rot qA(aA), qB(aB);
  V2 rA_ = mul(qA, localAnchorA - localCenterA), rB_ = mul(qB, localAnchorB - localCenterB);
  V2 uA_ = cA + rA_ - groundAnchorA, uB_ = cB + rB_ - groundAnchorB;
  float lenA = uA_.len(), lenB = uB_.len();
  if (lenA > 10.0f * kLinearSlop) uA_ *= 1.0f / lenA; else uA_ = V2(0, 0);
  if (lenB > 10.0f * kLinearSlop) uB_ *= 1.0f / lenB; else uB_ = V2(0, 0);
  float ruA = cross(rA_, uA_), ruB = cross(rB_, uB_);
  float mA = invMassA + invIA * ruA * ruA;
  float mB = invMassB + invIB * ruB * ruB;
  float m = mA + ratio * ratio * mB;
  if (m > 0.0f) m = 1.0f / m;
  float C = constant - lenA - ratio * lenB;
  float linearError = absT(C);
  float imp = -m * C;
  V2 PA = -imp * uA_;
  V2 PB = -ratio * imp * uB_;
  cA += invMassA * PA; aA += invIA * cross(rA_, PA);
  cB += invMassB * PB; aB += invIB * cross(rB_, PB);
  STORE_POS();
  return linearError < kLinearSlop;
}

// ------------------------------------------------------------------ gear (b2gearjoint.d:245-500)
void GearJoint::initVelocityConstraints(const SolverData& data) {
  indexA = bodyA->islandIndex; indexB = bodyB->islandIndex; indexC = bodyC->islandIndex; indexD = bodyD->islandIndex;
  lcA = bodyA->sweep.localCenter; lcB = bodyB->sweep.localCenter; lcC = bodyC->sweep.localCenter; lcD = bodyD->sweep.localCenter;
  mA = bodyA->invMass; mB = bodyB->invMass; mC = bodyC->invMass; mD = bodyD->invMass;
  iA = bodyA->invI; iB = bodyB->invI; iC = bodyC->invI; iD = bodyD->invI;
  float aA = data.positions[indexA].a; V2 vA = data.velocities[indexA].v; float wA = data.velocities[indexA].w;
  float aB = data.positions[indexB].a; V2 vB = data.velocities[indexB].v; float wB = data.velocities[indexB].w;
  float aC = data.positions[indexC].a; V2 vC = data.velocities[indexC].v; float wC = data.velocities[indexC].w;
  float aD = data.positions[indexD].a; V2 vD = data.velocities[indexD].v; float wD = data.velocities[indexD].w;
  Rot qA(aA), qB(aB), qC(aC), qD(aD);
  mass = 0.0f;
  if (typeA == jRevolute) {
    JvAC = V2(0, 0); JwA = 1.0f; JwC = 1.0f;
    mass += iA + iC;
  } else {
    V2 u = mul(qC, localAxisC);
    V2 rC = mul(qC, localAnchorC - lcC);
    V2 rA = mul(qA, localAnchorA - lcA);
    JvAC = u; JwC = cross(rC, u); JwA = cross(rA, u);
    mass += mC + mA + iC * JwC * JwC + iA * JwA * JwA;
  }
  if (typeB == jRevolute) {
    JvBD = V2(0, 0); JwB = ratio; JwD = ratio;
    mass += ratio * ratio * (iB + iD);
  } else {
    V2 u = mul(qD, localAxisD);
    V2 rD = mul(qD, localAnchorD - lcD);
    V2 rB = mul(qB, localAnchorB - lcB);
    JvBD = ratio * u; JwD = ratio * cross(rD, u); JwB = ratio * cross(rB, u);
    mass += ratio * ratio * (mD + mB) + iD * JwD * JwD + iB * JwB * JwB;
  }
  mass = mass > 0.0f ? 1.0f / mass : 0.0f;
  if (data.step.warmStarting) {
    vA += (mA * impulse) * JvAC; wA += iA * impulse * JwA;
    vB += (mB * impulse) * JvBD; wB += iB * impulse * JwB;
    vC -= (mC * impulse) * JvAC; wC -= iC * impulse * JwC;
    vD -= (mD * impulse) * JvBD; wD -= iD * impulse * JwD;
  } else impulse = 0.0f;
  data.velocities[indexA].v = vA; data.velocities[indexA].w = wA; data.velocities[indexB].v = vB; data.velocities[indexB].w = wB;
  data.velocities[indexC].v = vC; data.velocities[indexC].w = wC; data.velocities[indexD].v = vD; data.velocities[indexD].w = wD;
}
void GearJoint::solveVelocityConstraints(const SolverData& data) {
  V2 vA = data.velocities[indexA].v; float wA = data.velocities[indexA].w;
  V2 vB = data.velocities[indexB].v; float wB = data.velocities[indexB].w;
  V2 vC = data.velocities[indexC].v; float wC = data.velocities[indexC].w;
  V2 vD = data.velocities[indexD].v; float wD = data.velocities[indexD].w;
  float Cdot = dot(JvAC, vA - vC) + dot(JvBD, vB - vD);
  Cdot += (JwA * wA - JwC * wC) + (JwB * wB - JwD * wD);
  float imp = -mass * Cdot;
  impulse += imp;
  vA += (mA * imp) * JvAC; wA += iA * imp * JwA;
  vB += (mB * imp) * JvBD; wB += iB * imp * JwB;
  vC -= (mC * imp) * JvAC; wC -= iC * imp * JwC;
  vD -= (mD * imp) * JvBD; wD -= iD * imp * JwD;
  data.velocities[indexA].v = vA; data.velocities[indexA].w = wA; data.velocities[indexB].v = vB; data.velocities[indexB].w = wB;
  data.velocities[indexC].v = vC; data.velocities[indexC].w = wC; data.velocities[indexD].v = vD; data.velocities[indexD].w = wD;
}
bool GearJoint::solvePositionConstraints(const SolverData& data) {
  V2 cA = data.positions[indexA].c; float aA = data.positions[indexA].a;
  V2 cB = data.positions[indexB].c; float aB = data.positions[indexB].a;
  V2 cC = data.positions[indexC].c; float aC = data.positions[indexC].a;
  V2 cD = data.positions[indexD].c; float aD = data.positions[indexD].a;
  Rot qA(aA), qB(aB), qC(aC), qD(aD);
  float linearError = 0.0f;
  float coordinateA, coordinateB;
  V2 JvAC_, JvBD_; float JwA_, JwB_, JwC_, JwD_;
  float m = 0.0f;
  if (typeA == jRevolute) {
    JvAC_ = V2(0, 0); JwA_ = 1.0f; JwC_ = 1.0f;
    m += iA + iC;
    coordinateA = aA - aC - referenceAngleA;
  } else {
    V2 u = mul(qC, localAxisC);
    V2 rC = mul(qC, localAnchorC - lcC);
    V2 rA = mul(qA, localAnchorA - lcA);
    JvAC_ = u; JwC_ = cross(rC, u); JwA_ = cross(rA, u);
    m += mC + mA + iC * JwC_ * JwC_ + iA * JwA_ * JwA_;
    V2 pC = localAnchorC - lcC;
    V2 pA = mulT(qC, rA + (cA - cC));
    coordinateA = dot(pA - pC, localAxisC);
  }
  if (typeB == jRevolute) {
    JvBD_ = V2(0, 0); JwB_ = ratio; JwD_ = ratio;
    m += ratio * ratio * (iB + iD);
    coordinateB = aB - aD - referenceAngleB;
  } else {
    V2 u = mul(qD, localAxisD);
    V2 rD = mul(qD, localAnchorD - lcD);
    V2 rB = mul(qB, localAnchorB - lcB);
    JvBD_ = ratio * u; JwD_ = ratio * cross(rD, u); JwB_ = ratio * cross(rB, u);
    m += ratio * ratio * (mD + mB) + iD * JwD_ * JwD_ + iB * JwB_ * JwB_;
    V2 pD = localAnchorD - lcD;
    V2 pB = mulT(qD, rB + (cB - cD));
    coordinateB = dot(pB - pD, localAxisD);
  }
  float C = (coordinateA + ratio * coordinateB) - constant;
  float imp = 0.0f;
  if (m > 0.0f) imp = -C / m;
  cA += mA * imp * JvAC_; aA += iA * imp * JwA_;
  cB += mB * imp * JvBD_; aB += iB * imp * JwB_;
  cC -= mC * imp * JvAC_; aC -= iC * imp * JwC_;
  cD -= mD * imp * JvBD_; aD -= iD * imp * JwD_;
  data.positions[indexA].c = cA; data.positions[indexA].a = aA; data.positions[indexB].c = cB; data.positions[indexB].a = aB;
  data.positions[indexC].c = cC; data.positions[indexC].a = aC; data.positions[indexD].c = cD; data.positions[indexD].a = aD;
  return linearError < kLinearSlop;
}

}  // namespace orc
