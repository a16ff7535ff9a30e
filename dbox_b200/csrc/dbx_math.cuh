// dbx_math.cuh — fp32 2-D math for the device (and the host side of the shim).
// Semantics follow the reference's math layer (src/dbox/common/b2math.d:599-828, b2settings.d:61-149): every
// expression keeps the reference's evaluation order and the library is compiled with --fmad=false so that
// narrowphase results are bit-identical to the reference arithmetic for identical inputs.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

#define DBX_HD __host__ __device__ __forceinline__
#define DBX_D __device__ __forceinline__

namespace dbx {

// b2settings.d:61-149
constexpr float kMaxFloat = FLT_MAX;
constexpr float kEpsilon = FLT_EPSILON;
constexpr float kPi = 3.14159265359f;
constexpr int kMaxManifoldPoints = 2;
constexpr int kMaxPolygonVertices = 8;
constexpr float kAabbExtension = 0.1f;
constexpr float kAabbMultiplier = 2.0f;
constexpr float kLinearSlop = 0.005f;
constexpr float kAngularSlop = (2.0f / 180.0f * kPi);
constexpr float kPolygonRadius = (2.0f * kLinearSlop);
constexpr int kMaxSubSteps = 8;
constexpr int kMaxTOIContacts = 32;
constexpr float kVelocityThreshold = 1.0f;
constexpr float kMaxLinearCorrection = 0.2f;
constexpr float kMaxAngularCorrection = (8.0f / 180.0f * kPi);
constexpr float kMaxTranslation = 2.0f;
constexpr float kMaxTranslationSquared = (kMaxTranslation * kMaxTranslation);
constexpr float kMaxRotation = (0.5f * kPi);
constexpr float kMaxRotationSquared = (kMaxRotation * kMaxRotation);
constexpr float kBaumgarte = 0.2f;
constexpr float kToiBaumgarte = 0.75f;
constexpr float kTimeToSleep = 0.5f;
constexpr float kLinearSleepTolerance = 0.01f;
constexpr float kAngularSleepTolerance = (2.0f / 180.0f * kPi);

typedef float2 v2;
DBX_HD v2 V(float x, float y) { return make_float2(x, y); }
DBX_HD v2 operator+(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
DBX_HD v2 operator-(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
DBX_HD v2 operator-(v2 a) { return V(-a.x, -a.y); }
DBX_HD v2 operator*(float s, v2 a) { return V(s * a.x, s * a.y); }
DBX_HD void operator+=(v2& a, v2 b) { a.x += b.x; a.y += b.y; }
DBX_HD void operator-=(v2& a, v2 b) { a.x -= b.x; a.y -= b.y; }
DBX_HD void operator*=(v2& a, float s) { a.x *= s; a.y *= s; }
DBX_HD float dot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
DBX_HD float cross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
DBX_HD v2 cross(v2 a, float s) { return V(s * a.y, -s * a.x); }
DBX_HD v2 cross(float s, v2 a) { return V(-s * a.y, s * a.x); }
DBX_HD float len2(v2 a) { return a.x * a.x + a.y * a.y; }
DBX_HD float len(v2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
DBX_HD float dist2(v2 a, v2 b) { v2 c = a - b; return dot(c, c); }
DBX_HD float dist(v2 a, v2 b) { v2 c = a - b; return len(c); }
// b2Vec2.Normalize (b2math.d:162-175)
DBX_HD float normalize(v2& a) {
  float l = len(a);
  if (l < kEpsilon) return 0.0f;
  float inv = 1.0f / l;
  a.x *= inv; a.y *= inv;
  return l;
}
// b2Min/b2Max/b2Clamp/b2Abs exactly as the reference spells them (b2math.d:769-828)
DBX_HD float fminr(float a, float b) { return a < b ? a : b; }
DBX_HD float fmaxr(float a, float b) { return a > b ? a : b; }
DBX_HD float fclampr(float a, float lo, float hi) { return fmaxr(lo, fminr(a, hi)); }
DBX_HD float fabsr(float a) { return a > 0.0f ? a : -a; }
DBX_HD v2 vmin(v2 a, v2 b) { return V(fminr(a.x, b.x), fminr(a.y, b.y)); }
DBX_HD v2 vmax(v2 a, v2 b) { return V(fmaxr(a.x, b.x), fmaxr(a.y, b.y)); }

struct Rot { float s, c; };
DBX_HD Rot R(float s, float c) { Rot r; r.s = s; r.c = c; return r; }
// b2Rot.Set (b2math.d:483-488).  On the device one sincosf: the same bits as sinf and cosf apart (checked on B200 for every float
// of magnitude 2^-20 .. 2^20, tools/sincos_test.cu), one range reduction instead of two
DBX_HD Rot rot_from_angle(float a) {
  Rot r;
#ifdef __CUDA_ARCH__
  sincosf(a, &r.s, &r.c);
#else
  r.s = sinf(a); r.c = cosf(a);
#endif
  return r;
}
DBX_HD Rot mul(Rot q, Rot r) { return R(q.s * r.c + q.c * r.s, q.c * r.c - q.s * r.s); }
DBX_HD Rot mulT(Rot q, Rot r) { return R(q.c * r.s - q.s * r.c, q.c * r.c + q.s * r.s); }
DBX_HD v2 mul(Rot q, v2 v) { return V(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
DBX_HD v2 mulT(Rot q, v2 v) { return V(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }

struct Xf { v2 p; Rot q; };
DBX_HD Xf XF(float4 f) { Xf x; x.p = V(f.x, f.y); x.q = R(f.z, f.w); return x; }
DBX_HD float4 pack(Xf x) { return make_float4(x.p.x, x.p.y, x.q.s, x.q.c); }
DBX_HD v2 mul(Xf T, v2 v) { return V((T.q.c * v.x - T.q.s * v.y) + T.p.x, (T.q.s * v.x + T.q.c * v.y) + T.p.y); }
DBX_HD v2 mulT(Xf T, v2 v) {
  float px = v.x - T.p.x, py = v.y - T.p.y;
  return V((T.q.c * px + T.q.s * py), (-T.q.s * px + T.q.c * py));
}
DBX_HD Xf mulT(Xf A, Xf B) { Xf C; C.q = mulT(A.q, B.q); C.p = mulT(A.q, B.p - A.p); return C; }
// transform of a body whose centre of mass is at c with angle a (b2body.d:1143-1147)
DBX_HD Xf xf_from_sweep(v2 c, float a, v2 localCenter) {
  Xf x; x.q = rot_from_angle(a); x.p = c - mul(x.q, localCenter); return x;
}

struct M22 { v2 ex, ey; };
DBX_HD v2 mul(M22 A, v2 v) { return V(A.ex.x * v.x + A.ey.x * v.y, A.ex.y * v.x + A.ey.y * v.y); }
// b2Mat22.GetInverse (b2math.d:322-337)
DBX_HD M22 inverse(M22 m) {
  float a = m.ex.x, b = m.ey.x, c = m.ex.y, d = m.ey.y;
  M22 B;
  float det = a * d - b * c;
  if (det != 0.0f) det = 1.0f / det;
  B.ex.x = det * d;  B.ey.x = -det * b;
  B.ex.y = -det * c; B.ey.y = det * a;
  return B;
}
// b2Mat22.Solve / b2Mat33.Solve22 (b2math.d:341-354, 402-415)
DBX_HD v2 solve22(float a11, float a12, float a21, float a22, v2 b) {
  float det = a11 * a22 - a12 * a21;
  if (det != 0.0f) det = 1.0f / det;
  return V(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
}
struct v3 { float x, y, z; };
DBX_HD v3 V3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
DBX_HD float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DBX_HD v3 cross(v3 a, v3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// b2Mat33.Solve33 (b2math.d:385-398)
DBX_HD v3 solve33(v3 ex, v3 ey, v3 ez, v3 b) {
  float det = dot(ex, cross(ey, ez));
  if (det != 0.0f) det = 1.0f / det;
  return V3(det * dot(b, cross(ey, ez)), det * dot(ex, cross(b, ez)), det * dot(ex, cross(ey, b)));
}

struct Box { v2 lo, hi; };  // b2AABB
DBX_HD Box BX(float4 f) { Box b; b.lo = V(f.x, f.y); b.hi = V(f.z, f.w); return b; }
DBX_HD float4 pack(Box b) { return make_float4(b.lo.x, b.lo.y, b.hi.x, b.hi.y); }
// b2TestOverlap(b2AABB, b2AABB) (collision/b2collision.d:470-483): closed intervals
DBX_HD bool overlap(Box a, Box b) {
  v2 d1 = b.lo - a.hi, d2 = a.lo - b.hi;
  if (d1.x > 0.0f || d1.y > 0.0f) return false;
  if (d2.x > 0.0f || d2.y > 0.0f) return false;
  return true;
}
DBX_HD bool contains(Box o, Box a) { return o.lo.x <= a.lo.x && o.lo.y <= a.lo.y && a.hi.x <= o.hi.x && a.hi.y <= o.hi.y; }
DBX_HD Box combine(Box a, Box b) { Box r; r.lo = vmin(a.lo, b.lo); r.hi = vmax(a.hi, b.hi); return r; }

// b2Sweep (b2math.d:552-593)
struct Sweep { v2 localCenter, c0, c; float a0, a, alpha0; };
DBX_HD Xf sweep_xf(const Sweep& s, float beta) {
  Xf xf;
  xf.p = (1.0f - beta) * s.c0 + beta * s.c;
  float angle = (1.0f - beta) * s.a0 + beta * s.a;
  xf.q = rot_from_angle(angle);
  xf.p -= mul(xf.q, s.localCenter);
  return xf;
}
DBX_HD void sweep_advance(Sweep& s, float alpha) {
  float beta = (alpha - s.alpha0) / (1.0f - s.alpha0);
  s.c0 += beta * (s.c - s.c0);
  s.a0 += beta * (s.a - s.a0);
  s.alpha0 = alpha;
}
DBX_HD void sweep_normalize(Sweep& s) {
  float twoPi = 2.0f * kPi;
  float d = twoPi * floorf(s.a0 / twoPi);
  s.a0 -= d;
  s.a -= d;
}

}  // namespace dbx
