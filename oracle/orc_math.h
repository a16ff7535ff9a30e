// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the product
// (dbox_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it, and only as the checker / CPU baseline.
//
// CPU restatement (C++17, single-threaded, fp32, no FMA contraction) of the reference's math layer:
//   src/dbox/common/b2math.d     (b2Vec2 :60-190, b2Mat22 :278-362, b2Mat33 :365-469, b2Rot :472-517,
//                                 b2Transform :521-546, b2Sweep :552-593, free functions :599-858)
//   src/dbox/common/b2settings.d (tuning constants :61-149)
// PARITY UNPINNED: the reference ships no golden vectors and cannot be compiled in this image (no D
// toolchain), so this oracle is validated by construction + physical invariants only.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

namespace orc {

// b2settings.d:61-149.  Constants are written as the float nearest the exact expression.
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

struct V2 {
  float x = 0, y = 0;
  V2() = default;
  V2(float x_, float y_) : x(x_), y(y_) {}
  V2 operator-() const { return V2(-x, -y); }
  void operator+=(V2 v) { x += v.x; y += v.y; }
  void operator-=(V2 v) { x -= v.x; y -= v.y; }
  void operator*=(float s) { x *= s; y *= s; }
  float len() const { return sqrtf(x * x + y * y); }
  float len2() const { return x * x + y * y; }
  // b2math.d:162-175
  float normalize() {
    float l = len();
    if (l < kEpsilon) return 0.0f;
    float inv = 1.0f / l;
    x *= inv; y *= inv;
    return l;
  }
  V2 skew() const { return V2(-y, x); }
};
inline V2 operator+(V2 a, V2 b) { return V2(a.x + b.x, a.y + b.y); }
inline V2 operator-(V2 a, V2 b) { return V2(a.x - b.x, a.y - b.y); }
inline V2 operator*(float s, V2 a) { return V2(s * a.x, s * a.y); }
inline V2 operator*(V2 a, float s) { return V2(s * a.x, s * a.y); }  // b2math.d:113-117 (s op x)
inline bool operator==(V2 a, V2 b) { return a.x == b.x && a.y == b.y; }

struct V3 {
  float x = 0, y = 0, z = 0;
  V3() = default;
  V3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  V3 operator-() const { return V3(-x, -y, -z); }
  void operator+=(V3 v) { x += v.x; y += v.y; z += v.z; }
  void operator-=(V3 v) { x -= v.x; y -= v.y; z -= v.z; }
  void operator*=(float s) { x *= s; y *= s; z *= s; }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(float s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }

inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
inline float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
inline V2 cross(V2 a, float s) { return V2(s * a.y, -s * a.x); }
inline V2 cross(float s, V2 a) { return V2(-s * a.y, s * a.x); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float dist(V2 a, V2 b) { V2 c = a - b; return c.len(); }
inline float dist2(V2 a, V2 b) { V2 c = a - b; return dot(c, c); }

template <class T> inline T absT(T a) { return a > T(0) ? a : -a; }
inline V2 absv(V2 a) { return V2(absT(a.x), absT(a.y)); }
template <class T> inline T minT(T a, T b) { return a < b ? a : b; }
template <class T> inline T maxT(T a, T b) { return a > b ? a : b; }
inline V2 minv(V2 a, V2 b) { return V2(minT(a.x, b.x), minT(a.y, b.y)); }
inline V2 maxv(V2 a, V2 b) { return V2(maxT(a.x, b.x), maxT(a.y, b.y)); }
template <class T> inline T clampT(T a, T lo, T hi) { return maxT(lo, minT(a, hi)); }

struct M22 {
  V2 ex, ey;
  M22() = default;
  M22(V2 c1, V2 c2) : ex(c1), ey(c2) {}
  // b2math.d:322-337
  M22 inverse() const {
    float a = ex.x, b = ey.x, c = ex.y, d = ey.y;
    M22 B;
    float det = a * d - b * c;
    if (det != 0.0f) det = 1.0f / det;
    B.ex.x = det * d;  B.ey.x = -det * b;
    B.ex.y = -det * c; B.ey.y = det * a;
    return B;
  }
  // b2math.d:341-354
  V2 solve(V2 b) const {
    float a11 = ex.x, a12 = ey.x, a21 = ex.y, a22 = ey.y;
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = 1.0f / det;
    V2 x;
    x.x = det * (a22 * b.x - a12 * b.y);
    x.y = det * (a11 * b.y - a21 * b.x);
    return x;
  }
};
inline V2 mul(const M22& A, V2 v) { return V2(A.ex.x * v.x + A.ey.x * v.y, A.ex.y * v.x + A.ey.y * v.y); }
inline V2 mulT(const M22& A, V2 v) { return V2(dot(v, A.ex), dot(v, A.ey)); }

struct M33 {
  V3 ex, ey, ez;
  // b2math.d:385-398
  V3 solve33(V3 b) const {
    float det = dot(ex, cross(ey, ez));
    if (det != 0.0f) det = 1.0f / det;
    V3 x;
    x.x = det * dot(b, cross(ey, ez));
    x.y = det * dot(ex, cross(b, ez));
    x.z = det * dot(ex, cross(ey, b));
    return x;
  }
  // b2math.d:402-415
  V2 solve22(V2 b) const {
    float a11 = ex.x, a12 = ey.x, a21 = ex.y, a22 = ey.y;
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = 1.0f / det;
    V2 x;
    x.x = det * (a22 * b.x - a12 * b.y);
    x.y = det * (a11 * b.y - a21 * b.x);
    return x;
  }
  // b2math.d:419-438
  void inverse22(M33* M) const {
    float a = ex.x, b = ey.x, c = ex.y, d = ey.y;
    float det = a * d - b * c;
    if (det != 0.0f) det = 1.0f / det;
    M->ex.x = det * d;  M->ey.x = -det * b; M->ex.z = 0.0f;
    M->ex.y = -det * c; M->ey.y = det * a;  M->ey.z = 0.0f;
    M->ez.x = 0.0f; M->ez.y = 0.0f; M->ez.z = 0.0f;
  }
  // b2math.d:442-466
  void symInverse33(M33* M) const {
    float det = dot(ex, cross(ey, ez));
    if (det != 0.0f) det = 1.0f / det;
    float a11 = ex.x, a12 = ey.x, a13 = ez.x;
    float a22 = ey.y, a23 = ez.y;
    float a33 = ez.z;
    M->ex.x = det * (a22 * a33 - a23 * a23);
    M->ex.y = det * (a13 * a23 - a12 * a33);
    M->ex.z = det * (a12 * a23 - a13 * a22);
    M->ey.x = M->ex.y;
    M->ey.y = det * (a11 * a33 - a13 * a13);
    M->ey.z = det * (a13 * a12 - a11 * a23);
    M->ez.x = M->ex.z;
    M->ez.y = M->ey.z;
    M->ez.z = det * (a11 * a22 - a12 * a12);
  }
};
inline V3 mul(const M33& A, V3 v) { return v.x * A.ex + v.y * A.ey + v.z * A.ez; }
inline V2 mul22(const M33& A, V2 v) { return V2(A.ex.x * v.x + A.ey.x * v.y, A.ex.y * v.x + A.ey.y * v.y); }

struct Rot {
  float s = 0, c = 1;
  Rot() = default;
  explicit Rot(float angle) { s = sinf(angle); c = cosf(angle); }  // b2math.d:475-480 (libc sinf/cosf)
  void set(float angle) { s = sinf(angle); c = cosf(angle); }
  float angle() const { return atan2f(s, c); }
  V2 xAxis() const { return V2(c, s); }
  V2 yAxis() const { return V2(-s, c); }
};
inline Rot mul(Rot q, Rot r) { Rot o; o.s = q.s * r.c + q.c * r.s; o.c = q.c * r.c - q.s * r.s; return o; }
inline Rot mulT(Rot q, Rot r) { Rot o; o.s = q.c * r.s - q.s * r.c; o.c = q.c * r.c + q.s * r.s; return o; }
inline V2 mul(Rot q, V2 v) { return V2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
inline V2 mulT(Rot q, V2 v) { return V2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }

struct Xf {
  V2 p;
  Rot q;
  Xf() = default;
  Xf(V2 p_, Rot q_) : p(p_), q(q_) {}
  void set(V2 pos, float angle) { p = pos; q.set(angle); }
};
// b2math.d:729-766
inline V2 mul(const Xf& T, V2 v) {
  float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
  float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
  return V2(x, y);
}
inline V2 mulT(const Xf& T, V2 v) {
  float px = v.x - T.p.x, py = v.y - T.p.y;
  float x = (T.q.c * px + T.q.s * py);
  float y = (-T.q.s * px + T.q.c * py);
  return V2(x, y);
}
inline Xf mul(const Xf& A, const Xf& B) { Xf C; C.q = mul(A.q, B.q); C.p = mul(A.q, B.p) + A.p; return C; }
inline Xf mulT(const Xf& A, const Xf& B) { Xf C; C.q = mulT(A.q, B.q); C.p = mulT(A.q, B.p - A.p); return C; }

// b2math.d:552-593
struct Sweep {
  V2 localCenter, c0, c;
  float a0 = 0, a = 0, alpha0 = 0;
  void getTransform(Xf* xf, float beta) const {
    xf->p = (1.0f - beta) * c0 + beta * c;
    float angle = (1.0f - beta) * a0 + beta * a;
    xf->q.set(angle);
    xf->p -= mul(xf->q, localCenter);
  }
  void advance(float alpha) {
    float beta = (alpha - alpha0) / (1.0f - alpha0);
    c0 += beta * (c - c0);
    a0 += beta * (a - a0);
    alpha0 = alpha;
  }
  void normalize() {
    float twoPi = 2.0f * kPi;
    float d = twoPi * floorf(a0 / twoPi);
    a0 -= d;
    a -= d;
  }
};

struct AABB {
  V2 lo, hi;
  V2 center() const { return 0.5f * (lo + hi); }
  V2 extents() const { return 0.5f * (hi - lo); }
  float perimeter() const { float wx = hi.x - lo.x, wy = hi.y - lo.y; return 2.0f * (wx + wy); }
  void combine(const AABB& a) { lo = minv(lo, a.lo); hi = maxv(hi, a.hi); }
  void combine(const AABB& a, const AABB& b) { lo = minv(a.lo, b.lo); hi = maxv(a.hi, b.hi); }
  bool contains(const AABB& a) const {
    bool r = true;
    r = r && lo.x <= a.lo.x;
    r = r && lo.y <= a.lo.y;
    r = r && a.hi.x <= hi.x;
    r = r && a.hi.y <= hi.y;
    return r;
  }
};
// b2collision.d:470-483
inline bool overlap(const AABB& a, const AABB& b) {
  V2 d1 = b.lo - a.hi, d2 = a.lo - b.hi;
  if (d1.x > 0.0f || d1.y > 0.0f) return false;
  if (d2.x > 0.0f || d2.y > 0.0f) return false;
  return true;
}

}  // namespace orc
