// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// Shapes, manifolds, narrowphase, GJK distance and time-of-impact; restates
//   src/dbox/collision/shapes/b2{shape,circleshape,edgeshape,polygonshape,chainshape}.d
//   src/dbox/collision/b2collision.d, b2collidepolygon.d, b2collidecircle.d, b2collideedge.d,
//   src/dbox/collision/b2distance.d, b2timeofimpact.d
#pragma once
#include <vector>
#include "orc_math.h"

namespace orc {

enum ShapeType { kCircle = 0, kEdge = 1, kPolygon = 2, kChain = 3, kShapeTypeCount = 4 };  // b2shape.d:45-52

struct MassData { float mass = 0; V2 center; float I = 0; };

// One tagged struct instead of the reference's class hierarchy; fields follow
// b2circleshape.d:158, b2edgeshape.d:189-193, b2polygonshape.d:562-565, b2chainshape.d:257-263.
struct Shape {
  int type = kCircle;
  float radius = 0;
  // circle
  V2 p;
  // edge: v1,v2 (+ ghost v0,v3)
  V2 v0, v1, v2, v3;
  bool hasV0 = false, hasV3 = false;
  // polygon
  V2 centroid;
  V2 verts[kMaxPolygonVertices];
  V2 normals[kMaxPolygonVertices];
  int count = 0;
  // chain
  std::vector<V2> chain;
  V2 prevVertex, nextVertex;
  bool hasPrev = false, hasNext = false;

  static Shape circle(V2 p, float r);
  static Shape edge(V2 a, V2 b);
  static Shape box(float hx, float hy);
  static Shape box(float hx, float hy, V2 center, float angle);
  static Shape polygon(const V2* pts, int n);
  static Shape chainLoop(const V2* pts, int n);
  static Shape chainOpen(const V2* pts, int n);

  int childCount() const { return type == kChain ? (int)chain.size() - 1 : 1; }
  void childEdge(Shape* e, int index) const;               // b2chainshape.d:162-192
  void computeAABB(AABB* out, const Xf& xf, int child) const;
  void computeMass(MassData* md, float density) const;
  // b2Shape.RayCast (b2circleshape.d:67-94, b2edgeshape.d:96-150, b2polygonshape.d:279-332, b2chainshape.d:204-223)
  bool rayCast(float* fraction, V2* normal, V2 p1, V2 p2, float maxFraction, const Xf& xf, int child) const;
  // b2Shape.TestPoint (b2circleshape.d:60-65, b2polygonshape.d:265-279; false for edges b2edgeshape.d:84-87 and chains b2chainshape.d:196-199)
  bool testPoint(const Xf& xf, V2 pt) const;
};

// b2collision.d:38-114
struct ContactFeature { uint8_t indexA, indexB, typeA, typeB; };
union ContactID { ContactFeature cf; uint32_t key; };
enum { kFeatVertex = 0, kFeatFace = 1 };
struct ManifoldPoint { V2 localPoint; float normalImpulse = 0, tangentImpulse = 0; ContactID id{}; };
enum ManifoldType { kManCircles = 0, kManFaceA = 1, kManFaceB = 2 };
struct Manifold {
  ManifoldPoint points[kMaxManifoldPoints];
  V2 localNormal, localPoint;
  int type = 0;
  int pointCount = 0;
};
struct WorldManifold {
  V2 normal;
  V2 points[kMaxManifoldPoints];
  float separations[kMaxManifoldPoints] = {0, 0};
  void initialize(const Manifold* m, const Xf& xfA, float rA, const Xf& xfB, float rB);  // b2collision.d:123-191
};
struct ClipVertex { V2 v; ContactID id{}; };
int clipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], V2 normal, float offset, int vertexIndexA);

void collideCircles(Manifold* m, const Shape& a, const Xf& xfA, const Shape& b, const Xf& xfB);
void collidePolygonAndCircle(Manifold* m, const Shape& a, const Xf& xfA, const Shape& b, const Xf& xfB);
void collidePolygons(Manifold* m, const Shape& a, const Xf& xfA, const Shape& b, const Xf& xfB);
void collideEdgeAndCircle(Manifold* m, const Shape& a, const Xf& xfA, const Shape& b, const Xf& xfB);
void collideEdgeAndPolygon(Manifold* m, const Shape& a, const Xf& xfA, const Shape& b, const Xf& xfB);

// b2distance.d:30-180
struct DistanceProxy {
  V2 buffer[2];
  const V2* vertices = nullptr;
  int count = 0;
  float radius = 0;
  DistanceProxy() = default;
  DistanceProxy(const DistanceProxy& o) { *this = o; }
  DistanceProxy& operator=(const DistanceProxy& o) {
    buffer[0] = o.buffer[0]; buffer[1] = o.buffer[1]; count = o.count; radius = o.radius;
    vertices = (o.vertices == o.buffer) ? buffer : o.vertices;  // keep self-referencing proxies self-referencing
    return *this;
  }
  void set(const Shape& s, int index);
  int support(V2 d) const;
  V2 vertex(int i) const { return vertices[i]; }
};
struct SimplexCache { float metric = 0; uint16_t count = 0; uint8_t indexA[3] = {0, 0, 0}, indexB[3] = {0, 0, 0}; };
struct DistanceInput { DistanceProxy proxyA, proxyB; Xf transformA, transformB; bool useRadii = false; };
struct DistanceOutput { V2 pointA, pointB; float distance = 0; int iterations = 0; };
void distance(DistanceOutput* out, SimplexCache* cache, const DistanceInput* in);
bool testOverlap(const Shape& a, int ia, const Shape& b, int ib, const Xf& xfA, const Xf& xfB);  // b2collision.d:449-468

// b2timeofimpact.d:29-58
struct TOIInput { DistanceProxy proxyA, proxyB; Sweep sweepA, sweepB; float tMax = 0; };
enum TOIState { kToiUnknown, kToiFailed, kToiOverlapped, kToiTouching, kToiSeparated };
struct TOIOutput { int state = 0; float t = 0; };
void timeOfImpact(TOIOutput* out, const TOIInput* in);

}  // namespace orc
