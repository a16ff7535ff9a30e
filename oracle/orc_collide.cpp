// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
// Restates the reference's shapes + narrowphase + GJK + TOI; each function cites the lines it follows.
#include "orc_collide.h"
#include <cassert>
#include <cstring>

namespace orc {

// ------------------------------------------------------------------ shapes
Shape Shape::circle(V2 p, float r) { Shape s; s.type = kCircle; s.p = p; s.radius = r; return s; }

// b2edgeshape.d:32-53
Shape Shape::edge(V2 a, V2 b) {
  Shape s; s.type = kEdge; s.radius = kPolygonRadius; s.v1 = a; s.v2 = b; s.hasV0 = s.hasV3 = false; return s;
}

// b2polygonshape.d:220-232
Shape Shape::box(float hx, float hy) {
  Shape s; s.type = kPolygon; s.radius = kPolygonRadius; s.count = 4;
  s.verts[0] = V2(-hx, -hy); s.verts[1] = V2(hx, -hy); s.verts[2] = V2(hx, hy); s.verts[3] = V2(-hx, hy);
  s.normals[0] = V2(0.0f, -1.0f); s.normals[1] = V2(1.0f, 0.0f); s.normals[2] = V2(0.0f, 1.0f); s.normals[3] = V2(-1.0f, 0.0f);
  s.centroid = V2(0, 0);
  return s;
}

// b2polygonshape.d:239-262
Shape Shape::box(float hx, float hy, V2 center, float angle) {
  Shape s = box(hx, hy);
  s.centroid = center;
  Xf xf; xf.p = center; xf.q.set(angle);
  for (int i = 0; i < s.count; ++i) { s.verts[i] = mul(xf, s.verts[i]); s.normals[i] = mul(xf.q, s.normals[i]); }
  return s;
}

// b2polygonshape.d:506-554
static V2 computeCentroid(const V2* vs, int count) {
  V2 c(0.0f, 0.0f);
  float area = 0.0f;
  V2 pRef(0.0f, 0.0f);
  const float inv3 = 1.0f / 3.0f;
  for (int i = 0; i < count; ++i) {
    V2 p1 = pRef, p2 = vs[i], p3 = i + 1 < count ? vs[i + 1] : vs[0];
    V2 e1 = p2 - p1, e2 = p3 - p1;
    float D = cross(e1, e2);
    float triangleArea = 0.5f * D;
    area += triangleArea;
    c += triangleArea * inv3 * (p1 + p2 + p3);
  }
  c *= 1.0f / area;
  return c;
}

// b2polygonshape.d:74-209 (weld close points, gift-wrap hull, normals, centroid)
Shape Shape::polygon(const V2* vertices, int count) {
  if (count < 3) return box(1.0f, 1.0f);
  Shape s; s.type = kPolygon; s.radius = kPolygonRadius;
  int n = minT(count, kMaxPolygonVertices);
  V2 ps[kMaxPolygonVertices];
  int tempCount = 0;
  for (int i = 0; i < n; ++i) {
    V2 v = vertices[i];
    bool unique = true;
    for (int j = 0; j < tempCount; ++j) {
      if (dist2(v, ps[j]) < 0.5f * kLinearSlop) { unique = false; break; }
    }
    if (unique) ps[tempCount++] = v;
  }
  n = tempCount;
  if (n < 3) return box(1.0f, 1.0f);
  int i0 = 0;
  float x0 = ps[0].x;
  for (int i = 1; i < n; ++i) {
    float x = ps[i].x;
    if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
  }
  int hull[kMaxPolygonVertices];
  int m = 0, ih = i0;
  for (;;) {
    hull[m] = ih;
    int ie = 0;
    for (int j = 1; j < n; ++j) {
      if (ie == ih) { ie = j; continue; }
      V2 r = ps[ie] - ps[hull[m]];
      V2 v = ps[j] - ps[hull[m]];
      float c = cross(r, v);
      if (c < 0.0f) ie = j;
      if (c == 0.0f && v.len2() > r.len2()) ie = j;
    }
    ++m;
    ih = ie;
    if (ie == i0) break;
  }
  if (m < 3) return box(1.0f, 1.0f);
  s.count = m;
  for (int i = 0; i < m; ++i) s.verts[i] = ps[hull[i]];
  for (int i = 0; i < m; ++i) {
    int i1 = i, i2 = i + 1 < m ? i + 1 : 0;
    V2 e = s.verts[i2] - s.verts[i1];
    s.normals[i] = cross(e, 1.0f);
    s.normals[i].normalize();
  }
  s.centroid = computeCentroid(s.verts, m);
  return s;
}

// b2chainshape.d:64-86
Shape Shape::chainLoop(const V2* pts, int n) {
  Shape s; s.type = kChain; s.radius = kPolygonRadius;
  s.chain.assign(pts, pts + n);
  s.chain.push_back(s.chain[0]);
  s.prevVertex = s.chain[s.chain.size() - 2];
  s.nextVertex = s.chain[1];
  s.hasPrev = s.hasNext = true;
  return s;
}
// b2chainshape.d:97-117
Shape Shape::chainOpen(const V2* pts, int n) {
  Shape s; s.type = kChain; s.radius = kPolygonRadius;
  s.chain.assign(pts, pts + n);
  s.hasPrev = s.hasNext = false;
  return s;
}

void Shape::childEdge(Shape* e, int index) const {
  int cnt = (int)chain.size();
  e->type = kEdge;
  e->radius = radius;
  e->v1 = chain[index + 0];
  e->v2 = chain[index + 1];
  if (index > 0) { e->v0 = chain[index - 1]; e->hasV0 = true; }
  else { e->v0 = prevVertex; e->hasV0 = hasPrev; }
  if (index < cnt - 2) { e->v3 = chain[index + 2]; e->hasV3 = true; }
  else { e->v3 = nextVertex; e->hasV3 = hasNext; }
}

// b2circleshape.d:110-117, b2edgeshape.d:163-176, b2polygonshape.d:356-373, b2chainshape.d:226-243
void Shape::computeAABB(AABB* aabb, const Xf& xf, int child) const {
  switch (type) {
    case kCircle: {
      V2 c = xf.p + mul(xf.q, p);
      aabb->lo = V2(c.x - radius, c.y - radius);
      aabb->hi = V2(c.x + radius, c.y + radius);
    } break;
    case kEdge: {
      V2 a = mul(xf, v1), b = mul(xf, v2);
      V2 lower = minv(a, b), upper = maxv(a, b);
      V2 r(radius, radius);
      aabb->lo = lower - r;
      aabb->hi = upper + r;
    } break;
    case kPolygon: {
      V2 lower = mul(xf, verts[0]);
      V2 upper = lower;
      for (int i = 1; i < count; ++i) {
        V2 v = mul(xf, verts[i]);
        lower = minv(lower, v);
        upper = maxv(upper, v);
      }
      V2 r(radius, radius);
      aabb->lo = lower - r;
      aabb->hi = upper + r;
    } break;
    case kChain: {
      int i1 = child, i2 = child + 1;
      if (i2 == (int)chain.size()) i2 = 0;
      V2 a = mul(xf, chain[i1]), b = mul(xf, chain[i2]);
      aabb->lo = minv(a, b);
      aabb->hi = maxv(a, b);
    } break;
  }
}

// b2circleshape.d:120-127, b2edgeshape.d:179-186, b2polygonshape.d:376-459, b2chainshape.d:247-254
void Shape::computeMass(MassData* md, float density) const {
  switch (type) {
    case kCircle:
      md->mass = density * kPi * radius * radius;
      md->center = p;
      md->I = md->mass * (0.5f * radius * radius + dot(p, p));
      break;
    case kEdge:
      md->mass = 0.0f; md->center = 0.5f * (v1 + v2); md->I = 0.0f;
      break;
    case kChain:
      md->mass = 0.0f; md->center = V2(0, 0); md->I = 0.0f;
      break;
    case kPolygon: {
      V2 center(0.0f, 0.0f);
      float area = 0.0f, I = 0.0f;
      V2 s(0.0f, 0.0f);
      for (int i = 0; i < count; ++i) s += verts[i];
      s *= 1.0f / count;
      const float k_inv3 = 1.0f / 3.0f;
      for (int i = 0; i < count; ++i) {
        V2 e1 = verts[i] - s;
        V2 e2 = i + 1 < count ? verts[i + 1] - s : verts[0] - s;
        float D = cross(e1, e2);
        float triangleArea = 0.5f * D;
        area += triangleArea;
        center += triangleArea * k_inv3 * (e1 + e2);
        float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
        float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
        float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
        I += (0.25f * k_inv3 * D) * (intx2 + inty2);
      }
      md->mass = density * area;
      center *= 1.0f / area;
      md->center = center + s;
      md->I = density * I;
      md->I += md->mass * (dot(md->center, md->center) - dot(center, center));
    } break;
  }
}

// ------------------------------------------------------------------ manifolds
void WorldManifold::initialize(const Manifold* manifold, const Xf& xfA, float radiusA, const Xf& xfB, float radiusB) {
  if (manifold->pointCount == 0) return;
  switch (manifold->type) {
    case kManCircles: {
      normal = V2(1.0f, 0.0f);
      V2 pointA = mul(xfA, manifold->localPoint);
      V2 pointB = mul(xfB, manifold->points[0].localPoint);
      if (dist2(pointA, pointB) > kEpsilon * kEpsilon) { normal = pointB - pointA; normal.normalize(); }
      V2 cA = pointA + radiusA * normal;
      V2 cB = pointB - radiusB * normal;
      points[0] = 0.5f * (cA + cB);
      separations[0] = dot(cB - cA, normal);
    } break;
    case kManFaceA: {
      normal = mul(xfA.q, manifold->localNormal);
      V2 planePoint = mul(xfA, manifold->localPoint);
      for (int i = 0; i < manifold->pointCount; ++i) {
        V2 clipPoint = mul(xfB, manifold->points[i].localPoint);
        V2 cA = clipPoint + (radiusA - dot(clipPoint - planePoint, normal)) * normal;
        V2 cB = clipPoint - radiusB * normal;
        points[i] = 0.5f * (cA + cB);
        separations[i] = dot(cB - cA, normal);
      }
    } break;
    case kManFaceB: {
      normal = mul(xfB.q, manifold->localNormal);
      V2 planePoint = mul(xfB, manifold->localPoint);
      for (int i = 0; i < manifold->pointCount; ++i) {
        V2 clipPoint = mul(xfA, manifold->points[i].localPoint);
        V2 cB = clipPoint + (radiusB - dot(clipPoint - planePoint, normal)) * normal;
        V2 cA = clipPoint - radiusA * normal;
        points[i] = 0.5f * (cA + cB);
        separations[i] = dot(cA - cB, normal);
      }
      normal = -normal;
    } break;
  }
}

// b2collision.d:413-446
int clipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], V2 normal, float offset, int vertexIndexA) {
  int numOut = 0;
  float distance0 = dot(normal, vIn[0].v) - offset;
  float distance1 = dot(normal, vIn[1].v) - offset;
  if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
  if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];
  if (distance0 * distance1 < 0.0f) {
    float interp = distance0 / (distance0 - distance1);
    vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
    vOut[numOut].id.cf.indexA = (uint8_t)vertexIndexA;
    vOut[numOut].id.cf.indexB = vIn[0].id.cf.indexB;
    vOut[numOut].id.cf.typeA = kFeatVertex;
    vOut[numOut].id.cf.typeB = kFeatFace;
    ++numOut;
  }
  return numOut;
}

// b2collidecircle.d:29-56
void collideCircles(Manifold* manifold, const Shape& circleA, const Xf& xfA, const Shape& circleB, const Xf& xfB) {
  manifold->pointCount = 0;
  V2 pA = mul(xfA, circleA.p), pB = mul(xfB, circleB.p);
  V2 d = pB - pA;
  float distSqr = dot(d, d);
  float rA = circleA.radius, rB = circleB.radius;
  float radius = rA + rB;
  if (distSqr > radius * radius) return;
  manifold->type = kManCircles;
  manifold->localPoint = circleA.p;
  manifold->localNormal = V2(0, 0);
  manifold->pointCount = 1;
  manifold->points[0].localPoint = circleB.p;
  manifold->points[0].id.key = 0;
}

// b2collidecircle.d:59-164
void collidePolygonAndCircle(Manifold* manifold, const Shape& polygonA, const Xf& xfA, const Shape& circleB, const Xf& xfB) {
  manifold->pointCount = 0;
  V2 c = mul(xfB, circleB.p);
  V2 cLocal = mulT(xfA, c);
  int normalIndex = 0;
  float separation = -kMaxFloat;
  float radius = polygonA.radius + circleB.radius;
  int vertexCount = polygonA.count;
  const V2* vertices = polygonA.verts;
  const V2* normals = polygonA.normals;
  for (int i = 0; i < vertexCount; ++i) {
    float s = dot(normals[i], cLocal - vertices[i]);
    if (s > radius) return;
    if (s > separation) { separation = s; normalIndex = i; }
  }
  int vertIndex1 = normalIndex;
  int vertIndex2 = vertIndex1 + 1 < vertexCount ? vertIndex1 + 1 : 0;
  V2 v1 = vertices[vertIndex1], v2 = vertices[vertIndex2];
  if (separation < kEpsilon) {
    manifold->pointCount = 1;
    manifold->type = kManFaceA;
    manifold->localNormal = normals[normalIndex];
    manifold->localPoint = 0.5f * (v1 + v2);
    manifold->points[0].localPoint = circleB.p;
    manifold->points[0].id.key = 0;
    return;
  }
  float u1 = dot(cLocal - v1, v2 - v1);
  float u2 = dot(cLocal - v2, v1 - v2);
  if (u1 <= 0.0f) {
    if (dist2(cLocal, v1) > radius * radius) return;
    manifold->pointCount = 1;
    manifold->type = kManFaceA;
    manifold->localNormal = cLocal - v1;
    manifold->localNormal.normalize();
    manifold->localPoint = v1;
    manifold->points[0].localPoint = circleB.p;
    manifold->points[0].id.key = 0;
  } else if (u2 <= 0.0f) {
    if (dist2(cLocal, v2) > radius * radius) return;
    manifold->pointCount = 1;
    manifold->type = kManFaceA;
    manifold->localNormal = cLocal - v2;
    manifold->localNormal.normalize();
    manifold->localPoint = v2;
    manifold->points[0].localPoint = circleB.p;
    manifold->points[0].id.key = 0;
  } else {
    V2 faceCenter = 0.5f * (v1 + v2);
    float separation2 = dot(cLocal - faceCenter, normals[vertIndex1]);
    if (separation2 > radius) return;
    manifold->pointCount = 1;
    manifold->type = kManFaceA;
    manifold->localNormal = normals[vertIndex1];
    manifold->localPoint = faceCenter;
    manifold->points[0].localPoint = circleB.p;
    manifold->points[0].id.key = 0;
  }
}

// b2collidepolygon.d:29-71
static float findMaxSeparation(int* edgeIndex, const Shape& poly1, const Xf& xf1, const Shape& poly2, const Xf& xf2) {
  int count1 = poly1.count, count2 = poly2.count;
  const V2* n1s = poly1.normals;
  const V2* v1s = poly1.verts;
  const V2* v2s = poly2.verts;
  Xf xf = mulT(xf2, xf1);
  int bestIndex = 0;
  float maxSeparation = -kMaxFloat;
  for (int i = 0; i < count1; ++i) {
    V2 n = mul(xf.q, n1s[i]);
    V2 v1 = mul(xf, v1s[i]);
    float si = kMaxFloat;
    for (int j = 0; j < count2; ++j) {
      float sij = dot(n, v2s[j] - v1);
      if (sij < si) si = sij;
    }
    if (si > maxSeparation) { maxSeparation = si; bestIndex = i; }
  }
  *edgeIndex = bestIndex;
  return maxSeparation;
}

// b2collidepolygon.d:73-118
static void findIncidentEdge(ClipVertex c[2], const Shape& poly1, const Xf& xf1, int edge1, const Shape& poly2, const Xf& xf2) {
  const V2* normals1 = poly1.normals;
  int count2 = poly2.count;
  const V2* vertices2 = poly2.verts;
  const V2* normals2 = poly2.normals;
  V2 normal1 = mulT(xf2.q, mul(xf1.q, normals1[edge1]));
  int index = 0;
  float minDot = kMaxFloat;
  for (int i = 0; i < count2; ++i) {
    float d = dot(normal1, normals2[i]);
    if (d < minDot) { minDot = d; index = i; }
  }
  int i1 = index, i2 = i1 + 1 < count2 ? i1 + 1 : 0;
  c[0].v = mul(xf2, vertices2[i1]);
  c[0].id.cf.indexA = (uint8_t)edge1; c[0].id.cf.indexB = (uint8_t)i1;
  c[0].id.cf.typeA = kFeatFace; c[0].id.cf.typeB = kFeatVertex;
  c[1].v = mul(xf2, vertices2[i2]);
  c[1].id.cf.indexA = (uint8_t)edge1; c[1].id.cf.indexB = (uint8_t)i2;
  c[1].id.cf.typeA = kFeatFace; c[1].id.cf.typeB = kFeatVertex;
}

// b2collidepolygon.d:127-254
void collidePolygons(Manifold* manifold, const Shape& polyA, const Xf& xfA, const Shape& polyB, const Xf& xfB) {
  manifold->pointCount = 0;
  float totalRadius = polyA.radius + polyB.radius;
  int edgeA = 0;
  float separationA = findMaxSeparation(&edgeA, polyA, xfA, polyB, xfB);
  if (separationA > totalRadius) return;
  int edgeB = 0;
  float separationB = findMaxSeparation(&edgeB, polyB, xfB, polyA, xfA);
  if (separationB > totalRadius) return;

  const Shape* poly1; const Shape* poly2;
  Xf xf1, xf2;
  int edge1;
  uint8_t flip;
  const float k_tol = 0.1f * kLinearSlop;
  if (separationB > separationA + k_tol) {
    poly1 = &polyB; poly2 = &polyA; xf1 = xfB; xf2 = xfA; edge1 = edgeB; manifold->type = kManFaceB; flip = 1;
  } else {
    poly1 = &polyA; poly2 = &polyB; xf1 = xfA; xf2 = xfB; edge1 = edgeA; manifold->type = kManFaceA; flip = 0;
  }
  ClipVertex incidentEdge[2];
  findIncidentEdge(incidentEdge, *poly1, xf1, edge1, *poly2, xf2);
  int count1 = poly1->count;
  const V2* vertices1 = poly1->verts;
  int iv1 = edge1, iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
  V2 v11 = vertices1[iv1], v12 = vertices1[iv2];
  V2 localTangent = v12 - v11;
  localTangent.normalize();
  V2 localNormal = cross(localTangent, 1.0f);
  V2 planePoint = 0.5f * (v11 + v12);
  V2 tangent = mul(xf1.q, localTangent);
  V2 normal = cross(tangent, 1.0f);
  v11 = mul(xf1, v11);
  v12 = mul(xf1, v12);
  float frontOffset = dot(normal, v11);
  float sideOffset1 = -dot(tangent, v11) + totalRadius;
  float sideOffset2 = dot(tangent, v12) + totalRadius;
  ClipVertex clipPoints1[2], clipPoints2[2];
  int np = clipSegmentToLine(clipPoints1, incidentEdge, -tangent, sideOffset1, iv1);
  if (np < 2) return;
  np = clipSegmentToLine(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
  if (np < 2) return;
  manifold->localNormal = localNormal;
  manifold->localPoint = planePoint;
  int pointCount = 0;
  for (int i = 0; i < kMaxManifoldPoints; ++i) {
    float separation = dot(normal, clipPoints2[i].v) - frontOffset;
    if (separation <= totalRadius) {
      ManifoldPoint* cp = manifold->points + pointCount;
      cp->localPoint = mulT(xf2, clipPoints2[i].v);
      cp->id = clipPoints2[i].id;
      if (flip) {
        ContactFeature cf = cp->id.cf;
        cp->id.cf.indexA = cf.indexB; cp->id.cf.indexB = cf.indexA;
        cp->id.cf.typeA = cf.typeB;   cp->id.cf.typeB = cf.typeA;
      }
      ++pointCount;
    }
  }
  manifold->pointCount = pointCount;
}

// b2collideedge.d:31-160
void collideEdgeAndCircle(Manifold* manifold, const Shape& edgeA, const Xf& xfA, const Shape& circleB, const Xf& xfB) {
  manifold->pointCount = 0;
  V2 Q = mulT(xfA, mul(xfB, circleB.p));
  V2 A = edgeA.v1, B = edgeA.v2;
  V2 e = B - A;
  float u = dot(e, B - Q);
  float v = dot(e, Q - A);
  float radius = edgeA.radius + circleB.radius;
  ContactFeature cf; cf.indexA = 0; cf.typeA = 0;
  cf.indexB = 0; cf.typeB = kFeatVertex;
  if (v <= 0.0f) {
    V2 P = A, d = Q - P;
    float dd = dot(d, d);
    if (dd > radius * radius) return;
    if (edgeA.hasV0) {
      V2 A1 = edgeA.v0, B1 = A, e1 = B1 - A1;
      float u1 = dot(e1, B1 - Q);
      if (u1 > 0.0f) return;
    }
    cf.indexA = 0; cf.typeA = kFeatVertex;
    manifold->pointCount = 1;
    manifold->type = kManCircles;
    manifold->localNormal = V2(0, 0);
    manifold->localPoint = P;
    manifold->points[0].id.key = 0;
    manifold->points[0].id.cf = cf;
    manifold->points[0].localPoint = circleB.p;
    return;
  }
  if (u <= 0.0f) {
    V2 P = B, d = Q - P;
    float dd = dot(d, d);
    if (dd > radius * radius) return;
    if (edgeA.hasV3) {
      V2 B2 = edgeA.v3, A2 = B, e2 = B2 - A2;
      float v2 = dot(e2, Q - A2);
      if (v2 > 0.0f) return;
    }
    cf.indexA = 1; cf.typeA = kFeatVertex;
    manifold->pointCount = 1;
    manifold->type = kManCircles;
    manifold->localNormal = V2(0, 0);
    manifold->localPoint = P;
    manifold->points[0].id.key = 0;
    manifold->points[0].id.cf = cf;
    manifold->points[0].localPoint = circleB.p;
    return;
  }
  float den = dot(e, e);
  V2 P = (1.0f / den) * (u * A + v * B);
  V2 d = Q - P;
  float dd = dot(d, d);
  if (dd > radius * radius) return;
  V2 n(-e.y, e.x);
  if (dot(n, Q - A) < 0.0f) n = V2(-n.x, -n.y);
  n.normalize();
  cf.indexA = 0; cf.typeA = kFeatFace;
  manifold->pointCount = 1;
  manifold->type = kManFaceA;
  manifold->localNormal = n;
  manifold->localPoint = A;
  manifold->points[0].id.key = 0;
  manifold->points[0].id.cf = cf;
  manifold->points[0].localPoint = circleB.p;
}

// b2collideedge.d:163-715 (b2EPCollider)
namespace {
enum AxisType { kAxUnknown, kAxEdgeA, kAxEdgeB };
struct EPAxis { int type; int index; float separation; };
struct TempPolygon { V2 vertices[kMaxPolygonVertices], normals[kMaxPolygonVertices]; int count; };
struct ReferenceFace { int i1, i2; V2 v1, v2, normal, sideNormal1; float sideOffset1; V2 sideNormal2; float sideOffset2; };

struct EPCollider {
  TempPolygon polygonB;
  Xf xf;
  V2 centroidB, v0, v1, v2, v3, normal0, normal1, normal2, normal, lowerLimit, upperLimit;
  float radius = 0;
  bool front = false;

  EPAxis computeEdgeSeparation() const {
    EPAxis axis; axis.type = kAxEdgeA; axis.index = front ? 0 : 1; axis.separation = FLT_MAX;
    for (int i = 0; i < polygonB.count; ++i) {
      float s = dot(normal, polygonB.vertices[i] - v1);
      if (s < axis.separation) axis.separation = s;
    }
    return axis;
  }
  EPAxis computePolygonSeparation() const {
    EPAxis axis; axis.type = kAxUnknown; axis.index = -1; axis.separation = -FLT_MAX;
    V2 perp(-normal.y, normal.x);
    for (int i = 0; i < polygonB.count; ++i) {
      V2 n = -polygonB.normals[i];
      float s1 = dot(n, polygonB.vertices[i] - v1);
      float s2 = dot(n, polygonB.vertices[i] - v2);
      float s = minT(s1, s2);
      if (s > radius) { axis.type = kAxEdgeB; axis.index = i; axis.separation = s; return axis; }
      if (dot(n, perp) >= 0.0f) {
        if (dot(n - upperLimit, normal) < -kAngularSlop) continue;
      } else {
        if (dot(n - lowerLimit, normal) < -kAngularSlop) continue;
      }
      if (s > axis.separation) { axis.type = kAxEdgeB; axis.index = i; axis.separation = s; }
    }
    return axis;
  }

  void collide(Manifold* manifold, const Shape& edgeA, const Xf& xfA, const Shape& polyB, const Xf& xfB) {
    xf = mulT(xfA, xfB);
    centroidB = mul(xf, polyB.centroid);
    v0 = edgeA.v0; v1 = edgeA.v1; v2 = edgeA.v2; v3 = edgeA.v3;
    bool hasVertex0 = edgeA.hasV0, hasVertex3 = edgeA.hasV3;
    V2 edge1 = v2 - v1;
    edge1.normalize();
    normal1 = V2(edge1.y, -edge1.x);
    float offset1 = dot(normal1, centroidB - v1);
    float offset0 = 0.0f, offset2 = 0.0f;
    bool convex1 = false, convex2 = false;
    if (hasVertex0) {
      V2 edge0 = v1 - v0;
      edge0.normalize();
      normal0 = V2(edge0.y, -edge0.x);
      convex1 = cross(edge0, edge1) >= 0.0f;
      offset0 = dot(normal0, centroidB - v0);
    }
    if (hasVertex3) {
      V2 edge2 = v3 - v2;
      edge2.normalize();
      normal2 = V2(edge2.y, -edge2.x);
      convex2 = cross(edge1, edge2) > 0.0f;
      offset2 = dot(normal2, centroidB - v2);
    }
    if (hasVertex0 && hasVertex3) {
      if (convex1 && convex2) {
        front = offset0 >= 0.0f || offset1 >= 0.0f || offset2 >= 0.0f;
        if (front) { normal = normal1; lowerLimit = normal0; upperLimit = normal2; }
        else { normal = -normal1; lowerLimit = -normal1; upperLimit = -normal1; }
      } else if (convex1) {
        front = offset0 >= 0.0f || (offset1 >= 0.0f && offset2 >= 0.0f);
        if (front) { normal = normal1; lowerLimit = normal0; upperLimit = normal1; }
        else { normal = -normal1; lowerLimit = -normal2; upperLimit = -normal1; }
      } else if (convex2) {
        front = offset2 >= 0.0f || (offset0 >= 0.0f && offset1 >= 0.0f);
        if (front) { normal = normal1; lowerLimit = normal1; upperLimit = normal2; }
        else { normal = -normal1; lowerLimit = -normal1; upperLimit = -normal0; }
      } else {
        front = offset0 >= 0.0f && offset1 >= 0.0f && offset2 >= 0.0f;
        if (front) { normal = normal1; lowerLimit = normal1; upperLimit = normal1; }
        else { normal = -normal1; lowerLimit = -normal2; upperLimit = -normal0; }
      }
    } else if (hasVertex0) {
      if (convex1) {
        front = offset0 >= 0.0f || offset1 >= 0.0f;
        if (front) { normal = normal1; lowerLimit = normal0; upperLimit = -normal1; }
        else { normal = -normal1; lowerLimit = normal1; upperLimit = -normal1; }
      } else {
        front = offset0 >= 0.0f && offset1 >= 0.0f;
        if (front) { normal = normal1; lowerLimit = normal1; upperLimit = -normal1; }
        else { normal = -normal1; lowerLimit = normal1; upperLimit = -normal0; }
      }
    } else if (hasVertex3) {
      if (convex2) {
        front = offset1 >= 0.0f || offset2 >= 0.0f;
        if (front) { normal = normal1; lowerLimit = -normal1; upperLimit = normal2; }
        else { normal = -normal1; lowerLimit = -normal1; upperLimit = normal1; }
      } else {
        front = offset1 >= 0.0f && offset2 >= 0.0f;
        if (front) { normal = normal1; lowerLimit = -normal1; upperLimit = normal1; }
        else { normal = -normal1; lowerLimit = -normal2; upperLimit = normal1; }
      }
    } else {
      front = offset1 >= 0.0f;
      if (front) { normal = normal1; lowerLimit = -normal1; upperLimit = -normal1; }
      else { normal = -normal1; lowerLimit = normal1; upperLimit = normal1; }
    }
    polygonB.count = polyB.count;
    for (int i = 0; i < polyB.count; ++i) {
      polygonB.vertices[i] = mul(xf, polyB.verts[i]);
      polygonB.normals[i] = mul(xf.q, polyB.normals[i]);
    }
    radius = 2.0f * kPolygonRadius;
    manifold->pointCount = 0;
    EPAxis edgeAxis = computeEdgeSeparation();
    if (edgeAxis.type == kAxUnknown) return;
    if (edgeAxis.separation > radius) return;
    EPAxis polygonAxis = computePolygonSeparation();
    if (polygonAxis.type != kAxUnknown && polygonAxis.separation > radius) return;
    const float k_relativeTol = 0.98f;
    const float k_absoluteTol = 0.001f;
    EPAxis primaryAxis;
    if (polygonAxis.type == kAxUnknown) primaryAxis = edgeAxis;
    else if (polygonAxis.separation > k_relativeTol * edgeAxis.separation + k_absoluteTol) primaryAxis = polygonAxis;
    else primaryAxis = edgeAxis;

    ClipVertex ie[2];
    ReferenceFace rf;
    if (primaryAxis.type == kAxEdgeA) {
      manifold->type = kManFaceA;
      int bestIndex = 0;
      float bestValue = dot(normal, polygonB.normals[0]);
      for (int i = 1; i < polygonB.count; ++i) {
        float value = dot(normal, polygonB.normals[i]);
        if (value < bestValue) { bestValue = value; bestIndex = i; }
      }
      int i1 = bestIndex, i2 = i1 + 1 < polygonB.count ? i1 + 1 : 0;
      ie[0].v = polygonB.vertices[i1];
      ie[0].id.cf.indexA = 0; ie[0].id.cf.indexB = (uint8_t)i1; ie[0].id.cf.typeA = kFeatFace; ie[0].id.cf.typeB = kFeatVertex;
      ie[1].v = polygonB.vertices[i2];
      ie[1].id.cf.indexA = 0; ie[1].id.cf.indexB = (uint8_t)i2; ie[1].id.cf.typeA = kFeatFace; ie[1].id.cf.typeB = kFeatVertex;
      if (front) { rf.i1 = 0; rf.i2 = 1; rf.v1 = v1; rf.v2 = v2; rf.normal = normal1; }
      else { rf.i1 = 1; rf.i2 = 0; rf.v1 = v2; rf.v2 = v1; rf.normal = -normal1; }
    } else {
      manifold->type = kManFaceB;
      ie[0].v = v1;
      ie[0].id.cf.indexA = 0; ie[0].id.cf.indexB = (uint8_t)primaryAxis.index; ie[0].id.cf.typeA = kFeatVertex; ie[0].id.cf.typeB = kFeatFace;
      ie[1].v = v2;
      ie[1].id.cf.indexA = 0; ie[1].id.cf.indexB = (uint8_t)primaryAxis.index; ie[1].id.cf.typeA = kFeatVertex; ie[1].id.cf.typeB = kFeatFace;
      rf.i1 = primaryAxis.index;
      rf.i2 = rf.i1 + 1 < polygonB.count ? rf.i1 + 1 : 0;
      rf.v1 = polygonB.vertices[rf.i1];
      rf.v2 = polygonB.vertices[rf.i2];
      rf.normal = polygonB.normals[rf.i1];
    }
    rf.sideNormal1 = V2(rf.normal.y, -rf.normal.x);
    rf.sideNormal2 = -rf.sideNormal1;
    rf.sideOffset1 = dot(rf.sideNormal1, rf.v1);
    rf.sideOffset2 = dot(rf.sideNormal2, rf.v2);
    ClipVertex clipPoints1[2], clipPoints2[2];
    int np = clipSegmentToLine(clipPoints1, ie, rf.sideNormal1, rf.sideOffset1, rf.i1);
    if (np < kMaxManifoldPoints) return;
    np = clipSegmentToLine(clipPoints2, clipPoints1, rf.sideNormal2, rf.sideOffset2, rf.i2);
    if (np < kMaxManifoldPoints) return;
    if (primaryAxis.type == kAxEdgeA) { manifold->localNormal = rf.normal; manifold->localPoint = rf.v1; }
    else { manifold->localNormal = polyB.normals[rf.i1]; manifold->localPoint = polyB.verts[rf.i1]; }
    int pointCount = 0;
    for (int i = 0; i < kMaxManifoldPoints; ++i) {
      float separation = dot(rf.normal, clipPoints2[i].v - rf.v1);
      if (separation <= radius) {
        ManifoldPoint* cp = manifold->points + pointCount;
        if (primaryAxis.type == kAxEdgeA) {
          cp->localPoint = mulT(xf, clipPoints2[i].v);
          cp->id = clipPoints2[i].id;
        } else {
          cp->localPoint = clipPoints2[i].v;
          cp->id.cf.typeA = clipPoints2[i].id.cf.typeB;
          cp->id.cf.typeB = clipPoints2[i].id.cf.typeA;
          cp->id.cf.indexA = clipPoints2[i].id.cf.indexB;
          cp->id.cf.indexB = clipPoints2[i].id.cf.indexA;
        }
        ++pointCount;
      }
    }
    manifold->pointCount = pointCount;
  }
};
}  // namespace

void collideEdgeAndPolygon(Manifold* m, const Shape& edgeA, const Xf& xfA, const Shape& polyB, const Xf& xfB) {
  EPCollider c;
  c.collide(m, edgeA, xfA, polyB, xfB);
}

// ------------------------------------------------------------------ GJK (b2distance.d)
void DistanceProxy::set(const Shape& shape, int index) {
  switch (shape.type) {
    case kCircle: vertices = &shape.p; count = 1; radius = shape.radius; break;
    case kPolygon: vertices = shape.verts; count = shape.count; radius = shape.radius; break;
    case kChain: {
      buffer[0] = shape.chain[index];
      if (index + 1 < (int)shape.chain.size()) buffer[1] = shape.chain[index + 1];
      else buffer[1] = shape.chain[0];
      vertices = buffer; count = 2; radius = shape.radius;
    } break;
    case kEdge:
      buffer[0] = shape.v1; buffer[1] = shape.v2;  // the reference points at &m_vertex1 (v1,v2 adjacent)
      vertices = buffer; count = 2; radius = shape.radius;
      break;
  }
}
int DistanceProxy::support(V2 d) const {
  int bestIndex = 0;
  float bestValue = dot(vertices[0], d);
  for (int i = 1; i < count; ++i) {
    float value = dot(vertices[i], d);
    if (value > bestValue) { bestIndex = i; bestValue = value; }
  }
  return bestIndex;
}

namespace {
struct SimplexVertex { V2 wA, wB, w; float a = 0; int indexA = 0, indexB = 0; };
struct Simplex {
  SimplexVertex v[3];
  int count = 0;

  float metric() const {
    switch (count) {
      case 1: return 0.0f;
      case 2: return dist(v[0].w, v[1].w);
      case 3: return cross(v[1].w - v[0].w, v[2].w - v[0].w);
    }
    return 0.0f;
  }
  // b2distance.d:539-590
  void readCache(const SimplexCache* cache, const DistanceProxy* proxyA, const Xf& xfA, const DistanceProxy* proxyB, const Xf& xfB) {
    count = cache->count;
    for (int i = 0; i < count; ++i) {
      SimplexVertex* s = v + i;
      s->indexA = cache->indexA[i];
      s->indexB = cache->indexB[i];
      V2 wALocal = proxyA->vertex(s->indexA), wBLocal = proxyB->vertex(s->indexB);
      s->wA = mul(xfA, wALocal);
      s->wB = mul(xfB, wBLocal);
      s->w = s->wB - s->wA;
      s->a = 0.0f;
    }
    if (count > 1) {
      float metric1 = cache->metric, metric2 = metric();
      if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEpsilon) count = 0;
    }
    if (count == 0) {
      SimplexVertex* s = v + 0;
      s->indexA = 0; s->indexB = 0;
      V2 wALocal = proxyA->vertex(0), wBLocal = proxyB->vertex(0);
      s->wA = mul(xfA, wALocal);
      s->wB = mul(xfB, wBLocal);
      s->w = s->wB - s->wA;
      s->a = 1.0f;
      count = 1;
    }
  }
  void writeCache(SimplexCache* cache) const {
    cache->metric = metric();
    cache->count = (uint16_t)count;
    for (int i = 0; i < count; ++i) { cache->indexA[i] = (uint8_t)v[i].indexA; cache->indexB[i] = (uint8_t)v[i].indexB; }
  }
  V2 searchDirection() const {
    if (count == 1) return -v[0].w;
    V2 e12 = v[1].w - v[0].w;
    float sgn = cross(e12, -v[0].w);
    if (sgn > 0.0f) return cross(1.0f, e12);
    return cross(e12, 1.0f);
  }
  V2 closestPoint() const {
    switch (count) {
      case 1: return v[0].w;
      case 2: return v[0].a * v[0].w + v[1].a * v[1].w;
      default: return V2(0, 0);
    }
  }
  void witnessPoints(V2* pA, V2* pB) const {
    switch (count) {
      case 1: *pA = v[0].wA; *pB = v[0].wB; break;
      case 2:
        *pA = v[0].a * v[0].wA + v[1].a * v[1].wA;
        *pB = v[0].a * v[0].wB + v[1].a * v[1].wB;
        break;
      case 3:
        *pA = v[0].a * v[0].wA + v[1].a * v[1].wA + v[2].a * v[2].wA;
        *pB = *pA;
        break;
    }
  }
  // b2distance.d:389-423
  void solve2() {
    V2 w1 = v[0].w, w2 = v[1].w;
    V2 e12 = w2 - w1;
    float d12_2 = -dot(w1, e12);
    if (d12_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
    float d12_1 = dot(w2, e12);
    if (d12_1 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
    float inv_d12 = 1.0f / (d12_1 + d12_2);
    v[0].a = d12_1 * inv_d12;
    v[1].a = d12_2 * inv_d12;
    count = 2;
  }
  // b2distance.d:430-537
  void solve3() {
    V2 w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
    V2 e12 = w2 - w1;
    float w1e12 = dot(w1, e12), w2e12 = dot(w2, e12);
    float d12_1 = w2e12, d12_2 = -w1e12;
    V2 e13 = w3 - w1;
    float w1e13 = dot(w1, e13), w3e13 = dot(w3, e13);
    float d13_1 = w3e13, d13_2 = -w1e13;
    V2 e23 = w3 - w2;
    float w2e23 = dot(w2, e23), w3e23 = dot(w3, e23);
    float d23_1 = w3e23, d23_2 = -w2e23;
    float n123 = cross(e12, e13);
    float d123_1 = n123 * cross(w2, w3);
    float d123_2 = n123 * cross(w3, w1);
    float d123_3 = n123 * cross(w1, w2);
    if (d12_2 <= 0.0f && d13_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
    if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
      float inv = 1.0f / (d12_1 + d12_2);
      v[0].a = d12_1 * inv; v[1].a = d12_2 * inv; count = 2; return;
    }
    if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
      float inv = 1.0f / (d13_1 + d13_2);
      v[0].a = d13_1 * inv; v[2].a = d13_2 * inv; count = 2; v[1] = v[2]; return;
    }
    if (d12_1 <= 0.0f && d23_2 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
    if (d13_1 <= 0.0f && d23_1 <= 0.0f) { v[2].a = 1.0f; count = 1; v[0] = v[2]; return; }
    if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
      float inv = 1.0f / (d23_1 + d23_2);
      v[1].a = d23_1 * inv; v[2].a = d23_2 * inv; count = 2; v[0] = v[2]; return;
    }
    float inv = 1.0f / (d123_1 + d123_2 + d123_3);
    v[0].a = d123_1 * inv; v[1].a = d123_2 * inv; v[2].a = d123_3 * inv;
    count = 3;
  }
};
}  // namespace

// b2distance.d:185-347
void distance(DistanceOutput* output, SimplexCache* cache, const DistanceInput* input) {
  const DistanceProxy* proxyA = &input->proxyA;
  const DistanceProxy* proxyB = &input->proxyB;
  Xf transformA = input->transformA, transformB = input->transformB;
  Simplex simplex;
  simplex.readCache(cache, proxyA, transformA, proxyB, transformB);
  SimplexVertex* vertices = simplex.v;
  const int k_maxIters = 20;
  int saveA[3], saveB[3];
  int saveCount = 0;
  float distanceSqr1 = kMaxFloat, distanceSqr2 = distanceSqr1;
  int iter = 0;
  while (iter < k_maxIters) {
    saveCount = simplex.count;
    for (int i = 0; i < saveCount; ++i) { saveA[i] = vertices[i].indexA; saveB[i] = vertices[i].indexB; }
    switch (simplex.count) {
      case 1: break;
      case 2: simplex.solve2(); break;
      case 3: simplex.solve3(); break;
    }
    if (simplex.count == 3) break;
    V2 p = simplex.closestPoint();
    distanceSqr2 = p.len2();
    if (distanceSqr2 >= distanceSqr1) { /* reference: no-op (b2distance.d:255-258) */ }
    distanceSqr1 = distanceSqr2;
    V2 d = simplex.searchDirection();
    if (d.len2() < kEpsilon * kEpsilon) break;
    SimplexVertex* vertex = vertices + simplex.count;
    vertex->indexA = proxyA->support(mulT(transformA.q, -d));
    vertex->wA = mul(transformA, proxyA->vertex(vertex->indexA));
    vertex->indexB = proxyB->support(mulT(transformB.q, d));
    vertex->wB = mul(transformB, proxyB->vertex(vertex->indexB));
    vertex->w = vertex->wB - vertex->wA;
    ++iter;
    bool duplicate = false;
    for (int i = 0; i < saveCount; ++i) {
      if (vertex->indexA == saveA[i] && vertex->indexB == saveB[i]) { duplicate = true; break; }
    }
    if (duplicate) break;
    ++simplex.count;
  }
  simplex.witnessPoints(&output->pointA, &output->pointB);
  output->distance = dist(output->pointA, output->pointB);
  output->iterations = iter;
  simplex.writeCache(cache);
  if (input->useRadii) {
    float rA = proxyA->radius, rB = proxyB->radius;
    if (output->distance > rA + rB && output->distance > kEpsilon) {
      output->distance -= rA + rB;
      V2 normal = output->pointB - output->pointA;
      normal.normalize();
      output->pointA += rA * normal;
      output->pointB -= rB * normal;
    } else {
      V2 p = 0.5f * (output->pointA + output->pointB);
      output->pointA = p;
      output->pointB = p;
      output->distance = 0.0f;
    }
  }
}

bool testOverlap(const Shape& a, int ia, const Shape& b, int ib, const Xf& xfA, const Xf& xfB) {
  DistanceInput input;
  input.proxyA.set(a, ia);
  input.proxyB.set(b, ib);
  input.transformA = xfA;
  input.transformB = xfB;
  input.useRadii = true;
  SimplexCache cache;
  cache.count = 0;
  DistanceOutput output;
  distance(&output, &cache, &input);
  return output.distance < 10.0f * kEpsilon;
}

// ------------------------------------------------------------------ TOI (b2timeofimpact.d)
namespace {
enum SepType { kSepPoints, kSepFaceA, kSepFaceB };
struct SeparationFunction {
  const DistanceProxy* proxyA; const DistanceProxy* proxyB;
  Sweep sweepA, sweepB;
  int type;
  V2 localPoint, axis;

  // b2timeofimpact.d:324-404
  float initialize(const SimplexCache* cache, const DistanceProxy* pA, const Sweep& sA, const DistanceProxy* pB, const Sweep& sB, float t1) {
    proxyA = pA; proxyB = pB;
    int count = cache->count;
    sweepA = sA; sweepB = sB;
    Xf xfA, xfB;
    sweepA.getTransform(&xfA, t1);
    sweepB.getTransform(&xfB, t1);
    if (count == 1) {
      type = kSepPoints;
      V2 localPointA = proxyA->vertex(cache->indexA[0]);
      V2 localPointB = proxyB->vertex(cache->indexB[0]);
      V2 pointA = mul(xfA, localPointA), pointB = mul(xfB, localPointB);
      axis = pointB - pointA;
      return axis.normalize();
    } else if (cache->indexA[0] == cache->indexA[1]) {
      type = kSepFaceB;
      V2 localPointB1 = proxyB->vertex(cache->indexB[0]);
      V2 localPointB2 = proxyB->vertex(cache->indexB[1]);
      axis = cross(localPointB2 - localPointB1, 1.0f);
      axis.normalize();
      V2 normal = mul(xfB.q, axis);
      localPoint = 0.5f * (localPointB1 + localPointB2);
      V2 pointB = mul(xfB, localPoint);
      V2 localPointA = proxyA->vertex(cache->indexA[0]);
      V2 pointA = mul(xfA, localPointA);
      float s = dot(pointA - pointB, normal);
      if (s < 0.0f) { axis = -axis; s = -s; }
      return s;
    } else {
      type = kSepFaceA;
      V2 localPointA1 = proxyA->vertex(cache->indexA[0]);
      V2 localPointA2 = proxyA->vertex(cache->indexA[1]);
      axis = cross(localPointA2 - localPointA1, 1.0f);
      axis.normalize();
      V2 normal = mul(xfA.q, axis);
      localPoint = 0.5f * (localPointA1 + localPointA2);
      V2 pointA = mul(xfA, localPoint);
      V2 localPointB = proxyB->vertex(cache->indexB[0]);
      V2 pointB = mul(xfB, localPointB);
      float s = dot(pointB - pointA, normal);
      if (s < 0.0f) { axis = -axis; s = -s; }
      return s;
    }
  }
  // b2timeofimpact.d:407-472
  float findMinSeparation(int* indexA, int* indexB, float t) const {
    Xf xfA, xfB;
    sweepA.getTransform(&xfA, t);
    sweepB.getTransform(&xfB, t);
    switch (type) {
      case kSepPoints: {
        V2 axisA = mulT(xfA.q, axis), axisB = mulT(xfB.q, -axis);
        *indexA = proxyA->support(axisA);
        *indexB = proxyB->support(axisB);
        V2 pointA = mul(xfA, proxyA->vertex(*indexA));
        V2 pointB = mul(xfB, proxyB->vertex(*indexB));
        return dot(pointB - pointA, axis);
      }
      case kSepFaceA: {
        V2 normal = mul(xfA.q, axis);
        V2 pointA = mul(xfA, localPoint);
        V2 axisB = mulT(xfB.q, -normal);
        *indexA = -1;
        *indexB = proxyB->support(axisB);
        V2 pointB = mul(xfB, proxyB->vertex(*indexB));
        return dot(pointB - pointA, normal);
      }
      default: {
        V2 normal = mul(xfB.q, axis);
        V2 pointB = mul(xfB, localPoint);
        V2 axisA = mulT(xfA.q, -normal);
        *indexB = -1;
        *indexA = proxyA->support(axisA);
        V2 pointA = mul(xfA, proxyA->vertex(*indexA));
        return dot(pointA - pointB, normal);
      }
    }
  }
  // b2timeofimpact.d:475-522
  float evaluate(int indexA, int indexB, float t) const {
    Xf xfA, xfB;
    sweepA.getTransform(&xfA, t);
    sweepB.getTransform(&xfB, t);
    switch (type) {
      case kSepPoints: {
        V2 pointA = mul(xfA, proxyA->vertex(indexA));
        V2 pointB = mul(xfB, proxyB->vertex(indexB));
        return dot(pointB - pointA, axis);
      }
      case kSepFaceA: {
        V2 normal = mul(xfA.q, axis);
        V2 pointA = mul(xfA, localPoint);
        V2 pointB = mul(xfB, proxyB->vertex(indexB));
        return dot(pointB - pointA, normal);
      }
      default: {
        V2 normal = mul(xfB.q, axis);
        V2 pointB = mul(xfB, localPoint);
        V2 pointA = mul(xfA, proxyA->vertex(indexA));
        return dot(pointA - pointB, normal);
      }
    }
  }
};
}  // namespace

// b2timeofimpact.d:67-302
void timeOfImpact(TOIOutput* output, const TOIInput* input) {
  output->state = kToiUnknown;
  output->t = input->tMax;
  const DistanceProxy* proxyA = &input->proxyA;
  const DistanceProxy* proxyB = &input->proxyB;
  Sweep sweepA = input->sweepA, sweepB = input->sweepB;
  sweepA.normalize();
  sweepB.normalize();
  float tMax = input->tMax;
  float totalRadius = proxyA->radius + proxyB->radius;
  float target = maxT(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
  float tolerance = 0.25f * kLinearSlop;
  float t1 = 0.0f;
  const int k_maxIterations = 20;
  int iter = 0;
  SimplexCache cache;
  cache.count = 0;
  DistanceInput distanceInput;
  distanceInput.proxyA = input->proxyA;
  distanceInput.proxyB = input->proxyB;
  distanceInput.useRadii = false;
  for (;;) {
    Xf xfA, xfB;
    sweepA.getTransform(&xfA, t1);
    sweepB.getTransform(&xfB, t1);
    distanceInput.transformA = xfA;
    distanceInput.transformB = xfB;
    DistanceOutput distanceOutput;
    distance(&distanceOutput, &cache, &distanceInput);
    if (distanceOutput.distance <= 0.0f) { output->state = kToiOverlapped; output->t = 0.0f; break; }
    if (distanceOutput.distance < target + tolerance) { output->state = kToiTouching; output->t = t1; break; }
    SeparationFunction fcn;
    fcn.initialize(&cache, proxyA, sweepA, proxyB, sweepB, t1);
    bool done = false;
    float t2 = tMax;
    int pushBackIter = 0;
    for (;;) {
      int indexA, indexB;
      float s2 = fcn.findMinSeparation(&indexA, &indexB, t2);
      if (s2 > target + tolerance) { output->state = kToiSeparated; output->t = tMax; done = true; break; }
      if (s2 > target - tolerance) { t1 = t2; break; }
      float s1 = fcn.evaluate(indexA, indexB, t1);
      if (s1 < target - tolerance) { output->state = kToiFailed; output->t = t1; done = true; break; }
      if (s1 <= target + tolerance) { output->state = kToiTouching; output->t = t1; done = true; break; }
      int rootIterCount = 0;
      float a1 = t1, a2 = t2;
      for (;;) {
        float t;
        if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
        else t = 0.5f * (a1 + a2);
        ++rootIterCount;
        float s = fcn.evaluate(indexA, indexB, t);
        if (absT(s - target) < tolerance) { t2 = t; break; }
        if (s > target) { a1 = t; s1 = s; }
        else { a2 = t; s2 = s; }
        if (rootIterCount == 50) break;
      }
      ++pushBackIter;
      if (pushBackIter == kMaxPolygonVertices) break;
    }
    ++iter;
    if (done) break;
    if (iter == k_maxIterations) { output->state = kToiFailed; output->t = t1; break; }
  }
}

static bool rayCastEdge(float* fraction, V2* normalOut, V2 P1, V2 P2, float maxFraction, const Xf& xf, V2 v1, V2 v2) {
  V2 p1 = mulT(xf.q, P1 - xf.p);
  V2 p2 = mulT(xf.q, P2 - xf.p);
  V2 d = p2 - p1;
  V2 e = v2 - v1;
  V2 normal(e.y, -e.x);
  normal.normalize();
  float numerator = dot(normal, v1 - p1);
  float denominator = dot(normal, d);
  if (denominator == 0.0f) return false;
  float t = numerator / denominator;
  if (t < 0.0f || maxFraction < t) return false;
  V2 q = p1 + t * d;
  V2 r = v2 - v1;
  float rr = dot(r, r);
  if (rr == 0.0f) return false;
  float s = dot(q - v1, r) / rr;
  if (s < 0.0f || 1.0f < s) return false;
  *fraction = t;
  *normalOut = numerator > 0.0f ? -mul(xf.q, normal) : mul(xf.q, normal);
  return true;
}

bool Shape::testPoint(const Xf& xf, V2 pt) const {
  if (type == kCircle) {
    V2 center = xf.p + mul(xf.q, p);
    V2 d = pt - center;
    return dot(d, d) <= radius * radius;
  }
  if (type != kPolygon) return false;
  V2 pLocal = mulT(xf.q, pt - xf.p);
  for (int i = 0; i < count; ++i) {
    float d = dot(normals[i], pLocal - verts[i]);
    if (d > 0.0f) return false;
  }
  return true;
}

bool Shape::rayCast(float* fraction, V2* normalOut, V2 P1, V2 P2, float maxFraction, const Xf& xf, int child) const {
  if (type == kCircle) {
    V2 position = xf.p + mul(xf.q, p);
    V2 s = P1 - position;
    float b = dot(s, s) - radius * radius;
    V2 r = P2 - P1;
    float c = dot(s, r);
    float rr = dot(r, r);
    float sigma = c * c - rr * b;
    if (sigma < 0.0f || rr < kEpsilon) return false;
    float a = -(c + sqrtf(sigma));
    if (0.0f <= a && a <= maxFraction * rr) {
      a /= rr;
      *fraction = a;
      *normalOut = s + a * r;
      normalOut->normalize();
      return true;
    }
    return false;
  }
  if (type == kEdge) return rayCastEdge(fraction, normalOut, P1, P2, maxFraction, xf, v1, v2);
  if (type == kChain) {
    int i1 = child, i2 = child + 1;
    if (i2 == (int)chain.size()) i2 = 0;
    return rayCastEdge(fraction, normalOut, P1, P2, maxFraction, xf, chain[i1], chain[i2]);
  }
  // polygon
  V2 p1 = mulT(xf.q, P1 - xf.p);
  V2 p2 = mulT(xf.q, P2 - xf.p);
  V2 d = p2 - p1;
  float lower = 0.0f, upper = maxFraction;
  int index = -1;
  for (int i = 0; i < count; ++i) {
    float numerator = dot(normals[i], verts[i] - p1);
    float denominator = dot(normals[i], d);
    if (denominator == 0.0f) {
      if (numerator < 0.0f) return false;
    } else {
      if (denominator < 0.0f && numerator < lower * denominator) { lower = numerator / denominator; index = i; }
      else if (denominator > 0.0f && numerator < upper * denominator) upper = numerator / denominator;
    }
    if (upper < lower) return false;
  }
  if (index >= 0) { *fraction = lower; *normalOut = mul(xf.q, normals[index]); return true; }
  return false;
}

}  // namespace orc
