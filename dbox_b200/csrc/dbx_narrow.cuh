// dbx_narrow.cuh — per-shape-pair narrowphase, GJK distance and time of impact as device functions.
//
// Replaces (relative to /root/reference/src/dbox/):
//   collision/b2collidepolygon.d:29-254   b2FindMaxSeparation / b2FindIncidentEdge / b2CollidePolygons
//   collision/b2collidecircle.d:29-164    b2CollideCircles / b2CollidePolygonAndCircle
//   collision/b2collideedge.d:31-724      b2CollideEdgeAndCircle / b2EPCollider (edge & chain-child vs polygon)
//   collision/b2collision.d:413-446       b2ClipSegmentToLine
//   collision/b2distance.d:185-347        b2Distance (GJK with simplex cache)
//   collision/b2timeofimpact.d:67-530     b2TimeOfImpact / b2SeparationFunction
// Geometry is read straight from the de-duplicated DShape pool in global memory (read-only path, L1-resident for
// piles of identical boxes); nothing here touches world state, so one thread evaluates one contact.
#pragma once
#include "dbx_math.cuh"

namespace dbx {

enum { SH_CIRCLE = 0, SH_EDGE = 1, SH_POLYGON = 2, SH_CHAIN = 3 };
enum { SHF_HAS_V0 = 1, SHF_HAS_V3 = 2, SHF_CHAIN_CHILD = 4 };
enum { MAN_CIRCLES = 0, MAN_FACE_A = 1, MAN_FACE_B = 2 };
enum { FEAT_VERTEX = 0, FEAT_FACE = 1 };

// One geometry record per distinct (shape, child): polygons as given, circles, edges; chain children are expanded to
// edges with ghost vertices at fixture creation (b2chainshape.d:162-192) and flagged so their AABB skips the radius.
struct __align__(16) DShape {
  int type;
  float radius;
  int count;
  int flags;
  v2 c;     // circle centre | polygon centroid
  v2 _pad;
  v2 v[8];  // polygon vertices | edge: v[0..3] = vertex0, vertex1, vertex2, vertex3
  v2 n[8];  // polygon normals
};
static_assert(sizeof(DShape) == 160, "DShape layout");

struct Manifold {
  v2 localNormal, localPoint;
  v2 lp[2];
  uint32_t key[2];
  int type, pointCount;
};

// b2ContactID.key = indexA | indexB<<8 | typeA<<16 | typeB<<24 (b2collision.d:38-60, little-endian union)
DBX_HD uint32_t feat_key(int indexA, int indexB, int typeA, int typeB) {
  return (uint32_t)(indexA & 0xFF) | ((uint32_t)(indexB & 0xFF) << 8) | ((uint32_t)typeA << 16) | ((uint32_t)typeB << 24);
}
DBX_HD uint32_t feat_flip(uint32_t k) {  // swap (indexA,typeA) with (indexB,typeB)
  return ((k & 0xFF) << 8) | ((k >> 8) & 0xFF) | (((k >> 16) & 0xFF) << 24) | (((k >> 24) & 0xFF) << 16);
}

struct ClipV { v2 v; uint32_t key; };

// b2collision.d:413-446
DBX_HD int clip_segment(ClipV out[2], const ClipV in[2], v2 normal, float offset, int vertexIndexA) {
  int numOut = 0;
  float d0 = dot(normal, in[0].v) - offset;
  float d1 = dot(normal, in[1].v) - offset;
  if (d0 <= 0.0f) out[numOut++] = in[0];
  if (d1 <= 0.0f) out[numOut++] = in[1];
  if (d0 * d1 < 0.0f) {
    float interp = d0 / (d0 - d1);
    out[numOut].v = in[0].v + interp * (in[1].v - in[0].v);
    out[numOut].key = feat_key(vertexIndexA, (in[0].key >> 8) & 0xFF, FEAT_VERTEX, FEAT_FACE);
    ++numOut;
  }
  return numOut;
}

// b2collidecircle.d:29-56
DBX_HD void collide_circles(Manifold& m, const DShape* A, Xf xfA, const DShape* B, Xf xfB) {
  m.pointCount = 0;
  v2 pA = mul(xfA, A->c), pB = mul(xfB, B->c);
  v2 d = pB - pA;
  float distSqr = dot(d, d);
  float radius = A->radius + B->radius;
  if (distSqr > radius * radius) return;
  m.type = MAN_CIRCLES;
  m.localPoint = A->c;
  m.localNormal = V(0.0f, 0.0f);
  m.pointCount = 1;
  m.lp[0] = B->c;
  m.key[0] = 0;
}

// b2collidecircle.d:59-164
DBX_HD void collide_polygon_circle(Manifold& m, const DShape* A, Xf xfA, const DShape* B, Xf xfB) {
  m.pointCount = 0;
  v2 c = mul(xfB, B->c);
  v2 cLocal = mulT(xfA, c);
  int normalIndex = 0;
  float separation = -kMaxFloat;
  float radius = A->radius + B->radius;
  int vertexCount = A->count;
  for (int i = 0; i < vertexCount; ++i) {
    float s = dot(A->n[i], cLocal - A->v[i]);
    if (s > radius) return;
    if (s > separation) { separation = s; normalIndex = i; }
  }
  int i1 = normalIndex, i2 = i1 + 1 < vertexCount ? i1 + 1 : 0;
  v2 v1 = A->v[i1], v2_ = A->v[i2];
  m.type = MAN_FACE_A;
  m.lp[0] = B->c;
  m.key[0] = 0;
  if (separation < kEpsilon) {
    m.pointCount = 1;
    m.localNormal = A->n[normalIndex];
    m.localPoint = 0.5f * (v1 + v2_);
    return;
  }
  float u1 = dot(cLocal - v1, v2_ - v1);
  float u2 = dot(cLocal - v2_, v1 - v2_);
  if (u1 <= 0.0f) {
    if (dist2(cLocal, v1) > radius * radius) return;
    m.pointCount = 1;
    m.localNormal = cLocal - v1;
    normalize(m.localNormal);
    m.localPoint = v1;
  } else if (u2 <= 0.0f) {
    if (dist2(cLocal, v2_) > radius * radius) return;
    m.pointCount = 1;
    m.localNormal = cLocal - v2_;
    normalize(m.localNormal);
    m.localPoint = v2_;
  } else {
    v2 faceCenter = 0.5f * (v1 + v2_);
    float separation2 = dot(cLocal - faceCenter, A->n[i1]);
    if (separation2 > radius) return;
    m.pointCount = 1;
    m.localNormal = A->n[i1];
    m.localPoint = faceCenter;
  }
}

// b2collidepolygon.d:29-71
DBX_HD float find_max_separation(int* edgeIndex, const DShape* p1, Xf xf1, const DShape* p2, Xf xf2) {
  int count1 = p1->count, count2 = p2->count;
  Xf xf = mulT(xf2, xf1);
  int bestIndex = 0;
  float maxSeparation = -kMaxFloat;
  for (int i = 0; i < count1; ++i) {
    v2 n = mul(xf.q, p1->n[i]);
    v2 v1 = mul(xf, p1->v[i]);
    float si = kMaxFloat;
    for (int j = 0; j < count2; ++j) {
      float sij = dot(n, p2->v[j] - v1);
      if (sij < si) si = sij;
    }
    if (si > maxSeparation) { maxSeparation = si; bestIndex = i; }
  }
  *edgeIndex = bestIndex;
  return maxSeparation;
}

// b2collidepolygon.d:127-254 (incl. b2FindIncidentEdge :73-118)
DBX_HD void collide_polygons(Manifold& m, const DShape* A, Xf xfA, const DShape* B, Xf xfB) {
  m.pointCount = 0;
  float totalRadius = A->radius + B->radius;
  int edgeA = 0;
  float separationA = find_max_separation(&edgeA, A, xfA, B, xfB);
  if (separationA > totalRadius) return;
  int edgeB = 0;
  float separationB = find_max_separation(&edgeB, B, xfB, A, xfA);
  if (separationB > totalRadius) return;
  const DShape* poly1; const DShape* poly2;
  Xf xf1, xf2;
  int edge1;
  bool flip;
  const float k_tol = 0.1f * kLinearSlop;
  if (separationB > separationA + k_tol) { poly1 = B; poly2 = A; xf1 = xfB; xf2 = xfA; edge1 = edgeB; m.type = MAN_FACE_B; flip = true; }
  else { poly1 = A; poly2 = B; xf1 = xfA; xf2 = xfB; edge1 = edgeA; m.type = MAN_FACE_A; flip = false; }
  // incident edge
  ClipV incident[2];
  {
    int count2 = poly2->count;
    v2 normal1 = mulT(xf2.q, mul(xf1.q, poly1->n[edge1]));
    int index = 0;
    float minDot = kMaxFloat;
    for (int i = 0; i < count2; ++i) {
      float d = dot(normal1, poly2->n[i]);
      if (d < minDot) { minDot = d; index = i; }
    }
    int i1 = index, i2 = i1 + 1 < count2 ? i1 + 1 : 0;
    incident[0].v = mul(xf2, poly2->v[i1]); incident[0].key = feat_key(edge1, i1, FEAT_FACE, FEAT_VERTEX);
    incident[1].v = mul(xf2, poly2->v[i2]); incident[1].key = feat_key(edge1, i2, FEAT_FACE, FEAT_VERTEX);
  }
  int count1 = poly1->count;
  int iv1 = edge1, iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
  v2 v11 = poly1->v[iv1], v12 = poly1->v[iv2];
  v2 localTangent = v12 - v11;
  normalize(localTangent);
  v2 localNormal = cross(localTangent, 1.0f);
  v2 planePoint = 0.5f * (v11 + v12);
  v2 tangent = mul(xf1.q, localTangent);
  v2 normal = cross(tangent, 1.0f);
  v11 = mul(xf1, v11);
  v12 = mul(xf1, v12);
  float frontOffset = dot(normal, v11);
  float sideOffset1 = -dot(tangent, v11) + totalRadius;
  float sideOffset2 = dot(tangent, v12) + totalRadius;
  ClipV clip1[2], clip2[2];
  int np = clip_segment(clip1, incident, -tangent, sideOffset1, iv1);
  if (np < 2) return;
  np = clip_segment(clip2, clip1, tangent, sideOffset2, iv2);
  if (np < 2) return;
  m.localNormal = localNormal;
  m.localPoint = planePoint;
  int pointCount = 0;
  for (int i = 0; i < kMaxManifoldPoints; ++i) {
    float separation = dot(normal, clip2[i].v) - frontOffset;
    if (separation <= totalRadius) {
      m.lp[pointCount] = mulT(xf2, clip2[i].v);
      m.key[pointCount] = flip ? feat_flip(clip2[i].key) : clip2[i].key;
      ++pointCount;
    }
  }
  m.pointCount = pointCount;
}

// b2collideedge.d:31-160
DBX_HD void collide_edge_circle(Manifold& m, const DShape* E, Xf xfA, const DShape* Cc, Xf xfB) {
  m.pointCount = 0;
  v2 Q = mulT(xfA, mul(xfB, Cc->c));
  v2 A = E->v[1], B = E->v[2];
  v2 e = B - A;
  float u = dot(e, B - Q);
  float v = dot(e, Q - A);
  float radius = E->radius + Cc->radius;
  if (v <= 0.0f) {
    v2 P = A, d = Q - P;
    float dd = dot(d, d);
    if (dd > radius * radius) return;
    if (E->flags & SHF_HAS_V0) {
      v2 A1 = E->v[0], B1 = A, e1 = B1 - A1;
      float u1 = dot(e1, B1 - Q);
      if (u1 > 0.0f) return;
    }
    m.pointCount = 1; m.type = MAN_CIRCLES; m.localNormal = V(0.0f, 0.0f); m.localPoint = P;
    m.key[0] = feat_key(0, 0, FEAT_VERTEX, FEAT_VERTEX); m.lp[0] = Cc->c;
    return;
  }
  if (u <= 0.0f) {
    v2 P = B, d = Q - P;
    float dd = dot(d, d);
    if (dd > radius * radius) return;
    if (E->flags & SHF_HAS_V3) {
      v2 B2 = E->v[3], A2 = B, e2 = B2 - A2;
      float v2d = dot(e2, Q - A2);
      if (v2d > 0.0f) return;
    }
    m.pointCount = 1; m.type = MAN_CIRCLES; m.localNormal = V(0.0f, 0.0f); m.localPoint = P;
    m.key[0] = feat_key(1, 0, FEAT_VERTEX, FEAT_VERTEX); m.lp[0] = Cc->c;
    return;
  }
  float den = dot(e, e);
  v2 P = (1.0f / den) * (u * A + v * B);
  v2 d = Q - P;
  float dd = dot(d, d);
  if (dd > radius * radius) return;
  v2 n = V(-e.y, e.x);
  if (dot(n, Q - A) < 0.0f) n = V(-n.x, -n.y);
  normalize(n);
  m.pointCount = 1; m.type = MAN_FACE_A; m.localNormal = n; m.localPoint = A;
  m.key[0] = feat_key(0, 0, FEAT_FACE, FEAT_VERTEX); m.lp[0] = Cc->c;
}

// b2collideedge.d:217-695 (b2EPCollider.Collide + ComputeEdgeSeparation + ComputePolygonSeparation)
DBX_HD void collide_edge_polygon(Manifold& m, const DShape* E, Xf xfA, const DShape* Pg, Xf xfB) {
  Xf xf = mulT(xfA, xfB);
  v2 centroidB = mul(xf, Pg->c);
  v2 v0 = E->v[0], v1 = E->v[1], v2_ = E->v[2], v3 = E->v[3];
  bool hasVertex0 = (E->flags & SHF_HAS_V0) != 0, hasVertex3 = (E->flags & SHF_HAS_V3) != 0;
  v2 edge1 = v2_ - v1;
  normalize(edge1);
  v2 normal1 = V(edge1.y, -edge1.x);
  float offset1 = dot(normal1, centroidB - v1);
  float offset0 = 0.0f, offset2 = 0.0f;
  bool convex1 = false, convex2 = false;
  v2 normal0 = V(0.0f, 0.0f), normal2 = V(0.0f, 0.0f);
  if (hasVertex0) {
    v2 edge0 = v1 - v0;
    normalize(edge0);
    normal0 = V(edge0.y, -edge0.x);
    convex1 = cross(edge0, edge1) >= 0.0f;
    offset0 = dot(normal0, centroidB - v0);
  }
  if (hasVertex3) {
    v2 edge2 = v3 - v2_;
    normalize(edge2);
    normal2 = V(edge2.y, -edge2.x);
    convex2 = cross(edge1, edge2) > 0.0f;
    offset2 = dot(normal2, centroidB - v2_);
  }
  bool front;
  v2 normal, lowerLimit, upperLimit;
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
  // polygon B in the edge's frame
  v2 pv[kMaxPolygonVertices], pn[kMaxPolygonVertices];
  int pcount = Pg->count;
  for (int i = 0; i < pcount; ++i) { pv[i] = mul(xf, Pg->v[i]); pn[i] = mul(xf.q, Pg->n[i]); }
  const float radius = 2.0f * kPolygonRadius;
  m.pointCount = 0;
  // ComputeEdgeSeparation
  float edgeSep = FLT_MAX;
  for (int i = 0; i < pcount; ++i) {
    float s = dot(normal, pv[i] - v1);
    if (s < edgeSep) edgeSep = s;
  }
  if (edgeSep > radius) return;
  // ComputePolygonSeparation
  int polyType = 0 /*unknown*/, polyIndex = -1;
  float polySep = -FLT_MAX;
  {
    v2 perp = V(-normal.y, normal.x);
    for (int i = 0; i < pcount; ++i) {
      v2 n = -pn[i];
      float s1 = dot(n, pv[i] - v1);
      float s2 = dot(n, pv[i] - v2_);
      float s = fminr(s1, s2);
      if (s > radius) { polyType = 2; polyIndex = i; polySep = s; break; }
      if (dot(n, perp) >= 0.0f) {
        if (dot(n - upperLimit, normal) < -kAngularSlop) continue;
      } else {
        if (dot(n - lowerLimit, normal) < -kAngularSlop) continue;
      }
      if (s > polySep) { polyType = 2; polyIndex = i; polySep = s; }
    }
  }
  if (polyType != 0 && polySep > radius) return;
  const float k_relativeTol = 0.98f, k_absoluteTol = 0.001f;
  bool primaryIsEdge;
  if (polyType == 0) primaryIsEdge = true;
  else if (polySep > k_relativeTol * edgeSep + k_absoluteTol) primaryIsEdge = false;
  else primaryIsEdge = true;

  ClipV ie[2];
  int rf_i1, rf_i2;
  v2 rf_v1, rf_v2, rf_normal;
  if (primaryIsEdge) {
    m.type = MAN_FACE_A;
    int bestIndex = 0;
    float bestValue = dot(normal, pn[0]);
    for (int i = 1; i < pcount; ++i) {
      float value = dot(normal, pn[i]);
      if (value < bestValue) { bestValue = value; bestIndex = i; }
    }
    int i1 = bestIndex, i2 = i1 + 1 < pcount ? i1 + 1 : 0;
    ie[0].v = pv[i1]; ie[0].key = feat_key(0, i1, FEAT_FACE, FEAT_VERTEX);
    ie[1].v = pv[i2]; ie[1].key = feat_key(0, i2, FEAT_FACE, FEAT_VERTEX);
    if (front) { rf_i1 = 0; rf_i2 = 1; rf_v1 = v1; rf_v2 = v2_; rf_normal = normal1; }
    else { rf_i1 = 1; rf_i2 = 0; rf_v1 = v2_; rf_v2 = v1; rf_normal = -normal1; }
  } else {
    m.type = MAN_FACE_B;
    ie[0].v = v1; ie[0].key = feat_key(0, polyIndex, FEAT_VERTEX, FEAT_FACE);
    ie[1].v = v2_; ie[1].key = feat_key(0, polyIndex, FEAT_VERTEX, FEAT_FACE);
    rf_i1 = polyIndex;
    rf_i2 = rf_i1 + 1 < pcount ? rf_i1 + 1 : 0;
    rf_v1 = pv[rf_i1]; rf_v2 = pv[rf_i2]; rf_normal = pn[rf_i1];
  }
  v2 sideNormal1 = V(rf_normal.y, -rf_normal.x);
  v2 sideNormal2 = -sideNormal1;
  float sideOffset1 = dot(sideNormal1, rf_v1);
  float sideOffset2 = dot(sideNormal2, rf_v2);
  ClipV clip1[2], clip2[2];
  int np = clip_segment(clip1, ie, sideNormal1, sideOffset1, rf_i1);
  if (np < kMaxManifoldPoints) return;
  np = clip_segment(clip2, clip1, sideNormal2, sideOffset2, rf_i2);
  if (np < kMaxManifoldPoints) return;
  if (primaryIsEdge) { m.localNormal = rf_normal; m.localPoint = rf_v1; }
  else { m.localNormal = Pg->n[rf_i1]; m.localPoint = Pg->v[rf_i1]; }
  int pointCount = 0;
  for (int i = 0; i < kMaxManifoldPoints; ++i) {
    float separation = dot(rf_normal, clip2[i].v - rf_v1);
    if (separation <= radius) {
      if (primaryIsEdge) { m.lp[pointCount] = mulT(xf, clip2[i].v); m.key[pointCount] = clip2[i].key; }
      else { m.lp[pointCount] = clip2[i].v; m.key[pointCount] = feat_flip(clip2[i].key); }
      ++pointCount;
    }
  }
  m.pointCount = pointCount;
}

// Dispatch of the 7 contact classes (dynamics/contacts/b2*contact.d line 56 each); typeA is already the primary type.
DBX_HD void collide_dispatch(Manifold& m, const DShape* A, Xf xfA, const DShape* B, Xf xfB) {
  int tA = A->type, tB = B->type;
  if (tA == SH_POLYGON && tB == SH_POLYGON) collide_polygons(m, A, xfA, B, xfB);
  else if (tA == SH_EDGE && tB == SH_POLYGON) collide_edge_polygon(m, A, xfA, B, xfB);
  else if (tA == SH_POLYGON && tB == SH_CIRCLE) collide_polygon_circle(m, A, xfA, B, xfB);
  else if (tA == SH_CIRCLE && tB == SH_CIRCLE) collide_circles(m, A, xfA, B, xfB);
  else if (tA == SH_EDGE && tB == SH_CIRCLE) collide_edge_circle(m, A, xfA, B, xfB);
  else m.pointCount = 0;
}

// Shape.ComputeAABB: b2circleshape.d:110-117, b2edgeshape.d:163-176, b2polygonshape.d:356-373, b2chainshape.d:226-243
DBX_HD Box shape_aabb(const DShape* s, Xf xf) {
  Box b;
  if (s->type == SH_CIRCLE) {
    v2 p = xf.p + mul(xf.q, s->c);
    b.lo = V(p.x - s->radius, p.y - s->radius);
    b.hi = V(p.x + s->radius, p.y + s->radius);
  } else if (s->type == SH_EDGE) {
    v2 v1 = mul(xf, s->v[1]), v2_ = mul(xf, s->v[2]);
    v2 lower = vmin(v1, v2_), upper = vmax(v1, v2_);
    if (s->flags & SHF_CHAIN_CHILD) { b.lo = lower; b.hi = upper; }           // b2chainshape.d:241-242: no radius padding
    else { v2 r = V(s->radius, s->radius); b.lo = lower - r; b.hi = upper + r; }
  } else {
    v2 lower = mul(xf, s->v[0]);
    v2 upper = lower;
    for (int i = 1; i < s->count; ++i) { v2 v = mul(xf, s->v[i]); lower = vmin(lower, v); upper = vmax(upper, v); }
    v2 r = V(s->radius, s->radius);
    b.lo = lower - r; b.hi = upper + r;
  }
  return b;
}


// ------------------------------------------------------------------------------------------------ GJK
struct DProxy { const v2* verts; int count; float radius; };
// b2DistanceProxy.Set (b2distance.d:34-90); chain children were expanded to edges, so only three cases remain
DBX_HD DProxy make_proxy(const DShape* s) {
  DProxy p;
  p.radius = s->radius;
  if (s->type == SH_CIRCLE) { p.verts = &s->c; p.count = 1; }
  else if (s->type == SH_POLYGON) { p.verts = s->v; p.count = s->count; }
  else { p.verts = &s->v[1]; p.count = 2; }
  return p;
}
DBX_HD int proxy_support(const DProxy& p, v2 d) {
  int best = 0;
  float bestValue = dot(p.verts[0], d);
  for (int i = 1; i < p.count; ++i) {
    float value = dot(p.verts[i], d);
    if (value > bestValue) { best = i; bestValue = value; }
  }
  return best;
}
struct SimplexCache { float metric; int count; uint8_t indexA[3], indexB[3]; };
struct SimplexVertex { v2 wA, wB, w; float a; int indexA, indexB; };
struct DistanceOutput { v2 pointA, pointB; float distance; int iterations; };

DBX_HD float simplex_metric(const SimplexVertex* v, int count) {
  if (count == 2) return dist(v[0].w, v[1].w);
  if (count == 3) return cross(v[1].w - v[0].w, v[2].w - v[0].w);
  return 0.0f;
}

// b2Distance (b2distance.d:185-347) incl. ReadCache/WriteCache/Solve2/Solve3
DBX_HD void gjk_distance(DistanceOutput& out, SimplexCache& cache, const DProxy& pA, Xf xfA, const DProxy& pB, Xf xfB, bool useRadii) {
  SimplexVertex v[3];
  int count = cache.count;
  for (int i = 0; i < count; ++i) {
    v[i].indexA = cache.indexA[i]; v[i].indexB = cache.indexB[i];
    v[i].wA = mul(xfA, pA.verts[v[i].indexA]);
    v[i].wB = mul(xfB, pB.verts[v[i].indexB]);
    v[i].w = v[i].wB - v[i].wA;
    v[i].a = 0.0f;
  }
  if (count > 1) {
    float metric1 = cache.metric, metric2 = simplex_metric(v, count);
    if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEpsilon) count = 0;
  }
  if (count == 0) {
    v[0].indexA = 0; v[0].indexB = 0;
    v[0].wA = mul(xfA, pA.verts[0]);
    v[0].wB = mul(xfB, pB.verts[0]);
    v[0].w = v[0].wB - v[0].wA;
    v[0].a = 1.0f;
    count = 1;
  }
  const int k_maxIters = 20;
  int saveA[3], saveB[3];
  int iter = 0;
  while (iter < k_maxIters) {
    int saveCount = count;
    for (int i = 0; i < saveCount; ++i) { saveA[i] = v[i].indexA; saveB[i] = v[i].indexB; }
    if (count == 2) {
      // Solve2 (b2distance.d:389-423)
      v2 w1 = v[0].w, w2 = v[1].w, e12 = w2 - w1;
      float d12_2 = -dot(w1, e12);
      if (d12_2 <= 0.0f) { v[0].a = 1.0f; count = 1; }
      else {
        float d12_1 = dot(w2, e12);
        if (d12_1 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; }
        else { float inv = 1.0f / (d12_1 + d12_2); v[0].a = d12_1 * inv; v[1].a = d12_2 * inv; count = 2; }
      }
    } else if (count == 3) {
      // Solve3 (b2distance.d:430-537)
      v2 w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
      v2 e12 = w2 - w1;
      float d12_1 = dot(w2, e12), d12_2 = -dot(w1, e12);
      v2 e13 = w3 - w1;
      float d13_1 = dot(w3, e13), d13_2 = -dot(w1, e13);
      v2 e23 = w3 - w2;
      float d23_1 = dot(w3, e23), d23_2 = -dot(w2, e23);
      float n123 = cross(e12, e13);
      float d123_1 = n123 * cross(w2, w3), d123_2 = n123 * cross(w3, w1), d123_3 = n123 * cross(w1, w2);
      if (d12_2 <= 0.0f && d13_2 <= 0.0f) { v[0].a = 1.0f; count = 1; }
      else if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) { float inv = 1.0f / (d12_1 + d12_2); v[0].a = d12_1 * inv; v[1].a = d12_2 * inv; count = 2; }
      else if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) { float inv = 1.0f / (d13_1 + d13_2); v[0].a = d13_1 * inv; v[2].a = d13_2 * inv; count = 2; v[1] = v[2]; }
      else if (d12_1 <= 0.0f && d23_2 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; }
      else if (d13_1 <= 0.0f && d23_1 <= 0.0f) { v[2].a = 1.0f; count = 1; v[0] = v[2]; }
      else if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) { float inv = 1.0f / (d23_1 + d23_2); v[1].a = d23_1 * inv; v[2].a = d23_2 * inv; count = 2; v[0] = v[2]; }
      else { float inv = 1.0f / (d123_1 + d123_2 + d123_3); v[0].a = d123_1 * inv; v[1].a = d123_2 * inv; v[2].a = d123_3 * inv; count = 3; }
    }
    if (count == 3) break;
    // GetSearchDirection (b2distance.d:605-632)
    v2 d;
    if (count == 1) d = -v[0].w;
    else {
      v2 e12 = v[1].w - v[0].w;
      float sgn = cross(e12, -v[0].w);
      d = sgn > 0.0f ? cross(1.0f, e12) : cross(e12, 1.0f);
    }
    if (len2(d) < kEpsilon * kEpsilon) break;
    SimplexVertex* nv = v + count;
    nv->indexA = proxy_support(pA, mulT(xfA.q, -d));
    nv->wA = mul(xfA, pA.verts[nv->indexA]);
    nv->indexB = proxy_support(pB, mulT(xfB.q, d));
    nv->wB = mul(xfB, pB.verts[nv->indexB]);
    nv->w = nv->wB - nv->wA;
    ++iter;
    bool duplicate = false;
    for (int i = 0; i < saveCount; ++i) {
      if (nv->indexA == saveA[i] && nv->indexB == saveB[i]) { duplicate = true; break; }
    }
    if (duplicate) break;
    ++count;
  }
  // GetWitnessPoints (b2distance.d:655-680)
  if (count == 1) { out.pointA = v[0].wA; out.pointB = v[0].wB; }
  else if (count == 2) { out.pointA = v[0].a * v[0].wA + v[1].a * v[1].wA; out.pointB = v[0].a * v[0].wB + v[1].a * v[1].wB; }
  else { out.pointA = v[0].a * v[0].wA + v[1].a * v[1].wA + v[2].a * v[2].wA; out.pointB = out.pointA; }
  out.distance = dist(out.pointA, out.pointB);
  out.iterations = iter;
  cache.metric = simplex_metric(v, count);
  cache.count = count;
  for (int i = 0; i < count; ++i) { cache.indexA[i] = (uint8_t)v[i].indexA; cache.indexB[i] = (uint8_t)v[i].indexB; }
  if (useRadii) {
    float rA = pA.radius, rB = pB.radius;
    if (out.distance > rA + rB && out.distance > kEpsilon) {
      out.distance -= rA + rB;
      v2 normal = out.pointB - out.pointA;
      normalize(normal);
      out.pointA += rA * normal;
      out.pointB -= rB * normal;
    } else {
      v2 p = 0.5f * (out.pointA + out.pointB);
      out.pointA = p; out.pointB = p; out.distance = 0.0f;
    }
  }
}

// b2TestOverlap(shapes) (b2collision.d:449-468) — the sensor path of b2Contact.Update
DBX_HD bool shapes_overlap(const DShape* A, Xf xfA, const DShape* B, Xf xfB) {
  DProxy pA = make_proxy(A), pB = make_proxy(B);
  SimplexCache cache; cache.count = 0; cache.metric = 0.0f;
  DistanceOutput out;
  gjk_distance(out, cache, pA, xfA, pB, xfB, true);
  return out.distance < 10.0f * kEpsilon;
}

// ------------------------------------------------------------------------------------------------ TOI
enum { TOI_UNKNOWN = 0, TOI_FAILED = 1, TOI_OVERLAPPED = 2, TOI_TOUCHING = 3, TOI_SEPARATED = 4 };
enum { SEP_POINTS = 0, SEP_FACE_A = 1, SEP_FACE_B = 2 };

struct SepFn {
  DProxy pA, pB;
  Sweep sA, sB;
  int type;
  v2 localPoint, axis;
};
// b2SeparationFunction.Initialize (b2timeofimpact.d:324-404)
DBX_HD void sep_init(SepFn& f, const SimplexCache& cache, const DProxy& pA, const Sweep& sA, const DProxy& pB, const Sweep& sB, float t1) {
  f.pA = pA; f.pB = pB; f.sA = sA; f.sB = sB;
  Xf xfA = sweep_xf(sA, t1), xfB = sweep_xf(sB, t1);
  if (cache.count == 1) {
    f.type = SEP_POINTS;
    v2 pointA = mul(xfA, pA.verts[cache.indexA[0]]);
    v2 pointB = mul(xfB, pB.verts[cache.indexB[0]]);
    f.axis = pointB - pointA;
    normalize(f.axis);
  } else if (cache.indexA[0] == cache.indexA[1]) {
    f.type = SEP_FACE_B;
    v2 b1 = pB.verts[cache.indexB[0]], b2 = pB.verts[cache.indexB[1]];
    f.axis = cross(b2 - b1, 1.0f);
    normalize(f.axis);
    v2 normal = mul(xfB.q, f.axis);
    f.localPoint = 0.5f * (b1 + b2);
    v2 pointB = mul(xfB, f.localPoint);
    v2 pointA = mul(xfA, pA.verts[cache.indexA[0]]);
    float s = dot(pointA - pointB, normal);
    if (s < 0.0f) f.axis = -f.axis;
  } else {
    f.type = SEP_FACE_A;
    v2 a1 = pA.verts[cache.indexA[0]], a2 = pA.verts[cache.indexA[1]];
    f.axis = cross(a2 - a1, 1.0f);
    normalize(f.axis);
    v2 normal = mul(xfA.q, f.axis);
    f.localPoint = 0.5f * (a1 + a2);
    v2 pointA = mul(xfA, f.localPoint);
    v2 pointB = mul(xfB, pB.verts[cache.indexB[0]]);
    float s = dot(pointB - pointA, normal);
    if (s < 0.0f) f.axis = -f.axis;
  }
}
// b2SeparationFunction.FindMinSeparation (b2timeofimpact.d:407-472)
DBX_HD float sep_find_min(const SepFn& f, int* indexA, int* indexB, float t) {
  Xf xfA = sweep_xf(f.sA, t), xfB = sweep_xf(f.sB, t);
  if (f.type == SEP_POINTS) {
    v2 axisA = mulT(xfA.q, f.axis), axisB = mulT(xfB.q, -f.axis);
    *indexA = proxy_support(f.pA, axisA);
    *indexB = proxy_support(f.pB, axisB);
    v2 pointA = mul(xfA, f.pA.verts[*indexA]), pointB = mul(xfB, f.pB.verts[*indexB]);
    return dot(pointB - pointA, f.axis);
  } else if (f.type == SEP_FACE_A) {
    v2 normal = mul(xfA.q, f.axis);
    v2 pointA = mul(xfA, f.localPoint);
    v2 axisB = mulT(xfB.q, -normal);
    *indexA = -1;
    *indexB = proxy_support(f.pB, axisB);
    v2 pointB = mul(xfB, f.pB.verts[*indexB]);
    return dot(pointB - pointA, normal);
  } else {
    v2 normal = mul(xfB.q, f.axis);
    v2 pointB = mul(xfB, f.localPoint);
    v2 axisA = mulT(xfA.q, -normal);
    *indexB = -1;
    *indexA = proxy_support(f.pA, axisA);
    v2 pointA = mul(xfA, f.pA.verts[*indexA]);
    return dot(pointA - pointB, normal);
  }
}
// b2SeparationFunction.Evaluate (b2timeofimpact.d:475-522)
DBX_HD float sep_eval(const SepFn& f, int indexA, int indexB, float t) {
  Xf xfA = sweep_xf(f.sA, t), xfB = sweep_xf(f.sB, t);
  if (f.type == SEP_POINTS) {
    v2 pointA = mul(xfA, f.pA.verts[indexA]), pointB = mul(xfB, f.pB.verts[indexB]);
    return dot(pointB - pointA, f.axis);
  } else if (f.type == SEP_FACE_A) {
    v2 normal = mul(xfA.q, f.axis);
    v2 pointA = mul(xfA, f.localPoint);
    v2 pointB = mul(xfB, f.pB.verts[indexB]);
    return dot(pointB - pointA, normal);
  } else {
    v2 normal = mul(xfB.q, f.axis);
    v2 pointB = mul(xfB, f.localPoint);
    v2 pointA = mul(xfA, f.pA.verts[indexA]);
    return dot(pointA - pointB, normal);
  }
}

// b2TimeOfImpact (b2timeofimpact.d:67-302)
DBX_HD int time_of_impact(float* tOut, const DProxy& pA, Sweep sweepA, const DProxy& pB, Sweep sweepB, float tMax) {
  int state = TOI_UNKNOWN;
  float tRes = tMax;
  sweep_normalize(sweepA);
  sweep_normalize(sweepB);
  float totalRadius = pA.radius + pB.radius;
  float target = fmaxr(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
  float tolerance = 0.25f * kLinearSlop;
  float t1 = 0.0f;
  const int k_maxIterations = 20;
  int iter = 0;
  SimplexCache cache; cache.count = 0; cache.metric = 0.0f;
  for (;;) {
    Xf xfA = sweep_xf(sweepA, t1), xfB = sweep_xf(sweepB, t1);
    DistanceOutput dout;
    gjk_distance(dout, cache, pA, xfA, pB, xfB, false);
    if (dout.distance <= 0.0f) { state = TOI_OVERLAPPED; tRes = 0.0f; break; }
    if (dout.distance < target + tolerance) { state = TOI_TOUCHING; tRes = t1; break; }
    SepFn fcn;
    sep_init(fcn, cache, pA, sweepA, pB, sweepB, t1);
    bool done = false;
    float t2 = tMax;
    int pushBackIter = 0;
    for (;;) {
      int indexA, indexB;
      float s2 = sep_find_min(fcn, &indexA, &indexB, t2);
      if (s2 > target + tolerance) { state = TOI_SEPARATED; tRes = tMax; done = true; break; }
      if (s2 > target - tolerance) { t1 = t2; break; }
      float s1 = sep_eval(fcn, indexA, indexB, t1);
      if (s1 < target - tolerance) { state = TOI_FAILED; tRes = t1; done = true; break; }
      if (s1 <= target + tolerance) { state = TOI_TOUCHING; tRes = t1; done = true; break; }
      int rootIterCount = 0;
      float a1 = t1, a2 = t2;
      for (;;) {
        float t;
        if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
        else t = 0.5f * (a1 + a2);
        ++rootIterCount;
        float s = sep_eval(fcn, indexA, indexB, t);
        if (fabsr(s - target) < tolerance) { t2 = t; break; }
        if (s > target) { a1 = t; s1 = s; }
        else { a2 = t; s2 = s; }
        if (rootIterCount == 50) break;
      }
      ++pushBackIter;
      if (pushBackIter == kMaxPolygonVertices) break;
    }
    ++iter;
    if (done) break;
    if (iter == k_maxIterations) { state = TOI_FAILED; tRes = t1; break; }
  }
  *tOut = tRes;
  return state;
}

}  // namespace dbx
