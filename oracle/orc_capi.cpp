// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// C API over the oracle world.  It deliberately has the same shape as the product's C ABI
// (include/dbox_b200.h: same POD structs, `orc_` prefix instead of `dbx_`) so that one Python scene
// builder can drive either side in the parity tests.  Only the struct *definitions* are shared with the
// product header; no product code is linked here.
#include <algorithm>
#include <cstring>
#include <map>
#include <thread>
#include <vector>
#include "../include/dbox_b200.h"
#include "orc_world.h"

using namespace orc;

struct orc_world {
  World w;
  std::vector<std::unique_ptr<std::vector<V2>>> keep;  // nothing retained from defs; placeholder
  explicit orc_world(V2 g) : w(g) {}
};

static V2 v2(dbx_vec2 a) { return V2(a.x, a.y); }
static dbx_vec2 d2(V2 a) { dbx_vec2 r; r.x = a.x; r.y = a.y; return r; }

static Shape toShape(const dbx_shape* s) {
  Shape o;
  o.type = s->type;
  o.radius = s->radius;
  switch (s->type) {
    case kCircle: o.p = v2(s->p); break;
    case kEdge: o.v0 = v2(s->v0); o.v1 = v2(s->v1); o.v2 = v2(s->v2); o.v3 = v2(s->v3); o.hasV0 = s->hasV0 != 0; o.hasV3 = s->hasV3 != 0; break;
    case kPolygon:
      o.centroid = v2(s->centroid); o.count = s->count;
      for (int i = 0; i < s->count; ++i) { o.verts[i] = v2(s->vertices[i]); o.normals[i] = v2(s->normals[i]); }
      break;
    case kChain:
      for (int i = 0; i < s->chainCount; ++i) o.chain.push_back(v2(s->chainVertices[i]));
      o.prevVertex = v2(s->prevVertex); o.nextVertex = v2(s->nextVertex); o.hasPrev = s->hasPrev != 0; o.hasNext = s->hasNext != 0;
      break;
  }
  return o;
}
static void fromShape(const Shape& o, dbx_shape* s) {
  std::memset(s, 0, sizeof(*s));
  s->type = o.type; s->radius = o.radius; s->p = d2(o.p);
  s->v0 = d2(o.v0); s->v1 = d2(o.v1); s->v2 = d2(o.v2); s->v3 = d2(o.v3); s->hasV0 = o.hasV0; s->hasV3 = o.hasV3;
  s->centroid = d2(o.centroid); s->count = o.count;
  for (int i = 0; i < o.count; ++i) { s->vertices[i] = d2(o.verts[i]); s->normals[i] = d2(o.normals[i]); }
}

extern "C" {

// ---- shape helpers (the oracle's restatement of the shape setup functions) ----
void orc_shape_set_circle(dbx_shape* out, float px, float py, float r) { fromShape(Shape::circle(V2(px, py), r), out); }
void orc_shape_set_edge(dbx_shape* out, dbx_vec2 a, dbx_vec2 b) { fromShape(Shape::edge(v2(a), v2(b)), out); }
void orc_shape_set_box(dbx_shape* out, float hx, float hy) { fromShape(Shape::box(hx, hy), out); }
void orc_shape_set_box_at(dbx_shape* out, float hx, float hy, dbx_vec2 c, float angle) { fromShape(Shape::box(hx, hy, v2(c), angle), out); }
int32_t orc_shape_set_polygon(dbx_shape* out, const dbx_vec2* pts, int32_t n) {
  std::vector<V2> p(n);
  for (int i = 0; i < n; ++i) p[i] = v2(pts[i]);
  fromShape(Shape::polygon(p.data(), n), out);
  return out->count;
}
void orc_shape_set_chain(dbx_shape* out, const dbx_vec2* pts, int32_t n, int32_t loop) {
  std::memset(out, 0, sizeof(*out));
  out->type = kChain; out->radius = kPolygonRadius;
  out->chainVertices = pts; out->chainCount = n;  // caller's array; a loop must already repeat vertex 0
  if (loop) { out->prevVertex = pts[n - 2]; out->nextVertex = pts[1]; out->hasPrev = out->hasNext = 1; }
}
void orc_shape_mass(const dbx_shape* s, float density, float* mass, dbx_vec2* center, float* I) {
  MassData md; toShape(s).computeMass(&md, density);
  *mass = md.mass; *center = d2(md.center); *I = md.I;
}
void orc_shape_aabb(const dbx_shape* s, float px, float py, float angle, int32_t child, dbx_aabb* out) {
  Xf xf; xf.set(V2(px, py), angle);
  AABB a; toShape(s).computeAABB(&a, xf, child);
  out->lo = d2(a.lo); out->hi = d2(a.hi);
}

// ---- world ----
orc_world* orc_world_create(float gx, float gy) { return new orc_world(V2(gx, gy)); }
void orc_world_destroy(orc_world* w) { delete w; }
int32_t orc_world_set_flags(orc_world* w, uint32_t f) {
  w->w.allowSleep = (f & DBX_WORLD_ALLOW_SLEEP) != 0;
  if (!w->w.allowSleep) for (Body* b = w->w.bodyList; b; b = b->next) b->setAwake(true);  // b2world.d:622-640
  w->w.warmStarting = (f & DBX_WORLD_WARM_STARTING) != 0;
  w->w.continuousPhysics = (f & DBX_WORLD_CONTINUOUS) != 0;
  w->w.subStepping = (f & DBX_WORLD_SUB_STEPPING) != 0;
  w->w.clearForcesFlag = (f & DBX_WORLD_AUTO_CLEAR_FORCES) != 0;
  return 0;
}
int32_t orc_world_set_gravity(orc_world* w, float gx, float gy) { w->w.gravity = V2(gx, gy); return 0; }

int32_t orc_body_create(orc_world* w, const dbx_body_def* d) {
  BodyDef bd;
  bd.type = d->type; bd.position = v2(d->position); bd.angle = d->angle; bd.linearVelocity = v2(d->linearVelocity);
  bd.angularVelocity = d->angularVelocity; bd.linearDamping = d->linearDamping; bd.angularDamping = d->angularDamping;
  bd.allowSleep = d->allowSleep != 0; bd.awake = d->awake != 0; bd.fixedRotation = d->fixedRotation != 0;
  bd.bullet = d->bullet != 0; bd.active = d->active != 0; bd.gravityScale = d->gravityScale; bd.userData = d->userData;
  Body* b = w->w.createBody(bd);
  return b ? b->id : DBX_E_LOCKED;
}
int32_t orc_body_destroy(orc_world* w, int32_t body) {
  if (body < 0 || body >= (int)w->w.bodiesById.size() || !w->w.bodiesById[body]) return DBX_E_INVALID;
  w->w.destroyBody(w->w.bodiesById[body]);
  return 0;
}
int32_t orc_fixture_create(orc_world* w, int32_t body, const dbx_fixture_def* d, const dbx_shape* s) {
  if (body < 0 || body >= (int)w->w.bodiesById.size() || !w->w.bodiesById[body]) return DBX_E_INVALID;
  Shape shape = toShape(s);
  FixtureDef fd;
  fd.shape = &shape; fd.friction = d->friction; fd.restitution = d->restitution; fd.density = d->density;
  fd.isSensor = d->isSensor != 0; fd.filter.categoryBits = d->categoryBits; fd.filter.maskBits = d->maskBits; fd.filter.groupIndex = d->groupIndex;
  fd.userData = d->userData;
  Fixture* f = w->w.createFixture(w->w.bodiesById[body], fd);
  return f ? f->id : DBX_E_LOCKED;
}
int32_t orc_fixture_destroy(orc_world* w, int32_t fixture) {
  if (fixture < 0 || fixture >= (int)w->w.fixturesById.size() || !w->w.fixturesById[fixture]) return DBX_E_INVALID;
  w->w.destroyFixture(w->w.fixturesById[fixture]);
  return 0;
}
int32_t orc_joint_create(orc_world* w, const dbx_joint_def* d) {
  auto& B = w->w.bodiesById;
  if (d->bodyA < 0 || d->bodyB < 0 || d->bodyA >= (int)B.size() || d->bodyB >= (int)B.size() || !B[d->bodyA] || !B[d->bodyB]) return DBX_E_INVALID;
  Joint* j = nullptr;
  if (d->type == jRevolute) {
    // b2revolutejoint.d:118-137
    RevoluteJoint* r = new RevoluteJoint();
    r->localAnchorA = v2(d->localAnchorA); r->localAnchorB = v2(d->localAnchorB); r->referenceAngle = d->referenceAngle;
    r->lowerAngle = d->lowerAngle; r->upperAngle = d->upperAngle; r->maxMotorTorque = d->maxMotorTorque; r->motorSpeed = d->motorSpeed;
    r->enableLimit = d->enableLimit != 0; r->enableMotor = d->enableMotor != 0; r->limitState = kInactiveLimit;
    j = r;
  } else if (d->type == jDistance) {
    // b2distancejoint.d:98-109
    DistanceJoint* dj = new DistanceJoint();
    dj->localAnchorA = v2(d->localAnchorA); dj->localAnchorB = v2(d->localAnchorB); dj->length = d->length;
    dj->frequencyHz = d->frequencyHz; dj->dampingRatio = d->dampingRatio;
    j = dj;
  } else if (d->type == jRope) {             // b2ropejoint.d:86-99
    RopeJoint* r = new RopeJoint();
    r->localAnchorA = v2(d->localAnchorA); r->localAnchorB = v2(d->localAnchorB); r->maxLength = d->maxLength;
    j = r;
  } else if (d->type == jWeld) {             // b2weldjoint.d:89-102
    WeldJoint* r = new WeldJoint();
    r->localAnchorA = v2(d->localAnchorA); r->localAnchorB = v2(d->localAnchorB); r->referenceAngle = d->referenceAngle;
    r->frequencyHz = d->frequencyHz; r->dampingRatio = d->dampingRatio;
    j = r;
  } else if (d->type == jFriction) {         // b2frictionjoint.d:81-94
    FrictionJoint* r = new FrictionJoint();
    r->localAnchorA = v2(d->localAnchorA); r->localAnchorB = v2(d->localAnchorB); r->maxForce = d->maxForce; r->maxTorque = d->maxTorque;
    j = r;
  } else if (d->type == jMotor) {            // b2motorjoint.d:90-104
    MotorJoint* r = new MotorJoint();
    r->linearOffset = v2(d->linearOffset); r->angularOffset = d->angularOffset; r->maxForce = d->maxForce; r->maxTorque = d->maxTorque;
    r->correctionFactor = d->correctionFactor;
    j = r;
  } else if (d->type == jMouse) {            // b2mousejoint.d:62-82
    MouseJoint* r = new MouseJoint();
    r->targetA = v2(d->target);
    r->localAnchorB = mulT(B[d->bodyB]->xf, r->targetA);
    r->maxForce = d->maxForce; r->frequencyHz = d->frequencyHz; r->dampingRatio = d->dampingRatio;
    j = r;
  } else if (d->type == jPrismatic) {        // b2prismaticjoint.d:166-190
    PrismaticJoint* r = new PrismaticJoint();
    r->localAnchorA = v2(d->localAnchorA); r->localAnchorB = v2(d->localAnchorB);
    r->localXAxisA = v2(d->localAxisA); r->localXAxisA.normalize(); r->localYAxisA = cross(1.0f, r->localXAxisA);
    r->referenceAngle = d->referenceAngle; r->lowerTranslation = d->lowerTranslation; r->upperTranslation = d->upperTranslation;
    r->maxMotorForce = d->maxMotorForce; r->motorSpeed = d->motorSpeed; r->enableLimit = d->enableLimit != 0; r->enableMotor = d->enableMotor != 0;
    j = r;
  } else if (d->type == jWheel) {            // b2wheeljoint.d:108-135
    WheelJoint* r = new WheelJoint();
    r->localAnchorA = v2(d->localAnchorA); r->localAnchorB = v2(d->localAnchorB);
    r->localXAxisA = v2(d->localAxisA); r->localYAxisA = cross(1.0f, r->localXAxisA);
    r->maxMotorTorque = d->maxMotorTorque; r->motorSpeed = d->motorSpeed; r->enableMotor = d->enableMotor != 0;
    r->frequencyHz = d->frequencyHz; r->dampingRatio = d->dampingRatio;
    j = r;
  } else if (d->type == jPulley) {           // b2pulleyjoint.d:107-123
    PulleyJoint* r = new PulleyJoint();
    r->localAnchorA = v2(d->localAnchorA); r->localAnchorB = v2(d->localAnchorB);
    r->groundAnchorA = v2(d->groundAnchorA); r->groundAnchorB = v2(d->groundAnchorB);
    r->lengthA = d->lengthA; r->lengthB = d->lengthB; r->ratio = d->ratio; r->constant = d->lengthA + r->ratio * d->lengthB;
    j = r;
  } else if (d->type == jGear) {             // b2gearjoint.d:84-160
    auto& J = w->w.jointsById;
    if (d->joint1 < 0 || d->joint2 < 0 || d->joint1 >= (int)J.size() || d->joint2 >= (int)J.size() || !J[d->joint1] || !J[d->joint2]) return DBX_E_INVALID;
    Joint* j1 = J[d->joint1]; Joint* j2 = J[d->joint2];
    if ((j1->type != jRevolute && j1->type != jPrismatic) || (j2->type != jRevolute && j2->type != jPrismatic)) return DBX_E_INVALID;
    GearJoint* g = new GearJoint();
    g->typeA = j1->type; g->typeB = j2->type;
    float coordinateA, coordinateB;
    g->bodyC = j1->bodyA; Body* bA = j1->bodyB;
    Xf xfA = bA->xf; float aA = bA->sweep.a; Xf xfC = g->bodyC->xf; float aC = g->bodyC->sweep.a;
    if (g->typeA == jRevolute) {
      RevoluteJoint* r = (RevoluteJoint*)j1;
      g->localAnchorC = r->localAnchorA; g->localAnchorA = r->localAnchorB; g->referenceAngleA = r->referenceAngle; g->localAxisC = V2(0, 0);
      coordinateA = aA - aC - g->referenceAngleA;
    } else {
      PrismaticJoint* pj = (PrismaticJoint*)j1;
      g->localAnchorC = pj->localAnchorA; g->localAnchorA = pj->localAnchorB; g->referenceAngleA = pj->referenceAngle; g->localAxisC = pj->localXAxisA;
      V2 pC = g->localAnchorC;
      V2 pA = mulT(xfC.q, mul(xfA.q, g->localAnchorA) + (xfA.p - xfC.p));
      coordinateA = dot(pA - pC, g->localAxisC);
    }
    g->bodyD = j2->bodyA; Body* bB = j2->bodyB;
    Xf xfB = bB->xf; float aB = bB->sweep.a; Xf xfD = g->bodyD->xf; float aD = g->bodyD->sweep.a;
    if (g->typeB == jRevolute) {
      RevoluteJoint* r = (RevoluteJoint*)j2;
      g->localAnchorD = r->localAnchorA; g->localAnchorB = r->localAnchorB; g->referenceAngleB = r->referenceAngle; g->localAxisD = V2(0, 0);
      coordinateB = aB - aD - g->referenceAngleB;
    } else {
      PrismaticJoint* pj = (PrismaticJoint*)j2;
      g->localAnchorD = pj->localAnchorA; g->localAnchorB = pj->localAnchorB; g->referenceAngleB = pj->referenceAngle; g->localAxisD = pj->localXAxisA;
      V2 pD = g->localAnchorD;
      V2 pB = mulT(xfD.q, mul(xfB.q, g->localAnchorB) + (xfB.p - xfD.p));
      coordinateB = dot(pB - pD, g->localAxisD);
    }
    g->ratio = d->ratio;
    g->constant = coordinateA + g->ratio * coordinateB;
    g->type = d->type; g->bodyA = bA; g->bodyB = bB; g->collideConnected = d->collideConnected != 0; g->userData = d->userData;
    Joint* r = w->w.addJoint(g);
    return r ? r->id : DBX_E_LOCKED;
  } else {
    return DBX_E_UNSUPPORTED;
  }
  j->type = d->type; j->bodyA = B[d->bodyA]; j->bodyB = B[d->bodyB]; j->collideConnected = d->collideConnected != 0; j->userData = d->userData;
  Joint* r = w->w.addJoint(j);
  return r ? r->id : DBX_E_LOCKED;
}
static Body* bodyAt(orc_world* w, int32_t body) { return (body >= 0 && body < (int)w->w.bodiesById.size()) ? w->w.bodiesById[body] : nullptr; }
static Fixture* fixtureAt(orc_world* w, int32_t f) { return (f >= 0 && f < (int)w->w.fixturesById.size()) ? w->w.fixturesById[f] : nullptr; }
// b2Body.SetMassData (b2body.d:502-540)
int32_t orc_body_set_mass_data(orc_world* w, int32_t body, float mass, float cx, float cy, float I) {
  Body* b = bodyAt(w, body); if (!b) return DBX_E_INVALID;
  if (w->w.locked || b->type != kDynamic) return 0;
  b->invMass = 0.0f; b->I = 0.0f; b->invI = 0.0f;
  b->mass = mass;
  if (b->mass <= 0.0f) b->mass = 1.0f;
  b->invMass = 1.0f / b->mass;
  V2 center(cx, cy);
  if (I > 0.0f && (b->flags & bFixedRotation) == 0) { b->I = I - b->mass * dot(center, center); b->invI = 1.0f / b->I; }
  V2 oldCenter = b->sweep.c;
  b->sweep.localCenter = center;
  b->sweep.c0 = b->sweep.c = mul(b->xf, b->sweep.localCenter);
  b->linearVelocity += cross(b->angularVelocity, b->sweep.c - oldCenter);
  return 0;
}
int32_t orc_body_reset_mass_data(orc_world* w, int32_t body) { Body* b = bodyAt(w, body); if (!b) return DBX_E_INVALID; b->resetMassData(); return 0; }
int32_t orc_body_set_fixed_rotation(orc_world* w, int32_t body, int32_t flag) {      // b2body.d:924-945
  Body* b = bodyAt(w, body); if (!b) return DBX_E_INVALID;
  bool status = (b->flags & bFixedRotation) == bFixedRotation;
  if (status == (flag != 0)) return 0;
  if (flag) b->flags |= bFixedRotation; else b->flags &= ~bFixedRotation;
  b->angularVelocity = 0.0f;
  b->resetMassData();
  return 0;
}
int32_t orc_body_set_linear_damping(orc_world* w, int32_t body, float d) { Body* b = bodyAt(w, body); if (!b) return DBX_E_INVALID; b->linearDamping = d; return 0; }
int32_t orc_body_set_angular_damping(orc_world* w, int32_t body, float d) { Body* b = bodyAt(w, body); if (!b) return DBX_E_INVALID; b->angularDamping = d; return 0; }
int32_t orc_body_set_gravity_scale(orc_world* w, int32_t body, float s) { Body* b = bodyAt(w, body); if (!b) return DBX_E_INVALID; b->gravityScale = s; return 0; }
// b2Fixture.SetFilterData + Refilter (b2fixture.d:131-178)
int32_t orc_fixture_set_filter(orc_world* w, int32_t fixture, int32_t cat, int32_t mask, int32_t group) {
  Fixture* f = fixtureAt(w, fixture); if (!f) return DBX_E_INVALID;
  f->filter.categoryBits = (uint16_t)cat; f->filter.maskBits = (uint16_t)mask; f->filter.groupIndex = (int16_t)group;
  for (ContactEdge* e = f->body->contactList; e; e = e->next) { Contact* c = e->contact; if (c->fixtureA == f || c->fixtureB == f) c->flags |= cFilter; }
  for (int i = 0; i < f->proxyCount; ++i) w->w.broadPhase.touchProxy(f->proxies[i].proxyId);
  return 0;
}
int32_t orc_fixture_set_sensor(orc_world* w, int32_t fixture, int32_t flag) {        // b2fixture.d:108-115
  Fixture* f = fixtureAt(w, fixture); if (!f) return DBX_E_INVALID;
  if ((flag != 0) != f->isSensor) { f->body->setAwake(true); f->isSensor = flag != 0; }
  return 0;
}
int32_t orc_fixture_set_friction(orc_world* w, int32_t fixture, float v) { Fixture* f = fixtureAt(w, fixture); if (!f) return DBX_E_INVALID; f->friction = v; return 0; }
int32_t orc_fixture_set_restitution(orc_world* w, int32_t fixture, float v) { Fixture* f = fixtureAt(w, fixture); if (!f) return DBX_E_INVALID; f->restitution = v; return 0; }
int32_t orc_fixture_set_density(orc_world* w, int32_t fixture, float v) { Fixture* f = fixtureAt(w, fixture); if (!f) return DBX_E_INVALID; f->density = v; return 0; }
int32_t orc_body_set_type(orc_world* w, int32_t body, int32_t type) {
  if (body < 0 || body >= (int)w->w.bodiesById.size() || !w->w.bodiesById[body] || type < 0 || type > 2) return DBX_E_INVALID;
  w->w.setBodyType(w->w.bodiesById[body], type); return 0;
}
int32_t orc_body_set_active(orc_world* w, int32_t body, int32_t flag) {
  if (body < 0 || body >= (int)w->w.bodiesById.size() || !w->w.bodiesById[body]) return DBX_E_INVALID;
  w->w.setBodyActive(w->w.bodiesById[body], flag != 0); return 0;
}
// b2MouseJoint.SetTarget (b2mousejoint.d:112-120)
int32_t orc_joint_set_target(orc_world* w, int32_t joint, float x, float y) {
  if (joint < 0 || joint >= (int)w->w.jointsById.size() || !w->w.jointsById[joint] || w->w.jointsById[joint]->type != jMouse) return DBX_E_INVALID;
  MouseJoint* m = (MouseJoint*)w->w.jointsById[joint];
  if (m->bodyB->isAwake() == false) m->bodyB->setAwake(true);
  m->targetA = V2(x, y);
  return 0;
}
int32_t orc_joint_destroy(orc_world* w, int32_t joint) {
  if (joint < 0 || joint >= (int)w->w.jointsById.size() || !w->w.jointsById[joint]) return DBX_E_INVALID;
  w->w.destroyJoint(w->w.jointsById[joint]);
  return 0;
}

int32_t orc_world_step(orc_world* w, float dt, int32_t vi, int32_t pi) { w->w.step(dt, vi, pi); return 0; }
int32_t orc_world_step_n(orc_world* w, float dt, int32_t vi, int32_t pi, int32_t n) { for (int i = 0; i < n; ++i) w->w.step(dt, vi, pi); return 0; }

// ---- body access ----
static void fillBody(const Body* b, dbx_body_state* o) {
  o->type = b->type; o->flags = b->flags; o->p = d2(b->xf.p); o->qs = b->xf.q.s; o->qc = b->xf.q.c;
  o->localCenter = d2(b->sweep.localCenter); o->c0 = d2(b->sweep.c0); o->c = d2(b->sweep.c);
  o->a0 = b->sweep.a0; o->a = b->sweep.a; o->alpha0 = b->sweep.alpha0;
  o->v = d2(b->linearVelocity); o->w = b->angularVelocity; o->force = d2(b->force); o->torque = b->torque;
  o->mass = b->mass; o->invMass = b->invMass; o->I = b->I; o->invI = b->invI;
  o->linearDamping = b->linearDamping; o->angularDamping = b->angularDamping; o->gravityScale = b->gravityScale; o->sleepTime = b->sleepTime;
}
static Body* getBody(orc_world* w, int32_t id) { return (id >= 0 && id < (int)w->w.bodiesById.size()) ? w->w.bodiesById[id] : nullptr; }
int32_t orc_body_get_state(orc_world* w, int32_t body, dbx_body_state* out) { Body* b = getBody(w, body); if (!b) return DBX_E_INVALID; fillBody(b, out); return 0; }
int32_t orc_world_read_bodies(orc_world* w, dbx_body_state* out, int32_t cap) {
  int n = (int)w->w.bodiesById.size();
  for (int i = 0; i < n && i < cap; ++i) { if (w->w.bodiesById[i]) fillBody(w->w.bodiesById[i], out + i); else std::memset(out + i, 0, sizeof(*out)); }
  return n;
}
// b2body.d:261-285
int32_t orc_body_set_transform(orc_world* w, int32_t body, float x, float y, float angle) {
  Body* b = getBody(w, body); if (!b) return DBX_E_INVALID;
  b->xf.q.set(angle); b->xf.p = V2(x, y);
  b->sweep.c = mul(b->xf, b->sweep.localCenter); b->sweep.a = angle;
  b->sweep.c0 = b->sweep.c; b->sweep.a0 = angle;
  for (Fixture* f = b->fixtureList; f; f = f->next) {
    for (int i = 0; i < f->proxyCount; ++i) {
      FixtureProxy* p = &f->proxies[i];
      AABB a1, a2; f->shape.computeAABB(&a1, b->xf, p->childIndex); f->shape.computeAABB(&a2, b->xf, p->childIndex);
      p->aabb.combine(a1, a2);
      w->w.broadPhase.moveProxy(p->proxyId, p->aabb, b->xf.p - b->xf.p);
    }
  }
  return 0;
}
int32_t orc_body_set_linear_velocity(orc_world* w, int32_t body, float vx, float vy) {
  Body* b = getBody(w, body); if (!b) return DBX_E_INVALID;
  if (b->type == kStatic) return 0;
  V2 v(vx, vy);
  if (dot(v, v) > 0.0f) b->setAwake(true);
  b->linearVelocity = v; return 0;
}
int32_t orc_body_set_angular_velocity(orc_world* w, int32_t body, float om) {
  Body* b = getBody(w, body); if (!b) return DBX_E_INVALID;
  if (b->type == kStatic) return 0;
  if (om * om > 0.0f) b->setAwake(true);
  b->angularVelocity = om; return 0;
}
int32_t orc_body_apply_force(orc_world* w, int32_t body, float fx, float fy, float px, float py, int32_t wake) {
  Body* b = getBody(w, body); if (!b) return DBX_E_INVALID;
  if (b->type != kDynamic) return 0;
  if (wake && (b->flags & bAwake) == 0) b->setAwake(true);
  if (b->flags & bAwake) { b->force += V2(fx, fy); b->torque += cross(V2(px, py) - b->sweep.c, V2(fx, fy)); }
  return 0;
}
int32_t orc_body_apply_torque(orc_world* w, int32_t body, float t, int32_t wake) {
  Body* b = getBody(w, body); if (!b) return DBX_E_INVALID;
  if (b->type != kDynamic) return 0;
  if (wake && (b->flags & bAwake) == 0) b->setAwake(true);
  if (b->flags & bAwake) b->torque += t;
  return 0;
}
int32_t orc_body_apply_linear_impulse(orc_world* w, int32_t body, float ix, float iy, float px, float py, int32_t wake) {
  Body* b = getBody(w, body); if (!b) return DBX_E_INVALID;
  if (b->type != kDynamic) return 0;
  if (wake && (b->flags & bAwake) == 0) b->setAwake(true);
  if (b->flags & bAwake) { b->linearVelocity += b->invMass * V2(ix, iy); b->angularVelocity += b->invI * cross(V2(px, py) - b->sweep.c, V2(ix, iy)); }
  return 0;
}
int32_t orc_body_apply_angular_impulse(orc_world* w, int32_t body, float imp, int32_t wake) {
  Body* b = getBody(w, body); if (!b) return DBX_E_INVALID;
  if (b->type != kDynamic) return 0;
  if (wake && (b->flags & bAwake) == 0) b->setAwake(true);
  if (b->flags & bAwake) b->angularVelocity += b->invI * imp;
  return 0;
}
int32_t orc_body_set_awake(orc_world* w, int32_t body, int32_t flag) { Body* b = getBody(w, body); if (!b) return DBX_E_INVALID; b->setAwake(flag != 0); return 0; }
int32_t orc_body_set_bullet(orc_world* w, int32_t body, int32_t flag) { Body* b = getBody(w, body); if (!b) return DBX_E_INVALID; if (flag) b->flags |= bBullet; else b->flags &= ~bBullet; return 0; }
int32_t orc_body_set_sleeping_allowed(orc_world* w, int32_t body, int32_t flag) {
  Body* b = getBody(w, body); if (!b) return DBX_E_INVALID;
  if (flag) b->flags |= bAutoSleep; else { b->flags &= ~bAutoSleep; b->setAwake(true); }
  return 0;
}

// ---- bulk state ----
int32_t orc_world_counts(orc_world* w, dbx_counts* o) {
  std::memset(o, 0, sizeof(*o));
  o->bodies = w->w.bodyCount; o->joints = w->w.jointCount; o->contacts = w->w.contactCount; o->proxies = w->w.broadPhase.proxyCount();
  for (Fixture* f : w->w.fixturesById) if (f) ++o->fixtures;
  for (Contact* c = w->w.contactList; c; c = c->next) if (c->isTouching()) ++o->touching;
  for (Body* b = w->w.bodyList; b; b = b->next) if (b->isAwake() && b->type != kStatic) ++o->awakeBodies;
  o->islands = w->w.lastIslandCount; o->moves = (int)w->w.broadPhase.moveBuffer().size(); o->pairs = (int)w->w.lastPairs.size();
  o->colours = w->w.toiEvents;  // oracle has no colouring; slot reused for the cumulative TOI event count (tests)
  return 0;
}
int32_t orc_world_profile(orc_world* w, dbx_profile* o) {
  const Profile& p = w->w.profile;
  o->step = p.step; o->collide = p.collide; o->solve = p.solve; o->solveInit = p.solveInit; o->solveVelocity = p.solveVelocity;
  o->solvePosition = p.solvePosition; o->broadphase = p.broadphase; o->solveTOI = p.solveTOI;
  return 0;
}
static void fillContact(const Contact* c, dbx_contact_rec* o) {
  o->fixtureA = c->fixtureA->id; o->fixtureB = c->fixtureB->id; o->childA = c->indexA; o->childB = c->indexB; o->flags = c->flags;
  for (int i = 0; i < 2; ++i) {
    o->manifold.points[i].localPoint = d2(c->manifold.points[i].localPoint);
    o->manifold.points[i].normalImpulse = c->manifold.points[i].normalImpulse;
    o->manifold.points[i].tangentImpulse = c->manifold.points[i].tangentImpulse;
    o->manifold.points[i].key = c->manifold.points[i].id.key;
  }
  o->manifold.localNormal = d2(c->manifold.localNormal); o->manifold.localPoint = d2(c->manifold.localPoint);
  o->manifold.type = c->manifold.type; o->manifold.pointCount = c->manifold.pointCount;
  o->friction = c->friction; o->restitution = c->restitution; o->tangentSpeed = c->tangentSpeed; o->toiCount = c->toiCount; o->toi = c->toi;
}
// world contact list order (newest first), as GetContactList would traverse it (b2world.d:610-613)
int32_t orc_world_read_contacts(orc_world* w, dbx_contact_rec* out, int32_t cap) {
  int n = 0;
  for (Contact* c = w->w.contactList; c; c = c->next) { if (n < cap) fillContact(c, out + n); ++n; }
  return n;
}
// the step cut at PreSolve (see include/dbox_b200.h)
int32_t orc_world_step_begin(orc_world* w, float dt, int32_t vi, int32_t pi) {
  if (w->w.midStep) return DBX_E_INVALID;
  w->w.pendDt = dt; w->w.pendVi = vi; w->w.pendPi = pi; w->w.midStep = true;
  w->w.stepHalves(dt, vi, pi, 1);
  return 0;
}
int32_t orc_world_step_end(orc_world* w) {
  if (!w->w.midStep) return DBX_E_INVALID;
  w->w.midStep = false;
  w->w.stepHalves(w->w.pendDt, w->w.pendVi, w->w.pendPi, 2);
  return 0;
}
int32_t orc_world_patch_contacts(orc_world* w, const dbx_contact_patch* p, int32_t n) {
  for (int k = 0; k < n; ++k) {
    for (Contact* c = w->w.contactList; c; c = c->next) {
      const bool same = c->fixtureA->id == p[k].fixtureA && c->indexA == p[k].childA && c->fixtureB->id == p[k].fixtureB && c->indexB == p[k].childB;
      const bool swapped = c->fixtureA->id == p[k].fixtureB && c->indexA == p[k].childB && c->fixtureB->id == p[k].fixtureA && c->indexB == p[k].childA;
      if (!same && !swapped) continue;
      if (p[k].mask & DBX_PATCH_ENABLED) {                                                                          // b2contact.d:137-147
        if (p[k].enabled) { c->flags |= cEnabled; c->flags &= ~cPreSolveOff; } else { c->flags &= ~cEnabled; c->flags |= cPreSolveOff; }
      }
      if (p[k].mask & DBX_PATCH_FRICTION) c->friction = p[k].friction;
      if (p[k].mask & DBX_PATCH_RESTITUTION) c->restitution = p[k].restitution;
      if (p[k].mask & DBX_PATCH_TANGENT_SPEED) c->tangentSpeed = p[k].tangentSpeed;
      break;
    }
  }
  return n;
}
// b2World.RayCast (b2world.d:577-587, wrapper :1605-1624) with the callback that returns `fraction` (closest hit)
int32_t orc_world_raycast_closest(orc_world* w, const dbx_ray* rays, int32_t n, dbx_ray_hit* out) {
  for (int k = 0; k < n; ++k) {
    dbx_ray_hit& h = out[k];
    h.fixture = -1; h.child = 0; h.fraction = 1.0f; h.point = dbx_vec2{0, 0}; h.normal = dbx_vec2{0, 0};
    V2 P1 = v2(rays[k].p1), P2 = v2(rays[k].p2);
    if ((P2 - P1).len2() <= 0.0f) continue;
    w->w.broadPhase.tree().rayCast([&](V2 p1, V2 p2, float maxFraction, int proxyId) -> float {
      FixtureProxy* proxy = (FixtureProxy*)w->w.broadPhase.userData(proxyId);
      Fixture* f = proxy->fixture;
      float fraction; V2 normal;
      bool hit = f->shape.rayCast(&fraction, &normal, p1, p2, maxFraction, f->body->xf, proxy->childIndex);
      if (!hit) return maxFraction;
      V2 point = (1.0f - fraction) * p1 + fraction * p2;
      h.fixture = f->id; h.child = proxy->childIndex; h.fraction = fraction; h.point = d2(point); h.normal = d2(normal);
      return fraction;
    }, P1, P2, 1.0f);
  }
  return n;
}
// b2World.QueryAABB (b2world.d:563-570): fixtures whose FAT proxy box overlaps; reported here sorted by (fixture, child)
int32_t orc_world_query_aabb(orc_world* w, const dbx_aabb* boxes, int32_t n, int32_t capPerQuery, int32_t* counts, int32_t* fixture_child) {
  for (int k = 0; k < n; ++k) {
    AABB box; box.lo = v2(boxes[k].lo); box.hi = v2(boxes[k].hi);
    std::vector<std::pair<int, int>> hits;
    w->w.broadPhase.tree().query([&](int proxyId) {
      FixtureProxy* proxy = (FixtureProxy*)w->w.broadPhase.userData(proxyId);
      hits.push_back({proxy->fixture->id, proxy->childIndex});
      return true;
    }, box);
    std::sort(hits.begin(), hits.end());
    counts[k] = (int)hits.size();
    for (int i = 0; i < (int)hits.size() && i < capPerQuery; ++i) { fixture_child[2 * ((size_t)k * capPerQuery + i)] = hits[i].first; fixture_child[2 * ((size_t)k * capPerQuery + i) + 1] = hits[i].second; }
  }
  return n;
}
// b2World.RayCast with the callback that returns 1 (report every fixture, keep the full ray): sorted by (fraction, fixture, child)
int32_t orc_world_raycast_all(orc_world* w, const dbx_ray* rays, int32_t n, int32_t capPerRay, int32_t* counts, dbx_ray_hit* hits) {
  for (int k = 0; k < n; ++k) {
    std::vector<dbx_ray_hit> found;
    V2 P1 = v2(rays[k].p1), P2 = v2(rays[k].p2);
    if ((P2 - P1).len2() > 0.0f) {
      w->w.broadPhase.tree().rayCast([&](V2 p1, V2 p2, float maxFraction, int proxyId) -> float {
        FixtureProxy* proxy = (FixtureProxy*)w->w.broadPhase.userData(proxyId);
        Fixture* f = proxy->fixture;
        float fraction; V2 normal;
        bool hit = f->shape.rayCast(&fraction, &normal, p1, p2, maxFraction, f->body->xf, proxy->childIndex);
        if (!hit) return maxFraction;
        V2 point = (1.0f - fraction) * p1 + fraction * p2;
        dbx_ray_hit h; h.fixture = f->id; h.child = proxy->childIndex; h.fraction = fraction; h.point = d2(point); h.normal = d2(normal);
        found.push_back(h);
        return 1.0f;
      }, P1, P2, 1.0f);
    }
    std::sort(found.begin(), found.end(), [](const dbx_ray_hit& x, const dbx_ray_hit& y) {
      return x.fraction != y.fraction ? x.fraction < y.fraction : x.fixture != y.fixture ? x.fixture < y.fixture : x.child < y.child; });
    counts[k] = (int)found.size();
    for (int i = 0; i < (int)found.size() && i < capPerRay; ++i) hits[(size_t)k * capPerRay + i] = found[i];
  }
  return n;
}
// b2Fixture.TestPoint (b2fixture.d:209-212)
int32_t orc_world_test_points(orc_world* w, const int32_t* fixtures, const dbx_vec2* points, int32_t n, int32_t* inside) {
  for (int k = 0; k < n; ++k) {
    const int id = fixtures[k];
    if (id < 0 || id >= (int)w->w.fixturesById.size() || !w->w.fixturesById[id]) return DBX_E_INVALID;
    const Fixture* f = w->w.fixturesById[id];
    inside[k] = f->shape.testPoint(f->body->xf, v2(points[k])) ? 1 : 0;
  }
  return n;
}
int32_t orc_world_shift_origin(orc_world* w, float x, float y) { w->w.shiftOrigin(V2(x, y)); return 0; }
// b2Contact.GetWorldManifold (contacts/b2contact.d:77-91) for every contact, in orc_world_read_contacts order
int32_t orc_world_read_world_manifolds(orc_world* w, dbx_world_manifold* out, int32_t cap) {
  int n = 0;
  for (Contact* c = w->w.contactList; c; c = c->next) {
    if (n < cap) {
      WorldManifold wm;
      const Shape& sa = c->fixtureA->shape; const Shape& sb = c->fixtureB->shape;
      wm.initialize(&c->manifold, c->fixtureA->body->xf, sa.radius, c->fixtureB->body->xf, sb.radius);
      dbx_world_manifold& o = out[n];
      std::memset(&o, 0, sizeof(o));
      o.pointCount = c->manifold.pointCount;
      if (o.pointCount > 0) {
        o.normal = d2(wm.normal);
        for (int i = 0; i < o.pointCount && i < 2; ++i) { o.points[i] = d2(wm.points[i]); o.separations[i] = wm.separations[i]; }
      }
    }
    ++n;
  }
  return n;
}
// the joint classes' setters, grouped by `mask` as in include/dbox_b200.h; side effects as in the reference:
// b2revolutejoint.d:216-300, b2prismaticjoint.d:250-330, b2wheeljoint.d:170-230, b2motorjoint.d:110-170
int32_t orc_joint_set_params(orc_world* w, int32_t joint, const dbx_joint_def* d, uint32_t mask) {
  auto& J = w->w.jointsById;
  if (joint < 0 || joint >= (int)J.size() || !J[joint] || J[joint]->type != d->type) return DBX_E_INVALID;
  Joint* j = J[joint];
  auto wake = [&]() { j->bodyA->setAwake(true); j->bodyB->setAwake(true); };
  if (j->type == jRevolute) {
    RevoluteJoint* r = (RevoluteJoint*)j;
    if (mask & DBX_JP_MOTOR_SPEED) { wake(); r->motorSpeed = d->motorSpeed; }
    if (mask & DBX_JP_MAX_MOTOR) { wake(); r->maxMotorTorque = d->maxMotorTorque; }
    if (mask & DBX_JP_ENABLE_MOTOR) { wake(); r->enableMotor = d->enableMotor != 0; }
    if ((mask & DBX_JP_ENABLE_LIMIT) && (d->enableLimit != 0) != r->enableLimit) { wake(); r->enableLimit = d->enableLimit != 0; r->impulse.z = 0.0f; }
    if ((mask & DBX_JP_LIMITS) && (d->lowerAngle != r->lowerAngle || d->upperAngle != r->upperAngle)) { wake(); r->impulse.z = 0.0f; r->lowerAngle = d->lowerAngle; r->upperAngle = d->upperAngle; }
  } else if (j->type == jPrismatic) {
    PrismaticJoint* r = (PrismaticJoint*)j;
    if (mask & DBX_JP_MOTOR_SPEED) { wake(); r->motorSpeed = d->motorSpeed; }
    if (mask & DBX_JP_MAX_MOTOR) { wake(); r->maxMotorForce = d->maxMotorForce; }
    if (mask & DBX_JP_ENABLE_MOTOR) { wake(); r->enableMotor = d->enableMotor != 0; }
    if ((mask & DBX_JP_ENABLE_LIMIT) && (d->enableLimit != 0) != r->enableLimit) { wake(); r->enableLimit = d->enableLimit != 0; r->impulse.z = 0.0f; }
    if ((mask & DBX_JP_LIMITS) && (d->lowerTranslation != r->lowerTranslation || d->upperTranslation != r->upperTranslation)) { wake(); r->lowerTranslation = d->lowerTranslation; r->upperTranslation = d->upperTranslation; r->impulse.z = 0.0f; }
  } else if (j->type == jWheel) {
    WheelJoint* r = (WheelJoint*)j;
    if (mask & DBX_JP_MOTOR_SPEED) { wake(); r->motorSpeed = d->motorSpeed; }
    if (mask & DBX_JP_MAX_MOTOR) { wake(); r->maxMotorTorque = d->maxMotorTorque; }
    if (mask & DBX_JP_ENABLE_MOTOR) { wake(); r->enableMotor = d->enableMotor != 0; }
    if (mask & DBX_JP_SPRING) { r->frequencyHz = d->frequencyHz; r->dampingRatio = d->dampingRatio; }
  } else if (j->type == jDistance) {
    DistanceJoint* r = (DistanceJoint*)j;
    if (mask & DBX_JP_LENGTH) r->length = d->length;
    if (mask & DBX_JP_SPRING) { r->frequencyHz = d->frequencyHz; r->dampingRatio = d->dampingRatio; }
  } else if (j->type == jWeld) {
    WeldJoint* r = (WeldJoint*)j;
    if (mask & DBX_JP_SPRING) { r->frequencyHz = d->frequencyHz; r->dampingRatio = d->dampingRatio; }
  } else if (j->type == jRope) {
    if (mask & DBX_JP_LENGTH) ((RopeJoint*)j)->maxLength = d->maxLength;
  } else if (j->type == jFriction) {
    FrictionJoint* r = (FrictionJoint*)j;
    if (mask & DBX_JP_MAX_FORCE) { r->maxForce = d->maxForce; r->maxTorque = d->maxTorque; }
  } else if (j->type == jMouse) {
    MouseJoint* r = (MouseJoint*)j;
    if (mask & DBX_JP_MAX_FORCE) r->maxForce = d->maxForce;
    if (mask & DBX_JP_SPRING) { r->frequencyHz = d->frequencyHz; r->dampingRatio = d->dampingRatio; }
  } else if (j->type == jMotor) {
    MotorJoint* r = (MotorJoint*)j;
    if (mask & DBX_JP_MAX_FORCE) { r->maxForce = d->maxForce; r->maxTorque = d->maxTorque; }
    if ((mask & DBX_JP_OFFSETS) && (d->linearOffset.x != r->linearOffset.x || d->linearOffset.y != r->linearOffset.y || d->angularOffset != r->angularOffset)) {
      wake(); r->linearOffset = v2(d->linearOffset); r->angularOffset = d->angularOffset;
    }
    if (mask & DBX_JP_CORRECTION) r->correctionFactor = d->correctionFactor;
  }
  return 0;
}
int32_t orc_world_set_motor_speeds(orc_world* w, const int32_t* joints, const float* speeds, int32_t n) {
  for (int k = 0; k < n; ++k) {
    dbx_joint_def d{};
    auto& J = w->w.jointsById;
    if (joints[k] < 0 || joints[k] >= (int)J.size() || !J[joints[k]]) return DBX_E_INVALID;
    d.type = J[joints[k]]->type; d.motorSpeed = speeds[k];
    if (d.type != jRevolute && d.type != jPrismatic && d.type != jWheel) return DBX_E_INVALID;
    int rc = orc_joint_set_params(w, joints[k], &d, DBX_JP_MOTOR_SPEED); if (rc < 0) return rc;
  }
  return n;
}
// b2World.SetContactFilter with a C callback standing in for the user's b2ContactFilter subclass (tests)
int32_t orc_world_set_contact_filter(orc_world* w, int (*cb)(int, int, int)) { w->w.userFilter = cb; return 0; }
// the PostSolve call log of the last step, in the reference's CALL order
int32_t orc_world_enable_post_solve(orc_world* w, int32_t capacity) { w->w.recordPostSolve = capacity > 0; w->w.postSolveLog.clear(); return capacity; }
int32_t orc_world_read_post_solve(orc_world* w, dbx_post_solve* out, int32_t cap) {
  const int n = (int)w->w.postSolveLog.size();
  if (!out || cap <= 0) return n;
  for (int i = 0; i < n && i < cap; ++i) {
    const World::PostSolveRec& r = w->w.postSolveLog[i];
    dbx_post_solve& o = out[i];
    o.fixtureA = r.fixtureA; o.fixtureB = r.fixtureB; o.childA = r.childA; o.childB = r.childB; o.phase = r.phase; o.count = r.count;
    for (int j = 0; j < 2; ++j) { o.normalImpulses[j] = r.normalImpulses[j]; o.tangentImpulses[j] = r.tangentImpulses[j]; }
  }
  return n;
}
// the listener call log since the last poll, in the reference's CALL order (the CUDA library returns the same events sorted)
int32_t orc_world_enable_contact_events(orc_world* w, int32_t capacity) {
  w->w.recordContactEvents = capacity > 0; w->w.contactEvents.clear(); return capacity;
}
int32_t orc_world_poll_contact_events(orc_world* w, dbx_contact_event* out, int32_t cap) {
  const int n = (int)w->w.contactEvents.size();
  if (!out || cap <= 0) return n;
  for (int i = 0; i < n && i < cap; ++i) {
    const World::ContactEvt& e = w->w.contactEvents[i];
    out[i].type = e.type; out[i].phase = e.phase; out[i].stepsAgo = w->w.stepCount - e.step;
    out[i].fixtureA = e.fixtureA; out[i].fixtureB = e.fixtureB; out[i].childA = e.childA; out[i].childB = e.childB;
    out[i].bodyA = e.bodyA; out[i].bodyB = e.bodyB;
  }
  w->w.contactEvents.clear();
  return n;
}
// contacts in the order the islands solved them during the last Step (sequential Gauss-Seidel order)
int32_t orc_world_read_solve_order(orc_world* w, int32_t* fixA_childA_fixB_childB, int32_t cap) {
  int n = 0;
  for (Contact* c : w->w.lastSolveOrder) {
    if (n < cap) { int32_t* o = fixA_childA_fixB_childB + 4 * n; o[0] = c->fixtureA->id; o[1] = c->indexA; o[2] = c->fixtureB->id; o[3] = c->indexB; }
    ++n;
  }
  return n;
}
int32_t orc_world_read_joint_solve_order(orc_world* w, int32_t* jointIds, int32_t cap) {
  int n = 0;
  for (int id : w->w.lastJointOrder) { if (n < cap) jointIds[n] = id; ++n; }
  return n;
}
int32_t orc_world_read_proxies(orc_world* w, dbx_proxy_rec* out, int32_t cap) {
  int n = 0;
  for (Fixture* f : w->w.fixturesById) {
    if (!f) continue;
    for (int i = 0; i < f->proxyCount; ++i) {
      if (n < cap) {
        dbx_proxy_rec* o = out + n;
        o->fixture = f->id; o->child = f->proxies[i].childIndex; o->proxyId = f->proxies[i].proxyId;
        o->aabb.lo = d2(f->proxies[i].aabb.lo); o->aabb.hi = d2(f->proxies[i].aabb.hi);
        const AABB& fat = w->w.broadPhase.fatAABB(f->proxies[i].proxyId);
        o->fat.lo = d2(fat.lo); o->fat.hi = d2(fat.hi);
      }
      ++n;
    }
  }
  return n;
}
int32_t orc_world_read_joints(orc_world* w, dbx_joint_state* out, int32_t cap) {
  int n = (int)w->w.jointsById.size();
  for (int i = 0; i < n && i < cap; ++i) {
    dbx_joint_state* o = out + i; std::memset(o, 0, sizeof(*o));
    Joint* j = w->w.jointsById[i]; if (!j) continue;
    o->type = j->type;
    if (j->type == jRevolute) { auto* r = (RevoluteJoint*)j; o->impulse[0] = r->impulse.x; o->impulse[1] = r->impulse.y; o->impulse[2] = r->impulse.z; o->motorImpulse = r->motorImpulse; o->limitState = r->limitState; }
    else if (j->type == jDistance) { auto* d = (DistanceJoint*)j; o->impulse[0] = d->impulse; }
    else if (j->type == jRope) { auto* d = (RopeJoint*)j; o->impulse[0] = d->impulse; o->limitState = d->state; }
    else if (j->type == jWeld) { auto* d = (WeldJoint*)j; o->impulse[0] = d->impulse.x; o->impulse[1] = d->impulse.y; o->impulse[2] = d->impulse.z; }
    else if (j->type == jFriction) { auto* d = (FrictionJoint*)j; o->impulse[0] = d->linearImpulse.x; o->impulse[1] = d->linearImpulse.y; o->impulse[2] = d->angularImpulse; }
    else if (j->type == jMotor) { auto* d = (MotorJoint*)j; o->impulse[0] = d->linearImpulse.x; o->impulse[1] = d->linearImpulse.y; o->impulse[2] = d->angularImpulse; }
    else if (j->type == jMouse) { auto* d = (MouseJoint*)j; o->impulse[0] = d->impulse.x; o->impulse[1] = d->impulse.y; }
    else if (j->type == jPrismatic) { auto* d = (PrismaticJoint*)j; o->impulse[0] = d->impulse.x; o->impulse[1] = d->impulse.y; o->impulse[2] = d->impulse.z; o->motorImpulse = d->motorImpulse; o->limitState = d->limitState; }
    else if (j->type == jWheel) { auto* d = (WheelJoint*)j; o->impulse[0] = d->impulse; o->impulse[1] = d->springImpulse; o->motorImpulse = d->motorImpulse; }
    else if (j->type == jPulley) { auto* d = (PulleyJoint*)j; o->impulse[0] = d->impulse; }
    else if (j->type == jGear) { auto* d = (GearJoint*)j; o->impulse[0] = d->impulse; }
  }
  return n;
}
int32_t orc_world_read_moves(orc_world* w, int32_t* out, int32_t cap) {
  int n = 0;
  for (int id : w->w.broadPhase.moveBuffer()) {
    if (id == kNullNode) continue;
    FixtureProxy* p = (FixtureProxy*)w->w.broadPhase.userData(id);
    if (n < cap) { out[2 * n] = p->fixture->id; out[2 * n + 1] = p->childIndex; }
    ++n;
  }
  return n;
}
int32_t orc_world_read_pairs(orc_world* w, int32_t* out, int32_t cap) {
  int n = 0;
  for (auto& pr : w->w.lastPairs) {
    if (n < cap) { int32_t* o = out + 4 * n; o[0] = pr.first->fixture->id; o[1] = pr.first->childIndex; o[2] = pr.second->fixture->id; o[3] = pr.second->childIndex; }
    ++n;
  }
  return n;
}
// ------------------------------------------------------------------------------------------------ state import
// Mirror of dbx_world_write_* (include/dbox_b200.h): lets a test / bench.py transplant a device-resident world (same scene
// built on both sides) into the oracle, e.g. the settled 100,000-body pile the oracle would need ten minutes to settle itself.
// No reference counterpart: the records are the persistent state of b2Body (b2body.d:1182-1218), the tree's fat AABBs
// (b2dynamictree.d:146-180), b2Contact (b2contact.d:441-465), the joints' accumulated impulses and the move buffer
// (b2broadphase.d:244-257).
int32_t orc_world_write_bodies(orc_world* w, const dbx_body_state* in, int32_t n) {
  auto& B = w->w.bodiesById;
  if (n > (int)B.size()) return DBX_E_INVALID;
  for (int i = 0; i < n; ++i) {
    Body* b = B[i]; if (!b) continue;
    const dbx_body_state& s = in[i];
    if (s.type != b->type) return DBX_E_INVALID;
    b->flags = (uint16_t)(s.flags & 0x7F);
    b->xf.p = v2(s.p); b->xf.q.s = s.qs; b->xf.q.c = s.qc;
    b->sweep.localCenter = v2(s.localCenter); b->sweep.c0 = v2(s.c0); b->sweep.c = v2(s.c);
    b->sweep.a0 = s.a0; b->sweep.a = s.a; b->sweep.alpha0 = s.alpha0;
    b->linearVelocity = v2(s.v); b->angularVelocity = s.w; b->force = v2(s.force); b->torque = s.torque;
    b->mass = s.mass; b->invMass = s.invMass; b->I = s.I; b->invI = s.invI;
    b->linearDamping = s.linearDamping; b->angularDamping = s.angularDamping; b->gravityScale = s.gravityScale; b->sleepTime = s.sleepTime;
  }
  return n;
}
int32_t orc_world_write_proxies(orc_world* w, const dbx_proxy_rec* in, int32_t n) {
  for (int i = 0; i < n; ++i) {
    const dbx_proxy_rec& r = in[i];
    Fixture* f = fixtureAt(w, r.fixture);
    if (!f || r.child < 0 || r.child >= f->proxyCount) return DBX_E_INVALID;
    FixtureProxy& p = f->proxies[r.child];
    if (p.proxyId != r.proxyId) return DBX_E_INVALID;     // both sides must have lived the same history of proxy creations (b2dynamictree.d:516-564)
    p.aabb.lo = v2(r.aabb.lo); p.aabb.hi = v2(r.aabb.hi);
    AABB fat; fat.lo = v2(r.fat.lo); fat.hi = v2(r.fat.hi);
    w->w.broadPhase.importFatAABB(p.proxyId, fat);
  }
  return n;
}
int32_t orc_world_write_contacts(orc_world* w, const dbx_contact_rec* in, int32_t n) {
  w->w.clearContacts();
  for (int i = 0; i < n; ++i) {
    const dbx_contact_rec& r = in[i];
    Fixture* fA = fixtureAt(w, r.fixtureA); Fixture* fB = fixtureAt(w, r.fixtureB);
    if (!fA || !fB || r.childA < 0 || r.childB < 0 || r.childA >= fA->proxyCount || r.childB >= fB->proxyCount) return DBX_E_INVALID;
    Contact* c = w->w.importContact(fA, r.childA, fB, r.childB);
    c->flags = r.flags & 0x3F;
    for (int k = 0; k < 2; ++k) {
      c->manifold.points[k].localPoint = v2(r.manifold.points[k].localPoint);
      c->manifold.points[k].normalImpulse = r.manifold.points[k].normalImpulse;
      c->manifold.points[k].tangentImpulse = r.manifold.points[k].tangentImpulse;
      c->manifold.points[k].id.key = r.manifold.points[k].key;
    }
    c->manifold.localNormal = v2(r.manifold.localNormal); c->manifold.localPoint = v2(r.manifold.localPoint);
    c->manifold.type = r.manifold.type; c->manifold.pointCount = r.manifold.pointCount;
    c->friction = r.friction; c->restitution = r.restitution; c->tangentSpeed = r.tangentSpeed; c->toiCount = r.toiCount; c->toi = r.toi;
  }
  return n;
}
int32_t orc_world_write_joints(orc_world* w, const dbx_joint_state* in, int32_t n) {
  auto& J = w->w.jointsById;
  if (n > (int)J.size()) return DBX_E_INVALID;
  for (int i = 0; i < n; ++i) {
    Joint* j = J[i]; if (!j) continue;
    const dbx_joint_state* o = in + i;
    if (o->type != j->type) return DBX_E_INVALID;
    if (j->type == jRevolute) { auto* r = (RevoluteJoint*)j; r->impulse.x = o->impulse[0]; r->impulse.y = o->impulse[1]; r->impulse.z = o->impulse[2]; r->motorImpulse = o->motorImpulse; r->limitState = o->limitState; }
    else if (j->type == jDistance) { auto* d = (DistanceJoint*)j; d->impulse = o->impulse[0]; }
    else if (j->type == jRope) { auto* d = (RopeJoint*)j; d->impulse = o->impulse[0]; d->state = o->limitState; }
    else if (j->type == jWeld) { auto* d = (WeldJoint*)j; d->impulse.x = o->impulse[0]; d->impulse.y = o->impulse[1]; d->impulse.z = o->impulse[2]; }
    else if (j->type == jFriction) { auto* d = (FrictionJoint*)j; d->linearImpulse.x = o->impulse[0]; d->linearImpulse.y = o->impulse[1]; d->angularImpulse = o->impulse[2]; }
    else if (j->type == jMotor) { auto* d = (MotorJoint*)j; d->linearImpulse.x = o->impulse[0]; d->linearImpulse.y = o->impulse[1]; d->angularImpulse = o->impulse[2]; }
    else if (j->type == jMouse) { auto* d = (MouseJoint*)j; d->impulse.x = o->impulse[0]; d->impulse.y = o->impulse[1]; }
    else if (j->type == jPrismatic) { auto* d = (PrismaticJoint*)j; d->impulse.x = o->impulse[0]; d->impulse.y = o->impulse[1]; d->impulse.z = o->impulse[2]; d->motorImpulse = o->motorImpulse; d->limitState = o->limitState; }
    else if (j->type == jWheel) { auto* d = (WheelJoint*)j; d->impulse = o->impulse[0]; d->springImpulse = o->impulse[1]; d->motorImpulse = o->motorImpulse; }
    else if (j->type == jPulley) { auto* d = (PulleyJoint*)j; d->impulse = o->impulse[0]; }
    else if (j->type == jGear) { auto* d = (GearJoint*)j; d->impulse = o->impulse[0]; }
  }
  return n;
}
int32_t orc_world_write_moves(orc_world* w, const int32_t* fc, int32_t n) {
  std::vector<int> ids;
  for (int i = 0; i < n; ++i) {
    Fixture* f = fixtureAt(w, fc[2 * i]);
    if (!f || fc[2 * i + 1] < 0 || fc[2 * i + 1] >= f->proxyCount) return DBX_E_INVALID;
    ids.push_back(f->proxies[fc[2 * i + 1]].proxyId);
  }
  w->w.broadPhase.importMoveBuffer(ids);
  w->w.newFixture = false;
  return n;
}
int32_t orc_world_set_inv_dt0(orc_world* w, float v) { w->w.inv_dt0 = v; return 0; }
// Test hook (orc_world.h, World::orderOverride): the NEXT Solve walks every island's joints and contacts as ONE sequence by
// ascending rank -- rank[i] for contact i (identified by keys4[4 i ..] = fixtureA, childA, fixtureB, childB), jointRank[joint id]
// for a joint; at equal rank joints go first; contacts that are not listed go last.  Position iterations walk the same sequence
// forwards (positionMode 0), backwards (1), or its contacts and then its joints (2, the reference's own split).
int32_t orc_world_debug_set_solve_order(orc_world* w, const int32_t* keys4, const int32_t* rank, int32_t n, const int32_t* jointRank, int32_t nj, int32_t positionMode) {
  struct K { int a, b, c, d; bool operator<(const K& o) const { return a != o.a ? a < o.a : b != o.b ? b < o.b : c != o.c ? c < o.c : d < o.d; } };
  std::vector<std::pair<K, int>> tab((size_t)n);
  for (int i = 0; i < n; ++i) tab[i] = {K{keys4[4 * i], keys4[4 * i + 1], keys4[4 * i + 2], keys4[4 * i + 3]}, rank[i]};
  std::sort(tab.begin(), tab.end(), [](const std::pair<K, int>& x, const std::pair<K, int>& y) { return x.first < y.first; });
  int found = 0;
  for (Contact* c = w->w.contactList; c; c = c->next) {
    const K k{c->fixtureA->id, c->indexA, c->fixtureB->id, c->indexB};
    auto it = std::lower_bound(tab.begin(), tab.end(), k, [](const std::pair<K, int>& x, const K& y) { return x.first < y; });
    if (it != tab.end() && !(k < it->first)) { c->orderRank = it->second; ++found; } else c->orderRank = 0x7fffffff;
  }
  auto& J = w->w.jointsById;
  for (int i = 0; i < (int)J.size(); ++i) if (J[i]) J[i]->orderRank = i < nj ? jointRank[i] : 0x7fffffff;
  w->w.orderOverride = true; w->w.orderPositionMode = positionMode;
  return found;
}
int32_t orc_world_get_inv_dt0(orc_world* w, float* out) { *out = w->w.inv_dt0; return 0; }
int32_t orc_world_stage_find_new_contacts(orc_world* w) { w->w.findNewContacts(); w->w.newFixture = false; return 0; }
int32_t orc_world_stage_collide(orc_world* w) { w->w.collide(); return 0; }
int32_t orc_world_tree_height(orc_world* w) { return w->w.broadPhase.treeHeight(); }
int32_t orc_world_tree_validate(orc_world* w) { return w->w.broadPhase.tree().validate() ? 1 : 0; }

// ---- pure-function entry points for known-answer fixtures (polycollision.d, distancetest.d, timeofimpact.d) ----
int32_t orc_collide(const dbx_shape* sA, float ax, float ay, float aa, int32_t childA, const dbx_shape* sB, float bx, float by, float ba, int32_t childB, dbx_manifold* out) {
  Shape A = toShape(sA), B = toShape(sB);
  Fixture fA, fB; fA.shape = A; fB.shape = B;
  Contact c; c.fixtureA = &fA; c.fixtureB = &fB; c.indexA = childA; c.indexB = childB;
  Xf xfA, xfB; xfA.set(V2(ax, ay), aa); xfB.set(V2(bx, by), ba);
  Manifold m;
  c.evaluate(&m, xfA, xfB);
  std::memset(out, 0, sizeof(*out));
  out->type = m.type; out->pointCount = m.pointCount; out->localNormal = d2(m.localNormal); out->localPoint = d2(m.localPoint);
  for (int i = 0; i < m.pointCount; ++i) { out->points[i].localPoint = d2(m.points[i].localPoint); out->points[i].key = m.points[i].id.key; }
  return m.pointCount;
}
// same, with explicit rotation (sin, cos) so device-vs-libm sinf/cosf differences cannot leak into a bit-exact compare
int32_t orc_collide_xf(const dbx_shape* sA, const float* xfa, int32_t childA, const dbx_shape* sB, const float* xfb, int32_t childB, dbx_manifold* out) {
  Shape A = toShape(sA), B = toShape(sB);
  Fixture fA, fB; fA.shape = A; fB.shape = B;
  Contact c; c.fixtureA = &fA; c.fixtureB = &fB; c.indexA = childA; c.indexB = childB;
  Xf xfA, xfB; xfA.p = V2(xfa[0], xfa[1]); xfA.q.s = xfa[2]; xfA.q.c = xfa[3]; xfB.p = V2(xfb[0], xfb[1]); xfB.q.s = xfb[2]; xfB.q.c = xfb[3];
  Manifold m;
  c.evaluate(&m, xfA, xfB);
  std::memset(out, 0, sizeof(*out));
  out->type = m.type; out->pointCount = m.pointCount; out->localNormal = d2(m.localNormal); out->localPoint = d2(m.localPoint);
  for (int i = 0; i < m.pointCount; ++i) { out->points[i].localPoint = d2(m.points[i].localPoint); out->points[i].key = m.points[i].id.key; }
  return m.pointCount;
}
float orc_distance(const dbx_shape* sA, float ax, float ay, float aa, int32_t childA, const dbx_shape* sB, float bx, float by, float ba, int32_t childB,
                   int32_t useRadii, dbx_vec2* pA, dbx_vec2* pB, int32_t* iterations) {
  Shape A = toShape(sA), B = toShape(sB);
  DistanceInput in; in.proxyA.set(A, childA); in.proxyB.set(B, childB);
  in.transformA.set(V2(ax, ay), aa); in.transformB.set(V2(bx, by), ba); in.useRadii = useRadii != 0;
  SimplexCache cache; cache.count = 0;
  DistanceOutput o; distance(&o, &cache, &in);
  *pA = d2(o.pointA); *pB = d2(o.pointB); *iterations = o.iterations;
  return o.distance;
}
// sweeps given as {localCenter.x, localCenter.y, c0.x, c0.y, c.x, c.y, a0, a, alpha0}
int32_t orc_time_of_impact(const dbx_shape* sA, const float* sweepA, int32_t childA, const dbx_shape* sB, const float* sweepB, int32_t childB, float tMax, float* t) {
  Shape A = toShape(sA), B = toShape(sB);
  TOIInput in; in.proxyA.set(A, childA); in.proxyB.set(B, childB);
  auto rd = [](const float* s) { Sweep w; w.localCenter = V2(s[0], s[1]); w.c0 = V2(s[2], s[3]); w.c = V2(s[4], s[5]); w.a0 = s[6]; w.a = s[7]; w.alpha0 = s[8]; return w; };
  in.sweepA = rd(sweepA); in.sweepB = rd(sweepB); in.tMax = tMax;
  TOIOutput o; timeOfImpact(&o, &in);
  *t = o.t;
  return o.state;
}

// ---- CPU baseline helper: step `nWorlds` independent worlds `steps` times on `threads` host threads, one world per
// thread at a time.  Returns wall seconds of the stepping only. ----
double orc_batch_step(orc_world** worlds, int32_t nWorlds, float dt, int32_t vi, int32_t pi, int32_t steps, int32_t threads) {
  auto t0 = std::chrono::steady_clock::now();
  if (threads <= 1) {
    for (int i = 0; i < nWorlds; ++i) for (int s = 0; s < steps; ++s) worlds[i]->w.step(dt, vi, pi);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
      pool.emplace_back([=]() { for (int i = t; i < nWorlds; i += threads) for (int s = 0; s < steps; ++s) worlds[i]->w.step(dt, vi, pi); });
    for (auto& th : pool) th.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
