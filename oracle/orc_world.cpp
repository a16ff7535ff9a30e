// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
// Restates the reference's dynamics layer; each function cites the lines it follows.
#include "orc_world.h"
#include <chrono>
#include <cstring>

namespace orc {

static double nowMs() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------ Body
void Body::setAwake(bool flag) {
  if (flag) {
    if ((flags & bAwake) == 0) { flags |= bAwake; sleepTime = 0.0f; }
  } else {
    flags &= ~bAwake;
    sleepTime = 0.0f;
    linearVelocity = V2(0, 0);
    angularVelocity = 0.0f;
    force = V2(0, 0);
    torque = 0.0f;
  }
}
void Body::synchronizeTransform() {
  xf.q.set(sweep.a);
  xf.p = sweep.c - mul(xf.q, sweep.localCenter);
}
// b2fixture.d:480-502
static void fixtureSynchronize(Fixture* f, BroadPhase* bp, const Xf& xf1, const Xf& xf2) {
  if (f->proxyCount == 0) return;
  for (int i = 0; i < f->proxyCount; ++i) {
    FixtureProxy* proxy = &f->proxies[i];
    AABB aabb1, aabb2;
    f->shape.computeAABB(&aabb1, xf1, proxy->childIndex);
    f->shape.computeAABB(&aabb2, xf2, proxy->childIndex);
    proxy->aabb.combine(aabb1, aabb2);
    V2 displacement = xf2.p - xf1.p;
    bp->moveProxy(proxy->proxyId, proxy->aabb, displacement);
  }
}
void Body::synchronizeFixtures() {
  Xf xf1;
  xf1.q.set(sweep.a0);
  xf1.p = sweep.c0 - mul(xf1.q, sweep.localCenter);
  for (Fixture* f = fixtureList; f; f = f->next) fixtureSynchronize(f, &world->broadPhase, xf1, xf);
}
void Body::advance(float alpha) {
  sweep.advance(alpha);
  sweep.c = sweep.c0;
  sweep.a = sweep.a0;
  xf.q.set(sweep.a);
  xf.p = sweep.c - mul(xf.q, sweep.localCenter);
}
bool Body::shouldCollide(const Body* other) const {
  if (type != kDynamic && other->type != kDynamic) return false;
  for (JointEdge* jn = jointList; jn; jn = jn->next) {
    if (jn->other == other) {
      if (jn->joint->collideConnected == false) return false;
    }
  }
  return true;
}
void Body::resetMassData() {
  mass = 0.0f; invMass = 0.0f; I = 0.0f; invI = 0.0f;
  sweep.localCenter = V2(0, 0);
  if (type == kStatic || type == kKinematic) {
    sweep.c0 = xf.p; sweep.c = xf.p; sweep.a0 = sweep.a;
    return;
  }
  V2 localCenter(0, 0);
  for (Fixture* f = fixtureList; f; f = f->next) {
    if (f->density == 0.0f) continue;
    MassData md;
    f->shape.computeMass(&md, f->density);
    mass += md.mass;
    localCenter += md.mass * md.center;
    I += md.I;
  }
  if (mass > 0.0f) { invMass = 1.0f / mass; localCenter *= invMass; }
  else { mass = 1.0f; invMass = 1.0f; }
  if (I > 0.0f && (flags & bFixedRotation) == 0) {
    I -= mass * dot(localCenter, localCenter);
    invI = 1.0f / I;
  } else { I = 0.0f; invI = 0.0f; }
  V2 oldCenter = sweep.c;
  sweep.localCenter = localCenter;
  sweep.c0 = sweep.c = mul(xf, sweep.localCenter);
  linearVelocity += cross(angularVelocity, sweep.c - oldCenter);
}

// ------------------------------------------------------------------ Contact
// Dispatch table of b2contact.d:425-437: after Create's swap, typeA is the "primary" type.
void Contact::evaluate(Manifold* m, const Xf& xfA, const Xf& xfB) const {
  const Shape& sA = fixtureA->shape;
  const Shape& sB = fixtureB->shape;
  int tA = sA.type, tB = sB.type;
  if (tA == kCircle && tB == kCircle) collideCircles(m, sA, xfA, sB, xfB);                       // b2circlecontact.d:56
  else if (tA == kPolygon && tB == kCircle) collidePolygonAndCircle(m, sA, xfA, sB, xfB);         // b2polygonandcirclecontact.d:56
  else if (tA == kPolygon && tB == kPolygon) collidePolygons(m, sA, xfA, sB, xfB);                // b2polygoncontact.d:56
  else if (tA == kEdge && tB == kCircle) collideEdgeAndCircle(m, sA, xfA, sB, xfB);               // b2edgeandcirclecontact.d:56
  else if (tA == kEdge && tB == kPolygon) collideEdgeAndPolygon(m, sA, xfA, sB, xfB);             // b2edgeandpolygoncontact.d:56
  else if (tA == kChain && tB == kCircle) { Shape e; sA.childEdge(&e, indexA); collideEdgeAndCircle(m, e, xfA, sB, xfB); }   // b2chainandcirclecontact.d:56-66
  else if (tA == kChain && tB == kPolygon) { Shape e; sA.childEdge(&e, indexA); collideEdgeAndPolygon(m, e, xfA, sB, xfB); } // b2chainandpolygoncontact.d:56-66
}

void Contact::update(World* w) {
  Manifold oldManifold = manifold;
  flags |= cEnabled;
  if (!w || w->evPhase != 2) flags &= ~cPreSolveOff;       // the decision of a step's PreSolve ends with that step
  bool touching = false;
  bool wasTouching = (flags & cTouching) == cTouching;
  bool sensor = fixtureA->isSensor || fixtureB->isSensor;
  Body* bodyA = fixtureA->body;
  Body* bodyB = fixtureB->body;
  const Xf xfA = bodyA->xf, xfB = bodyB->xf;
  if (sensor) {
    touching = testOverlap(fixtureA->shape, indexA, fixtureB->shape, indexB, xfA, xfB);
    manifold.pointCount = 0;
  } else {
    evaluate(&manifold, xfA, xfB);
    touching = manifold.pointCount > 0;
    for (int i = 0; i < manifold.pointCount; ++i) {
      ManifoldPoint* mp2 = manifold.points + i;
      mp2->normalImpulse = 0.0f;
      mp2->tangentImpulse = 0.0f;
      ContactID id2 = mp2->id;
      for (int j = 0; j < oldManifold.pointCount; ++j) {
        ManifoldPoint* mp1 = oldManifold.points + j;
        if (mp1->id.key == id2.key) {
          mp2->normalImpulse = mp1->normalImpulse;
          mp2->tangentImpulse = mp1->tangentImpulse;
          break;
        }
      }
    }
    if (touching != wasTouching) { bodyA->setAwake(true); bodyB->setAwake(true); }
  }
  if (touching) flags |= cTouching; else flags &= ~cTouching;
  if (!sensor && touching && (flags & cPreSolveOff)) flags &= ~cEnabled;       // PreSolve (b2contact.d:348-355), see cPreSolveOff
  // listener callbacks (b2contact.d:338-355): BeginContact / EndContact are logged, PreSolve is the default no-op
  if (w && wasTouching == false && touching == true) w->logContactEvent(1, this);
  if (w && wasTouching == true && touching == false) w->logContactEvent(2, this);
}

void World::logContactEvent(int type, const Contact* c) {
  if (!recordContactEvents) return;
  contactEvents.push_back({type, evPhase, stepCount + (evPhase == 3 ? 0 : 1), c->fixtureA->id, c->indexA, c->fixtureB->id, c->indexB, c->fixtureA->body->id, c->fixtureB->body->id});
}

// b2contact.d:375-400 + ctor :208-239
static Contact* createContact(Fixture* fA, int iA, Fixture* fB, int iB) {
  static const struct { bool has, primary; } reg[4][4] = {
      /* circle  */ {{true, true}, {true, false}, {true, false}, {true, false}},
      /* edge    */ {{true, true}, {false, false}, {true, true}, {false, false}},
      /* polygon */ {{true, true}, {true, false}, {true, true}, {true, false}},
      /* chain   */ {{true, true}, {false, false}, {true, true}, {false, false}},
  };
  int t1 = fA->shape.type, t2 = fB->shape.type;
  if (!reg[t1][t2].has) return nullptr;
  Contact* c = new Contact();
  if (reg[t1][t2].primary) { c->fixtureA = fA; c->indexA = iA; c->fixtureB = fB; c->indexB = iB; }
  else { c->fixtureA = fB; c->indexA = iB; c->fixtureB = fA; c->indexB = iA; }
  c->flags = cEnabled;
  c->manifold.pointCount = 0;
  c->toiCount = 0;
  c->friction = sqrtf(c->fixtureA->friction * c->fixtureB->friction);                        // b2contact.d:32-35
  c->restitution = c->fixtureA->restitution > c->fixtureB->restitution ? c->fixtureA->restitution : c->fixtureB->restitution;  // :39-42
  c->tangentSpeed = 0.0f;
  return c;
}

// ------------------------------------------------------------------ World: lifecycle
World::World(V2 g) : gravity(g) {}
World::~World() {
  while (jointList) { Joint* j = jointList; jointList = j->next; delete j; }
  while (contactList) { Contact* c = contactList; contactList = c->next; delete c; }
  while (bodyList) {
    Body* b = bodyList; bodyList = b->next;
    while (b->fixtureList) { Fixture* f = b->fixtureList; b->fixtureList = f->next; delete f; }
    delete b;
  }
}

Body* World::createBody(const BodyDef& bd) {
  if (locked) return nullptr;
  Body* b = new Body();
  b->id = (int)bodiesById.size();
  bodiesById.push_back(b);
  // b2body.d:1030-1115
  b->flags = 0;
  if (bd.bullet) b->flags |= bBullet;
  if (bd.fixedRotation) b->flags |= bFixedRotation;
  if (bd.allowSleep) b->flags |= bAutoSleep;
  if (bd.awake) b->flags |= bAwake;
  if (bd.active) b->flags |= bActive;
  b->world = this;
  b->xf.p = bd.position;
  b->xf.q.set(bd.angle);
  b->sweep.localCenter = V2(0, 0);
  b->sweep.c0 = b->xf.p; b->sweep.c = b->xf.p;
  b->sweep.a0 = bd.angle; b->sweep.a = bd.angle; b->sweep.alpha0 = 0.0f;
  b->linearVelocity = bd.linearVelocity; b->angularVelocity = bd.angularVelocity;
  b->linearDamping = bd.linearDamping; b->angularDamping = bd.angularDamping; b->gravityScale = bd.gravityScale;
  b->type = bd.type;
  if (b->type == kDynamic) { b->mass = 1.0f; b->invMass = 1.0f; } else { b->mass = 0.0f; b->invMass = 0.0f; }
  b->userData = bd.userData;
  b->prev = nullptr; b->next = bodyList;
  if (bodyList) bodyList->prev = b;
  bodyList = b;
  ++bodyCount;
  return b;
}

Fixture* World::createFixture(Body* body, const FixtureDef& def) {
  if (locked) return nullptr;
  Fixture* f = new Fixture();
  f->id = (int)fixturesById.size();
  fixturesById.push_back(f);
  // b2fixture.d:367-395
  f->userData = def.userData; f->friction = def.friction; f->restitution = def.restitution;
  f->body = body; f->next = nullptr; f->filter = def.filter; f->isSensor = def.isSensor;
  f->shape = *def.shape;
  int childCount = f->shape.childCount();
  f->proxies.resize(childCount);
  f->proxyCount = 0;
  f->density = def.density;
  if (body->flags & bActive) {
    // b2fixture.d:450-465
    f->proxyCount = childCount;
    for (int i = 0; i < f->proxyCount; ++i) {
      FixtureProxy* proxy = &f->proxies[i];
      f->shape.computeAABB(&proxy->aabb, body->xf, i);
      proxy->proxyId = broadPhase.createProxy(proxy->aabb, proxy);
      proxy->fixture = f;
      proxy->childIndex = i;
    }
  }
  f->next = body->fixtureList;
  body->fixtureList = f;
  ++body->fixtureCount;
  if (f->density > 0.0f) body->resetMassData();
  newFixture = true;
  return f;
}

void World::destroyFixture(Fixture* fixture) {
  if (locked) return;
  Body* b = fixture->body;
  Fixture** node = &b->fixtureList;
  while (*node) {
    if (*node == fixture) { *node = fixture->next; break; }
    node = &(*node)->next;
  }
  ContactEdge* edge = b->contactList;
  while (edge) {
    Contact* c = edge->contact;
    edge = edge->next;
    if (fixture == c->fixtureA || fixture == c->fixtureB) destroyContact(c);
  }
  if (b->flags & bActive) {
    for (int i = 0; i < fixture->proxyCount; ++i) { broadPhase.destroyProxy(fixture->proxies[i].proxyId); fixture->proxies[i].proxyId = -1; }
    fixture->proxyCount = 0;
  }
  fixturesById[fixture->id] = nullptr;
  delete fixture;
  --b->fixtureCount;
  b->resetMassData();
}

// b2Body.SetType (b2body.d:867-914)
void World::setBodyType(Body* b, int type) {
  if (locked) return;
  if (b->type == type) return;
  b->type = type;
  b->resetMassData();
  if (b->type == kStatic) {
    b->linearVelocity = V2(0, 0); b->angularVelocity = 0.0f;
    b->sweep.a0 = b->sweep.a; b->sweep.c0 = b->sweep.c;
    b->synchronizeFixtures();
  }
  b->setAwake(true);
  b->force = V2(0, 0); b->torque = 0.0f;
  ContactEdge* ce = b->contactList;
  while (ce) { ContactEdge* ce0 = ce; ce = ce->next; destroyContact(ce0->contact); }
  b->contactList = nullptr;
  for (Fixture* f = b->fixtureList; f; f = f->next)
    for (int i = 0; i < f->proxyCount; ++i) broadPhase.touchProxy(f->proxies[i].proxyId);
}
// b2Body.SetActive (b2body.d:718-775)
void World::setBodyActive(Body* b, bool flag) {
  if (flag == ((b->flags & bActive) == bActive)) return;
  if (flag) {
    b->flags |= bActive;
    for (Fixture* f = b->fixtureList; f; f = f->next) {
      f->proxyCount = f->shape.childCount();
      for (int i = 0; i < f->proxyCount; ++i) {
        FixtureProxy* proxy = &f->proxies[i];
        f->shape.computeAABB(&proxy->aabb, b->xf, i);
        proxy->proxyId = broadPhase.createProxy(proxy->aabb, proxy);
        proxy->fixture = f; proxy->childIndex = i;
      }
    }
  } else {
    b->flags &= ~bActive;
    for (Fixture* f = b->fixtureList; f; f = f->next) {
      for (int i = 0; i < f->proxyCount; ++i) { broadPhase.destroyProxy(f->proxies[i].proxyId); f->proxies[i].proxyId = -1; }
      f->proxyCount = 0;
    }
    ContactEdge* ce = b->contactList;
    while (ce) { ContactEdge* ce0 = ce; ce = ce->next; destroyContact(ce0->contact); }
    b->contactList = nullptr;
  }
}

void World::destroyBody(Body* b) {
  if (locked) return;
  JointEdge* je = b->jointList;
  while (je) { JointEdge* je0 = je; je = je->next; destroyJoint(je0->joint); b->jointList = je; }
  b->jointList = nullptr;
  ContactEdge* ce = b->contactList;
  while (ce) { ContactEdge* ce0 = ce; ce = ce->next; destroyContact(ce0->contact); }
  b->contactList = nullptr;
  Fixture* f = b->fixtureList;
  while (f) {
    Fixture* f0 = f; f = f->next;
    for (int i = 0; i < f0->proxyCount; ++i) broadPhase.destroyProxy(f0->proxies[i].proxyId);
    fixturesById[f0->id] = nullptr;
    delete f0;
    b->fixtureList = f; b->fixtureCount -= 1;
  }
  if (b->prev) b->prev->next = b->next;
  if (b->next) b->next->prev = b->prev;
  if (b == bodyList) bodyList = b->next;
  --bodyCount;
  bodiesById[b->id] = nullptr;
  delete b;
}

Joint* World::addJoint(Joint* j) {
  if (locked) { delete j; return nullptr; }
  j->id = (int)jointsById.size();
  jointsById.push_back(j);
  j->prev = nullptr; j->next = jointList;
  if (jointList) jointList->prev = j;
  jointList = j;
  ++jointCount;
  j->edgeA.joint = j; j->edgeA.other = j->bodyB; j->edgeA.prev = nullptr; j->edgeA.next = j->bodyA->jointList;
  if (j->bodyA->jointList) j->bodyA->jointList->prev = &j->edgeA;
  j->bodyA->jointList = &j->edgeA;
  j->edgeB.joint = j; j->edgeB.other = j->bodyA; j->edgeB.prev = nullptr; j->edgeB.next = j->bodyB->jointList;
  if (j->bodyB->jointList) j->bodyB->jointList->prev = &j->edgeB;
  j->bodyB->jointList = &j->edgeB;
  Body* bodyA = j->bodyA; Body* bodyB = j->bodyB;
  if (j->collideConnected == false) {
    for (ContactEdge* edge = bodyB->contactList; edge; edge = edge->next)
      if (edge->other == bodyA) edge->contact->flags |= cFilter;
  }
  return j;
}

void World::destroyJoint(Joint* j) {
  if (locked) return;
  bool collideConnected = j->collideConnected;
  if (j->prev) j->prev->next = j->next;
  if (j->next) j->next->prev = j->prev;
  if (j == jointList) jointList = j->next;
  Body* bodyA = j->bodyA; Body* bodyB = j->bodyB;
  bodyA->setAwake(true);
  bodyB->setAwake(true);
  if (j->edgeA.prev) j->edgeA.prev->next = j->edgeA.next;
  if (j->edgeA.next) j->edgeA.next->prev = j->edgeA.prev;
  if (&j->edgeA == bodyA->jointList) bodyA->jointList = j->edgeA.next;
  if (j->edgeB.prev) j->edgeB.prev->next = j->edgeB.next;
  if (j->edgeB.next) j->edgeB.next->prev = j->edgeB.prev;
  if (&j->edgeB == bodyB->jointList) bodyB->jointList = j->edgeB.next;
  jointsById[j->id] = nullptr;
  delete j;
  --jointCount;
  if (collideConnected == false) {
    for (ContactEdge* edge = bodyB->contactList; edge; edge = edge->next)
      if (edge->other == bodyA) edge->contact->flags |= cFilter;
  }
}

// ------------------------------------------------------------------ contact manager
// b2worldcallbacks.d:52-64
static bool defaultShouldCollide(const Fixture* fA, const Fixture* fB) {
  const Filter& a = fA->filter; const Filter& b = fB->filter;
  if (a.groupIndex == b.groupIndex && a.groupIndex != 0) return a.groupIndex > 0;
  return (a.maskBits & b.categoryBits) != 0 && (a.categoryBits & b.maskBits) != 0;
}

// b2contactmanager.d:52-176
void World::addPair(void* udA, void* udB) {
  FixtureProxy* proxyA = (FixtureProxy*)udA;
  FixtureProxy* proxyB = (FixtureProxy*)udB;
  lastPairs.push_back({proxyA, proxyB});
  Fixture* fixtureA = proxyA->fixture; Fixture* fixtureB = proxyB->fixture;
  int indexA = proxyA->childIndex, indexB = proxyB->childIndex;
  Body* bodyA = fixtureA->body; Body* bodyB = fixtureB->body;
  if (bodyA == bodyB) return;
  for (ContactEdge* edge = bodyB->contactList; edge; edge = edge->next) {
    if (edge->other == bodyA) {
      Fixture* fA = edge->contact->fixtureA; Fixture* fB = edge->contact->fixtureB;
      int iA = edge->contact->indexA, iB = edge->contact->indexB;
      if (fA == fixtureA && fB == fixtureB && iA == indexA && iB == indexB) return;
      if (fA == fixtureB && fB == fixtureA && iA == indexB && iB == indexA) return;
    }
  }
  if (bodyB->shouldCollide(bodyA) == false) return;
  if ((userFilter ? userFilter(fixtureA->id, fixtureB->id, defaultShouldCollide(fixtureA, fixtureB)) != 0 : defaultShouldCollide(fixtureA, fixtureB)) == false) return;
  Contact* c = createContact(fixtureA, indexA, fixtureB, indexB);
  if (c == nullptr) return;
  fixtureA = c->fixtureA; fixtureB = c->fixtureB;
  bodyA = fixtureA->body; bodyB = fixtureB->body;
  c->prev = nullptr; c->next = contactList;
  if (contactList) contactList->prev = c;
  contactList = c;
  c->nodeA.contact = c; c->nodeA.other = bodyB; c->nodeA.prev = nullptr; c->nodeA.next = bodyA->contactList;
  if (bodyA->contactList) bodyA->contactList->prev = &c->nodeA;
  bodyA->contactList = &c->nodeA;
  c->nodeB.contact = c; c->nodeB.other = bodyA; c->nodeB.prev = nullptr; c->nodeB.next = bodyB->contactList;
  if (bodyB->contactList) bodyB->contactList->prev = &c->nodeB;
  bodyB->contactList = &c->nodeB;
  if (fixtureA->isSensor == false && fixtureB->isSensor == false) { bodyA->setAwake(true); bodyB->setAwake(true); }
  ++contactCount;
}

void World::findNewContacts() {
  lastPairs.clear();
  broadPhase.updatePairs([this](void* a, void* b) { addPair(a, b); });
}

// State import (tests, bench transplant; no reference counterpart): link a contact exactly as recorded -- fixture A/B as
// given (the record already carries the type-registry order of b2contact.d:375-400), no filter test, no wake-up -- with
// the same push-front list surgery as AddPair (b2contactmanager.d:134-166).
Contact* World::importContact(Fixture* fA, int iA, Fixture* fB, int iB) {
  Contact* c = new Contact();
  c->fixtureA = fA; c->indexA = iA; c->fixtureB = fB; c->indexB = iB;
  Body* bodyA = fA->body; Body* bodyB = fB->body;
  c->prev = nullptr; c->next = contactList;
  if (contactList) contactList->prev = c;
  contactList = c;
  c->nodeA.contact = c; c->nodeA.other = bodyB; c->nodeA.prev = nullptr; c->nodeA.next = bodyA->contactList;
  if (bodyA->contactList) bodyA->contactList->prev = &c->nodeA;
  bodyA->contactList = &c->nodeA;
  c->nodeB.contact = c; c->nodeB.other = bodyA; c->nodeB.prev = nullptr; c->nodeB.next = bodyB->contactList;
  if (bodyB->contactList) bodyB->contactList->prev = &c->nodeB;
  bodyB->contactList = &c->nodeB;
  ++contactCount;
  return c;
}
void World::clearContacts() {
  while (contactList) { Contact* c = contactList; contactList = c->next; delete c; }
  contactCount = 0;
  for (Body* b = bodyList; b; b = b->next) b->contactList = nullptr;
  lastSolveOrder.clear();
}

// b2contactmanager.d:183-246 + b2contact.d:402-423
void World::destroyContact(Contact* c) {
  Fixture* fixtureA = c->fixtureA; Fixture* fixtureB = c->fixtureB;
  Body* bodyA = fixtureA->body; Body* bodyB = fixtureB->body;
  if (c->isTouching()) logContactEvent(2, c);   // b2contactmanager.d:60-63
  if (c->prev) c->prev->next = c->next;
  if (c->next) c->next->prev = c->prev;
  if (c == contactList) contactList = c->next;
  if (c->nodeA.prev) c->nodeA.prev->next = c->nodeA.next;
  if (c->nodeA.next) c->nodeA.next->prev = c->nodeA.prev;
  if (&c->nodeA == bodyA->contactList) bodyA->contactList = c->nodeA.next;
  if (c->nodeB.prev) c->nodeB.prev->next = c->nodeB.next;
  if (c->nodeB.next) c->nodeB.next->prev = c->nodeB.prev;
  if (&c->nodeB == bodyB->contactList) bodyB->contactList = c->nodeB.next;
  if (c->manifold.pointCount > 0 && fixtureA->isSensor == false && fixtureB->isSensor == false) {
    bodyA->setAwake(true);
    bodyB->setAwake(true);
  }
  delete c;
  --contactCount;
}

// b2contactmanager.d:251-317
void World::collide() {
  Contact* c = contactList;
  while (c) {
    Fixture* fixtureA = c->fixtureA; Fixture* fixtureB = c->fixtureB;
    int indexA = c->indexA, indexB = c->indexB;
    Body* bodyA = fixtureA->body; Body* bodyB = fixtureB->body;
    if (c->flags & cFilter) {
      if (bodyB->shouldCollide(bodyA) == false) { Contact* n = c; c = n->next; destroyContact(n); continue; }
      if ((userFilter ? userFilter(fixtureA->id, fixtureB->id, defaultShouldCollide(fixtureA, fixtureB)) != 0 : defaultShouldCollide(fixtureA, fixtureB)) == false) { Contact* n = c; c = n->next; destroyContact(n); continue; }
      c->flags &= ~cFilter;
    }
    bool activeA = bodyA->isAwake() && bodyA->type != kStatic;
    bool activeB = bodyB->isAwake() && bodyB->type != kStatic;
    if (activeA == false && activeB == false) { c = c->next; continue; }
    int proxyIdA = fixtureA->proxies[indexA].proxyId;
    int proxyIdB = fixtureB->proxies[indexB].proxyId;
    bool ov = broadPhase.testOverlap(proxyIdA, proxyIdB);
    if (ov == false) { Contact* n = c; c = n->next; destroyContact(n); continue; }
    c->update(this);
    c = c->next;
  }
}

// ------------------------------------------------------------------ contact solver (b2contactsolver.d)
namespace {
struct VelocityConstraintPoint { V2 rA, rB; float normalImpulse = 0, tangentImpulse = 0, normalMass = 0, tangentMass = 0, velocityBias = 0; };
struct ContactVelocityConstraint {
  VelocityConstraintPoint points[kMaxManifoldPoints];
  V2 normal; M22 normalMass; M22 K;
  int indexA, indexB; float invMassA, invMassB, invIA, invIB, friction, restitution, tangentSpeed;
  int pointCount, contactIndex;
};
struct ContactPositionConstraint {
  V2 localPoints[kMaxManifoldPoints]; V2 localNormal, localPoint;
  int indexA, indexB; float invMassA, invMassB; V2 localCenterA, localCenterB; float invIA, invIB;
  int type; float radiusA, radiusB; int pointCount;
};
// b2contactsolver.d:816-868
struct PositionSolverManifold {
  V2 normal, point; float separation = 0;
  void initialize(const ContactPositionConstraint* pc, const Xf& xfA, const Xf& xfB, int index) {
    switch (pc->type) {
      case kManCircles: {
        V2 pointA = mul(xfA, pc->localPoint);
        V2 pointB = mul(xfB, pc->localPoints[0]);
        normal = pointB - pointA;
        normal.normalize();
        point = 0.5f * (pointA + pointB);
        separation = dot(pointB - pointA, normal) - pc->radiusA - pc->radiusB;
      } break;
      case kManFaceA: {
        normal = mul(xfA.q, pc->localNormal);
        V2 planePoint = mul(xfA, pc->localPoint);
        V2 clipPoint = mul(xfB, pc->localPoints[index]);
        separation = dot(clipPoint - planePoint, normal) - pc->radiusA - pc->radiusB;
        point = clipPoint;
      } break;
      case kManFaceB: {
        normal = mul(xfB.q, pc->localNormal);
        V2 planePoint = mul(xfB, pc->localPoint);
        V2 clipPoint = mul(xfA, pc->localPoints[index]);
        separation = dot(clipPoint - planePoint, normal) - pc->radiusA - pc->radiusB;
        point = clipPoint;
        normal = -normal;
      } break;
    }
  }
};

struct ContactSolver {
  TimeStep step;
  Position* positions; Velocity* velocities;
  std::vector<ContactPositionConstraint> pcs;
  std::vector<ContactVelocityConstraint> vcs;
  Contact** contacts; int count;

  // b2contactsolver.d:244-329
  ContactSolver(const TimeStep& st, Contact** cs, int n, Position* p, Velocity* v) : step(st), positions(p), velocities(v), contacts(cs), count(n) {
    pcs.resize(n); vcs.resize(n);
    for (int i = 0; i < count; ++i) {
      Contact* contact = contacts[i];
      Fixture* fixtureA = contact->fixtureA; Fixture* fixtureB = contact->fixtureB;
      float radiusA = fixtureA->shape.radius, radiusB = fixtureB->shape.radius;
      Body* bodyA = fixtureA->body; Body* bodyB = fixtureB->body;
      Manifold* manifold = &contact->manifold;
      int pointCount = manifold->pointCount;
      ContactVelocityConstraint* vc = &vcs[i];
      vc->friction = contact->friction; vc->restitution = contact->restitution; vc->tangentSpeed = contact->tangentSpeed;
      vc->indexA = bodyA->islandIndex; vc->indexB = bodyB->islandIndex;
      vc->invMassA = bodyA->invMass; vc->invMassB = bodyB->invMass; vc->invIA = bodyA->invI; vc->invIB = bodyB->invI;
      vc->contactIndex = i; vc->pointCount = pointCount;
      vc->K = M22(); vc->normalMass = M22();
      ContactPositionConstraint* pc = &pcs[i];
      pc->indexA = bodyA->islandIndex; pc->indexB = bodyB->islandIndex;
      pc->invMassA = bodyA->invMass; pc->invMassB = bodyB->invMass;
      pc->localCenterA = bodyA->sweep.localCenter; pc->localCenterB = bodyB->sweep.localCenter;
      pc->invIA = bodyA->invI; pc->invIB = bodyB->invI;
      pc->localNormal = manifold->localNormal; pc->localPoint = manifold->localPoint;
      pc->pointCount = pointCount; pc->radiusA = radiusA; pc->radiusB = radiusB; pc->type = manifold->type;
      for (int j = 0; j < pointCount; ++j) {
        ManifoldPoint* cp = manifold->points + j;
        VelocityConstraintPoint* vcp = vc->points + j;
        if (step.warmStarting) {
          vcp->normalImpulse = step.dtRatio * cp->normalImpulse;
          vcp->tangentImpulse = step.dtRatio * cp->tangentImpulse;
        } else { vcp->normalImpulse = 0.0f; vcp->tangentImpulse = 0.0f; }
        vcp->rA = V2(0, 0); vcp->rB = V2(0, 0);
        vcp->normalMass = 0.0f; vcp->tangentMass = 0.0f; vcp->velocityBias = 0.0f;
        pc->localPoints[j] = cp->localPoint;
      }
    }
  }

  // b2contactsolver.d:338-450
  void initializeVelocityConstraints() {
    for (int i = 0; i < count; ++i) {
      ContactVelocityConstraint* vc = &vcs[i];
      ContactPositionConstraint* pc = &pcs[i];
      float radiusA = pc->radiusA, radiusB = pc->radiusB;
      Manifold* manifold = &contacts[vc->contactIndex]->manifold;
      int indexA = vc->indexA, indexB = vc->indexB;
      float mA = vc->invMassA, mB = vc->invMassB, iA = vc->invIA, iB = vc->invIB;
      V2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
      V2 cA = positions[indexA].c; float aA = positions[indexA].a;
      V2 vA = velocities[indexA].v; float wA = velocities[indexA].w;
      V2 cB = positions[indexB].c; float aB = positions[indexB].a;
      V2 vB = velocities[indexB].v; float wB = velocities[indexB].w;
      Xf xfA, xfB;
      xfA.q.set(aA); xfB.q.set(aB);
      xfA.p = cA - mul(xfA.q, localCenterA);
      xfB.p = cB - mul(xfB.q, localCenterB);
      WorldManifold worldManifold;
      worldManifold.initialize(manifold, xfA, radiusA, xfB, radiusB);
      vc->normal = worldManifold.normal;
      int pointCount = vc->pointCount;
      for (int j = 0; j < pointCount; ++j) {
        VelocityConstraintPoint* vcp = vc->points + j;
        vcp->rA = worldManifold.points[j] - cA;
        vcp->rB = worldManifold.points[j] - cB;
        float rnA = cross(vcp->rA, vc->normal), rnB = cross(vcp->rB, vc->normal);
        float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        vcp->normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
        V2 tangent = cross(vc->normal, 1.0f);
        float rtA = cross(vcp->rA, tangent), rtB = cross(vcp->rB, tangent);
        float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
        vcp->tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
        vcp->velocityBias = 0.0f;
        float vRel = dot(vc->normal, vB + cross(wB, vcp->rB) - vA - cross(wA, vcp->rA));
        if (vRel < -kVelocityThreshold) vcp->velocityBias = -vc->restitution * vRel;
      }
      if (vc->pointCount == 2) {  // g_blockSolve == true (b2contactsolver.d:799)
        VelocityConstraintPoint* vcp1 = vc->points + 0;
        VelocityConstraintPoint* vcp2 = vc->points + 1;
        float rn1A = cross(vcp1->rA, vc->normal), rn1B = cross(vcp1->rB, vc->normal);
        float rn2A = cross(vcp2->rA, vc->normal), rn2B = cross(vcp2->rB, vc->normal);
        float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
        float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
        float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
        const float k_maxConditionNumber = 1000.0f;
        if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
          vc->K.ex = V2(k11, k12);
          vc->K.ey = V2(k12, k22);
          vc->normalMass = vc->K.inverse();
        } else {
          vc->pointCount = 1;
        }
      }
    }
  }

  // b2contactsolver.d:452-490
  void warmStart() {
    for (int i = 0; i < count; ++i) {
      ContactVelocityConstraint* vc = &vcs[i];
      int indexA = vc->indexA, indexB = vc->indexB;
      float mA = vc->invMassA, iA = vc->invIA, mB = vc->invMassB, iB = vc->invIB;
      int pointCount = vc->pointCount;
      V2 vA = velocities[indexA].v; float wA = velocities[indexA].w;
      V2 vB = velocities[indexB].v; float wB = velocities[indexB].w;
      V2 normal = vc->normal;
      V2 tangent = cross(normal, 1.0f);
      for (int j = 0; j < pointCount; ++j) {
        VelocityConstraintPoint* vcp = vc->points + j;
        V2 P = vcp->normalImpulse * normal + vcp->tangentImpulse * tangent;
        wA -= iA * cross(vcp->rA, P);
        vA -= mA * P;
        wB += iB * cross(vcp->rB, P);
        vB += mB * P;
      }
      velocities[indexA].v = vA; velocities[indexA].w = wA;
      velocities[indexB].v = vB; velocities[indexB].w = wB;
    }
  }

  // b2contactsolver.d:492-772
  void solveVelocityConstraints() { solveVelocityRange(0, count); }
  // constraints [i0, i1) of the loop above (:492-772); a single constraint at a time serves the merged-order test hook
  void solveVelocityRange(int i0, int i1) {
    for (int i = i0; i < i1; ++i) {
      ContactVelocityConstraint* vc = &vcs[i];
      int indexA = vc->indexA, indexB = vc->indexB;
      float mA = vc->invMassA, iA = vc->invIA, mB = vc->invMassB, iB = vc->invIB;
      int pointCount = vc->pointCount;
      V2 vA = velocities[indexA].v; float wA = velocities[indexA].w;
      V2 vB = velocities[indexB].v; float wB = velocities[indexB].w;
      V2 normal = vc->normal;
      V2 tangent = cross(normal, 1.0f);
      float friction = vc->friction;
      for (int j = 0; j < pointCount; ++j) {
        VelocityConstraintPoint* vcp = vc->points + j;
        V2 dv = vB + cross(wB, vcp->rB) - vA - cross(wA, vcp->rA);
        float vt = dot(dv, tangent) - vc->tangentSpeed;
        float lambda = vcp->tangentMass * (-vt);
        float maxFriction = friction * vcp->normalImpulse;
        float newImpulse = clampT(vcp->tangentImpulse + lambda, -maxFriction, maxFriction);
        lambda = newImpulse - vcp->tangentImpulse;
        vcp->tangentImpulse = newImpulse;
        V2 P = lambda * tangent;
        vA -= mA * P;
        wA -= iA * cross(vcp->rA, P);
        vB += mB * P;
        wB += iB * cross(vcp->rB, P);
      }
      if (pointCount == 1) {
        for (int idx = 0; idx < pointCount; ++idx) {
          VelocityConstraintPoint* vcp = vc->points + idx;
          V2 dv = vB + cross(wB, vcp->rB) - vA - cross(wA, vcp->rA);
          float vn = dot(dv, normal);
          float lambda = -vcp->normalMass * (vn - vcp->velocityBias);
          float newImpulse = maxT(vcp->normalImpulse + lambda, 0.0f);
          lambda = newImpulse - vcp->normalImpulse;
          vcp->normalImpulse = newImpulse;
          V2 P = lambda * normal;
          vA -= mA * P;
          wA -= iA * cross(vcp->rA, P);
          vB += mB * P;
          wB += iB * cross(vcp->rB, P);
        }
      } else {
        VelocityConstraintPoint* cp1 = vc->points + 0;
        VelocityConstraintPoint* cp2 = vc->points + 1;
        V2 a(cp1->normalImpulse, cp2->normalImpulse);
        V2 dv1 = vB + cross(wB, cp1->rB) - vA - cross(wA, cp1->rA);
        V2 dv2 = vB + cross(wB, cp2->rB) - vA - cross(wA, cp2->rA);
        float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
        V2 b;
        b.x = vn1 - cp1->velocityBias;
        b.y = vn2 - cp2->velocityBias;
        b -= mul(vc->K, a);
        auto apply = [&](V2 x) {
          V2 d = x - a;
          V2 P1 = d.x * normal, P2 = d.y * normal;
          vA -= mA * (P1 + P2);
          wA -= iA * (cross(cp1->rA, P1) + cross(cp2->rA, P2));
          vB += mB * (P1 + P2);
          wB += iB * (cross(cp1->rB, P1) + cross(cp2->rB, P2));
          cp1->normalImpulse = x.x;
          cp2->normalImpulse = x.y;
        };
        for (;;) {
          V2 x = -mul(vc->normalMass, b);
          if (x.x >= 0.0f && x.y >= 0.0f) { apply(x); break; }
          x.x = -cp1->normalMass * b.x;
          x.y = 0.0f;
          vn1 = 0.0f;
          vn2 = vc->K.ex.y * x.x + b.y;
          if (x.x >= 0.0f && vn2 >= 0.0f) { apply(x); break; }
          x.x = 0.0f;
          x.y = -cp2->normalMass * b.y;
          vn1 = vc->K.ey.x * x.y + b.x;
          vn2 = 0.0f;
          if (x.y >= 0.0f && vn1 >= 0.0f) { apply(x); break; }
          x.x = 0.0f;
          x.y = 0.0f;
          vn1 = b.x;
          vn2 = b.y;
          if (vn1 >= 0.0f && vn2 >= 0.0f) { apply(x); break; }
          break;
        }
      }
      velocities[indexA].v = vA; velocities[indexA].w = wA;
      velocities[indexB].v = vB; velocities[indexB].w = wB;
    }
  }

  // b2contactsolver.d:774-787
  void storeImpulses() {
    for (int i = 0; i < count; ++i) {
      ContactVelocityConstraint* vc = &vcs[i];
      Manifold* manifold = &contacts[vc->contactIndex]->manifold;
      for (int j = 0; j < vc->pointCount; ++j) {
        manifold->points[j].normalImpulse = vc->points[j].normalImpulse;
        manifold->points[j].tangentImpulse = vc->points[j].tangentImpulse;
      }
    }
  }

  // b2contactsolver.d:73-149 (toi=false) and :152-242 (toi=true)
  bool solvePositionImpl(bool toi, int toiIndexA, int toiIndexB) {
    const float minSeparation = solvePositionRange(toi, toiIndexA, toiIndexB, 0, count, 0.0f);
    return minSeparation >= (toi ? -1.5f : -3.0f) * kLinearSlop;
  }
  // constraints [i0, i1) of SolvePositionConstraints (:73-149) / SolveTOIPositionConstraints (:152-242); carries the running minimum
  float solvePositionRange(bool toi, int toiIndexA, int toiIndexB, int i0, int i1, float minSeparation) {
    for (int i = i0; i < i1; ++i) {
      ContactPositionConstraint* pc = &pcs[i];
      int indexA = pc->indexA, indexB = pc->indexB;
      V2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
      int pointCount = pc->pointCount;
      float mA, iA, mB, iB;
      if (toi) {
        mA = 0.0f; iA = 0.0f;
        if (indexA == toiIndexA || indexA == toiIndexB) { mA = pc->invMassA; iA = pc->invIA; }
        mB = 0.0f; iB = 0.0f;
        if (indexB == toiIndexA || indexB == toiIndexB) { mB = pc->invMassB; iB = pc->invIB; }
      } else { mA = pc->invMassA; iA = pc->invIA; mB = pc->invMassB; iB = pc->invIB; }
      V2 cA = positions[indexA].c; float aA = positions[indexA].a;
      V2 cB = positions[indexB].c; float aB = positions[indexB].a;
      for (int j = 0; j < pointCount; ++j) {
        Xf xfA, xfB;
        xfA.q.set(aA); xfB.q.set(aB);
        xfA.p = cA - mul(xfA.q, localCenterA);
        xfB.p = cB - mul(xfB.q, localCenterB);
        PositionSolverManifold psm;
        psm.initialize(pc, xfA, xfB, j);
        V2 normal = psm.normal;
        V2 point = psm.point;
        float separation = psm.separation;
        V2 rA = point - cA, rB = point - cB;
        minSeparation = minT(minSeparation, separation);
        float C = clampT((toi ? kToiBaumgarte : kBaumgarte) * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
        float rnA = cross(rA, normal), rnB = cross(rB, normal);
        float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        float impulse = K > 0.0f ? -C / K : 0.0f;
        V2 P = impulse * normal;
        cA -= mA * P;
        aA -= iA * cross(rA, P);
        cB += mB * P;
        aB += iB * cross(rB, P);
      }
      positions[indexA].c = cA; positions[indexA].a = aA;
      positions[indexB].c = cB; positions[indexB].a = aB;
    }
    return minSeparation;
  }
};

// b2island.d
struct Island {
  std::vector<Body*> bodies; std::vector<Contact*> contacts; std::vector<Joint*> joints;
  std::vector<Position> positions; std::vector<Velocity> velocities;
  size_t bodyCapacity, contactCapacity;
  Island(size_t bc, size_t cc) : bodyCapacity(bc), contactCapacity(cc) {}
  void clear() { bodies.clear(); contacts.clear(); joints.clear(); }
  void add(Body* b) { b->islandIndex = (int)bodies.size(); bodies.push_back(b); }
  void add(Contact* c) { contacts.push_back(c); }
  void add(Joint* j) { joints.push_back(j); }
  // b2island.d:438-462: hand every contact of the island its b2ContactImpulse (here: append to the world's call log)
  std::vector<World::PostSolveRec>* postSolveLog = nullptr; int reportPhase = 1;
  void report(const ContactVelocityConstraint* constraints) {
    if (postSolveLog == nullptr) return;
    for (size_t i = 0; i < contacts.size(); ++i) {
      const Contact* c = contacts[i];
      const ContactVelocityConstraint* vc = constraints + i;
      World::PostSolveRec r{};
      r.phase = reportPhase; r.fixtureA = c->fixtureA->id; r.childA = c->indexA; r.fixtureB = c->fixtureB->id; r.childB = c->indexB;
      r.count = vc->pointCount;
      for (int j = 0; j < vc->pointCount; ++j) { r.normalImpulses[j] = vc->points[j].normalImpulse; r.tangentImpulses[j] = vc->points[j].tangentImpulse; }
      postSolveLog->push_back(r);
    }
  }

  // b2island.d:75-280
  // seq != nullptr: test hook (World::orderOverride) -- one merged Gauss-Seidel order over joints (entry ~index) and contacts
  // (entry index), walked forwards by the velocity passes and, with reversePosition, backwards by the position passes, instead
  // of "all joints, then all contacts" (:153-161) / "all contacts, then all joints" (:206-216)
  void solve(Profile* profile, const TimeStep& step, V2 gravity, bool allowSleep, const std::vector<int>* seq = nullptr, int positionMode = 0) {
    double t0 = nowMs();
    float h = step.dt;
    int bodyCount = (int)bodies.size();
    positions.resize(bodyCount); velocities.resize(bodyCount);
    for (int i = 0; i < bodyCount; ++i) {
      Body* b = bodies[i];
      V2 c = b->sweep.c; float a = b->sweep.a;
      V2 v = b->linearVelocity; float w = b->angularVelocity;
      b->sweep.c0 = b->sweep.c;
      b->sweep.a0 = b->sweep.a;
      if (b->type == kDynamic) {
        v += h * (b->gravityScale * gravity + b->invMass * b->force);
        w += h * b->invI * b->torque;
        v *= 1.0f / (1.0f + h * b->linearDamping);
        w *= 1.0f / (1.0f + h * b->angularDamping);
      }
      positions[i].c = c; positions[i].a = a;
      velocities[i].v = v; velocities[i].w = w;
    }
    SolverData solverData; solverData.step = step; solverData.positions = positions.data(); solverData.velocities = velocities.data();
    ContactSolver contactSolver(step, contacts.data(), (int)contacts.size(), positions.data(), velocities.data());
    contactSolver.initializeVelocityConstraints();
    if (step.warmStarting) contactSolver.warmStart();
    if (seq) { for (int e : *seq) if (e < 0) joints[~e]->initVelocityConstraints(solverData); }
    else for (Joint* j : joints) j->initVelocityConstraints(solverData);
    double t1 = nowMs(); profile->solveInit = (float)(t1 - t0);
    for (int i = 0; i < step.velocityIterations; ++i) {
      if (seq) {
        for (int e : *seq) { if (e < 0) joints[~e]->solveVelocityConstraints(solverData); else contactSolver.solveVelocityRange(e, e + 1); }
        continue;
      }
      for (Joint* j : joints) j->solveVelocityConstraints(solverData);
      contactSolver.solveVelocityConstraints();
    }
    contactSolver.storeImpulses();
    double t2 = nowMs(); profile->solveVelocity = (float)(t2 - t1);
    for (int i = 0; i < bodyCount; ++i) {
      V2 c = positions[i].c; float a = positions[i].a;
      V2 v = velocities[i].v; float w = velocities[i].w;
      V2 translation = h * v;
      if (dot(translation, translation) > kMaxTranslationSquared) {
        float ratio = kMaxTranslation / translation.len();
        v *= ratio;
      }
      float rotation = h * w;
      if (rotation * rotation > kMaxRotationSquared) {
        float ratio = kMaxRotation / absT(rotation);
        w *= ratio;
      }
      c += h * v;
      a += h * w;
      positions[i].c = c; positions[i].a = a;
      velocities[i].v = v; velocities[i].w = w;
    }
    bool positionSolved = false;
    for (int i = 0; i < step.positionIterations; ++i) {
      bool contactsOkay, jointsOkay = true;
      if (seq) {
        float minSeparation = 0.0f;
        const int m = (int)seq->size();
        // positionMode 0: the sequence forwards; 1: backwards; 2: its contacts, then its joints (the reference's own split, :206-216)
        for (int pass = 0; pass < (positionMode == 2 ? 2 : 1); ++pass)
          for (int k = 0; k < m; ++k) {
            const int e = (*seq)[positionMode == 1 ? m - 1 - k : k];
            if (positionMode == 2 && (e < 0) != (pass == 1)) continue;
            if (e < 0) { bool jointOkay = joints[~e]->solvePositionConstraints(solverData); jointsOkay = jointsOkay && jointOkay; }
            else minSeparation = contactSolver.solvePositionRange(false, 0, 0, e, e + 1, minSeparation);
          }
        contactsOkay = minSeparation >= -3.0f * kLinearSlop;
      } else {
        contactsOkay = contactSolver.solvePositionImpl(false, 0, 0);
        for (Joint* j : joints) { bool jointOkay = j->solvePositionConstraints(solverData); jointsOkay = jointsOkay && jointOkay; }
      }
      if (contactsOkay && jointsOkay) { positionSolved = true; break; }
    }
    for (int i = 0; i < bodyCount; ++i) {
      Body* b = bodies[i];
      b->sweep.c = positions[i].c; b->sweep.a = positions[i].a;
      b->linearVelocity = velocities[i].v; b->angularVelocity = velocities[i].w;
      b->synchronizeTransform();
    }
    profile->solvePosition = (float)(nowMs() - t2);
    report(contactSolver.vcs.data());            // b2island.d:239
    if (allowSleep) {
      float minSleepTime = kMaxFloat;
      const float linTolSqr = kLinearSleepTolerance * kLinearSleepTolerance;
      const float angTolSqr = kAngularSleepTolerance * kAngularSleepTolerance;
      for (int i = 0; i < bodyCount; ++i) {
        Body* b = bodies[i];
        if (b->type == kStatic) continue;
        if ((b->flags & bAutoSleep) == 0 || b->angularVelocity * b->angularVelocity > angTolSqr ||
            dot(b->linearVelocity, b->linearVelocity) > linTolSqr) {
          b->sleepTime = 0.0f;
          minSleepTime = 0.0f;
        } else {
          b->sleepTime += h;
          minSleepTime = minT(minSleepTime, b->sleepTime);
        }
      }
      if (minSleepTime >= kTimeToSleep && positionSolved) {
        for (int i = 0; i < bodyCount; ++i) bodies[i]->setAwake(false);
      }
    }
  }

  // b2island.d:282-416
  void solveTOI(const TimeStep& subStep, int toiIndexA, int toiIndexB) {
    int bodyCount = (int)bodies.size();
    positions.resize(bodyCount); velocities.resize(bodyCount);
    for (int i = 0; i < bodyCount; ++i) {
      Body* b = bodies[i];
      positions[i].c = b->sweep.c; positions[i].a = b->sweep.a;
      velocities[i].v = b->linearVelocity; velocities[i].w = b->angularVelocity;
    }
    ContactSolver contactSolver(subStep, contacts.data(), (int)contacts.size(), positions.data(), velocities.data());
    for (int i = 0; i < subStep.positionIterations; ++i) {
      bool contactsOkay = contactSolver.solvePositionImpl(true, toiIndexA, toiIndexB);
      if (contactsOkay) break;
    }
    bodies[toiIndexA]->sweep.c0 = positions[toiIndexA].c;
    bodies[toiIndexA]->sweep.a0 = positions[toiIndexA].a;
    bodies[toiIndexB]->sweep.c0 = positions[toiIndexB].c;
    bodies[toiIndexB]->sweep.a0 = positions[toiIndexB].a;
    contactSolver.initializeVelocityConstraints();
    for (int i = 0; i < subStep.velocityIterations; ++i) contactSolver.solveVelocityConstraints();
    float h = subStep.dt;
    for (int i = 0; i < bodyCount; ++i) {
      V2 c = positions[i].c; float a = positions[i].a;
      V2 v = velocities[i].v; float w = velocities[i].w;
      V2 translation = h * v;
      if (dot(translation, translation) > kMaxTranslationSquared) {
        float ratio = kMaxTranslation / translation.len();
        v *= ratio;
      }
      float rotation = h * w;
      if (rotation * rotation > kMaxRotationSquared) {
        float ratio = kMaxRotation / absT(rotation);
        w *= ratio;
      }
      c += h * v;
      a += h * w;
      positions[i].c = c; positions[i].a = a;
      velocities[i].v = v; velocities[i].w = w;
      Body* b = bodies[i];
      b->sweep.c = c; b->sweep.a = a;
      b->linearVelocity = v; b->angularVelocity = w;
      b->synchronizeTransform();
    }
    report(contactSolver.vcs.data());            // b2island.d:414
  }
};
}  // namespace

// ------------------------------------------------------------------ World::solve (b2world.d:930-1124)
void World::solve(const TimeStep& step) {
  profile.solveInit = 0.0f; profile.solveVelocity = 0.0f; profile.solvePosition = 0.0f;
  Island island(bodyCount, contactCount);
  if (recordPostSolve) { island.postSolveLog = &postSolveLog; island.reportPhase = 1; }
  for (Body* b = bodyList; b; b = b->next) b->flags &= ~bIsland;
  for (Contact* c = contactList; c; c = c->next) c->flags &= ~cIsland;
  for (Joint* j = jointList; j; j = j->next) j->islandFlag = false;
  lastIslandCount = 0;
  lastSolveOrder.clear(); lastJointOrder.clear();
  std::vector<Body*> stack(bodyCount);
  for (Body* seed = bodyList; seed; seed = seed->next) {
    if (seed->flags & bIsland) continue;
    if (seed->isAwake() == false || seed->isActive() == false) continue;
    if (seed->type == kStatic) continue;
    island.clear();
    int stackCount = 0;
    stack[stackCount++] = seed;
    seed->flags |= bIsland;
    while (stackCount > 0) {
      Body* b = stack[--stackCount];
      island.add(b);
      b->setAwake(true);
      if (b->type == kStatic) continue;
      for (ContactEdge* ce = b->contactList; ce; ce = ce->next) {
        Contact* contact = ce->contact;
        if (contact->flags & cIsland) continue;
        if (contact->isEnabled() == false || contact->isTouching() == false) continue;
        if (contact->fixtureA->isSensor || contact->fixtureB->isSensor) continue;
        island.add(contact);
        contact->flags |= cIsland;
        Body* other = ce->other;
        if (other->flags & bIsland) continue;
        stack[stackCount++] = other;
        other->flags |= bIsland;
      }
      for (JointEdge* je = b->jointList; je; je = je->next) {
        if (je->joint->islandFlag == true) continue;
        Body* other = je->other;
        if (other->isActive() == false) continue;
        island.add(je->joint);
        je->joint->islandFlag = true;
        if (other->flags & bIsland) continue;
        stack[stackCount++] = other;
        other->flags |= bIsland;
      }
    }
    std::vector<int> seq;
    if (orderOverride) {   // test hook (orc_world.h): caller-supplied Gauss-Seidel order instead of DFS order, joints and contacts merged
      // the arrays themselves go into rank order too: the contacts' warm start (:138-141) walks the contact array
      std::stable_sort(island.contacts.begin(), island.contacts.end(), [](const Contact* a, const Contact* b) { return a->orderRank < b->orderRank; });
      std::stable_sort(island.joints.begin(), island.joints.end(), [](const Joint* a, const Joint* b) { return a->orderRank < b->orderRank; });
      const int nc = (int)island.contacts.size(), nj = (int)island.joints.size();
      seq.resize((size_t)nc + nj);
      for (int k = 0; k < nj; ++k) seq[k] = ~k;
      for (int k = 0; k < nc; ++k) seq[nj + k] = k;
      auto rank = [&](int e) { return e < 0 ? island.joints[~e]->orderRank : island.contacts[e]->orderRank; };
      std::stable_sort(seq.begin(), seq.end(), [&](int a, int b) { return rank(a) < rank(b); });
    }
    Profile p;
    island.solve(&p, step, gravity, allowSleep, orderOverride ? &seq : nullptr, orderPositionMode);
    ++lastIslandCount;
    lastSolveOrder.insert(lastSolveOrder.end(), island.contacts.begin(), island.contacts.end());
    for (Joint* j : island.joints) lastJointOrder.push_back(j->id);
    profile.solveInit += p.solveInit; profile.solveVelocity += p.solveVelocity; profile.solvePosition += p.solvePosition;
    for (Body* b : island.bodies) if (b->type == kStatic) b->flags &= ~bIsland;
  }
  orderOverride = false;
  double t0 = nowMs();
  for (Body* b = bodyList; b; b = b->next) {
    if ((b->flags & bIsland) == 0) continue;
    if (b->type == kStatic) continue;
    b->synchronizeFixtures();
  }
  findNewContacts();
  profile.broadphase = (float)(nowMs() - t0);
}

// ------------------------------------------------------------------ World::solveTOI (b2world.d:1127-1452)
void World::solveTOI(const TimeStep& step) {
  Island island(2 * kMaxTOIContacts, kMaxTOIContacts);
  if (recordPostSolve) { island.postSolveLog = &postSolveLog; island.reportPhase = 2; }
  if (stepComplete) {
    for (Body* b = bodyList; b; b = b->next) { b->flags &= ~bIsland; b->sweep.alpha0 = 0.0f; }
    for (Contact* c = contactList; c; c = c->next) { c->flags &= ~(cToi | cIsland); c->toiCount = 0; c->toi = 1.0f; }
  }
  for (;;) {
    Contact* minContact = nullptr;
    float minAlpha = 1.0f;
    for (Contact* c = contactList; c; c = c->next) {
      if (c->isEnabled() == false) continue;
      if (c->toiCount > kMaxSubSteps) continue;
      float alpha = 1.0f;
      if (c->flags & cToi) {
        alpha = c->toi;
      } else {
        Fixture* fA = c->fixtureA; Fixture* fB = c->fixtureB;
        if (fA->isSensor || fB->isSensor) continue;
        Body* bA = fA->body; Body* bB = fB->body;
        int typeA = bA->type, typeB = bB->type;
        bool activeA = bA->isAwake() && typeA != kStatic;
        bool activeB = bB->isAwake() && typeB != kStatic;
        if (activeA == false && activeB == false) continue;
        bool collideA = bA->isBullet() || typeA != kDynamic;
        bool collideB = bB->isBullet() || typeB != kDynamic;
        if (collideA == false && collideB == false) continue;
        float alpha0 = bA->sweep.alpha0;
        if (bA->sweep.alpha0 < bB->sweep.alpha0) { alpha0 = bB->sweep.alpha0; bA->sweep.advance(alpha0); }
        else if (bB->sweep.alpha0 < bA->sweep.alpha0) { alpha0 = bA->sweep.alpha0; bB->sweep.advance(alpha0); }
        int indexA = c->indexA, indexB = c->indexB;
        TOIInput input;
        input.proxyA.set(fA->shape, indexA);
        input.proxyB.set(fB->shape, indexB);
        input.sweepA = bA->sweep;
        input.sweepB = bB->sweep;
        input.tMax = 1.0f;
        TOIOutput output;
        timeOfImpact(&output, &input);
        float beta = output.t;
        if (output.state == kToiTouching) alpha = minT(alpha0 + (1.0f - alpha0) * beta, 1.0f);
        else alpha = 1.0f;
        c->toi = alpha;
        c->flags |= cToi;
      }
      if (alpha < minAlpha) { minContact = c; minAlpha = alpha; }
    }
    if (minContact == nullptr || 1.0f - 10.0f * kEpsilon < minAlpha) { stepComplete = true; break; }
    ++toiEvents;
    Fixture* fA = minContact->fixtureA; Fixture* fB = minContact->fixtureB;
    Body* bA = fA->body; Body* bB = fB->body;
    Sweep backup1 = bA->sweep, backup2 = bB->sweep;
    bA->advance(minAlpha);
    bB->advance(minAlpha);
    minContact->update(this);
    minContact->flags &= ~cToi;
    ++minContact->toiCount;
    if (minContact->isEnabled() == false || minContact->isTouching() == false) {
      minContact->flags &= ~cEnabled;
      bA->sweep = backup1; bB->sweep = backup2;
      bA->synchronizeTransform(); bB->synchronizeTransform();
      continue;
    }
    bA->setAwake(true);
    bB->setAwake(true);
    island.clear();
    island.add(bA); island.add(bB); island.add(minContact);
    bA->flags |= bIsland; bB->flags |= bIsland; minContact->flags |= cIsland;
    Body* pair[2] = {bA, bB};
    for (int i = 0; i < 2; ++i) {
      Body* body = pair[i];
      if (body->type == kDynamic) {
        for (ContactEdge* ce = body->contactList; ce; ce = ce->next) {
          if (island.bodies.size() == island.bodyCapacity) break;
          if (island.contacts.size() == island.contactCapacity) break;
          Contact* contact = ce->contact;
          if (contact->flags & cIsland) continue;
          Body* other = ce->other;
          if (other->type == kDynamic && body->isBullet() == false && other->isBullet() == false) continue;
          if (contact->fixtureA->isSensor || contact->fixtureB->isSensor) continue;
          Sweep backup = other->sweep;
          if ((other->flags & bIsland) == 0) other->advance(minAlpha);
          contact->update(this);
          if (contact->isEnabled() == false) { other->sweep = backup; other->synchronizeTransform(); continue; }
          if (contact->isTouching() == false) { other->sweep = backup; other->synchronizeTransform(); continue; }
          contact->flags |= cIsland;
          island.add(contact);
          if (other->flags & bIsland) continue;
          other->flags |= bIsland;
          if (other->type != kStatic) other->setAwake(true);
          island.add(other);
        }
      }
    }
    TimeStep subStep;
    subStep.dt = (1.0f - minAlpha) * step.dt;
    subStep.inv_dt = 1.0f / subStep.dt;
    subStep.dtRatio = 1.0f;
    subStep.positionIterations = 20;
    subStep.velocityIterations = step.velocityIterations;
    subStep.warmStarting = false;
    island.solveTOI(subStep, bA->islandIndex, bB->islandIndex);
    for (Body* body : island.bodies) {
      body->flags &= ~bIsland;
      if (body->type != kDynamic) continue;
      body->synchronizeFixtures();
      for (ContactEdge* ce = body->contactList; ce; ce = ce->next) ce->contact->flags &= ~(cToi | cIsland);
    }
    findNewContacts();
    if (subStepping) { stepComplete = false; break; }
  }
}

// ------------------------------------------------------------------ World::step (b2world.d:367-434)
// b2World.Step (b2world.d:367-434) cut after Collide, which is where PreSolve edits land (b2contact.d:348-355); halves = 3 is the
// plain step
void World::step(float dt, int velocityIterations, int positionIterations) { stepHalves(dt, velocityIterations, positionIterations, 3); }
void World::stepHalves(float dt, int velocityIterations, int positionIterations, int halves) {
  double t0 = nowMs();
  if (halves & 1) { if (newFixture) { findNewContacts(); newFixture = false; } }
  locked = true;
  TimeStep step;
  step.dt = dt;
  step.velocityIterations = velocityIterations;
  step.positionIterations = positionIterations;
  step.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
  step.dtRatio = inv_dt0 * dt;
  step.warmStarting = warmStarting;
  evPhase = 1;
  if (halves & 1) { double t = nowMs(); collide(); profile.collide = (float)(nowMs() - t); }
  if (!(halves & 2)) { locked = false; return; }
  postSolveLog.clear();
  if (stepComplete && step.dt > 0.0f) { double t = nowMs(); solve(step); profile.solve = (float)(nowMs() - t); }
  evPhase = 2;
  if (continuousPhysics && step.dt > 0.0f) { double t = nowMs(); solveTOI(step); profile.solveTOI = (float)(nowMs() - t); }
  evPhase = 3; ++stepCount;
  if (step.dt > 0.0f) inv_dt0 = step.inv_dt;
  if (clearForcesFlag) clearForces();
  locked = false;
  profile.step = (float)(nowMs() - t0);
}

// b2world.d:758-780
void World::shiftOrigin(V2 newOrigin) {
  if (locked) return;
  for (Body* b = bodyList; b; b = b->next) {
    b->xf.p -= newOrigin;
    b->sweep.c0 -= newOrigin;
    b->sweep.c -= newOrigin;
  }
  for (Joint* j = jointList; j; j = j->next) {
    // b2Joint.ShiftOrigin is a no-op except b2mousejoint.d:174-177 and b2pulleyjoint.d:227-231
    if (j->type == jMouse) static_cast<MouseJoint*>(j)->targetA -= newOrigin;
    else if (j->type == jPulley) { PulleyJoint* pj = static_cast<PulleyJoint*>(j); pj->groundAnchorA -= newOrigin; pj->groundAnchorB -= newOrigin; }
  }
  broadPhase.tree().shiftOrigin(newOrigin);          // b2broadphase.d:237-240
}

void World::clearForces() {
  for (Body* b = bodyList; b; b = b->next) { b->force = V2(0, 0); b->torque = 0.0f; }
}

// ------------------------------------------------------------------ RevoluteJoint (b2revolutejoint.d)
void RevoluteJoint::initVelocityConstraints(const SolverData& data) {
  indexA = bodyA->islandIndex; indexB = bodyB->islandIndex;
  localCenterA = bodyA->sweep.localCenter; localCenterB = bodyB->sweep.localCenter;
  invMassA = bodyA->invMass; invMassB = bodyB->invMass; invIA = bodyA->invI; invIB = bodyB->invI;
  float aA = data.positions[indexA].a; V2 vA = data.velocities[indexA].v; float wA = data.velocities[indexA].w;
  float aB = data.positions[indexB].a; V2 vB = data.velocities[indexB].v; float wB = data.velocities[indexB].w;
  Rot qA(aA), qB(aB);
  rA = mul(qA, localAnchorA - localCenterA);
  rB = mul(qB, localAnchorB - localCenterB);
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  bool fixedRotation = (iA + iB == 0.0f);
  mass.ex.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
  mass.ey.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
  mass.ez.x = -rA.y * iA - rB.y * iB;
  mass.ex.y = mass.ey.x;
  mass.ey.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
  mass.ez.y = rA.x * iA + rB.x * iB;
  mass.ex.z = mass.ez.x;
  mass.ey.z = mass.ez.y;
  mass.ez.z = iA + iB;
  motorMass = iA + iB;
  if (motorMass > 0.0f) motorMass = 1.0f / motorMass;
  if (enableMotor == false || fixedRotation) motorImpulse = 0.0f;
  if (enableLimit && fixedRotation == false) {
    float jointAngle = aB - aA - referenceAngle;
    if (absT(upperAngle - lowerAngle) < 2.0f * kAngularSlop) limitState = kEqualLimits;
    else if (jointAngle <= lowerAngle) { if (limitState != kAtLowerLimit) impulse.z = 0.0f; limitState = kAtLowerLimit; }
    else if (jointAngle >= upperAngle) { if (limitState != kAtUpperLimit) impulse.z = 0.0f; limitState = kAtUpperLimit; }
    else { limitState = kInactiveLimit; impulse.z = 0.0f; }
  } else {
    limitState = kInactiveLimit;
  }
  if (data.step.warmStarting) {
    impulse *= data.step.dtRatio;
    motorImpulse *= data.step.dtRatio;
    V2 P(impulse.x, impulse.y);
    vA -= mA * P;
    wA -= iA * (cross(rA, P) + motorImpulse + impulse.z);
    vB += mB * P;
    wB += iB * (cross(rB, P) + motorImpulse + impulse.z);
  } else {
    impulse = V3(0, 0, 0);
    motorImpulse = 0.0f;
  }
  data.velocities[indexA].v = vA; data.velocities[indexA].w = wA;
  data.velocities[indexB].v = vB; data.velocities[indexB].w = wB;
}

void RevoluteJoint::solveVelocityConstraints(const SolverData& data) {
  V2 vA = data.velocities[indexA].v; float wA = data.velocities[indexA].w;
  V2 vB = data.velocities[indexB].v; float wB = data.velocities[indexB].w;
  float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
  bool fixedRotation = (iA + iB == 0.0f);
  if (enableMotor && limitState != kEqualLimits && fixedRotation == false) {
    float Cdot = wB - wA - motorSpeed;
    float imp = -motorMass * Cdot;
    float oldImpulse = motorImpulse;
    float maxImpulse = data.step.dt * maxMotorTorque;
    motorImpulse = clampT(motorImpulse + imp, -maxImpulse, maxImpulse);
    imp = motorImpulse - oldImpulse;
    wA -= iA * imp;
    wB += iB * imp;
  }
  if (enableLimit && limitState != kInactiveLimit && fixedRotation == false) {
    V2 Cdot1 = vB + cross(wB, rB) - vA - cross(wA, rA);
    float Cdot2 = wB - wA;
    V3 Cdot(Cdot1.x, Cdot1.y, Cdot2);
    V3 imp = -mass.solve33(Cdot);
    if (limitState == kEqualLimits) {
      impulse += imp;
    } else if (limitState == kAtLowerLimit) {
      float newImpulse = impulse.z + imp.z;
      if (newImpulse < 0.0f) {
        V2 rhs = -Cdot1 + impulse.z * V2(mass.ez.x, mass.ez.y);
        V2 reduced = mass.solve22(rhs);
        imp.x = reduced.x; imp.y = reduced.y; imp.z = -impulse.z;
        impulse.x += reduced.x; impulse.y += reduced.y; impulse.z = 0.0f;
      } else {
        impulse += imp;
      }
    } else if (limitState == kAtUpperLimit) {
      float newImpulse = impulse.z + imp.z;
      if (newImpulse > 0.0f) {
        V2 rhs = -Cdot1 + impulse.z * V2(mass.ez.x, mass.ez.y);
        V2 reduced = mass.solve22(rhs);
        imp.x = reduced.x; imp.y = reduced.y; imp.z = -impulse.z;
        impulse.x += reduced.x; impulse.y += reduced.y; impulse.z = 0.0f;
      } else {
        impulse += imp;
      }
    }
    V2 P(imp.x, imp.y);
    vA -= mA * P;
    wA -= iA * (cross(rA, P) + imp.z);
    vB += mB * P;
    wB += iB * (cross(rB, P) + imp.z);
  } else {
    V2 Cdot = vB + cross(wB, rB) - vA - cross(wA, rA);
    V2 imp = mass.solve22(-Cdot);
    impulse.x += imp.x;
    impulse.y += imp.y;
    vA -= mA * imp;
    wA -= iA * cross(rA, imp);
    vB += mB * imp;
    wB += iB * cross(rB, imp);
  }
  data.velocities[indexA].v = vA; data.velocities[indexA].w = wA;
  data.velocities[indexB].v = vB; data.velocities[indexB].w = wB;
}

bool RevoluteJoint::solvePositionConstraints(const SolverData& data) {
  V2 cA = data.positions[indexA].c; float aA = data.positions[indexA].a;
  V2 cB = data.positions[indexB].c; float aB = data.positions[indexB].a;
  Rot qA(aA), qB(aB);
  float angularError = 0.0f, positionError = 0.0f;
  bool fixedRotation = (invIA + invIB == 0.0f);
  if (enableLimit && limitState != kInactiveLimit && fixedRotation == false) {
    float angle = aB - aA - referenceAngle;
    float limitImpulse = 0.0f;
    if (limitState == kEqualLimits) {
      float C = clampT(angle - lowerAngle, -kMaxAngularCorrection, kMaxAngularCorrection);
      limitImpulse = -motorMass * C;
      angularError = absT(C);
    } else if (limitState == kAtLowerLimit) {
      float C = angle - lowerAngle;
      angularError = -C;
      C = clampT(C + kAngularSlop, -kMaxAngularCorrection, 0.0f);
      limitImpulse = -motorMass * C;
    } else if (limitState == kAtUpperLimit) {
      float C = angle - upperAngle;
      angularError = C;
      C = clampT(C - kAngularSlop, 0.0f, kMaxAngularCorrection);
      limitImpulse = -motorMass * C;
    }
    aA -= invIA * limitImpulse;
    aB += invIB * limitImpulse;
  }
  {
    qA.set(aA); qB.set(aB);
    V2 rA_ = mul(qA, localAnchorA - localCenterA);
    V2 rB_ = mul(qB, localAnchorB - localCenterB);
    V2 C = cB + rB_ - cA - rA_;
    positionError = C.len();
    float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
    M22 K;
    K.ex.x = mA + mB + iA * rA_.y * rA_.y + iB * rB_.y * rB_.y;
    K.ex.y = -iA * rA_.x * rA_.y - iB * rB_.x * rB_.y;
    K.ey.x = K.ex.y;
    K.ey.y = mA + mB + iA * rA_.x * rA_.x + iB * rB_.x * rB_.x;
    V2 imp = -K.solve(C);
    cA -= mA * imp;
    aA -= iA * cross(rA_, imp);
    cB += mB * imp;
    aB += iB * cross(rB_, imp);
  }
  data.positions[indexA].c = cA; data.positions[indexA].a = aA;
  data.positions[indexB].c = cB; data.positions[indexB].a = aB;
  return positionError <= kLinearSlop && angularError <= kAngularSlop;
}

// ------------------------------------------------------------------ DistanceJoint (b2distancejoint.d)
void DistanceJoint::initVelocityConstraints(const SolverData& data) {
  indexA = bodyA->islandIndex; indexB = bodyB->islandIndex;
  localCenterA = bodyA->sweep.localCenter; localCenterB = bodyB->sweep.localCenter;
  invMassA = bodyA->invMass; invMassB = bodyB->invMass; invIA = bodyA->invI; invIB = bodyB->invI;
  V2 cA = data.positions[indexA].c; float aA = data.positions[indexA].a;
  V2 vA = data.velocities[indexA].v; float wA = data.velocities[indexA].w;
  V2 cB = data.positions[indexB].c; float aB = data.positions[indexB].a;
  V2 vB = data.velocities[indexB].v; float wB = data.velocities[indexB].w;
  Rot qA(aA), qB(aB);
  rA = mul(qA, localAnchorA - localCenterA);
  rB = mul(qB, localAnchorB - localCenterB);
  u = cB + rB - cA - rA;
  float len = u.len();
  if (len > kLinearSlop) u *= 1.0f / len; else u = V2(0.0f, 0.0f);
  float crAu = cross(rA, u), crBu = cross(rB, u);
  float invMass = invMassA + invIA * crAu * crAu + invMassB + invIB * crBu * crBu;
  mass = invMass != 0.0f ? 1.0f / invMass : 0.0f;
  if (frequencyHz > 0.0f) {
    float C = len - length;
    float omega = 2.0f * kPi * frequencyHz;
    float d = 2.0f * mass * dampingRatio * omega;
    float k = mass * omega * omega;
    float h = data.step.dt;
    gamma = h * (d + h * k);
    gamma = gamma != 0.0f ? 1.0f / gamma : 0.0f;
    bias = C * h * k * gamma;
    invMass += gamma;
    mass = invMass != 0.0f ? 1.0f / invMass : 0.0f;
  } else {
    gamma = 0.0f;
    bias = 0.0f;
  }
  if (data.step.warmStarting) {
    impulse *= data.step.dtRatio;
    V2 P = impulse * u;
    vA -= invMassA * P;
    wA -= invIA * cross(rA, P);
    vB += invMassB * P;
    wB += invIB * cross(rB, P);
  } else {
    impulse = 0.0f;
  }
  data.velocities[indexA].v = vA; data.velocities[indexA].w = wA;
  data.velocities[indexB].v = vB; data.velocities[indexB].w = wB;
}

void DistanceJoint::solveVelocityConstraints(const SolverData& data) {
  V2 vA = data.velocities[indexA].v; float wA = data.velocities[indexA].w;
  V2 vB = data.velocities[indexB].v; float wB = data.velocities[indexB].w;
  V2 vpA = vA + cross(wA, rA);
  V2 vpB = vB + cross(wB, rB);
  float Cdot = dot(u, vpB - vpA);
  float imp = -mass * (Cdot + bias + gamma * impulse);
  impulse += imp;
  V2 P = imp * u;
  vA -= invMassA * P;
  wA -= invIA * cross(rA, P);
  vB += invMassB * P;
  wB += invIB * cross(rB, P);
  data.velocities[indexA].v = vA; data.velocities[indexA].w = wA;
  data.velocities[indexB].v = vB; data.velocities[indexB].w = wB;
}

bool DistanceJoint::solvePositionConstraints(const SolverData& data) {
  if (frequencyHz > 0.0f) return true;
  V2 cA = data.positions[indexA].c; float aA = data.positions[indexA].a;
  V2 cB = data.positions[indexB].c; float aB = data.positions[indexB].a;
  Rot qA(aA), qB(aB);
  V2 rA_ = mul(qA, localAnchorA - localCenterA);
  V2 rB_ = mul(qB, localAnchorB - localCenterB);
  V2 u_ = cB + rB_ - cA - rA_;
  float len = u_.normalize();
  float C = len - length;
  C = clampT(C, -kMaxLinearCorrection, kMaxLinearCorrection);
  float imp = -mass * C;
  V2 P = imp * u_;
  cA -= invMassA * P;
  aA -= invIA * cross(rA_, P);
  cB += invMassB * P;
  aB += invIB * cross(rB_, P);
  data.positions[indexA].c = cA; data.positions[indexA].a = aA;
  data.positions[indexB].c = cB; data.positions[indexB].a = aB;
  return absT(C) < kLinearSlop;
}

}  // namespace orc
