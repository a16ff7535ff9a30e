// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// World / body / fixture / contact / joint object graph, contact manager, contact solver, islands and
// the TOI sub-stepper; restates
//   src/dbox/dynamics/b2world.d, b2body.d, b2fixture.d, b2contactmanager.d, b2island.d, b2timestep.d,
//   src/dbox/dynamics/b2worldcallbacks.d (default filter), dynamics/contacts/b2contact.d, b2contactsolver.d,
//   src/dbox/dynamics/joints/b2joint.d, b2revolutejoint.d, b2distancejoint.d
// It keeps the reference's intrusive push-front lists and DFS island order because those define the
// sequential Gauss-Seidel order the reference results depend on.
#pragma once
#include <memory>
#include <vector>
#include "orc_collide.h"
#include "orc_tree.h"

namespace orc {

struct Body; struct Fixture; struct Contact; struct Joint; struct World;

enum BodyType { kStatic = 0, kKinematic = 1, kDynamic = 2 };
enum BodyFlags { bIsland = 1, bAwake = 2, bAutoSleep = 4, bBullet = 8, bFixedRotation = 0x10, bActive = 0x20, bToi = 0x40 };
enum ContactFlags { cIsland = 1, cTouching = 2, cEnabled = 4, cFilter = 8, cBulletHit = 0x10, cToi = 0x20,
                    cPreSolveOff = 0x40 /* not in the reference: this step's PreSolve said SetEnabled(false) (orc_world_patch_contacts);
                                           the Update calls of the TOI loop (b2world.d:1295,1379) re-apply it, standing in for a
                                           listener that answers the same when asked again */ };
enum JointType { jUnknown, jRevolute, jPrismatic, jDistance, jPulley, jMouse, jGear, jWheel, jWeld, jFriction, jRope, jMotor };
enum LimitState { kInactiveLimit, kAtLowerLimit, kAtUpperLimit, kEqualLimits };

struct Filter { uint16_t categoryBits = 1, maskBits = 0xFFFF; int16_t groupIndex = 0; };
struct FixtureProxy { AABB aabb; Fixture* fixture = nullptr; int childIndex = 0; int proxyId = -1; };

struct ContactEdge { Body* other = nullptr; Contact* contact = nullptr; ContactEdge* prev = nullptr; ContactEdge* next = nullptr; };
struct JointEdge { Body* other = nullptr; Joint* joint = nullptr; JointEdge* prev = nullptr; JointEdge* next = nullptr; };

struct TimeStep { float dt = 0, inv_dt = 0, dtRatio = 0; int velocityIterations = 0, positionIterations = 0; bool warmStarting = false; };
struct Position { V2 c; float a = 0; };
struct Velocity { V2 v; float w = 0; };
struct SolverData { TimeStep step; Position* positions; Velocity* velocities; };

struct Fixture {
  int id = -1;
  float density = 0;
  Fixture* next = nullptr;
  Body* body = nullptr;
  Shape shape;
  float friction = 0.2f, restitution = 0;
  std::vector<FixtureProxy> proxies;  // sized to childCount at creation, never resized
  int proxyCount = 0;
  Filter filter;
  bool isSensor = false;
  uint64_t userData = 0;
};

struct Body {
  int id = -1;
  int type = kStatic;
  uint16_t flags = 0;
  int islandIndex = 0;
  Xf xf;
  Sweep sweep;
  V2 linearVelocity; float angularVelocity = 0;
  V2 force; float torque = 0;
  World* world = nullptr;
  Body* prev = nullptr; Body* next = nullptr;
  Fixture* fixtureList = nullptr; int fixtureCount = 0;
  JointEdge* jointList = nullptr;
  ContactEdge* contactList = nullptr;
  float mass = 0, invMass = 0, I = 0, invI = 0;
  float linearDamping = 0, angularDamping = 0, gravityScale = 1;
  float sleepTime = 0;
  uint64_t userData = 0;

  bool isAwake() const { return (flags & bAwake) == bAwake; }
  bool isActive() const { return (flags & bActive) == bActive; }
  bool isBullet() const { return (flags & bBullet) == bBullet; }
  void setAwake(bool flag);                // b2body.d:827-846
  void synchronizeTransform();             // b2body.d:1143-1147
  void synchronizeFixtures();              // b2body.d:1129-1141
  void advance(float alpha);               // b2body.d:1172-1180
  bool shouldCollide(const Body* other) const;  // b2body.d:1149-1170
  void resetMassData();                    // b2body.d:555-625
};

struct Contact {
  uint32_t flags = cEnabled;
  Contact* prev = nullptr; Contact* next = nullptr;
  ContactEdge nodeA, nodeB;
  Fixture* fixtureA = nullptr; Fixture* fixtureB = nullptr;
  int indexA = 0, indexB = 0;
  Manifold manifold;
  int toiCount = 0;
  float toi = 0;
  float friction = 0, restitution = 0, tangentSpeed = 0;
  int orderRank = 0x7fffffff;   // test hook (World::orderOverride): position in the caller-supplied Gauss-Seidel order
  bool isTouching() const { return (flags & cTouching) == cTouching; }
  bool isEnabled() const { return (flags & cEnabled) == cEnabled; }
  void evaluate(Manifold* m, const Xf& xfA, const Xf& xfB) const;  // the 7 b2*contact.d Evaluate overrides (line 56 each)
  void update(World* w);                                           // b2contact.d:270-356
};

struct Joint {
  int id = -1;
  int type = jUnknown;
  Joint* prev = nullptr; Joint* next = nullptr;
  JointEdge edgeA, edgeB;
  Body* bodyA = nullptr; Body* bodyB = nullptr;
  bool islandFlag = false, collideConnected = false;
  uint64_t userData = 0;
  int orderRank = 0x7fffffff;   // test hook, see Contact::orderRank
  virtual ~Joint() {}
  virtual void initVelocityConstraints(const SolverData& data) = 0;
  virtual void solveVelocityConstraints(const SolverData& data) = 0;
  virtual bool solvePositionConstraints(const SolverData& data) = 0;
};

struct RevoluteJoint : Joint {
  V2 localAnchorA, localAnchorB; V3 impulse; float motorImpulse = 0;
  bool enableMotor = false; float maxMotorTorque = 0, motorSpeed = 0;
  bool enableLimit = false; float referenceAngle = 0, lowerAngle = 0, upperAngle = 0;
  int indexA = 0, indexB = 0; V2 rA, rB, localCenterA, localCenterB;
  float invMassA = 0, invMassB = 0, invIA = 0, invIB = 0;
  M33 mass; float motorMass = 0; int limitState = kInactiveLimit;
  void initVelocityConstraints(const SolverData& data) override;
  void solveVelocityConstraints(const SolverData& data) override;
  bool solvePositionConstraints(const SolverData& data) override;
};

struct DistanceJoint : Joint {
  float frequencyHz = 0, dampingRatio = 0, bias = 0;
  V2 localAnchorA, localAnchorB; float gamma = 0, impulse = 0, length = 0;
  int indexA = 0, indexB = 0; V2 u, rA, rB, localCenterA, localCenterB;
  float invMassA = 0, invMassB = 0, invIA = 0, invIB = 0, mass = 0;
  void initVelocityConstraints(const SolverData& data) override;
  void solveVelocityConstraints(const SolverData& data) override;
  bool solvePositionConstraints(const SolverData& data) override;
};

// ---- second-wave joints, restated in orc_joints.cpp
#define ORC_JOINT_HOOKS \
  void initVelocityConstraints(const SolverData& data) override; \
  void solveVelocityConstraints(const SolverData& data) override; \
  bool solvePositionConstraints(const SolverData& data) override;
struct JointCommon : Joint {   // the solver-temp block every reference joint repeats
  int indexA = 0, indexB = 0; V2 rA, rB, localCenterA, localCenterB;
  float invMassA = 0, invMassB = 0, invIA = 0, invIB = 0;
  V2 localAnchorA, localAnchorB;
  void loadBodies();
};
struct RopeJoint : JointCommon {        // b2ropejoint.d
  float maxLength = 0, length = 0, impulse = 0, mass = 0; V2 u; int state = kInactiveLimit;
  ORC_JOINT_HOOKS
};
struct WeldJoint : JointCommon {        // b2weldjoint.d
  float frequencyHz = 0, dampingRatio = 0, bias = 0, referenceAngle = 0, gamma = 0; V3 impulse; M33 mass;
  ORC_JOINT_HOOKS
};
struct FrictionJoint : JointCommon {    // b2frictionjoint.d
  V2 linearImpulse; float angularImpulse = 0, maxForce = 0, maxTorque = 0; M22 linearMass; float angularMass = 0;
  ORC_JOINT_HOOKS
};
struct MotorJoint : JointCommon {       // b2motorjoint.d (no anchors: rA = -qA*lcA, rB = -qB*lcB)
  V2 linearOffset; float angularOffset = 0; V2 linearImpulse; float angularImpulse = 0, maxForce = 0, maxTorque = 0, correctionFactor = 0;
  V2 linearError; float angularError = 0; M22 linearMass; float angularMass = 0;
  ORC_JOINT_HOOKS
};
struct MouseJoint : JointCommon {       // b2mousejoint.d (drives bodyB only)
  V2 targetA; float frequencyHz = 0, dampingRatio = 0, beta = 0; V2 impulse; float maxForce = 0, gamma = 0; M22 mass; V2 C;
  ORC_JOINT_HOOKS
};
struct PrismaticJoint : JointCommon {   // b2prismaticjoint.d
  V2 localXAxisA, localYAxisA; float referenceAngle = 0; V3 impulse; float motorImpulse = 0, lowerTranslation = 0, upperTranslation = 0;
  float maxMotorForce = 0, motorSpeed = 0; bool enableLimit = false, enableMotor = false; int limitState = kInactiveLimit;
  V2 axis, perp; float s1 = 0, s2 = 0, a1 = 0, a2 = 0; M33 K; float motorMass = 0;
  ORC_JOINT_HOOKS
};
struct WheelJoint : JointCommon {       // b2wheeljoint.d
  float frequencyHz = 0, dampingRatio = 0; V2 localXAxisA, localYAxisA; float impulse = 0, motorImpulse = 0, springImpulse = 0;
  float maxMotorTorque = 0, motorSpeed = 0; bool enableMotor = false;
  V2 ax, ay; float sAx = 0, sBx = 0, sAy = 0, sBy = 0, mass = 0, motorMass = 0, springMass = 0, bias = 0, gamma = 0;
  ORC_JOINT_HOOKS
};
struct PulleyJoint : JointCommon {      // b2pulleyjoint.d
  V2 groundAnchorA, groundAnchorB; float lengthA = 0, lengthB = 0, constant = 0, ratio = 1, impulse = 0; V2 uA, uB; float mass = 0;
  ORC_JOINT_HOOKS
};

struct GearJoint : Joint {             // b2gearjoint.d (four bodies: A, B driven through joint1 / joint2 against C, D)
  int typeA = 0, typeB = 0; Body* bodyC = nullptr; Body* bodyD = nullptr;
  V2 localAnchorA, localAnchorB, localAnchorC, localAnchorD, localAxisC, localAxisD;
  float referenceAngleA = 0, referenceAngleB = 0, constant = 0, ratio = 1, impulse = 0;
  int indexA = 0, indexB = 0, indexC = 0, indexD = 0; V2 lcA, lcB, lcC, lcD;
  float mA = 0, mB = 0, mC = 0, mD = 0, iA = 0, iB = 0, iC = 0, iD = 0;
  V2 JvAC, JvBD; float JwA = 0, JwB = 0, JwC = 0, JwD = 0, mass = 0;
  ORC_JOINT_HOOKS
};

struct Profile { float step = 0, collide = 0, solve = 0, solveInit = 0, solveVelocity = 0, solvePosition = 0, broadphase = 0, solveTOI = 0; };

struct BodyDef {
  int type = kStatic; V2 position; float angle = 0; V2 linearVelocity; float angularVelocity = 0;
  float linearDamping = 0, angularDamping = 0;
  bool allowSleep = true, awake = true, fixedRotation = false, bullet = false, active = true;
  float gravityScale = 1.0f; uint64_t userData = 0;
};
struct FixtureDef { const Shape* shape = nullptr; float friction = 0.2f, restitution = 0, density = 0; bool isSensor = false; Filter filter; uint64_t userData = 0; };

struct World {
  explicit World(V2 gravity);
  ~World();
  Body* createBody(const BodyDef& def);               // b2world.d:75-99
  void destroyBody(Body* b);                          // b2world.d:105-191
  void setBodyType(Body* b, int type);                // b2body.d:867-914
  void setBodyActive(Body* b, bool flag);             // b2body.d:718-775
  Fixture* createFixture(Body* b, const FixtureDef& def);  // b2body.d:116-155
  void destroyFixture(Fixture* f);                    // b2body.d:179-247
  Joint* addJoint(Joint* j);                          // b2world.d:196-261 (takes ownership)
  void destroyJoint(Joint* j);                        // b2world.d:265-360
  void step(float dt, int velocityIterations, int positionIterations);  // b2world.d:367-434
  void stepHalves(float dt, int velocityIterations, int positionIterations, int halves);
  float pendDt = 0; int pendVi = 0, pendPi = 0; bool midStep = false;
  void clearForces();

  // contact manager (b2contactmanager.d)
  void addPair(void* proxyUserDataA, void* proxyUserDataB);
  void findNewContacts();
  void destroyContact(Contact* c);
  void collide();

  void solve(const TimeStep& step);
  void solveTOI(const TimeStep& step);

  // state
  BroadPhase broadPhase;
  Contact* contactList = nullptr; int contactCount = 0;
  Body* bodyList = nullptr; Joint* jointList = nullptr;
  int bodyCount = 0, jointCount = 0;
  V2 gravity; bool allowSleep = true;
  bool newFixture = false, locked = false, clearForcesFlag = true;
  float inv_dt0 = 0;
  bool warmStarting = true, continuousPhysics = true, subStepping = false, stepComplete = true;
  Profile profile;

  // id tables (C API handles)
  std::vector<Body*> bodiesById; std::vector<Fixture*> fixturesById; std::vector<Joint*> jointsById;
  // diagnostics for tests
  int lastIslandCount = 0; int toiEvents = 0;
  // b2ContactListener.BeginContact / EndContact call log, in the reference's call order (b2contact.d:338-346,
  // b2contactmanager.d:60-63); phase 1 = Collide, 2 = SolveTOI, 3 = API call between steps
  struct ContactEvt { int type, phase, step, fixtureA, childA, fixtureB, childB, bodyA, bodyB; };
  std::vector<ContactEvt> contactEvents; bool recordContactEvents = false; int evPhase = 3; int stepCount = 0;
  void logContactEvent(int type, const Contact* c);
  // b2ContactListener.PostSolve call log of the last step (b2Island.Report, b2island.d:438-462; call sites :239 and :414)
  struct PostSolveRec { int phase, fixtureA, childA, fixtureB, childB, count; float normalImpulses[2], tangentImpulses[2]; };
  std::vector<PostSolveRec> postSolveLog; bool recordPostSolve = false;
  void shiftOrigin(V2 newOrigin);                     // b2world.d:758-780
  // b2World.SetContactFilter (b2world.d:52-56): a user b2ContactFilter.ShouldCollide (b2worldcallbacks.d:55-66) REPLACES the
  // default one at both call sites (b2contactmanager.d:110-114, 274-281); callback(fixtureA id, fixtureB id, default answer)
  typedef int (*UserFilter)(int fixtureA, int fixtureB, int defaultAnswer);
  UserFilter userFilter = nullptr;
  std::vector<std::pair<FixtureProxy*, FixtureProxy*>> lastPairs;  // unique pairs handed to AddPair by the last UpdatePairs
  std::vector<Contact*> lastSolveOrder;   // contacts in the order islands solved them in the last Solve
  std::vector<int> lastJointOrder;        // joint ids, likewise
  // Test hook: solve every island's contacts / joints in a caller-supplied order instead of DFS order (one-shot, consumed by
  // the next Solve).  A coloured Gauss-Seidel sweep is a topological re-ordering of SOME sequential sweep; handing that sweep
  // to the oracle lets a test compare the device solver with the sequential algorithm at sizes where the DFS order would
  // need more levels than the device's level override holds.  Contact::orderRank and Joint::orderRank live in ONE rank space:
  // the island walks joints and contacts merged by rank (ties: joints first, then island order).  orderPositionMode: the
  // position iterations walk the same sequence forwards (0), backwards (1: the device runs its position phases downwards), or
  // its contacts and then its joints (2: the reference's own split).
  bool orderOverride = false; int orderPositionMode = 0;
  // state import (tests, bench transplant): a contact exactly as recorded, no filtering, no Evaluate
  Contact* importContact(Fixture* fA, int iA, Fixture* fB, int iB);
  void clearContacts();
};

}  // namespace orc
