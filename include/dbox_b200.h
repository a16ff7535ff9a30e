/* dbox_b200.h — the drop-in boundary: a C ABI for dbox's per-step world pipeline on one B200.
 *
 * The reference (d-gamedev-team/dbox, a D port of Box2D 2.3.1) has no FFI seam: D user code calls
 * methods on b2World / b2Body / b2Fixture / b2Joint directly.  The seam is therefore introduced one
 * level below that public API: the D classes keep their signatures and forward to these functions
 * (the `extern(C)` block a maintainer adds is shown in INTEGRATION.md).  Each entry point cites the
 * reference method it stands in for, relative to /root/reference/src/dbox/.
 *
 * Conventions
 *  - plain C: opaque world pointer, int32 handles (ids are dense, never reused while alive), POD
 *    structs, caller-owned buffers; no C++/torch types.
 *  - status returns: >= 0 success (often an id or a count), < 0 a DBX_E_* code.  The reference
 *    asserts / silently returns null when the world is locked (dynamics/b2world.d:77-82); here the
 *    same calls return DBX_E_LOCKED.
 *  - one CUDA stream per world; calls are synchronous with respect to the host unless noted.
 *  - all world state (bodies, fixtures, proxies, contacts, manifolds, joints, constraints) lives in
 *    SoA device buffers; there is NO CPU fallback: without a CUDA device dbx_world_create fails with
 *    DBX_E_NO_DEVICE and every other call on a null world returns DBX_E_INVALID.
 */
#ifndef DBOX_B200_H_
#define DBOX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBX_ABI_VERSION 1

typedef struct dbx_world dbx_world; /* opaque */

enum {
  DBX_OK = 0,
  DBX_E_INVALID = -1,    /* bad handle / argument */
  DBX_E_LOCKED = -2,     /* world is inside Step (reference: IsLocked()) */
  DBX_E_NO_DEVICE = -3,  /* no usable CUDA device */
  DBX_E_CUDA = -4,       /* a CUDA call failed; see dbx_last_error() */
  DBX_E_CAPACITY = -5,   /* a device pool overflowed (pairs / contacts); see dbx_world_caps */
  DBX_E_UNSUPPORTED = -6 /* feature outside the hot-path scope of this build */
};

/* common/b2math.d:60-190 */
typedef struct { float x, y; } dbx_vec2;
typedef struct { dbx_vec2 lo, hi; } dbx_aabb; /* collision/b2collision.d:283-409 */

/* dynamics/b2body.d:35-43 */
enum { DBX_STATIC_BODY = 0, DBX_KINEMATIC_BODY = 1, DBX_DYNAMIC_BODY = 2 };
/* dynamics/b2body.d:1118-1127 */
enum {
  DBX_BODY_ISLAND = 0x0001, DBX_BODY_AWAKE = 0x0002, DBX_BODY_AUTOSLEEP = 0x0004, DBX_BODY_BULLET = 0x0008,
  DBX_BODY_FIXED_ROTATION = 0x0010, DBX_BODY_ACTIVE = 0x0020, DBX_BODY_TOI = 0x0040
};
/* dynamics/contacts/b2contact.d:242-261 */
enum {
  DBX_CONTACT_ISLAND = 0x0001, DBX_CONTACT_TOUCHING = 0x0002, DBX_CONTACT_ENABLED = 0x0004,
  DBX_CONTACT_FILTER = 0x0008, DBX_CONTACT_BULLET_HIT = 0x0010, DBX_CONTACT_TOI = 0x0020
};
/* collision/shapes/b2shape.d:45-52 */
enum { DBX_SHAPE_CIRCLE = 0, DBX_SHAPE_EDGE = 1, DBX_SHAPE_POLYGON = 2, DBX_SHAPE_CHAIN = 3 };
/* collision/b2collision.d:98-103 */
enum { DBX_MANIFOLD_CIRCLES = 0, DBX_MANIFOLD_FACE_A = 1, DBX_MANIFOLD_FACE_B = 2 };
/* dynamics/joints/b2joint.d:28-42 */
enum {
  DBX_JOINT_UNKNOWN = 0, DBX_JOINT_REVOLUTE = 1, DBX_JOINT_PRISMATIC = 2, DBX_JOINT_DISTANCE = 3, DBX_JOINT_PULLEY = 4,
  DBX_JOINT_MOUSE = 5, DBX_JOINT_GEAR = 6, DBX_JOINT_WHEEL = 7, DBX_JOINT_WELD = 8, DBX_JOINT_FRICTION = 9,
  DBX_JOINT_ROPE = 10, DBX_JOINT_MOTOR = 11
};
/* world toggles: dynamics/b2world.d:622-753 (SetAllowSleeping / SetWarmStarting / SetContinuousPhysics /
 * SetSubStepping / SetAutoClearForces).  Sub-stepping: one TOI event per Step, the following Steps run Collide and resume
 * SolveTOI without Solve until no event is left (b2world.d:1127-1146, 1441-1446); not available on replicated worlds. */
enum {
  DBX_WORLD_ALLOW_SLEEP = 0x01, DBX_WORLD_WARM_STARTING = 0x02, DBX_WORLD_CONTINUOUS = 0x04,
  DBX_WORLD_SUB_STEPPING = 0x08, DBX_WORLD_AUTO_CLEAR_FORCES = 0x10,
  DBX_WORLD_DEFAULT_FLAGS = 0x01 | 0x02 | 0x04 | 0x10 /* dynamics/b2world.d:876-885 */
};

/* dynamics/b2body.d:51-104 (b2BodyDef; same defaults via dbx_default_body_def) */
typedef struct {
  int32_t type;
  dbx_vec2 position;
  float angle;
  dbx_vec2 linearVelocity;
  float angularVelocity;
  float linearDamping, angularDamping;
  int32_t allowSleep, awake, fixedRotation, bullet, active;
  float gravityScale;
  uint64_t userData;
} dbx_body_def;

/* One POD for the four shape classes (collision/shapes/b2circleshape.d:158, b2edgeshape.d:189-193,
 * b2polygonshape.d:562-565, b2chainshape.d:257-263).  Polygon vertices/normals/centroid are the
 * already-processed values (b2PolygonShape.Set / SetAsBox run on the caller's side; helpers below). */
typedef struct {
  int32_t type;
  float radius;
  dbx_vec2 p;                 /* circle centre */
  dbx_vec2 v0, v1, v2, v3;    /* edge: v1-v2 with optional ghost vertices */
  int32_t hasV0, hasV3;
  dbx_vec2 centroid;          /* polygon */
  dbx_vec2 vertices[8];
  dbx_vec2 normals[8];
  int32_t count;
  const dbx_vec2* chainVertices; /* chain: chainCount vertices (a loop repeats vertex 0 at the end) */
  int32_t chainCount;
  dbx_vec2 prevVertex, nextVertex;
  int32_t hasPrev, hasNext;
} dbx_shape;

/* dynamics/b2fixture.d:32-73 (b2Filter + b2FixtureDef) */
typedef struct {
  float friction, restitution, density;
  int32_t isSensor;
  uint16_t categoryBits, maskBits;
  int16_t groupIndex;
  uint16_t _pad;
  uint64_t userData;
} dbx_fixture_def;

/* dynamics/joints/b2joint.d:77-93 + b2revolutejoint.d:39-107 + b2distancejoint.d:36-90 */
typedef struct {
  int32_t type;
  int32_t bodyA, bodyB;
  int32_t collideConnected;
  dbx_vec2 localAnchorA, localAnchorB;
  /* revolute */
  float referenceAngle;
  int32_t enableLimit;
  float lowerAngle, upperAngle;
  int32_t enableMotor;
  float motorSpeed, maxMotorTorque;
  /* distance (frequencyHz / dampingRatio also: weld, wheel, mouse) */
  float length, frequencyHz, dampingRatio;
  uint64_t userData;
  /* ---- second-wave joints; a def only reads the fields of its own type (+ the common head above) ------------------
   * prismatic b2prismaticjoint.d:39-113 (localAxisA, referenceAngle, enableLimit, lower/upperTranslation, enableMotor,
   *   maxMotorForce, motorSpeed) · wheel b2wheeljoint.d:39-92 (localAxisA, enableMotor, maxMotorTorque, motorSpeed,
   *   frequencyHz, dampingRatio) · weld b2weldjoint.d:38-80 (referenceAngle, frequencyHz, dampingRatio) ·
   * rope b2ropejoint.d:40-66 (maxLength) · friction b2frictionjoint.d:36-72 (maxForce, maxTorque) ·
   * motor b2motorjoint.d:36-80 (linearOffset, angularOffset, maxForce, maxTorque, correctionFactor; anchors unused) ·
   * mouse b2mousejoint.d:36-66 (target, maxForce, frequencyHz, dampingRatio; bodyA is only a placeholder) ·
   * pulley b2pulleyjoint.d:41-100 (groundAnchorA/B, lengthA, lengthB, ratio) */
  dbx_vec2 localAxisA;
  float lowerTranslation, upperTranslation, maxMotorForce;
  float maxLength;
  float maxForce, maxTorque;
  dbx_vec2 linearOffset;
  float angularOffset, correctionFactor;
  dbx_vec2 target;
  dbx_vec2 groundAnchorA, groundAnchorB;
  float lengthA, lengthB, ratio;
  int32_t joint1, joint2;     /* gear b2gearjoint.d:36-58: ids of two revolute / prismatic joints (+ ratio); bodyA / bodyB are derived */
  int32_t _pad;
} dbx_joint_def;

/* full per-body state (dynamics/b2body.d:1182-1218) */
typedef struct {
  int32_t type;
  uint32_t flags;
  dbx_vec2 p;            /* m_xf.p */
  float qs, qc;          /* m_xf.q */
  dbx_vec2 localCenter, c0, c; /* m_sweep */
  float a0, a, alpha0;
  dbx_vec2 v;            /* m_linearVelocity */
  float w;               /* m_angularVelocity */
  dbx_vec2 force;
  float torque;
  float mass, invMass, I, invI;
  float linearDamping, angularDamping, gravityScale;
  float sleepTime;
} dbx_body_state;

/* collision/b2collision.d:72-114 (b2ManifoldPoint, b2Manifold) */
typedef struct { dbx_vec2 localPoint; float normalImpulse, tangentImpulse; uint32_t key; } dbx_manifold_point;
typedef struct {
  dbx_manifold_point points[2];
  dbx_vec2 localNormal, localPoint;
  int32_t type, pointCount;
} dbx_manifold;

/* one persistent contact (dynamics/contacts/b2contact.d:441-465), identified by (fixture, child) pairs */
typedef struct {
  int32_t fixtureA, fixtureB, childA, childB;
  uint32_t flags;
  dbx_manifold manifold;
  float friction, restitution, tangentSpeed;
  int32_t toiCount;
  float toi;
} dbx_contact_rec;

/* one broadphase proxy (dynamics/b2fixture.d:76-82 + the tree's fat AABB, collision/b2dynamictree.d:146-180) */
typedef struct {
  int32_t fixture, child;
  int32_t proxyId;  /* reference tree-node id: decides A/B order of pairs (collision/b2broadphase.d:289-290) */
  dbx_aabb aabb;    /* tight swept AABB */
  dbx_aabb fat;     /* persistent fat AABB */
} dbx_proxy_rec;

/* joint solver state that persists across steps */
typedef struct { int32_t type; float impulse[3]; float motorImpulse; int32_t limitState; } dbx_joint_state;

/* dynamics/b2world.d:677-716 counts */
typedef struct { int32_t bodies, fixtures, proxies, contacts, touching, joints, awakeBodies, colours, islands, moves, pairs; } dbx_counts;

/* dynamics/b2timestep.d:37-47 (b2Profile, ms) — filled from CUDA events per phase */
typedef struct { float step, collide, solve, solveInit, solveVelocity, solvePosition, broadphase, solveTOI; } dbx_profile;

/* pool sizing (0 = default).  All pools are device-resident; overflow => DBX_E_CAPACITY. */
typedef struct {
  int32_t maxBodies, maxProxies, maxContacts, maxJoints, maxPairs;
} dbx_caps;

/* ---- library ---- */
int32_t dbx_abi_version(void);
const char* dbx_last_error(void);
int32_t dbx_device_count(void);

/* ---- defaults: dynamics/b2body.d:55-103, dynamics/b2fixture.d:59-72, joint def ctors ---- */
void dbx_default_body_def(dbx_body_def* out);
void dbx_default_fixture_def(dbx_fixture_def* out);
void dbx_default_joint_def(dbx_joint_def* out, int32_t type);

/* ---- shape helpers (host side of the shim; run at setup time) ---- */
void dbx_shape_set_circle(dbx_shape* out, float px, float py, float radius);               /* b2circleshape.d */
void dbx_shape_set_edge(dbx_shape* out, dbx_vec2 v1, dbx_vec2 v2);                         /* b2edgeshape.d:47-53 */
void dbx_shape_set_box(dbx_shape* out, float hx, float hy);                                /* b2polygonshape.d:220-232 */
void dbx_shape_set_box_at(dbx_shape* out, float hx, float hy, dbx_vec2 center, float angle); /* :239-262 */
int32_t dbx_shape_set_polygon(dbx_shape* out, const dbx_vec2* pts, int32_t n);             /* :74-209 */
void dbx_shape_set_chain(dbx_shape* out, const dbx_vec2* pts, int32_t n, int32_t loop);    /* b2chainshape.d:64-117; pts must outlive fixture creation */

/* ---- world lifecycle: dynamics/b2world.d:865-919 ---- */
dbx_world* dbx_world_create(float gx, float gy, int32_t device, const dbx_caps* caps);
void dbx_world_destroy(dbx_world* w);
int32_t dbx_world_set_flags(dbx_world* w, uint32_t flags);            /* b2world.d:622-753 */
uint32_t dbx_world_get_flags(dbx_world* w);
int32_t dbx_world_set_gravity(dbx_world* w, float gx, float gy);      /* b2world.d:719-728 */

/* ---- object lifecycle ---- */
int32_t dbx_body_create(dbx_world* w, const dbx_body_def* def);                                   /* b2world.d:75-99 */
int32_t dbx_body_destroy(dbx_world* w, int32_t body);                                             /* b2world.d:105-191 */
int32_t dbx_fixture_create(dbx_world* w, int32_t body, const dbx_fixture_def* def, const dbx_shape* shape); /* b2body.d:116-155 */
int32_t dbx_fixture_destroy(dbx_world* w, int32_t fixture);                                       /* b2body.d:179-247 */
int32_t dbx_joint_create(dbx_world* w, const dbx_joint_def* def);                                 /* b2world.d:196-261 */
int32_t dbx_joint_destroy(dbx_world* w, int32_t joint);                                           /* b2world.d:265-360 */
int32_t dbx_joint_set_target(dbx_world* w, int32_t joint, float x, float y);                       /* b2MouseJoint.SetTarget b2mousejoint.d:112-120 (wakes bodyB) */

/* ---- the hot path: dynamics/b2world.d:367-434 (Collide -> Solve -> SolveTOI -> ClearForces) ---- */
int32_t dbx_world_step(dbx_world* w, float dt, int32_t velocityIterations, int32_t positionIterations);
/* n consecutive steps without returning to the host in between (RL / benchmark loops) */
int32_t dbx_world_step_n(dbx_world* w, float dt, int32_t velocityIterations, int32_t positionIterations, int32_t n);
int32_t dbx_world_clear_forces(dbx_world* w);                          /* b2world.d:443-450 */

/* measurement + bulk per-step I/O (RL loops): n steps each bracketed by CUDA events on the world's stream, optional >L2
 * flush between steps (untimed); stageMs[9] = average ms of {collide, islands, colour+sort, prepare, solve, sync_fixtures,
 * find_new_contacts, toi, clear_forces}.  apply_forces adds (fx, fy, torque, -) per body like b2Body.ApplyForceToCenter /
 * ApplyTorque with wake=false (b2body.d:390-431); read_transforms returns (p.x, p.y, sin, cos) per body. */
int32_t dbx_world_time_steps(dbx_world* w, float dt, int32_t velocityIterations, int32_t positionIterations, int32_t n, int32_t flushL2,
                             float* totalMs, float* stageMs);
int32_t dbx_world_apply_forces(dbx_world* w, const float* fx_fy_torque_pad, int32_t n);
/* Bulk b2Body.SetTransform (dynamics/b2body.d:261-285: pose, sweep c0 = c, b2Fixture.Synchronize with zero displacement)
 * and SetLinearVelocity / SetAngularVelocity (:296-326: a non-zero velocity wakes the body) for n bodies from host arrays
 * of 4 floats per body; either array may be NULL (that part is left alone); ids NULL = bodies 0..n-1.  Works on replicated
 * worlds (body r*B+b is body b of replica r): the reset / randomisation call of a batched-worlds loop.  Returns n. */
int32_t dbx_world_set_body_states(dbx_world* w, const int32_t* ids, const float* x_y_angle_pad, const float* vx_vy_w_pad, int32_t n);
int32_t dbx_world_read_transforms(dbx_world* w, float* out, int32_t n);
int64_t dbx_world_launch_count(dbx_world* w);   /* kernels of this library launched so far on this world */

/* ---- body accessors / mutators: dynamics/b2body.d ---- */
int32_t dbx_body_get_state(dbx_world* w, int32_t body, dbx_body_state* out);
int32_t dbx_body_set_transform(dbx_world* w, int32_t body, float x, float y, float angle);   /* b2body.d:261-285 */
int32_t dbx_body_set_linear_velocity(dbx_world* w, int32_t body, float vx, float vy);       /* :322-335 */
int32_t dbx_body_set_angular_velocity(dbx_world* w, int32_t body, float omega);             /* :346-359 */
int32_t dbx_body_apply_force(dbx_world* w, int32_t body, float fx, float fy, float px, float py, int32_t wake); /* :367-385 */
int32_t dbx_body_apply_torque(dbx_world* w, int32_t body, float torque, int32_t wake);      /* :414-431 */
int32_t dbx_body_apply_linear_impulse(dbx_world* w, int32_t body, float ix, float iy, float px, float py, int32_t wake); /* :439-457 */
int32_t dbx_body_apply_angular_impulse(dbx_world* w, int32_t body, float impulse, int32_t wake); /* :462-479 */
int32_t dbx_body_set_awake(dbx_world* w, int32_t body, int32_t flag);                       /* :827-846 */
int32_t dbx_body_set_bullet(dbx_world* w, int32_t body, int32_t flag);                      /* :784-794 */
int32_t dbx_body_set_sleeping_allowed(dbx_world* w, int32_t body, int32_t flag);            /* :804-815 */
int32_t dbx_body_set_mass_data(dbx_world* w, int32_t body, float mass, float centerX, float centerY, float I); /* :502-540 */
int32_t dbx_body_reset_mass_data(dbx_world* w, int32_t body);                                  /* :555-625 */
int32_t dbx_body_set_fixed_rotation(dbx_world* w, int32_t body, int32_t flag);                 /* :924-945 */
int32_t dbx_body_set_linear_damping(dbx_world* w, int32_t body, float damping);                /* :653-656 */
int32_t dbx_body_set_angular_damping(dbx_world* w, int32_t body, float damping);               /* :665-668 */
int32_t dbx_body_set_gravity_scale(dbx_world* w, int32_t body, float scale);                   /* :677-680 */
/* fixture mutators (dynamics/b2fixture.d): SetFilterData + Refilter :131-178 (contacts flagged for filtering, proxies touched),
 * SetSensor :108-115 (wakes the body), SetFriction / SetRestitution :225-249 (existing contacts keep their mixed values),
 * SetDensity :257-262 (counts at the next ResetMassData) */
int32_t dbx_fixture_set_filter(dbx_world* w, int32_t fixture, int32_t categoryBits, int32_t maskBits, int32_t groupIndex);
int32_t dbx_fixture_set_sensor(dbx_world* w, int32_t fixture, int32_t flag);
int32_t dbx_fixture_set_friction(dbx_world* w, int32_t fixture, float friction);
int32_t dbx_fixture_set_restitution(dbx_world* w, int32_t fixture, float restitution);
int32_t dbx_fixture_set_density(dbx_world* w, int32_t fixture, float density);
int32_t dbx_body_set_type(dbx_world* w, int32_t body, int32_t type);                           /* :867-914: mass reset, contacts destroyed, proxies touched */
int32_t dbx_body_set_active(dbx_world* w, int32_t body, int32_t flag);                         /* :718-775: proxies created / destroyed, contacts destroyed */

/* ---- bulk state access (one D2H / H2D copy each; also the checkpoint / parity-injection path) ---- */
int32_t dbx_world_counts(dbx_world* w, dbx_counts* out);
int32_t dbx_world_profile(dbx_world* w, dbx_profile* out);                                    /* b2world.d:789-792 */
int32_t dbx_world_read_bodies(dbx_world* w, dbx_body_state* out, int32_t cap);                /* index = body id */
int32_t dbx_world_write_bodies(dbx_world* w, const dbx_body_state* in, int32_t n);
int32_t dbx_world_read_contacts(dbx_world* w, dbx_contact_rec* out, int32_t cap);             /* b2world.d:610-613 (GetContactList) */
int32_t dbx_world_write_contacts(dbx_world* w, const dbx_contact_rec* in, int32_t n);         /* replaces the pair cache */
int32_t dbx_world_read_proxies(dbx_world* w, dbx_proxy_rec* out, int32_t cap);
int32_t dbx_world_write_proxies(dbx_world* w, const dbx_proxy_rec* in, int32_t n);            /* fat/tight AABBs by (fixture, child) */
int32_t dbx_world_read_joints(dbx_world* w, dbx_joint_state* out, int32_t cap);
int32_t dbx_world_write_joints(dbx_world* w, const dbx_joint_state* in, int32_t n);
int32_t dbx_world_read_moves(dbx_world* w, int32_t* fixture_child_pairs, int32_t cap);        /* pending move buffer (b2broadphase.d:244-257) */
int32_t dbx_world_write_moves(dbx_world* w, const int32_t* fixture_child_pairs, int32_t n);
int32_t dbx_world_get_inv_dt0(dbx_world* w, float* out);                                      /* b2world.d:421-424 */
int32_t dbx_world_set_inv_dt0(dbx_world* w, float inv_dt0);

/* ---- staged stepping for parity tests and listener round-trips (same kernels as dbx_world_step) ---- */
int32_t dbx_world_stage_find_new_contacts(dbx_world* w);   /* b2contactmanager.d:178-181 */
int32_t dbx_world_stage_collide(dbx_world* w);             /* b2contactmanager.d:251-317 */
int32_t dbx_world_read_pairs(dbx_world* w, int32_t* fixA_childA_fixB_childB, int32_t cap); /* last UpdatePairs' unique pair set */
/* override the constraint colouring with a caller-supplied schedule: level[i] for contact rec i of the last
 * dbx_world_read_contacts order; lets a test replay the reference's exact sequential order. n=0 clears. */
int32_t dbx_world_debug_set_contact_levels(dbx_world* w, const int32_t* levels, int32_t n);

/* the Gauss-Seidel schedule the LAST step actually ran, for handing it to a sequential checker.  Joints and contacts share ONE
 * rank space: contactRank[i] for contact rec i of the last dbx_world_read_contacts order, jointRank[j] for joint id j (-1 = not
 * in the solver); a velocity pass walks the constraints by ascending rank (equal ranks never share a dynamic body; a body's
 * joints come before its contacts, dynamics/b2island.d:153-161).  info[0] = 1: position passes walk the same order backwards
 * (a body's contacts before its joints, :206-216); info[1] = contact colours, info[2] = joint colours, info[3] = tiles of the
 * tile solver (0: the grid-phase solver ran). */
int32_t dbx_world_debug_read_solve_order(dbx_world* w, int32_t* contactRank, int32_t capContacts, int32_t* jointRank, int32_t capJoints, int32_t* info4);

/* number of solver-contact pairs that shared a dynamic body AND a colour in the last step (must be 0) */
int32_t dbx_world_debug_colour_conflicts(dbx_world* w);

/* ---- parity hooks: the device narrowphase / GJK / TOI functions on caller-supplied inputs, one CUDA thread per item.
 * xf = {p.x, p.y, sin, cos} per item; sweep = {localCenter.xy, c0.xy, c.xy, a0, a, alpha0} per item; chain shapes are
 * passed as their child edge.  Replaces, for tests: collision/b2collide*.d, b2distance.d:185-347, b2timeofimpact.d:67-302 */
int32_t dbx_debug_collide(int32_t device, int32_t n, const dbx_shape* shapesA, const float* xfA, const dbx_shape* shapesB, const float* xfB, dbx_manifold* out);
int32_t dbx_debug_distance(int32_t device, int32_t n, const dbx_shape* shapesA, const float* xfA, const dbx_shape* shapesB, const float* xfB, int32_t useRadii,
                           float* outDistance, dbx_vec2* outA, dbx_vec2* outB, int32_t* outIterations);
int32_t dbx_debug_time_of_impact(int32_t device, int32_t n, const dbx_shape* shapesA, const float* sweepsA, const dbx_shape* shapesB, const float* sweepsB,
                                 float tMax, int32_t* outState, float* outT);

/* first call enables %globaltimer stamps after every barrier of the persistent solve kernel; later calls return the last step's stamps (ns) */
int32_t dbx_world_debug_phase_times(dbx_world* w, uint64_t* out, int32_t cap);
int32_t dbx_world_debug_header(dbx_world* w, void* out, int32_t bytes); /* raw copy of the device-side counter block (debug) */
float dbx_debug_barrier_us(int32_t device, int32_t blocks, int32_t threads, int32_t iters); /* grid-barrier latency microbenchmark */

/* ---- batched independent worlds (config 5): replicas of a template share one device world ---- */
/* The one collective of the batched path: the end-of-run statistics of the ranks of a batch, reduced in
 * place through the caller's NCCL communicator -- sums[0 .. nSums) summed, maxima[0 .. nMax) max-reduced (counts and times:
 * world-steps, contacts, awake bodies; the slowest rank's seconds).  `ncclComm` is an ncclComm_t, `cudaStream` a cudaStream_t
 * (NULL: the default stream).  The library does not link NCCL: it resolves ncclAllReduce from the NCCL the host program has
 * already loaded (or libnccl.so.2).  Nothing in the stepping path communicates; the reference has no counterpart (a dbox user
 * with many worlds steps many b2World objects, dynamics/b2world.d:34-40). */
int32_t dbx_stats_allreduce(void* ncclComm, void* cudaStream, double* sums, int32_t nSums, double* maxima, int32_t nMax);
int32_t dbx_world_replicate(dbx_world* w, int32_t copies);  /* world becomes `copies` disjoint replicas of its current content */
int32_t dbx_world_replica_count(dbx_world* w);

/* ---- snapshot / restore ---------------------------------------------------------------------------------------------
 * The reference can only Dump() D source text that rebuilds a scene (dynamics/b2world.d:796-855).  Here the dynamic state of
 * a world -- body states, proxy boxes (tight + fat), the contact cache with manifolds, impulses and solver colours, joint
 * impulses, the pending move buffer, inv_dt0 -- goes to / comes from one flat buffer.  The scene itself (bodies, fixtures, joints) is not
 * in the blob: import into a world built the same way.  export returns the bytes needed (call with buf = NULL to size it);
 * stepping an imported world continues bit for bit where the exported one would have. */
int64_t dbx_world_export_state(dbx_world* w, void* buf, int64_t cap);
int32_t dbx_world_import_state(dbx_world* w, const void* buf, int64_t n);

/* ---- world queries on the device tree, batched ---------------------------------------------
 * b2World.RayCast (dynamics/b2world.d:577-587; wrapper :1605-1624; b2DynamicTree.RayCast collision/b2dynamictree.d:237-331;
 * b2Shape.RayCast b2circleshape.d:67-94, b2edgeshape.d:96-150, b2polygonshape.d:279-332, b2chainshape.d:204-223) and
 * b2World.QueryAABB (b2world.d:563-570, b2dynamictree.d:203-235).  The reference hands one ray / one box to a user
 * callback; here a whole batch goes to the device (one thread per ray, one warp per box) and the answers come back in
 * arrays.  raycast_closest is the callback that returns `fraction` (the testbed's RayCastClosestCallback): the nearest
 * hit along p1 -> p2, fixture = -1 for a miss.  query_aabb reports the (fixture, child) pairs whose FAT proxy box overlaps,
 * as the reference does, sorted by (fixture, child); counts[k] may exceed capPerQuery (the surplus is dropped).
 * Both see the world as it is after the last step / edit. */
typedef struct dbx_ray { dbx_vec2 p1, p2; } dbx_ray;
typedef struct dbx_ray_hit { int32_t fixture, child; float fraction; dbx_vec2 point, normal; } dbx_ray_hit;
int32_t dbx_world_raycast_closest(dbx_world* w, const dbx_ray* rays, int32_t n, dbx_ray_hit* out);
int32_t dbx_world_query_aabb(dbx_world* w, const dbx_aabb* boxes, int32_t n, int32_t capPerQuery, int32_t* counts, int32_t* fixture_child);

/* ---- contact listener, deferred ------------------------------------------------------------
 * b2World.SetContactListener (dynamics/b2world.d:62-66) + b2ContactListener.BeginContact / EndContact
 * (dynamics/b2worldcallbacks.d:87-95).  The reference calls the listener in the middle of the step, from b2Contact.Update
 * (contacts/b2contact.d:338-346; also inside the TOI loop, dynamics/b2world.d:1295,1379) and from b2ContactManager.Destroy
 * (dynamics/b2contactmanager.d:60-63).  Device code cannot call back into the host, so the same call sites append records
 * to a device buffer and the host shim delivers them right after dbx_world_step returns: same events, same fixtures,
 * later in time.  Consequence: a listener cannot change the step it is told about through THESE calls; PreSolve's
 * SetEnabled(false) / friction edits go through the split step below (dbx_world_step_begin / patch_contacts / step_end),
 * PostSolve's impulses through dbx_world_enable_post_solve / dbx_world_read_post_solve.  Events are ordered (step, phase, pair key, type); phase 1 = Collide,
 * 2 = TOI sub-steps, 3 = contact destroyed by an API call after that step (DestroyBody / DestroyFixture / CreateJoint). */
typedef struct dbx_contact_event {
  int32_t type;               /* DBX_CONTACT_BEGIN / DBX_CONTACT_END */
  int32_t phase;
  int32_t stepsAgo;           /* 0 = the step that just ran */
  int32_t fixtureA, fixtureB; /* in the contact's A/B order (b2contact.d:375-400) */
  int32_t childA, childB;
  int32_t bodyA, bodyB;
} dbx_contact_event;
enum { DBX_CONTACT_BEGIN = 1, DBX_CONTACT_END = 2 };
/* capacity > 0: start recording (at most `capacity` events between two polls), 0: stop.  Returns the capacity in use. */
int32_t dbx_world_enable_contact_events(dbx_world* w, int32_t capacity);
/* Copies the events recorded since the last poll (sorted) and clears the buffer; returns their number, or
 * DBX_E_CAPACITY if more than `capacity` were produced (the surplus is lost).  out == NULL: returns the count only. */
int32_t dbx_world_poll_contact_events(dbx_world* w, dbx_contact_event* out, int32_t cap);

/* ---- PreSolve: the step cut where the reference calls it ---------------------------------------------------------------
 * b2ContactListener.PreSolve runs inside b2Contact.Update for every touching non-sensor contact (contacts/b2contact.d:348-355)
 * and may call b2Contact.SetEnabled(false) (this step only: Update re-enables, :272), SetFriction, SetRestitution,
 * SetTangentSpeed (:137-205).  A shim that has a PreSolve listener steps like this:
 *     dbx_world_step_begin(w, dt, vi, pi);          // FindNewContacts-if-needed + Collide (b2world.d:372-399)
 *     dbx_world_read_contacts(...)                  // touching, non-sensor contacts -> listener.PreSolve(contact, oldManifold = none)
 *     dbx_world_patch_contacts(w, patches, n);      // what the listener changed
 *     dbx_world_step_end(w);                        // Solve, SolveTOI, ClearForces (b2world.d:401-431)
 * and is bit-identical to dbx_world_step when nothing is patched.
 * The TOI loop calls b2Contact.Update -- and with it PreSolve -- again for the contacts of a TOI event (b2world.d:1295,1379), in
 * the middle of step_end where no listener can be reached.  There the answer already given stands: a contact patched with
 * enabled = 0 stays disabled through every re-evaluation of this step (the listener, asked again about the same contact in the
 * same step, is taken to answer the same) and is enabled again by the next step's Collide, as in the reference (b2contact.d:272).
 * A fast body may TOUCH for the first time inside the TOI loop; patches are accepted for contacts that are not touching yet
 * (they exist as soon as the fat AABBs overlap), so a shim can ask its listener ahead of time ("if this pair touches during
 * this step ...") -- dbox_b200.world.b2World.StepWithPreSolve(..., toi_lookahead=True) does.  Not covered: contacts that are
 * CREATED inside the TOI loop (its own FindNewContacts, b2world.d:1444) and used by a later TOI event of the same step, and the
 * old-manifold argument. */
typedef struct dbx_contact_patch {
  int32_t fixtureA, childA, fixtureB, childB;   /* either order */
  int32_t mask;                                 /* DBX_PATCH_* of the fields to apply */
  int32_t enabled;
  float friction, restitution, tangentSpeed;
} dbx_contact_patch;
enum { DBX_PATCH_ENABLED = 1, DBX_PATCH_FRICTION = 2, DBX_PATCH_RESTITUTION = 4, DBX_PATCH_TANGENT_SPEED = 8,
       DBX_PATCH_DESTROY = 16 /* the user's b2ContactFilter said no: see "user contact filter" below */ };
int32_t dbx_world_step_begin(dbx_world* w, float dt, int32_t velocityIterations, int32_t positionIterations);
int32_t dbx_world_patch_contacts(dbx_world* w, const dbx_contact_patch* patches, int32_t n);   /* unknown pairs are ignored */
int32_t dbx_world_step_end(dbx_world* w);

/* ---- more queries, accessors and world edits either side of the step ---------------------- */
/* b2Fixture.TestPoint (dynamics/b2fixture.d:209-212) -> b2Shape.TestPoint (polygon b2polygonshape.d:265-279, circle
 * b2circleshape.d:60-65; edges and chains contain no point, b2edgeshape.d:84-87, b2chainshape.d:196-199): n (fixture, point)
 * pairs, one device thread each, against the bodies' current transforms; inside[k] = 1 / 0.  The testbed's mouse pick
 * (QueryAABB over a small box, then TestPoint: demos/tests/test.d MouseDown) is dbx_world_query_aabb + this call. */
int32_t dbx_world_test_points(dbx_world* w, const int32_t* fixtures, const dbx_vec2* points, int32_t n, int32_t* inside);
/* b2World.RayCast with the callback that returns 1 ("report every fixture on the ray, do not clip": the testbed's
 * RayCastMultipleCallback, demos/tests/raycast.d): every (fixture, child) the ray p1 -> p2 hits, with fraction, point and
 * normal, sorted by (fraction, fixture, child); counts[k] may exceed capPerRay (the surplus is dropped).  The callback that
 * returns 0 (any hit) is counts[k] > 0. */
int32_t dbx_world_raycast_all(dbx_world* w, const dbx_ray* rays, int32_t n, int32_t capPerRay, int32_t* counts, dbx_ray_hit* hits);
/* b2World.ShiftOrigin (dynamics/b2world.d:758-780): body transforms and sweeps, joint anchors kept in world coordinates
 * (b2mousejoint.d:174-177, b2pulleyjoint.d:227-231) and the broadphase boxes (b2dynamictree.d:503-511) all move by -newOrigin,
 * on the device; the pair cache, manifolds and impulses are local quantities and stay. */
int32_t dbx_world_shift_origin(dbx_world* w, float newOriginX, float newOriginY);
/* b2Contact.GetWorldManifold (contacts/b2contact.d:77-91) -> b2WorldManifold.Initialize (collision/b2collision.d:123-191) for
 * every contact, computed on the device from the bodies' current transforms; records come in dbx_world_read_contacts order. */
typedef struct dbx_world_manifold { dbx_vec2 normal; dbx_vec2 points[2]; float separations[2]; int32_t pointCount; int32_t _pad; } dbx_world_manifold;
int32_t dbx_world_read_world_manifolds(dbx_world* w, dbx_world_manifold* out, int32_t cap);
/* b2ContactListener.PostSolve (dynamics/b2worldcallbacks.d:120-128), deferred like Begin/EndContact: b2Island.Report
 * (dynamics/b2island.d:438-462) runs once per island solve (:239) and once per TOI sub-step (:414) and hands every contact of
 * that island its b2ContactImpulse.  With recording on, the same call sites append one record per contact; read returns the
 * records of the LAST step sorted by (phase, pair key, order of arrival): phase 1 = the island solve, 2 = TOI sub-steps
 * (a contact may appear several times there). */
typedef struct dbx_post_solve {
  int32_t fixtureA, fixtureB, childA, childB;
  int32_t phase, count;                         /* b2ContactImpulse.count = the constraint's point count */
  float normalImpulses[2], tangentImpulses[2];
} dbx_post_solve;
int32_t dbx_world_enable_post_solve(dbx_world* w, int32_t capacity);   /* > 0: record (at most `capacity` per step), 0: stop */
int32_t dbx_world_read_post_solve(dbx_world* w, dbx_post_solve* out, int32_t cap);   /* out == NULL: the count only */

/* ---- user contact filter, deferred ------------------------------------------------------------
 * b2World.SetContactFilter (dynamics/b2world.d:52-56) + b2ContactFilter.ShouldCollide (dynamics/b2worldcallbacks.d:55-66).
 * The reference asks the filter when the broadphase reports a new pair (b2contactmanager.d:110-114: no contact is created on
 * "no") and when a contact flagged by Refilter is visited by Collide (:274-281: the contact is destroyed on "no").  Device
 * code cannot call the filter, so with logging on every contact the broadphase creates is also listed; the shim polls the
 * list after the step (and after any call that runs the broadphase), asks the user's filter and destroys what it rejects
 * with dbx_world_patch_contacts(mask = DBX_PATCH_DESTROY) before the next Collide can evaluate it -- such a contact never
 * touches, never reaches the solver and raises no Begin/EndContact.  Refilter: the shim walks the fixture's contacts
 * (dbx_world_read_contacts) at SetFilterData time and vetoes the same way.  Not covered: contacts the TOI sub-steps create
 * and use inside the same step.
 * mode: 0 = off; DBX_FILTER_LOG = list new contacts; | DBX_FILTER_REPLACES_DEFAULT = the device skips the category / mask /
 * group test (b2worldcallbacks.d:40-52), because the user's ShouldCollide does not call the default one. */
enum { DBX_FILTER_LOG = 1, DBX_FILTER_REPLACES_DEFAULT = 2 };
int32_t dbx_world_set_user_filter(dbx_world* w, int32_t mode);
/* contacts created since the last poll as (fixtureA, childA, fixtureB, childB), ascending pair key; returns their number
 * (out == NULL: the count only, nothing is consumed) */
int32_t dbx_world_poll_new_contacts(dbx_world* w, int32_t* fixA_childA_fixB_childB, int32_t cap);

/* ---- pipelined stepping and bulk I/O (act / step / observe loops) ------------------------------------------------------
 * dbx_world_step, dbx_world_apply_forces and dbx_world_read_transforms are synchronous: the caller gets its answer (and
 * any device error) when they return, and a step's host<->device copies sit between the steps.  The calls below only
 * ENQUEUE: the step goes to the world's stream; the copies go to two copy streams (host -> device, device -> host) that run
 * beside it, ordered against the step by events.  forces must be pinned host memory and stay untouched until the step that
 * consumes them has been enqueued AND a later dbx_world_io_wait / dbx_world_sync has returned; read_transforms_async takes
 * a snapshot of the transforms at its place in the queue (device-to-device, so the next step is free to move on), copies
 * it to pinned `out` on the copy stream and returns a ticket; the data is there once dbx_world_io_wait(w, ticket) returns.
 * Two reads may be in flight (double-buffered snapshot); a third waits for the first.  dbx_world_sync waits for everything
 * and returns the sticky device error (DBX_E_CAPACITY ...) like a synchronous step would have. */
int32_t dbx_world_step_async(dbx_world* w, float dt, int32_t velocityIterations, int32_t positionIterations);
int32_t dbx_world_apply_forces_async(dbx_world* w, const float* pinned_fx_fy_torque_pad, int32_t n);
int32_t dbx_world_read_transforms_async(dbx_world* w, float* pinned_out, int32_t n);
/* record format of the four bulk I/O calls above (apply_forces / read_transforms and their _async forms).  DBX_IO_FULL (default):
 * 16-byte records -- forces (fx, fy, torque, -), transforms (p.x, p.y, sin, cos).  DBX_IO_COMPACT: 12-byte records -- forces
 * (fx, fy, torque), poses (p.x, p.y, angle): a quarter less traffic over PCIe for an act / step / observe loop that moves every
 * body's record both ways each step. */
enum { DBX_IO_FULL = 0, DBX_IO_COMPACT = 1 };
int32_t dbx_world_set_io_format(dbx_world* w, int32_t format);
int32_t dbx_world_io_wait(dbx_world* w, int32_t ticket);
int32_t dbx_world_sync(dbx_world* w);

/* ---- joint parameters at run time -----------------------------------------------------------------
 * The setters of the joint classes: b2RevoluteJoint.EnableLimit / SetLimits / EnableMotor / SetMotorSpeed / SetMaxMotorTorque
 * (joints/b2revolutejoint.d:216-300), the same five of b2PrismaticJoint (b2prismaticjoint.d:250-330), b2WheelJoint.EnableMotor /
 * SetMotorSpeed / SetMaxMotorTorque / SetSpringFrequencyHz / SetSpringDampingRatio (b2wheeljoint.d:170-230), b2DistanceJoint.SetLength /
 * SetFrequency / SetDampingRatio, b2WeldJoint.SetFrequency / SetDampingRatio, b2RopeJoint.SetMaxLength, b2FrictionJoint / b2MouseJoint /
 * b2MotorJoint.SetMaxForce / SetMaxTorque, b2MotorJoint.SetLinearOffset / SetAngularOffset / SetCorrectionFactor.
 * `def` carries the new values in the fields of the joint's own type; `mask` says which groups to take.  Side effects as in the
 * reference: the motor groups wake both bodies; the limit groups, when they change something, wake both bodies and zero the limit
 * impulse; the motor joint's offsets wake both bodies when they change; the other groups wake nobody.  One joint's record is patched
 * in place on the device (no re-upload of the joint set). */
enum {
  DBX_JP_MOTOR_SPEED = 1, DBX_JP_MAX_MOTOR = 2 /* maxMotorTorque | maxMotorForce */, DBX_JP_ENABLE_MOTOR = 4, DBX_JP_ENABLE_LIMIT = 8,
  DBX_JP_LIMITS = 16 /* lower / upper angle | translation */, DBX_JP_SPRING = 32 /* frequencyHz, dampingRatio */,
  DBX_JP_LENGTH = 64 /* distance length | rope maxLength */, DBX_JP_MAX_FORCE = 128 /* maxForce, maxTorque */,
  DBX_JP_OFFSETS = 256 /* motor joint: linearOffset, angularOffset */, DBX_JP_CORRECTION = 512
};
int32_t dbx_joint_set_params(dbx_world* w, int32_t joint, const dbx_joint_def* def, uint32_t mask);
/* Bulk SetMotorSpeed for n revolute / prismatic / wheel joints from host arrays (the actuation call of a control loop), on the
 * device; wakes the bodies of every joint named, like the setter.  Works on replicated worlds: joint r * J + j is joint j of
 * replica r (J = joints per replica). */
int32_t dbx_world_set_motor_speeds(dbx_world* w, const int32_t* joints, const float* speeds, int32_t n);

/* b2World.GetTreeHeight / GetTreeBalance / GetTreeQuality (dynamics/b2world.d:694-716 -> collision/b2dynamictree.d:354-419), the
 * diagnostics some demos print (tiles.d:118-120) -- of THIS library's broadphase tree, the LBVH over the fat AABBs (rebuilt or
 * widened as for a world query), so the numbers describe a different tree than the reference's incremental one: height = edges
 * on the longest root-to-leaf path, balance = largest height difference between the two children of a node, quality = sum of
 * all node perimeters / root perimeter.  "Should not be called often" there; here it costs one download of the tree. */
int32_t dbx_world_tree_stats(dbx_world* w, int32_t* height, int32_t* maxBalance, float* quality);

#ifdef __cplusplus
}
#endif
#endif /* DBOX_B200_H_ */
