// dbx_device.cuh — the SoA device world: every array the step pipeline touches, and the device-side header of
// counters that lets the whole step run without returning to the host.
//
// Layout in HBM (one array per field group, float4-packed so a warp's loads are 512 B coalesced transactions):
//   bodies    (index = body id)      xf, xf0, pos(c,a), pos0(c0,a0,alpha0), vel(v,w), force(f,torque), mass, lc/damping, ...
//   proxies   (dense slot)           fixture, child, body, shape, refKey, tight AABB, fat AABB, moved flag
//   contacts  (slot, free-listed)    pair key, ids, fixtures/shapes, flags, manifold (4 x float4), material, colour
//   joints    (index = joint id)     definition + persistent impulses + per-step temporaries
//   solver    (colour-sorted)        velocity / position constraint blocks built each step from touching contacts
// Reference state being replaced: dynamics/b2body.d:1182-1218, b2fixture.d:76-82,504-521, contacts/b2contact.d:441-465,
// contacts/b2contactsolver.d:32-58,801-814, collision/b2dynamictree.d:31-54.
#pragma once
#include "dbx_narrow.cuh"

namespace dbx {

enum { BODY_STATIC = 0, BODY_KINEMATIC = 1, BODY_DYNAMIC = 2 };
// low 16 bits = b2Body flags (b2body.d:1118-1127); bits 16-17 = body type; bit 20 = slot in use
enum : uint32_t {
  BF_ISLAND = 0x0001, BF_AWAKE = 0x0002, BF_AUTOSLEEP = 0x0004, BF_BULLET = 0x0008, BF_FIXEDROT = 0x0010, BF_ACTIVE = 0x0020, BF_TOI = 0x0040,
  BF_TYPE_SHIFT = 16, BF_TYPE_MASK = 0x30000, BF_ALIVE = 0x100000
};
DBX_HD int body_type(uint32_t f) { return (int)((f & BF_TYPE_MASK) >> BF_TYPE_SHIFT); }
// low bits = b2Contact flags (b2contact.d:242-261); bit 8 = slot alive; bit 9 = either fixture is a sensor
enum : uint32_t {
  CF_ISLAND = 0x0001, CF_TOUCHING = 0x0002, CF_ENABLED = 0x0004, CF_FILTER = 0x0008, CF_BULLET_HIT = 0x0010, CF_TOI = 0x0020,
  CF_ALIVE = 0x0100, CF_SENSOR = 0x0200, CF_SOLVE = 0x0400 /* in an awake island this step */,
  CF_FRESH = 0x0800 /* created by this step's FindNewContacts: the overlapped TOI pre-evaluation must not look at it */,
  CF_NEW = 0x1000 /* created since the host last polled the new-contact list (user contact filter, deferred) */,
  CF_PRESOLVE_OFF = 0x2000 /* this step's PreSolve patch said SetEnabled(false): the Update calls of the TOI loop re-apply it */
};
// what a kernel stores in Header::error: DBX_E_CAPACITY (-5) in the low bits, the pool that overflowed above (host: checkDeviceError)
enum { E_CONTACTS = -5 - 16 * 1, E_PAIRS = -5 - 16 * 2, E_MOVES = -5 - 16 * 3, E_COLOURS = -5 - 16 * 4, E_SOLVER_ROWS = -5 - 16 * 5,
       E_HASH = -5 - 16 * 6, E_QUERY_STACK = -5 - 16 * 7, E_TOI_CANDIDATES = -5 - 16 * 8 };
enum { FXF_SENSOR = 1 };
enum { PF_ALIVE = 1, PF_MOVED = 2 };
enum { JT_REVOLUTE = 1, JT_DISTANCE = 3 };
enum { LIM_INACTIVE = 0, LIM_LOWER = 1, LIM_UPPER = 2, LIM_EQUAL = 3 };

constexpr int kMaxColours = 1024;        // contact colours: 0..63 tracked in per-body bit masks, 64.. = per-body overflow lanes
constexpr int kMaskColours = 64;
constexpr int kMaxJointColours = 64;
constexpr int kSortBlocks = 296;         // 2 CTAs per SM for the colour counting sort
constexpr int kMaxPosIters = 8;
constexpr int kTailContacts = 1024;       // tail colours holding at most this many constraints share one CTA-local phase
constexpr int kTileColours = 64;         // tile solver: colours a tile walks locally (higher ones, the overflow lanes, run as global phases)
enum { XF_FOREIGN = 1 /* moved by the left neighbour's boundary rows */, XF_OWNB = 2 /* by this tile's own boundary rows */, XF_G = 4 /* by global rows */,
       XF_ISLAND = 8 /* alive and in an island this step (BF_ALIVE | BF_ISLAND, copied here by k_mark_solve) */ };
constexpr int kToiCand = 64;             // candidate contacts per event side (mini-island holds at most 32)
enum { TF_INVAL = 1, TF_SYNC = 2 };
constexpr unsigned long long kHashEmpty = ~0ull;
constexpr unsigned long long kHashTomb = ~0ull - 1;

// device-resident counters and per-step scalars (one cache line group; written by kernels, read back rarely)
struct Header {
  int cHigh;          // contact slots in use are [0, cHigh)
  int nFree;          // entries on the contact free stack
  int nContacts;      // alive contacts
  int nMoved;         // entries in the move list
  int nPairs;         // entries in the candidate pair buffer
  int nSolve;         // contacts handed to the solver this step
  int nColours;       // number of contact colours in use this step
  int nTouching;
  int nUncoloured;    // worklist sizes of the colouring loop
  int nUncoloured2;
  int error;          // sticky DBX_E_* raised on the device (capacity overflow, ...)
  int nTomb;          // tombstones in the pair hash
  int nIslands;
  int nAwake;
  int toiEvents;      // cumulative TOI events processed
  int nEvents;        // events selected in the current TOI pass
  int tailStart;      // first colour of the tail that k_solve runs inside one CTA
  int nCtEvents;      // contact begin/end events recorded since the last poll (may exceed evCap: the surplus is lost and reported)
  unsigned barrier;   // grid barrier ticket counter for the persistent kernels
  unsigned epoch;     // colouring round stamp (k_colour leaves the next one in epochNext; k_island_init moves it here, so that the
                      // word every CTA of k_colour reads when it starts is not written while that kernel runs)
  int nFresh;         // contacts created by the current FindNewContacts, listed in c_work for k_toi's first pass
  int maxColour;      // largest colour ever handed out (monotonic): bounds the key width of the world-major sort
  float bounds[4];    // world bounds of fat AABB centres (Morton normalisation), as ordered ints
  int colourOff[kMaxColours + 1];   // solver order: contacts of colour c are [colourOff[c], colourOff[c+1])
  int jointColourOff[kMaxJointColours + 1];
  int nToi;           // entries of c_toiList: contacts the TOI pass can ever care about this step (listed by k_collide)
  int nPostSolve;     // PostSolve records of the current step (may exceed psCap: the surplus is lost and reported)
  int nNewContacts;   // scratch counter of k_list_new_contacts
  int stepIncomplete; // sub-stepping (b2World.SetSubStepping): k_toi stopped after one solved TOI event; the next Step resumes
  int toiSolved;      // TOI mini-islands solved by the current k_toi launch (sub-stepping stops at the first)
  unsigned long long toiGlobalMin;   // sub-stepping: smallest event priority of the current pass (the ONE event to handle)
  int nTileB, nTileG; // tile solver: boundary / global constraints (contacts + joints) of this step
  unsigned long long solveStamp[4];   // %globaltimer of CTA 0 in the island solver: start, velocity passes begin, position passes begin, end (b2Profile split)
  unsigned epochNext; // see epoch
};

struct DevWorld {
  Header* hdr;
  // ---- bodies
  int nBodies;
  float4* b_xf;      // p.x p.y q.s q.c
  float4* b_xf0;     // transform at (c0, a0): what b2Body.SynchronizeFixtures recomputes as xf1 (b2body.d:1131-1133)
  float4* b_pos;     // c.x c.y a  -
  float4* b_pos0;    // c0.x c0.y a0 alpha0
  float4* b_vel;     // v.x v.y w  -
  unsigned long long* b_acc;   // [3 * nBodies] contact warm-start velocity deltas (x, y, w) in 32.32 fixed point, zero between steps
  float4* b_force;   // f.x f.y torque -
  float4* b_mass;    // invMass invI mass I
  float4* b_lc;      // localCenter.x localCenter.y linearDamping angularDamping
  float2* b_gs;      // gravityScale sleepTime
  uint32_t* b_flags;
  int* b_wake;       // wake requests raised during Collide (applied when islands are built)
  int* b_root;       // union-find parent, then island root
  int* b_islAwake;   // per root: island contains an awake seed body
  int* b_islMinSleep;// per root: min sleepTime over the island (float bits, non-negative)
  int* b_posNotOk;   // [kMaxPosIters][nBodies] per root: some constraint still violated after iteration i
  unsigned long long* b_mask;   // colours used by the touching contacts of a dynamic body
  unsigned long long* b_claim;  // colouring arbitration word
  int* b_ovf;        // overflow colour counter (bodies with more than 64 touching contacts)
  int* b_world;      // replica index (batched independent worlds)
  // TOI sub-stepping (dynamics/b2world.d:1127-1452)
  unsigned long long* b_toiMin;    // per non-static body: min (alpha bits << 32 | contact slot) over its candidate events
  unsigned long long* b_toiOther;  // per non-static body: min priority of the events that would pull it into their mini-island
  int* b_toiEvt;     // event index that owns the body in this pass (-1 none)
  int* b_toiFlags;   // bit 0: invalidate cached TOIs of its contacts; bit 1: synchronize its fixtures
  int* e_contact;    // [events] contact slot
  int* e_ncand;      // [events][2]
  int* e_cand;       // [events][2][kToiCand] contacts of bodyA / bodyB that may join the mini-island
  int eventCap;
  // ---- fixtures
  int nFixtures;
  int* f_body;
  float2* f_mat;     // friction restitution
  uint32_t* f_filter;// categoryBits | maskBits << 16
  int* f_group;      // groupIndex (low 16, signed) | flags << 16
  // ---- shapes (de-duplicated geometry pool)
  int nShapes;
  const DShape* shapes;
  // ---- proxies
  int nProxies;
  int4* p_ids;       // fixture child body shape
  int* p_key;        // reference tree-node id: (lo, hi) order of a pair decides fixture A/B (b2broadphase.d:289-290)
  float4* p_aabb;    // tight swept AABB
  float4* p_fat;     // persistent fat AABB
  uint32_t* p_flags;
  int* moveList; int moveCap;
  // ---- LBVH over the fat AABBs
  unsigned long long* bv_key; unsigned long long* bv_keyAlt;
  int* bv_leaf; int* bv_leafAlt;     // sorted leaf -> proxy slot
  float4* bv_box;    // [2n-1]: internal nodes 0..n-2, leaves n-1..2n-2
  int2* bv_child;    // [n-1]
  int* bv_parent;    // [2n-1]
  int* bv_visit;     // [n-1]
  int4* ev_a;        // [evCap] contact events: (type | phase << 8 | step << 16, fixtureA, fixtureB, childA | childB << 16)
  int4* ev_b;        //         (bodyA, bodyB, pair key lo, pair key hi)
  int evCap;         // 0 = contact events off
  int4* ps_a;        // [psCap] PostSolve records: (fixtureA, fixtureB, childA | childB << 16, count | phase << 8)
  float4* ps_b;      //         (normalImpulse0, tangentImpulse0, normalImpulse1, tangentImpulse1)
  unsigned long long* ps_key;   //  pair key
  int psCap;         // 0 = PostSolve recording off
  int userFilter;    // DBX_FILTER_* (user b2ContactFilter, deferred): bit 0 tag new contacts, bit 1 skip the default filter
  // world-local solve (batched replicas without joints): solver slots sorted by (replica, colour) instead of colour alone
  const unsigned* sw_key;   // [nSolve] sorted (replica << swColourBits | colour)
  int swColourBits;
  int* w_start;             // [nWorlds] first solver slot of a replica
  int* w_end;               // [nWorlds] one past its last
  int toiReset;      // k_toi: the per-body TOI scratch may be dirty (first step, bodies added) -> full reset phase
  int toiClearMoves; // k_toi also empties the move buffer FindNewContacts left (saves two launches)
  int toiClearForces;// k_toi also runs ClearForces (b2world.d:443-450) in its final body pass
  int toiMode;       // k_toi: 1 = only the first evaluation, launched as a plain kernel on the second stream (no grid barrier is reached)
  int toiPre;        // k_toi: k_toi_pre already did the first evaluation of the contacts that existed before FindNewContacts
  int subStep;       // b2World.SetSubStepping(true): handle ONE TOI event per Step (dynamics/b2world.d:1441-1446)
  int toiResume;     // the previous Step left SolveTOI unfinished (m_stepComplete == false): no reset of the TOI state (:1131-1146)
  int stepIndex;     // low 16 bits stamp the events of this step
  int2* bv_wr;       // [n-1] replica-index range of the leaves under an internal node
  int* bv_pos;       // proxy slot -> sorted leaf index
  const int* bv_sorted;  // sorted leaf index -> proxy slot (whichever CUB buffer is current)
  // ---- candidate pairs
  int2* pairs; int pairCap;   // proxy slots (lo, hi) in reference key order
  // ---- joints with collideConnected == false, as sorted (bodyLo << 32 | bodyHi)
  int nJointPairs; const unsigned long long* jp_keys;
  const unsigned long long* b_jmask;   // per body: all colours up to its highest joint colour (contacts there must stay above)
  int unifiedColours;                  // joint colour c and contact colour c share a solver phase (see k_solve)
  const uint32_t* jp_bits;   // one bit per body: set if some joint forbids a collision of that body (skips the search above)
  // ---- contacts
  int cCap;
  unsigned long long* c_key;   // (refKeyLo << 32 | refKeyHi)
  int4* c_ids;       // proxyA proxyB bodyA bodyB   (A/B after the type-registry swap, b2contact.d:387-394)
  int4* c_fix;       // fixtureA fixtureB shapeA shapeB
  uint32_t* c_flags;
  float4* c_m0;      // localNormal.xy localPoint.xy
  float4* c_m1;      // points[0].localPoint points[1].localPoint
  float4* c_imp;     // n0 t0 n1 t1
  uint4* c_mk;       // key0 key1 type pointCount
  float4* c_mat;     // friction restitution tangentSpeed toi
  int* c_toiList;    // [cCap] see Header::nToi
  int* c_toiCount;
  int* c_colour;
  int* c_free;
  int* c_work; int* c_work2;   // colouring worklists
  // pair hash
  int hCap;          // power of two
  unsigned long long* h_key; int* h_val;
  // ---- solver (colour-sorted constraints)
  int sCap;
  int* s_contact;    // solver index -> contact slot
  int* s_hist;       // [kSortBlocks][kMaxColours]
  int2* s_body;      // bodyA bodyB
  float4* s_v0;      // normal.xy friction tangentSpeed
  float4* s_v1;      // invMassA invIA invMassB invIB
  float4* s_r0;      // point 0: rA.xy rB.xy
  float4* s_r1;
  float4* s_q0;      // point 0: normalMass tangentMass velocityBias -
  float4* s_q1;
  float4* s_imp;     // n0 t0 n1 t1 (working impulses)
  float4* s_nm;      // normalMass 2x2: ex.x ex.y ey.x ey.y
  float4* s_k;       // K 2x2
  int* s_pc;         // pointCount (possibly reduced to 1 by the block solver) | type << 8 | manifold pointCount << 16
  float4* s_p0;      // localPoints[0].xy localPoints[1].xy
  float4* s_p1;      // localNormal.xy localPoint.xy
  float4* s_p2;      // localCenterA.xy localCenterB.xy
  float2* s_p3;      // radiusA radiusB
  int* s_root;       // island root of the constraint
  // ---- joints
  int nJoints;
  int4* j_ids;       // type bodyA bodyB flags(collideConnected | enableLimit<<1 | enableMotor<<2 | alive<<3)
  float4* j_anchor;  // localAnchorA.xy localAnchorB.xy
  float4* j_p0;      // revolute: referenceAngle lowerAngle upperAngle maxMotorTorque | distance: length frequencyHz dampingRatio -
  float4* j_p1;      // revolute: motorSpeed - - -
  float4* j_imp;     // impulse.xyz motorImpulse   (distance: impulse - - -)
  int* j_limit;      // limit state
  int* j_colour;
  int* j_order;      // colour-sorted joint indices
  int* j_root;
  // per-step temporaries (b2revolutejoint.d:654-666, b2distancejoint.d:387-398)
  float4* j_r;       // rA.xy rB.xy
  float4* j_lc;      // localCenterA.xy localCenterB.xy
  float4* j_m;       // invMassA invIA invMassB invIB
  float4* j_k0;      // revolute mass.ex.xyz motorMass | distance: u.x u.y mass gamma
  float4* j_k1;      // revolute mass.ey.xyz -        | distance: bias - - -
  float4* j_k2;      // revolute mass.ez.xyz -
  float4* j_k3;      // prismatic / wheel / gear (dbx_joints2.cuh)
  float4* j_p2;      // gear: referenceAngleA referenceAngleB ratio constant
  int4* j_ids2;      // gear: bodyC bodyD typeA typeB
  // ---- step parameters
  float dt, inv_dt, dtRatio; int velIters, posIters; int warmStarting; int allowSleep; int continuous; float gx, gy;
  int nWorlds;
  int keyStride;        // reference-key stride between replicas (0 when not replicated)
  unsigned long long* phaseTimes; int phaseCap;   // debug: globaltimer stamp after every barrier of k_solve (null = off)
  int dbgFlags;         // experiments only (DBX_DEBUG env): bit 0 = joint velocity phases do no work
  int nJointColours;    // joint colours in use (host-side greedy colouring)
  int jointBlocks;      // CTAs of the persistent solver dedicated to joint phases
  int colourOverride;   // debug: keep caller-supplied contact levels instead of colouring (dbx_world_debug_set_contact_levels)
  // ---- tile solver (dbx_tiles.cu): dynamic bodies in x order cut into nTiles tiles of tileBodies; constraints by (class, tile, colour)
  int tiled;            // this step's rows / joints carry body references (dbx_solver.cuh, BodyView) and k_solve_tiles runs them
  int nTiles, tileBodies, nTileBodies;
  int tileKinematic;    // the world holds kinematic bodies (they are integrated outside the tiles)
  int* b_tslot;         // body -> position in tile order (-1: not a dynamic body)
  int* t_body;          // position in tile order -> body
  int* b_tclaim;        // per body: lowest boundary straddled by one of its constraints (0x7fffffff at rest)
  int* b_xflag;         // per TILE SLOT: XF_* (0 at rest)
  float2* t_mass;       // per tile slot: inverse mass, inverse inertia (refreshed by k_mark_solve every step)
  int* c_tkey; int2* c_bref;     // per contact slot: bin and body references of this step
  int* c_tcol; int* j_tcol;      // boundary constraints: the local colour their tile gave them (k_solve_tiles), -1 otherwise
  int* j_tkey; int2* j_bref;     // per joint slot likewise
  int* t_off; int* t_cur;        // [bins + 1] solver-slot offsets per bin, [bins] histogram / scatter cursors
  int* tj_off; int* tj_cur; int* tj_order;   // joints: offsets, cursors, joint slots in bin order
  int* t_flag;          // [2 nTiles] k_solve_tiles' neighbour handshakes: passes published / boundary passes done, per tile
};

}  // namespace dbx
