// dbx_world.h — host side of the device world: object tables, the reference-compatible id allocators, lazy
// host<->device mirroring and the per-step launch sequence.  Host code here is setup/boundary work only
// (b2World.CreateBody / CreateFixture / CreateJoint, mass data, proxy-id order); the step itself never touches it.
#pragma once
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include "../../include/dbox_b200.h"
#include "dbx_kernels.cuh"

namespace dbx {

template <class T> struct DevBuf {
  T* p = nullptr; size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }           // (World::~World releases most pools by name; whatever it does not list goes here)
  cudaError_t reserve(size_t n, bool keep, cudaStream_t st) {
    if (n <= cap) return cudaSuccess;
    size_t ncap = cap ? cap : 64;
    while (ncap < n) ncap *= 2;
    if (n > (1u << 22)) ncap = n;   // very large pools (replicated worlds): no power-of-two slack
    T* q = nullptr;
    cudaError_t e = cudaMalloc((void**)&q, ncap * sizeof(T));
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(q, 0, ncap * sizeof(T), st);
    if (e != cudaSuccess) return e;
    if (keep && p && cap) { e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st); if (e != cudaSuccess) return e; }
    if (p) { cudaStreamSynchronize(st); cudaFree(p); }
    p = q; cap = ncap;
    return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HShape {            // host copy of a fixture's shape (b2fixture.d:380 clones it)
  dbx_shape s;
  std::vector<dbx_vec2> chain;
};
struct HFixture {
  bool alive = false; int body = -1; dbx_fixture_def def{}; HShape shape; std::vector<int> proxies;
};
struct HProxy {
  bool alive = false; int fixture = -1, child = 0, body = -1, shape = -1, key = -1; uint32_t flags = 0; float4 aabb{}, fat{};
};
struct HBody {
  bool alive = false; dbx_body_state st{}; float4 xf0{}; std::vector<int> fixtures, joints; int world = 0;
  unsigned long long validEpoch = 0;   // host row read back at this World::bodyEpoch_ (row-granular sync: World::mutBodyRow)
  bool dirty = false;                  // host row newer than the device's (listed in World::dirtyBodies_)
};
struct HJoint {
  bool alive = false; dbx_joint_def def{}; float imp[4] = {0, 0, 0, 0}; int limit = 0; int colour = -1;
  int bodyC = -1, bodyD = -1, typeA = 0, typeB = 0;   // gear only: the far ends and kinds of joint1 / joint2
};

class World {
 public:
  World(float gx, float gy, int device, const dbx_caps* caps);
  ~World();
  bool ok() const { return ok_; }

  int createBody(const dbx_body_def& d);
  int destroyBody(int b);
  int createFixture(int b, const dbx_fixture_def& d, const dbx_shape& s);
  int destroyFixture(int f);
  int createJoint(const dbx_joint_def& d);
  int createGearJoint(const dbx_joint_def& d);
  int destroyJoint(int j);
  int setJointTarget(int j, float x, float y);
  int setJointParams(int j, const dbx_joint_def& d, uint32_t mask);
  int setMotorSpeeds(const int32_t* joints, const float* speeds, int n);
  int step(float dt, int vi, int pi, int n);
  int enqueueStep(float dt, int vi, int pi, bool fineEvents, int halves = 3);
  int stepBegin(float dt, int vi, int pi);
  int stepEnd();
  int patchContacts(const dbx_contact_patch* in, int n);
  int timeSteps(float dt, int vi, int pi, int n, bool flushL2, float* totalMs, float* stageMs);
  int applyForces(const float* f4, int n);
  int setBodyStates(const int* ids, const float* pose4, const float* vel4, int n);
  int rayCastClosest(const dbx_ray* rays, int n, dbx_ray_hit* out);
  int queryAabb(const dbx_aabb* boxes, int n, int capPer, int32_t* counts, int32_t* fixtureChild);
  int refreshTreeForQuery();
  int treeStats(int32_t* height, int32_t* maxBalance, float* quality);
  int rayCastAll(const dbx_ray* rays, int n, int capPer, int32_t* counts, dbx_ray_hit* hits);
  int testPoints(const int32_t* fixtures, const dbx_vec2* points, int n, int32_t* inside);
  int shiftOrigin(float x, float y);
  int readWorldManifolds(dbx_world_manifold* out, int cap);
  int enablePostSolve(int capacity);
  int readPostSolve(dbx_post_solve* out, int cap);
  int setUserFilter(int mode);
  int pollNewContacts(int32_t* out, int cap);
  int enableContactEvents(int capacity);
  int pollContactEvents(dbx_contact_event* out, int cap);
  int readTransforms(float* out, int n);
  int stepAsync(float dt, int vi, int pi);
  int applyForcesAsync(const float* f4, int n);
  int readTransformsAsync(float* out, int n);
  int ioWait(int ticket);
  int sync();
  long launchCount() const { return L_.launches; }
  int clearForces();
  int setFlags(uint32_t f);
  uint32_t flags() const { return flags_; }
  int setGravity(float gx, float gy) { gx_ = gx; gy_ = gy; return 0; }

  int getBody(int b, dbx_body_state* out);
  int readBodiesDevice(int from, int count, dbx_body_state* out);
  HBody* mutBody(int b);
  // pulls, marks dirty; nullptr if invalid
  HBody* mutBodyRow(int b);            // as mutBody, for edits that touch this body's row only
  int setTransform(int b, float x, float y, float angle);
  int setBodyType(int b, int type);
  int setBodyActive(int b, bool flag);
  int setMassData(int b, float mass, float cx, float cy, float I);
  int resetMass(int b);
  int setFixedRotation(int b, bool flag);
  int setBodyScalars(int b, const float* linearDamping, const float* angularDamping, const float* gravityScale);
  int setFixtureFilter(int f, int category, int mask, int group);
  int setFixtureSensor(int f, bool flag);
  int setFixtureMaterial(int f, const float* friction, const float* restitution, const float* density);
  int createProxiesFor(int fid);
  void wake(HBody& hb, bool flag);

  int counts(dbx_counts* out);
  int profile(dbx_profile* out);
  int readBodies(dbx_body_state* out, int cap);
  int writeBodies(const dbx_body_state* in, int n);
  int readContacts(dbx_contact_rec* out, int cap);
  int writeContacts(const dbx_contact_rec* in, int n);
  int readContactColours(int32_t* out, int cap);      // in the order of the last readContacts
  int writeContactColours(const int32_t* in, int n);  // for the records of the last writeContacts
  int readProxies(dbx_proxy_rec* out, int cap);
  int writeProxies(const dbx_proxy_rec* in, int n);
  int readJoints(dbx_joint_state* out, int cap);
  int writeJoints(const dbx_joint_state* in, int n);
  int readMoves(int32_t* out, int cap);
  int writeMoves(const int32_t* in, int n);
  int readPairs(int32_t* out, int cap);
  int stageFindNewContacts();
  int stageCollide();
  int setIoFormat(int format) { if (format != 0 && format != 1) return DBX_E_INVALID; ioCompact_ = format == 1; return 0; }
  int setContactLevels(const int32_t* levels, int n);
  // snapshot / restore: the tile solver's body-to-tile assignment is not part of a snapshot; both sides of an export / import
  // re-derive it from the body positions at that moment, so a restored world keeps stepping bit for bit like the original
  void resetSolverSchedule() { tilesValid_ = false; sinceTileSort_ = 0; }
  int readSolveOrder(int32_t* contactRank, int capC, int32_t* jointRank, int capJ, int32_t* info4);
  int colourConflicts();
  int readHeader(void* out, int bytes) { cudaStreamSynchronize(stream_); int n = bytes < (int)sizeof(Header) ? bytes : (int)sizeof(Header); return cudaMemcpy(out, hdr_.p, n, cudaMemcpyDeviceToHost) == cudaSuccess ? n : DBX_E_CUDA; }
  int phaseTimes(unsigned long long* out, int cap);   // debug: enable + fetch the last step's k_solve barrier stamps
  int replicate(int copies);
  int replicaCount() const { return nWorlds_; }
  float inv_dt0 = 0.0f;

 private:
  int fail(cudaError_t e, const char* where);
  int push();              // upload pending host-side changes, (re)size device pools
  int pullBodies();
  int pullProxies();
  int pullJoints();
  int checkDeviceError(bool sync);
  void refreshView();
  int allocProxyKey(); void freeProxyKey(int key);
  int internShape(const DShape& s);
  void buildChildShape(const HShape& hs, int child, DShape* out) const;
  void computeMass(const HShape& hs, float density, float* mass, dbx_vec2* center, float* I) const;
  void resetMassData(HBody& hb);
  void hostAabb(const DShape& s, const Xf& xf, Box* out) const;
  int destroyContactsWhere(int body, int fixture, int otherBody, bool flagOnly);
  int recolourJoints();
  static float4 jointParams(const dbx_joint_def& d, int which);
  int reserveDevice(bool& rehash);
  int findNewContacts(bool deferClear = false);
  int compactContacts();
  void setStepParams(float dt, int vi, int pi);

  bool ok_ = false;
  int device_ = 0;
  cudaStream_t stream_ = 0;
  LaunchCfg L_;
  float gx_, gy_;
  uint32_t flags_ = DBX_WORLD_DEFAULT_FLAGS;
  bool newFixture_ = false, stepComplete_ = true;
  long stepCount_ = 0;
  int nWorlds_ = 1; bool replicated_ = false; int keyStride_ = 0;
  dbx_caps caps_{};

  // host tables
  std::vector<HBody> bodies_; std::vector<HFixture> fixtures_; std::vector<HProxy> proxies_; std::vector<HJoint> joints_;
  std::vector<DShape> shapes_; std::unordered_map<std::string, int> shapeIndex_;
  std::vector<int> proxyFree_;
  std::vector<int> jointAt_, jointPos_;   // device joint slot <-> joint id (device arrays are colour-sorted)
  // reference leaf-id allocator (collision/b2dynamictree.d:516-564 replayed)
  std::vector<int> keyFree_; int keyFresh_ = 0; int keyLeaves_ = 0;
  // mirror state
  size_t bodiesSynced_ = 0, fixturesSynced_ = 0, proxiesSynced_ = 0, shapesSynced_ = 0, jointsSynced_ = 0;
  bool hostBodiesValid_ = true, hostProxiesValid_ = true, hostJointsValid_ = true;
  // Row-granular body sync for the per-body calls a game makes every frame (b2Body.ApplyForce, SetLinearVelocity, SetAwake ...):
  // one body's row is read back and written instead of every array of the world.  bodyEpoch_ counts the device-side changes;
  // a host row is current when the whole mirror is (hostBodiesValid_) or when it was read at this epoch or edited since.
  unsigned long long bodyEpoch_ = 1, rowPullEpoch_ = 0; int rowPulls_ = 0;
  std::vector<int> dirtyBodies_;
  int pullBodyRow(int b);
  int pushBodyRows();
  DevBuf<float4> rowStage_; DevBuf<int> rowIds_;
  bool fullPushBodies_ = false, fullPushProxies_ = false, fullPushJoints_ = false, fullPushFixtures_ = false;
  bool jointsChanged_ = false;
  std::vector<int> movesOnDevice_; std::unordered_set<int> movesUploaded_;   // host moves already in the device move list since the last FindNewContacts
  std::vector<int> pendingMoves_;   // proxies buffered on the host since the last push (b2broadphase.d:244-257)

  // device
  DevWorld dw_{};
  DevBuf<Header> hdr_;
  DevBuf<float4> b_xf, b_xf0, b_pos, b_pos0, b_vel, b_force, b_mass, b_lc; DevBuf<float2> b_gs; DevBuf<uint32_t> b_flags;
  DevBuf<unsigned long long> b_toiMin, b_toiOther, b_acc; DevBuf<int> b_toiEvt, b_toiFlags, e_contact, e_ncand, e_cand, bv_pos;
  DevBuf<int> b_wake, b_root, b_islAwake, b_islMinSleep, b_posNotOk, b_ovf, b_world; DevBuf<unsigned long long> b_mask, b_claim;
  DevBuf<int> f_body, f_group; DevBuf<float2> f_mat; DevBuf<uint32_t> f_filter;
  DevBuf<DShape> d_shapes;
  DevBuf<int4> p_ids; DevBuf<int> p_key, moveList; DevBuf<float4> p_aabb, p_fat; DevBuf<uint32_t> p_flags;
  DevBuf<unsigned long long> bv_key, bv_keyAlt; DevBuf<int> bv_leaf, bv_leafAlt, bv_parent, bv_visit; DevBuf<float4> bv_box; DevBuf<int2> bv_child, bv_wr;
  DevBuf<int2> pairs; DevBuf<unsigned long long> jp_keys; DevBuf<uint32_t> jp_bits;
  int uploadJointBits(const std::vector<unsigned long long>& keys);
  std::vector<unsigned long long> jpHost_; size_t jpBitsBodies_ = 0;
  DevBuf<unsigned long long> b_jmask; std::vector<unsigned long long> jmaskHost_; size_t jmaskBodies_ = 0;
  int uploadJointMasks();
  DevBuf<unsigned long long> c_key, h_key; DevBuf<int4> c_ids, c_fix; DevBuf<uint32_t> c_flags; DevBuf<float4> c_m0, c_m1, c_imp, c_mat; DevBuf<uint4> c_mk;
  DevBuf<int> c_toiList, c_toiCount, c_colour, c_free, c_work, c_work2, h_val;
  DevBuf<int> s_contact, s_hist, s_pc, s_root; DevBuf<int2> s_body; DevBuf<float4> s_v0, s_v1, s_r0, s_r1, s_q0, s_q1, s_imp, s_nm, s_k, s_p0, s_p1, s_p2; DevBuf<float2> s_p3;
  DevBuf<int4> j_ids; DevBuf<float4> j_anchor, j_p0, j_p1, j_imp, j_r, j_lc, j_m, j_k0, j_k1, j_k2, j_k3, j_p2; DevBuf<int4> j_ids2; DevBuf<int> j_limit, j_colour, j_order, j_root;
  DevBuf<char> cubTemp; DevBuf<int> d_levels;
  DevBuf<unsigned long long> cmpKeyA_, cmpKeyB_; DevBuf<int> cmpValA_, cmpValB_;
  int nJointPairs_ = 0; int jointBlocks_ = 0, nJointColours_ = 0;
  cudaEvent_t ev_[10]{};
  bool evValid_ = false, evFine_ = false;
  DevBuf<char> flushBuf_; DevBuf<float4> ioBuf_, ioBuf2_; DevBuf<int> ioIds_; DevBuf<unsigned> swKeyA_, swKeyB_; DevBuf<int> swValA_, swValB_, wStart_, wEnd_;
  DevBuf<int4> ev_a_, ev_b_; DevBuf<unsigned long long> patchKeys_; bool midStep_ = false; float stepDt_ = 0; int stepVi_ = 0, stepPi_ = 0; DevBuf<float4> qIn_, qOut_; DevBuf<int> qCount_; DevBuf<int2> qPairs_; DevBuf<unsigned long long> phaseBuf_;
  DevBuf<int4> ps_a_; DevBuf<float4> ps_b_; DevBuf<unsigned long long> ps_key_;
  // pipelined I/O: copy streams beside the step, double-buffered staging on both sides
  int ensureIoStreams();
  cudaStream_t h2d_ = nullptr, d2h_ = nullptr;
  DevBuf<float4> inStage_[2], outSnap_[2];
  cudaEvent_t inCopied_[2]{}, inRead_[2]{}, snapReady_[2]{}, outDone_[2]{};
  bool inReadValid_[2] = {false, false}, outDoneValid_[2] = {false, false};
  int inFlip_ = 0, ioTicket_ = 0;
  bool overrideLevels_ = false;
  bool lastUnified_ = false;   // the last step ran joint colour c and contact colour c in one phase
  bool treeValid_ = false; int sinceRebuild_ = 0;
  // contact-pool watermark: every 8th step the header is copied to pinned memory without waiting; a later step looks at
  // the copy that has landed and doubles the pool before it can overflow
  int* wm_ = nullptr; cudaEvent_t wmEv_ = nullptr; bool wmPending_ = false; size_t contactFloor_ = 0;
  int growContactsIfNeeded();
  bool ioCompact_ = false; size_t ioRecordBytes() const { return ioCompact_ ? 12 : 16; }
  int seenMaxColour_ = 0;
  // second stream for the overlapped TOI pre-evaluation (fork after the solver, join before k_toi)
  cudaStream_t aux_ = nullptr; cudaEvent_t evFork_ = nullptr, evJoin_ = nullptr; bool toiClean_ = false; size_t toiBodies_ = 0;
  std::vector<int> lastReadSlots_;
  // tile solver (dbx_tiles.cu): dynamic bodies of a big single world in x order, cut into one tile per CTA
  int prepareTiles();            // > 0: this step runs the tile solver (buffers sized, tiles assigned, dw_ filled in)
  DevBuf<float2> t_mass_;
  DevBuf<int> b_tslot_, t_body_, b_tclaim_, b_xflag_, c_tkey_, j_tkey_, c_tcol_, j_tcol_, t_flag_, t_off_, t_cur_, tj_off_, tj_cur_, tj_order_, tValA_, tValB_;
  DevBuf<int2> c_bref_, j_bref_; DevBuf<unsigned> tKeyA_, tKeyB_;
  bool tilesDirty_ = true, tilesValid_ = false, lastTiled_ = false, lastWorldsPath_ = false; int sinceTileSort_ = 0, nDynamic_ = 0; size_t tileBodyCap_ = 0;
};

void set_last_error(const std::string& s);
const char* get_last_error();

}  // namespace dbx
