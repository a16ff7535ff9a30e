// dbx_kernels.cuh — the step pipeline as CUDA kernels for sm_100a (declared here, defined in dbx_kernels.cu).
#pragma once
#include "dbx_device.cuh"

namespace dbx {

struct LaunchCfg {
  int sms = 148;
  int gridWide = 148 * 8;   // grid-stride kernels: multiple of the SM count
  int coopBlocks = 148;     // persistent cooperative kernels: one CTA per SM
  int coopThreads = 512;
  cudaStream_t stream = 0;
  void* cubTemp = nullptr; size_t cubTempBytes = 0;
  mutable long coopLaunches = 0;
  mutable long launches = 0;   // kernels of this library launched so far (CUB's radix-sort kernels are not counted)
};

// stages of b2World.Step (dynamics/b2world.d:367-434); each returns the first CUDA error
cudaError_t stage_collide(const DevWorld& W, const LaunchCfg& L);                 // b2ContactManager.Collide
cudaError_t stage_islands_and_integrate(const DevWorld& W, const LaunchCfg& L);   // island discovery + wake + integrate velocities
cudaError_t stage_colour_and_sort(const DevWorld& W, const LaunchCfg& L);
cudaError_t stage_colour_and_sort_worlds(DevWorld& W, const LaunchCfg& L, unsigned* keysA, unsigned* keysB, int* valsA, int* valsB, int n, int colourBits);
cudaError_t launch_mark_and_colour(const DevWorld& W, const LaunchCfg& L);
cudaError_t stage_solve_worlds(const DevWorld& W, const LaunchCfg& L, int bodiesPerWorld);
size_t cub_temp_bytes_u32(int n);         // graph colouring + colour counting sort
cudaError_t stage_prepare(const DevWorld& W, const LaunchCfg& L);                 // contact constraint setup
cudaError_t stage_solve(const DevWorld& W, const LaunchCfg& L);                   // warm start + iterations + finalize + sleep (one persistent kernel)
// tile solver (dbx_tiles.cu)
cudaError_t stage_tile_assign(const DevWorld& W, const LaunchCfg& L, unsigned* keysA, unsigned* keysB, int* valsA, int* valsB);
cudaError_t stage_colour_and_sort_tiles(const DevWorld& W, const LaunchCfg& L);
cudaError_t stage_solve_tiles(const DevWorld& W, const LaunchCfg& L);
size_t tile_smem_bytes(int tileBodies);
cudaError_t stage_sync_fixtures(const DevWorld& W, const LaunchCfg& L);           // b2Body.SynchronizeFixtures / MoveProxy
cudaError_t stage_find_new_contacts(DevWorld& W, const LaunchCfg& L, bool rebuild, bool deferClear);       // LBVH rebuild + pair query + AddPair
cudaError_t launch_api_resensor(const DevWorld& W, const LaunchCfg& L, int fixture);
cudaError_t launch_patch_contacts(const DevWorld& W, const LaunchCfg& L, const unsigned long long* keys, const float4* vals, const int* masks, int n);
cudaError_t stage_refresh_tree(DevWorld& W, const LaunchCfg& L, bool rebuild);
cudaError_t launch_raycast(const DevWorld& W, const LaunchCfg& L, const float4* rays, int n, float4* out);
cudaError_t launch_query_aabb(const DevWorld& W, const LaunchCfg& L, const float4* boxes, int n, int capPer, int* counts, int2* out);
cudaError_t launch_raycast_all(const DevWorld& W, const LaunchCfg& L, const float4* rays, int n, int capPer, int* counts, float4* out);
cudaError_t launch_test_points(const DevWorld& W, const LaunchCfg& L, const float4* q, int n, int* inside);
cudaError_t launch_shift_origin(const DevWorld& W, const LaunchCfg& L, float ox, float oy);
cudaError_t launch_world_manifolds(const DevWorld& W, const LaunchCfg& L, int high, float4* out);
cudaError_t launch_post_solve(const DevWorld& W, const LaunchCfg& L);
cudaError_t launch_list_new_contacts(const DevWorld& W, const LaunchCfg& L, int4* out, unsigned long long* keys, int cap, int peek);
cudaError_t stage_toi(DevWorld& W, const LaunchCfg& L);
cudaError_t stage_toi_pre(const DevWorld& W, const LaunchCfg& L, cudaStream_t aux);                               // b2World.SolveTOI
cudaError_t stage_rebuild_hash(const DevWorld& W, const LaunchCfg& L);
cudaError_t stage_compact_contacts(DevWorld& W, const LaunchCfg& L, int high, int nAlive, void* scratch, unsigned long long* keyA, unsigned long long* keyB, int* valA, int* valB);
cudaError_t stage_count(const DevWorld& W, const LaunchCfg& L);                   // refresh hdr->nContacts / nTouching / nAwake
size_t cub_temp_bytes(int maxProxies);

// helpers used by the state import path
cudaError_t launch_insert_contacts(const DevWorld& W, const LaunchCfg& L, int n);  // (re)build hash + free list for slots [0,n)
cudaError_t launch_api_contacts(const DevWorld& W, const LaunchCfg& L, int body, int fixture, int otherBody, int flagOnly);
cudaError_t launch_api_wake(const DevWorld& W, const LaunchCfg& L, int a, int b);
cudaError_t launch_body_rows(const DevWorld& W, const LaunchCfg& L, const int* ids, float4* rows, int n, bool set);   // nine float4 per row
cudaError_t launch_set_motor_speeds(const DevWorld& W, const LaunchCfg& L, const int* slots, const float* speeds, int n);
cudaError_t launch_set_states(const DevWorld& W, const LaunchCfg& L, const int* ids, const float4* pose, const float4* vel, int n);
cudaError_t launch_apply_forces(const DevWorld& W, const LaunchCfg& L, const float4* forces, int n);
cudaError_t launch_apply_forces3(const DevWorld& W, const LaunchCfg& L, const float* forces, int n);   // DBX_IO_COMPACT
cudaError_t launch_pack_poses(const DevWorld& W, const LaunchCfg& L, float* out, int n);
cudaError_t launch_clear_forces(const DevWorld& W, const LaunchCfg& L);
cudaError_t launch_replicate(const DevWorld& W, const LaunchCfg& L, int nB, int nF, int nP, int nMoved, int keyStride, int copies);
cudaError_t launch_set_levels(const DevWorld& W, const LaunchCfg& L, const int* d_levels, int n);

}  // namespace dbx
