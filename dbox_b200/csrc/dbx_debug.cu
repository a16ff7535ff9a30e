// dbx_debug.cu — parity hooks: run the device narrowphase / GJK / TOI functions on caller-supplied inputs, one thread per
// item.  They exist so tests can compare the CUDA arithmetic with the CPU oracle on arbitrary (random) inputs without a
// world around them; the step pipeline calls the same __device__ functions (dbx_narrow.cuh).
#include <cstring>
#include <vector>
#include "../../include/dbox_b200.h"
#include "dbx_device.cuh"
#include "dbx_world.h"

using namespace dbx;

static bool to_dshape(const dbx_shape& s, DShape* out) {
  std::memset(out, 0, sizeof(DShape));
  out->radius = s.radius;
  switch (s.type) {
    case DBX_SHAPE_CIRCLE: out->type = SH_CIRCLE; out->c = V(s.p.x, s.p.y); return true;
    case DBX_SHAPE_EDGE:
      out->type = SH_EDGE;
      out->v[0] = V(s.v0.x, s.v0.y); out->v[1] = V(s.v1.x, s.v1.y); out->v[2] = V(s.v2.x, s.v2.y); out->v[3] = V(s.v3.x, s.v3.y);
      out->flags = (s.hasV0 ? SHF_HAS_V0 : 0) | (s.hasV3 ? SHF_HAS_V3 : 0);
      return true;
    case DBX_SHAPE_POLYGON:
      out->type = SH_POLYGON; out->count = s.count; out->c = V(s.centroid.x, s.centroid.y);
      for (int i = 0; i < s.count && i < 8; ++i) { out->v[i] = V(s.vertices[i].x, s.vertices[i].y); out->n[i] = V(s.normals[i].x, s.normals[i].y); }
      return true;
    default: return false;   // chains: pass the child edge
  }
}

__global__ void k_dbg_collide(int n, const DShape* A, const float4* xfA, const DShape* B, const float4* xfB, dbx_manifold* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Manifold m;
  m.type = 0; m.pointCount = 0; m.localNormal = V(0, 0); m.localPoint = V(0, 0); m.lp[0] = m.lp[1] = V(0, 0); m.key[0] = m.key[1] = 0;
  collide_dispatch(m, A + i, XF(xfA[i]), B + i, XF(xfB[i]));
  dbx_manifold o;
  memset(&o, 0, sizeof(o));
  o.type = m.type; o.pointCount = m.pointCount;
  o.localNormal.x = m.localNormal.x; o.localNormal.y = m.localNormal.y; o.localPoint.x = m.localPoint.x; o.localPoint.y = m.localPoint.y;
  for (int k = 0; k < m.pointCount; ++k) { o.points[k].localPoint.x = m.lp[k].x; o.points[k].localPoint.y = m.lp[k].y; o.points[k].key = m.key[k]; }
  out[i] = o;
}
__global__ void k_dbg_distance(int n, const DShape* A, const float4* xfA, const DShape* B, const float4* xfB, int useRadii, float* dist, float2* pA, float2* pB, int* iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  DProxy a = make_proxy(A + i), b = make_proxy(B + i);
  SimplexCache cache; cache.count = 0; cache.metric = 0.0f;
  DistanceOutput o;
  gjk_distance(o, cache, a, XF(xfA[i]), b, XF(xfB[i]), useRadii != 0);
  dist[i] = o.distance; pA[i] = o.pointA; pB[i] = o.pointB; iters[i] = o.iterations;
}
__global__ void k_dbg_toi(int n, const DShape* A, const float* swA, const DShape* B, const float* swB, float tMax, int* state, float* t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto rd = [](const float* s) { Sweep w; w.localCenter = V(s[0], s[1]); w.c0 = V(s[2], s[3]); w.c = V(s[4], s[5]); w.a0 = s[6]; w.a = s[7]; w.alpha0 = s[8]; return w; };
  DProxy a = make_proxy(A + i), b = make_proxy(B + i);
  float tt;
  state[i] = time_of_impact(&tt, a, rd(swA + 9 * i), b, rd(swB + 9 * i), tMax);
  t[i] = tt;
}

namespace {
template <class T> struct Tmp {
  T* p = nullptr;
  cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T)); }
  ~Tmp() { if (p) cudaFree(p); }
};
int shapes_up(int n, const dbx_shape* a, const dbx_shape* b, Tmp<DShape>& dA, Tmp<DShape>& dB) {
  std::vector<DShape> ha(n), hb(n);
  for (int i = 0; i < n; ++i) if (!to_dshape(a[i], &ha[i]) || !to_dshape(b[i], &hb[i])) return DBX_E_INVALID;
  if (dA.alloc(n) != cudaSuccess || dB.alloc(n) != cudaSuccess) return DBX_E_CUDA;
  cudaMemcpy(dA.p, ha.data(), (size_t)n * sizeof(DShape), cudaMemcpyHostToDevice);
  cudaMemcpy(dB.p, hb.data(), (size_t)n * sizeof(DShape), cudaMemcpyHostToDevice);
  return 0;
}
int dev_ok(int device) {
  int nd = 0;
  if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) { set_last_error("no CUDA device: dbox_b200 has no CPU fallback"); return DBX_E_NO_DEVICE; }
  if (device < 0 || device >= nd || cudaSetDevice(device) != cudaSuccess) return DBX_E_INVALID;
  return 0;
}
int finish(const char* what) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_last_error(std::string(what) + ": " + cudaGetErrorString(e)); return DBX_E_CUDA; }
  return 0;
}
}  // namespace

__global__ void k_dbg_barrier(unsigned* counter, int iters, int mode, float4* scratch) {
  const unsigned nb = gridDim.x;
  unsigned* flag = counter + 64;   // separate 128 B line
  unsigned gen = 0;
  for (int i = 0; i < iters; ++i) {
    if (scratch) __stcg(&scratch[blockIdx.x * blockDim.x + threadIdx.x], make_float4(i, 0, 0, 0));   // stores in flight, like the solver
    __syncthreads();
    if (threadIdx.x == 0) {
      if (mode == 0) {
        __threadfence();
        unsigned ticket = atomicAdd(counter, 1u);
        unsigned target = (ticket / nb + 1u) * nb;
        while (*((volatile unsigned*)counter) < target) { }
        __threadfence();
      } else if (mode == 1) {
        ++gen;
        __threadfence();
        unsigned ticket = atomicAdd(counter, 1u);
        if (ticket == gen * nb - 1u) { *((volatile unsigned*)flag) = gen; }
        else { while (*((volatile unsigned*)flag) < gen) { } }
        __threadfence();
      } else {
        ++gen;
        unsigned ticket;
        asm volatile("atom.add.release.gpu.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(counter) : "memory");
        if (ticket == gen * nb - 1u) { asm volatile("st.release.gpu.u32 [%0], %1;" :: "l"(flag), "r"(gen) : "memory"); }
        else { unsigned v; do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while (v < gen); }
      }
    }
    __syncthreads();
  }
}

extern "C" {

/* microbenchmark: average microseconds per grid barrier of `blocks` co-resident CTAs (design input for the persistent solver) */
float dbx_debug_barrier_us(int32_t device, int32_t blocks, int32_t threads, int32_t iters) {
  int mode = iters / 100000; iters = iters % 100000; int withStores = mode / 10; mode = mode % 10;
  Tmp<float4> scr; if (withStores && scr.alloc((size_t)blocks * threads) != cudaSuccess) return -1.0f;
  float4* scrp = withStores ? scr.p : nullptr;
  if (dev_ok(device) < 0) return -1.0f;
  Tmp<unsigned> ctr; if (ctr.alloc(256) != cudaSuccess) return -1.0f;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(ctr.p, 0, 1024);
    void* args[] = {(void*)&ctr.p, (void*)&iters, (void*)&mode, (void*)&scrp};
    cudaEventRecord(a);
    if (cudaLaunchCooperativeKernel((const void*)k_dbg_barrier, dim3(blocks), dim3(threads), args, 0, 0) != cudaSuccess) return -1.0f;
    cudaEventRecord(b);
    if (cudaEventSynchronize(b) != cudaSuccess) return -1.0f;
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return best * 1000.0f / iters;
}

int32_t dbx_debug_collide(int32_t device, int32_t n, const dbx_shape* shapesA, const float* xfA, const dbx_shape* shapesB, const float* xfB, dbx_manifold* out) {
  int rc = dev_ok(device); if (rc < 0) return rc;
  if (n <= 0) return 0;
  Tmp<DShape> dA, dB; Tmp<float4> xa, xb; Tmp<dbx_manifold> o;
  rc = shapes_up(n, shapesA, shapesB, dA, dB); if (rc < 0) return rc;
  if (xa.alloc(n) != cudaSuccess || xb.alloc(n) != cudaSuccess || o.alloc(n) != cudaSuccess) return DBX_E_CUDA;
  cudaMemcpy(xa.p, xfA, (size_t)n * 16, cudaMemcpyHostToDevice); cudaMemcpy(xb.p, xfB, (size_t)n * 16, cudaMemcpyHostToDevice);
  k_dbg_collide<<<(n + 127) / 128, 128>>>(n, dA.p, xa.p, dB.p, xb.p, o.p);
  rc = finish("debug_collide"); if (rc < 0) return rc;
  cudaMemcpy(out, o.p, (size_t)n * sizeof(dbx_manifold), cudaMemcpyDeviceToHost);
  return n;
}

int32_t dbx_debug_distance(int32_t device, int32_t n, const dbx_shape* shapesA, const float* xfA, const dbx_shape* shapesB, const float* xfB, int32_t useRadii,
                           float* outDistance, dbx_vec2* outA, dbx_vec2* outB, int32_t* outIterations) {
  int rc = dev_ok(device); if (rc < 0) return rc;
  if (n <= 0) return 0;
  Tmp<DShape> dA, dB; Tmp<float4> xa, xb; Tmp<float> d; Tmp<float2> pa, pb; Tmp<int> it;
  rc = shapes_up(n, shapesA, shapesB, dA, dB); if (rc < 0) return rc;
  if (xa.alloc(n) != cudaSuccess || xb.alloc(n) != cudaSuccess || d.alloc(n) != cudaSuccess || pa.alloc(n) != cudaSuccess || pb.alloc(n) != cudaSuccess || it.alloc(n) != cudaSuccess) return DBX_E_CUDA;
  cudaMemcpy(xa.p, xfA, (size_t)n * 16, cudaMemcpyHostToDevice); cudaMemcpy(xb.p, xfB, (size_t)n * 16, cudaMemcpyHostToDevice);
  k_dbg_distance<<<(n + 127) / 128, 128>>>(n, dA.p, xa.p, dB.p, xb.p, useRadii, d.p, pa.p, pb.p, it.p);
  rc = finish("debug_distance"); if (rc < 0) return rc;
  cudaMemcpy(outDistance, d.p, (size_t)n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(outA, pa.p, (size_t)n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(outB, pb.p, (size_t)n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(outIterations, it.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
  return n;
}

int32_t dbx_debug_time_of_impact(int32_t device, int32_t n, const dbx_shape* shapesA, const float* sweepsA, const dbx_shape* shapesB, const float* sweepsB, float tMax,
                                 int32_t* outState, float* outT) {
  int rc = dev_ok(device); if (rc < 0) return rc;
  if (n <= 0) return 0;
  Tmp<DShape> dA, dB; Tmp<float> sa, sb, t; Tmp<int> st;
  rc = shapes_up(n, shapesA, shapesB, dA, dB); if (rc < 0) return rc;
  if (sa.alloc(9 * (size_t)n) != cudaSuccess || sb.alloc(9 * (size_t)n) != cudaSuccess || t.alloc(n) != cudaSuccess || st.alloc(n) != cudaSuccess) return DBX_E_CUDA;
  cudaMemcpy(sa.p, sweepsA, (size_t)n * 36, cudaMemcpyHostToDevice); cudaMemcpy(sb.p, sweepsB, (size_t)n * 36, cudaMemcpyHostToDevice);
  k_dbg_toi<<<(n + 63) / 64, 64>>>(n, dA.p, sa.p, dB.p, sb.p, tMax, st.p, t.p);
  rc = finish("debug_time_of_impact"); if (rc < 0) return rc;
  cudaMemcpy(outState, st.p, (size_t)n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(outT, t.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
  return n;
}

}  // extern "C"
