// Shared host/device helpers for the pointsea_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/pointsea_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pointsea_b200 kernels are written for sm_100a (B200) only"
#endif

struct ps_comm;

namespace ps {

typedef unsigned long long u64;

// ---- error plumbing -------------------------------------------------------------------------
char* err_buf();                 // thread-local 512-byte buffer
long long& launch_counter();     // thread-local launch counter
int set_error(int code, const char* fmt, ...);

#define PS_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ps::set_error(PS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                              \
  } while (0)

#define PS_REQUIRE(cond, ...)                                        \
  do {                                                               \
    if (!(cond)) return ps::set_error(PS_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

// Counts the launch and converts a launch error into PS_ERR_CUDA.
#define PS_LAUNCH_CHECK()                                                                    \
  do {                                                                                       \
    ps::launch_counter()++;                                                                  \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return ps::set_error(PS_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                  \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                      \
  } while (0)

// RAII device guard: the reference needs torch.cuda.set_device around Chamfer
// (dist_chamfer_3D.py:43); here every entry point switches and restores explicitly.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

int sm_count(int dev);
// stream-ordered scratch allocation from the library's own memory pool (freed blocks stay cached up to a bounded
// threshold; torch's allocator and the device's default pool are left alone); free with cudaFreeAsync
int scratch_alloc(void** ptr, size_t bytes, int dev, cudaStream_t stream);
// creates the device's scratch pool now (entry points that capture call this BEFORE cudaStreamBeginCapture:
// creating a pool is not allowed while the thread is capturing)
void scratch_pool_init(int dev);
// A caller-owned arena that replaces the stream-ordered allocations of the launchers running on this host
// thread while it is installed (the graph-captured step entry points install one: their graphs then hold no
// memory nodes and stay updatable).  Requests beyond it fall back to the pool and record what was needed.
struct Workspace {
  char* base = nullptr;
  size_t bytes = 0, used = 0, needed = 0;
};
Workspace*& tls_workspace();
// RAII owner of one scratch block: every early return of a launcher gives the block back to the pool
struct ScratchGuard {
  void* ptr = nullptr;
  cudaStream_t stream = nullptr;
  bool pooled = false;
  int alloc(size_t bytes, int dev, cudaStream_t s) {
    stream = s;
    if (Workspace* w = tls_workspace()) {
      const size_t need = (bytes + 255) & ~(size_t)255;
      w->needed += need;
      if (w->used + need <= w->bytes) {
        ptr = w->base + w->used;
        w->used += need;
        return PS_OK;
      }
    }
    pooled = true;
    return scratch_alloc(&ptr, bytes, dev, s);
  }
  int release() {
    void* q = ptr;
    ptr = nullptr;
    if (q && pooled) PS_CUDA(cudaFreeAsync(q, stream));
    return PS_OK;
  }
  ~ScratchGuard() { if (ptr && pooled) cudaFreeAsync(ptr, stream); }
  ScratchGuard() = default;
  ScratchGuard(const ScratchGuard&) = delete;
  ScratchGuard& operator=(const ScratchGuard&) = delete;
};
// fills `bytes` (multiple of 4) with the 32-bit pattern `value` by a kernel launch (see runtime.cu for why not memset)
int fill32_async(void* ptr, unsigned value, size_t bytes, cudaStream_t stream);

// chamfer_sym.cu: PS_OK when handled, 1 when the two-pass kernel should run instead
// sums6 != null: the six loss sums of ps_chamfer_sums come out of the same epilogue launch; comm != null (needs sums6):
// that launch also publishes them to the peers' mailboxes (comm.cuh)
int chamfer_fwd_symmetric(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int* idx1,
                          int* idx2, double* sums6, const struct ::ps_comm* comm, int B, int N, int M, int dev,
                          cudaStream_t stream);

// chamfer.cu: the six loss sums of ps_chamfer_sums; accumulate != 0 adds to out6 instead of overwriting it
int chamfer_sums_launch(const float* dist1, const float* dist2, double* out6, long long n1, long long n2, int accumulate,
                        int dev, cudaStream_t stream);

// knn_select.cu: PS_OK when handled, 1 when the streaming kernel in neighbors.cu should run instead
int knn_select_launch(const float* xyz, const float* new_xyz, int* idx, float* gxyz, int B, int N, int S, int k,
                      int skip, int order, int var, int nsm, cudaStream_t stream);

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device arithmetic ----------------------------------------------------------------------
#ifdef __CUDACC__

// The reference's squared distance as nvcc contracts it for every kernel on this path
// (verified in the sm_100a SASS of chamfer3D.cu:34-37, sampling_gpu.cu:103-104,
// ball_query_gpu.cu:30-31, interpolate_gpu.cu:33): FMUL on y, FFMA on x, FFMA on z.
__device__ __forceinline__ float dist2_ref(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2).  Each half rounds exactly like the
// scalar .rn instruction, so results are bit-identical to the scalar expression above.
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ float lo2(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi2(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
// 3-input min / max (sm_100 FMNMX3); NaN handling as fminf/fmaxf (returns the non-NaN operand).
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

// packed squared distance of one query against two targets: d = {dist2(b0-q), dist2(b1-q)}
// nq* hold the NEGATED query coordinate in both halves (b + (-q) == b - q exactly).
__device__ __forceinline__ u64 dist2x2(u64 bx, u64 by, u64 bz, u64 nqx, u64 nqy, u64 nqz) {
  u64 dx = add2(bx, nqx), dy = add2(by, nqy), dz = add2(bz, nqz);
  return fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- thread-block cluster primitives ---------------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned mapa_shared(unsigned addr, unsigned rank) { unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void st_cluster_v4(unsigned addr, unsigned a, unsigned b, unsigned c, unsigned d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(unsigned addr, unsigned a, unsigned b) {
  asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

#endif  // __CUDACC__

}  // namespace ps
