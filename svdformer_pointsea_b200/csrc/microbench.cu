// Microbenchmarks that set the roofline denominators and the FPS latency floor on B200.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench microbench.cu
// Run  : ./microbench            (prints one JSON object per line)
//
// What it measures (all on the current device, CUDA-event or clock64 timed):
//   ffma        scalar FFMA issue rate (3 distinct source registers)
//   ffma2       packed fma.rn.f32x2 rate (sm_100 FFMA2)
//   chamfer_mix the exact instruction mix of the Chamfer inner loop out of registers:
//               3 FADD2 + FMUL2 + 2 FFMA2 per two pairs + one FMNMX3
//   chamfer_mix_scalar  same mix with scalar FADD/FMUL/FFMA/FMNMX
//   redux       redux.sync.max.s32 dependent-chain latency
//   bar         __syncthreads() period at 512 threads
//   cluster_bar barrier.cluster arrive+wait period at cluster size 4/8/16
//   st_async    all-to-all st.async + mbarrier round per iteration at cluster size 4/8/16
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

typedef unsigned long long u64;

__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { return ((u64)__float_as_uint(hi) << 32) | __float_as_uint(lo); }

constexpr int ITERS = 2048;

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = __fmaf_rn(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, float a, float b) {
  u64 acc[16];
  u64 a2 = pack2(a, a), b2 = pack2(b, b);
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = pack2(threadIdx.x * 0.001f + i, i);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = fma2(acc[i], a2, b2);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Chamfer mix, packed: per (query, 2 targets): 3 FADD2, 1 FMUL2, 2 FFMA2, 1 FMNMX3.
template <int Q>
__global__ void __launch_bounds__(256) k_mix2(float* out, const float* in) {
  u64 nq[Q][3];
  float best[Q];
#pragma unroll
  for (int q = 0; q < Q; q++) {
    best[q] = 1e30f;
#pragma unroll
    for (int c = 0; c < 3; c++) { float v = in[(threadIdx.x * Q + q) * 3 + c]; nq[q][c] = pack2(-v, -v); }
  }
  u64 bx = pack2(in[0], in[1]), by = pack2(in[2], in[3]), bz = pack2(in[4], in[5]);
  u64 inc = pack2(in[6], in[7]);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int q = 0; q < Q; q++) {
      u64 dx = add2(bx, nq[q][0]), dy = add2(by, nq[q][1]), dz = add2(bz, nq[q][2]);
      u64 d = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
      best[q] = min3(best[q], __uint_as_float((unsigned)d), __uint_as_float((unsigned)(d >> 32)));
    }
    bx = add2(bx, inc);  // keep the loop from being hoisted (1 extra op per Q groups)
  }
  float s = 0;
#pragma unroll
  for (int q = 0; q < Q; q++) s += best[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int Q>
__global__ void __launch_bounds__(256) k_mix1(float* out, const float* in) {
  float qx[Q], qy[Q], qz[Q], best[Q];
#pragma unroll
  for (int q = 0; q < Q; q++) {
    best[q] = 1e30f;
    qx[q] = in[(threadIdx.x * Q + q) * 3 + 0]; qy[q] = in[(threadIdx.x * Q + q) * 3 + 1]; qz[q] = in[(threadIdx.x * Q + q) * 3 + 2];
  }
  float bx0 = in[0], bx1 = in[1], by0 = in[2], by1 = in[3], bz0 = in[4], bz1 = in[5], inc = in[6];
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int q = 0; q < Q; q++) {
      float dx0 = bx0 - qx[q], dy0 = by0 - qy[q], dz0 = bz0 - qz[q];
      float dx1 = bx1 - qx[q], dy1 = by1 - qy[q], dz1 = bz1 - qz[q];
      float d0 = __fmaf_rn(dz0, dz0, __fmaf_rn(dx0, dx0, __fmul_rn(dy0, dy0)));
      float d1 = __fmaf_rn(dz1, dz1, __fmaf_rn(dx1, dx1, __fmul_rn(dy1, dy1)));
      best[q] = min3(best[q], d0, d1);
    }
    bx0 += inc; bx1 += inc;
  }
  float s = 0;
#pragma unroll
  for (int q = 0; q < Q; q++) s += best[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_redux_lat(long long* cyc, int* out) {
  int v = threadIdx.x * 7919;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; i++) { v = __reduce_max_sync(0xffffffffu, v ^ i) + (threadIdx.x & 1); }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = v;
}

__global__ void __launch_bounds__(512) k_bar_lat(long long* cyc, int* out) {
  __shared__ int s[16];
  int v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < 1024; i++) {
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v + i;
    __syncthreads();
    v += s[(threadIdx.x + i) & 15];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = v;
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }

__global__ void __launch_bounds__(512) k_cluster_bar_lat(long long* cyc, int* out) {
  int v = threadIdx.x;
  cluster_arrive(); cluster_wait();
  long long t0 = clock64();
  for (int i = 0; i < 512; i++) { cluster_arrive(); cluster_wait(); v += i; }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

// all-to-all message round: lane c of warp 0 in each CTA pushes 16+4 bytes to CTA c's slot, every
// thread waits on the local mbarrier (double-buffered), reads the C slots and goes round again.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) { unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_async_v4(unsigned raddr, unsigned rbar, unsigned a, unsigned b, unsigned c, unsigned d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_b32(unsigned raddr, unsigned rbar, unsigned a) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(a), "r"(rbar) : "memory");
}

__global__ void __launch_bounds__(512) k_st_async_lat(long long* cyc, int* out, int with_bar) {
  __shared__ __align__(16) unsigned slots[2][16][8];
  __shared__ __align__(8) u64 bars[2];
  __shared__ int wres[16];
  const unsigned C = cluster_nctarank(), me = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(smem_u32(&bars[0]), C * 20); mbar_expect_tx(smem_u32(&bars[1]), C * 20);
  }
  cluster_arrive(); cluster_wait();
  int v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < 1024; i++) {
    const int b = i & 1;
    if (with_bar) {
      if ((threadIdx.x & 31) == 0) wres[threadIdx.x >> 5] = v;
      __syncthreads();
    }
    if (threadIdx.x < C) {
      int m = with_bar ? wres[threadIdx.x & 15] : v;
      unsigned ra = mapa(smem_u32(&slots[b][me][0]), threadIdx.x);
      unsigned rb = mapa(smem_u32(&bars[b]), threadIdx.x);
      st_async_v4(ra, rb, (unsigned)m, me, i, 7);
      st_async_b32(ra + 16, rb, i);
    }
    mbar_wait(smem_u32(&bars[b]), (i >> 1) & 1);
    int s = 0;
    for (unsigned c = 0; c < C; c++) s += slots[b][c][0] + slots[b][c][4];
    v += s;
    if (threadIdx.x == 0) mbar_expect_tx(smem_u32(&bars[b]), C * 20);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && me == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  cluster_arrive(); cluster_wait();
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

template <typename K>
static void launch_cluster(K kern, int nblocks, int csize, long long* cyc, int* out, int extra, bool has_extra) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nblocks); cfg.blockDim = dim3(512);
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  (void)extra; (void)has_extra;
  void* args3[] = {&cyc, &out, &extra};
  CK(cudaLaunchKernelExC(&cfg, (const void*)kern, args3));
}

int main() {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
  printf("{\"bench\": \"device\", \"name\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_mhz\": %.0f}\n", p.name, p.multiProcessorCount, p.major, p.minor, clk_khz / 1000.0);
  const int nsm = p.multiProcessorCount;
  const int grid = nsm * 16, threads = 256;
  float *out, *in; CK(cudaMalloc(&out, (size_t)grid * 512 * 4)); CK(cudaMalloc(&in, 1 << 20)); CK(cudaMemset(in, 0, 1 << 20));
  {
    float ms = time_ms([&] { k_ffma<<<grid, threads>>>(out, 1.0001f, 0.5f); });
    double fl = (double)grid * threads * ITERS * 16 * 2;
    printf("{\"bench\": \"ffma\", \"ms\": %.4f, \"tflops\": %.2f}\n", ms, fl / ms / 1e9);
  }
  {
    float ms = time_ms([&] { k_ffma2<<<grid, threads>>>(out, 1.0001f, 0.5f); });
    double fl = (double)grid * threads * ITERS * 16 * 4;
    printf("{\"bench\": \"ffma2\", \"ms\": %.4f, \"tflops\": %.2f}\n", ms, fl / ms / 1e9);
  }
#define MIX(K, Q, NAME) { float ms = time_ms([&] { K<Q><<<grid, threads>>>(out, in); }); double pairs = (double)grid * threads * ITERS * Q * 2; \
    printf("{\"bench\": \"%s\", \"Q\": %d, \"ms\": %.4f, \"gpairs_per_s\": %.1f, \"tflops_8\": %.2f}\n", NAME, Q, ms, pairs / ms / 1e6, pairs * 8 / ms / 1e9); }
  MIX(k_mix2, 4, "chamfer_mix_f32x2") MIX(k_mix2, 8, "chamfer_mix_f32x2") MIX(k_mix2, 12, "chamfer_mix_f32x2")
  MIX(k_mix1, 4, "chamfer_mix_scalar") MIX(k_mix1, 8, "chamfer_mix_scalar")
  CK(cudaGetLastError());

  long long* cyc; int* iout; CK(cudaMalloc(&cyc, 64)); CK(cudaMalloc(&iout, 16 * 512 * 16 * 4));
  long long h;
  k_redux_lat<<<1, 32>>>(cyc, iout); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("{\"bench\": \"redux_chain\", \"cycles_per_op\": %.1f}\n", h / 1024.0);
  k_bar_lat<<<1, 512>>>(cyc, iout); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("{\"bench\": \"bar_sync_512\", \"cycles_per_iter\": %.1f}\n", h / 1024.0);
  CK(cudaFuncSetAttribute(k_cluster_bar_lat, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(k_st_async_lat, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (int cs : {1, 2, 4, 8, 16}) {
    launch_cluster(k_cluster_bar_lat, cs, cs, cyc, iout, 0, false);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"bench\": \"cluster_bar\", \"cluster\": %d, \"error\": \"%s\"}\n", cs, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"cluster_bar\", \"cluster\": %d, \"cycles_per_iter\": %.1f}\n", cs, h / 512.0);
  }
  for (int wb : {0, 1}) for (int cs : {1, 2, 4, 8, 16}) {
    launch_cluster(k_st_async_lat, cs, cs, cyc, iout, wb, true);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"bench\": \"st_async_round\", \"cluster\": %d, \"error\": \"%s\"}\n", cs, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"bench\": \"st_async_round\", \"cluster\": %d, \"with_cta_bar\": %d, \"cycles_per_iter\": %.1f}\n", cs, wb, h / 1024.0);
  }
  return 0;
}
