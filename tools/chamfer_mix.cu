// Pipe-mix laboratory for the Chamfer scan (sm_100a): how fast can the inner loop go, piece by piece?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_bin/chamfer_mix tools/chamfer_mix.cu && build_bin/chamfer_mix
// Every kernel scans the same shared-memory tile of B points with 8 A points per thread (256 threads, 2 CTAs/SM,
// grid = 2 x SMs x WAVES) and differs only in what it does with the distances:
//   F = 1  row minimum only (FMNMX3)                       -> the FP32-pipe ceiling of the mix (6 pipe slots / pair)
//   F = 2  + argmin bookkeeping (FSETP + SEL per A point per step)
//   F = 4  + column minima c0..c3 (FMNMX3), folded into a dummy without cross-lane work
//   F = 8  + REDUX.MIN + ballot + owner select per B point (the real column side, no shared-memory merge)
// PACK = 0 uses scalar FADD/FMUL/FFMA (same arithmetic), PACK = 1 the packed f32x2 forms.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ float lo2(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi2(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ u64 dist2x2(u64 bx, u64 by, u64 bz, u64 nqx, u64 nqy, u64 nqz) {
  u64 dx = add2(bx, nqx), dy = add2(by, nqy), dz = add2(bz, nqz);
  return fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
}
__device__ __forceinline__ u64 dist2x2_scalar(u64 bx, u64 by, u64 bz, float qx, float qy, float qz) {
  const float dx0 = lo2(bx) - qx, dy0 = lo2(by) - qy, dz0 = lo2(bz) - qz;
  const float dx1 = hi2(bx) - qx, dy1 = hi2(by) - qy, dz1 = hi2(bz) - qz;
  return pack2(__fmaf_rn(dz0, dz0, __fmaf_rn(dx0, dx0, __fmul_rn(dy0, dy0))), __fmaf_rn(dz1, dz1, __fmaf_rn(dx1, dx1, __fmul_rn(dy1, dy1))));
}

constexpr int TILE = 2048, Q = 8;

template <int F, int PACK>
__global__ void __launch_bounds__(256, 2) mix_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int reps) {
  __shared__ __align__(16) float sx[TILE], sy[TILE], sz[TILE];
  __shared__ unsigned sink[256];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int j = tid; j < TILE; j += 256) { sx[j] = b[j * 3]; sy[j] = b[j * 3 + 1]; sz[j] = b[j * 3 + 2]; }
  float qx[Q], qy[Q], qz[Q], best[Q];
  u64 nqx[Q], nqy[Q], nqz[Q];
  int cstep[Q];
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int i = (blockIdx.x * 256 + tid) * Q + q;
    qx[q] = a[(i % 65536) * 3]; qy[q] = a[(i % 65536) * 3 + 1]; qz[q] = a[(i % 65536) * 3 + 2];
    nqx[q] = pack2(-qx[q], -qx[q]); nqy[q] = pack2(-qy[q], -qy[q]); nqz[q] = pack2(-qz[q], -qz[q]);
    best[q] = 1e30f; cstep[q] = 0;
  }
  __syncthreads();
  unsigned acc = 0, key_m = 0xffffffffu, key_b = 1u;
  int step = 0;
  for (int r = 0; r < reps; r++) {
    for (int j32 = 0; j32 < TILE; j32 += 32) {
#pragma unroll 2
      for (int j = j32; j < j32 + 32; j += 4, step++) {
        const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(&sx[j]);
        const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(&sy[j]);
        const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(&sz[j]);
        float c0 = 1e30f, c1 = 1e30f, c2 = 1e30f, c3 = 1e30f;
#pragma unroll
        for (int q = 0; q < Q; q += 2) {
          u64 d01, d23, e01, e23;
          if (PACK) {
            d01 = dist2x2(X.x, Y.x, Z.x, nqx[q], nqy[q], nqz[q]); d23 = dist2x2(X.y, Y.y, Z.y, nqx[q], nqy[q], nqz[q]);
            e01 = dist2x2(X.x, Y.x, Z.x, nqx[q + 1], nqy[q + 1], nqz[q + 1]); e23 = dist2x2(X.y, Y.y, Z.y, nqx[q + 1], nqy[q + 1], nqz[q + 1]);
          } else {
            d01 = dist2x2_scalar(X.x, Y.x, Z.x, qx[q], qy[q], qz[q]); d23 = dist2x2_scalar(X.y, Y.y, Z.y, qx[q], qy[q], qz[q]);
            e01 = dist2x2_scalar(X.x, Y.x, Z.x, qx[q + 1], qy[q + 1], qz[q + 1]); e23 = dist2x2_scalar(X.y, Y.y, Z.y, qx[q + 1], qy[q + 1], qz[q + 1]);
          }
          float nb0 = min3(best[q], lo2(d01), hi2(d01)); nb0 = min3(nb0, lo2(d23), hi2(d23));
          if (F & 2) { if (nb0 < best[q]) cstep[q] = step; }
          best[q] = nb0;
          float nb1 = min3(best[q + 1], lo2(e01), hi2(e01)); nb1 = min3(nb1, lo2(e23), hi2(e23));
          if (F & 2) { if (nb1 < best[q + 1]) cstep[q + 1] = step; }
          best[q + 1] = nb1;
          if (F & 12) {
            c0 = min3(c0, lo2(d01), lo2(e01)); c1 = min3(c1, hi2(d01), hi2(e01));
            c2 = min3(c2, lo2(d23), lo2(e23)); c3 = min3(c3, hi2(d23), hi2(e23));
          }
        }
        if ((F & 12) == 4) acc ^= __float_as_uint(c0) ^ __float_as_uint(c1) ^ __float_as_uint(c2) ^ __float_as_uint(c3);
        if (F & 8) {
          const unsigned b0 = __float_as_uint(c0), b1 = __float_as_uint(c1), b2 = __float_as_uint(c2), b3 = __float_as_uint(c3);
          const unsigned m0 = __reduce_min_sync(0xffffffffu, b0), m1 = __reduce_min_sync(0xffffffffu, b1);
          const unsigned m2 = __reduce_min_sync(0xffffffffu, b2), m3 = __reduce_min_sync(0xffffffffu, b3);
          const unsigned l0 = __ballot_sync(0xffffffffu, b0 == m0), l1 = __ballot_sync(0xffffffffu, b1 == m1);
          const unsigned l2 = __ballot_sync(0xffffffffu, b2 == m2), l3 = __ballot_sync(0xffffffffu, b3 == m3);
          const int o = lane - (j - j32);
          if (o == 0) { key_m = m0; key_b = l0; }
          if (o == 1) { key_m = m1; key_b = l1; }
          if (o == 2) { key_m = m2; key_b = l2; }
          if (o == 3) { key_m = m3; key_b = l3; }
        }
      }
      if (F & 8) acc ^= key_m + __ffs(key_b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < Q; q++) s += best[q] + (float)cstep[q];
  sink[tid] = acc;
  out[blockIdx.x * 256 + tid] = s + (float)sink[(tid + 1) & 255];
}

// Dispatch probe: NF packed FFMA2 (or 2*NF scalar FFMA) + NA independent ALU-pipe ops (LOP3) per iteration, one
// warp per scheduler and four; cycles per iteration tell whether a packed instruction holds the dispatch port for
// one cycle or two (i.e. whether ALU work can be issued "under" the packed FP32 work).
template <int NF, int NA, int PACK>
__global__ void __launch_bounds__(512) dispatch_probe(float* out, long long* cyc, float a, float b, unsigned m) {
  u64 acc[16];
  float sacc[32];
  unsigned x[16];
  const u64 a2 = pack2(a, a), b2 = pack2(b, b);
#pragma unroll
  for (int i = 0; i < 16; i++) { acc[i] = pack2(threadIdx.x * 0.001f + i, (float)i); x[i] = threadIdx.x * 7919u + i; }
#pragma unroll
  for (int i = 0; i < 32; i++) sacc[i] = threadIdx.x * 0.001f + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < 4096; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (i < NF) {
        if (PACK) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a2), "l"(b2));
        else {
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(sacc[2 * i]) : "f"(a), "f"(b));
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(sacc[2 * i + 1]) : "f"(a), "f"(b));
        }
      }
      if (i < NA) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(m), "r"(it));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  unsigned y = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) { s += lo2(acc[i]) + hi2(acc[i]); y ^= x[i]; }
#pragma unroll
  for (int i = 0; i < 32; i++) s += sacc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)y;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NF, int NA, int PACK>
static void probe(const char* what, float* out, int threads) {
  long long* cyc; CK(cudaMalloc(&cyc, 64));
  dispatch_probe<NF, NA, PACK><<<1, threads>>>(out, cyc, 1.0001f, 0.5f, 0x5a5a5a5au);
  CK(cudaDeviceSynchronize());
  long long h = 0; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("{\"probe\": \"%s\", \"fp32_lane_ops_per_iter\": %d, \"packed\": %d, \"alu_ops_per_iter\": %d, \"warps_per_scheduler\": %d, \"cycles_per_iter_per_warp_slot\": %.2f}\n",
         what, 2 * NF, PACK, NA, threads / 128, (double)h / 4096.0 / (threads / 128));
  cudaFree(cyc);
}

template <int F, int PACK>
static void run(const char* name, const float* a, const float* b, float* out, int sms) {
  const int waves = 3, reps = 4;
  const int grid = 2 * sms * waves;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  mix_kernel<F, PACK><<<grid, 256>>>(a, b, out, reps);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < 5; i++) {
    CK(cudaEventRecord(e0));
    mix_kernel<F, PACK><<<grid, 256>>>(a, b, out, reps);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double pairs = (double)grid * 256 * Q * TILE * reps;
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double peak_pairs = (double)sms * 128 * (clk * 1e3) / 6.0;  // 6 FP32-pipe lane-slots per pair
  printf("{\"mix\": \"%s\", \"flags\": %d, \"packed\": %d, \"ms\": %.4f, \"unique_tpair_per_s\": %.3f, \"frac_of_fp32_pipe_ceiling\": %.3f}\n", name, F, PACK, best,
         pairs / best / 1e9, pairs / (best * 1e-3) / peak_pairs);
}

int main() {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float *a, *b, *out;
  CK(cudaMalloc(&a, 65536 * 3 * 4)); CK(cudaMalloc(&b, TILE * 3 * 4)); CK(cudaMalloc(&out, (size_t)2 * sms * 3 * 256 * 4));
  float* h = (float*)malloc(65536 * 3 * 4);
  srand(1);
  for (int i = 0; i < 65536 * 3; i++) h[i] = rand() / (float)RAND_MAX - 0.5f;
  CK(cudaMemcpy(a, h, 65536 * 3 * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b, h + 1000, TILE * 3 * 4, cudaMemcpyHostToDevice));
  run<1, 1>("row min only", a, b, out, sms);
  run<1, 0>("row min only, scalar", a, b, out, sms);
  run<3, 1>("+ argmin bookkeeping", a, b, out, sms);
  run<7, 1>("+ column minima (no cross-lane)", a, b, out, sms);
  run<11, 1>("+ REDUX/ballot/owner (full column side)", a, b, out, sms);
  run<11, 0>("full, scalar", a, b, out, sms);
  run<9, 1>("row min + full column, no bookkeeping", a, b, out, sms);
  for (int threads : {128, 512}) {
    if (threads == 128) {
      probe<16, 0, 1>("16 FFMA2", out, 128); probe<16, 16, 1>("16 FFMA2 + 16 LOP3", out, 128); probe<16, 8, 1>("16 FFMA2 + 8 LOP3", out, 128);
      probe<16, 0, 0>("32 FFMA", out, 128); probe<16, 16, 0>("32 FFMA + 16 LOP3", out, 128); probe<0, 16, 1>("16 LOP3", out, 128);
    } else {
      probe<16, 0, 1>("16 FFMA2", out, 512); probe<16, 16, 1>("16 FFMA2 + 16 LOP3", out, 512); probe<16, 8, 1>("16 FFMA2 + 8 LOP3", out, 512);
      probe<16, 0, 0>("32 FFMA", out, 512); probe<16, 16, 0>("32 FFMA + 16 LOP3", out, 512); probe<0, 16, 1>("16 LOP3", out, 512);
    }
  }
  return 0;
}
