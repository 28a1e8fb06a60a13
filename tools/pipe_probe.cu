// Pipe-rate probe for the Chamfer inner loop on sm_100a: warp-instructions per cycle per SM
// sub-partition for single opcodes and for the mixes the kernels use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_bin/pipe_probe tools/pipe_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float min2(float a, float b) { float d; asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { return ((u64)__float_as_uint(hi) << 32) | __float_as_uint(lo); }
__device__ __forceinline__ float lo2(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi2(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

constexpr int ITERS = 4096;
constexpr int NACC = 16;

// MODE: 0 FADD2, 1 FMUL2, 2 FFMA2 (3 distinct), 3 FFMA2 (a,a,c), 4 FMNMX3, 5 FMNMX2, 6 FSETP+SEL, 7 FFMA2+FMNMX3 1:1,
//       8 FADD2+FMNMX3 1:1, 9 full chamfer pair mix (6 FP2 + 2 FMNMX3 per 2 pairs), 10 = 9 + FSETP/SEL bookkeeping (per 4 pairs)
template <int MODE>
__global__ void __launch_bounds__(1024) probe(float* out, const float* in, long long* cyc) {
  u64 acc[NACC];
  float f[NACC];
  int s[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { acc[i] = pack2(in[i] + threadIdx.x, in[i + 1]); f[i] = in[i + 2] + threadIdx.x; s[i] = 0; }
  const u64 c1 = pack2(in[40], in[41]), c2 = pack2(in[42], in[43]);
  const float g1 = in[44], g2 = in[45];
  __syncthreads();
  const long long t0 = clock64();
  u64 v1 = c1, v2 = c2, v3 = pack2(in[46], in[47]);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (MODE == 0) acc[i] = add2(acc[i], c1);
      if (MODE == 1) acc[i] = mul2(acc[i], c1);
      if (MODE == 2) acc[i] = fma2(acc[i], c1, c2);
      if (MODE == 3) acc[i] = fma2(c1, c1, acc[i]);
      if (MODE == 4) f[i] = min3(f[i], g1, g2);
      if (MODE == 5) f[i] = min2(f[i], g1);
      if (MODE == 6) { acc[i] = fma2(acc[i], c1, c2); asm volatile("{.reg .pred p; setp.lt.f32 p, %1, %2; selp.b32 %0, %3, %0, p;}" : "+r"(s[i]) : "f"(f[i]), "f"(g1), "r"(it)); }
      if (MODE == 7) { acc[i] = fma2(acc[i], c1, c2); f[i] = min3(f[i], g1, g2); }
      if (MODE == 8) { acc[i] = add2(acc[i], c1); f[i] = min3(f[i], g1, g2); }
      if (MODE == 11) { acc[i] = fma2(acc[i], c1, c2); f[i] = min2(f[i], g1); }
      if (MODE == 12) { acc[i] = fma2(acc[i], c1, c2); asm volatile("add.s32 %0, %0, %1;" : "+r"(s[i]) : "r"(it)); }
      if (MODE == 13) { acc[i] = fma2(acc[i], c1, c2); acc[i] = fma2(acc[i], c1, c2); f[i] = min3(f[i], g1, g2); }
      if (MODE == 14) { acc[i] = fma2(acc[i], c1, c2); acc[i] = fma2(acc[i], c1, c2); acc[i] = fma2(acc[i], c1, c2); f[i] = min3(f[i], g1, g2); }
      if (MODE == 9 || MODE == 10) {
        // two pairs: the targets (v1, v2) change every iteration, so nothing hoists
        u64 dx = add2(v1, acc[i]), dy = add2(v2, acc[i]), dz = add2(v3, acc[i]);
        u64 d = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
        float nb = min3(f[i], lo2(d), hi2(d));            // row side
        f[(i + 5) % NACC] = min3(f[(i + 5) % NACC], lo2(d), hi2(d));  // column side stand-in
        if (MODE == 10 && (i & 1)) asm volatile("{.reg .pred p; setp.lt.f32 p, %1, %2; selp.b32 %0, %3, %0, p;}" : "+r"(s[i]) : "f"(nb), "f"(f[i]), "r"(it));
        f[i] = nb;
      }
    }
    if (MODE == 9 || MODE == 10) { v1 = add2(v1, c2); v2 = add2(v2, c1); v3 = add2(v3, c1); }
  }
  const long long t1 = clock64();
  float r = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) r += lo2(acc[i]) + hi2(acc[i]) + f[i] + s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
static void run(const char* name, double inst_per_inner, int nsm, float* out, float* in, long long* cyc, int warps_per_smsp) {
  const int threads = 32 * 4 * warps_per_smsp;  // one CTA per SM
  probe<MODE><<<nsm, threads>>>(out, in, cyc);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  probe<MODE><<<nsm, threads>>>(out, in, cyc);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  const double inst = (double)ITERS * NACC * inst_per_inner * warps_per_smsp;  // warp-instructions per SMSP
  printf("{\"probe\": \"%s\", \"warps_per_smsp\": %d, \"cycles\": %lld, \"inst_per_cycle_per_smsp\": %.3f, \"cycles_per_inner\": %.3f}\n",
         name, warps_per_smsp, h, inst / h, (double)h / ((double)ITERS * NACC * warps_per_smsp));
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  float *out, *in; long long* cyc;
  CK(cudaMalloc(&out, (size_t)p.multiProcessorCount * 1024 * 4)); CK(cudaMalloc(&in, 4096)); CK(cudaMemset(in, 0, 4096)); CK(cudaMalloc(&cyc, 8));
  const int n = p.multiProcessorCount;
  for (int w : {1, 2, 4, 8}) {
    run<0>("FADD2", 1, n, out, in, cyc, w);
    run<2>("FFMA2 abc", 1, n, out, in, cyc, w);
    run<3>("FFMA2 aac", 1, n, out, in, cyc, w);
    run<4>("FMNMX3", 1, n, out, in, cyc, w);
    run<5>("FMNMX2", 1, n, out, in, cyc, w);
    run<6>("FFMA2+FSETP+SEL", 3, n, out, in, cyc, w);
    run<7>("FFMA2+FMNMX3", 2, n, out, in, cyc, w);
    run<8>("FADD2+FMNMX3", 2, n, out, in, cyc, w);
    run<11>("FFMA2+FMNMX2", 2, n, out, in, cyc, w);
    run<12>("FFMA2+IADD", 2, n, out, in, cyc, w);
    run<13>("2 FFMA2+FMNMX3", 3, n, out, in, cyc, w);
    run<14>("3 FFMA2+FMNMX3", 4, n, out, in, cyc, w);
    run<9>("pair mix 6FP2+2FMNMX3 (2 pairs)", 8, n, out, in, cyc, w);
    run<10>("pair mix + bookkeeping/2", 9, n, out, in, cyc, w);
  }
  return 0;
}
