// kNN by threshold selection: the fast path of ps_knn for N <= 2048 candidates and k+skip <= 16
// (every kNN call of SVDFormer / PointSea: N in {512, 2048}, k = 16).
//
// Replaces the torch expression query_knn / square_distance (models/model_utils.py:258-286) like
// knn_kernel in neighbors.cu, with a cheaper selection.  One warp per query:
//   1. every lane evaluates its candidates (j = 32u + lane) and tracks only its own minimum (one
//      FMNMX per candidate);
//   2. the 32 lane minima are sorted across the warp (values only); T = their KK-th smallest.  The
//      KK smallest lane minima are KK distinct candidates, so the true KK-th smallest distance is
//      <= T: the set {d <= T} CONTAINS the exact answer.  Its expected size for k=16 is
//      sum_{i<16} 32/(32-i) ~= 22 candidates;
//   3. the distances are evaluated again and the candidates with d <= T are compacted (ballot + prefix popcount) into a per-warp
//      shared-memory buffer of 32, sorted ascending by (distance, index) with one warp bitonic
//      sort, and the first KK are the answer (skip dropped) — the same (dist, index) order as the
//      streaming kernel and as torch's stable radix sort;
//   4. if more than 32 candidates pass (heavy duplication), the query falls back to the streaming
//      insertion over the same shared-memory tile.
// ~1100 warp instructions per query instead of ~2700 for the streaming insertion.
#include "common.cuh"
#include "select.cuh"

namespace ps {
namespace {

constexpr int KS_THREADS = 256;
constexpr int KS_WARPS = KS_THREADS / 32;
constexpr int KS_TILE = 2048;
// register allocation: plain bound (58 registers, 4 CTAs/SM).  Forcing 5 CTAs/SM (48 registers, small spills)
// measured slower: 0.168 vs 0.158 ms at C3 on the same box; an explicit minimum of 1 let ptxas use 84 registers.
#ifdef KS_MIN_CTAS
#define KS_BOUNDS __launch_bounds__(KS_THREADS, KS_MIN_CTAS)
#else
#define KS_BOUNDS __launch_bounds__(KS_THREADS)
#endif

// identical arithmetic to neighbors.cu (see the comments there and DESIGN.md "kNN arithmetic")
__device__ __forceinline__ float ks_sumsq(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(z, z)), __fmul_rn(y, y));
}
// Candidates are staged as PAIRS (j, j + 32) so that one lane evaluates two of them with the packed fp32x2
// instructions: element (64u + l, 64u + 32 + l) lives in A[32u + l] = {x_a, x_b, y_a, y_b} and
// B[32u + l] = {z_a, z_b, |p_a|^2, |p_b|^2}.  The coordinates are stored multiplied by -2: every step of the
// dot-product chain then carries the exact factor -2 (a power of two), so the chain yields (-2 * dot) with the
// same rounding as the reference's  -2 * matmul  and the separate multiplication disappears.
struct KsPair { u64 x, y, z, pp; };
__device__ __forceinline__ KsPair ks_load(const float4* __restrict__ A, const float4* __restrict__ Bv, int e) {
  const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(&A[e]);
  const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(&Bv[e]);
  return {a.x, a.y, b.x, b.y};
}
// {dist(q, cand_a), dist(q, cand_b)}; q* hold the query coordinate in both halves
template <int VAR>
__device__ __forceinline__ u64 ks_dist2(u64 qx, u64 qy, u64 qz, u64 qq, const KsPair& c) {
  u64 dot;
  if (VAR == 0) dot = fma2(qz, c.z, fma2(qy, c.y, mul2(qx, c.x)));
  else if (VAR == 1) dot = fma2(qx, c.x, fma2(qy, c.y, mul2(qz, c.z)));
  else dot = add2(add2(mul2(qx, c.x), mul2(qy, c.y)), mul2(qz, c.z));
  return add2(add2(dot, qq), c.pp);
}
// one candidate by index (fallback path, fused epilogue): {-2x, -2y, -2z, |p|^2}
__device__ __forceinline__ float4 ks_cand(const float4* __restrict__ A, const float4* __restrict__ Bv, int j) {
  const int e = (j >> 6) * 32 + (j & 31);
  const float4 a = A[e], b = Bv[e];
  return (j & 32) ? make_float4(a.y, a.w, b.y, b.w) : make_float4(a.x, a.z, b.x, b.z);
}
template <int VAR>
__device__ __forceinline__ float ks_dist1(float qx, float qy, float qz, float qq, float4 c) {
  float dot;  // c holds -2 * coordinates
  if (VAR == 0) dot = __fmaf_rn(qz, c.z, __fmaf_rn(qy, c.y, __fmul_rn(qx, c.x)));
  else if (VAR == 1) dot = __fmaf_rn(qx, c.x, __fmaf_rn(qy, c.y, __fmul_rn(qz, c.z)));
  else dot = __fadd_rn(__fadd_rn(__fmul_rn(qx, c.x), __fmul_rn(qy, c.y)), __fmul_rn(qz, c.z));
  return __fadd_rn(__fadd_rn(dot, qq), c.w);
}

// values only, ascending (NaN never reaches here: lane minima come from fminf)
__device__ __forceinline__ float ks_sort_values(float v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool keep_min = (((lane & k) == 0) == ((lane & j) == 0));
      v = keep_min ? fminf(v, o) : fmaxf(v, o);
    }
  }
  return v;
}

// QW queries per warp share every candidate load: the kernel is bound by shared-memory bandwidth
// otherwise (each query streams the 32 KB tile twice; measured 0.20 ms at C3 with QW = 1).
// EPI = 0: plain (dist, index) order, indices only (the C3 kernel, 48 registers); EPI = 1: result order and
// fused coordinate grouping selected at run time
template <int VAR, int QW, int EPI>
__global__ void KS_BOUNDS knn_select_kernel(const float* __restrict__ xyz,
                                                                const float* __restrict__ new_xyz,
                                                                int* __restrict__ idx, float* __restrict__ gxyz,
                                                                int N, int S, int k, int skip, int qpc, int order) {
  __shared__ float4 spa[KS_TILE / 2];  // {x_a, x_b, y_a, y_b} (times -2)
  __shared__ float4 spb[KS_TILE / 2];  // {z_a, z_b, |p_a|^2, |p_b|^2}
  __shared__ float bufd[KS_WARPS][QW][32];
  __shared__ int bufi[KS_WARPS][QW][32];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* cloud = xyz + (size_t)b * N * 3;
  const float INF = __int_as_float(0x7f800000);
  const int KK = k + skip;
  for (int e = tid; e < KS_TILE / 2; e += KS_THREADS) {
    const int ja = (e >> 5) * 64 + (e & 31), jb = ja + 32;
    float xa = 0.f, ya = 0.f, za = 0.f, pa = INF, xb = 0.f, yb = 0.f, zb = 0.f, pb = INF;  // padding: distance +inf
    if (ja < N) {
      const float x = __ldg(cloud + (size_t)ja * 3 + 0), y = __ldg(cloud + (size_t)ja * 3 + 1), z = __ldg(cloud + (size_t)ja * 3 + 2);
      xa = -2.0f * x; ya = -2.0f * y; za = -2.0f * z; pa = ks_sumsq(x, y, z);
    }
    if (jb < N) {
      const float x = __ldg(cloud + (size_t)jb * 3 + 0), y = __ldg(cloud + (size_t)jb * 3 + 1), z = __ldg(cloud + (size_t)jb * 3 + 2);
      xb = -2.0f * x; yb = -2.0f * y; zb = -2.0f * z; pb = ks_sumsq(x, y, z);
    }
    spa[e] = make_float4(xa, xb, ya, yb);
    spb[e] = make_float4(za, zb, pa, pb);
  }
  __syncthreads();
  const int s_begin = blockIdx.x * qpc;
  const int s_end = min(S, s_begin + qpc);
  const int nsteps = (N + 63) / 64;  // pair steps that contain real candidates

  for (int s0 = s_begin + warp * QW; s0 < s_end; s0 += KS_WARPS * QW) {
    float qx[QW], qy[QW], qz[QW], qq[QW], lmin[QW], T[QW];
    int cnt[QW];
#pragma unroll
    for (int w = 0; w < QW; w++) {
      const int s = min(s0 + w, s_end - 1);  // tail: duplicate the last query, its result is not stored
      const float* qp = new_xyz + ((size_t)b * S + s) * 3;
      qx[w] = __ldg(qp + 0); qy[w] = __ldg(qp + 1); qz[w] = __ldg(qp + 2);
      qq[w] = ks_sumsq(qx[w], qy[w], qz[w]);
      lmin[w] = INF;
      cnt[w] = 0;
    }
    // 1. lane minima (distances are not kept: 64 registers per query would cost the occupancy)
#pragma unroll 2
    for (int u = 0; u < nsteps; u++) {
      const KsPair c = ks_load(spa, spb, u * 32 + lane);
#pragma unroll
      for (int w = 0; w < QW; w++) {
        const u64 d = ks_dist2<VAR>(pack2(qx[w], qx[w]), pack2(qy[w], qy[w]), pack2(qz[w], qz[w]), pack2(qq[w], qq[w]), c);
        lmin[w] = fminf(lmin[w], fminf(lo2(d), hi2(d)));
      }
    }
    // 2. thresholds = KK-th smallest lane minimum of each query
#pragma unroll
    for (int w = 0; w < QW; w++) T[w] = __shfl_sync(0xffffffffu, ks_sort_values(lmin[w], lane), KK - 1);
    // 3. re-evaluate and compact {d <= T} in index order (first the 32 candidates 64u + lane, then 64u + 32 + lane)
#pragma unroll 2
    for (int u = 0; u < nsteps; u++) {
      const KsPair c = ks_load(spa, spb, u * 32 + lane);
#pragma unroll
      for (int w = 0; w < QW; w++) {
        const u64 d = ks_dist2<VAR>(pack2(qx[w], qx[w]), pack2(qy[w], qy[w]), pack2(qz[w], qz[w]), pack2(qq[w], qq[w]), c);
        const float d0 = lo2(d), d1 = hi2(d);
        const bool p0 = d0 <= T[w], p1 = d1 <= T[w];
        if (__any_sync(0xffffffffu, p0 || p1)) {
          const unsigned m0 = __ballot_sync(0xffffffffu, p0);
          const unsigned m1 = __ballot_sync(0xffffffffu, p1);
          const unsigned below = (1u << lane) - 1u;
          const int pos0 = cnt[w] + __popc(m0 & below);
          const int pos1 = cnt[w] + __popc(m0) + __popc(m1 & below);
          if (p0 && pos0 < 32) { bufd[warp][w][pos0] = d0; bufi[warp][w][pos0] = u * 64 + lane; }
          if (p1 && pos1 < 32) { bufd[warp][w][pos1] = d1; bufi[warp][w][pos1] = u * 64 + 32 + lane; }
          cnt[w] += __popc(m0) + __popc(m1);
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int w = 0; w < QW; w++) {
      float d = INF;
      int ci = 0x7fffffff;
      if (cnt[w] <= 32) {
        if (lane < cnt[w]) { d = bufd[warp][w][lane]; ci = bufi[warp][w][lane]; }
        warp_sort_pairs(d, ci, lane);
      } else {
        // 4. rare: more than 32 candidates at or below the threshold (many equal distances):
        // streaming insertion into a sorted warp list, candidates in index order
        float thr = INF;
        for (int j0 = 0; j0 < N; j0 += 32) {
          float dj = INF;
          if (j0 + lane < N) dj = ks_dist1<VAR>(qx[w], qy[w], qz[w], qq[w], ks_cand(spa, spb, j0 + lane));
          unsigned mask = __ballot_sync(0xffffffffu, dj < thr);
          while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const float cd = __shfl_sync(0xffffffffu, dj, src);
            const bool ok = cd < thr;
            const float ud = __shfl_up_sync(0xffffffffu, d, 1);
            const int ui = __shfl_up_sync(0xffffffffu, ci, 1);
            const bool shift = ok && (lane > 0) && (ud > cd);
            const bool ins = ok && (d > cd);
            d = shift ? ud : (ins ? cd : d);
            ci = shift ? ui : (ins ? j0 + src : ci);
            thr = __shfl_sync(0xffffffffu, d, KK - 1);
          }
        }
      }
      if (EPI && order == PS_ORDER_TOPK) warp_torch_topk_order(d, ci, lane, KK);
      const int s = s0 + w;
      if (s < s_end && lane >= skip && lane < KK) {
        idx[((size_t)b * S + s) * k + (lane - skip)] = ci;
        if (EPI && gxyz) {
          // fused grouping of the coordinates + centre subtraction (models/model_utils.py:344-345):
          // grouped_xyz[b,c,s,j] = xyz[b,idx[b,s,j],c] - new_xyz[b,s,c]   (staged coordinates are -2 * xyz: exact to undo)
          const float4 c = ks_cand(spa, spb, ci);
          float* o = gxyz + ((size_t)b * 3 * S + s) * k + (lane - skip);
          o[0] = __fsub_rn(-0.5f * c.x, qx[w]);
          o[(size_t)S * k] = __fsub_rn(-0.5f * c.y, qy[w]);
          o[(size_t)2 * S * k] = __fsub_rn(-0.5f * c.z, qz[w]);
        }
      }
    }
    __syncwarp();  // the buffers are reused by this warp's next group of queries
  }
}

}  // namespace

// PS_OK when handled, 1 when the shape needs the streaming kernel, negative on error.
int knn_select_launch(const float* xyz, const float* new_xyz, int* idx, float* gxyz, int B, int N, int S, int k,
                      int skip, int order, int var, int nsm, cudaStream_t stream) {
  if (N > KS_TILE || k + skip > 16) return 1;
  int qpc = 64;  // queries per CTA (a multiple of 8 warps x 4 queries while it stays >= 32)
  while (qpc > 32 && (long long)B * ceil_div(S, qpc) < (long long)nsm * 4) qpc /= 2;
  const dim3 grid(ceil_div(S, qpc), B);
  constexpr int QW = 4;
#define PS_KS(VARV, EPIV) knn_select_kernel<VARV, QW, EPIV><<<grid, KS_THREADS, 0, stream>>>(xyz, new_xyz, idx, gxyz, N, S, k, skip, qpc, order)
  const bool epi = gxyz != nullptr || order != PS_ORDER_SORT;
  if (var == 1) { if (epi) PS_KS(1, 1); else PS_KS(1, 0); }
  else if (var == 2) { if (epi) PS_KS(2, 1); else PS_KS(2, 0); }
  else { if (epi) PS_KS(0, 1); else PS_KS(0, 0); }
#undef PS_KS
  PS_LAUNCH_CHECK();
  return PS_OK;
}

}  // namespace ps
