// Neighbourhood queries for sm_100a: streaming kNN, ball query, 3-NN and 3-point interpolation.
//
//   ps_knn         replaces the torch expression query_knn/square_distance
//                  (models/model_utils.py:258-286): no (B,S,N) matrix, no full sort.
//   ps_ball_query  replaces query_ball_point_kernel (pointnet2_ops/_ext-src/src/ball_query_gpu.cu:9-44)
//   ps_three_nn / ps_three_interpolate_*  replace interpolate_gpu.cu:9-59, :72-101, :116-143
//
// kNN design: one warp per query.  The candidate cloud is staged through shared memory as SoA
// (x, y, z, |p|^2); each lane evaluates one candidate per step with the reference's *expanded*
// fp32 expression (so the ordering of near-ties matches torch), and the warp keeps the k best
// in a register list distributed over the lanes (entry e lives in lane e%32, register e/32),
// sorted ascending by (distance, index).  A candidate is inserted only when it beats the current
// k-th distance (ballot + shuffle-shift), which happens ~k*ln(N/k) times per query.
#include "common.cuh"
#include "select.cuh"

namespace ps {

constexpr int NB_THREADS = 256;
constexpr int NB_WARPS = NB_THREADS / 32;
constexpr int NB_TILE = 2048;

// |p|^2 as torch.sum(p ** 2, -1) evaluates it: three separately rounded squares combined as (x^2 + z^2) + y^2 — the order torch's CUDA reduction uses for a length-3 row (measured on B200, tests/golden/knn.npz).
__device__ __forceinline__ float sumsq_torch(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(z, z)), __fmul_rn(y, y));
}

// square_distance (models/model_utils.py:276-278): dist = -2*matmul; dist += |q|^2; dist += |p|^2.
// VAR selects the K=3 accumulation order of the fp32 GEMM (decided by measurement on the box,
// see DESIGN.md "kNN arithmetic"): 0 = x,y,z ascending FMA chain, 1 = z,y,x, 2 = unfused.
template <int VAR>
__device__ __forceinline__ float knn_dist(float qx, float qy, float qz, float qq, float px, float py,
                                          float pz, float pp) {
  float dot;
  if (VAR == 0) dot = __fmaf_rn(qz, pz, __fmaf_rn(qy, py, __fmul_rn(qx, px)));
  else if (VAR == 1) dot = __fmaf_rn(qx, px, __fmaf_rn(qy, py, __fmul_rn(qz, pz)));
  else dot = __fadd_rn(__fadd_rn(__fmul_rn(qx, px), __fmul_rn(qy, py)), __fmul_rn(qz, pz));
  return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), qq), pp);
}

// Warp-wide bitonic sort of one (distance, index) pair per lane, ascending by (d, i).
__device__ __forceinline__ void warp_bitonic_sort(float& d, int& i, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, j);
      const int oi = __shfl_xor_sync(0xffffffffu, i, j);
      const bool up = ((lane & k) == 0);          // this k-block sorts ascending
      const bool lower = ((lane & j) == 0);       // this lane keeps the smaller of the pair
      const bool other_less = (od < d) || (od == d && oi < i);
      const bool take = (lower == up) ? other_less : !other_less;
      if (take) { d = od; i = oi; }
    }
  }
}

// One warp per query; candidates staged as float4 {x, y, z, |p|^2}.  List of the KK = k + skip
// best lives in lanes 0..KK-1 (R registers per lane when KK > 32), ascending by (dist, index).
template <int R, int VAR>
__global__ void __launch_bounds__(NB_THREADS) knn_kernel(const float* __restrict__ xyz,
                                                         const float* __restrict__ new_xyz,
                                                         int* __restrict__ idx, float* __restrict__ gxyz,
                                                         int N, int S, int k, int skip, int qpc, int order) {
  __shared__ float4 sp[NB_TILE];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* cloud = xyz + (size_t)b * N * 3;
  const float INF = __int_as_float(0x7f800000);
  const int KK = k + skip;
  const int s_begin = blockIdx.x * qpc;
  const int s_end = min(S, s_begin + qpc);
  const bool single_tile = N <= NB_TILE;

  auto stage = [&](int ts, int cnt) {
    for (int i = tid; i < cnt; i += NB_THREADS) {
      const float x = __ldg(cloud + (size_t)(ts + i) * 3 + 0);
      const float y = __ldg(cloud + (size_t)(ts + i) * 3 + 1);
      const float z = __ldg(cloud + (size_t)(ts + i) * 3 + 2);
      sp[i] = make_float4(x, y, z, sumsq_torch(x, y, z));
    }
  };
  if (single_tile) { stage(0, N); __syncthreads(); }

  for (int s0 = s_begin; s0 < s_end; s0 += NB_WARPS) {
    const int s = s0 + warp;
    const bool valid = s < s_end;
    float qx = 0.f, qy = 0.f, qz = 0.f, qq = 0.f;
    if (valid) {
      const float* qp = new_xyz + ((size_t)b * S + s) * 3;
      qx = __ldg(qp + 0); qy = __ldg(qp + 1); qz = __ldg(qp + 2);
      qq = sumsq_torch(qx, qy, qz);
    }
    float ld[R];
    int li[R];
#pragma unroll
    for (int r = 0; r < R; r++) { ld[r] = INF; li[r] = 0; }
    float thr = INF;

    for (int ts = 0; ts < N; ts += NB_TILE) {
      const int cnt = min(NB_TILE, N - ts);
      if (!single_tile) { __syncthreads(); stage(ts, cnt); __syncthreads(); }
      if (!valid) continue;
      int j0 = 0;
      if (R == 1 && ts == 0) {
        // the first 32 candidates seed the list with one bitonic sort instead of 32 insertions
        float d = INF;
        int ci = lane;
        if (lane < cnt) { const float4 c = sp[lane]; d = knn_dist<VAR>(qx, qy, qz, qq, c.x, c.y, c.z, c.w); }
        warp_bitonic_sort(d, ci, lane);
        ld[0] = d; li[0] = ci;
        thr = __shfl_sync(0xffffffffu, ld[0], (KK - 1) & 31);
        j0 = 32;
      }
      if (R == 1) {
        // two candidates per lane per step (indices j0+lane and j0+32+lane); insertion is
        // branch-free: entries > cd move one lane up, the first of them takes the candidate; equal
        // distances stay in front (they have lower indices: candidates arrive in index order)
        for (; j0 < cnt; j0 += 64) {
          float d0 = INF, d1 = INF;
          if (j0 + lane < cnt) { const float4 c = sp[j0 + lane]; d0 = knn_dist<VAR>(qx, qy, qz, qq, c.x, c.y, c.z, c.w); }
          if (j0 + 32 + lane < cnt) { const float4 c = sp[j0 + 32 + lane]; d1 = knn_dist<VAR>(qx, qy, qz, qq, c.x, c.y, c.z, c.w); }
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const float d = h ? d1 : d0;
            unsigned mask = __ballot_sync(0xffffffffu, d < thr);
            while (mask) {
              const int src = __ffs(mask) - 1;
              mask &= mask - 1;
              const float cd = __shfl_sync(0xffffffffu, d, src);
              const int ci = ts + j0 + 32 * h + src;
              const bool ok = cd < thr;  // warp-uniform; the threshold may have tightened in this group
              const float ud = __shfl_up_sync(0xffffffffu, ld[0], 1);
              const int ui = __shfl_up_sync(0xffffffffu, li[0], 1);
              const bool shift = ok && (lane > 0) && (ud > cd);
              const bool ins = ok && (ld[0] > cd);
              ld[0] = shift ? ud : (ins ? cd : ld[0]);
              li[0] = shift ? ui : (ins ? ci : li[0]);
              thr = __shfl_sync(0xffffffffu, ld[0], (KK - 1) & 31);
            }
          }
        }
      } else {
        for (; j0 < cnt; j0 += 32) {
          const int j = j0 + lane;
          float d = INF;
          if (j < cnt) { const float4 c = sp[j]; d = knn_dist<VAR>(qx, qy, qz, qq, c.x, c.y, c.z, c.w); }
          unsigned mask = __ballot_sync(0xffffffffu, d < thr);
          while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const float cd = __shfl_sync(0xffffffffu, d, src);
            if (!(cd < thr)) continue;  // the threshold may have tightened inside this group
            const int ci = ts + j0 + src;
            int pos = 0;
#pragma unroll
            for (int r = 0; r < R; r++) pos += __popc(__ballot_sync(0xffffffffu, ld[r] <= cd));
#pragma unroll
            for (int r = R - 1; r >= 0; r--) {
              float ud = __shfl_up_sync(0xffffffffu, ld[r], 1);
              int ui = __shfl_up_sync(0xffffffffu, li[r], 1);
              if (r > 0) {
                const float pd = __shfl_sync(0xffffffffu, ld[r - 1], 31);
                const int pi = __shfl_sync(0xffffffffu, li[r - 1], 31);
                if (lane == 0) { ud = pd; ui = pi; }
              }
              const int e = r * 32 + lane;
              if (e > pos) { ld[r] = ud; li[r] = ui; }
              else if (e == pos) { ld[r] = cd; li[r] = ci; }
            }
            float tv = ld[0];
#pragma unroll
            for (int r = 1; r < R; r++)
              if (r == ((KK - 1) >> 5)) tv = ld[r];
            thr = __shfl_sync(0xffffffffu, tv, (KK - 1) & 31);
          }
        }
      }
    }
    if (valid) {
      if (R == 1 && order == PS_ORDER_TOPK) warp_torch_topk_order(ld[0], li[0], lane, KK);
      int* o = idx + ((size_t)b * S + s) * k;
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int e = r * 32 + lane;
        if (e >= skip && e < KK) {
          o[e - skip] = li[r];
          if (gxyz) {  // fused coordinate grouping + centre subtraction (models/model_utils.py:344-345)
            const float* cp = cloud + (size_t)li[r] * 3;
            float* g = gxyz + ((size_t)b * 3 * S + s) * k + (e - skip);
            g[0] = __fsub_rn(__ldg(cp + 0), qx);
            g[(size_t)S * k] = __fsub_rn(__ldg(cp + 1), qy);
            g[(size_t)2 * S * k] = __fsub_rn(__ldg(cp + 2), qz);
          }
        }
      }
    }
  }
}

// ---- ball query: one warp per centre ---------------------------------------------------------
__global__ void __launch_bounds__(NB_THREADS) ball_query_kernel(const float* __restrict__ new_xyz,
                                                                const float* __restrict__ xyz,
                                                                int* __restrict__ idx, int N, int S,
                                                                float radius2, int nsample, int qpc) {
  __shared__ float sx[NB_TILE], sy[NB_TILE], sz[NB_TILE];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* cloud = xyz + (size_t)b * N * 3;
  const int s_begin = blockIdx.x * qpc;
  const int s_end = min(S, s_begin + qpc);
  const bool single_tile = N <= NB_TILE;
  auto stage = [&](int ts, int cnt) {
    for (int i = tid; i < cnt; i += NB_THREADS) {
      sx[i] = __ldg(cloud + (size_t)(ts + i) * 3 + 0);
      sy[i] = __ldg(cloud + (size_t)(ts + i) * 3 + 1);
      sz[i] = __ldg(cloud + (size_t)(ts + i) * 3 + 2);
    }
  };
  if (single_tile) { stage(0, N); __syncthreads(); }
  for (int s0 = s_begin; s0 < s_end; s0 += NB_WARPS) {
    const int s = s0 + warp;
    const bool valid = s < s_end;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    int* o = idx + ((size_t)b * S + (valid ? s : 0)) * nsample;
    if (valid) {
      const float* qp = new_xyz + ((size_t)b * S + s) * 3;
      nx = __ldg(qp + 0); ny = __ldg(qp + 1); nz = __ldg(qp + 2);
    }
    int cnt_hits = 0, first = 0;
    for (int ts = 0; ts < N; ts += NB_TILE) {
      const int cnt = min(NB_TILE, N - ts);
      if (!single_tile) { __syncthreads(); stage(ts, cnt); __syncthreads(); }
      if (!valid) continue;
      for (int j0 = 0; j0 < cnt && cnt_hits < nsample; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        if (j < cnt) hit = dist2_ref(nx - sx[j], ny - sy[j], nz - sz[j]) < radius2;
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask) {
          if (cnt_hits == 0) first = ts + j0 + __ffs(mask) - 1;
          const int pos = cnt_hits + __popc(mask & ((1u << lane) - 1u));
          if (hit && pos < nsample) o[pos] = ts + j;
          cnt_hits += __popc(mask);
        }
      }
    }
    if (valid) {
      // pad with the first hit (ball_query_gpu.cu:34-38); all zeros when nothing was in range
      // (the reference's output buffer is torch::zeros, ball_query.cpp:19-21)
      const int fill = cnt_hits > 0 ? first : 0;
      for (int l = min(cnt_hits, nsample) + lane; l < nsample; l += 32) o[l] = fill;
    }
  }
}

// ---- three_nn: one thread per unknown point, known cloud tiled through shared memory ------------
__global__ void __launch_bounds__(NB_THREADS) three_nn_kernel(const float* __restrict__ unknown,
                                                              const float* __restrict__ known,
                                                              float* __restrict__ dist2,
                                                              int* __restrict__ idx, int n, int m) {
  __shared__ float sx[NB_TILE], sy[NB_TILE], sz[NB_TILE];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int j = blockIdx.x * NB_THREADS + tid;
  const bool valid = j < n;
  const float* kc = known + (size_t)b * m * 3;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (valid) {
    const float* up = unknown + ((size_t)b * n + j) * 3;
    ux = __ldg(up + 0); uy = __ldg(up + 1); uz = __ldg(up + 2);
  }
  // interpolate_gpu.cu:27 keeps the running bests in double initialised to 1e40; every value
  // ever stored is an fp32 distance, so fp32 with +inf start compares identically and the
  // final (float)1e40 is +inf as well.
  const float INF = __int_as_float(0x7f800000);
  float b1 = INF, b2 = INF, b3 = INF;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int ts = 0; ts < m; ts += NB_TILE) {
    const int cnt = min(NB_TILE, m - ts);
    __syncthreads();
    for (int i = tid; i < cnt; i += NB_THREADS) {
      sx[i] = __ldg(kc + (size_t)(ts + i) * 3 + 0);
      sy[i] = __ldg(kc + (size_t)(ts + i) * 3 + 1);
      sz[i] = __ldg(kc + (size_t)(ts + i) * 3 + 2);
    }
    __syncthreads();
    if (!valid) continue;
#pragma unroll 4
    for (int k = 0; k < cnt; k++) {
      const float d = dist2_ref(ux - sx[k], uy - sy[k], uz - sz[k]);
      const int kk = ts + k;
      if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk; }
      else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = kk; }
      else if (d < b3) { b3 = d; i3 = kk; }
    }
  }
  if (valid) {
    float* od = dist2 + ((size_t)b * n + j) * 3;
    int* oi = idx + ((size_t)b * n + j) * 3;
    od[0] = b1; od[1] = b2; od[2] = b3;
    oi[0] = i1; oi[1] = i2; oi[2] = i3;
  }
}

// out[b,c,j] = p1*w1 + p2*w2 + p3*w3 contracted as nvcc does for the reference
// (interpolate_gpu.cu:97-98 -> SASS: FMUL(p2,w2), FFMA(p1,w1,.), FFMA(p3,w3,.))
template <int CT>
__global__ void __launch_bounds__(NB_THREADS) three_interp_fwd_kernel(
    const float* __restrict__ points, const int* __restrict__ idx, const float* __restrict__ weight,
    float* __restrict__ out, int C, int m, int n) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT;
  const int j = blockIdx.x * NB_THREADS + threadIdx.x;
  if (j >= n) return;
  const int* ip = idx + ((size_t)b * n + j) * 3;
  const float* wp = weight + ((size_t)b * n + j) * 3;
  const int i1 = __ldg(ip + 0), i2 = __ldg(ip + 1), i3 = __ldg(ip + 2);
  const float w1 = __ldg(wp + 0), w2 = __ldg(wp + 1), w3 = __ldg(wp + 2);
#pragma unroll
  for (int r = 0; r < CT; r++) {
    const int c = c0 + r;
    if (c < C) {
      const float* row = points + ((size_t)b * C + c) * m;
      out[((size_t)b * C + c) * n + j] =
          __fmaf_rn(__ldg(row + i3), w3, __fmaf_rn(__ldg(row + i1), w1, __fmul_rn(__ldg(row + i2), w2)));
    }
  }
}

template <int CT>
__global__ void __launch_bounds__(NB_THREADS) three_interp_bwd_kernel(
    const float* __restrict__ gout, const int* __restrict__ idx, const float* __restrict__ weight,
    float* __restrict__ gpoints, int C, int n, int m) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT;
  const int j = blockIdx.x * NB_THREADS + threadIdx.x;
  if (j >= n) return;
  const int* ip = idx + ((size_t)b * n + j) * 3;
  const float* wp = weight + ((size_t)b * n + j) * 3;
  const int i1 = __ldg(ip + 0), i2 = __ldg(ip + 1), i3 = __ldg(ip + 2);
  const float w1 = __ldg(wp + 0), w2 = __ldg(wp + 1), w3 = __ldg(wp + 2);
#pragma unroll
  for (int r = 0; r < CT; r++) {
    const int c = c0 + r;
    if (c < C) {
      const float g = __ldg(gout + ((size_t)b * C + c) * n + j);
      float* row = gpoints + ((size_t)b * C + c) * m;
      atomicAdd(row + i1, __fmul_rn(g, w1));
      atomicAdd(row + i2, __fmul_rn(g, w2));
      atomicAdd(row + i3, __fmul_rn(g, w3));
    }
  }
}

static int knn_variant() {
  if (const char* e = getenv("PS_KNN_VARIANT")) return atoi(e);
  return 0;
}

template <int R>
static void launch_knn(int var, dim3 grid, cudaStream_t st, const float* xyz, const float* new_xyz,
                       int* idx, float* gxyz, int N, int S, int k, int skip, int qpc, int order) {
  if (var == 1) knn_kernel<R, 1><<<grid, NB_THREADS, 0, st>>>(xyz, new_xyz, idx, gxyz, N, S, k, skip, qpc, order);
  else if (var == 2) knn_kernel<R, 2><<<grid, NB_THREADS, 0, st>>>(xyz, new_xyz, idx, gxyz, N, S, k, skip, qpc, order);
  else knn_kernel<R, 0><<<grid, NB_THREADS, 0, st>>>(xyz, new_xyz, idx, gxyz, N, S, k, skip, qpc, order);
}

}  // namespace ps

using namespace ps;

static int knn_impl(const float* xyz, const float* new_xyz, int* idx, float* gxyz, int B, int N, int S, int k,
                    int skip, int order, int dev, void* stream_, const char* who) {
  PS_REQUIRE(B >= 0 && N > 0 && S >= 0 && k > 0 && skip >= 0, "%s: bad sizes B=%d N=%d S=%d k=%d skip=%d", who, B, N, S, k, skip);
  PS_REQUIRE(k + skip <= N, "%s: k+skip=%d exceeds the number of points N=%d", who, k + skip, N);
  PS_REQUIRE(order == PS_ORDER_SORT || order == PS_ORDER_TOPK, "%s: unknown result order %d", who, order);
  if (k + skip > 128) return set_error(PS_ERR_UNSUPPORTED, "%s: k+skip=%d > 128 not supported", who, k + skip);
  if (order == PS_ORDER_TOPK && (k > 32 || skip != 0))
    return set_error(PS_ERR_UNSUPPORTED, "%s: torch.topk result order is reproduced for k <= 32, skip == 0 (k=%d skip=%d)", who, k, skip);
  if (B == 0 || S == 0) return PS_OK;
  PS_REQUIRE(xyz && new_xyz && idx, "%s: null pointer", who);
  PS_REQUIRE(B <= 65535, "%s: B=%d exceeds the grid y limit", who, B);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "%s: cannot select device %d", who, dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nsm = sm_count(dev);
  // queries per CTA: a multiple of 8 (one per warp per round); enough CTAs for ~4 per SM
  int qpc = 64;
  while (qpc > 8 && (long long)B * ceil_div(S, qpc) < (long long)nsm * 4) qpc /= 2;
  const dim3 grid(ceil_div(S, qpc), B);
  const int KK = k + skip;
  const int var = knn_variant();
  {
    const char* e = getenv("PS_KNN_SELECT");  // PS_KNN_SELECT=0 forces the streaming kernel (A/B, tests)
    if (!(e && atoi(e) == 0)) {
      const int rc = knn_select_launch(xyz, new_xyz, idx, gxyz, B, N, S, k, skip, order, var, nsm, stream);
      if (rc <= 0) return rc;
    }
  }
  if (KK <= 32) launch_knn<1>(var, grid, stream, xyz, new_xyz, idx, gxyz, N, S, k, skip, qpc, order);
  else if (KK <= 64) launch_knn<2>(var, grid, stream, xyz, new_xyz, idx, gxyz, N, S, k, skip, qpc, order);
  else launch_knn<4>(var, grid, stream, xyz, new_xyz, idx, gxyz, N, S, k, skip, qpc, order);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

extern "C" int ps_knn(const float* xyz, const float* new_xyz, int* idx, int B, int N, int S, int k,
                      int skip, int dev, void* stream) {
  return knn_impl(xyz, new_xyz, idx, nullptr, B, N, S, k, skip, PS_ORDER_SORT, dev, stream, "ps_knn");
}

extern "C" int ps_knn_point(const float* xyz, const float* new_xyz, int* idx, int B, int N, int S, int k,
                            int dev, void* stream) {
  return knn_impl(xyz, new_xyz, idx, nullptr, B, N, S, k, 0, PS_ORDER_TOPK, dev, stream, "ps_knn_point");
}

extern "C" int ps_knn_group_xyz(const float* xyz, const float* new_xyz, int* idx, float* grouped_xyz, int B,
                                int N, int S, int k, int dev, void* stream) {
  PS_REQUIRE(grouped_xyz || B == 0 || S == 0, "ps_knn_group_xyz: null grouped_xyz");
  return knn_impl(xyz, new_xyz, idx, grouped_xyz, B, N, S, k, 0, PS_ORDER_SORT, dev, stream, "ps_knn_group_xyz");
}

extern "C" int ps_ball_query(const float* new_xyz, const float* xyz, int* idx, int B, int N, int S,
                             float radius, int nsample, int dev, void* stream_) {
  PS_REQUIRE(B >= 0 && N > 0 && S >= 0 && nsample >= 0, "ps_ball_query: bad sizes B=%d N=%d S=%d nsample=%d", B, N, S, nsample);
  if (B == 0 || S == 0 || nsample == 0) return PS_OK;
  PS_REQUIRE(xyz && new_xyz && idx, "ps_ball_query: null pointer");
  PS_REQUIRE(B <= 65535, "ps_ball_query: B=%d exceeds the grid y limit", B);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_ball_query: cannot select device %d", dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nsm = sm_count(dev);
  int qpc = 64;
  while (qpc > 8 && (long long)B * ceil_div(S, qpc) < (long long)nsm * 4) qpc /= 2;
  const float radius2 = radius * radius;  // fp32 product, as ball_query_gpu.cu:22
  ball_query_kernel<<<dim3(ceil_div(S, qpc), B), NB_THREADS, 0, stream>>>(new_xyz, xyz, idx, N, S, radius2, nsample, qpc);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

extern "C" int ps_three_nn(const float* unknown, const float* known, float* dist2, int* idx, int B,
                           int n, int m, int dev, void* stream_) {
  PS_REQUIRE(B >= 0 && n >= 0 && m > 0, "ps_three_nn: bad sizes B=%d n=%d m=%d", B, n, m);
  if (B == 0 || n == 0) return PS_OK;
  PS_REQUIRE(unknown && known && dist2 && idx, "ps_three_nn: null pointer");
  PS_REQUIRE(B <= 65535, "ps_three_nn: B=%d exceeds the grid y limit", B);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_three_nn: cannot select device %d", dev);
  three_nn_kernel<<<dim3(ceil_div(n, NB_THREADS), B), NB_THREADS, 0, (cudaStream_t)stream_>>>(unknown, known, dist2, idx, n, m);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

extern "C" int ps_three_interpolate_fwd(const float* points, const int* idx, const float* weight,
                                        float* out, int B, int C, int m, int n, int dev, void* stream_) {
  PS_REQUIRE(B >= 0 && C >= 0 && m > 0 && n >= 0, "ps_three_interpolate_fwd: bad sizes B=%d C=%d m=%d n=%d", B, C, m, n);
  if (B == 0 || C == 0 || n == 0) return PS_OK;
  PS_REQUIRE(points && idx && weight && out, "ps_three_interpolate_fwd: null pointer");
  PS_REQUIRE(B <= 65535, "ps_three_interpolate_fwd: B=%d exceeds the grid z limit", B);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_three_interpolate_fwd: cannot select device %d", dev);
  constexpr int CT = 8;
  three_interp_fwd_kernel<CT><<<dim3(ceil_div(n, NB_THREADS), ceil_div(C, CT), B), NB_THREADS, 0, (cudaStream_t)stream_>>>(points, idx, weight, out, C, m, n);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

extern "C" int ps_three_interpolate_bwd(const float* grad_out, const int* idx, const float* weight,
                                        float* grad_points, int B, int C, int n, int m, int dev, void* stream_) {
  PS_REQUIRE(B >= 0 && C >= 0 && m > 0 && n >= 0, "ps_three_interpolate_bwd: bad sizes B=%d C=%d n=%d m=%d", B, C, n, m);
  if (B == 0 || C == 0) return PS_OK;
  PS_REQUIRE(grad_points && (n == 0 || (grad_out && idx && weight)), "ps_three_interpolate_bwd: null pointer");
  PS_REQUIRE(B <= 65535, "ps_three_interpolate_bwd: B=%d exceeds the grid z limit", B);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_three_interpolate_bwd: cannot select device %d", dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  PS_CUDA(cudaMemsetAsync(grad_points, 0, (size_t)B * C * m * sizeof(float), stream));
  if (n == 0) return PS_OK;
  constexpr int CT = 8;
  three_interp_bwd_kernel<CT><<<dim3(ceil_div(n, NB_THREADS), ceil_div(C, CT), B), NB_THREADS, 0, stream>>>(grad_out, idx, weight, grad_points, C, n, m);
  PS_LAUNCH_CHECK();
  return PS_OK;
}
