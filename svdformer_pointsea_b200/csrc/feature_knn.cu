// Feature-space kNN for sm_100a: query_knn_point / group_local / EdgeConv neighbourhoods on C-dimensional
// features (models/model_utils.py:258-279 square_distance, :807-826; EdgeConv(64,256,8), EdgeConv(256,512,4)
// in models/SVDFormer.py:171-172 and models_PointSea/PointSea.py:234-236).
//
// The reference materialises dist = -2 * (Q @ R^T) + |q|^2 + |r|^2 as a (B,S,N) fp32 matrix through cuBLAS and
// two torch reductions, then runs torch.topk on every row.  Here one CTA owns a tile of TQ queries of one
// cloud: the TQ x N distance block is produced in shared memory by an fp32 FMA tile loop and each row is
// reduced to its k smallest by one warp (threshold selection as in knn_select.cu) — nothing of size S*N
// ever reaches HBM.
//
// Arithmetic (measured bit-exactly on B200 against torch 2.11 / cuBLAS, tools + tests/golden/next.npz):
//   dot   = ascending-channel FMA chain: acc = fma(q_c, r_c, acc), c = 0..C-1  (cuBLAS fp32 SIMT GEMM order)
//   |x|^2 = torch.sum(x ** 2, -1) in the order of ATen's reduce kernel (Reduce.cuh): squares rounded
//           separately; "thread" t of a row owns elements t, t+bw, ... (or float4 chunks t, t+bw, ... when
//           C >= 128) in 4 interleaved accumulators, combined ((a0+a1)+a2)+a3, then a shuffle-down tree
//           over bw = min(pow2floor(C or C/4), 32) threads.  C = 3 gives (x^2 + z^2) + y^2, the order
//           found earlier for the coordinate kNN.
//   dist  = ((-2 * dot) + |q|^2) + |r|^2
// Result order: PS_ORDER_SORT or PS_ORDER_TOPK (select.cuh).
#include "common.cuh"
#include "select.cuh"

#include <cstdlib>

namespace ps {
namespace {

constexpr int FK_THREADS = 256;
constexpr int FK_WARPS = FK_THREADS / 32;
constexpr int FK_KC = 16;   // channels per shared-memory stage
constexpr int FK_TR = 256;  // reference points per accumulator pass (each lane: 2 x 4 consecutive)

// One warp per row; lane t plays torch's reduce thread t (block width bw <= 32).
__global__ void __launch_bounds__(FK_THREADS) rowsumsq_torch_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                    int C, int N, long long sb, long long sn, long long sc,
                                                                    int vec, int bw, long long rows) {
  const long long row = (long long)blockIdx.x * FK_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(row / N), n = (int)(row % N);
  const float* p = x + (size_t)b * sb + (size_t)n * sn;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (lane < bw) {
    if (vec) {
      for (int c = lane; c < C / 4; c += bw) {
        const float v0 = __ldg(p + (size_t)(4 * c + 0) * sc), v1 = __ldg(p + (size_t)(4 * c + 1) * sc);
        const float v2 = __ldg(p + (size_t)(4 * c + 2) * sc), v3 = __ldg(p + (size_t)(4 * c + 3) * sc);
        a0 = __fadd_rn(a0, __fmul_rn(v0, v0));
        a1 = __fadd_rn(a1, __fmul_rn(v1, v1));
        a2 = __fadd_rn(a2, __fmul_rn(v2, v2));
        a3 = __fadd_rn(a3, __fmul_rn(v3, v3));
      }
    } else {
      int j = 0;
      for (int e = lane; e < C; e += bw, j++) {
        const float v = __ldg(p + (size_t)e * sc);
        const float s = __fmul_rn(v, v);
        const int a = j & 3;
        if (a == 0) a0 = __fadd_rn(a0, s);
        else if (a == 1) a1 = __fadd_rn(a1, s);
        else if (a == 2) a2 = __fadd_rn(a2, s);
        else a3 = __fadd_rn(a3, s);
      }
    }
  }
  float v = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
  for (int off = bw >> 1; off > 0; off >>= 1) v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, off));
  if (lane == 0) out[row] = v;
}

struct FkArgs {
  const float* xq; const float* xr;    // queries / references
  const float* qq; const float* pp;    // their torch-order squared norms, (B,S) and (B,N)
  int* idx;                            // (B,S,k)
  int C, N, S, k, order, npad;
  long long sbq, snq, scq, sbr, snr, scr;
};

// values only, ascending
__device__ __forceinline__ float fk_sort_values(float v, int lane) {
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

// One warp reduces one row of npad distances (shared memory, padding at +inf) to its k smallest and writes them
// in the requested order.
__device__ __forceinline__ void fk_select_row(const FkArgs& a, const float* row, int b, int q, int npad, float* bufd, int* bufi) {
  const int lane = threadIdx.x & 31;
  const float INF = __int_as_float(0x7f800000);
  const int k = a.k, N = a.N;
  const int nsteps = npad / 32;
  float lmin = INF;
  for (int u = 0; u < nsteps; u++) lmin = fminf(lmin, row[u * 32 + lane]);
  const float T = __shfl_sync(0xffffffffu, fk_sort_values(lmin, lane), k - 1);
  int cnt = 0;
  for (int u = 0; u < nsteps; u++) {
    const float dj = row[u * 32 + lane];
    const bool pass = dj <= T;
    const unsigned mask = __ballot_sync(0xffffffffu, pass);
    if (mask) {
      const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
      if (pass && pos < 32) { bufd[pos] = dj; bufi[pos] = u * 32 + lane; }
      cnt += __popc(mask);
    }
  }
  __syncwarp();
  float d = INF;
  int ci = 0x7fffffff;
  if (cnt <= 32) {
    if (lane < cnt) { d = bufd[lane]; ci = bufi[lane]; }
    warp_sort_pairs(d, ci, lane);
  } else {
    // many equal distances: streaming insertion into a sorted warp list, candidates in index order
    float thr = INF;
    for (int j0 = 0; j0 < N; j0 += 32) {
      const float dj = (j0 + lane < N) ? row[j0 + lane] : INF;
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
        thr = __shfl_sync(0xffffffffu, d, k - 1);
      }
    }
  }
  if (a.order == PS_ORDER_TOPK) warp_torch_topk_order(d, ci, lane, k);
  if (lane < k) a.idx[((size_t)b * a.S + q) * k + lane] = ci;
  __syncwarp();
}

template <int TQ>
__global__ void __launch_bounds__(FK_THREADS) knn_feat_kernel(const FkArgs a) {
  constexpr int QPT = TQ / FK_WARPS;  // queries per thread in the tile loop = rows per warp in the selection
  constexpr int NQ = FK_KC * TQ / FK_THREADS;     // staged query values per thread (TQ=32: 2, 16: 1, 8: 0.5 -> 1)
  constexpr int NQR = NQ > 0 ? NQ : 1;
  constexpr int NR = FK_KC * FK_TR / FK_THREADS;  // staged reference values per thread
  extern __shared__ __align__(16) float smem[];
  float* dist = smem;                          // TQ x npad
  float* qs = dist + (size_t)TQ * a.npad;      // FK_KC x TQ
  float* rs = qs + FK_KC * TQ;                 // FK_KC x FK_TR
  __shared__ float bufd[FK_WARPS][32];
  __shared__ int bufi[FK_WARPS][32];
  const int b = blockIdx.y, q0 = blockIdx.x * TQ, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float INF = __int_as_float(0x7f800000);
  const float* xq = a.xq + (size_t)b * a.sbq;
  const float* xr = a.xr + (size_t)b * a.sbr;
  const int C = a.C, N = a.N, S = a.S, npad = a.npad;

  float qn[QPT];
#pragma unroll
  for (int i = 0; i < QPT; i++) {
    const int q = q0 + warp * QPT + i;
    qn[i] = q < S ? __ldg(a.qq + (size_t)b * S + q) : 0.f;
  }

  // global -> register prefetch of one (FK_KC channels) x (TQ queries | FK_TR references) stage; the loads
  // of stage s+1 are in flight while stage s is consumed from shared memory
  float pq[NQR], pr[NR];
  auto fetch = [&](int r0, int c0) {
#pragma unroll
    for (int j = 0; j < NQR; j++) {
      const int i = tid + j * FK_THREADS;
      const int kc = i / TQ, q = i % TQ;
      pq[j] = (i < FK_KC * TQ && c0 + kc < C && q0 + q < S) ? __ldg(xq + (size_t)(q0 + q) * a.snq + (size_t)(c0 + kc) * a.scq) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < NR; j++) {
      const int i = tid + j * FK_THREADS;
      const int kc = i / FK_TR, r = i % FK_TR;
      pr[j] = (c0 + kc < C && r0 + r < N) ? __ldg(xr + (size_t)(r0 + r) * a.snr + (size_t)(c0 + kc) * a.scr) : 0.f;
    }
  };
  auto commit = [&]() {
#pragma unroll
    for (int j = 0; j < NQR; j++) {
      const int i = tid + j * FK_THREADS;
      if (i < FK_KC * TQ) qs[i] = pq[j];  // qs[kc * TQ + q], i = kc * TQ + q
    }
#pragma unroll
    for (int j = 0; j < NR; j++) rs[tid + j * FK_THREADS] = pr[j];  // rs[kc * FK_TR + r]
  };

  const int nstage = (C + FK_KC - 1) / FK_KC;
  for (int r0 = 0; r0 < npad; r0 += FK_TR) {
    // accumulators as packed pairs of references: one FFMA2 per (query, 2 references, channel); each half rounds
    // like the scalar FFMA, so the ascending-channel chain is unchanged
    u64 acc2[QPT][4];
#pragma unroll
    for (int i = 0; i < QPT; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc2[i][j] = 0ull;
    fetch(r0, 0);
    for (int st = 0; st < nstage; st++) {
      const int c0 = st * FK_KC;
      const int kcn = min(FK_KC, C - c0);
      __syncthreads();  // the previous stage has been consumed
      commit();
      __syncthreads();
      if (st + 1 < nstage) fetch(r0, c0 + FK_KC);
      // ascending-channel FMA chain per (query, reference) pair; a partial last stage stops at kcn so that
      // no padded product enters the chain
      auto step = [&](int kc) {
        const ulonglong2 r0v = *reinterpret_cast<const ulonglong2*>(&rs[kc * FK_TR + lane * 4]);
        const ulonglong2 r1v = *reinterpret_cast<const ulonglong2*>(&rs[kc * FK_TR + 128 + lane * 4]);
        float qv[QPT];
        if (QPT == 4) {
          const float4 t = *reinterpret_cast<const float4*>(&qs[kc * TQ + warp * 4]);
          qv[0] = t.x; qv[1 % QPT] = t.y; qv[2 % QPT] = t.z; qv[3 % QPT] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < QPT; i++) qv[i] = qs[kc * TQ + warp * QPT + i];
        }
#pragma unroll
        for (int i = 0; i < QPT; i++) {
          const u64 qd = pack2(qv[i], qv[i]);
          acc2[i][0] = fma2(qd, r0v.x, acc2[i][0]);
          acc2[i][1] = fma2(qd, r0v.y, acc2[i][1]);
          acc2[i][2] = fma2(qd, r1v.x, acc2[i][2]);
          acc2[i][3] = fma2(qd, r1v.y, acc2[i][3]);
        }
      };
      if (kcn == FK_KC) {
#pragma unroll
        for (int kc = 0; kc < FK_KC; kc++) step(kc);
      } else {
        for (int kc = 0; kc < kcn; kc++) step(kc);
      }
    }
    float acc[QPT][8];
#pragma unroll
    for (int i = 0; i < QPT; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) { acc[i][2 * j] = lo2(acc2[i][j]); acc[i][2 * j + 1] = hi2(acc2[i][j]); }
    // dist = ((-2 * dot) + |q|^2) + |r|^2 ; padding columns at +inf
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int rb = r0 + h * 128 + lane * 4;
      float pn[4];
#pragma unroll
      for (int j = 0; j < 4; j++) pn[j] = rb + j < N ? __ldg(a.pp + (size_t)b * N + rb + j) : INF;
#pragma unroll
      for (int i = 0; i < QPT; i++) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          o[j] = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, acc[i][h * 4 + j]), qn[i]), pn[j]);
          if (rb + j >= N) o[j] = INF;
        }
        *reinterpret_cast<float4*>(&dist[(size_t)(warp * QPT + i) * npad + rb]) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  __syncwarp();  // each warp selects from the rows it wrote itself

  // ---- per-row selection of the k smallest (one warp per row, rows live in shared memory) ----------
#pragma unroll 1
  for (int i = 0; i < QPT; i++) {
    const int q = q0 + warp * QPT + i;
    if (q >= S) break;  // warp-uniform
    fk_select_row(a, dist + (size_t)(warp * QPT + i) * npad, b, q, npad, bufd[warp], bufi[warp]);
  }
}

// ---- 8 x 8 register tiles, cp.async double buffering (TQ = 32, N <= 1024) ---------------------------------------
// knn_feat_kernel<32> above is bound by SHARED-MEMORY BANDWIDTH, not by the FP32 pipe (ncu, B=32 C=256 N=512:
// lsu wavefronts 57 %, fma 52 %, short_scoreboard the first stall): its 4 x 8 tile reads 12 floats per 32 FMAs, and
// the SM moves one 128-byte wavefront per cycle for all four schedulers.  Here a warp owns 8 queries x 256
// references of a 512-reference pass (8 x 8 per lane): 16 floats per 64 FMAs, 10 wavefronts per 64 FP32-pipe
// cycles, so the LSU stays below the FMA time.  Stages of 8 channels arrive by cp.async into two shared-memory
// buffers (no prefetch registers), ONE barrier per stage.  Same ascending-channel FMA chain, same results.
constexpr int F8_KC = 8;
constexpr int F8_TR = 512;
constexpr int F8_TQ = 32;

__device__ __forceinline__ void cp_async4(unsigned dst, const float* src, bool valid) {
  const int n = valid ? 4 : 0;  // src-size 0: zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned dst, const float* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

// KC channels per stage.  UNION (single-pass shapes, npad == F8_TR): the distance block overlays the stage buffers —
// it is only written after the last stage has been consumed — so KC = 16 fits two CTAs per SM (70 KB each) and the
// per-stage costs (barrier, copy issue, the accumulator moves ptxas places at the loop's back edge: 35 IMAD.MOV on
// the FP32 pipe per stage) are paid half as often.  The copy addresses are loop constants kept in registers.
template <int VEC, int KC, bool UNION>
__global__ void __launch_bounds__(FK_THREADS, 2) knn_feat8_kernel(const FkArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* dist = smem;                                                   // 32 x npad
  float* qs = UNION ? smem : dist + (size_t)F8_TQ * a.npad;             // 2 x KC x 32
  float* rs = qs + 2 * KC * F8_TQ;                                      // 2 x KC x F8_TR
  __shared__ float bufd[FK_WARPS][32];
  __shared__ int bufi[FK_WARPS][32];
  const int b = blockIdx.y, q0 = blockIdx.x * F8_TQ, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qg = warp & 3, rh = warp >> 2;  // this warp: queries qg*8..+8, reference half rh of the pass
  const float INF = __int_as_float(0x7f800000);
  const float* xq = a.xq + (size_t)b * a.sbq;
  const float* xr = a.xr + (size_t)b * a.sbr;
  const int C = a.C, N = a.N, S = a.S, npad = a.npad;

  float qn[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int q = q0 + qg * 8 + i;
    qn[i] = q < S ? __ldg(a.qq + (size_t)b * S + q) : 0.f;
  }
  // copy plan of this thread (constant over the stages): references as KC x F8_TR, queries as KC x 32
  constexpr int NRV = KC * F8_TR / 4 / FK_THREADS;  // 16-byte reference copies per thread and stage (VEC)
  constexpr int NRS = KC * F8_TR / FK_THREADS;      // 4-byte copies otherwise
  constexpr int NQC = KC * F8_TQ / FK_THREADS;      // query copies
  const int rv_kc = tid / (F8_TR / 4), rv_r = (tid % (F8_TR / 4)) * 4;  // copy j: channel rv_kc + j * (FK_THREADS * 4 / F8_TR)
  const int rs_kc = tid / F8_TR, rs_r = tid % F8_TR;                    // only when FK_THREADS >= F8_TR is false: see below
  const int qc_kc = tid / F8_TQ, qc_q = tid % F8_TQ;                    // copy j: channel qc_kc + j * (FK_THREADS / F8_TQ)
  const bool q_ok = q0 + qc_q < S;
  const float* q_src = xq + (q_ok ? (size_t)(q0 + qc_q) * a.snq : 0);
  const unsigned q_dst = smem_u32(qs + qc_kc * F8_TQ + qc_q);
  const unsigned rv_dst = smem_u32(rs + rv_kc * F8_TR + rv_r);
  auto load = [&](int r0, int c0, int buf) {
    const unsigned boff_r = (unsigned)(buf * KC * F8_TR * 4), boff_q = (unsigned)(buf * KC * F8_TQ * 4);
    if (VEC) {  // references contiguous along r (channel-major tensors, N % 4 == 0, 16-byte aligned rows)
      const bool r_ok = r0 + rv_r < N;
      const float* src = xr + (r_ok ? (size_t)(r0 + rv_r) : 0);
#pragma unroll
      for (int j = 0; j < NRV; j++) {
        const int kc = rv_kc + j * (FK_THREADS * 4 / F8_TR);
        const bool ok = r_ok && c0 + kc < C;
        cp_async16(rv_dst + boff_r + (unsigned)(j * (FK_THREADS * 4 / F8_TR) * F8_TR * 4), src + (ok ? (size_t)(c0 + kc) * a.scr : 0), ok);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NRS; j++) {
        const int i = tid + j * FK_THREADS;
        const int kc = i / F8_TR, r = i % F8_TR;
        const bool ok = c0 + kc < C && r0 + r < N;
        cp_async4(smem_u32(rs + kc * F8_TR + r) + boff_r, xr + (ok ? (size_t)(r0 + r) * a.snr + (size_t)(c0 + kc) * a.scr : 0), ok);
      }
    }
#pragma unroll
    for (int j = 0; j < NQC; j++) {
      const int kc = qc_kc + j * (FK_THREADS / F8_TQ);
      const bool ok = q_ok && c0 + kc < C;
      cp_async4(q_dst + boff_q + (unsigned)(j * (FK_THREADS / F8_TQ) * F8_TQ * 4), q_src + (ok ? (size_t)(c0 + kc) * a.scq : 0), ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  (void)rs_kc; (void)rs_r;

  const int nstage = (C + KC - 1) / KC;
  for (int r0 = 0; r0 < npad; r0 += F8_TR) {
    u64 acc2[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc2[i][j] = 0ull;
    load(r0, 0, 0);
    for (int st = 0; st < nstage; st++) {
      const int c0 = st * KC;
      const int kcn = min(KC, C - c0);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();  // stage st has landed for everyone; everyone is done with stage st-1 (the other buffer)
      if (st + 1 < nstage) load(r0, c0 + KC, (st + 1) & 1);
      const float* rb = rs + (st & 1) * KC * F8_TR + rh * 256 + lane * 4;
      const float* qb = qs + (st & 1) * KC * F8_TQ + qg * 8;
      auto step = [&](int kc) {
        const ulonglong2 r0v = *reinterpret_cast<const ulonglong2*>(rb + kc * F8_TR);
        const ulonglong2 r1v = *reinterpret_cast<const ulonglong2*>(rb + kc * F8_TR + 128);
        const float4 qa = *reinterpret_cast<const float4*>(qb + kc * F8_TQ);
        const float4 qc = *reinterpret_cast<const float4*>(qb + kc * F8_TQ + 4);
        const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qc.x, qc.y, qc.z, qc.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const u64 qd = pack2(qv[i], qv[i]);
          acc2[i][0] = fma2(qd, r0v.x, acc2[i][0]);
          acc2[i][1] = fma2(qd, r0v.y, acc2[i][1]);
          acc2[i][2] = fma2(qd, r1v.x, acc2[i][2]);
          acc2[i][3] = fma2(qd, r1v.y, acc2[i][3]);
        }
      };
      if (kcn == KC) {
#pragma unroll
        for (int kc = 0; kc < KC; kc++) step(kc);
      } else {
        for (int kc = 0; kc < kcn; kc++) step(kc);  // no padded product enters the chain
      }
    }
    if (UNION) __syncthreads();  // every warp is done with the last stage before the distance block overwrites it
    // dist = ((-2 * dot) + |q|^2) + |r|^2 ; padding columns at +inf
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int rbase = r0 + rh * 256 + h * 128 + lane * 4;
      float pn[4];
#pragma unroll
      for (int j = 0; j < 4; j++) pn[j] = rbase + j < N ? __ldg(a.pp + (size_t)b * N + rbase + j) : INF;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float d0 = lo2(acc2[i][2 * h]), d1 = hi2(acc2[i][2 * h]), d2 = lo2(acc2[i][2 * h + 1]), d3 = hi2(acc2[i][2 * h + 1]);
        float o[4] = {d0, d1, d2, d3};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          o[j] = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, o[j]), qn[i]), pn[j]);
          if (rbase + j >= N) o[j] = INF;
        }
        *reinterpret_cast<float4*>(&dist[(size_t)(qg * 8 + i) * npad + rbase]) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    __syncthreads();  // the next pass refills buffer 0
  }
  // rows were written by two warps each (the two reference halves): the barrier above makes them complete
#pragma unroll 1
  for (int i = 0; i < F8_TQ / FK_WARPS; i++) {
    const int row = warp * (F8_TQ / FK_WARPS) + i;
    const int q = q0 + row;
    if (q >= S) break;  // warp-uniform
    fk_select_row(a, dist + (size_t)row * npad, b, q, npad, bufd[warp], bufi[warp]);
  }
}

static int pow2floor(long long v) { int p = 1; while ((long long)p * 2 <= v) p *= 2; return p; }

// block width of ATen's reduce kernel for `rows` outputs of `d0` (vector) elements each
// (ReduceConfig::set_block_dimension, Reduce.cuh:100-108, max_num_threads = 512 for float)
static int torch_reduce_block_width(long long d0, long long rows) {
  const int maxt = 512;
  const int d0p = d0 < maxt ? pow2floor(d0) : maxt;
  const int d1p = rows < maxt ? pow2floor(rows) : maxt;
  int bw = d0p < 32 ? d0p : 32;
  const int bh = d1p < maxt / bw ? d1p : maxt / bw;
  bw = d0p < maxt / bh ? d0p : maxt / bh;
  return bw;
}

static int rowsumsq_launch(const float* x, float* out, int B, int C, int N, long long sb, long long sn, long long sc,
                           cudaStream_t stream, const char* who) {
  const long long rows = (long long)B * N;
  const int vec = (C >= 128) ? 1 : 0;  // "vectorize along input": dim0 >= 128 (Reduce.cuh:1099)
  if (vec && (C & 3))
    return set_error(PS_ERR_UNSUPPORTED, "%s: C=%d >= 128 must be a multiple of 4 (torch's vectorised row sum handles "
                     "ragged rows with a head/tail split that is not reproduced)", who, C);
  const int bw = torch_reduce_block_width(vec ? C / 4 : C, rows);
  if (bw > 32)
    return set_error(PS_ERR_UNSUPPORTED, "%s: fewer than 16 rows with C=%d: torch reduces such rows across warps, "
                     "not reproduced", who, C);
  rowsumsq_torch_kernel<<<ceil_div(rows, FK_WARPS), FK_THREADS, 0, stream>>>(x, out, C, N, sb, sn, sc, vec, bw, rows);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

}  // namespace
}  // namespace ps

using namespace ps;

// x layout flags: channel_major != 0 -> (B,C,N) (EdgeConv's tensors), else (B,N,C) (what query_knn_point receives)
extern "C" int ps_knn_feat(const float* xr, const float* xq, int* idx, int B, int C, int N, int S, int k,
                           int channel_major, int order, int dev, void* stream_) {
  PS_REQUIRE(B >= 0 && C > 0 && N > 0 && S >= 0 && k > 0, "ps_knn_feat: bad sizes B=%d C=%d N=%d S=%d k=%d", B, C, N, S, k);
  PS_REQUIRE(k <= N, "ps_knn_feat: k=%d exceeds the number of points N=%d", k, N);
  PS_REQUIRE(order == PS_ORDER_SORT || order == PS_ORDER_TOPK, "ps_knn_feat: unknown result order %d", order);
  if (k > 32) return set_error(PS_ERR_UNSUPPORTED, "ps_knn_feat: k=%d > 32 not supported", k);
  if (B == 0 || S == 0) return PS_OK;
  PS_REQUIRE(xr && xq && idx, "ps_knn_feat: null pointer");
  PS_REQUIRE(B <= 65535, "ps_knn_feat: B=%d exceeds the grid y limit", B);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_knn_feat: cannot select device %d", dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int npad = (N + FK_TR - 1) / FK_TR * FK_TR;
  int TQ = 0;
  const size_t tile_bytes = (size_t)FK_KC * FK_TR * 4;
  for (int t : {32, 16, 8})
    if (!TQ && (size_t)t * npad * 4 + (size_t)FK_KC * t * 4 + tile_bytes <= 200 * 1024) TQ = t;
  if (!TQ) return set_error(PS_ERR_UNSUPPORTED, "ps_knn_feat: N=%d reference points exceed the shared-memory distance block (max 6144)", N);
  // a finer query tile when the grid would not fill the GPU
  const int nsm = sm_count(dev);
  while (TQ > 8 && (long long)B * ceil_div(S, TQ) < 2ll * nsm) TQ /= 2;

  const bool self = (xq == xr) && (S == N);
  const size_t nn = (size_t)B * N + (self ? 0 : (size_t)B * S);
  ScratchGuard norms_mem;
  if (int rc = norms_mem.alloc(nn * sizeof(float), dev, stream)) return rc;
  float* norms = static_cast<float*>(norms_mem.ptr);
  FkArgs a;
  a.xq = xq; a.xr = xr; a.idx = idx; a.C = C; a.N = N; a.S = S; a.k = k; a.order = order; a.npad = npad;
  if (channel_major) { a.sbr = (long long)C * N; a.snr = 1; a.scr = N; a.sbq = (long long)C * S; a.snq = 1; a.scq = S; }
  else { a.sbr = (long long)N * C; a.snr = C; a.scr = 1; a.sbq = (long long)S * C; a.snq = C; a.scq = 1; }
  a.pp = norms;
  a.qq = self ? norms : norms + (size_t)B * N;
  if (int rc = rowsumsq_launch(xr, norms, B, C, N, a.sbr, a.snr, a.scr, stream, "ps_knn_feat")) return rc;
  if (!self)
    if (int rc = rowsumsq_launch(xq, norms + (size_t)B * N, B, C, S, a.sbq, a.snq, a.scq, stream, "ps_knn_feat")) return rc;
  // 8 x 8 register tiles for the models' EdgeConv shapes (full query tiles, N <= 1024): see knn_feat8_kernel
  bool use8 = TQ == 32 && N <= 1024;
  if (const char* e = getenv("PS_KNN_FEAT8")) use8 = use8 && atoi(e) != 0;
  if (use8) {
    a.npad = (N + F8_TR - 1) / F8_TR * F8_TR;
    const bool vec = a.snr == 1 && (N & 3) == 0 && (a.scr & 3) == 0 && (a.sbr & 3) == 0 && (reinterpret_cast<uintptr_t>(xr) & 15) == 0;
    const dim3 grid8(ceil_div(S, F8_TQ), B);
    // single-pass shapes: stages of 16 channels under the distance block (C=256 N=512: 0.166 -> see profiles/feature_knn_r2.jsonl)
    bool uni = a.npad == F8_TR && C >= 32;
    if (const char* e = getenv("PS_KNN_FEAT_UNION")) uni = uni && atoi(e) != 0;
#define PS_F8(VECV, KCV, UNI)                                                                                          \
  {                                                                                                                    \
    const size_t stages = (size_t)2 * KCV * (F8_TQ + F8_TR) * 4, block = (size_t)F8_TQ * a.npad * 4;                   \
    const size_t smem = UNI ? (stages > block ? stages : block) : stages + block;                                      \
    auto kern = knn_feat8_kernel<VECV, KCV, UNI>;                                                                      \
    PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                       \
    kern<<<grid8, FK_THREADS, smem, stream>>>(a);                                                                      \
  }
    // two-pass shapes (N <= 1024) hold one CTA per SM either way (128 KB distance block): 16-channel stages fit beside it
    const bool wide = !uni && C >= 32 && (size_t)F8_TQ * a.npad * 4 + (size_t)2 * 16 * (F8_TQ + F8_TR) * 4 <= 200 * 1024;
    if (uni) { if (vec) PS_F8(1, 16, true) else PS_F8(0, 16, true) }
    else if (wide) { if (vec) PS_F8(1, 16, false) else PS_F8(0, 16, false) }
    else { if (vec) PS_F8(1, 8, false) else PS_F8(0, 8, false) }
#undef PS_F8
    PS_LAUNCH_CHECK();
    return norms_mem.release();
  }
  const dim3 grid(ceil_div(S, TQ), B);
#define PS_FK(TQV)                                                                                  \
  {                                                                                                 \
    const size_t smem = (size_t)TQV * npad * 4 + (size_t)FK_KC * TQV * 4 + tile_bytes;              \
    auto kern = knn_feat_kernel<TQV>;                                                               \
    PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    kern<<<grid, FK_THREADS, smem, stream>>>(a);                                                    \
  }
  if (TQ == 32) PS_FK(32) else if (TQ == 16) PS_FK(16) else PS_FK(8)
#undef PS_FK
  PS_LAUNCH_CHECK();
  return norms_mem.release();
}
