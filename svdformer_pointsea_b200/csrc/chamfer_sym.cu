// Chamfer forward, symmetric single-pass variant: every pair distance is evaluated ONCE and used
// for both directions.
//
// The reference evaluates d(a_i, b_j) twice, once per direction (chamfer3D.cu:142-143).  Its
// expression d = fma(dz,dz, fma(dx,dx, dy*dy)) only squares the differences, and
// (b - a) == -(a - b) exactly in IEEE arithmetic, so both evaluations are bit-identical.  This
// kernel therefore halves the FP32-pipe work (the bound of the two-pass kernel in chamfer.cu):
//   * cloud A (the larger one) is held in registers, Q consecutive points per thread
//     (i = tile*256*Q + tid*Q + q); cloud B streams through shared memory as SoA tiles;
//   * row side (min over B for each A point): FMNMX3 running minimum + per-step bookkeeping and a
//     one-step re-evaluation for the argmin, exactly as in chamfer.cu;
//   * column side (min over A for each B point): per 4-target step every lane folds its Q
//     distances per target (FMNMX3), the warp folds the 32 lanes with REDUX.MIN on the fp32 bit
//     pattern (distances are >= 0, so unsigned order == float order), the lowest lane holding the
//     minimum is found with one ballot; the results of 8 steps are kept in registers (lane l owns
//     B point 32k+l) and merged as (min_bits << 32 | thread id) into a per-tile shared-memory
//     key array with one 64-bit atomicMin per lane.  Because a thread's points are
//     consecutive in i and thread ids order the same way, the key's minimum is the lowest-index
//     holder of the minimum distance inside the CTA.  After the tile, each B point re-evaluates
//     the Q points of the recorded thread (A tile kept in shared memory) to get the exact index;
//   * partial results of both sides (B splits for rows, A tiles for columns) meet in global
//     64-bit keys (dist_bits << 32 | idx) merged with atomicMin, then an unpack kernel.
// Index semantics are the reference's on both sides: lowest index among exact minima.
#include "comm.cuh"

#include <cstdlib>
#include <cstring>

namespace ps {

#ifndef PS_CS_THREADS
#define PS_CS_THREADS 256  // A/B builds: -DPS_CS_THREADS=128 makes four 4-warp CTAs per SM instead of two 8-warp ones
#endif
constexpr int CS_THREADS = PS_CS_THREADS;
constexpr int CS_PER_SM = 512 / CS_THREADS;  // resident CTAs per SM (128 registers per thread)
constexpr int CS_TILE = 2048;  // B points per shared-memory tile
constexpr int CS_STEP = 4;

struct SymParams {
  const float* a;   // (B, na, 3)  register side
  const float* b;   // (B, nb, 3)  streamed side
  u64* keys_a;      // (B, na) row-side merge keys   (min over b)
  u64* keys_b;      // (B, nb) column-side merge keys (min over a)
  int na, nb;
  int natiles;      // A tiles per cloud
  int nsplit;       // B splits per cloud
  int split_len;    // multiple of CS_STEP
  // Tail balancing: CTAs are dispatched in blockIdx order, so the last `units - ubig` units form the final,
  // partially filled wave.  Each of them is cut into `fsub` sub-units of sub_len targets (own CTAs), which
  // lets that wave finish in ~1/fsub of a full unit's time instead of a whole one.
  int ubig;         // units [0, ubig) run whole
  int fsub;         // sub-units per tail unit (1: none)
  int sub_len;      // multiple of CS_STEP
};

// blockIdx -> (cloud, A tile, [t0, t1) of B); false when the sub-unit is empty
__device__ __forceinline__ bool sym_decode_unit(const SymParams& p, int& b, int& at, int& t0, int& t1) {
  int unit = blockIdx.x, sub = -1;
  if (unit >= p.ubig) {
    const int r = unit - p.ubig;
    unit = p.ubig + r / p.fsub;
    sub = r % p.fsub;
  }
  const int split = unit % p.nsplit;
  unit /= p.nsplit;
  at = unit % p.natiles;
  b = unit / p.natiles;
  t0 = split * p.split_len;
  t1 = min(p.nb, t0 + p.split_len);
  if (sub >= 0) {
    t0 += sub * p.sub_len;
    t1 = min(t1, t0 + p.sub_len);
  }
  return t0 < t1;
}

// Epilogue of the forward: ONE launch unpacks both key arrays into (dist, idx), and — when the caller wants
// them — reduces the distances to the six loss sums of ps_chamfer_sums (utils/loss_utils.py:10-31) without
// re-reading them from HBM: per-block partials in a fixed order, the last block (ticket) folds them, again
// in a fixed order, so the sums are bit-reproducible run to run.  With a communicator the same last block
// PUBLISHES the sums into every peer's mailbox over NVLink (comm.cuh): the step's collective is fused here.
struct EpiParams {
  const u64* keys_a; const u64* keys_b;
  float* da; int* ia; float* db; int* ib;
  size_t nka, nkb;
  int a_is_1;        // cloud A is xyz1 (sums are reported in the caller's order)
  double* partial;   // [gridDim.x][4] scratch, or null: no sums
  unsigned* ticket;  // all-ones before the launch (the key fill covers it)
  double* out6;
  int publish;
};

__global__ void __launch_bounds__(256) sym_epilogue_kernel(const EpiParams p, const CommDev c) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double sa_sqrt = 0.0, sa = 0.0, sb_sqrt = 0.0, sb = 0.0;
  for (size_t i = i0; i < p.nka; i += stride) {
    const u64 k = p.keys_a[i];
    const float d = __uint_as_float((unsigned)(k >> 32));
    p.da[i] = d;
    p.ia[i] = (int)(unsigned)k;
    sa_sqrt += (double)sqrtf(d);
    sa += (double)d;
  }
  for (size_t i = i0; i < p.nkb; i += stride) {
    const u64 k = p.keys_b[i];
    const float d = __uint_as_float((unsigned)(k >> 32));
    p.db[i] = d;
    p.ib[i] = (int)(unsigned)k;
    sb_sqrt += (double)sqrtf(d);
    sb += (double)d;
  }
  if (p.partial == nullptr) return;
  // caller order: [sum sqrt d1, sum sqrt d2, sum d1, sum d2]
  double s[4];
  s[0] = p.a_is_1 ? sa_sqrt : sb_sqrt; s[1] = p.a_is_1 ? sb_sqrt : sa_sqrt;
  s[2] = p.a_is_1 ? sa : sb;           s[3] = p.a_is_1 ? sb : sa;
  __shared__ double sh[8][4];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  }
  if (lane == 0) { sh[warp][0] = s[0]; sh[warp][1] = s[1]; sh[warp][2] = s[2]; sh[warp][3] = s[3]; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += sh[w][threadIdx.x];
    p.partial[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(p.ticket, 1u) + 2u == gridDim.x);  // the ticket starts at 0xffffffff
  __syncthreads();
  if (!last) return;
  __threadfence();
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  for (unsigned g = threadIdx.x; g < gridDim.x; g += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 4; k++) t[k] += __ldcg(p.partial + (size_t)g * 4 + k);
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
  }
  __syncthreads();
  if (lane == 0) { sh[warp][0] = t[0]; sh[warp][1] = t[1]; sh[warp][2] = t[2]; sh[warp][3] = t[3]; }
  __syncthreads();
  __shared__ double fin[6];
  if (threadIdx.x < 4) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) v += sh[w][threadIdx.x];
    fin[threadIdx.x] = v;
  }
  if (threadIdx.x == 4) fin[4] = (double)(p.a_is_1 ? p.nka : p.nkb);
  if (threadIdx.x == 5) fin[5] = (double)(p.a_is_1 ? p.nkb : p.nka);
  __syncthreads();
  if (threadIdx.x < 6) p.out6[threadIdx.x] = fin[threadIdx.x];
  if (p.publish && warp == 0) comm_publish(c, fin, 6);
}

template <int Q>
__global__ void __launch_bounds__(CS_THREADS, CS_PER_SM) chamfer_sym_kernel(const SymParams p) {
  constexpr int TA = CS_THREADS * Q;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sx = reinterpret_cast<float*>(smem_raw);
  float* sy = sx + CS_TILE;
  float* sz = sy + CS_TILE;
  u64* colkey = reinterpret_cast<u64*>(sz + CS_TILE);
  float* ax = reinterpret_cast<float*>(colkey + CS_TILE);
  float* ay = ax + TA;
  float* az = ay + TA;

  int b, at, t0, t1;
  if (!sym_decode_unit(p, b, at, t0, t1)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int na = p.na, nb = p.nb;
  const float INF = __int_as_float(0x7f800000);

  // ---- A points into registers (and shared memory for the column-side index recovery) -------
  u64 nqx[Q], nqy[Q], nqz[Q];
  float best[Q];
  int cstep[Q];
  const float* abase = p.a + (size_t)b * na * 3;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int i = at * TA + tid * Q + q;
    const bool valid = i < na;
    // invalid slots sit at +inf: every distance to them is +inf (or NaN against +inf padding),
    // which neither FMNMX nor the integer REDUX ever selects over a finite value
    const float x = valid ? __ldg(abase + (size_t)i * 3 + 0) : INF;
    const float y = valid ? __ldg(abase + (size_t)i * 3 + 1) : 0.f;
    const float z = valid ? __ldg(abase + (size_t)i * 3 + 2) : 0.f;
    ax[tid * Q + q] = x; ay[tid * Q + q] = y; az[tid * Q + q] = z;
    nqx[q] = pack2(-x, -x); nqy[q] = pack2(-y, -y); nqz[q] = pack2(-z, -z);
    best[q] = INF;
    cstep[q] = 0;
  }

  const float* bcloud = p.b + (size_t)b * nb * 3;

  for (int ts = t0; ts < t1; ts += CS_TILE) {
    const int cnt = min(CS_TILE, t1 - ts);
    const int cnt_pad = (cnt + CS_STEP - 1) / CS_STEP * CS_STEP;
    const float* tb = bcloud + (size_t)ts * 3;
    const bool vec = (reinterpret_cast<uintptr_t>(tb) & 15) == 0;
    __syncthreads();  // previous tile (and its column pass) fully consumed
    for (int g = tid; g < cnt_pad / 4; g += CS_THREADS) {
      float4 X, Y, Z;
      if (vec && g * 4 + 4 <= cnt) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(tb + g * 12));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(tb + g * 12 + 4));
        const float4 c = __ldg(reinterpret_cast<const float4*>(tb + g * 12 + 8));
        X = make_float4(a.x, a.w, bb.z, c.y);
        Y = make_float4(a.y, bb.x, bb.w, c.z);
        Z = make_float4(a.z, bb.y, c.x, c.w);
      } else {
        float xs[4], ys[4], zs[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int pi = g * 4 + e;
          const bool in = pi < cnt;
          xs[e] = in ? __ldg(tb + pi * 3 + 0) : INF;
          ys[e] = in ? __ldg(tb + pi * 3 + 1) : 0.f;
          zs[e] = in ? __ldg(tb + pi * 3 + 2) : 0.f;
        }
        X = make_float4(xs[0], xs[1], xs[2], xs[3]);
        Y = make_float4(ys[0], ys[1], ys[2], ys[3]);
        Z = make_float4(zs[0], zs[1], zs[2], zs[3]);
      }
      *reinterpret_cast<float4*>(&sx[g * 4]) = X;
      *reinterpret_cast<float4*>(&sy[g * 4]) = Y;
      *reinterpret_cast<float4*>(&sz[g * 4]) = Z;
    }
    for (int j = tid; j < cnt_pad; j += CS_THREADS) colkey[j] = ~0ull;
    __syncthreads();

    // ---- scan: 4 B points x Q A points per lane per step ---------------------------------------
    // Column results are collected in registers for 8 steps (32 B points: lane l keeps the key of
    // B point j32 + l) and merged into shared memory with ONE 64-bit atomicMin per lane.
    int step = (ts - t0) / CS_STEP;
    for (int j32 = 0; j32 < cnt_pad; j32 += 32) {
      unsigned key_m = 0xffffffffu, key_b = 1u;  // owner lane keeps the ballot MASK; ffs once per 32 B points
      const int jend = min(cnt_pad, j32 + 32);
#pragma unroll 2
      for (int j = j32; j < jend; j += CS_STEP, step++) {
        const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(&sx[j]);
        const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(&sy[j]);
        const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(&sz[j]);
        float c0 = INF, c1 = INF, c2 = INF, c3 = INF;  // this lane's minimum over its Q points, per B point
#pragma unroll
        for (int q = 0; q < Q; q += 2) {
          const u64 d01 = dist2x2(X.x, Y.x, Z.x, nqx[q], nqy[q], nqz[q]);
          const u64 d23 = dist2x2(X.y, Y.y, Z.y, nqx[q], nqy[q], nqz[q]);
          float nb0 = min3(best[q], lo2(d01), hi2(d01));
          nb0 = min3(nb0, lo2(d23), hi2(d23));
          if (nb0 < best[q]) cstep[q] = step;
          best[q] = nb0;
          const u64 e01 = dist2x2(X.x, Y.x, Z.x, nqx[q + 1], nqy[q + 1], nqz[q + 1]);
          const u64 e23 = dist2x2(X.y, Y.y, Z.y, nqx[q + 1], nqy[q + 1], nqz[q + 1]);
          float nb1 = min3(best[q + 1], lo2(e01), hi2(e01));
          nb1 = min3(nb1, lo2(e23), hi2(e23));
          if (nb1 < best[q + 1]) cstep[q + 1] = step;
          best[q + 1] = nb1;
          c0 = min3(c0, lo2(d01), lo2(e01));
          c1 = min3(c1, hi2(d01), hi2(e01));
          c2 = min3(c2, lo2(d23), lo2(e23));
          c3 = min3(c3, hi2(d23), hi2(e23));
        }
        // warp minimum per B point + lowest lane holding it; the owner lane of each B point keeps it
        const unsigned b0 = __float_as_uint(c0), b1 = __float_as_uint(c1), b2 = __float_as_uint(c2), b3 = __float_as_uint(c3);
        const unsigned m0 = __reduce_min_sync(0xffffffffu, b0);
        const unsigned m1 = __reduce_min_sync(0xffffffffu, b1);
        const unsigned m2 = __reduce_min_sync(0xffffffffu, b2);
        const unsigned m3 = __reduce_min_sync(0xffffffffu, b3);
        const unsigned l0 = __ballot_sync(0xffffffffu, b0 == m0);
        const unsigned l1 = __ballot_sync(0xffffffffu, b1 == m1);
        const unsigned l2 = __ballot_sync(0xffffffffu, b2 == m2);
        const unsigned l3 = __ballot_sync(0xffffffffu, b3 == m3);
        const int o = lane - (j - j32);  // 0..3 for the four owner lanes of this step
        if (o == 0) { key_m = m0; key_b = l0; }
        if (o == 1) { key_m = m1; key_b = l1; }
        if (o == 2) { key_m = m2; key_b = l2; }
        if (o == 3) { key_m = m3; key_b = l3; }
      }
      if (j32 + lane < cnt_pad)
        atomicMin(&colkey[j32 + lane], ((u64)key_m << 32) | (unsigned)(warp * 32 + __ffs(key_b) - 1));
    }
    __syncthreads();  // all column keys of this tile are final

    // ---- column side: exact index inside the recorded thread's Q points, then global merge ----
    for (int jj = tid; jj < cnt; jj += CS_THREADS) {
      const u64 key = colkey[jj];
      const unsigned mbits = (unsigned)(key >> 32);
      const int tcand = (int)(unsigned)key;
      const float bx = sx[jj], by = sy[jj], bz = sz[jj];
      int found = 0;
#pragma unroll
      for (int q = Q - 1; q >= 0; q--) {
        const float d = dist2_ref(bx - ax[tcand * Q + q], by - ay[tcand * Q + q], bz - az[tcand * Q + q]);
        if (__float_as_uint(d) == mbits) found = q;
      }
      const int ia = at * TA + tcand * Q + found;
      atomicMin(&p.keys_b[(size_t)b * nb + ts + jj], ((u64)mbits << 32) | (unsigned)ia);
    }
  }

  // ---- row side: first B point of the remembered step that reproduces `best` --------------------
  const int last_ts = t0 + ((t1 - t0 - 1) / CS_TILE) * CS_TILE;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int i = at * TA + tid * Q + q;
    const int base = t0 + cstep[q] * CS_STEP;
    u64 d01, d23;
    if (base >= last_ts) {
      const int off = base - last_ts;
      const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(&sx[off]);
      const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(&sy[off]);
      const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(&sz[off]);
      d01 = dist2x2(X.x, Y.x, Z.x, nqx[q], nqy[q], nqz[q]);
      d23 = dist2x2(X.y, Y.y, Z.y, nqx[q], nqy[q], nqz[q]);
    } else {
      const float* tp = bcloud + (size_t)base * 3;
      float px[4], py[4], pz[4];
#pragma unroll
      for (int e = 0; e < 4; e++) { px[e] = __ldg(tp + e * 3 + 0); py[e] = __ldg(tp + e * 3 + 1); pz[e] = __ldg(tp + e * 3 + 2); }
      d01 = dist2x2(pack2(px[0], px[1]), pack2(py[0], py[1]), pack2(pz[0], pz[1]), nqx[q], nqy[q], nqz[q]);
      d23 = dist2x2(pack2(px[2], px[3]), pack2(py[2], py[3]), pack2(pz[2], pz[3]), nqx[q], nqy[q], nqz[q]);
    }
    int found = 0;
    if (hi2(d23) == best[q]) found = 3;
    if (lo2(d23) == best[q]) found = 2;
    if (hi2(d01) == best[q]) found = 1;
    if (lo2(d01) == best[q]) found = 0;
    if (i >= na) continue;
    atomicMin(&p.keys_a[(size_t)b * na + i], ((u64)__float_as_uint(best[q]) << 32) | (unsigned)(base + found));
  }
}

// ------------------------------------------------------------------------------------------------
// Software-pipelined variant of chamfer_sym_kernel<8> (same algorithm, same results).
//
// The scan has two kinds of work per 4-target step: ~96 packed FP32-pipe instructions (the distances, 192 pipe
// cycles) and ~70 ALU-pipe instructions (FMNMX3 row / column minima, the argmin bookkeeping, REDUX / VOTE of the
// column fold; the ALU pipe is half rate).  ptxas schedules a step as "all distances, then all minima", the four
// warps of a scheduler run the same code in near lockstep, and so the two pipes are used in ALTERNATION: ncu on
// chamfer_sym_kernel shows fma 61 % / alu 49 % / issue 61 % with math_pipe_throttle as the first stall reason —
// both pipes idle half of the time although neither is saturated.  Here the step is cut into two halves of four
// A points (A: q 0-3, B: q 4-7) and pipelined by hand: the distances of one half are computed WHILE the minima
// of the previous half are folded, in source order interleaved instruction group by instruction group, so every
// single warp feeds both pipes at once.  Live distance registers stay at 32 (16 produced + 16 consumed).
// GRAN: argmin bookkeeping once per GRAN steps (1 or 2): the final re-evaluation then looks at 4*GRAN candidates
template <int Q, int UNROLL, int GRAN>
__global__ void __launch_bounds__(CS_THREADS, CS_PER_SM) chamfer_sym2_kernel(const SymParams p) {
  static_assert(Q == 8, "the pipelined variant is written for 8 A points per thread");
  constexpr int TA = CS_THREADS * Q;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sx = reinterpret_cast<float*>(smem_raw);
  float* sy = sx + CS_TILE;
  float* sz = sy + CS_TILE;
  u64* colkey = reinterpret_cast<u64*>(sz + CS_TILE);
  float* ax = reinterpret_cast<float*>(colkey + CS_TILE);
  float* ay = ax + TA;
  float* az = ay + TA;

  int b, at, t0, t1;
  if (!sym_decode_unit(p, b, at, t0, t1)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int na = p.na, nb = p.nb;
  const float INF = __int_as_float(0x7f800000);

  u64 nqx[Q], nqy[Q], nqz[Q];
  float best[Q];
  int cstep[Q];
  const float* abase = p.a + (size_t)b * na * 3;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int i = at * TA + tid * Q + q;
    const bool valid = i < na;
    const float x = valid ? __ldg(abase + (size_t)i * 3 + 0) : INF;  // see chamfer_sym_kernel: +inf slots never win
    const float y = valid ? __ldg(abase + (size_t)i * 3 + 1) : 0.f;
    const float z = valid ? __ldg(abase + (size_t)i * 3 + 2) : 0.f;
    ax[tid * Q + q] = x; ay[tid * Q + q] = y; az[tid * Q + q] = z;
    nqx[q] = pack2(-x, -x); nqy[q] = pack2(-y, -y); nqz[q] = pack2(-z, -z);
    best[q] = INF;
    cstep[q] = 0;
  }
  const float* bcloud = p.b + (size_t)b * nb * 3;

  for (int ts = t0; ts < t1; ts += CS_TILE) {
    const int cnt = min(CS_TILE, t1 - ts);
    const int cnt_pad = (cnt + 31) / 32 * 32;  // whole 32-point blocks: the padding sits at +inf and never wins
    const float* tb = bcloud + (size_t)ts * 3;
    const bool vec = (reinterpret_cast<uintptr_t>(tb) & 15) == 0;
    __syncthreads();  // previous tile (and its column pass) fully consumed
    for (int g = tid; g < cnt_pad / 4; g += CS_THREADS) {
      float4 X, Y, Z;
      if (vec && g * 4 + 4 <= cnt) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(tb + g * 12));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(tb + g * 12 + 4));
        const float4 c = __ldg(reinterpret_cast<const float4*>(tb + g * 12 + 8));
        X = make_float4(a.x, a.w, bb.z, c.y);
        Y = make_float4(a.y, bb.x, bb.w, c.z);
        Z = make_float4(a.z, bb.y, c.x, c.w);
      } else {
        float xs[4], ys[4], zs[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int pi = g * 4 + e;
          const bool in = pi < cnt;
          xs[e] = in ? __ldg(tb + pi * 3 + 0) : INF;
          ys[e] = in ? __ldg(tb + pi * 3 + 1) : 0.f;
          zs[e] = in ? __ldg(tb + pi * 3 + 2) : 0.f;
        }
        X = make_float4(xs[0], xs[1], xs[2], xs[3]);
        Y = make_float4(ys[0], ys[1], ys[2], ys[3]);
        Z = make_float4(zs[0], zs[1], zs[2], zs[3]);
      }
      *reinterpret_cast<float4*>(&sx[g * 4]) = X;
      *reinterpret_cast<float4*>(&sy[g * 4]) = Y;
      *reinterpret_cast<float4*>(&sz[g * 4]) = Z;
    }
    for (int j = tid; j < cnt_pad; j += CS_THREADS) colkey[j] = ~0ull;
    __syncthreads();

    const int step0 = (ts - t0) / CS_STEP;
    // pipeline prologue: distances of half A (q 0-3) for step 0
    ulonglong2 X = *reinterpret_cast<const ulonglong2*>(&sx[0]);
    ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(&sy[0]);
    ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(&sz[0]);
    u64 dA01[4], dA23[4], dB01[4], dB23[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      dA01[i] = dist2x2(X.x, Y.x, Z.x, nqx[i], nqy[i], nqz[i]);
      dA23[i] = dist2x2(X.y, Y.y, Z.y, nqx[i], nqy[i], nqz[i]);
    }
    float old[Q];  // GRAN == 2: the running minima before the current pair of steps
#pragma unroll
    for (int q = 0; q < Q; q++) old[q] = INF;
    for (int j32 = 0; j32 < cnt_pad; j32 += 32) {
      unsigned key_m = 0xffffffffu, key_b = 1u;  // owner lane keeps the ballot MASK; ffs once per 32 B points
#pragma unroll UNROLL
      for (int s8 = 0; s8 < 8; s8++) {
        const int step = step0 + j32 / CS_STEP + s8;
        float c0 = INF, c1 = INF, c2 = INF, c3 = INF;
        // ---- half 1: distances of q 4-7 at this step  ||  minima of q 0-3 at this step ----------------
#pragma unroll
        for (int i = 0; i < 4; i++) {
          dB01[i] = dist2x2(X.x, Y.x, Z.x, nqx[4 + i], nqy[4 + i], nqz[4 + i]);
          dB23[i] = dist2x2(X.y, Y.y, Z.y, nqx[4 + i], nqy[4 + i], nqz[4 + i]);
          float nbv = min3(best[i], lo2(dA01[i]), hi2(dA01[i]));
          nbv = min3(nbv, lo2(dA23[i]), hi2(dA23[i]));
          if (GRAN == 1) { if (nbv < best[i]) cstep[i] = step; }
          else if (s8 & 1) { if (nbv < old[i]) cstep[i] = step >> 1; }
          else old[i] = best[i];
          best[i] = nbv;
          if (i & 1) {
            c0 = min3(c0, lo2(dA01[i - 1]), lo2(dA01[i]));
            c1 = min3(c1, hi2(dA01[i - 1]), hi2(dA01[i]));
            c2 = min3(c2, lo2(dA23[i - 1]), lo2(dA23[i]));
            c3 = min3(c3, hi2(dA23[i - 1]), hi2(dA23[i]));
          }
        }
        // B points of the next step.  Past the tile's last step this reads the bytes behind the arrays (still
        // inside this CTA's shared memory): those distances are computed and never consumed.
        const int jn = j32 + (s8 + 1) * CS_STEP;
        X = *reinterpret_cast<const ulonglong2*>(&sx[jn]);
        Y = *reinterpret_cast<const ulonglong2*>(&sy[jn]);
        Z = *reinterpret_cast<const ulonglong2*>(&sz[jn]);
        // ---- half 2: distances of q 0-3 at the next step  ||  minima of q 4-7 at this step -------------
#pragma unroll
        for (int i = 0; i < 4; i++) {
          dA01[i] = dist2x2(X.x, Y.x, Z.x, nqx[i], nqy[i], nqz[i]);
          dA23[i] = dist2x2(X.y, Y.y, Z.y, nqx[i], nqy[i], nqz[i]);
          float nbv = min3(best[4 + i], lo2(dB01[i]), hi2(dB01[i]));
          nbv = min3(nbv, lo2(dB23[i]), hi2(dB23[i]));
          if (GRAN == 1) { if (nbv < best[4 + i]) cstep[4 + i] = step; }
          else if (s8 & 1) { if (nbv < old[4 + i]) cstep[4 + i] = step >> 1; }
          else old[4 + i] = best[4 + i];
          best[4 + i] = nbv;
          if (i & 1) {
            c0 = min3(c0, lo2(dB01[i - 1]), lo2(dB01[i]));
            c1 = min3(c1, hi2(dB01[i - 1]), hi2(dB01[i]));
            c2 = min3(c2, lo2(dB23[i - 1]), lo2(dB23[i]));
            c3 = min3(c3, hi2(dB23[i - 1]), hi2(dB23[i]));
          }
        }
        // ---- column fold of this step: warp minimum per B point + the lanes holding it -----------------
        const unsigned b0 = __float_as_uint(c0), b1 = __float_as_uint(c1), b2 = __float_as_uint(c2), b3 = __float_as_uint(c3);
        const unsigned m0 = __reduce_min_sync(0xffffffffu, b0);
        const unsigned m1 = __reduce_min_sync(0xffffffffu, b1);
        const unsigned m2 = __reduce_min_sync(0xffffffffu, b2);
        const unsigned m3 = __reduce_min_sync(0xffffffffu, b3);
        const unsigned l0 = __ballot_sync(0xffffffffu, b0 == m0);
        const unsigned l1 = __ballot_sync(0xffffffffu, b1 == m1);
        const unsigned l2 = __ballot_sync(0xffffffffu, b2 == m2);
        const unsigned l3 = __ballot_sync(0xffffffffu, b3 == m3);
        const int o = lane - s8 * CS_STEP;  // 0..3 for the four owner lanes of this step
        if (o == 0) { key_m = m0; key_b = l0; }
        if (o == 1) { key_m = m1; key_b = l1; }
        if (o == 2) { key_m = m2; key_b = l2; }
        if (o == 3) { key_m = m3; key_b = l3; }
      }
      atomicMin(&colkey[j32 + lane], ((u64)key_m << 32) | (unsigned)(warp * 32 + __ffs(key_b) - 1));
    }
    __syncthreads();  // all column keys of this tile are final

    for (int jj = tid; jj < cnt; jj += CS_THREADS) {
      const u64 key = colkey[jj];
      const unsigned mbits = (unsigned)(key >> 32);
      const int tcand = (int)(unsigned)key;
      const float bx = sx[jj], by = sy[jj], bz = sz[jj];
      int found = 0;
#pragma unroll
      for (int q = Q - 1; q >= 0; q--) {
        const float d = dist2_ref(bx - ax[tcand * Q + q], by - ay[tcand * Q + q], bz - az[tcand * Q + q]);
        if (__float_as_uint(d) == mbits) found = q;
      }
      const int ia = at * TA + tcand * Q + found;
      atomicMin(&p.keys_b[(size_t)b * nb + ts + jj], ((u64)mbits << 32) | (unsigned)ia);
    }
  }

  // ---- row side: first B point of the remembered group of GRAN steps that reproduces `best` -----------------
  const int last_ts = t0 + ((t1 - t0 - 1) / CS_TILE) * CS_TILE;
  constexpr int NC = CS_STEP * GRAN;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int i = at * TA + tid * Q + q;
    const int base = t0 + cstep[q] * NC;
    const float qx = -lo2(nqx[q]), qy = -lo2(nqy[q]), qz = -lo2(nqz[q]);
    int found = 0;
    if (base >= last_ts) {
      const int off = base - last_ts;  // inside the staged tile (its padding sits at +inf)
#pragma unroll
      for (int e = NC - 1; e >= 0; e--)
        if (dist2_ref(sx[off + e] - qx, sy[off + e] - qy, sz[off + e] - qz) == best[q]) found = e;
    } else {
      const float* tp = bcloud + (size_t)base * 3;
#pragma unroll
      for (int e = NC - 1; e >= 0; e--)
        if (dist2_ref(__ldg(tp + e * 3 + 0) - qx, __ldg(tp + e * 3 + 1) - qy, __ldg(tp + e * 3 + 2) - qz) == best[q]) found = e;
    }
    if (i >= na) continue;
    atomicMin(&p.keys_a[(size_t)b * na + i], ((u64)__float_as_uint(best[q]) << 32) | (unsigned)(base + found));
  }
}

// ------------------------------------------------------------------------------------------------
// Measured and NOT kept (round 2, profiles/chamfer_lab_r2.jsonl variants 8-11): a "column store" version of the
// pipelined kernel in which the holders of each warp minimum store it to per-warp shared-memory rows (predicated by
// the ISETP that feeds the ballot) and lane 0 stores the four ballots with one 128-bit store, instead of routing
// them to owner lanes (ISETP + 2 SEL per target) and merging the warps with a 64-bit shared atomicMin per 32
// targets.  SASS per 4-target step: 171 -> 167 instructions, ALU pipe 68 -> 62 (53 with the argmin bookkeeping once
// per 32 targets), +5 LSU.  Bit-identical, but SLOWER on every shape: C1 291 -> 297 us, 32 x 16384^2 2040 -> 2083 us;
// with bookkeeping groups of 4 / 8 steps 318-324 / 406 us (the fully unrolled 20 KB loop body and the 16- / 32-candidate
// re-evaluation cost more than the saved compares).  The scan is not bound by the ALU instruction count.
// ------------------------------------------------------------------------------------------------
// A-packed variant.  Same algorithm, other register layout: the two halves of every packed
// operand are two DIFFERENT A points of the thread (a_2p, a_2p+1) and the streamed B point is
// duplicated in shared memory ((x_j, x_j) pairs, read as broadcast LDS.128).  An A point then
// costs 3 registers instead of 6, so a thread holds Q = 16 points at the same 128 registers and
// every per-B-point cost of the warp (column REDUX + ballot + owner selects, B loads, loop
// overhead) is amortised over twice as many pair evaluations; FLO/BREV (XU pipe, 1/8 rate) leave
// the loop because the owner lane keeps the ballot mask and takes ffs once per 32 B points.
constexpr int CP_TILE = 1024;  // B points per shared-memory tile (2 floats per coordinate)

// if (a < b) shared[addr] = v;  one FSETP + one predicated STS, the address stays in its register
__device__ __forceinline__ void st_shared_if_less(float a, float b, unsigned addr, int v) {
  // no "memory" clobber on purpose: it would pin the B-tile loads of the next step behind these stores.
  // The slots are only read back through ld_shared_volatile (volatile asm statements keep their order).
  asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %0, %1;\n\t@p st.shared.u32 [%2], %3;\n\t}" ::"f"(a), "f"(b), "r"(addr), "r"(v));
}
__device__ __forceinline__ void st_shared_volatile(unsigned addr, int v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ int ld_shared_volatile(unsigned addr) { int v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }

template <int QP>
__global__ void __launch_bounds__(CS_THREADS, QP >= 8 ? 2 : (QP == 4 ? 3 : 4)) chamfer_symp_kernel(const SymParams p) {
  constexpr int Q = 2 * QP;
  constexpr int TA = CS_THREADS * Q;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sxd = reinterpret_cast<float*>(smem_raw);
  float* syd = sxd + 2 * CP_TILE;
  float* szd = syd + 2 * CP_TILE;
  u64* colkey = reinterpret_cast<u64*>(szd + 2 * CP_TILE);
  float* ax = reinterpret_cast<float*>(colkey + CP_TILE);
  float* ay = ax + TA;
  float* az = ay + TA;
  // [q][thread]: step of the last strict improvement of best[q]
  const unsigned cstep_addr = smem_u32(reinterpret_cast<int*>(az + TA) + threadIdx.x);

  int b, at, t0, t1;
  if (!sym_decode_unit(p, b, at, t0, t1)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int na = p.na, nb = p.nb;
  const float INF = __int_as_float(0x7f800000);

  // ---- A points: negated and packed two per register pair; plain copy in shared memory --------
  u64 nax[QP], nay[QP], naz[QP];
  float best[Q];
  const float* abase = p.a + (size_t)b * na * 3;
#pragma unroll
  for (int pp = 0; pp < QP; pp++) {
    float x[2], y[2], z[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int i = at * TA + tid * Q + 2 * pp + h;
      const bool valid = i < na;
      x[h] = valid ? __ldg(abase + (size_t)i * 3 + 0) : INF;  // see chamfer_sym_kernel: +inf slots never win
      y[h] = valid ? __ldg(abase + (size_t)i * 3 + 1) : 0.f;
      z[h] = valid ? __ldg(abase + (size_t)i * 3 + 2) : 0.f;
      ax[tid * Q + 2 * pp + h] = x[h]; ay[tid * Q + 2 * pp + h] = y[h]; az[tid * Q + 2 * pp + h] = z[h];
      best[2 * pp + h] = INF;
      st_shared_volatile(cstep_addr + (2 * pp + h) * CS_THREADS * 4, 0);
    }
    nax[pp] = pack2(-x[0], -x[1]); nay[pp] = pack2(-y[0], -y[1]); naz[pp] = pack2(-z[0], -z[1]);
  }

  const float* bcloud = p.b + (size_t)b * nb * 3;

  for (int ts = t0; ts < t1; ts += CP_TILE) {
    const int cnt = min(CP_TILE, t1 - ts);
    const int cnt_pad = (cnt + CS_STEP - 1) / CS_STEP * CS_STEP;
    const float* tb = bcloud + (size_t)ts * 3;
    const bool vec = (reinterpret_cast<uintptr_t>(tb) & 15) == 0;
    __syncthreads();  // previous tile (and its column pass) fully consumed
    for (int g = tid; g < cnt_pad / 4; g += CS_THREADS) {
      float xs[4], ys[4], zs[4];
      if (vec && g * 4 + 4 <= cnt) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(tb + g * 12));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(tb + g * 12 + 4));
        const float4 c = __ldg(reinterpret_cast<const float4*>(tb + g * 12 + 8));
        xs[0] = a.x; xs[1] = a.w; xs[2] = bb.z; xs[3] = c.y;
        ys[0] = a.y; ys[1] = bb.x; ys[2] = bb.w; ys[3] = c.z;
        zs[0] = a.z; zs[1] = bb.y; zs[2] = c.x; zs[3] = c.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int pi = g * 4 + e;
          const bool in = pi < cnt;
          xs[e] = in ? __ldg(tb + pi * 3 + 0) : INF;
          ys[e] = in ? __ldg(tb + pi * 3 + 1) : 0.f;
          zs[e] = in ? __ldg(tb + pi * 3 + 2) : 0.f;
        }
      }
      *reinterpret_cast<float4*>(&sxd[g * 8]) = make_float4(xs[0], xs[0], xs[1], xs[1]);
      *reinterpret_cast<float4*>(&sxd[g * 8 + 4]) = make_float4(xs[2], xs[2], xs[3], xs[3]);
      *reinterpret_cast<float4*>(&syd[g * 8]) = make_float4(ys[0], ys[0], ys[1], ys[1]);
      *reinterpret_cast<float4*>(&syd[g * 8 + 4]) = make_float4(ys[2], ys[2], ys[3], ys[3]);
      *reinterpret_cast<float4*>(&szd[g * 8]) = make_float4(zs[0], zs[0], zs[1], zs[1]);
      *reinterpret_cast<float4*>(&szd[g * 8 + 4]) = make_float4(zs[2], zs[2], zs[3], zs[3]);
    }
    for (int j = tid; j < cnt_pad; j += CS_THREADS) colkey[j] = ~0ull;
    __syncthreads();

    int step = (ts - t0) / CS_STEP;
    for (int j32 = 0; j32 < cnt_pad; j32 += 32) {
      unsigned key_m = 0xffffffffu, key_b = 1u;
      const int jend = min(cnt_pad, j32 + 32);
#pragma unroll 2
      for (int j = j32; j < jend; j += CS_STEP, step++) {
        const ulonglong2 X01 = *reinterpret_cast<const ulonglong2*>(&sxd[2 * j]);
        const ulonglong2 X23 = *reinterpret_cast<const ulonglong2*>(&sxd[2 * j + 4]);
        const ulonglong2 Y01 = *reinterpret_cast<const ulonglong2*>(&syd[2 * j]);
        const ulonglong2 Y23 = *reinterpret_cast<const ulonglong2*>(&syd[2 * j + 4]);
        const ulonglong2 Z01 = *reinterpret_cast<const ulonglong2*>(&szd[2 * j]);
        const ulonglong2 Z23 = *reinterpret_cast<const ulonglong2*>(&szd[2 * j + 4]);
        float c0 = INF, c1 = INF, c2 = INF, c3 = INF;  // this lane's minimum over its Q points, per B point
#pragma unroll
        for (int pp = 0; pp < QP; pp++) {
          const u64 d0 = dist2x2(X01.x, Y01.x, Z01.x, nax[pp], nay[pp], naz[pp]);  // {d(a_2p, b_j), d(a_2p+1, b_j)}
          const u64 d1 = dist2x2(X01.y, Y01.y, Z01.y, nax[pp], nay[pp], naz[pp]);
          const u64 d2 = dist2x2(X23.x, Y23.x, Z23.x, nax[pp], nay[pp], naz[pp]);
          const u64 d3 = dist2x2(X23.y, Y23.y, Z23.y, nax[pp], nay[pp], naz[pp]);
          float nl = min3(best[2 * pp], lo2(d0), lo2(d1));
          nl = min3(nl, lo2(d2), lo2(d3));
          st_shared_if_less(nl, best[2 * pp], cstep_addr + (2 * pp) * CS_THREADS * 4, step);
          best[2 * pp] = nl;
          float nh = min3(best[2 * pp + 1], hi2(d0), hi2(d1));
          nh = min3(nh, hi2(d2), hi2(d3));
          st_shared_if_less(nh, best[2 * pp + 1], cstep_addr + (2 * pp + 1) * CS_THREADS * 4, step);
          best[2 * pp + 1] = nh;
          c0 = min3(c0, lo2(d0), hi2(d0));
          c1 = min3(c1, lo2(d1), hi2(d1));
          c2 = min3(c2, lo2(d2), hi2(d2));
          c3 = min3(c3, lo2(d3), hi2(d3));
        }
        // warp minimum per B point + the lanes holding it; the owner lane of each B point keeps both
        const unsigned b0 = __float_as_uint(c0), b1 = __float_as_uint(c1), b2 = __float_as_uint(c2), b3 = __float_as_uint(c3);
        const unsigned m0 = __reduce_min_sync(0xffffffffu, b0);
        const unsigned m1 = __reduce_min_sync(0xffffffffu, b1);
        const unsigned m2 = __reduce_min_sync(0xffffffffu, b2);
        const unsigned m3 = __reduce_min_sync(0xffffffffu, b3);
        const unsigned l0 = __ballot_sync(0xffffffffu, b0 == m0);
        const unsigned l1 = __ballot_sync(0xffffffffu, b1 == m1);
        const unsigned l2 = __ballot_sync(0xffffffffu, b2 == m2);
        const unsigned l3 = __ballot_sync(0xffffffffu, b3 == m3);
        const int o = lane - (j - j32);  // 0..3 for the four owner lanes of this step
        if (o == 0) { key_m = m0; key_b = l0; }
        if (o == 1) { key_m = m1; key_b = l1; }
        if (o == 2) { key_m = m2; key_b = l2; }
        if (o == 3) { key_m = m3; key_b = l3; }
      }
      if (j32 + lane < cnt_pad)
        atomicMin(&colkey[j32 + lane], ((u64)key_m << 32) | (unsigned)(warp * 32 + __ffs(key_b) - 1));
    }
    __syncthreads();  // all column keys of this tile are final

    // ---- column side: exact index inside the recorded thread's Q points, then global merge ----
    for (int jj = tid; jj < cnt; jj += CS_THREADS) {
      const u64 key = colkey[jj];
      const unsigned mbits = (unsigned)(key >> 32);
      const int tcand = (int)(unsigned)key;
      const float bx = sxd[2 * jj], by = syd[2 * jj], bz = szd[2 * jj];
      int found = 0;
#pragma unroll
      for (int q = Q - 1; q >= 0; q--) {
        const float d = dist2_ref(bx - ax[tcand * Q + q], by - ay[tcand * Q + q], bz - az[tcand * Q + q]);
        if (__float_as_uint(d) == mbits) found = q;
      }
      const int ia = at * TA + tcand * Q + found;
      atomicMin(&p.keys_b[(size_t)b * nb + ts + jj], ((u64)mbits << 32) | (unsigned)ia);
    }
  }

  // ---- row side: first B point of the remembered step that reproduces `best` --------------------
  const int last_ts = t0 + ((t1 - t0 - 1) / CP_TILE) * CP_TILE;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int i = at * TA + tid * Q + q;
    const int base = t0 + ld_shared_volatile(cstep_addr + q * CS_THREADS * 4) * CS_STEP;
    const float qx = (q & 1) ? hi2(nax[q >> 1]) : lo2(nax[q >> 1]);  // negated coordinates
    const float qy = (q & 1) ? hi2(nay[q >> 1]) : lo2(nay[q >> 1]);
    const float qz = (q & 1) ? hi2(naz[q >> 1]) : lo2(naz[q >> 1]);
    float d[4];
    if (base >= last_ts) {
      const int off = base - last_ts;
#pragma unroll
      for (int e = 0; e < 4; e++) d[e] = dist2_ref(sxd[2 * (off + e)] + qx, syd[2 * (off + e)] + qy, szd[2 * (off + e)] + qz);
    } else {
      const float* tp = bcloud + (size_t)base * 3;
#pragma unroll
      for (int e = 0; e < 4; e++) d[e] = dist2_ref(__ldg(tp + e * 3 + 0) + qx, __ldg(tp + e * 3 + 1) + qy, __ldg(tp + e * 3 + 2) + qz);
    }
    int found = 0;
    if (d[3] == best[q]) found = 3;
    if (d[2] == best[q]) found = 2;
    if (d[1] == best[q]) found = 1;
    if (d[0] == best[q]) found = 0;
    if (i >= na) continue;
    atomicMin(&p.keys_a[(size_t)b * na + i], ((u64)__float_as_uint(best[q]) << 32) | (unsigned)(base + found));
  }
}

template <int QP>
static int launch_symp(const SymParams& p, int grid, cudaStream_t stream) {
  const size_t smem = (size_t)6 * CP_TILE * 4 + (size_t)CP_TILE * 8 + (size_t)4 * CS_THREADS * 2 * QP * 4;
  auto kern = chamfer_symp_kernel<QP>;
  PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, CS_THREADS, smem, stream>>>(p);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

template <int Q>
static int launch_sym(const SymParams& p, int grid, cudaStream_t stream) {
  const size_t smem = (size_t)3 * CS_TILE * 4 + (size_t)CS_TILE * 8 + (size_t)3 * CS_THREADS * Q * 4;
  auto kern = chamfer_sym_kernel<Q>;
  PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, CS_THREADS, smem, stream>>>(p);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

template <int UNROLL, int GRAN>
static int launch_sym2(const SymParams& p, int grid, cudaStream_t stream) {
  const size_t smem = (size_t)3 * CS_TILE * 4 + (size_t)CS_TILE * 8 + (size_t)3 * CS_THREADS * 8 * 4;
  auto kern = chamfer_sym2_kernel<8, UNROLL, GRAN>;
  PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, CS_THREADS, smem, stream>>>(p);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

// Returns PS_OK when it handled the call, 1 when the shape is better served by the two-pass kernel.
int chamfer_fwd_symmetric(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int* idx1,
                          int* idx2, double* sums6, const ps_comm* comm, int B, int N, int M, int dev, cudaStream_t stream) {
  // 0: two-pass kernel (chamfer.cu); 1: chamfer_sym_kernel; 2: A-packed chamfer_symp_kernel; 3-7: software-pipelined
  // chamfer_sym2_kernel<8, UNROLL, GRAN> (3: 8,1  4: 4,1  5: 2,1  6: 4,2  7: 2,2); default -1 = 6 where it applies
  // (measured on B200, tools/chamfer_lab.py: C1 305 -> 291 us, 32 x 16384^2 2181 -> 2044 us), else 1
  int variant = -1;
  if (const char* e = getenv("PS_CHAMFER_SYM")) variant = atoi(e);
  if (variant == 0) return 1;
  const int big = N > M ? N : M, small = N > M ? M : N;
  if (small < 256 || big < 1024) return 1;  // tiny clouds: launch-bound either way
  const int nsm = sm_count(dev);
  const bool a_is_1 = N >= M;  // the larger cloud sits in registers
  SymParams p;
  p.a = a_is_1 ? xyz1 : xyz2;
  p.b = a_is_1 ? xyz2 : xyz1;
  p.na = big;
  p.nb = small;
  const int qmin = variant == 2 ? 4 : 2;
  int Q = variant == 2 ? 16 : 8;
  if (const char* e = getenv("PS_CHAMFER_SYM_Q")) { const int v = atoi(e); if (v >= qmin && v <= Q && (v & (v - 1)) == 0) Q = v; }
  while (Q > qmin && CS_THREADS * Q / 2 >= big) Q /= 2;
  p.natiles = ceil_div(big, CS_THREADS * Q);
  // B-side split: enough units for >= 3 waves of the 2 resident CTAs per SM (measured on C1:
  // L=256 0.339 ms, L=512 0.330, L=1024 0.367, L=2048 0.375), but never below 256 targets per
  // unit so the per-unit fixed cost (A load, column pass, merge atomics) stays small.
  const int slots = nsm * CS_PER_SM;
  const long long base_units = (long long)B * p.natiles;
  int nsplit = (int)((3ll * slots + base_units - 1) / base_units);
  int max_split = small / 256 > 0 ? small / 256 : 1;
  // small clouds leave the GPU under-filled at 256 targets per unit (512 x 2048, B=32: 64 CTAs on 296 slots):
  // there, units of 128 targets win (45.4 -> 35.2 us) although their fixed cost is relatively larger
  // (only below half a wave: at 2048 x 2048, 256 CTAs of 256 targets beat 512 CTAs of 128: 67.9 vs 69.9 us)
  if (2 * base_units * max_split < slots && small / 128 > max_split) max_split = small / 128;
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  int bestL = ceil_div(small, nsplit);
  bestL = (bestL + CS_STEP - 1) / CS_STEP * CS_STEP;
  if (const char* e = getenv("PS_CHAMFER_SPLIT")) { const int v = atoi(e); if (v >= CS_STEP) bestL = (v + 3) / 4 * 4; }
  p.split_len = bestL;
  p.nsplit = ceil_div(small, bestL);

  const size_t nka = (size_t)B * big, nkb = (size_t)B * small;
  // scratch: [keys_a][keys_b][ticket, 16 bytes][epilogue partials]; the fill covers keys + ticket
  int egrid = ceil_div((long long)(nka + nkb), 256 * 4);
  if (egrid > nsm * 4) egrid = nsm * 4;
  const size_t fill_bytes = (nka + nkb) * sizeof(u64) + 16;
  ScratchGuard scratch_mem;
  if (int rc = scratch_mem.alloc(fill_bytes + (size_t)egrid * 4 * sizeof(double), dev, stream)) return rc;
  u64* scratch = static_cast<u64*>(scratch_mem.ptr);
  if (int rc = fill32_async(scratch, 0xffffffffu, fill_bytes, stream)) return rc;
  p.keys_a = scratch;
  p.keys_b = scratch + nka;
  // tail balancing (see SymParams): cut the units of the last, partially filled wave
  const long long units = (long long)B * p.natiles * p.nsplit;
  p.ubig = (int)units;
  p.fsub = 1;
  p.sub_len = p.split_len;
  {
    int tail = 1;
    if (const char* e = getenv("PS_CHAMFER_TAIL")) tail = atoi(e);
    const long long full = units / slots, rem = units % slots;
    if (tail && full >= 1 && rem > 0) {
      double best = (double)full + 1.0;
      for (int f = 2; f <= 4; f *= 2) {
        const int sl = (ceil_div(p.split_len, f) + CS_STEP - 1) / CS_STEP * CS_STEP;
        if (sl < 128) break;
        const double cost = (double)full + (double)ceil_div(rem * f, slots) / f + 0.02 * f;  // small bias against needless cuts
        if (cost < best - 1e-9) { best = cost; p.fsub = f; p.sub_len = sl; }
      }
      if (p.fsub > 1) p.ubig = (int)(units - rem);
    }
  }
  const int grid = (int)(p.ubig + (units - p.ubig) * p.fsub);
  if (variant < 0) {
    // the pipelined kernel scans whole 32-point blocks: ragged splits pay for their padding (8 x 1000 x 5000: +5 %)
    const bool blocks = Q == 8 && (small % 32) == 0 && (p.split_len % 32) == 0 && (p.sub_len % 32) == 0;
    variant = blocks ? 6 : 1;
  }
  int rc;
  if (variant == 2) {
    if (Q == 16) rc = launch_symp<8>(p, grid, stream);
    else if (Q == 8) rc = launch_symp<4>(p, grid, stream);
    else rc = launch_symp<2>(p, grid, stream);
  } else {
    if (Q == 8 && variant == 3) rc = launch_sym2<8, 1>(p, grid, stream);
    else if (Q == 8 && variant == 4) rc = launch_sym2<4, 1>(p, grid, stream);
    else if (Q == 8 && variant == 5) rc = launch_sym2<2, 1>(p, grid, stream);
    else if (Q == 8 && variant == 6) rc = launch_sym2<4, 2>(p, grid, stream);
    else if (Q == 8 && variant == 7) rc = launch_sym2<2, 2>(p, grid, stream);
    else if (Q == 8) rc = launch_sym<8>(p, grid, stream);
    else if (Q == 4) rc = launch_sym<4>(p, grid, stream);
    else rc = launch_sym<2>(p, grid, stream);
  }
  if (rc) return rc;
  EpiParams e;
  e.keys_a = p.keys_a; e.keys_b = p.keys_b;
  e.da = a_is_1 ? dist1 : dist2; e.ia = a_is_1 ? idx1 : idx2;
  e.db = a_is_1 ? dist2 : dist1; e.ib = a_is_1 ? idx2 : idx1;
  e.nka = nka; e.nkb = nkb;
  e.a_is_1 = a_is_1 ? 1 : 0;
  e.ticket = reinterpret_cast<unsigned*>(scratch + nka + nkb);
  e.partial = sums6 ? reinterpret_cast<double*>(scratch + nka + nkb + 2) : nullptr;
  e.out6 = sums6;
  e.publish = (sums6 && comm) ? 1 : 0;
  CommDev cd;
  if (e.publish) cd = comm->d; else memset(&cd, 0, sizeof(cd));
  sym_epilogue_kernel<<<egrid, 256, 0, stream>>>(e, cd);
  PS_LAUNCH_CHECK();
  return scratch_mem.release();
}

}  // namespace ps
