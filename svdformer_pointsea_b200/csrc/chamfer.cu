// Chamfer distance forward/backward for sm_100a.
//
// Replaces NmDistanceKernel / NmDistanceGradKernel of the reference
// (metrics/CD/chamfer3D/chamfer3D.cu:12-134, :155-174) behind the C ABI in pointsea_b200.h.
//
// Forward design (fp32 CUDA-core bound, no tensor cores: the contraction has K=3):
//   * one launch covers BOTH directions; a CTA owns (direction, cloud, query tile, target split)
//   * every thread keeps Q queries in registers; the target split is staged through shared
//     memory as SoA tiles (x[], y[], z[]) so one broadcast LDS.128 per coordinate feeds
//     4 targets x Q queries
//   * distances are evaluated two targets at a time with the packed sm_100 instructions
//     FADD2 / FMUL2 / FFMA2 in exactly the reference's rounding order
//     d = fma(dz,dz, fma(dx,dx, dy*dy)), dx = target - query; the running minimum is one
//     FMNMX3 per two pairs
//   * the argmin index is recovered lazily: per 4-target step we only note whether the running
//     minimum improved (strict <, so the EARLIEST step holding the final minimum is remembered;
//     one FSETP + one SEL on the otherwise idle ALU pipe, the FP32 pipe is the bottleneck);
//     after the scan each thread re-evaluates that one step and takes the first of its four
//     targets whose distance equals the minimum bit-for-bit.  This is the reference's "lowest
//     index among exact minima" (strict < within a tile, chamfer3D.cu:36-70, strict > across
//     tiles, :126).
//   * when the targets are split across CTAs the partial results are merged with a 64-bit
//     atomicMin on (dist_bits << 32 | idx): distances are >= 0 so the bit pattern is
//     monotone, and equal distances resolve to the lower index.
#include "comm.cuh"

#include <cstdlib>

namespace ps {

constexpr int CH_THREADS = 256;
constexpr int CH_TILE = 2048;  // targets per shared-memory tile (24 KB SoA)
constexpr int CH_STEP = 4;     // targets per inner step = argmin bookkeeping granularity

struct ChamferDir {
  const float* q;  // queries (B, nq, 3)
  const float* t;  // targets (B, nt, 3)
  float* dist;     // (B, nq)
  int* idx;        // (B, nq)
  u64* keys;       // (B, nq) merge scratch when nsplit > 1
  int nq, nt;
  int nqtiles;    // query tiles per cloud
  int nsplit;     // target splits per cloud
  int split_len;  // targets per split (multiple of CH_STEP)
  int units;      // B * nqtiles * nsplit
};
struct ChamferParams {
  ChamferDir d[2];
};

template <int Q>
__global__ void __launch_bounds__(CH_THREADS) chamfer_nn_kernel(const ChamferParams p) {
  __shared__ __align__(16) float sx[CH_TILE];
  __shared__ __align__(16) float sy[CH_TILE];
  __shared__ __align__(16) float sz[CH_TILE];

  int unit = blockIdx.x;
  const int dir = unit >= p.d[0].units ? 1 : 0;
  if (dir) unit -= p.d[0].units;
  // field-wise select keeps the parameter struct in constant space (no local copy)
  struct {
    const float *q, *t; float* dist; int* idx; u64* keys; int nq, nt, nqtiles, nsplit, split_len;
  } D;
  D.q = dir ? p.d[1].q : p.d[0].q;
  D.t = dir ? p.d[1].t : p.d[0].t;
  D.dist = dir ? p.d[1].dist : p.d[0].dist;
  D.idx = dir ? p.d[1].idx : p.d[0].idx;
  D.keys = dir ? p.d[1].keys : p.d[0].keys;
  D.nq = dir ? p.d[1].nq : p.d[0].nq;
  D.nt = dir ? p.d[1].nt : p.d[0].nt;
  D.nqtiles = dir ? p.d[1].nqtiles : p.d[0].nqtiles;
  D.nsplit = dir ? p.d[1].nsplit : p.d[0].nsplit;
  D.split_len = dir ? p.d[1].split_len : p.d[0].split_len;
  const int split = unit % D.nsplit;
  const int rest = unit / D.nsplit;
  const int qt = rest % D.nqtiles;
  const int b = rest / D.nqtiles;
  const int tid = threadIdx.x;
  const int nq = D.nq, nt = D.nt;
  const float INF = __int_as_float(0x7f800000);

  // ---- queries into registers -------------------------------------------------------------
  float qx[Q], qy[Q], qz[Q];
  u64 nqx[Q], nqy[Q], nqz[Q];
  float best[Q];
  int cstep[Q];
  const float* qbase = D.q + (size_t)b * nq * 3;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int qi = qt * (CH_THREADS * Q) + q * CH_THREADS + tid;
    const bool valid = qi < nq;
    qx[q] = valid ? __ldg(qbase + (size_t)qi * 3 + 0) : 0.f;
    qy[q] = valid ? __ldg(qbase + (size_t)qi * 3 + 1) : 0.f;
    qz[q] = valid ? __ldg(qbase + (size_t)qi * 3 + 2) : 0.f;
    nqx[q] = pack2(-qx[q], -qx[q]);
    nqy[q] = pack2(-qy[q], -qy[q]);
    nqz[q] = pack2(-qz[q], -qz[q]);
    best[q] = INF;
    cstep[q] = 0;
  }

  const int t0 = split * D.split_len;
  const int t1 = min(nt, t0 + D.split_len);
  const float* tcloud = D.t + (size_t)b * nt * 3;

  for (int ts = t0; ts < t1; ts += CH_TILE) {
    const int cnt = min(CH_TILE, t1 - ts);
    const int cnt_pad = (cnt + CH_STEP - 1) / CH_STEP * CH_STEP;
    const float* tb = tcloud + (size_t)ts * 3;
    const bool vec = (reinterpret_cast<uintptr_t>(tb) & 15) == 0;
    __syncthreads();  // previous tile fully consumed
    // ---- stage AoS (x,y,z) points as SoA; pad the last step with +inf targets ---------------
    for (int g = tid; g < cnt_pad / 4; g += CH_THREADS) {
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
    __syncthreads();

    // ---- scan the tile: 4 targets x Q queries per step ----------------------------------------
    int step = (ts - t0) / CH_STEP;
#pragma unroll 4
    for (int j = 0; j < cnt_pad; j += CH_STEP, step++) {
      const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(&sx[j]);
      const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(&sy[j]);
      const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(&sz[j]);
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const u64 d01 = dist2x2(X.x, Y.x, Z.x, nqx[q], nqy[q], nqz[q]);
        const u64 d23 = dist2x2(X.y, Y.y, Z.y, nqx[q], nqy[q], nqz[q]);
        float nb = min3(best[q], lo2(d01), hi2(d01));
        nb = min3(nb, lo2(d23), hi2(d23));
        if (nb < best[q]) cstep[q] = step;
        best[q] = nb;
      }
    }
  }

  // ---- recover the argmin: first target of the remembered step that reproduces `best` --------
  // Steps inside the last tile of the split are re-evaluated from shared memory (still resident);
  // earlier tiles (multi-tile splits only) are re-read from global memory (L2-resident).
  const int last_ts = t0 + ((t1 - t0 - 1) / CH_TILE) * CH_TILE;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int qi = qt * (CH_THREADS * Q) + q * CH_THREADS + tid;
    const int base = t0 + cstep[q] * CH_STEP;
    u64 d01, d23;
    if (base >= last_ts) {
      const int off = base - last_ts;
      const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(&sx[off]);
      const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(&sy[off]);
      const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(&sz[off]);
      d01 = dist2x2(X.x, Y.x, Z.x, nqx[q], nqy[q], nqz[q]);
      d23 = dist2x2(X.y, Y.y, Z.y, nqx[q], nqy[q], nqz[q]);
    } else {
      // earlier tiles are full (CH_TILE is a multiple of CH_STEP), so all four targets exist
      const float* tp = tcloud + (size_t)base * 3;
      float px[4], py[4], pz[4];
#pragma unroll
      for (int e = 0; e < 4; e++) { px[e] = __ldg(tp + e * 3 + 0); py[e] = __ldg(tp + e * 3 + 1); pz[e] = __ldg(tp + e * 3 + 2); }
      d01 = dist2x2(pack2(px[0], px[1]), pack2(py[0], py[1]), pack2(pz[0], pz[1]), nqx[q], nqy[q], nqz[q]);
      d23 = dist2x2(pack2(px[2], px[3]), pack2(py[2], py[3]), pack2(pz[2], pz[3]), nqx[q], nqy[q], nqz[q]);
    }
    int found = 0;  // stays 0 only for non-finite inputs
    if (hi2(d23) == best[q]) found = 3;
    if (lo2(d23) == best[q]) found = 2;
    if (hi2(d01) == best[q]) found = 1;
    if (lo2(d01) == best[q]) found = 0;
    if (qi >= nq) continue;
    const int gi = base + found;
    const size_t o = (size_t)b * nq + qi;
    if (D.nsplit == 1) {
      D.dist[o] = best[q];
      D.idx[o] = gi;
    } else {
      atomicMin(&D.keys[o], ((u64)__float_as_uint(best[q]) << 32) | (unsigned)gi);
    }
  }
}

__global__ void chamfer_unpack_kernel(const u64* __restrict__ keys, float* __restrict__ dist,
                                      int* __restrict__ idx, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const u64 k = keys[i];
    dist[i] = __uint_as_float((unsigned)(k >> 32));
    idx[i] = (int)(unsigned)k;
  }
}

// ---- backward ---------------------------------------------------------------------------------
// Pass A writes every point's own term with plain stores (covers both buffers completely, so the
// caller does not have to zero them); pass B adds the scattered terms with RED.ADD.
//   side 1: g = 2*gd1[i]; v = g*(a_i - b_j), j = idx1[i]; grad1[i] = v (A); grad2[j] += -v (B)
//   side 2: symmetric with idx2 (reference launches the same kernel with the roles swapped,
//   chamfer3D.cu:184-185).
struct ChamferBwdParams {
  const float *xyz1, *xyz2, *gd1, *gd2;
  const int *idx1, *idx2;
  float *g1, *g2;
  int B, N, M;
};

template <bool SCATTER>
__global__ void __launch_bounds__(256) chamfer_bwd_kernel(const ChamferBwdParams p) {
  const size_t tot1 = (size_t)p.B * p.N, tot2 = (size_t)p.B * p.M;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tot1 + tot2) return;
  const float *a, *bt, *gd;
  const int* ix;
  float *ga, *gb;
  int na, nb;
  if (i < tot1) {
    a = p.xyz1; bt = p.xyz2; gd = p.gd1; ix = p.idx1; ga = p.g1; gb = p.g2; na = p.N; nb = p.M;
  } else {
    i -= tot1;
    a = p.xyz2; bt = p.xyz1; gd = p.gd2; ix = p.idx2; ga = p.g2; gb = p.g1; na = p.M; nb = p.N;
  }
  const size_t b = i / na;
  const int j = __ldg(ix + i);
  const float* pa = a + i * 3;
  const float* pb = bt + (b * nb + j) * 3;
  const float g0 = __ldg(gd + i);
  const float g = g0 + g0;  // reference: grad_dist*2 (SASS: FADD g,g)
  const float vx = __fmul_rn(g, __ldg(pa + 0) - __ldg(pb + 0));
  const float vy = __fmul_rn(g, __ldg(pa + 1) - __ldg(pb + 1));
  const float vz = __fmul_rn(g, __ldg(pa + 2) - __ldg(pb + 2));
  if (!SCATTER) {
    ga[i * 3 + 0] = vx;
    ga[i * 3 + 1] = vy;
    ga[i * 3 + 2] = vz;
  } else {
    float* o = gb + (b * nb + j) * 3;
    atomicAdd(o + 0, -vx);
    atomicAdd(o + 1, -vy);
    atomicAdd(o + 2, -vz);
  }
}

// Unit planning.  A unit = (cloud, tile of 256*Q queries, split of L targets).  All units of a
// launch get (nearly) the same L so they take the same time; L is chosen so that the unit count
// lands just below a whole number of waves of resident CTAs (the tail wave otherwise costs up to a
// full unit time) while the per-unit fixed cost (query load, staging, merge atomics) stays small.
static void set_split(ChamferDir& D, int B, int Q, int L) {
  D.nqtiles = ceil_div(D.nq, CH_THREADS * Q);
  int nsplit = ceil_div(D.nt, L);
  int len = ceil_div(D.nt, nsplit);
  len = (len + CH_STEP - 1) / CH_STEP * CH_STEP;
  D.split_len = len;
  D.nsplit = ceil_div(D.nt, len);
  D.units = B * D.nqtiles * D.nsplit;
}

static void plan_units(ChamferParams& p, int B, int Q, int nsm) {
  const int slots = nsm * 4;  // 64 registers x 256 threads -> 4 CTAs per SM
  const int maxt = p.d[0].nt > p.d[1].nt ? p.d[0].nt : p.d[1].nt;
  int bestL = maxt;
  double best_score = -1.0;
  for (int L = 256; L <= 16384; L += 256) {
    ChamferDir a = p.d[0], b = p.d[1];
    set_split(a, B, Q, L);
    set_split(b, B, Q, L);
    // work in units of (queries x targets); waves by the longest-processing-time bound
    const double w0 = (double)a.split_len, w1 = (double)b.split_len;
    const double total = a.units * w0 + b.units * w1;
    const double wmax = w0 > w1 ? w0 : w1;
    const double units = (double)a.units + b.units;
    double makespan;
    if (units <= slots) makespan = wmax;
    else {
      const double waves = units / slots;
      const double avg = total / units;
      makespan = ((double)(long long)waves + ((waves > (long long)waves) ? 1.0 : 0.0)) * avg;
      if (makespan < total / slots) makespan = total / slots;
    }
    const double fixed = 48.0;  // per-unit fixed cost expressed in "targets"
    const double score = (total / slots) / (makespan * (1.0 + fixed / (w0 < w1 ? w0 : w1)));
    if (score > best_score + 1e-9) { best_score = score; bestL = L; }
    if (L >= maxt) break;
  }
  if (const char* e = getenv("PS_CHAMFER_SPLIT")) { const int v = atoi(e); if (v >= CH_STEP) bestL = v; }
  set_split(p.d[0], B, Q, bestL);
  set_split(p.d[1], B, Q, bestL);
}

}  // namespace ps

using namespace ps;

namespace ps {
// Forward (+ optional fused loss sums, + optional publication of those sums to the peers' mailboxes)
int chamfer_fwd_impl(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int* idx1, int* idx2,
                     double* sums6, const ps_comm* comm, int B, int N, int M, int dev, void* stream_, const char* who) {
  PS_REQUIRE(B >= 0 && N > 0 && M > 0, "%s: bad sizes B=%d N=%d M=%d", who, B, N, M);
  if (B == 0) {
    if (sums6) {
      PS_CUDA(cudaMemsetAsync(sums6, 0, 6 * sizeof(double), (cudaStream_t)stream_));
      if (comm) return comm_publish_launch(comm, sums6, 6, (cudaStream_t)stream_);
    }
    return PS_OK;
  }
  PS_REQUIRE(xyz1 && xyz2 && dist1 && dist2 && idx1 && idx2, "%s: null pointer", who);
  PS_REQUIRE((long long)B * N < (1ll << 31) / 3 * 3 && (long long)B * M < (1ll << 31) / 3 * 3,
             "%s: B*N or B*M too large", who);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "%s: cannot select device %d", who, dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nsm = sm_count(dev);

  {
    const int rc = chamfer_fwd_symmetric(xyz1, xyz2, dist1, dist2, idx1, idx2, sums6, comm, B, N, M, dev, stream);
    if (rc <= 0) return rc;  // handled (PS_OK) or failed; 1 = shape better served by the two-pass kernel
  }
  const int maxq = N > M ? N : M;
  const int Q = maxq <= 256 ? 1 : (maxq <= 1024 ? 2 : 4);
  ChamferParams p;
  p.d[0].q = xyz1; p.d[0].t = xyz2; p.d[0].dist = dist1; p.d[0].idx = idx1; p.d[0].nq = N; p.d[0].nt = M;
  p.d[1].q = xyz2; p.d[1].t = xyz1; p.d[1].dist = dist2; p.d[1].idx = idx2; p.d[1].nq = M; p.d[1].nt = N;
  plan_units(p, B, Q, nsm);

  // merge scratch for split directions
  size_t need[2] = {0, 0};
  for (int d = 0; d < 2; d++)
    if (p.d[d].nsplit > 1) need[d] = (size_t)B * p.d[d].nq;
  ScratchGuard scratch_mem;
  u64* scratch = nullptr;
  if (need[0] + need[1]) {
    if (int rc = scratch_mem.alloc((need[0] + need[1]) * sizeof(u64), dev, stream)) return rc;
    scratch = static_cast<u64*>(scratch_mem.ptr);
    if (int rc = fill32_async(scratch, 0xffffffffu, (need[0] + need[1]) * sizeof(u64), stream)) return rc;
  }
  p.d[0].keys = need[0] ? scratch : nullptr;
  p.d[1].keys = need[1] ? scratch + need[0] : nullptr;

  const int grid = p.d[0].units + p.d[1].units;
  switch (Q) {
    case 1: chamfer_nn_kernel<1><<<grid, CH_THREADS, 0, stream>>>(p); break;
    case 2: chamfer_nn_kernel<2><<<grid, CH_THREADS, 0, stream>>>(p); break;
    default: chamfer_nn_kernel<4><<<grid, CH_THREADS, 0, stream>>>(p); break;
  }
  PS_LAUNCH_CHECK();
  for (int d = 0; d < 2; d++) {
    if (!need[d]) continue;
    chamfer_unpack_kernel<<<ceil_div(need[d], 256), 256, 0, stream>>>(p.d[d].keys, p.d[d].dist, p.d[d].idx, need[d]);
    PS_LAUNCH_CHECK();
  }
  if (int rc = scratch_mem.release()) return rc;
  if (sums6) {
    if (int rc = chamfer_sums_launch(dist1, dist2, sums6, (long long)B * N, (long long)B * M, 0, dev, stream)) return rc;
    if (comm) return comm_publish_launch(comm, sums6, 6, stream);
  }
  return PS_OK;
}
}  // namespace ps

extern "C" int ps_chamfer_fwd(const float* xyz1, const float* xyz2, float* dist1, float* dist2,
                              int* idx1, int* idx2, int B, int N, int M, int dev, void* stream) {
  return ps::chamfer_fwd_impl(xyz1, xyz2, dist1, dist2, idx1, idx2, nullptr, nullptr, B, N, M, dev, stream, "ps_chamfer_fwd");
}

extern "C" int ps_chamfer_fwd_sums(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int* idx1,
                                   int* idx2, double* sums6, int B, int N, int M, int dev, void* stream) {
  PS_REQUIRE(sums6 != nullptr, "ps_chamfer_fwd_sums: null sums6");
  return ps::chamfer_fwd_impl(xyz1, xyz2, dist1, dist2, idx1, idx2, sums6, nullptr, B, N, M, dev, stream, "ps_chamfer_fwd_sums");
}

extern "C" int ps_chamfer_bwd(const float* xyz1, const float* xyz2, const float* graddist1,
                              const float* graddist2, const int* idx1, const int* idx2,
                              float* gradxyz1, float* gradxyz2, int B, int N, int M, int dev,
                              void* stream_) {
  PS_REQUIRE(B >= 0 && N > 0 && M > 0, "ps_chamfer_bwd: bad sizes B=%d N=%d M=%d", B, N, M);
  if (B == 0) return PS_OK;
  PS_REQUIRE(xyz1 && xyz2 && graddist1 && graddist2 && idx1 && idx2 && gradxyz1 && gradxyz2,
             "ps_chamfer_bwd: null pointer");
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_chamfer_bwd: cannot select device %d", dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  ChamferBwdParams p{xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1, gradxyz2, B, N, M};
  const size_t tot = (size_t)B * N + (size_t)B * M;
  const int grid = ceil_div(tot, 256);
  chamfer_bwd_kernel<false><<<grid, 256, 0, stream>>>(p);
  PS_LAUNCH_CHECK();
  chamfer_bwd_kernel<true><<<grid, 256, 0, stream>>>(p);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

// ---- fused loss/metric partial sums ---------------------------------------------------------------
// The reductions the callers apply to Chamfer's outputs (utils/loss_utils.py:10-31,98-103:
// mean(d), mean(sqrt(d)) per side) as ONE launch instead of ~10 small torch kernels:
// out[0] = sum sqrt(dist1), out[1] = sum sqrt(dist2), out[2] = sum dist1, out[3] = sum dist2,
// accumulated in double; out[4], out[5] = the element counts.  These four numbers (plus the element counts) are exactly what the
// multi-GPU path all-reduces.
namespace ps {
__global__ void __launch_bounds__(256) chamfer_sums_kernel(const float* __restrict__ d1, const float* __restrict__ d2,
                                                           double* __restrict__ out, long long n1, long long n2, int accumulate) {
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += stride) {
    const float v = __ldg(d1 + i);
    s[0] += (double)sqrtf(v);
    s[2] += (double)v;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const float v = __ldg(d2 + i);
    s[1] += (double)sqrtf(v);
    s[3] += (double)v;
  }
  __shared__ double sh[8][4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[warp][0] = s[0]; sh[warp][1] = s[1]; sh[warp][2] = s[2]; sh[warp][3] = s[3]; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += sh[w][threadIdx.x];
    atomicAdd(out + threadIdx.x, t);
  }
  if (blockIdx.x == 0 && threadIdx.x == 4) {  // element counts
    if (accumulate) { atomicAdd(out + 4, (double)n1); atomicAdd(out + 5, (double)n2); }
    else { out[4] = (double)n1; out[5] = (double)n2; }
  }
}

// accumulate != 0: adds this call's six numbers to `out` (the chunked host pipeline zeroes it once per step)
int chamfer_sums_launch(const float* dist1, const float* dist2, double* out6, long long n1, long long n2, int accumulate,
                        int dev, cudaStream_t stream) {
  if (!accumulate) PS_CUDA(cudaMemsetAsync(out6, 0, 6 * sizeof(double), stream));
  const long long n = n1 > n2 ? n1 : n2;
  if (n == 0) return PS_OK;
  int grid = ceil_div(n, 256 * 8);
  const int cap = sm_count(dev) * 4;
  if (grid > cap) grid = cap;
  chamfer_sums_kernel<<<grid, 256, 0, stream>>>(dist1, dist2, out6, n1, n2, accumulate);
  PS_LAUNCH_CHECK();
  return PS_OK;
}
}  // namespace ps

extern "C" int ps_chamfer_sums(const float* dist1, const float* dist2, double* out4, long long n1,
                               long long n2, int dev, void* stream_) {
  PS_REQUIRE(n1 >= 0 && n2 >= 0 && out4 && (n1 == 0 || dist1) && (n2 == 0 || dist2), "ps_chamfer_sums: bad arguments");
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_chamfer_sums: cannot select device %d", dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  return chamfer_sums_launch(dist1, dist2, out4, n1, n2, 0, dev, stream);
}
