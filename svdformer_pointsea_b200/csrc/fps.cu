// Furthest point sampling for sm_100a: one thread-block cluster per cloud.
//
// Replaces furthest_point_sampling_kernel (pointnet2_ops/_ext-src/src/sampling_gpu.cu:69-173).
//
// The reference keeps the running min-distance array `temp` in GLOBAL memory and re-reads the
// coordinates every iteration (sampling_gpu.cu:95-110), with one 512-thread block per cloud and
// a 9-level __syncthreads tree per iteration.  Here:
//   * a cluster of C CTAs x 512 threads owns one cloud; thread g = ctarank*512 + tid owns the
//     points k = g + i*C*512 (i < P); their coordinates AND their running min distance live in
//     registers for the whole kernel.  Global memory is read once.
//   * per iteration every thread updates its P distances and keeps its best; a warp resolves
//     its best with two REDUX ops (max of the distance bits, then min of the tie-break rank);
//   * MODE 0 (C == 1): warp winners meet in shared memory, one __syncthreads per iteration;
//   * MODE 1 (C <= 4): every warp winner is pushed straight into the slot arrays of ALL CTAs of
//     the cluster through distributed shared memory (st.shared::cluster), then ONE
//     barrier.cluster (arrive.release / wait.acquire) per iteration, then every warp reduces the
//     16*C candidates redundantly, so no second exchange is needed;
//   * MODE 2 (C >= 8): CTA-level reduce first (one __syncthreads), then one candidate per CTA is
//     pushed to all CTAs, then the cluster barrier.
//   The message carries the winner's coordinates, so the next iteration starts without touching
//   global memory.
//
// Bit-exact tie-breaking.  The reference's winner among equal distances is decided by its
// per-thread strict `>` scan (lowest k within thread tid = k mod bs) followed by the shared-
// memory tree whose __update keeps the LEFT operand on ties (sampling_gpu.cu:59-65,115-168):
// the tied thread with the smallest bit-reversed tid wins.  We reproduce that order with
//     rank(k) = bitrev_L(k mod bs) * ceil(N / bs) + floor(k / bs),   bs = 2^L = opt_n_threads(N)
// and select (distance desc, rank asc).  Points with |p|^2 <= 1e-3 (compared in double, as
// the reference does, sampling_gpu.cu:100-101) never update and are never selected; they are
// carried as distance -1, which orders below every real distance when the fp32 bit patterns
// are compared as signed integers.
#include "common.cuh"

#include <cmath>
#include <cstdlib>

namespace ps {

constexpr int FPS_T = 512;
constexpr int FPS_WARPS = FPS_T / 32;

struct __align__(16) FpsEntry {  // two 16-byte vectors, each written by ONE st.v4 and carrying a tag
  int tb;          // distance bits (signed compare)
  unsigned rank;   // tie-break rank, smaller wins
  float x;
  unsigned tag_a;  // iteration number (flag-polling modes)
  float y, z;
  int k;           // point index
  unsigned tag_b;
};

struct FpsArgs {
  const float* xyz;
  int* idx;
  int N, npoint;
  int L;     // log2(bs) of the reference launch
  int nper;  // ceil(N / bs)
};

__device__ __forceinline__ unsigned fps_rank(int k, int L, int nper) {
  const unsigned low = (unsigned)k & ((1u << L) - 1u);
  const unsigned rev = L ? (__brev(low) >> (32 - L)) : 0u;
  return rev * (unsigned)nper + ((unsigned)k >> L);
}

// lexicographic "a better than b": larger tb, then smaller rank
__device__ __forceinline__ bool fps_better(int tb_a, unsigned r_a, int tb_b, unsigned r_b) {
  return tb_a > tb_b || (tb_a == tb_b && r_a < r_b);
}

// Warp-wide argbest over (tb desc, rank asc); returns the winning lane.
__device__ __forceinline__ int warp_argbest(int tb, unsigned rank, int& tb_max) {
  tb_max = __reduce_max_sync(0xffffffffu, tb);
  const unsigned r = (tb == tb_max) ? rank : 0xffffffffu;
  const unsigned rmin = __reduce_min_sync(0xffffffffu, r);
  const unsigned m = __ballot_sync(0xffffffffu, r == rmin);
  return __ffs(m) - 1;
}

// ---- st.async + mbarrier: the remote store and its completion signal travel together -------------
__device__ __forceinline__ void fps_mbar_init(unsigned bar, unsigned cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt) : "memory");
}
__device__ __forceinline__ void fps_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool fps_mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void st_async_v4(unsigned raddr, unsigned rbar, unsigned a, unsigned b, unsigned c, unsigned d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar) : "memory");
}

// MODE 0: single CTA, __syncthreads.
// MODE 1: cluster, every warp pushes to all CTAs, barrier.cluster per iteration.
// MODE 2: cluster, CTA-level reduce then one push per CTA, barrier.cluster per iteration.
// MODE 3: as 1 but NO cluster barrier (barrier.cluster.arrive.release costs a MEMBAR.ALL.GPU per
//         iteration): candidates are pushed with st.async, whose completion is counted in bytes
//         on the DESTINATION CTA's mbarrier; every warp sleeps on its own CTA's mbarrier
//         (try_wait) until all 16*C candidates of the iteration have landed.
// MODE 4: as 2 with the cluster exchange done by st.async + mbarrier.
// Slot reuse is safe with two buffers in every mode: a writer can only be at iteration j+2 after
// it has consumed every candidate of iteration j+1, and each reader sends its j+1 candidate only
// after it has finished reading the slots of iteration j.
template <int P, int MODE>
__global__ void __launch_bounds__(FPS_T, 1) fps_kernel(const FpsArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: px[P*T] py[P*T] pz[P*T] | mbar[2] (16 B) | slots[2][E] | wslots[2][16] (MODE 2/4)
  float* px = reinterpret_cast<float*>(smem_raw);
  float* py = px + P * FPS_T;
  float* pz = py + P * FPS_T;
  u64* mbar = reinterpret_cast<u64*>(pz + P * FPS_T);
  FpsEntry* slots = reinterpret_cast<FpsEntry*>(mbar + 2);

  constexpr bool CLUSTER = MODE != 0;
  constexpr bool DIRECT = MODE == 1 || MODE == 3;
  constexpr bool POLL = MODE == 3 || MODE == 4;
  const unsigned C = CLUSTER ? cluster_nctarank() : 1u;
  const unsigned crank = CLUSTER ? cluster_ctarank() : 0u;
  const int b = CLUSTER ? (int)cluster_id_x() : (int)blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int E = DIRECT ? FPS_WARPS * (int)C : (CLUSTER ? (int)C : FPS_WARPS);
  FpsEntry* wslots = slots + 2 * E;  // CTA-level staging (MODE 2/4)

  const int N = a.N;
  const float* cloud = a.xyz + (size_t)b * N * 3;
  int* out = a.idx + (size_t)b * a.npoint;
  const int g = (int)crank * FPS_T + tid;
  const int stride = (int)C * FPS_T;

  constexpr int PP = (P + 1) / 2;  // packed pairs (P == 1 keeps its second half inert)
  u64 x2[PP], y2[PP], z2[PP];
  float t[2 * PP];
#pragma unroll
  for (int i = 0; i < 2 * PP; i++) {
    const int k = g + i * stride;
    float xv = 0.f, yv = 0.f, zv = 0.f;
    t[i] = -1.0f;
    if (i < P && k < N) {
      xv = __ldg(cloud + (size_t)k * 3 + 0);
      yv = __ldg(cloud + (size_t)k * 3 + 1);
      zv = __ldg(cloud + (size_t)k * 3 + 2);
      const float mag = dist2_ref(xv, yv, zv);
      t[i] = ((double)mag <= 1e-3) ? -1.0f : 1e10f;
    }
    if (i < P) {
      px[i * FPS_T + tid] = xv;
      py[i * FPS_T + tid] = yv;
      pz[i * FPS_T + tid] = zv;
    }
    if (i & 1) {
      x2[i / 2] = pack2(lo2(x2[i / 2]), xv); y2[i / 2] = pack2(lo2(y2[i / 2]), yv); z2[i / 2] = pack2(lo2(z2[i / 2]), zv);
    } else {
      x2[i / 2] = pack2(xv, 0.f); y2[i / 2] = pack2(yv, 0.f); z2[i / 2] = pack2(zv, 0.f);
    }
  }
  if (POLL && tid == 0) {
    // one local arrive (the expect_tx below) + E*32 bytes of st.async traffic complete a phase
    fps_mbar_init(smem_u32(&mbar[0]), 1);
    fps_mbar_init(smem_u32(&mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fps_mbar_expect_tx(smem_u32(&mbar[0]), (unsigned)E * 32u);
    fps_mbar_expect_tx(smem_u32(&mbar[1]), (unsigned)E * 32u);
  }
  const float p0x = __ldg(cloud + 0), p0y = __ldg(cloud + 1), p0z = __ldg(cloud + 2);
  float lx = p0x, ly = p0y, lz = p0z;
  if (g == 0 && a.npoint > 0) out[0] = 0;
  if (CLUSTER) { cluster_arrive_release(); cluster_wait_acquire(); }  // peers resident, tags cleared
  else __syncthreads();

  for (int j = 1; j < a.npoint; j++) {
    // ---- per-thread update + best: d = |p - last|^2 two points at a time (FADD2/FMUL2/FFMA2) --
    const u64 nlx = pack2(-lx, -lx), nly = pack2(-ly, -ly), nlz = pack2(-lz, -lz);
    float best = -1.0f;
    int bi = 0;
#pragma unroll
    for (int i = 0; i < PP; i++) {
      const u64 d = dist2x2(x2[i], y2[i], z2[i], nlx, nly, nlz);
      t[2 * i] = fminf(lo2(d), t[2 * i]);
      t[2 * i + 1] = fminf(hi2(d), t[2 * i + 1]);  // inert halves stay at -1
      if (t[2 * i] > best) { best = t[2 * i]; bi = 2 * i; }
      if (t[2 * i + 1] > best) { best = t[2 * i + 1]; bi = 2 * i + 1; }
    }
    const int k_mine = g + bi * stride;
    const int tb = __float_as_int(best);
    const unsigned rk = fps_rank(k_mine, a.L, a.nper);
    int tbw;
    const int wl = warp_argbest(tb, rk, tbw);
    const int buf = j & 1;
    const unsigned tag = (unsigned)j;

    if (DIRECT) {
      // the winning lane pushes its candidate to slot (crank*16+warp) of every CTA
      if (lane == wl) {
        const unsigned e = crank * FPS_WARPS + warp;
        const unsigned local = smem_u32(&slots[buf * E + e]);
        const int pi = (bi < P ? bi : 0) * FPS_T + tid;
        const float wx = px[pi], wy = py[pi], wz = pz[pi];
        const unsigned lbar = smem_u32(&mbar[buf]);
        for (unsigned c = 0; c < C; c++) {
          const unsigned ra = mapa_shared(local, c);
          if (POLL) {
            const unsigned rb = mapa_shared(lbar, c);
            st_async_v4(ra, rb, (unsigned)tb, rk, __float_as_uint(wx), tag);
            st_async_v4(ra + 16, rb, __float_as_uint(wy), __float_as_uint(wz), (unsigned)k_mine, tag);
          } else {
            st_cluster_v4(ra, (unsigned)tb, rk, __float_as_uint(wx), tag);
            st_cluster_v4(ra + 16, __float_as_uint(wy), __float_as_uint(wz), (unsigned)k_mine, tag);
          }
        }
      }
      if (!POLL) { cluster_arrive_release(); cluster_wait_acquire(); }
    } else {
      // CTA-level staging
      FpsEntry* ws = (CLUSTER ? wslots : slots) + buf * FPS_WARPS;
      if (lane == wl) {
        const int pi = (bi < P ? bi : 0) * FPS_T + tid;
        FpsEntry en;
        en.tb = tb; en.rank = rk; en.k = k_mine;
        en.x = px[pi]; en.y = py[pi]; en.z = pz[pi];
        en.tag_a = en.tag_b = tag;
        ws[warp] = en;
      }
      __syncthreads();
      if (CLUSTER) {
        // every warp reduces the 16 warp winners; warp 0 forwards the CTA winner to all CTAs
        FpsEntry en;
        en.tb = (int)0x80000000; en.rank = 0xffffffffu; en.x = en.y = en.z = 0.f; en.k = 0;
        if (lane < FPS_WARPS) en = ws[lane];
        int tbc;
        const int cl = warp_argbest(en.tb, en.rank, tbc);
        if (warp == 0) {
          const unsigned stb = (unsigned)tbc;
          const unsigned srk = __shfl_sync(0xffffffffu, en.rank, cl);
          const unsigned sx = __shfl_sync(0xffffffffu, __float_as_uint(en.x), cl);
          const unsigned sy = __shfl_sync(0xffffffffu, __float_as_uint(en.y), cl);
          const unsigned sz = __shfl_sync(0xffffffffu, __float_as_uint(en.z), cl);
          const unsigned sk = __shfl_sync(0xffffffffu, (unsigned)en.k, cl);
          if ((unsigned)lane < C) {
            const unsigned ra = mapa_shared(smem_u32(&slots[buf * E + crank]), (unsigned)lane);
            if (POLL) {
              const unsigned rb = mapa_shared(smem_u32(&mbar[buf]), (unsigned)lane);
              st_async_v4(ra, rb, stb, srk, sx, tag);
              st_async_v4(ra + 16, rb, sy, sz, sk, tag);
            } else {
              st_cluster_v4(ra, stb, srk, sx, tag);
              st_cluster_v4(ra + 16, sy, sz, sk, tag);
            }
          }
        }
        if (!POLL) { cluster_arrive_release(); cluster_wait_acquire(); }
      }
    }

    // ---- final reduce over the E candidates (identical in every warp of every CTA) --------
    int ftb = (int)0x80000000;
    unsigned frk = 0xffffffffu;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    int fk = 0;
    const FpsEntry* sl = slots + buf * E;
    if (POLL) {
      // buffer `buf` is used by iterations buf, buf+2, ... (buf=1: j=1,3,..; buf=0: j=2,4,..)
      const unsigned parity = (unsigned)((j - 1) >> 1) & 1u;
      while (!fps_mbar_try_wait(smem_u32(&mbar[buf]), parity)) {}
      // re-arm for iteration j+2.  Safe before our other warps have observed this phase: the next
      // phase cannot complete until they have sent their j+1 and j+2 candidates.
      if (tid == 0) fps_mbar_expect_tx(smem_u32(&mbar[buf]), (unsigned)E * 32u);
    }
    for (int e = lane; e < E; e += 32) {
      const int4 va = *reinterpret_cast<const int4*>(&sl[e]);
      const int4 vb = *reinterpret_cast<const int4*>(reinterpret_cast<const char*>(&sl[e]) + 16);
      if (fps_better(va.x, (unsigned)va.y, ftb, frk)) {
        ftb = va.x; frk = (unsigned)va.y; fx = __int_as_float(va.z);
        fy = __int_as_float(vb.x); fz = __int_as_float(vb.y); fk = vb.z;
      }
    }
    int tbf;
    const int fl = warp_argbest(ftb, frk, tbf);
    lx = __shfl_sync(0xffffffffu, fx, fl);
    ly = __shfl_sync(0xffffffffu, fy, fl);
    lz = __shfl_sync(0xffffffffu, fz, fl);
    int kf = __shfl_sync(0xffffffffu, fk, fl);
    if (tbf < 0) {  // no eligible point anywhere: the reference's tree returns thread 0's besti = 0
      kf = 0; lx = p0x; ly = p0y; lz = p0z;
    }
    if (g == 0) out[j] = kf;
  }
  if (CLUSTER) { cluster_arrive_release(); cluster_wait_acquire(); }  // no CTA exits while peers may still write to it
}

// Generic fallback for clouds too large for the register-resident kernel: one CTA per cloud,
// running min distances in a global scratch array (still the exact reference order).
__global__ void __launch_bounds__(FPS_T, 1) fps_generic_kernel(const FpsArgs a, float* temp_all) {
  __shared__ FpsEntry ws[2][FPS_WARPS];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const float* cloud = a.xyz + (size_t)b * N * 3;
  float* temp = temp_all + (size_t)b * N;
  int* out = a.idx + (size_t)b * a.npoint;
  for (int k = tid; k < N; k += FPS_T) {
    const float mag = dist2_ref(cloud[k * 3 + 0], cloud[k * 3 + 1], cloud[k * 3 + 2]);
    temp[k] = ((double)mag <= 1e-3) ? -1.0f : 1e10f;
  }
  float lx = cloud[0], ly = cloud[1], lz = cloud[2];
  if (tid == 0 && a.npoint > 0) out[0] = 0;
  __syncthreads();
  for (int j = 1; j < a.npoint; j++) {
    float best = -1.0f;
    int bk = tid < N ? tid : 0;
    for (int k = tid; k < N; k += FPS_T) {
      const float d = dist2_ref(cloud[k * 3 + 0] - lx, cloud[k * 3 + 1] - ly, cloud[k * 3 + 2] - lz);
      const float t = fminf(d, temp[k]);
      temp[k] = t;
      if (t > best) { best = t; bk = k; }
    }
    int tbw;
    const unsigned rk = fps_rank(bk, a.L, a.nper);
    const int wl = warp_argbest(__float_as_int(best), rk, tbw);
    const int buf = j & 1;
    if (lane == wl) {
      FpsEntry en;
      en.tb = __float_as_int(best); en.rank = rk; en.k = bk;
      en.x = cloud[bk * 3 + 0]; en.y = cloud[bk * 3 + 1]; en.z = cloud[bk * 3 + 2];
      en.tag_a = en.tag_b = 0u;
      ws[buf][warp] = en;
    }
    __syncthreads();
    FpsEntry en;
    en.tb = (int)0x80000000; en.rank = 0xffffffffu; en.x = en.y = en.z = 0.f; en.k = 0;
    if (lane < FPS_WARPS) en = ws[buf][lane];
    int tbf;
    const int fl = warp_argbest(en.tb, en.rank, tbf);
    lx = __shfl_sync(0xffffffffu, en.x, fl);
    ly = __shfl_sync(0xffffffffu, en.y, fl);
    lz = __shfl_sync(0xffffffffu, en.z, fl);
    int kf = __shfl_sync(0xffffffffu, en.k, fl);
    if (tbf < 0) { kf = 0; lx = cloud[0]; ly = cloud[1]; lz = cloud[2]; }
    if (tid == 0) out[j] = kf;
  }
}

// opt_n_threads of the reference (include/cuda_utils.h:15-19), evaluated the same way
// (double log ratio truncated) so bs matches even where the quotient is inexact.
static int ref_block_log2(int n) {
  const int pow_2 = (int)(std::log(static_cast<double>(n)) / std::log(2.0));
  int bs = 1 << pow_2;
  if (bs > 512) bs = 512;
  if (bs < 1) bs = 1;
  int L = 0;
  while ((1 << L) < bs) L++;
  return L;
}

template <int P, int MODE>
static int launch_fps(const FpsArgs& a, int B, int C, cudaStream_t stream) {
  const int E = (MODE == 1 || MODE == 3) ? FPS_WARPS * C : (MODE == 0 ? FPS_WARPS : C);
  const size_t smem = (size_t)3 * P * FPS_T * sizeof(float) + 16 + (size_t)2 * E * sizeof(FpsEntry) +
                      ((MODE == 2 || MODE == 4) ? (size_t)2 * FPS_WARPS * sizeof(FpsEntry) : 0);
  auto kern = fps_kernel<P, MODE>;
  if (smem > 48 * 1024) PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (MODE == 0) {
    kern<<<B, FPS_T, smem, stream>>>(a);
  } else {
    if (C > 8) PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * C);
    cfg.blockDim = dim3(FPS_T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    PS_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  }
  PS_LAUNCH_CHECK();
  return PS_OK;
}

template <int MODE>
static int dispatch_p(int P, const FpsArgs& a, int B, int C, cudaStream_t s) {
  switch (P) {
    case 1: return launch_fps<1, MODE>(a, B, C, s);
    case 2: return launch_fps<2, MODE>(a, B, C, s);
    case 4: return launch_fps<4, MODE>(a, B, C, s);
    case 8: return launch_fps<8, MODE>(a, B, C, s);
    case 16: return launch_fps<16, MODE>(a, B, C, s);
  }
  return set_error(PS_ERR_UNSUPPORTED, "ps_fps: no kernel for P=%d", P);
}

// Per-iteration cost model in SM cycles (measured on B200, profiles/microbench_r1.jsonl):
// ~32 cycles of issue per resident point-per-thread, ~250 for the single-CTA exchange,
// ~700 for the cluster exchange (barrier.cluster alone is ~450-480).
static int choose_cluster(int B, int N, int nsm) {
  if (const char* e = getenv("PS_FPS_CLUSTER")) {
    const int c = atoi(e);
    if (c == 1 || c == 2 || c == 4 || c == 8 || c == 16) return c;
  }
  double best_cost = 1e30;
  int best_c = 1;
  for (int c = 1; c <= 16; c *= 2) {
    const int p = ceil_div(N, c * FPS_T);
    if (c > 1 && p < 1) break;
    if (p > 16) continue;
    const int waves = ceil_div((long long)B * c, nsm);
    const double cost = waves * (32.0 * p + (c == 1 ? 250.0 : 700.0));
    if (cost < best_cost - 1e-9) { best_cost = cost; best_c = c; }
  }
  return best_c;
}

}  // namespace ps

using namespace ps;

extern "C" int ps_fps(const float* xyz, int* idx, int B, int N, int npoint, int dev, void* stream_) {
  PS_REQUIRE(B >= 0 && N > 0 && npoint >= 0, "ps_fps: bad sizes B=%d N=%d npoint=%d", B, N, npoint);
  if (B == 0 || npoint == 0) return PS_OK;
  PS_REQUIRE(xyz && idx, "ps_fps: null pointer");
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_fps: cannot select device %d", dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nsm = sm_count(dev);

  FpsArgs a;
  a.xyz = xyz; a.idx = idx; a.N = N; a.npoint = npoint;
  a.L = ref_block_log2(N);
  a.nper = ceil_div(N, 1 << a.L);

  if (ceil_div(N, 16 * FPS_T) > 16) {
    // beyond the register-resident kernels: generic one-CTA-per-cloud path with global scratch
    float* temp = nullptr;
    if (int rc = scratch_alloc((void**)&temp, (size_t)B * N * sizeof(float), dev, stream)) return rc;
    fps_generic_kernel<<<B, FPS_T, 0, stream>>>(a, temp);
    PS_LAUNCH_CHECK();
    PS_CUDA(cudaFreeAsync(temp, stream));
    return PS_OK;
  }
  const int C = choose_cluster(B, N, nsm);
  int P = 1;
  while (P * C * FPS_T < N) P *= 2;
  // PS_FPS_SYNC=barrier selects the barrier.cluster variants (kept for A/B measurements)
  const char* sync_env = getenv("PS_FPS_SYNC");
  const bool poll = !(sync_env && sync_env[0] == 'b');
  if (C == 1) return dispatch_p<0>(P, a, B, C, stream);
  if (C <= 4) return poll ? dispatch_p<3>(P, a, B, C, stream) : dispatch_p<1>(P, a, B, C, stream);
  return poll ? dispatch_p<4>(P, a, B, C, stream) : dispatch_p<2>(P, a, B, C, stream);
}
