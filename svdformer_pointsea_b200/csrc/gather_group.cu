// gather_operation / grouping_operation forward + backward for sm_100a (HBM-bound).
//
// Replaces gather_points(_grad)_kernel (pointnet2_ops/_ext-src/src/sampling_gpu.cu:8-57) and
// group_points(_grad)_kernel (group_points_gpu.cu:8-75).  Grouping IS gathering with the
// (S,K) index tensor flattened to M' = S*K positions, so both ops share the kernels below.
//
//   forward   out[b,c,p] = feat[b,c,idx[b,p]]             feat (B,C,N), idx (B,M'), out (B,C,M')
//   backward  gfeat[b,c,idx[b,p]] += gout[b,c,p]
//
// Algorithmic HBM bytes: 4*(B*M' + B*C*N + B*C*M') for group, 4*(B*M' + 2*B*C*M')
// for a sparse gather.  The reference launches only B blocks for group_points and writes 4-byte
// stores nsample floats apart; here:
//   * "staged" kernel (output >> input): a CTA owns (cloud, CT-channel tile, position slice).
//     The CT feature rows (contiguous N floats each) are pulled into shared memory with 1-D bulk
//     async copies (cp.async.bulk, i.e. the TMA engine, completion on an mbarrier); every
//     thread then loads 4 consecutive indices once (128-bit), gathers from shared memory and
//     emits one 128-bit coalesced store per channel.  The feature tensor is read from HBM
//     once, the index tensor once per channel tile, the output is written once.
//   * "direct" kernel (sparse gather, or rows too long for shared memory): same thread
//     mapping, gathers through L1/L2 with read-only loads.
//   * backward mirrors it: the CT accumulator rows live in shared memory (RED.shared adds),
//     and are written out once with plain coalesced stores (no global atomics, no pre-zeroed
//     buffer); rows too long for shared memory fall back to memset + global RED.
#include "common.cuh"

#include <cstdlib>

namespace ps {

constexpr int GG_THREADS = 256;

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- forward, shared-memory staged -----------------------------------------------------------
// grid = (position slices, channel tiles, B); dynamic smem = CT * Npad floats (+ mbarrier)
template <int CT>
__global__ void __launch_bounds__(GG_THREADS) gather_staged_kernel(
    const float* __restrict__ feat, const int* __restrict__ idx, float* __restrict__ out, int C,
    int N, int Mp, int slice_len, int use_bulk, int vec_ok) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  u64* bar = reinterpret_cast<u64*>(smem_raw);
  float* rows = reinterpret_cast<float*>(smem_raw + 128);
  const int b = blockIdx.z, c0 = blockIdx.y * CT, tid = threadIdx.x;
  const int nrows = min(CT, C - c0);
  const float* src = feat + ((size_t)b * C + c0) * N;

  if (use_bulk) {
    if (tid == 0) {
      mbar_init(smem_u32(bar), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(smem_u32(bar), (unsigned)(nrows * N * 4));
      // rows are contiguous in global memory: one bulk copy per row keeps each request <= 64 KB
      for (int r = 0; r < nrows; r++)
        bulk_g2s(smem_u32(rows + (size_t)r * N), src + (size_t)r * N, (unsigned)(N * 4), smem_u32(bar));
    }
  } else {
    for (int i = tid; i < nrows * N; i += GG_THREADS) rows[i] = __ldg(src + i);
  }

  // indices for this thread's 4 positions (loaded while the bulk copies are in flight)
  const int p_begin = blockIdx.x * slice_len;
  const int p_end = min(Mp, p_begin + slice_len);
  const int* ib = idx + (size_t)b * Mp;
  const bool vec = vec_ok != 0;  // Mp % 4 == 0 and 16-byte aligned idx/out bases (host-checked)

  if (use_bulk) {
    while (!mbar_try_wait(smem_u32(bar), 0)) {}
  } else {
    __syncthreads();
  }

  for (int p = p_begin + tid * 4; p < p_end; p += GG_THREADS * 4) {
    int i0, i1, i2, i3;
    const bool full = vec && (p + 4 <= p_end);
    if (full) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(ib + p));
      i0 = v.x; i1 = v.y; i2 = v.z; i3 = v.w;
    } else {
      i0 = __ldg(ib + p);
      i1 = p + 1 < p_end ? __ldg(ib + p + 1) : 0;
      i2 = p + 2 < p_end ? __ldg(ib + p + 2) : 0;
      i3 = p + 3 < p_end ? __ldg(ib + p + 3) : 0;
    }
    float* ob = out + ((size_t)b * C + c0) * Mp + p;
#pragma unroll
    for (int r = 0; r < CT; r++) {
      if (r < nrows) {
        const float* row = rows + (size_t)r * N;
        const float4 v = make_float4(row[i0], row[i1], row[i2], row[i3]);
        float* o = ob + (size_t)r * Mp;
        if (full) {
          __stcs(reinterpret_cast<float4*>(o), v);  // streaming store: output is written once
        } else {
          o[0] = v.x;
          if (p + 1 < p_end) o[1] = v.y;
          if (p + 2 < p_end) o[2] = v.z;
          if (p + 3 < p_end) o[3] = v.w;
        }
      }
    }
  }
}

// ---- forward, direct ----------------------------------------------------------------------------
// grid = (position blocks, channel tiles, B)
template <int CT>
__global__ void __launch_bounds__(GG_THREADS) gather_direct_kernel(
    const float* __restrict__ feat, const int* __restrict__ idx, float* __restrict__ out, int C,
    int N, int Mp, int vec_ok) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT;
  const int p = (blockIdx.x * GG_THREADS + threadIdx.x) * 4;
  if (p >= Mp) return;
  const int nrows = min(CT, C - c0);
  const int* ib = idx + (size_t)b * Mp;
  const bool full = (vec_ok != 0) && (p + 4 <= Mp);
  int i0, i1, i2, i3;
  if (full) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(ib + p));
    i0 = v.x; i1 = v.y; i2 = v.z; i3 = v.w;
  } else {
    i0 = __ldg(ib + p);
    i1 = p + 1 < Mp ? __ldg(ib + p + 1) : 0;
    i2 = p + 2 < Mp ? __ldg(ib + p + 2) : 0;
    i3 = p + 3 < Mp ? __ldg(ib + p + 3) : 0;
  }
  const float* src = feat + ((size_t)b * C + c0) * N;
  float* ob = out + ((size_t)b * C + c0) * Mp + p;
#pragma unroll
  for (int r = 0; r < CT; r++) {
    if (r < nrows) {
      const float* row = src + (size_t)r * N;
      const float4 v = make_float4(__ldg(row + i0), __ldg(row + i1), __ldg(row + i2), __ldg(row + i3));
      float* o = ob + (size_t)r * Mp;
      if (full) {
        *reinterpret_cast<float4*>(o) = v;
      } else {
        o[0] = v.x;
        if (p + 1 < Mp) o[1] = v.y;
        if (p + 2 < Mp) o[2] = v.z;
        if (p + 3 < Mp) o[3] = v.w;
      }
    }
  }
}

// ---- backward, shared-memory accumulators -------------------------------------------------------
// grid = (1, channel tiles, B); dynamic smem = CT * N floats
template <int CT>
__global__ void __launch_bounds__(GG_THREADS) scatter_staged_kernel(
    const float* __restrict__ gout, const int* __restrict__ idx, float* __restrict__ gfeat, int C,
    int N, int Mp, int vec_ok, size_t gbs) {
  extern __shared__ __align__(16) float acc[];
  const int b = blockIdx.z, c0 = blockIdx.y * CT, tid = threadIdx.x;
  const int nrows = min(CT, C - c0);
  for (int i = tid; i < nrows * N; i += GG_THREADS) acc[i] = 0.f;
  __syncthreads();
  const int* ib = idx + (size_t)b * Mp;
  const bool vec = vec_ok != 0;
  for (int p = tid * 4; p < Mp; p += GG_THREADS * 4) {
    int ii[4];
    const bool full = vec && (p + 4 <= Mp);
    if (full) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(ib + p));
      ii[0] = v.x; ii[1] = v.y; ii[2] = v.z; ii[3] = v.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; e++) ii[e] = p + e < Mp ? __ldg(ib + p + e) : -1;
    }
    const float* gb = gout + (size_t)b * gbs + (size_t)c0 * Mp + p;
#pragma unroll
    for (int r = 0; r < CT; r++) {
      if (r < nrows) {
        float v[4];
        if (full) {
          const float4 t = __ldcs(reinterpret_cast<const float4*>(gb + (size_t)r * Mp));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; e++) v[e] = p + e < Mp ? gb[(size_t)r * Mp + e] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (ii[e] >= 0) atomicAdd(&acc[(size_t)r * N + ii[e]], v[e]);
      }
    }
  }
  __syncthreads();
  float* dst = gfeat + ((size_t)b * C + c0) * N;
  for (int i = tid; i < nrows * N; i += GG_THREADS) dst[i] = acc[i];
}

// ---- backward, global atomics (rows too long for shared memory; gfeat pre-zeroed by us) --------
__global__ void __launch_bounds__(GG_THREADS) scatter_direct_kernel(
    const float* __restrict__ gout, const int* __restrict__ idx, float* __restrict__ gfeat, int C,
    int N, int Mp, size_t gbs) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int p = blockIdx.x * GG_THREADS + threadIdx.x;
  if (p >= Mp) return;
  const int i = __ldg(idx + (size_t)b * Mp + p);
  atomicAdd(gfeat + ((size_t)b * C + c) * N + i, __ldg(gout + (size_t)b * gbs + (size_t)c * Mp + p));
}

// ---- backward through an inverse index (dense grouping: every source point appears many times) ---
// grad_features[b,c,n] = sum over the positions p with idx[b,p] == n of grad_out[b,c,p].
// The index tensor is shared by all C channels, so its inverse is built ONCE per call and the
// reduction becomes a gather: no atomics in the channel loop, deterministic summation order
// (ascending p).  Positions are handled in chunks of CSR_CHUNK (one 64 KB shared-memory stage of a
// grad_out row); per (cloud, chunk) the inverse is a counting sort of the chunk's positions by
// source: `list` (u16 position-in-chunk, sorted by (source, position)) + `bnd` (N+1 offsets).
//   csr_build_kernel   one CTA per (chunk, cloud): histogram / scan / fill / per-list sort, all in
//                      shared memory, coalesced write-out
//   scatter_csr_kernel persistent CTA per (cloud, CPB channels): per chunk the list + offsets are
//                      staged in shared memory and each thread caches the entries of its NPT
//                      sources in registers; the CPB grad_out row chunks then stream through a
//                      2-stage ring filled by 1-D bulk async copies (TMA engine, mbarrier
//                      completion) and are gathered from shared memory into register accumulators.
constexpr int CSR_CHUNK = 16384;   // positions per stage (64 KB of fp32)
constexpr int CSR_THREADS = 1024;  // measured on C3: 512 thr 0.195 ms, 1024 thr 0.172 ms
constexpr int CSR_BUILD_THREADS = 1024;
constexpr int CSR_HEAVY = 32;      // a source with more entries than this in one chunk is summed by a whole warp
constexpr int CSR_MAX_HEAVY = CSR_CHUNK / (CSR_HEAVY + 1) + 1;

__global__ void __launch_bounds__(CSR_BUILD_THREADS) csr_build_kernel(const int* __restrict__ idx, int* __restrict__ bnd_all,
                                                                     unsigned short* __restrict__ list_all, int N, int Mp,
                                                                     int nchunks) {
  extern __shared__ int sm_i[];
  int* cnt = sm_i;             // N + 1
  int* cur = sm_i + (N + 1);   // N
  unsigned short* list = reinterpret_cast<unsigned short*>(cur + N);  // CSR_CHUNK
  __shared__ int warp_tot[CSR_BUILD_THREADS / 32];
  __shared__ int heavy[CSR_CHUNK / (CSR_HEAVY + 1) + 1];
  __shared__ int nheavy;
  const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p0 = k * CSR_CHUNK, len = min(CSR_CHUNK, Mp - p0);
  const int* ib = idx + (size_t)b * Mp + p0;
  int* bnd = bnd_all + ((size_t)b * nchunks + k) * (N + 1);
  unsigned short* out = list_all + (size_t)b * Mp + p0;
  for (int n = tid; n <= N; n += CSR_BUILD_THREADS) cnt[n] = 0;
  __syncthreads();
  for (int p = tid; p < len; p += CSR_BUILD_THREADS) atomicAdd(&cnt[__ldg(ib + p)], 1);
  __syncthreads();
  // exclusive scan of cnt[0..N): each thread owns a contiguous run
  const int per = (N + CSR_BUILD_THREADS - 1) / CSR_BUILD_THREADS;
  const int lo = min(N, tid * per), hi = min(N, lo + per);
  int run = 0;
  for (int n = lo; n < hi; n++) run += cnt[n];
  int incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
    warp_tot[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  int base = incl - run + (warp ? warp_tot[warp - 1] : 0);
  for (int n = lo; n < hi; n++) { const int c = cnt[n]; cur[n] = base; cnt[n] = base; base += c; }  // cnt now holds the offsets
  __syncthreads();
  if (tid == 0) { cnt[N] = len; nheavy = 0; }
  __syncthreads();
  // Light sources (<= CSR_HEAVY entries): cursor fill + a short insertion sort per list.  Heavy sources
  // (hubs of a kNN graph in feature space; index 0 of a zero-filled ball query) would make that sort
  // quadratic, so their lists are produced already sorted by a warp that scans the chunk in order.
  for (int p = tid; p < len; p += CSR_BUILD_THREADS) {
    const int src = __ldg(ib + p);
    if (cnt[src + 1] - cnt[src] <= CSR_HEAVY) list[atomicAdd(&cur[src], 1)] = (unsigned short)p;
  }
  __syncthreads();
  for (int n = tid; n < N; n += CSR_BUILD_THREADS) {  // ascending positions inside each list
    const int s = cnt[n], e = cnt[n + 1];
    if (e - s > CSR_HEAVY) { heavy[atomicAdd(&nheavy, 1)] = n; continue; }
    for (int i = s + 1; i < e; i++) {
      const unsigned short v = list[i];
      int j = i - 1;
      while (j >= s && list[j] > v) { list[j + 1] = list[j]; j--; }
      list[j + 1] = v;
    }
  }
  __syncthreads();
  for (int h = warp; h < nheavy; h += CSR_BUILD_THREADS / 32) {
    const int n = heavy[h];
    int base = cnt[n];
    for (int p0w = 0; p0w < len; p0w += 32) {
      const int p = p0w + lane;
      const bool hit = p < len && __ldg(ib + p) == n;
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (hit) list[base + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)p;
      base += __popc(mask);
    }
  }
  __syncthreads();
  for (int n = tid; n <= N; n += CSR_BUILD_THREADS) bnd[n] = cnt[n];
  for (int i = tid; i < len; i += CSR_BUILD_THREADS) out[i] = list[i];
}

// Hubs of one staged channel chunk: each warp sums the entries of a heavy source with lane-strided partial sums
// and a fixed shuffle tree (deterministic).  Kept out of line so that its registers do not add to the pressure
// of the unrolled channel loop (64-register budget at 1024 threads).
__device__ __noinline__ void csr_heavy_pass(const float* st, const unsigned short* list, const int* h_lo, const int* h_cnt,
                                            float* h_sum, int nh) {
  const int lane = threadIdx.x & 31;
  for (int h = threadIdx.x >> 5; h < nh; h += CSR_THREADS / 32) {
    const int hl = h_lo[h], hc = h_cnt[h];
    float hs = 0.f;
    for (int e = lane; e < hc; e += 32) hs += st[list[hl + e]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hs += __shfl_xor_sync(0xffffffffu, hs, o);
    if (lane == 0) h_sum[h] = hs;
  }
}

// NPT sources per thread, PF list entries per source cached in registers, CPB channels per CTA
template <int NPT, int PF, int CSR_CPB>
__global__ void __launch_bounds__(CSR_THREADS, 1) scatter_csr_kernel(const float* __restrict__ gout, const int* __restrict__ bnd_all,
                                                                     const unsigned short* __restrict__ list_all,
                                                                     float* __restrict__ gfeat, int C, int N, int Mp, int nchunks,
                                                                     size_t gbs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  u64* bars = reinterpret_cast<u64*>(smem_raw);          // full[2]
  float* stage0 = reinterpret_cast<float*>(smem_raw + 128);
  float* stage1 = stage0 + CSR_CHUNK;
  unsigned short* list = reinterpret_cast<unsigned short*>(stage1 + CSR_CHUNK);  // CSR_CHUNK
  int* bnd = reinterpret_cast<int*>(list + CSR_CHUNK);                           // N + 1
  __shared__ int h_lo[CSR_MAX_HEAVY], h_cnt[CSR_MAX_HEAVY];
  __shared__ float h_sum[2][CSR_MAX_HEAVY];
  __shared__ int nh;
  const int b = blockIdx.y, c0 = blockIdx.x * CSR_CPB, tid = threadIdx.x;
  const int nch = min(CSR_CPB, C - c0);
  const int items = nch * nchunks;
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int item) {  // thread 0 only; items are chunk-major: item = k * nch + c
    const int k = item / nch, c = c0 + item % nch;
    const int p0 = k * CSR_CHUNK, len = min(CSR_CHUNK, Mp - p0);
    const unsigned bar = smem_u32(&bars[item & 1]);
    mbar_arrive_expect_tx(bar, (unsigned)len * 4u);
    bulk_g2s(smem_u32((item & 1) ? stage1 : stage0), gout + (size_t)b * gbs + (size_t)c * Mp + p0, (unsigned)len * 4u, bar);
  };
  if (tid == 0) { issue(0); if (items > 1) issue(1); }
  float acc[NPT][CSR_CPB];
#pragma unroll
  for (int i = 0; i < NPT; i++)
#pragma unroll
    for (int c = 0; c < CSR_CPB; c++) acc[i][c] = 0.f;

  int item = 0;
  for (int k = 0; k < nchunks; k++) {
    const int p0 = k * CSR_CHUNK, len = min(CSR_CHUNK, Mp - p0);
    // stage this chunk's inverse (shared by all channels)
    const int* gb = bnd_all + ((size_t)b * nchunks + k) * (N + 1);
    const unsigned short* gl = list_all + (size_t)b * Mp + p0;
    for (int n = tid; n <= N; n += CSR_THREADS) bnd[n] = __ldg(gb + n);
    for (int i = tid; i < len / 2; i += CSR_THREADS)
      reinterpret_cast<unsigned*>(list)[i] = __ldg(reinterpret_cast<const unsigned*>(gl) + i);
    if ((len & 1) && tid == 0) list[len - 1] = __ldg(gl + len - 1);
    if (tid == 0) nh = 0;
    __syncthreads();
    int lo[NPT], cntv[NPT];
    int pp[NPT][PF];
#pragma unroll
    for (int i = 0; i < NPT; i++) {
      const int n = tid + i * CSR_THREADS;
      lo[i] = 0; cntv[i] = 0;
      if (n < N) { lo[i] = bnd[n]; cntv[i] = bnd[n + 1] - lo[i]; }
      if (cntv[i] > CSR_HEAVY) {
        // hub: its entries are summed by a whole warp per channel (below); this thread only keeps the slot
        const int slot = atomicAdd(&nh, 1);
        h_lo[slot] = lo[i]; h_cnt[slot] = cntv[i];
        lo[i] = slot; cntv[i] = -1;
      }
#pragma unroll
      for (int u = 0; u < PF; u++) pp[i][u] = (u < cntv[i]) ? (int)list[lo[i] + u] : -1;
    }
    __syncthreads();  // heavy slots visible
#pragma unroll
    for (int c = 0; c < CSR_CPB; c++) {
      if (c < nch) {
        const float* st = (item & 1) ? stage1 : stage0;
        while (!mbar_try_wait(smem_u32(&bars[item & 1]), (unsigned)(item >> 1) & 1u)) {}
#pragma unroll
        for (int i = 0; i < NPT; i++) {
          float a = acc[i][c];
#pragma unroll
          for (int u = 0; u < PF; u++)
            if (pp[i][u] >= 0) a += st[pp[i][u]];
          for (int q = PF; q < cntv[i]; q++) a += st[list[lo[i] + q]];  // long lists: remainder from shared memory
          acc[i][c] = a;
        }
        if (nh > 0) csr_heavy_pass(st, list, h_lo, h_cnt, h_sum[item & 1], nh);
        __syncthreads();  // every thread is done with this stage before it is refilled
#pragma unroll
        for (int i = 0; i < NPT; i++)
          if (cntv[i] < 0) acc[i][c] += h_sum[item & 1][lo[i]];
        if (tid == 0 && item + 2 < items) issue(item + 2);
        item++;
      }
    }
    // the barrier above also protects `list` / `bnd` before the next chunk overwrites them
  }
#pragma unroll
  for (int c = 0; c < CSR_CPB; c++) {
    if (c < nch) {
      float* dst = gfeat + ((size_t)b * C + c0 + c) * N;
#pragma unroll
      for (int i = 0; i < NPT; i++) {
        const int n = tid + i * CSR_THREADS;
        if (n < N) dst[n] = acc[i][c];
      }
    }
  }
}

static const size_t GG_SMEM_MAX = 200 * 1024;

static int gather_fwd_impl(const float* feat, const int* idx, float* out, int B, int C, int N,
                           int Mp, int dev, cudaStream_t stream, const char* who) {
  PS_REQUIRE(B >= 0 && C >= 0 && N > 0 && Mp >= 0, "%s: bad sizes B=%d C=%d N=%d M=%d", who, B, C, N, Mp);
  if (B == 0 || C == 0 || Mp == 0) return PS_OK;
  PS_REQUIRE(feat && idx && out, "%s: null pointer", who);
  PS_REQUIRE(B <= 65535, "%s: B=%d exceeds the grid z limit", who, B);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "%s: cannot select device %d", who, dev);
  const int nsm = sm_count(dev);
  const int vec_ok = ((Mp & 3) == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  // Stage through shared memory when each feature element is reused (output >= 2x the input)
  // and 8 rows fit; otherwise gather straight from L1/L2.
  int CT = 8;
  int per_sm = 16;  // measured on C3 (B200): 4 -> 73 %, 8 -> 81 %, 16 -> 83 % of HBM peak (CT=8)
  if (const char* e = getenv("PS_GATHER_CT")) { const int v = atoi(e); if (v == 4 || v == 8) CT = v; }
  if (const char* e = getenv("PS_GATHER_PER_SM")) { const int v = atoi(e); if (v > 0) per_sm = v; }
  const size_t row_bytes = (size_t)N * 4;
  const bool staged = (long long)Mp >= 2ll * N && row_bytes * CT + 128 <= GG_SMEM_MAX / 2 && C >= 2;
  if (staged) {
    const size_t smem = 128 + row_bytes * CT;
    const int ctiles = ceil_div(C, CT);
    // slices so that the grid covers ~per_sm CTAs per SM; each slice a multiple of 1024 positions
    int nslice = ceil_div((long long)nsm * per_sm, (long long)B * ctiles);
    // a slice re-stages the CT rows (N floats each): keep it >= 2N positions so staging stays a
    // fraction of the output traffic, and >= 2048 positions
    const int min_len = 2 * N > 2048 ? 2 * N : 2048;
    const int max_slice = Mp / min_len > 0 ? Mp / min_len : 1;
    if (nslice > max_slice) nslice = max_slice;
    if (nslice < 1) nslice = 1;
    int slice_len = ceil_div(Mp, nslice);
    slice_len = (slice_len + 1023) / 1024 * 1024;
    nslice = ceil_div(Mp, slice_len);
    const int use_bulk = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0) && row_bytes <= 65536 * 4;
    if (CT == 8) {
      auto kern = gather_staged_kernel<8>;
      PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<dim3(nslice, ctiles, B), GG_THREADS, smem, stream>>>(feat, idx, out, C, N, Mp, slice_len, use_bulk, vec_ok);
    } else {
      auto kern = gather_staged_kernel<4>;
      PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<dim3(nslice, ctiles, B), GG_THREADS, smem, stream>>>(feat, idx, out, C, N, Mp, slice_len, use_bulk, vec_ok);
    }
    PS_LAUNCH_CHECK();
  } else {
    constexpr int CTD = 4;
    gather_direct_kernel<CTD><<<dim3(ceil_div(Mp, GG_THREADS * 4), ceil_div(C, CTD), B), GG_THREADS, 0, stream>>>(feat, idx, out, C, N, Mp, vec_ok);
    PS_LAUNCH_CHECK();
  }
  return PS_OK;
}

// gbs = batch stride of gout in elements (C*Mp for a dense (B,C,Mp) tensor; 2*C*Mp when gout is the first
// half of an EdgeConv feature gradient)
static int gather_bwd_impl(const float* gout, const int* idx, float* gfeat, int B, int C, int N,
                           int Mp, int dev, cudaStream_t stream, const char* who, size_t gbs = 0) {
  if (gbs == 0) gbs = (size_t)C * Mp;
  PS_REQUIRE(B >= 0 && C >= 0 && N > 0 && Mp >= 0, "%s: bad sizes B=%d C=%d N=%d M=%d", who, B, C, N, Mp);
  if (B == 0 || C == 0) return PS_OK;
  PS_REQUIRE(gfeat && (Mp == 0 || (gout && idx)), "%s: null pointer", who);
  PS_REQUIRE(B <= 65535 && C <= 65535, "%s: B or C exceeds the grid limit", who);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "%s: cannot select device %d", who, dev);
  const size_t row_bytes = (size_t)N * 4;
  if (Mp == 0) {
    PS_CUDA(cudaMemsetAsync(gfeat, 0, (size_t)B * C * row_bytes, stream));
    return PS_OK;
  }
  const int vec_ok = ((Mp & 3) == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) && ((reinterpret_cast<uintptr_t>(gout) & 15) == 0);
  // Dense grouping (each source referenced >= 4 times on average, several channels): inverse-index
  // path.  Needs 16-byte aligned rows for the bulk copies and N small enough for the build kernel.
  // (measured at the models' shapes, tools/sweep_scatter.py: below ~4096 positions per cloud the fixed cost of
  // the build + the 1024-thread persistent CTAs loses to the shared-memory accumulators)
  bool use_csr = (long long)Mp >= 4ll * N && Mp >= 4096 && C >= 4 && N <= 8 * CSR_THREADS && (Mp & 3) == 0 &&
                 (reinterpret_cast<uintptr_t>(gout) & 15) == 0;
  if (const char* e = getenv("PS_SCATTER_CSR")) use_csr = use_csr && atoi(e) != 0;
  if (use_csr) {
    const int nchunks = ceil_div(Mp, CSR_CHUNK);
    const size_t bnd_bytes = ((size_t)B * nchunks * (N + 1) * sizeof(int) + 15) / 16 * 16;
    const size_t list_bytes = (size_t)B * Mp * sizeof(unsigned short);
    ScratchGuard scratch_mem;
    if (int rc = scratch_mem.alloc(bnd_bytes + list_bytes, dev, stream)) return rc;
    unsigned char* scratch = static_cast<unsigned char*>(scratch_mem.ptr);
    int* bnd = reinterpret_cast<int*>(scratch);
    unsigned short* list = reinterpret_cast<unsigned short*>(scratch + bnd_bytes);
    const size_t smem_b = (size_t)(2 * N + 1) * sizeof(int) + (size_t)CSR_CHUNK * sizeof(unsigned short);
    PS_CUDA(cudaFuncSetAttribute(csr_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    csr_build_kernel<<<dim3(nchunks, B), CSR_BUILD_THREADS, smem_b, stream>>>(idx, bnd, list, N, Mp, nchunks);
    PS_LAUNCH_CHECK();
    const size_t smem = 128 + (size_t)2 * CSR_CHUNK * sizeof(float) + (size_t)CSR_CHUNK * sizeof(unsigned short) +
                        (size_t)(N + 1) * sizeof(int);
    const int npt = ceil_div(N, CSR_THREADS);
    // register budget at 1024 threads is 64: accumulators NPT*CPB + cached entries NPT*PF
#define PS_CSR(NPTV, PFV, CPBV)                                                                    \
  {                                                                                                \
    auto kern = scatter_csr_kernel<NPTV, PFV, CPBV>;                                               \
    PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    kern<<<dim3(ceil_div(C, CPBV), B), CSR_THREADS, smem, stream>>>(gout, bnd, list, gfeat, C, N, Mp, nchunks, gbs); \
  }
    if (npt <= 1) PS_CSR(1, 12, 8) else if (npt <= 2) PS_CSR(2, 12, 8) else if (npt <= 4) PS_CSR(4, 8, 4) else PS_CSR(8, 4, 2)
#undef PS_CSR
    PS_LAUNCH_CHECK();
    return scratch_mem.release();
  }
  int ct = 0;
  if (row_bytes * 8 <= GG_SMEM_MAX / 2) ct = 8;
  else if (row_bytes * 4 <= GG_SMEM_MAX) ct = 4;
  else if (row_bytes * 2 <= GG_SMEM_MAX) ct = 2;
  else if (row_bytes <= GG_SMEM_MAX) ct = 1;
  if (ct) {
    const size_t smem = row_bytes * ct;
    const dim3 grid(1, ceil_div(C, ct), B);
#define PS_SCATTER(CTV)                                                                          \
  {                                                                                              \
    auto kern = scatter_staged_kernel<CTV>;                                                      \
    PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, GG_THREADS, smem, stream>>>(gout, idx, gfeat, C, N, Mp, vec_ok, gbs);                   \
  }
    if (ct == 8) PS_SCATTER(8) else if (ct == 4) PS_SCATTER(4) else if (ct == 2) PS_SCATTER(2) else PS_SCATTER(1)
#undef PS_SCATTER
    PS_LAUNCH_CHECK();
  } else {
    PS_CUDA(cudaMemsetAsync(gfeat, 0, (size_t)B * C * row_bytes, stream));
    scatter_direct_kernel<<<dim3(ceil_div(Mp, GG_THREADS), C, B), GG_THREADS, 0, stream>>>(gout, idx, gfeat, C, N, Mp, gbs);
    PS_LAUNCH_CHECK();
  }
  return PS_OK;
}

// ---- EdgeConv front: cat(central - neighbour, central) -------------------------------------------
// Replaces group_local + the tensor algebra at the top of EdgeConv.forward (models/model_utils.py:812-826,
// 869-877: query_knn_point -> index_points -> permute -> .contiguous() -> unsqueeze/repeat -> subtract -> cat,
// six full passes over (B,C,N,K)-sized tensors) by one pass:
//   out[b, c,     n, k] = x[b,c,n] - x[b,c,idx[b,n,k]]
//   out[b, C + c, n, k] = x[b,c,n]                              x (B,C,N), idx (B,N,K), out (B,2C,N,K)
// Same CTA layout as gather_staged_kernel (rows staged by cp.async.bulk, 128-bit index loads and streaming
// stores).  Algorithmic HBM bytes: 4*(B*N*K + B*C*N + 2*B*C*N*K).
template <int CT>
__global__ void __launch_bounds__(GG_THREADS) edge_staged_kernel(
    const float* __restrict__ feat, const int* __restrict__ idx, float* __restrict__ out, int C,
    int N, int K, int Mp, int slice_len, int use_bulk, int vec_ok) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  u64* bar = reinterpret_cast<u64*>(smem_raw);
  float* rows = reinterpret_cast<float*>(smem_raw + 128);
  const int b = blockIdx.z, c0 = blockIdx.y * CT, tid = threadIdx.x;
  const int nrows = min(CT, C - c0);
  const float* src = feat + ((size_t)b * C + c0) * N;
  if (use_bulk) {
    if (tid == 0) {
      mbar_init(smem_u32(bar), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(smem_u32(bar), (unsigned)(nrows * N * 4));
      for (int r = 0; r < nrows; r++)
        bulk_g2s(smem_u32(rows + (size_t)r * N), src + (size_t)r * N, (unsigned)(N * 4), smem_u32(bar));
    }
  } else {
    for (int i = tid; i < nrows * N; i += GG_THREADS) rows[i] = __ldg(src + i);
  }
  const int p_begin = blockIdx.x * slice_len;
  const int p_end = min(Mp, p_begin + slice_len);
  const int* ib = idx + (size_t)b * Mp;
  const bool vec = vec_ok != 0;  // K % 4 == 0 (so 4 consecutive positions share their centre) and aligned bases
  if (use_bulk) {
    while (!mbar_try_wait(smem_u32(bar), 0)) {}
  } else {
    __syncthreads();
  }
  for (int p = p_begin + tid * 4; p < p_end; p += GG_THREADS * 4) {
    int ii[4], nn[4];
    const bool full = vec && (p + 4 <= p_end);
    if (full) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(ib + p));
      ii[0] = v.x; ii[1] = v.y; ii[2] = v.z; ii[3] = v.w;
      nn[0] = nn[1] = nn[2] = nn[3] = p / K;
    } else {
#pragma unroll
      for (int e = 0; e < 4; e++) {
        ii[e] = p + e < p_end ? __ldg(ib + p + e) : 0;
        nn[e] = p + e < p_end ? (p + e) / K : 0;
      }
    }
    float* oe = out + ((size_t)b * 2 * C + c0) * Mp + p;       // edge half
    float* oc = out + ((size_t)b * 2 * C + C + c0) * Mp + p;   // central half
#pragma unroll
    for (int r = 0; r < CT; r++) {
      if (r < nrows) {
        const float* row = rows + (size_t)r * N;
        const float4 cen = make_float4(row[nn[0]], row[nn[1]], row[nn[2]], row[nn[3]]);
        const float4 edg = make_float4(__fsub_rn(cen.x, row[ii[0]]), __fsub_rn(cen.y, row[ii[1]]),
                                       __fsub_rn(cen.z, row[ii[2]]), __fsub_rn(cen.w, row[ii[3]]));
        float* e_ = oe + (size_t)r * Mp;
        float* c_ = oc + (size_t)r * Mp;
        if (full) {
          __stcs(reinterpret_cast<float4*>(e_), edg);
          __stcs(reinterpret_cast<float4*>(c_), cen);
        } else {
          const float ev[4] = {edg.x, edg.y, edg.z, edg.w}, cv[4] = {cen.x, cen.y, cen.z, cen.w};
#pragma unroll
          for (int e = 0; e < 4; e++)
            if (p + e < p_end) { e_[e] = ev[e]; c_[e] = cv[e]; }
        }
      }
    }
  }
}

// rows too long for shared memory: straight from L1/L2
__global__ void __launch_bounds__(GG_THREADS) edge_direct_kernel(const float* __restrict__ feat, const int* __restrict__ idx,
                                                                 float* __restrict__ out, int C, int N, int K, int Mp) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int p = blockIdx.x * GG_THREADS + threadIdx.x;
  if (p >= Mp) return;
  const float* row = feat + ((size_t)b * C + c) * N;
  const float cen = __ldg(row + p / K);
  const float nb = __ldg(row + __ldg(idx + (size_t)b * Mp + p));
  out[((size_t)b * 2 * C + c) * Mp + p] = __fsub_rn(cen, nb);
  out[((size_t)b * 2 * C + C + c) * Mp + p] = cen;
}

// backward finish: gx[b,c,n] = sum_k (gE[b,c,n,k] + gC[b,c,n,k]) - gx[b,c,n]   (gx holds the scattered gE on entry)
// LPR = K/4 lanes share one (b,c,n) row: every lane loads one float4 of each half (fully coalesced), the
// row sum is finished with shuffles.  LPR in {1,2,4,8}; other K take the scalar kernel below.
template <int LPR>
__global__ void __launch_bounds__(GG_THREADS) edge_bwd_finish_vec_kernel(const float* __restrict__ gout, float* __restrict__ gx,
                                                                         int C, int N, long long rows_per_b, long long total_rows) {
  const long long t = (long long)blockIdx.x * GG_THREADS + threadIdx.x;  // one thread per float4 of a row
  const long long row = t / LPR;                                         // (b, c, n) flattened
  const bool valid = row < total_rows;
  float s = 0.f;
  if (valid) {
    const long long b = row / rows_per_b, r = row % rows_per_b;         // rows_per_b = C*N
    const size_t base = ((size_t)b * 2 * rows_per_b + r) * (LPR * 4) + (size_t)(t % LPR) * 4;
    const float4 e = __ldcs(reinterpret_cast<const float4*>(gout + base));
    const float4 c = __ldcs(reinterpret_cast<const float4*>(gout + base + (size_t)rows_per_b * (LPR * 4)));
    s = ((e.x + c.x) + (e.y + c.y)) + ((e.z + c.z) + (e.w + c.w));
  }
#pragma unroll
  for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (valid && (t % LPR) == 0) gx[row] = s - gx[row];
}
__global__ void __launch_bounds__(GG_THREADS) edge_bwd_finish_kernel(const float* __restrict__ gout, float* __restrict__ gx,
                                                                     int C, int N, int K, long long total) {
  const long long t = (long long)blockIdx.x * GG_THREADS + threadIdx.x;
  if (t >= total) return;
  const int n = (int)(t % N);
  const int c = (int)((t / N) % C);
  const int b = (int)(t / ((long long)N * C));
  const float* ge = gout + (((size_t)b * 2 * C + c) * N + n) * K;
  const float* gc = gout + (((size_t)b * 2 * C + C + c) * N + n) * K;
  float s = 0.f;
  for (int k = 0; k < K; k++) s += __ldg(ge + k) + __ldg(gc + k);
  gx[t] = s - gx[t];
}

// ---- index_points: row gather on point-major tensors (models/model_utils.py:828-845) -----------------
//   out[b,m,:] = points[b, idx[b,m], :]      points (B,N,C), idx (B,M) (any trailing shape flattened), out (B,M,C)
// grid = (row blocks, B); a CTA copies IP_ROWS consecutive output rows; one thread per VEC elements, 4 in flight
constexpr int IP_ROWS = 64;
template <int VEC>
__global__ void __launch_bounds__(GG_THREADS) index_points_kernel(const float* __restrict__ pts, const int* __restrict__ idx,
                                                                  float* __restrict__ out, int N, int M, int C) {
  __shared__ int sidx[IP_ROWS];
  const int b = blockIdx.y, m0 = blockIdx.x * IP_ROWS, tid = threadIdx.x;
  const int rows = min(IP_ROWS, M - m0);
  if (tid < rows) sidx[tid] = __ldg(idx + (size_t)b * M + m0 + tid);
  __syncthreads();
  const int cv = C / VEC;              // vector elements per row
  const int total = rows * cv;
  const float* src = pts + (size_t)b * N * C;
  float* dst = out + ((size_t)b * M + m0) * C;
  for (int e0 = tid; e0 < total; e0 += GG_THREADS * 4) {
    if (VEC == 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = e0 + u * GG_THREADS;
        if (e < total) { const int r = e / cv, c = (e - r * cv) * 4; v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)sidx[r] * C + c)); }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = e0 + u * GG_THREADS;
        if (e < total) __stcs(reinterpret_cast<float4*>(dst + (size_t)e * 4), v[u]);
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = e0 + u * GG_THREADS;
        if (e < total) { const int r = e / cv, c = e - r * cv; dst[e] = __ldg(src + (size_t)sidx[r] * C + c); }
      }
    }
  }
}
template <int VEC>
__global__ void __launch_bounds__(GG_THREADS) index_points_grad_kernel(const float* __restrict__ gout, const int* __restrict__ idx,
                                                                       float* __restrict__ gpts, int N, int M, int C, long long total) {
  const long long t = (long long)blockIdx.x * GG_THREADS + threadIdx.x;
  if (t >= total) return;
  const int cv = C / VEC;
  const int c = (int)(t % cv) * VEC;
  const long long row = t / cv;
  const int b = (int)(row / M);
  const int i = __ldg(idx + row);
  float* d = gpts + ((size_t)b * N + i) * C + c;
  const float* g = gout + (size_t)row * C + c;
  if (VEC == 4) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(g));
    atomicAdd(reinterpret_cast<float4*>(d), v);  // red.global.add.v4.f32 (sm_90+)
  } else {
    atomicAdd(d, __ldg(g));
  }
}
}  // namespace ps

using namespace ps;

extern "C" int ps_gather_fwd(const float* features, const int* idx, float* out, int B, int C,
                             int N, int M, int dev, void* stream) {
  return gather_fwd_impl(features, idx, out, B, C, N, M, dev, (cudaStream_t)stream, "ps_gather_fwd");
}
extern "C" int ps_gather_bwd(const float* grad_out, const int* idx, float* grad_features, int B,
                             int C, int N, int M, int dev, void* stream) {
  return gather_bwd_impl(grad_out, idx, grad_features, B, C, N, M, dev, (cudaStream_t)stream, "ps_gather_bwd");
}
extern "C" int ps_group_fwd(const float* features, const int* idx, float* out, int B, int C, int N,
                            int S, int K, int dev, void* stream) {
  PS_REQUIRE(S >= 0 && K >= 0 && (long long)S * K < (1ll << 31), "ps_group_fwd: bad S=%d K=%d", S, K);
  return gather_fwd_impl(features, idx, out, B, C, N, S * K, dev, (cudaStream_t)stream, "ps_group_fwd");
}
extern "C" int ps_group_bwd(const float* grad_out, const int* idx, float* grad_features, int B,
                            int C, int N, int S, int K, int dev, void* stream) {
  PS_REQUIRE(S >= 0 && K >= 0 && (long long)S * K < (1ll << 31), "ps_group_bwd: bad S=%d K=%d", S, K);
  return gather_bwd_impl(grad_out, idx, grad_features, B, C, N, S * K, dev, (cudaStream_t)stream, "ps_group_bwd");
}

extern "C" int ps_edge_features_fwd(const float* x, const int* idx, float* out, int B, int C, int N, int K,
                                    int dev, void* stream_) {
  PS_REQUIRE(B >= 0 && C >= 0 && N > 0 && K >= 0 && (long long)N * K < (1ll << 31), "ps_edge_features_fwd: bad sizes B=%d C=%d N=%d K=%d", B, C, N, K);
  if (B == 0 || C == 0 || K == 0) return PS_OK;
  PS_REQUIRE(x && idx && out, "ps_edge_features_fwd: null pointer");
  PS_REQUIRE(B <= 65535 && C <= 65535, "ps_edge_features_fwd: B or C exceeds the grid limit");
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_edge_features_fwd: cannot select device %d", dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nsm = sm_count(dev);
  const int Mp = N * K;
  const size_t row_bytes = (size_t)N * 4;
  constexpr int CT = 8;
  if (row_bytes * CT + 128 <= GG_SMEM_MAX / 2) {
    const int vec_ok = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int ctiles = ceil_div(C, CT);
    int nslice = ceil_div((long long)nsm * 16, (long long)B * ctiles);
    const int min_len = 2 * N > 2048 ? 2 * N : 2048;
    const int max_slice = Mp / min_len > 0 ? Mp / min_len : 1;
    if (nslice > max_slice) nslice = max_slice;
    if (nslice < 1) nslice = 1;
    int slice_len = ceil_div(Mp, nslice);
    slice_len = (slice_len + 1023) / 1024 * 1024;
    nslice = ceil_div(Mp, slice_len);
    const int use_bulk = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const size_t smem = 128 + row_bytes * CT;
    auto kern = edge_staged_kernel<CT>;
    PS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(nslice, ctiles, B), GG_THREADS, smem, stream>>>(x, idx, out, C, N, K, Mp, slice_len, use_bulk, vec_ok);
  } else {
    edge_direct_kernel<<<dim3(ceil_div(Mp, GG_THREADS), C, B), GG_THREADS, 0, stream>>>(x, idx, out, C, N, K, Mp);
  }
  PS_LAUNCH_CHECK();
  return PS_OK;
}

extern "C" int ps_edge_features_bwd(const float* grad_out, const int* idx, float* grad_x, int B, int C, int N,
                                    int K, int dev, void* stream_) {
  PS_REQUIRE(B >= 0 && C >= 0 && N > 0 && K >= 0 && (long long)N * K < (1ll << 31), "ps_edge_features_bwd: bad sizes B=%d C=%d N=%d K=%d", B, C, N, K);
  if (B == 0 || C == 0) return PS_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  // 1. scatter of the edge half through the grouping backward (grad_out batch stride = 2*C*N*K)
  if (int rc = gather_bwd_impl(grad_out, idx, grad_x, B, C, N, N * K, dev, stream, "ps_edge_features_bwd", (size_t)2 * C * N * K)) return rc;
  if (K == 0) return PS_OK;
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_edge_features_bwd: cannot select device %d", dev);
  // 2. centre terms minus the scatter
  const long long total = (long long)B * C * N;
  const bool vec = (K == 4 || K == 8 || K == 16 || K == 32) && (reinterpret_cast<uintptr_t>(grad_out) & 15) == 0;
  if (vec) {
    const int lpr = K / 4;
    const long long threads = total * lpr;
    PS_REQUIRE(threads / GG_THREADS < (1ll << 31) - 1, "ps_edge_features_bwd: tensor too large");
    const int grid = ceil_div(threads, GG_THREADS);
    if (lpr == 1) edge_bwd_finish_vec_kernel<1><<<grid, GG_THREADS, 0, stream>>>(grad_out, grad_x, C, N, (long long)C * N, total);
    else if (lpr == 2) edge_bwd_finish_vec_kernel<2><<<grid, GG_THREADS, 0, stream>>>(grad_out, grad_x, C, N, (long long)C * N, total);
    else if (lpr == 4) edge_bwd_finish_vec_kernel<4><<<grid, GG_THREADS, 0, stream>>>(grad_out, grad_x, C, N, (long long)C * N, total);
    else edge_bwd_finish_vec_kernel<8><<<grid, GG_THREADS, 0, stream>>>(grad_out, grad_x, C, N, (long long)C * N, total);
  } else {
    edge_bwd_finish_kernel<<<ceil_div(total, GG_THREADS), GG_THREADS, 0, stream>>>(grad_out, grad_x, C, N, K, total);
  }
  PS_LAUNCH_CHECK();
  return PS_OK;
}

extern "C" int ps_index_points_fwd(const float* points, const int* idx, float* out, int B, int N, int M, int C,
                                   int dev, void* stream_) {
  PS_REQUIRE(B >= 0 && N > 0 && M >= 0 && C >= 0, "ps_index_points_fwd: bad sizes B=%d N=%d M=%d C=%d", B, N, M, C);
  if (B == 0 || M == 0 || C == 0) return PS_OK;
  PS_REQUIRE(points && idx && out, "ps_index_points_fwd: null pointer");
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_index_points_fwd: cannot select device %d", dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool v4 = ((C & 3) == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  PS_REQUIRE(B <= 65535 && (long long)IP_ROWS * C < (1ll << 30), "ps_index_points_fwd: B or C too large");
  const dim3 grid(ceil_div(M, IP_ROWS), B);
  if (v4) index_points_kernel<4><<<grid, GG_THREADS, 0, stream>>>(points, idx, out, N, M, C);
  else index_points_kernel<1><<<grid, GG_THREADS, 0, stream>>>(points, idx, out, N, M, C);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

extern "C" int ps_index_points_bwd(const float* grad_out, const int* idx, float* grad_points, int B, int N, int M,
                                   int C, int dev, void* stream_) {
  PS_REQUIRE(B >= 0 && N > 0 && M >= 0 && C >= 0, "ps_index_points_bwd: bad sizes B=%d N=%d M=%d C=%d", B, N, M, C);
  if (B == 0 || C == 0) return PS_OK;
  PS_REQUIRE(grad_points && (M == 0 || (grad_out && idx)), "ps_index_points_bwd: null pointer");
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_index_points_bwd: cannot select device %d", dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  PS_CUDA(cudaMemsetAsync(grad_points, 0, (size_t)B * N * C * sizeof(float), stream));
  if (M == 0) return PS_OK;
  const bool v4 = ((C & 3) == 0) && ((reinterpret_cast<uintptr_t>(grad_points) & 15) == 0) && ((reinterpret_cast<uintptr_t>(grad_out) & 15) == 0);
  const long long total = (long long)B * M * (v4 ? C / 4 : C);
  PS_REQUIRE(total / GG_THREADS < (1ll << 31) - 1, "ps_index_points_bwd: tensor too large");
  if (v4) index_points_grad_kernel<4><<<ceil_div(total, GG_THREADS), GG_THREADS, 0, stream>>>(grad_out, idx, grad_points, N, M, C, total);
  else index_points_grad_kernel<1><<<ceil_div(total, GG_THREADS), GG_THREADS, 0, stream>>>(grad_out, idx, grad_points, N, M, C, total);
  PS_LAUNCH_CHECK();
  return PS_OK;
}
