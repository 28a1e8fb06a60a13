// Peer-memory exchange of the loss/metric partial sums (the path's only collective).
//
// The reduction the reference's callers apply to Chamfer's outputs (utils/loss_utils.py:10-31,50-57) needs
// a few doubles per rank per step.  Instead of a library all-reduce (a host-issued c10d/NCCL call per step:
// ~1 ms of host time at 8 ranks against a 0.35 ms device step, SCALE_r01), every rank owns a MAILBOX in its
// own HBM that all peers can store into over NVLink (CUDA IPC / peer access).  The kernel that produces the
// sums PUBLISHES them — one warp, lane r stores the payload and then a sequence number into rank r's mailbox
// (st.release.sys) — and a one-warp WAIT step (acquire loads on the local mailbox only, nothing polls across
// NVLink) adds the world's payloads in rank order, so every rank gets bit-identical sums.  No host call, no
// stream other than the caller's: the whole step is replayable as one CUDA graph.
//
// Mailbox = CHANNELS x DEPTH x world slots.  Step s of a channel uses row s % DEPTH.  A rank can run at most one
// publish ahead of the slowest rank's wait (its own wait of step s needs everybody's publish of s), so DEPTH = 4
// leaves slack.  The per-rank sequence counters live in device memory and are advanced by the kernels themselves.
// CHANNELS: independent instances of the protocol (own rows, own counters) inside one communicator.  Exchanges of
// one channel must be issued in the same order on every rank; exchanges of different channels may interleave
// freely.  Channel 0 serves everything stream-ordered; channels 1.. belong to the lanes of the asynchronous
// host-buffer steps (host_pipeline.cu), which run concurrently on every rank.
#pragma once
#include "common.cuh"

namespace ps {

constexpr int COMM_MAX_WORLD = 16;
constexpr int COMM_MAX_N = 30;  // doubles per message
constexpr int COMM_DEPTH = 4;
constexpr int COMM_CHANNELS = 5;
constexpr unsigned long long COMM_TIMEOUT_NS = 4000000000ull;  // a lost peer raises an error instead of hanging the GPU

struct __align__(16) CommSlot {
  double v[COMM_MAX_N];
  u64 n;
  u64 seq;
};
static_assert(sizeof(CommSlot) == 256, "slot is 256 bytes");

struct CommDev {  // passed by value to kernels
  CommSlot* peer[COMM_MAX_WORLD];  // peer[r]: rank r's mailbox (peer[rank]: our own)
  u64* pub_seq;                    // messages published by this rank so far
  u64* wait_seq;                   // messages consumed by this rank so far
  int* err;                        // set to 1 when a wait timed out
  int rank, world;
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys_u64(u64* p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys_u64(const u64* p) { u64 v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) { asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) { double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ u64 globaltimer_ns() { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// Called by ONE full warp.  vals: n doubles readable by every lane (shared or global memory).
__device__ __forceinline__ void comm_publish(const CommDev& c, const double* vals, int n) {
  const int lane = threadIdx.x & 31;
  u64 s = 0;
  if (lane == 0) { s = *c.pub_seq + 1; *c.pub_seq = s; }
  s = __shfl_sync(0xffffffffu, s, 0);
  if (lane < c.world) {
    CommSlot* d = c.peer[lane] + (size_t)(s % COMM_DEPTH) * c.world + c.rank;
    for (int i = 0; i < n; i++) st_relaxed_sys_f64(&d->v[i], vals[i]);
    d->n = (u64)n;
    __threadfence_system();
    st_release_sys_u64(&d->seq, s);
  }
}

// Called by ONE full warp.  out: n doubles = sum over ranks (rank order) of the step's payloads; NaN after a timeout.
__device__ __forceinline__ void comm_wait_reduce(const CommDev& c, double* out, int n) {
  const int lane = threadIdx.x & 31;
  u64 w = 0;
  if (lane == 0) { w = *c.wait_seq + 1; *c.wait_seq = w; }
  w = __shfl_sync(0xffffffffu, w, 0);
  const CommSlot* row = c.peer[c.rank] + (size_t)(w % COMM_DEPTH) * c.world;
  bool ok = true;
  if (lane < c.world) {
    const u64 t0 = globaltimer_ns();
    while (ld_acquire_sys_u64(&row[lane].seq) != w) {
      if (globaltimer_ns() - t0 > COMM_TIMEOUT_NS) { ok = false; break; }
      __nanosleep(64);
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  __threadfence_system();
  if (lane < n) {
    double acc = 0.0;
    for (int r = 0; r < c.world; r++) acc += ld_relaxed_sys_f64(&row[r].v[lane]);
    out[lane] = ok ? acc : __longlong_as_double(0x7ff8000000000000ll);
  }
  if (!ok && lane == 0) *c.err = 1;
}
#endif

}  // namespace ps

namespace ps {
// device view of channel `ch` of a communicator (channel 0: comm->d itself)
CommDev comm_channel(const ::ps_comm* comm, int ch);
int comm_allreduce_launch(const ::ps_comm* comm, int ch, const double* in, double* out, int n, cudaStream_t stream);
// stand-alone halves of the exchange for producers that cannot publish from their own last block
int comm_publish_launch(const ::ps_comm* comm, const double* in, int n, cudaStream_t stream);
int comm_wait_launch(const ::ps_comm* comm, double* out, int n, cudaStream_t stream);
// chamfer.cu
int chamfer_fwd_impl(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int* idx1, int* idx2,
                     double* sums6, const ::ps_comm* comm, int B, int N, int M, int dev, void* stream, const char* who);
}  // namespace ps

// Host handle behind the opaque ps_comm of include/pointsea_b200.h
struct ps_comm {
  int rank = 0, world = 1, dev = 0;
  ps::CommSlot* mailbox = nullptr;     // our own (cudaMalloc)
  unsigned long long* counters = nullptr;  // [2c] pub_seq, [2c+1] wait_seq of channel c, [2*COMM_CHANNELS] err (as int)
  void* opened[ps::COMM_MAX_WORLD] = {nullptr};  // cudaIpcOpenMemHandle results (to close)
  ps::CommDev d;
  bool connected = false;
};
