// Host-buffer entry point for the Chamfer step: the call a host-side caller (data loader, metric
// loop, the reference's CPU-allocating wrapper dist_chamfer_3D.py:33-42) makes when its clouds
// live in host memory.  The batch is cut into chunks of clouds and three streams run
//     H2D(chunk c+1)  ||  forward+backward kernels(chunk c)  ||  D2H(chunk c-1)
// so PCIe in, compute and PCIe out overlap; only the first upload and the last download are
// exposed.  Every kernel is per-cloud, so chunking changes no result bit.
//
// The call only ENQUEUES (like every other entry point): the host output buffers are complete when
// `stream` reaches the point of the call, i.e. after the caller synchronises `stream` or an event
// recorded on it.  Host buffers should be pinned (cudaHostAlloc / torch pin_memory); pageable memory
// works but the runtime then stages every copy synchronously and nothing overlaps.
//
// Per-device context (created on first use, guarded by a mutex): three non-blocking streams,
// SLOTS device staging slots (grown on demand, never shrunk) and their events.
//
// The pipeline is ~25 runtime calls per chunk (10 copies, ~8 launches, events); issued one by one
// the HOST becomes the bottleneck (measured: 0.64 ms per C1 step against 0.78 ms without any
// overlap).  So the whole multi-stream pipeline of one call is captured ONCE into a CUDA graph,
// keyed by the buffer addresses and sizes, and replayed with a single cudaGraphLaunch on the
// caller's stream while the key repeats (steady-state training / evaluation loops with persistent
// pinned buffers).  PS_HOST_GRAPH=0 disables the graphs.
//
// LANES: a call occupies the staging slots of its lane from its first upload to its last download, so calls on
// one lane run back to back.  The asynchronous entry points (ps_chamfer_host_submit / ps_chamfer_host_wait)
// alternate between LANES independent lanes, each with its own copy streams, slots and events, which is what a
// loader that prefetches the next batches does: PCIe is full duplex and the two copy engines and the SMs are three
// separate resources.  Two shapes of an asynchronous step were measured on C1 (one call at a time: 0.554 ms):
//   * "lanes": the chunked multi-stream graph of the stream-ordered call, one per lane on the lane's own launch
//     stream — 0.419 / 0.392 / 0.397 ms with 2 / 3 / 4 steps in flight.  Kernels of different steps run side by
//     side, and the three small forward launches of a chunked step cost 0.375 ms of kernels instead of 0.33;
//   * "FIFO" (the default, submit_fifo): a step = uploads on the lane's copy stream -> ALL its kernels as one graph
//     on the stream that serves every lane in submission order (one step at a time, each with the whole GPU) ->
//     downloads on the lane's copy stream, chained by events — 0.428 / 0.342 / **0.335 ms** with 2 / 3 / 4 steps in
//     flight, within 8 % of the 0.309 ms the same kernels take on device-resident buffers.
// PS_HOST_ASYNC=lanes selects the first.  The stream-ordered entry points keep their contract ("complete when
// `stream` reaches the call") and always use lane 0 and the chunked graph.
#include "comm.cuh"
#include "graph_cache.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace ps {

constexpr int SLOTS = 3;
constexpr int MAX_CHUNKS = 64;
constexpr int LANES = 4;
static_assert(LANES + 1 <= COMM_CHANNELS, "every lane needs its own communicator channel");

struct Lane {
  GraphCache graphs;
  cudaStream_t s_cap = nullptr;     // origin stream of the captures
  cudaStream_t s_launch = nullptr;  // launch stream of the asynchronous entry points
  cudaEvent_t ev_last = nullptr;    // end of the most recent call on this lane (any caller stream)
  cudaEvent_t ev_sub = nullptr;     // the submitting stream's position at ps_chamfer_host_submit
  cudaEvent_t ev_cmp = nullptr;     // end of the kernels of the lane's latest submission (FIFO mode)
  unsigned long long calls = 0;
  bool ready = false;
  cudaStream_t s_in = nullptr, s_run = nullptr, s_bwd = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[SLOTS], ev_gin[SLOTS], ev_fwd[SLOTS], ev_run[SLOTS], ev_out[SLOTS], ev_begin = nullptr, ev_end = nullptr;
  void* slot[SLOTS] = {nullptr, nullptr, nullptr};
  size_t slot_bytes = 0;
  double* d_sums = nullptr;  // 12 doubles: loss partial sums of the current call (ps_chamfer_host_step) [+ world-wide sums]
};

struct HostPipe {
  std::mutex mu;
  Lane lane[LANES];
  unsigned long long submits = 0;
  cudaStream_t s_compute = nullptr;  // FIFO mode: the kernels of ALL lanes' submissions, one step after the other
  GraphCache cgraphs;                // their graphs (kernels only; the copies around them are plain async copies)
};

static HostPipe* pipe_for(int dev) {
  static HostPipe pipes[64];
  return (dev >= 0 && dev < 64) ? &pipes[dev] : nullptr;
}

static int pipe_init(Lane& hp, int hp_dev) {
  if (hp.ready) return PS_OK;
  PS_CUDA(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
  PS_CUDA(cudaStreamCreateWithFlags(&hp.s_run, cudaStreamNonBlocking));
  PS_CUDA(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
  PS_CUDA(cudaStreamCreateWithFlags(&hp.s_bwd, cudaStreamNonBlocking));
  for (int i = 0; i < SLOTS; i++) {
    PS_CUDA(cudaEventCreateWithFlags(&hp.ev_in[i], cudaEventDisableTiming));
    PS_CUDA(cudaEventCreateWithFlags(&hp.ev_gin[i], cudaEventDisableTiming));
    PS_CUDA(cudaEventCreateWithFlags(&hp.ev_fwd[i], cudaEventDisableTiming));
    PS_CUDA(cudaEventCreateWithFlags(&hp.ev_run[i], cudaEventDisableTiming));
    PS_CUDA(cudaEventCreateWithFlags(&hp.ev_out[i], cudaEventDisableTiming));
  }
  PS_CUDA(cudaStreamCreateWithFlags(&hp.s_cap, cudaStreamNonBlocking));
  PS_CUDA(cudaStreamCreateWithFlags(&hp.s_launch, cudaStreamNonBlocking));
  PS_CUDA(cudaEventCreateWithFlags(&hp.ev_last, cudaEventDisableTiming));
  PS_CUDA(cudaEventCreateWithFlags(&hp.ev_sub, cudaEventDisableTiming));
  PS_CUDA(cudaEventCreateWithFlags(&hp.ev_cmp, cudaEventDisableTiming));
  PS_CUDA(cudaEventCreateWithFlags(&hp.ev_begin, cudaEventDisableTiming));
  PS_CUDA(cudaEventCreateWithFlags(&hp.ev_end, cudaEventDisableTiming));
  PS_CUDA(cudaMalloc(&hp.d_sums, 12 * sizeof(double)));
  scratch_pool_init(hp_dev);
  hp.ready = true;
  return PS_OK;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Chunk plan.  chunk > 0: equal chunks of that many clouds (ragged last one).  chunk <= 0: the
// library plan — a first chunk of B/5 clouds (its upload is the only exposed one) and the rest in
// two halves: few, large chunks keep the forward grid efficient and the number of copies low
// (PCIe moves the ~1 MB pieces of a chunk at well under its streaming rate).  Measured on C1
// (B=32): 6+13+13 0.545 ms, 3+11+11+7 0.564, 4x8 0.592, one chunk 0.70, no overlap 0.78.
// PS_HOST_PLAN="4,12,12,4" overrides (sizes must add up to B).
static int make_plan(int B, int chunk, int* sizes, int* nchunks, int* largest) {
  int n = 0;
  if (chunk <= 0) {
    if (const char* e = getenv("PS_HOST_PLAN")) {
      int sum = 0;
      const char* q = e;
      while (*q && n < MAX_CHUNKS) {
        const int v = atoi(q);
        if (v <= 0) { n = 0; break; }
        sizes[n++] = v;
        sum += v;
        while (*q && *q != ',') q++;
        if (*q == ',') q++;
      }
      if (sum != B) n = 0;
    }
    if (n == 0) {
      if (B < 5) {
        for (int i = 0; i < B; i++) sizes[n++] = 1;
      } else {
        const int first = B / 5;
        const int rest = B - first;
        sizes[n++] = first;
        sizes[n++] = rest - rest / 2;
        sizes[n++] = rest / 2;
      }
    }
  } else {
    if (chunk > B) chunk = B;
    if ((B + chunk - 1) / chunk > MAX_CHUNKS) chunk = (B + MAX_CHUNKS - 1) / MAX_CHUNKS;
    for (int b0 = 0; b0 < B; b0 += chunk) sizes[n++] = (B - b0 < chunk) ? B - b0 : chunk;
  }
  int mx = 0;
  for (int i = 0; i < n; i++) mx = sizes[i] > mx ? sizes[i] : mx;
  *nchunks = n;
  *largest = mx;
  return PS_OK;
}

struct PipeArgs {
  const float *xyz1, *xyz2, *graddist1, *graddist2;
  float *dist1, *dist2, *gradxyz1, *gradxyz2;
  int *idx1, *idx2;
  int B, N, M, chunk, dev;  // chunk = largest chunk (sizes the staging slots)
  int nchunks;
  int sizes[MAX_CHUNKS];
  bool with_bwd;
  // ps_chamfer_host_step: gradients written straight into the caller's DEVICE buffers (never downloaded), the six
  // loss sums accumulated on the device and downloaded once (48 bytes); dist/idx stay in the staging slots
  float *dev_g1 = nullptr, *dev_g2 = nullptr;
  double* h_sums = nullptr;
  ps_comm* comm = nullptr;  // with h_sums: the sums are exchanged with the peers and h_sums receives the world-wide ones
  int channel = 0;          // the communicator's channel: 0 stream-ordered calls, 1 + lane for submissions
  size_t o_x1, o_x2, o_d1, o_d2, o_i1, o_i2, o_gd1, o_gd2, o_g1, o_g2;
};

// Enqueues the chunked three-stream pipeline behind `origin` and joins it back into `origin`.
// Works both eagerly and under stream capture of `origin` (the side streams join the capture
// through the event waits; every branch ends in s_out, which is joined back at the end).
static int enqueue_pipeline(Lane& hp, const PipeArgs& a, cudaStream_t origin) {
#define COPY(...) PS_CUDA(cudaMemcpyAsync(__VA_ARGS__))
  PS_CUDA(cudaEventRecord(hp.ev_begin, origin));
  PS_CUDA(cudaStreamWaitEvent(hp.s_in, hp.ev_begin, 0));
  int b0 = 0;
  for (int c = 0; c < a.nchunks; b0 += a.sizes[c], c++) {
    const int s = c % SLOTS;
    const int nb = a.sizes[c];
    char* base = static_cast<char*>(hp.slot[s]);
    float* d_x1 = reinterpret_cast<float*>(base + a.o_x1);
    float* d_x2 = reinterpret_cast<float*>(base + a.o_x2);
    float* d_d1 = reinterpret_cast<float*>(base + a.o_d1);
    float* d_d2 = reinterpret_cast<float*>(base + a.o_d2);
    int* d_i1 = reinterpret_cast<int*>(base + a.o_i1);
    int* d_i2 = reinterpret_cast<int*>(base + a.o_i2);
    float* d_gd1 = reinterpret_cast<float*>(base + a.o_gd1);
    float* d_gd2 = reinterpret_cast<float*>(base + a.o_gd2);
    float* d_g1 = reinterpret_cast<float*>(base + a.o_g1);
    float* d_g2 = reinterpret_cast<float*>(base + a.o_g2);
    const size_t c1 = (size_t)nb * a.N, c2 = (size_t)nb * a.M;
    const size_t h1 = (size_t)b0 * a.N, h2 = (size_t)b0 * a.M;

    // upload: the slot's inputs are free once the kernels of its previous use (chunk c - SLOTS) are
    // done; uses by earlier CALLS are ordered by `origin` (every call starts behind ev_last)
    if (c >= SLOTS) PS_CUDA(cudaStreamWaitEvent(hp.s_in, hp.ev_run[s], 0));
    COPY(d_x1, a.xyz1 + h1 * 3, c1 * 12, cudaMemcpyHostToDevice, hp.s_in);
    COPY(d_x2, a.xyz2 + h2 * 3, c2 * 12, cudaMemcpyHostToDevice, hp.s_in);
    PS_CUDA(cudaEventRecord(hp.ev_in[s], hp.s_in));
    if (a.with_bwd) {  // only the backward kernels wait for the upstream gradients
      COPY(d_gd1, a.graddist1 + h1, c1 * 4, cudaMemcpyHostToDevice, hp.s_in);
      COPY(d_gd2, a.graddist2 + h2, c2 * 4, cudaMemcpyHostToDevice, hp.s_in);
      PS_CUDA(cudaEventRecord(hp.ev_gin[s], hp.s_in));
    }

    // kernels: need the upload, and the slot's outputs of its previous use downloaded
    PS_CUDA(cudaStreamWaitEvent(hp.s_run, hp.ev_in[s], 0));
    if (c >= SLOTS) PS_CUDA(cudaStreamWaitEvent(hp.s_run, hp.ev_out[s], 0));
    if (int rc = ps_chamfer_fwd(d_x1, d_x2, d_d1, d_d2, d_i1, d_i2, nb, a.N, a.M, a.dev, hp.s_run)) return rc;
    PS_CUDA(cudaEventRecord(hp.ev_fwd[s], hp.s_run));
    // loss sums + backward on their own stream: a handful of small, latency-bound kernels that fill the gaps of
    // the NEXT chunk's forward instead of delaying it (the forward of chunk c+1 only needs its own upload)
    PS_CUDA(cudaStreamWaitEvent(hp.s_bwd, hp.ev_fwd[s], 0));
    if (a.h_sums) {
      if (c == 0) PS_CUDA(cudaMemsetAsync(hp.d_sums, 0, 6 * sizeof(double), hp.s_bwd));
      if (int rc = chamfer_sums_launch(d_d1, d_d2, hp.d_sums, (long long)c1, (long long)c2, 1, a.dev, hp.s_bwd)) return rc;
    }
    if (a.with_bwd) {
      PS_CUDA(cudaStreamWaitEvent(hp.s_bwd, hp.ev_gin[s], 0));
      float* g1 = a.dev_g1 ? a.dev_g1 + h1 * 3 : d_g1;
      float* g2 = a.dev_g2 ? a.dev_g2 + h2 * 3 : d_g2;
      if (int rc = ps_chamfer_bwd(d_x1, d_x2, d_gd1, d_gd2, d_i1, d_i2, g1, g2, nb, a.N, a.M, a.dev, hp.s_bwd)) return rc;
    }
    // the path's single collective, after the last chunk's kernels: publish + wait over peer memory
    if (a.h_sums && a.comm && c == a.nchunks - 1)
      if (int rc = comm_allreduce_launch(a.comm, a.channel, hp.d_sums, hp.d_sums + 6, 6, hp.s_bwd)) return rc;
    PS_CUDA(cudaEventRecord(hp.ev_run[s], hp.s_bwd));

    // download: distances and indices as soon as the forward is done, gradients after the backward
    PS_CUDA(cudaStreamWaitEvent(hp.s_out, hp.ev_fwd[s], 0));
    if (a.dist1) COPY(a.dist1 + h1, d_d1, c1 * 4, cudaMemcpyDeviceToHost, hp.s_out);
    if (a.dist2) COPY(a.dist2 + h2, d_d2, c2 * 4, cudaMemcpyDeviceToHost, hp.s_out);
    if (a.idx1) COPY(a.idx1 + h1, d_i1, c1 * 4, cudaMemcpyDeviceToHost, hp.s_out);
    if (a.idx2) COPY(a.idx2 + h2, d_i2, c2 * 4, cudaMemcpyDeviceToHost, hp.s_out);
    PS_CUDA(cudaStreamWaitEvent(hp.s_out, hp.ev_run[s], 0));
    if (a.with_bwd && a.gradxyz1) COPY(a.gradxyz1 + h1 * 3, d_g1, c1 * 12, cudaMemcpyDeviceToHost, hp.s_out);
    if (a.with_bwd && a.gradxyz2) COPY(a.gradxyz2 + h2 * 3, d_g2, c2 * 12, cudaMemcpyDeviceToHost, hp.s_out);
    if (a.h_sums && c == a.nchunks - 1) COPY(a.h_sums, hp.d_sums + (a.comm ? 6 : 0), 6 * sizeof(double), cudaMemcpyDeviceToHost, hp.s_out);
    PS_CUDA(cudaEventRecord(hp.ev_out[s], hp.s_out));
  }
  // downloads are in order on s_out: `origin` continues once the last one has landed
  PS_CUDA(cudaEventRecord(hp.ev_end, hp.s_out));
  PS_CUDA(cudaStreamWaitEvent(origin, hp.ev_end, 0));
  return PS_OK;
}

// Submission in FIFO mode.  With several steps in flight the chunked multi-stream graph of enqueue_pipeline is the wrong
// shape: its purpose — overlapping the copies of a step with its own kernels — is now served by the OTHER steps, while
// its price stays (three small forward launches instead of one: 0.375 ms of kernels per C1 step instead of 0.33), and
// kernels of different steps that run side by side delay each other's downloads.  Here a step is three stages:
//     uploads (lane's copy stream)  ->  kernels (ONE stream shared by all lanes: strictly one step after the other,
//     each with the whole GPU, replayed as a graph)  ->  downloads (lane's copy stream)
// chained by events, so the copy engines work on the neighbouring steps while the SMs work on this one.
static int submit_fifo(HostPipe& P, Lane& L, const PipeArgs& a, int channel, cudaStream_t caller) {
#define COPY(...) PS_CUDA(cudaMemcpyAsync(__VA_ARGS__))
  if (!P.s_compute) PS_CUDA(cudaStreamCreateWithFlags(&P.s_compute, cudaStreamNonBlocking));
  char* base = static_cast<char*>(L.slot[0]);
  float* d_x1 = reinterpret_cast<float*>(base + a.o_x1);
  float* d_x2 = reinterpret_cast<float*>(base + a.o_x2);
  float* d_d1 = reinterpret_cast<float*>(base + a.o_d1);
  float* d_d2 = reinterpret_cast<float*>(base + a.o_d2);
  int* d_i1 = reinterpret_cast<int*>(base + a.o_i1);
  int* d_i2 = reinterpret_cast<int*>(base + a.o_i2);
  float* d_gd1 = reinterpret_cast<float*>(base + a.o_gd1);
  float* d_gd2 = reinterpret_cast<float*>(base + a.o_gd2);
  float* d_g1 = reinterpret_cast<float*>(base + a.o_g1);
  float* d_g2 = reinterpret_cast<float*>(base + a.o_g2);
  const size_t c1 = (size_t)a.B * a.N, c2 = (size_t)a.B * a.M;

  // uploads: behind the submitting stream's position and behind the lane's previous step (its slot is free again)
  PS_CUDA(cudaEventRecord(L.ev_sub, caller));
  PS_CUDA(cudaStreamWaitEvent(L.s_in, L.ev_sub, 0));
  PS_CUDA(cudaStreamWaitEvent(L.s_in, L.ev_last, 0));
  COPY(d_x1, a.xyz1, c1 * 12, cudaMemcpyHostToDevice, L.s_in);
  COPY(d_x2, a.xyz2, c2 * 12, cudaMemcpyHostToDevice, L.s_in);
  if (a.with_bwd) {
    COPY(d_gd1, a.graddist1, c1 * 4, cudaMemcpyHostToDevice, L.s_in);
    COPY(d_gd2, a.graddist2, c2 * 4, cudaMemcpyHostToDevice, L.s_in);
  }
  PS_CUDA(cudaEventRecord(L.ev_in[0], L.s_in));

  // kernels: forward (+ sums from its epilogue), backward, the exchange on the lane's channel
  PS_CUDA(cudaStreamWaitEvent(P.s_compute, L.ev_in[0], 0));
  auto kernels = [&](cudaStream_t s) -> int {
    if (int rc = chamfer_fwd_impl(d_x1, d_x2, d_d1, d_d2, d_i1, d_i2, a.h_sums ? L.d_sums : nullptr, nullptr, a.B, a.N, a.M, a.dev, s,
                                  "ps_chamfer_host_submit"))
      return rc;
    if (a.with_bwd)
      if (int rc = ps_chamfer_bwd(d_x1, d_x2, d_gd1, d_gd2, d_i1, d_i2, d_g1, d_g2, a.B, a.N, a.M, a.dev, s)) return rc;
    if (a.h_sums && a.comm) return comm_allreduce_launch(a.comm, channel, L.d_sums, L.d_sums + 6, 6, s);
    return PS_OK;
  };
  bool use_graph = true;
  if (const char* e = getenv("PS_HOST_GRAPH")) use_graph = atoi(e) != 0;
  int rc = PS_OK;
  GraphCache::Entry* hit = nullptr;
  if (use_graph) {
    GraphKey key;
    key.ptr(base);
    key.begin_shape();
    key.ptr(a.comm);
    key.val(a.B); key.val(a.N); key.val(a.M); key.val(a.with_bwd); key.val(a.h_sums != nullptr); key.val(channel);
    hit = P.cgraphs.get(key, L.s_cap, kernels, &rc);
    if (rc != PS_OK) return rc;
  }
  if (hit) {
    PS_CUDA(cudaGraphLaunch(hit->exec, P.s_compute));
    launch_counter() += hit->kernels;
  } else if (int rc2 = kernels(P.s_compute)) {
    return rc2;
  }
  PS_CUDA(cudaEventRecord(L.ev_cmp, P.s_compute));

  // downloads
  PS_CUDA(cudaStreamWaitEvent(L.s_out, L.ev_cmp, 0));
  if (a.dist1) COPY(a.dist1, d_d1, c1 * 4, cudaMemcpyDeviceToHost, L.s_out);
  if (a.dist2) COPY(a.dist2, d_d2, c2 * 4, cudaMemcpyDeviceToHost, L.s_out);
  if (a.idx1) COPY(a.idx1, d_i1, c1 * 4, cudaMemcpyDeviceToHost, L.s_out);
  if (a.idx2) COPY(a.idx2, d_i2, c2 * 4, cudaMemcpyDeviceToHost, L.s_out);
  if (a.with_bwd && a.gradxyz1) COPY(a.gradxyz1, d_g1, c1 * 12, cudaMemcpyDeviceToHost, L.s_out);
  if (a.with_bwd && a.gradxyz2) COPY(a.gradxyz2, d_g2, c2 * 12, cudaMemcpyDeviceToHost, L.s_out);
  if (a.h_sums) COPY(a.h_sums, L.d_sums + (a.comm ? 6 : 0), 6 * sizeof(double), cudaMemcpyDeviceToHost, L.s_out);
  PS_CUDA(cudaEventRecord(L.ev_last, L.s_out));
  return PS_OK;
#undef COPY
}

static void drop_graphs(Lane& hp) { hp.graphs.drop(); }

}  // namespace ps

using namespace ps;

static int host_pipeline_call(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int* idx1,
                              int* idx2, const float* graddist1, const float* graddist2, float* gradxyz1,
                              float* gradxyz2, float* dev_g1, float* dev_g2, double* h_sums, ps_comm* comm, int B, int N,
                              int M, int chunk, int dev, void* stream_, long long* ticket = nullptr) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool with_bwd = graddist1 != nullptr || graddist2 != nullptr;
  HostPipe* hpp = pipe_for(dev);
  PS_REQUIRE(hpp != nullptr, "ps_chamfer_host: bad device %d", dev);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_chamfer_host: cannot select device %d", dev);
  std::lock_guard<std::mutex> lock(hpp->mu);
  // stream-ordered call: lane 0 on the caller's stream.  Submission (ticket != null): the next lane in turn, on
  // the lane's own launch stream, behind the submitting stream's current position.
  const int li = ticket ? (int)(hpp->submits++ % LANES) : 0;
  Lane& hp = hpp->lane[li];
  if (int rc = pipe_init(hp, dev)) return rc;
  if (ticket) {
    PS_CUDA(cudaEventRecord(hp.ev_sub, stream));
    stream = hp.s_launch;
    PS_CUDA(cudaStreamWaitEvent(stream, hp.ev_sub, 0));
    // with a communicator every lane exchanges on its own channel (1 + lane): the steps of one lane run in
    // submission order on every rank, and lanes never wait for each other — provided every rank submits the same
    // sequence of steps, as SPMD ranks do
    *ticket = (long long)(((++hp.calls) << 8) | (unsigned)(li + 1));
  }

  // submissions: FIFO mode (one chunk, kernels of all lanes on one stream) unless PS_HOST_ASYNC=lanes asks for the
  // chunked multi-stream graph per lane
  bool fifo = ticket != nullptr;
  if (const char* e = getenv("PS_HOST_ASYNC")) fifo = fifo && e[0] != 'l';
  PipeArgs a;
  if (fifo) chunk = B;
  make_plan(B, chunk, a.sizes, &a.nchunks, &chunk);
  a.xyz1 = xyz1; a.xyz2 = xyz2; a.graddist1 = graddist1; a.graddist2 = graddist2;
  a.dist1 = dist1; a.dist2 = dist2; a.gradxyz1 = gradxyz1; a.gradxyz2 = gradxyz2;
  a.idx1 = idx1; a.idx2 = idx2;
  a.dev_g1 = dev_g1; a.dev_g2 = dev_g2; a.h_sums = h_sums; a.comm = comm;
  a.channel = (comm && ticket) ? 1 + li : 0;
  a.B = B; a.N = N; a.M = M; a.chunk = chunk; a.dev = dev; a.with_bwd = with_bwd;
  // slot layout (all sub-buffers 256-byte aligned so the vectorised kernels see aligned clouds)
  const size_t n1 = (size_t)chunk * N, n2 = (size_t)chunk * M;
  size_t off = 0;
  a.o_x1 = off; off += align256(n1 * 12);
  a.o_x2 = off; off += align256(n2 * 12);
  a.o_d1 = off; off += align256(n1 * 4);
  a.o_d2 = off; off += align256(n2 * 4);
  a.o_i1 = off; off += align256(n1 * 4);
  a.o_i2 = off; off += align256(n2 * 4);
  a.o_gd1 = off; off += with_bwd ? align256(n1 * 4) : 0;
  a.o_gd2 = off; off += with_bwd ? align256(n2 * 4) : 0;
  a.o_g1 = off; off += with_bwd ? align256(n1 * 12) : 0;
  a.o_g2 = off; off += with_bwd ? align256(n2 * 12) : 0;
  if (off > hp.slot_bytes) {
    PS_CUDA(cudaDeviceSynchronize());
    drop_graphs(hp);  // they hold the old slot addresses
    for (int i = 0; i < SLOTS; i++) {
      if (hp.slot[i]) PS_CUDA(cudaFree(hp.slot[i]));
      hp.slot[i] = nullptr;
    }
    hp.slot_bytes = 0;
    for (int i = 0; i < SLOTS; i++) PS_CUDA(cudaMalloc(&hp.slot[i], off));
    hp.slot_bytes = off;
  }

  if (fifo) return submit_fifo(*hpp, hp, a, a.channel, static_cast<cudaStream_t>(stream_));

  bool use_graph = true;
  if (const char* e = getenv("PS_HOST_GRAPH")) use_graph = atoi(e) != 0;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  const bool caller_capturing = cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone;
  if (caller_capturing) {
    use_graph = false;  // become part of the caller's graph instead
    cudaGetLastError();
  } else {
    // calls on one device share the staging slots: each starts behind the end of the previous one
    PS_CUDA(cudaStreamWaitEvent(stream, hp.ev_last, 0));
  }
  if (use_graph) {
    // copies from / to pageable memory cannot be captured (the runtime stages them synchronously)
    const void* hostp[11] = {xyz1, xyz2, dist1, dist2, idx1, idx2, graddist1, graddist2, gradxyz1, gradxyz2, h_sums};
    for (const void* hptr : hostp) {
      if (!hptr) continue;
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, hptr) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
        use_graph = false;
        cudaGetLastError();
        break;
      }
    }
  }
  int rc = PS_OK;
  if (use_graph) {
    GraphKey key;
    const void* ptrs[13] = {xyz1, xyz2, dist1, dist2, idx1, idx2, graddist1, graddist2, gradxyz1, gradxyz2, dev_g1, dev_g2, h_sums};
    for (const void* q : ptrs) key.ptr(q);
    key.begin_shape();
    key.ptr(comm);
    key.val(B); key.val(N); key.val(M); key.val(a.nchunks);
    for (int i = 0; i < a.nchunks; i++) key.val(a.sizes[i]);
    GraphCache::Entry* hit = hp.graphs.get(key, hp.s_cap, [&](cudaStream_t s) { return enqueue_pipeline(hp, a, s); }, &rc);
    if (rc != PS_OK) return rc;
    if (hit) {
      PS_CUDA(cudaGraphLaunch(hit->exec, stream));
      launch_counter() += hit->kernels;
      PS_CUDA(cudaEventRecord(hp.ev_last, stream));
      return PS_OK;
    }
  }
  rc = enqueue_pipeline(hp, a, stream);
  if (rc) return rc;
  if (!caller_capturing) PS_CUDA(cudaEventRecord(hp.ev_last, stream));
  return PS_OK;
}

extern "C" int ps_chamfer_host(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int* idx1,
                               int* idx2, const float* graddist1, const float* graddist2, float* gradxyz1,
                               float* gradxyz2, int B, int N, int M, int chunk, int dev, void* stream) {
  PS_REQUIRE(B >= 0 && N >= 0 && M >= 0, "ps_chamfer_host: negative size");
  if (B == 0 || (N == 0 && M == 0)) return PS_OK;
  PS_REQUIRE(N > 0 && M > 0, "ps_chamfer_host: both clouds need at least one point (N=%d, M=%d)", N, M);
  PS_REQUIRE(xyz1 && xyz2 && dist1 && dist2 && idx1 && idx2, "ps_chamfer_host: null pointer");
  if (graddist1 != nullptr || graddist2 != nullptr)
    PS_REQUIRE(graddist1 && graddist2 && gradxyz1 && gradxyz2, "ps_chamfer_host: backward needs graddist1, graddist2, gradxyz1, gradxyz2");
  return host_pipeline_call(xyz1, xyz2, dist1, dist2, idx1, idx2, graddist1, graddist2, gradxyz1, gradxyz2, nullptr, nullptr,
                            nullptr, nullptr, B, N, M, chunk, dev, stream);
}

// General form: every output of the step downloaded (dist, idx, gradients) AND the loss sums, which are world-wide
// when a communicator is given.  sums6 / comm may be NULL.
extern "C" int ps_chamfer_host_full(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int* idx1, int* idx2,
                                    const float* graddist1, const float* graddist2, float* gradxyz1, float* gradxyz2,
                                    double* sums6, ps_comm* comm, int B, int N, int M, int chunk, int dev, void* stream) {
  PS_REQUIRE(B >= 0 && N >= 0 && M >= 0, "ps_chamfer_host_full: negative size");
  if (B == 0 || (N == 0 && M == 0)) return PS_OK;
  PS_REQUIRE(N > 0 && M > 0, "ps_chamfer_host_full: both clouds need at least one point (N=%d, M=%d)", N, M);
  PS_REQUIRE(xyz1 && xyz2 && dist1 && dist2 && idx1 && idx2, "ps_chamfer_host_full: null pointer");
  if (graddist1 != nullptr || graddist2 != nullptr)
    PS_REQUIRE(graddist1 && graddist2 && gradxyz1 && gradxyz2, "ps_chamfer_host_full: backward needs graddist1, graddist2, gradxyz1, gradxyz2");
  if (comm) PS_REQUIRE(sums6 && comm->connected && comm->dev == dev, "ps_chamfer_host_full: communicator needs sums6, a connection and device %d", dev);
  return host_pipeline_call(xyz1, xyz2, dist1, dist2, idx1, idx2, graddist1, graddist2, gradxyz1, gradxyz2, nullptr, nullptr,
                            sums6, comm, B, N, M, chunk, dev, stream);
}

static int host_step_call(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                          float* dev_gradxyz1, float* dev_gradxyz2, double* sums6, ps_comm* comm, int B, int N, int M,
                          int chunk, int dev, void* stream) {
  PS_REQUIRE(B >= 0 && N >= 0 && M >= 0, "ps_chamfer_host_step: negative size");
  if (B == 0 || (N == 0 && M == 0)) return PS_OK;
  PS_REQUIRE(N > 0 && M > 0, "ps_chamfer_host_step: both clouds need at least one point (N=%d, M=%d)", N, M);
  PS_REQUIRE(xyz1 && xyz2 && sums6, "ps_chamfer_host_step: null pointer");
  const bool with_bwd = graddist1 != nullptr || graddist2 != nullptr;
  if (with_bwd)
    PS_REQUIRE(graddist1 && graddist2 && dev_gradxyz1 && dev_gradxyz2,
               "ps_chamfer_host_step: backward needs graddist1, graddist2 (host) and dev_gradxyz1, dev_gradxyz2 (device)");
  if (comm) PS_REQUIRE(comm->connected && comm->dev == dev, "ps_chamfer_host_step_dist: communicator not connected or on another device");
  return host_pipeline_call(xyz1, xyz2, nullptr, nullptr, nullptr, nullptr, graddist1, graddist2, nullptr, nullptr,
                            with_bwd ? dev_gradxyz1 : nullptr, with_bwd ? dev_gradxyz2 : nullptr, sums6, comm, B, N, M, chunk, dev, stream);
}

extern "C" int ps_chamfer_host_step(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                                    float* dev_gradxyz1, float* dev_gradxyz2, double* sums6, int B, int N, int M,
                                    int chunk, int dev, void* stream) {
  return host_step_call(xyz1, xyz2, graddist1, graddist2, dev_gradxyz1, dev_gradxyz2, sums6, nullptr, B, N, M, chunk, dev, stream);
}

extern "C" int ps_chamfer_host_step_dist(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                                         float* dev_gradxyz1, float* dev_gradxyz2, double* sums6, ps_comm* comm, int B, int N,
                                         int M, int chunk, int dev, void* stream) {
  PS_REQUIRE(comm != nullptr, "ps_chamfer_host_step_dist: null communicator");
  return host_step_call(xyz1, xyz2, graddist1, graddist2, dev_gradxyz1, dev_gradxyz2, sums6, comm, B, N, M, chunk, dev, stream);
}

// Graph-cache statistics of the host-buffer entry points on `dev` (see ps_chamfer_step_stats).
extern "C" int ps_chamfer_host_stats(int dev, long long* hits, long long* updates, long long* instantiations) {
  HostPipe* hp = pipe_for(dev);
  PS_REQUIRE(hp != nullptr, "ps_chamfer_host_stats: bad device %d", dev);
  std::lock_guard<std::mutex> lock(hp->mu);
  long long h = 0, u = 0, n = 0;
  for (const Lane& l : hp->lane) { h += l.graphs.hits; u += l.graphs.updates; n += l.graphs.instantiations; }
  if (hits) *hits = h;
  if (updates) *updates = u;
  if (instantiations) *instantiations = n;
  return PS_OK;
}

// Asynchronous form of ps_chamfer_host_full for loops that keep more than one step in flight (a loader that
// prefetches the next batch while the previous results are read): the call is ordered behind the current position
// of `stream`, runs on one of the library's lanes (in turn) and does NOT join `stream` again.  *ticket identifies
// it for ps_chamfer_host_wait.  Consecutive submissions overlap: upload and kernels of step i+1 with the download of
// step i.  With a communicator each lane exchanges on its own channel: every rank must submit the same sequence.
extern "C" int ps_chamfer_host_submit(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int* idx1, int* idx2,
                                      const float* graddist1, const float* graddist2, float* gradxyz1, float* gradxyz2,
                                      double* sums6, ps_comm* comm, int B, int N, int M, int chunk, int dev, void* stream,
                                      long long* ticket) {
  PS_REQUIRE(ticket != nullptr, "ps_chamfer_host_submit: null ticket");
  *ticket = 0;
  PS_REQUIRE(B >= 0 && N >= 0 && M >= 0, "ps_chamfer_host_submit: negative size");
  if (B == 0 || (N == 0 && M == 0)) return PS_OK;
  PS_REQUIRE(N > 0 && M > 0, "ps_chamfer_host_submit: both clouds need at least one point (N=%d, M=%d)", N, M);
  PS_REQUIRE(xyz1 && xyz2 && dist1 && dist2 && idx1 && idx2, "ps_chamfer_host_submit: null pointer");
  if (graddist1 != nullptr || graddist2 != nullptr)
    PS_REQUIRE(graddist1 && graddist2 && gradxyz1 && gradxyz2, "ps_chamfer_host_submit: backward needs graddist1, graddist2, gradxyz1, gradxyz2");
  if (comm) PS_REQUIRE(sums6 && comm->connected && comm->dev == dev, "ps_chamfer_host_submit: communicator needs sums6, a connection and device %d", dev);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  PS_REQUIRE(cudaStreamIsCapturing(static_cast<cudaStream_t>(stream), &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone,
             "ps_chamfer_host_submit: `stream` is capturing (use ps_chamfer_host_full inside a capture)");
  return host_pipeline_call(xyz1, xyz2, dist1, dist2, idx1, idx2, graddist1, graddist2, gradxyz1, gradxyz2, nullptr, nullptr,
                            sums6, comm, B, N, M, chunk, dev, stream, ticket);
}

// Joins a submission: `stream` (when wait_stream != 0) continues only after the submitted step has written its last
// output byte; block != 0 also blocks the calling host thread until then.  ticket 0 (an empty submission) is a no-op.
// A lane is reused every LANES submissions: waiting for an old ticket waits for the lane's latest step, which is
// later in the same stream order and therefore still sufficient.
extern "C" int ps_chamfer_host_wait(long long ticket, int dev, void* stream, int wait_stream, int block) {
  if (ticket == 0) return PS_OK;
  const int li = (int)(ticket & 0xff) - 1;
  HostPipe* hpp = pipe_for(dev);
  PS_REQUIRE(hpp != nullptr && li >= 0 && li < LANES, "ps_chamfer_host_wait: bad ticket or device %d", dev);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_chamfer_host_wait: cannot select device %d", dev);
  cudaEvent_t ev;
  {
    std::lock_guard<std::mutex> lock(hpp->mu);
    PS_REQUIRE(hpp->lane[li].ready, "ps_chamfer_host_wait: nothing was submitted on device %d", dev);
    ev = hpp->lane[li].ev_last;
  }
  if (wait_stream) PS_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), ev, 0));
  if (block) PS_CUDA(cudaEventSynchronize(ev));
  return PS_OK;
}
