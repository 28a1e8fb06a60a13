// One-launch Chamfer training step on DEVICE buffers:
//     forward (both directions, argmin indices) -> loss sums (same epilogue launch) -> [publish the sums to the
//     peers' mailboxes from that launch's last block] -> backward -> [one-warp wait: world-wide sums]
// i.e. what utils/loss_utils.py:10-19 + dist_chamfer_3D.py:26-64 do per loss term, plus the single collective of
// the batch-sharded path.  The five or six kernels are captured once per argument set into a CUDA
// graph (graph_cache.cuh) and replayed with ONE cudaGraphLaunch: the host cost of a step is a few microseconds,
// which is what keeps 8 ranks in lockstep (issuing ~7 launches + a c10d all-reduce per step took ~1 ms of host
// time per 0.35 ms device step on an 8-GPU box: SCALE_r01, efficiency 0.56).
#include "comm.cuh"
#include "graph_cache.cuh"

#include <cstdlib>
#include <mutex>

namespace ps {

struct StepCtx {
  std::mutex mu;
  bool ready = false;
  cudaStream_t s_cap = nullptr;
  GraphCache graphs;
};

static StepCtx* step_ctx(int dev) {
  static StepCtx ctx[64];
  return (dev >= 0 && dev < 64) ? &ctx[dev] : nullptr;
}

struct StepArgs {
  const float *xyz1, *xyz2, *gd1, *gd2;
  float *dist1, *dist2, *g1, *g2;
  int *idx1, *idx2;
  double *sums_local, *sums_global;
  ps_comm* comm;
  int B, N, M, dev;
};

static int enqueue_step(const StepArgs& a, cudaStream_t s) {
  if (int rc = chamfer_fwd_impl(a.xyz1, a.xyz2, a.dist1, a.dist2, a.idx1, a.idx2, a.sums_local, a.comm, a.B, a.N, a.M, a.dev, s,
                                "ps_chamfer_step"))
    return rc;
  if (a.gd1 && a.B > 0)
    if (int rc = ps_chamfer_bwd(a.xyz1, a.xyz2, a.gd1, a.gd2, a.idx1, a.idx2, a.g1, a.g2, a.B, a.N, a.M, a.dev, s)) return rc;
  // the wait comes last: the backward does not consume the world-wide sums, so a rank that is ahead computes
  // its backward while the slower ranks are still publishing
  if (a.comm) return comm_wait_launch(a.comm, a.sums_global, 6, s);
  return PS_OK;
}

}  // namespace ps

using namespace ps;

static int chamfer_step_impl(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                             float* dist1, float* dist2, int* idx1, int* idx2, float* gradxyz1, float* gradxyz2,
                             double* sums_local6, double* sums_global6, ps_comm* comm, int B, int N, int M, int dev,
                             void* stream_, bool launch) {
  PS_REQUIRE(B >= 0 && N > 0 && M > 0, "ps_chamfer_step: bad sizes B=%d N=%d M=%d", B, N, M);
  PS_REQUIRE(sums_local6 != nullptr, "ps_chamfer_step: null sums_local6");
  PS_REQUIRE(B == 0 || (xyz1 && xyz2 && dist1 && dist2 && idx1 && idx2), "ps_chamfer_step: null pointer");
  const bool with_bwd = graddist1 != nullptr || graddist2 != nullptr;
  if (with_bwd) PS_REQUIRE(graddist1 && graddist2 && gradxyz1 && gradxyz2, "ps_chamfer_step: backward needs graddist1, graddist2, gradxyz1, gradxyz2");
  if (comm) {
    PS_REQUIRE(sums_global6 != nullptr, "ps_chamfer_step: a communicator needs sums_global6");
    PS_REQUIRE(comm->connected && comm->dev == dev, "ps_chamfer_step: communicator not connected or on another device (%d vs %d)", comm->dev, dev);
  }
  StepCtx* cp = step_ctx(dev);
  PS_REQUIRE(cp != nullptr, "ps_chamfer_step: bad device %d", dev);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_chamfer_step: cannot select device %d", dev);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  StepCtx& cx = *cp;
  std::lock_guard<std::mutex> lock(cx.mu);
  if (!cx.ready) {
    PS_CUDA(cudaStreamCreateWithFlags(&cx.s_cap, cudaStreamNonBlocking));
    scratch_pool_init(dev);
    cx.ready = true;
  }
  StepArgs a{xyz1, xyz2, with_bwd ? graddist1 : nullptr, with_bwd ? graddist2 : nullptr, dist1, dist2,
             with_bwd ? gradxyz1 : nullptr, with_bwd ? gradxyz2 : nullptr, idx1, idx2, sums_local6, comm ? sums_global6 : nullptr,
             comm, B, N, M, dev};

  bool use_graph = true;
  if (const char* e = getenv("PS_STEP_GRAPH")) use_graph = atoi(e) != 0;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (launch && (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)) {
    cudaGetLastError();
    return enqueue_step(a, stream);  // become part of the caller's own graph (stream-ordered scratch)
  }
  if (!use_graph) return launch ? enqueue_step(a, stream) : PS_OK;

  GraphKey key;
  key.ptr(xyz1); key.ptr(xyz2); key.ptr(a.gd1); key.ptr(a.gd2); key.ptr(dist1); key.ptr(dist2); key.ptr(idx1); key.ptr(idx2);
  key.ptr(a.g1); key.ptr(a.g2); key.ptr(sums_local6); key.ptr(a.sums_global);
  key.begin_shape();
  key.ptr(comm);  // the mailbox addresses are baked into the kernel parameters
  key.val(B); key.val(N); key.val(M);

  int rc = PS_OK;
  GraphCache::Entry* exec = cx.graphs.get(key, cx.s_cap, [&](cudaStream_t s) { return enqueue_step(a, s); }, &rc);
  if (rc != PS_OK) return rc;
  if (!launch) return PS_OK;
  if (exec) {
    PS_CUDA(cudaGraphLaunch(exec->exec, stream));
    launch_counter() += exec->kernels;
    return PS_OK;
  }
  return enqueue_step(a, stream);
}

extern "C" int ps_chamfer_step(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                               float* dist1, float* dist2, int* idx1, int* idx2, float* gradxyz1, float* gradxyz2,
                               double* sums_local6, double* sums_global6, ps_comm* comm, int B, int N, int M, int dev,
                               void* stream) {
  return chamfer_step_impl(xyz1, xyz2, graddist1, graddist2, dist1, dist2, idx1, idx2, gradxyz1, gradxyz2, sums_local6,
                           sums_global6, comm, B, N, M, dev, stream, true);
}

// Captures and instantiates the graph of this exact call without launching anything (no stream work at all).
extern "C" int ps_chamfer_step_prepare(const float* xyz1, const float* xyz2, const float* graddist1, const float* graddist2,
                                       float* dist1, float* dist2, int* idx1, int* idx2, float* gradxyz1, float* gradxyz2,
                                       double* sums_local6, double* sums_global6, ps_comm* comm, int B, int N, int M, int dev) {
  return chamfer_step_impl(xyz1, xyz2, graddist1, graddist2, dist1, dist2, idx1, idx2, gradxyz1, gradxyz2, sums_local6,
                           sums_global6, comm, B, N, M, dev, nullptr, false);
}

// Cache statistics of ps_chamfer_step on `dev`: exact replays, in-place retargets (cudaGraphExecUpdate), instantiations.
extern "C" int ps_chamfer_step_stats(int dev, long long* hits, long long* updates, long long* instantiations) {
  StepCtx* cp = step_ctx(dev);
  PS_REQUIRE(cp != nullptr, "ps_chamfer_step_stats: bad device %d", dev);
  std::lock_guard<std::mutex> lock(cp->mu);
  if (hits) *hits = cp->graphs.hits;
  if (updates) *updates = cp->graphs.updates;
  if (instantiations) *instantiations = cp->graphs.instantiations;
  return PS_OK;
}
