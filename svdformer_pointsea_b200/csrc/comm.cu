// Host side of the peer-memory sums exchange (see comm.cuh) + the stand-alone all-reduce entry point.
#include "comm.cuh"

#include <cstring>

namespace ps {

__global__ void __launch_bounds__(32) comm_allreduce_kernel(const CommDev c, const double* __restrict__ in, double* __restrict__ out, int n) {
  __shared__ double vals[COMM_MAX_N];
  if ((int)threadIdx.x < n) vals[threadIdx.x] = in[threadIdx.x];
  __syncwarp();
  comm_publish(c, vals, n);
  comm_wait_reduce(c, out, n);
}

__global__ void __launch_bounds__(32) comm_publish_kernel(const CommDev c, const double* __restrict__ in, int n) {
  __shared__ double vals[COMM_MAX_N];
  if ((int)threadIdx.x < n) vals[threadIdx.x] = in[threadIdx.x];
  __syncwarp();
  comm_publish(c, vals, n);
}

CommDev comm_channel(const ps_comm* comm, int ch) {
  CommDev d = comm->d;
  if (ch > 0 && ch < COMM_CHANNELS) {
    for (int r = 0; r < comm->world; r++)
      if (d.peer[r]) d.peer[r] += (size_t)ch * COMM_DEPTH * comm->world;
    d.pub_seq = comm->counters + 2 * ch;
    d.wait_seq = comm->counters + 2 * ch + 1;
  }
  return d;
}

int comm_allreduce_launch(const ps_comm* comm, int ch, const double* in, double* out, int n, cudaStream_t stream) {
  comm_allreduce_kernel<<<1, 32, 0, stream>>>(comm_channel(comm, ch), in, out, n);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

int comm_publish_launch(const ps_comm* comm, const double* in, int n, cudaStream_t stream) {
  comm_publish_kernel<<<1, 32, 0, stream>>>(comm->d, in, n);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

__global__ void __launch_bounds__(32) comm_wait_kernel(const CommDev c, double* __restrict__ out, int n) { comm_wait_reduce(c, out, n); }

int comm_wait_launch(const ps_comm* comm, double* out, int n, cudaStream_t stream) {
  comm_wait_kernel<<<1, 32, 0, stream>>>(comm->d, out, n);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

}  // namespace ps

using namespace ps;

extern "C" int ps_comm_create(int rank, int world, int dev, ps_comm** out) {
  PS_REQUIRE(out != nullptr, "ps_comm_create: null output");
  PS_REQUIRE(world >= 1 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world, "ps_comm_create: bad rank %d / world %d (max %d)", rank,
             world, COMM_MAX_WORLD);
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_comm_create: cannot select device %d", dev);
  ps_comm* c = new ps_comm();
  c->rank = rank; c->world = world; c->dev = dev;
  const size_t bytes = sizeof(CommSlot) * COMM_CHANNELS * COMM_DEPTH * world;
  // plain cudaMalloc (not a pool allocation): CUDA IPC can only export such memory
  cudaError_t e = cudaMalloc((void**)&c->mailbox, bytes);
  if (e == cudaSuccess) e = cudaMemset(c->mailbox, 0, bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->counters, 128);
  if (e == cudaSuccess) e = cudaMemset(c->counters, 0, 128);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (c->mailbox) cudaFree(c->mailbox);
    if (c->counters) cudaFree(c->counters);
    delete c;
    return set_error(PS_ERR_CUDA, "ps_comm_create: %s", cudaGetErrorString(e));
  }
  memset(&c->d, 0, sizeof(c->d));
  c->d.rank = rank; c->d.world = world;
  c->d.pub_seq = c->counters; c->d.wait_seq = c->counters + 1; c->d.err = reinterpret_cast<int*>(c->counters + 2 * COMM_CHANNELS);
  c->d.peer[rank] = c->mailbox;
  c->connected = world == 1;
  *out = c;
  return PS_OK;
}

extern "C" int ps_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int ps_comm_export(ps_comm* c, void* handle) {
  PS_REQUIRE(c && handle, "ps_comm_export: null argument");
  DeviceGuard guard(c->dev);
  cudaIpcMemHandle_t h;
  PS_CUDA(cudaIpcGetMemHandle(&h, c->mailbox));
  memcpy(handle, &h, sizeof(h));
  return PS_OK;
}

extern "C" int ps_comm_connect(ps_comm* c, const void* handles) {
  PS_REQUIRE(c && handles, "ps_comm_connect: null argument");
  DeviceGuard guard(c->dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_comm_connect: cannot select device %d", c->dev);
  const char* hp = static_cast<const char*>(handles);
  for (int r = 0; r < c->world; r++) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hp + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    PS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->opened[r] = p;
    c->d.peer[r] = static_cast<CommSlot*>(p);
  }
  c->connected = true;
  return PS_OK;
}

extern "C" int ps_comm_connect_local(ps_comm* const* comms, int world) {
  PS_REQUIRE(comms && world >= 1 && world <= COMM_MAX_WORLD, "ps_comm_connect_local: bad arguments");
  for (int r = 0; r < world; r++) PS_REQUIRE(comms[r] && comms[r]->world == world && comms[r]->rank == r, "ps_comm_connect_local: comms[%d] is not rank %d of %d", r, r, world);
  for (int a = 0; a < world; a++) {
    DeviceGuard guard(comms[a]->dev);
    for (int b = 0; b < world; b++) {
      if (comms[b]->dev != comms[a]->dev) {
        int can = 0;
        PS_CUDA(cudaDeviceCanAccessPeer(&can, comms[a]->dev, comms[b]->dev));
        PS_REQUIRE(can, "ps_comm_connect_local: device %d cannot access device %d", comms[a]->dev, comms[b]->dev);
        const cudaError_t e = cudaDeviceEnablePeerAccess(comms[b]->dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) PS_CUDA(e);
        cudaGetLastError();
      }
      comms[a]->d.peer[b] = comms[b]->mailbox;
    }
    comms[a]->connected = true;
  }
  return PS_OK;
}

extern "C" int ps_comm_allreduce(ps_comm* c, const double* in, double* out, int n, void* stream) {
  PS_REQUIRE(c && in && out, "ps_comm_allreduce: null argument");
  PS_REQUIRE(n >= 1 && n <= COMM_MAX_N, "ps_comm_allreduce: n=%d outside 1..%d", n, COMM_MAX_N);
  PS_REQUIRE(c->connected, "ps_comm_allreduce: communicator is not connected (ps_comm_connect)");
  DeviceGuard guard(c->dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_comm_allreduce: cannot select device %d", c->dev);
  comm_allreduce_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(c->d, in, out, n);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

extern "C" int ps_comm_status(ps_comm* c, long long* published, long long* consumed, int* timed_out) {
  PS_REQUIRE(c != nullptr, "ps_comm_status: null communicator");
  DeviceGuard guard(c->dev);
  unsigned long long h[2 * COMM_CHANNELS + 1] = {0};
  PS_CUDA(cudaMemcpy(h, c->counters, sizeof(h), cudaMemcpyDeviceToHost));  // synchronises: diagnostics only
  long long pub = 0, con = 0;
  for (int ch = 0; ch < COMM_CHANNELS; ch++) { pub += (long long)h[2 * ch]; con += (long long)h[2 * ch + 1]; }
  if (published) *published = pub;  // all channels
  if (consumed) *consumed = con;
  if (timed_out) *timed_out = (int)(h[2 * COMM_CHANNELS] & 0xffffffffu);
  return PS_OK;
}

extern "C" int ps_comm_destroy(ps_comm* c) {
  if (!c) return PS_OK;
  DeviceGuard guard(c->dev);
  cudaDeviceSynchronize();
  for (int r = 0; r < COMM_MAX_WORLD; r++)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->opened[r]);
  if (c->mailbox) cudaFree(c->mailbox);
  if (c->counters) cudaFree(c->counters);
  cudaGetLastError();
  delete c;
  return PS_OK;
}
