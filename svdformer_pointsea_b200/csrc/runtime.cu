// Library plumbing: error strings, device queries, launch counter, fp32 peak probe.
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace ps {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
long long& launch_counter() {
  static thread_local long long n = 0;
  return n;
}
Workspace*& tls_workspace() {
  static thread_local Workspace* w = nullptr;
  return w;
}
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count(int dev) {
  // per-device cache; relaxed atomics: every writer stores the same value
  static std::atomic<int> cache[64];
  if (dev >= 0 && dev < 64) {
    const int v = cache[dev].load(std::memory_order_relaxed);
    if (v) return v;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

// Stream-ordered scratch from a PRIVATE pool per device.  A pool gives memory back to the driver at every
// synchronisation unless a release threshold is set, which turns each cudaMallocAsync after a sync into a
// real (100+ us) allocation; the library's pool keeps up to PS_SCRATCH_KEEP_MB (default 1024) cached and
// leaves the device's default pool (shared with everything else in the process) untouched.
static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64];
static bool g_pool_ready[64];

static cudaMemPool_t scratch_pool(int dev) {
  if (dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_pool_mu);
  if (!g_pool_ready[dev]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
      unsigned long long keep = 1024ull << 20;
      if (const char* e = getenv("PS_SCRATCH_KEEP_MB")) keep = (unsigned long long)atoll(e) << 20;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      g_pools[dev] = pool;
    } else {
      cudaGetLastError();
      g_pools[dev] = nullptr;  // fall back to the default pool, untouched
    }
    g_pool_ready[dev] = true;
  }
  return g_pools[dev];
}

void scratch_pool_init(int dev) { (void)scratch_pool(dev); }

int scratch_alloc(void** ptr, size_t bytes, int dev, cudaStream_t stream) {
  cudaMemPool_t pool = scratch_pool(dev);
  if (pool) PS_CUDA(cudaMallocFromPoolAsync(ptr, bytes, pool, stream));
  else PS_CUDA(cudaMallocAsync(ptr, bytes, stream));
  return PS_OK;
}

// cudaMemsetAsync on memory that came from cudaMallocAsync costs ~130 us of HOST time per call on this stack
// (measured inside ps_chamfer_fwd: malloc 14, memset 133, main launch 5, unpack 9 us); a fill kernel is one
// ordinary launch (~4 us) and runs at memset speed.
__global__ void __launch_bounds__(256) fill32_kernel(uint4* __restrict__ p, unsigned v, size_t n16, unsigned* __restrict__ tail, int ntail) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = i; j < n16; j += stride) p[j] = make_uint4(v, v, v, v);
  if (i < (size_t)ntail) tail[i] = v;
}
int fill32_async(void* ptr, unsigned value, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return PS_OK;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (bytes & 3) != 0) {  // odd cases keep the runtime's memset
    PS_CUDA(cudaMemsetAsync(ptr, (int)(value & 0xff), bytes, stream));
    return PS_OK;
  }
  const size_t n16 = bytes / 16;
  const int ntail = (int)((bytes - n16 * 16) / 4);
  size_t blocks = (n16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  fill32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<uint4*>(ptr), value, n16,
                                                     reinterpret_cast<unsigned*>(static_cast<char*>(ptr) + n16 * 16), ntail);
  PS_LAUNCH_CHECK();
  return PS_OK;
}

// FFMA2-only kernel: 16 independent packed accumulators per thread.
constexpr int PEAK_ITERS = 2048;
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, float a, float b) {
  u64 acc[16];
  const u64 a2 = pack2(a, a), b2 = pack2(b, b);
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = pack2(threadIdx.x * 0.001f + i, (float)i);
  for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a2), "l"(b2));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) s += lo2(acc[i]) + hi2(acc[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace ps

using namespace ps;

extern "C" int ps_version(void) { return 100; }  // 0.1.0

extern "C" const char* ps_last_error(void) { return err_buf(); }

extern "C" int ps_device_info(int dev, int* sm, int* major, int* minor) {
  int a = 0, b = 0, c = 0;
  PS_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  PS_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  PS_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm) *sm = a;
  if (major) *major = b;
  if (minor) *minor = c;
  return PS_OK;
}

extern "C" long long ps_launch_count(int reset) {
  const long long n = launch_counter();
  if (reset) launch_counter() = 0;
  return n;
}

extern "C" int ps_measure_fp32_peak(int dev, int reps, double* tflops) {
  PS_REQUIRE(tflops != nullptr && reps > 0, "ps_measure_fp32_peak: bad arguments");
  DeviceGuard guard(dev);
  if (!guard.ok) return set_error(PS_ERR_CUDA, "ps_measure_fp32_peak: cannot select device %d", dev);
  const int grid = sm_count(dev) * 16, threads = 256;
  float* out = nullptr;
  PS_CUDA(cudaMalloc((void**)&out, (size_t)grid * threads * sizeof(float)));
  cudaEvent_t e0, e1;
  PS_CUDA(cudaEventCreate(&e0));
  PS_CUDA(cudaEventCreate(&e1));
  fp32_peak_kernel<<<grid, threads>>>(out, 1.0001f, 0.5f);
  PS_CUDA(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    PS_CUDA(cudaEventRecord(e0));
    fp32_peak_kernel<<<grid, threads>>>(out, 1.0001f, 0.5f);
    PS_CUDA(cudaEventRecord(e1));
    PS_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    PS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  const double flops = (double)grid * threads * PEAK_ITERS * 16.0 * 4.0;  // 2 lanes x (mul+add)
  *tflops = flops / (best * 1e-3) / 1e12;
  return PS_OK;
}
