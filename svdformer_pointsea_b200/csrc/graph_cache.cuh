// Small LRU cache of instantiated CUDA graphs, shared by the one-launch step entry points
// (step.cu: device buffers, host_pipeline.cu: host buffers).
//
// A call is described by a KEY (buffer addresses + sizes: a replay is only valid for identical arguments) and a
// SHAPE (sizes and which optional buffers are present: calls of equal shape capture graphs of identical
// topology).  Exact key hit: one cudaGraphLaunch.  Shape hit with other addresses (a data loader that hands
// out fresh buffers every step): the call is re-captured — cheap, nothing is submitted — and the existing
// executable is UPDATED in place with cudaGraphExecUpdate instead of being re-instantiated (the expensive
// part).  Miss: capture + instantiate, evicting the least recently used entry.
#pragma once
#include "common.cuh"

#include <cstring>

namespace ps {

constexpr int GC_WORDS = 96;
constexpr int GC_ENTRIES = 8;

struct GraphKey {
  unsigned long long w[GC_WORDS];
  int nkey = 0;     // words [0, nkey) are compared for an exact hit
  int shape0 = 0;   // words [shape0, nkey) describe the shape
  GraphKey() { memset(w, 0, sizeof(w)); }
  void ptr(const void* p) { w[nkey++] = (unsigned long long)(uintptr_t)p; }
  void begin_shape() { shape0 = nkey; }
  void val(long long v) { w[nkey++] = (unsigned long long)v; }
  bool same(const GraphKey& o) const { return nkey == o.nkey && shape0 == o.shape0 && memcmp(w, o.w, sizeof(unsigned long long) * nkey) == 0; }
  bool same_shape(const GraphKey& o) const {
    if (nkey != o.nkey || shape0 != o.shape0) return false;
    for (int i = 0; i < shape0; i++)
      if ((w[i] == 0) != (o.w[i] == 0)) return false;  // optional buffers present in both or in neither
    return memcmp(w + shape0, o.w + shape0, sizeof(unsigned long long) * (nkey - shape0)) == 0;
  }
};

struct GraphCache {
  struct Entry {
    GraphKey key;
    cudaGraphExec_t exec = nullptr;
    unsigned long long stamp = 0;
    long long kernels = 0;  // kernel launches recorded while capturing (what one replay launches)
    Workspace ws;           // scratch arena of this graph's kernels (no memory nodes: the graph stays updatable)
  };
  Entry e[GC_ENTRIES];
  unsigned long long clock = 0;
  long long hits = 0, updates = 0, instantiations = 0;

  void drop() {
    for (auto& x : e) {
      if (x.exec) cudaGraphExecDestroy(x.exec);
      x.exec = nullptr;
      x.stamp = 0;
    }
  }

  template <class F>
  static cudaGraph_t capture(Entry* into, cudaStream_t cap_stream, F&& enqueue, int* rc, long long* kernels) {
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    into->ws.used = 0;
    into->ws.needed = 0;
    tls_workspace() = &into->ws;
    const long long k0 = launch_counter();
    *rc = enqueue(cap_stream);
    *kernels = launch_counter() - k0;
    launch_counter() = k0;  // captured, not launched
    tls_workspace() = nullptr;
    const cudaError_t ee = cudaStreamEndCapture(cap_stream, &graph);
    if (*rc != PS_OK || ee != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      return nullptr;
    }
    return graph;
  }

  // Returns the entry holding the executable for `key`, capturing `enqueue(cap_stream)` when needed; nullptr when
  // capture or instantiation failed (the caller then runs eagerly).  *rc receives enqueue's own error, if any.
  template <class F>
  Entry* get(const GraphKey& key, cudaStream_t cap_stream, F&& enqueue, int* rc) {
    *rc = PS_OK;
    Entry* victim = &e[0];
    Entry* shape_hit = nullptr;
    int used = 0;
    for (auto& x : e) {
      if (x.exec && x.key.same(key)) {
        x.stamp = ++clock;
        hits++;
        return &x;
      }
      if (x.exec && x.key.same_shape(key) && (!shape_hit || x.stamp < shape_hit->stamp)) shape_hit = &x;
      if (x.stamp < victim->stamp) victim = &x;
      used += x.exec != nullptr;
    }
    // a full cache keeps one executable per shape and retargets it; below that, instantiating keeps every
    // address set replayable with no per-call work at all (rotating double / triple buffers)
    const bool try_update = shape_hit && used == GC_ENTRIES;
    Entry* into = try_update ? shape_hit : victim;
    long long kernels = 0;
    cudaGraph_t graph = capture(into, cap_stream, enqueue, rc, &kernels);
    if (graph && into->ws.needed > into->ws.bytes) {
      // the capture fell back to pool allocations (memory nodes): give the arena the room and capture again
      cudaGraphDestroy(graph);
      graph = nullptr;
      if (into->exec) cudaGraphExecDestroy(into->exec);
      into->exec = nullptr;
      into->stamp = 0;
      if (into->ws.base) cudaFree(into->ws.base);  // synchronises with the launches still using it
      into->ws.base = nullptr;
      into->ws.bytes = 0;
      const size_t want = into->ws.needed;
      if (cudaMalloc((void**)&into->ws.base, want) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
      into->ws.bytes = want;
      graph = capture(into, cap_stream, enqueue, rc, &kernels);
    }
    if (!graph) return nullptr;
    if (into->exec && try_update) {
      cudaGraphExecUpdateResultInfo info;
      if (cudaGraphExecUpdate(into->exec, graph, &info) == cudaSuccess) {
        cudaGraphDestroy(graph);
        into->key = key;
        into->stamp = ++clock;
        into->kernels = kernels;
        updates++;
        return into;
      }
      cudaGetLastError();
    }
    if (into->exec) cudaGraphExecDestroy(into->exec);
    into->exec = nullptr;
    into->stamp = 0;
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    into->exec = exec;
    into->key = key;
    into->stamp = ++clock;
    into->kernels = kernels;
    instantiations++;
    return into;
  }
};

}  // namespace ps
