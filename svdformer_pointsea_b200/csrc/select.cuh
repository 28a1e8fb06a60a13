// Warp-level ordering helpers shared by the kNN kernels (knn_select.cu, neighbors.cu, feature_knn.cu).
#pragma once
#include "common.cuh"

namespace ps {

// Result order of a kNN query.
//   PS_ORDER_SORT  ascending by (distance, index): what torch.argsort (stable radix sort on CUDA)
//                  gives for query_knn (models/model_utils.py:281-286).
//   PS_ORDER_TOPK  the order of torch.topk(k, largest=False, sorted=True) on CUDA, which
//                  query_knn_point uses (models/model_utils.py:807-810).  Values ascend; among EQUAL
//                  values the order is whatever torch's gather + unstable bitonic network produce:
//                  the k results are first written as [values < kth in index order] ++ [values ==
//                  kth in index order], then sorted by a 32-slot bitonic network that also swaps on
//                  equality in its merge passes (ATen/native/cuda/SortUtils.cuh bitonicSort /
//                  bitonicSwap, sort size 32, slots >= k invalid).  Reproduced exactly below;
//                  tests/golden/next.npz "edge3dup" pins it (1400/1400 rows with duplicated points).
// (the two values are #defined in include/pointsea_b200.h)

#ifdef __CUDACC__

// one (distance, index) pair per lane, ascending by (d, i)
__device__ __forceinline__ void warp_sort_pairs(float& d, int& i, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, d, j);
      const int oi = __shfl_xor_sync(0xffffffffu, i, j);
      const bool up = ((lane & k) == 0);
      const bool lower = ((lane & j) == 0);
      const bool other_less = (od < d) || (od == d && oi < i);
      const bool take = (lower == up) ? other_less : !other_less;
      if (take) { d = od; i = oi; }
    }
  }
}

// Lanes 0..k-1 hold the k best in ascending (d, i) order (k <= 32; other lanes are ignored).
// Afterwards lanes 0..k-1 hold the same set in torch.topk's order (see PS_ORDER_TOPK above).
__device__ __forceinline__ void warp_torch_topk_order(float& d, int& i, int lane, int k) {
  // All k distances distinct (the common case): any correct sorting network gives the ascending order we
  // already hold, so torch's result is identical and the emulation below is skipped.
  const float prev = __shfl_up_sync(0xffffffffu, d, 1);
  if (__ballot_sync(0xffffffffu, lane > 0 && lane < k && prev == d) == 0u) return;
  const float kth = __shfl_sync(0xffffffffu, d, k - 1);
  // 1. torch's gather order: (d == kth, index) ascending; invalid lanes to the end
  unsigned key_hi = lane < k ? (d == kth ? 1u : 0u) : 2u;
  int key_lo = i;
#pragma unroll
  for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      const unsigned oh = __shfl_xor_sync(0xffffffffu, key_hi, j);
      const int ol = __shfl_xor_sync(0xffffffffu, key_lo, j);
      const float od = __shfl_xor_sync(0xffffffffu, d, j);
      const bool up = ((lane & kk) == 0);
      const bool lower = ((lane & j) == 0);
      const bool other_less = (oh < key_hi) || (oh == key_hi && ol < key_lo);
      const bool take = (lower == up) ? other_less : !other_less;
      if (take) { key_hi = oh; key_lo = ol; d = od; }
    }
  }
  i = key_lo;
  bool valid = key_hi < 2u;
  // 2. torch's 32-slot bitonic network: slot = lane; pair (A = lane without bit `stride`, B = A + stride);
  //    swap = (kA < kB && validA) || !validB; exchanged when swap == dir
  auto pass = [&](int stride, bool dir) {
    const float od = __shfl_xor_sync(0xffffffffu, d, stride);
    const int oi = __shfl_xor_sync(0xffffffffu, i, stride);
    const bool ov = __shfl_xor_sync(0xffffffffu, (int)valid, stride) != 0;
    const bool is_a = (lane & stride) == 0;
    const float ka = is_a ? d : od, kb = is_a ? od : d;
    const bool va = is_a ? valid : ov, vb = is_a ? ov : valid;
    const bool swap = ((ka < kb) && va) || !vb;
    if (swap == dir) { d = od; i = oi; valid = ov; }
  };
#pragma unroll
  for (int size = 2; size < 32; size <<= 1) {
    const bool flag = (lane & size) != 0;
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) pass(stride, flag);
  }
#pragma unroll
  for (int stride = 16; stride > 0; stride >>= 1) pass(stride, false);
}

#endif  // __CUDACC__

}  // namespace ps
