// Argmax pieces shared by k_select (sipp_select.cu) and the fused tail of k_gain (sipp_gain.cu): the (value, index) order of
// [upstream] SubsequentBest::selectNextBest evaluated as a sequential scan with a strict '>' (lowest index wins ties; NaN never
// beats anything), the last-block-done workspace, and the flat-tree corner cases of that scan.
#ifndef SIPP_ARGMAX_CUH_
#define SIPP_ARGMAX_CUH_

#include <math_constants.h>

#include <climits>

#include "sipp_internal.h"

constexpr int kSelMaxBlocks = 296;

struct SelectWs {
  unsigned int done;       // last-block-done ticket
  unsigned int pad[3];
  BestRec partial[kSelMaxBlocks];
};

__device__ __forceinline__ bool better(double v, long long i, double bv, long long bi) {
  // "v beats bv": strictly greater, or equal with a lower index. NaN never beats anything.
  return (v > bv) || (v == bv && i < bi);
}

__device__ __forceinline__ void warp_argmax(double& v, long long& i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(0xFFFFFFFFu, v, off);
    const long long oi = __shfl_xor_sync(0xFFFFFFFFu, i, off);
    if (better(ov, oi, v, i)) { v = ov; i = oi; }
  }
}

// Flat scan over value[0..n): what the sequential scan returns given the order-independent argmax (bv, bi) over the
// comparable elements. A NaN first element sticks (nothing beats it with '>'), and when no element was comparable
// (bi == LLONG_MAX: all NaN / -inf) the first element is returned.
__device__ __forceinline__ BestRec finalize_flat(int n, const double* value, double bv, long long bi, long long index_offset) {
  if (n <= 0) { bv = 0.0; bi = -1; }
  else {
    const double v0 = __ldcg(value);     // possibly written by another CTA of the same launch: read it from L2
    if (v0 != v0 || bi == LLONG_MAX) { bv = v0; bi = 0; }
  }
  BestRec r;
  r.value = bv;
  r.index = bi >= 0 ? bi + index_offset : bi;
  return r;
}

#endif  // SIPP_ARGMAX_CUH_
