// Value + selection kernels (hot-path row A14).
//
// [upstream] GlobalNormalizedGain::computeValue: value(n) = max over the subtree of n of
//   (sum of gain root..m) / (sum of cost root..m), 0 where the cost sum is <= 0,
// [upstream] SubsequentBest::selectNextBest: argmax over the root's children of the subtree-max value,
// evaluated as a sequential scan with a strict '>' (so the lowest index wins ties; the reference then picks
// randomly among exact ties, which no test can pin). Call sites in the reference:
// advanced_planner_ros/cfg/planners/example_config.yaml:100-106.
//
// NaN handling mirrors the sequential scan: a NaN can only "win" when it is the very first element; NaN
// elements met later never replace the running maximum.
#include <math_constants.h>

#include <climits>

#include "sipp_argmax.cuh"
#include "sipp_internal.h"
#include "sipp_peer.cuh"

namespace {

constexpr int kSelThreads = 256;

__device__ void block_argmax(double& v, long long& i) {
  __shared__ double sv[kSelThreads / 32];
  __shared__ long long si[kSelThreads / 32];
  warp_argmax(v, i);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = v; si[warp] = i; }
  __syncthreads();
  if (warp == 0) {
    v = lane < kSelThreads / 32 ? sv[lane] : -CUDART_INF;
    i = lane < kSelThreads / 32 ? si[lane] : LLONG_MAX;
    warp_argmax(v, i);
  }
  __syncthreads();
}

// flat argmax over value[0..n): candidates are the root's children in index order.
// `first` = index of the first element of the scan (its NaN-ness decides the NaN corner case), `stride`/`sel`
// allow scanning only the root's children in tree mode (sel[i] != 0).
__global__ void __launch_bounds__(kSelThreads)
k_select(int n, const double* __restrict__ value, const int32_t* __restrict__ parent, int tree_mode, long long index_offset,
         SelectWs* ws, BestRec* out, double* root_value, const int* overflow_flag, const PeerDev peer, int peer_mode) {
  PeerCounters pcnt = {0, 0};
  if (peer_mode >= 0 && threadIdx.x == 0) pcnt = peer_preload(peer);
  double bv = -CUDART_INF;
  long long bi = LLONG_MAX;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (tree_mode && (i == 0 || parent[i] != 0)) continue;   // only the root's children take part
    const double v = value[i];
    if (better(v, i, bv, bi)) { bv = v; bi = i; }
  }
  block_argmax(bv, bi);
  __shared__ bool is_last;
  __shared__ BestRec s_mine;
  if (threadIdx.x == 0) {
    ws->partial[blockIdx.x].value = bv;
    ws->partial[blockIdx.x].index = bi;
    __threadfence();
    const unsigned int ticket = atomicAdd(&ws->done, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  bv = -CUDART_INF;
  bi = LLONG_MAX;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
    const double v = ws->partial[b].value;
    const long long i = ws->partial[b].index;
    if (better(v, i, bv, bi)) { bv = v; bi = i; }
  }
  block_argmax(bv, bi);
  if (threadIdx.x == 0) {
    // flat tree: value(root) = findBest(root) = max(0, children) with NaN children ignored by std::max [upstream]
    if (root_value) *root_value = (bv > 0.0) ? bv : 0.0;
    // sequential-scan corner cases: find the first participating element
    long long first = -1;
    if (!tree_mode) first = n > 0 ? 0 : -1;
    else
      for (int i = 1; i < n; ++i)
        if (parent[i] == 0) { first = i; break; }
    if (first < 0) { bv = 0.0; bi = -1; }
    else if (value[first] != value[first]) { bv = value[first]; bi = first; }   // NaN first element sticks
    else if (bi == LLONG_MAX) { bv = value[first]; bi = first; }                // every other element was NaN / -inf ties
    if (overflow_flag && *overflow_flag) { bv = 0.0; bi = -2; }   // tree deeper than kMaxTreeDepth: in-band error
    s_mine.value = bv;
    s_mine.index = bi >= 0 ? bi + index_offset : bi;
    if (peer_mode < 0) *out = s_mine;
    ws->done = 0;   // re-arm for the next launch on this stream
  }
  if (peer_mode >= 0) {       // multi-GPU: publish the shard winner into every peer's mailbox, wait for theirs, merge (sipp_peer.cuh)
    __syncthreads();
    peer_exchange_block(peer, s_mine, out, peer_mode, pcnt);
  }
}

// ---- tree mode ----------------------------------------------------------------------------------------
// GlobalNormalizedGain [upstream]: computeValue(i) first sums gain and cost over i's ancestors, nearest first
// (parent, grandparent, ..., root), then findBest descends from i adding each node on the way down; the value of i is
// the maximum over its subtree of G/C with exactly that association. The normalised gain of a descendant m therefore
// depends (in the last ulps) on which ancestor i it is evaluated for, so every (ancestor, descendant) pair is summed in
// the reference's order: O(depth^2) adds per node, one thread per node m, results merged with atomicMax on an
// order-preserving integer key. std::max(a, b) = (a < b) ? b : a ignores a NaN b and keeps a NaN a: a NaN own value
// makes value(i) NaN, NaN descendants are skipped.
constexpr int kMaxTreeDepth = 512;

__device__ __forceinline__ unsigned long long dkey(double v) {   // monotone map double -> uint64 (no NaN)
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void k_tree_init(int n, unsigned long long* __restrict__ best, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) best[i] = 0ull;   // below every key
  if (i == 0) { flags[0] = 0; }
}

// gain_below (optional): gains to use for the nodes strictly BELOW the evaluated node i. The reference re-evaluates the tree
// in pre-order (OnlinePlanner::updateEvaluatorStep [upstream]): when computeValue(i) runs, i and its ancestors already carry
// this cycle's gains while i's descendants still carry last cycle's. gain_below == nullptr: one consistent gain array.
// root_live: the root's gain / cost are used as given (an updater pass that re-scored the root); otherwise they count as 0 (the
// planner zeroes the segment that becomes the root [upstream OnlinePlanner::requestNextTrajectory]).
// cost_below (optional): likewise for the costs — with update_cost the pass also recomputes the cost of every visited node.
__global__ void k_tree_pairs(int n, const double* __restrict__ gain, const double* __restrict__ gain_below, const double* __restrict__ cost,
                             const double* __restrict__ cost_below, const int32_t* __restrict__ parent, unsigned long long* __restrict__ best,
                             uint8_t* __restrict__ own_nan, int* __restrict__ flags, int root_live) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  int chain[kMaxTreeDepth];   // chain[0] = m, chain[depth-1] = root
  int depth = 0;
  for (int cur = m; cur >= 0; cur = (cur == 0) ? -1 : parent[cur]) {
    if (depth == kMaxTreeDepth) { atomicExch(&flags[0], 1); return; }
    chain[depth++] = cur;
  }
  own_nan[m] = 0;
  for (int a = 0; a < depth; ++a) {
    double g = 0.0, c = 0.0;
    for (int b = a + 1; b < depth; ++b) {   // computeValue: ancestors of i = chain[a], nearest first
      const int k = chain[b];
      g += (k == 0 && !root_live) ? 0.0 : gain[k];        // the root (current segment) has gain = cost = 0 [upstream OnlinePlanner]
      c += (k == 0 && !root_live) ? 0.0 : cost[k];
    }
    for (int b = a; b >= 0; --b) {          // findBest: i itself, then down to m
      const int k = chain[b];
      g += (k == 0 && !root_live) ? 0.0 : ((b == a || !gain_below) ? gain[k] : gain_below[k]);
      c += (k == 0 && !root_live) ? 0.0 : ((b == a || !cost_below) ? cost[k] : cost_below[k]);
    }
    const double v = c > 0 ? g / c : 0.0;
    if (v != v) {
      if (a == 0) own_nan[m] = 1;
      continue;
    }
    atomicMax(&best[chain[a]], dkey(v));
  }
}

__global__ void k_tree_final(int n, const unsigned long long* __restrict__ best, const uint8_t* __restrict__ own_nan,
                             double* __restrict__ value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  value[i] = own_nan[i] ? __longlong_as_double(0x7FF8000000000000ll) : dunkey(best[i]);
}

}  // namespace

// merge of the gathered per-shard records (multi-GPU): one warp, sequential-scan semantics (see sipp.h)
__global__ void k_merge_best(const BestRec* __restrict__ recs, int world, long long stride, BestRec* __restrict__ out) {
  if (threadIdx.x != 0) return;
  BestRec best = recs[0];
  if (best.index >= 0) best.index += 0;      // rank 0: offset 0
  bool have = best.index >= 0 && !(best.value != best.value);
  for (int r = 1; r < world; ++r) {
    BestRec c = recs[r];
    if (c.index < 0 || c.value != c.value) continue;
    c.index += (long long)r * stride;
    if (!have || c.value > best.value || (c.value == best.value && c.index < best.index)) { best = c; have = true; }
  }
  *out = best;
}

cudaError_t launch_merge_best(sipp_ctx* ctx, const BestRec* d_recs, int world, long long stride, BestRec* d_out, cudaStream_t s) {
  k_merge_best<<<1, 32, 0, s>>>(d_recs, world, stride, d_out);
  ctx->launches += 1;
  return cudaGetLastError();
}

size_t select_workspace_bytes() { return sizeof(SelectWs); }

static int sel_grid(int n) {
  int g = (n + kSelThreads * 4 - 1) / (kSelThreads * 4);
  if (g < 1) g = 1;
  if (g > kSelMaxBlocks) g = kSelMaxBlocks;
  return g;
}

cudaError_t launch_select_flat(sipp_ctx* ctx, int n, const double* d_value, long long index_offset, BestRec* d_best, double* d_root_value,
                               cudaStream_t s, const PeerDev* d_peer, int peer_mode) {
  static const PeerDev no_peer = {};
  k_select<<<sel_grid(n), kSelThreads, 0, s>>>(n, d_value, nullptr, 0, index_offset, reinterpret_cast<SelectWs*>(ctx->d_select_ws), d_best,
                                               d_root_value, nullptr, d_peer ? *d_peer : no_peer, d_peer ? peer_mode : -1);
  ctx->launches += 1;
  return cudaGetLastError();
}

// flat tree: every node i >= 1 is a child of the root (node 0, gain = cost = 0)
__global__ void k_value_flat(int n, const double* __restrict__ gain, const double* __restrict__ cost, double* __restrict__ value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0) { value[0] = 0.0; return; }
  const double g = 0.0 + gain[i], c = 0.0 + cost[i];
  value[i] = c > 0 ? g / c : 0.0;
}

cudaError_t launch_value_flat(sipp_ctx* ctx, int n, const double* d_gain, const double* d_cost, double* d_value, cudaStream_t s) {
  k_value_flat<<<(n + 255) / 256, 256, 0, s>>>(n, d_gain, d_cost, d_value);
  ctx->launches += 1;
  return cudaGetLastError();
}

// ws: best[n] u64 + own_nan[n] bytes + flags (int) -> 9 n + 64 bytes
cudaError_t launch_value_tree(sipp_ctx* ctx, int n, const double* d_gain, const double* d_gain_below, const double* d_cost,
                              const int32_t* d_parent, double* d_value, BestRec* d_best, void* ws, cudaStream_t s, int root_live,
                              const double* d_cost_below) {
  unsigned long long* best = reinterpret_cast<unsigned long long*>(ws);
  uint8_t* own_nan = reinterpret_cast<uint8_t*>(best + n);
  int* flags = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + (((size_t)n * 9 + 15) & ~(size_t)15));
  const int threads = 128, blocks = (n + threads - 1) / threads;
  k_tree_init<<<blocks, threads, 0, s>>>(n, best, flags);
  k_tree_pairs<<<blocks, threads, 0, s>>>(n, d_gain, d_gain_below, d_cost, d_cost_below, d_parent, best, own_nan, flags, root_live);
  k_tree_final<<<blocks, threads, 0, s>>>(n, best, own_nan, d_value);
  k_select<<<sel_grid(n), kSelThreads, 0, s>>>(n, d_value, d_parent, 1, 0, reinterpret_cast<SelectWs*>(ctx->d_select_ws), d_best, nullptr, flags, PeerDev{}, -1);
  ctx->launches += 4;
  return cudaGetLastError();
}
