// Evaluator-updater policy on the device (hot-path row A13): which tree nodes are re-scored in a planning cycle.
//
// The reference walks the tree in pre-order ([upstream] OnlinePlanner::updateEvaluatorStep) and calls, per node, the updater
// chain   RRTStarUpdater / [upstream] ConstrainedUpdater  ->  [upstream] UpdateAll / UpdateAllNonTerminal
// (advanced_planner_modules/src/evaluator_updater/rrt_star_updater.cpp:19-42, update_all_non_terminal.cpp:16-27). Every
// decision of that chain depends only on the node's OWN gain / terminal flag at the time it is visited (nothing a previously
// visited node wrote), so the predicate is evaluated for all nodes at once; the eligible nodes are compacted into an index
// list that k_gain consumes directly (candidate count read on the device: no host round trip), and the values are then
// computed with the pre-order semantics by k_tree_pairs (new gains on the chain root..i, old gains below i).
#include "sipp_internal.h"

namespace {

__global__ void k_update_select(sipp_updater_params u, int n, const double* __restrict__ end_xy, const double* __restrict__ gain,
                                const uint8_t* __restrict__ terminal, double cur_x, double cur_y, int path_changed,
                                uint8_t* __restrict__ do_gain, uint8_t* __restrict__ follow_out, int* __restrict__ sel, int* __restrict__ n_sel) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool follow = false, dg = false;
  // Node 0 is the root: the patched planner visits it first (trajectory_evaluator_->updateSegment(current_segment),
  // mav_active_3d_planning.patch:56). RRTStarUpdater only compares path ids there ("cannot update the root segment",
  // rrt_star_updater.cpp:21,30-38); a ConstrainedUpdater / a following updater used directly treat it like any other node.
  if (i < n && !(i == 0 && u.updater == SIPP_UPD_RRT_STAR)) {
    const double dx = cur_x - end_xy[2 * i], dy = cur_y - end_xy[2 * i + 1];
    // (current_position - trajectory.back().position_W).norm() with z = 0                                        :25
    const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), 0.0));
    const bool term = terminal[i] != 0;
    const double g = gain[i];
    if (u.updater == SIPP_UPD_RRT_STAR) {
      if (!term) follow = (g > u.minimum_gain || path_changed) && (dist <= u.update_range || path_changed);     // :22-28
    } else if (u.updater == SIPP_UPD_CONSTRAINED) {
      follow = g > u.minimum_gain && dist <= u.update_range;                                                     // [upstream]
    } else {
      follow = true;                                                                                            // following updater used directly
    }
    if (follow) dg = (u.following == SIPP_FOLLOW_UPDATE_ALL) ? (u.update_gain != 0) : (u.update_gain != 0 && !term);   // update_all_non_terminal.cpp:17-19
  }
  if (i < n) { do_gain[i] = dg ? 1 : 0; follow_out[i] = follow ? 1 : 0; }
  // compaction: one atomic per warp, the order of the list is irrelevant (results are scattered back by node index)
  const unsigned m = __ballot_sync(0xFFFFFFFFu, dg);
  if (m) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(n_sel, __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
    if (dg) sel[base + __popc(m & ((1u << lane) - 1u))] = i;
  }
}

__global__ void k_update_apply(sipp_updater_params u, int n, int sets_terminal, const uint8_t* __restrict__ do_gain, const uint8_t* __restrict__ follow,
                               const double* __restrict__ gain_old, const double* __restrict__ gain_tmp, const uint8_t* __restrict__ term_tmp,
                               double* __restrict__ gain_new, uint8_t* __restrict__ terminal, uint8_t* __restrict__ do_value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double g = gain_old[i];
  bool term = terminal[i] != 0;
  if (do_gain[i]) {
    g = gain_tmp[i];
    if (sets_terminal && term_tmp[i]) { term = true; terminal[i] = 1; }     // the evaluators only ever SET gain_terminal
  }
  gain_new[i] = g;
  bool dv = follow[i] != 0 && u.update_value != 0;
  if (dv && u.following == SIPP_FOLLOW_UPDATE_ALL_NON_TERMINAL && term && !u.update_cost) dv = false;   // update_all_non_terminal.cpp:23
  do_value[i] = dv ? 1 : 0;
}

__global__ void k_update_value(int n, const uint8_t* __restrict__ do_value, const double* __restrict__ value_tmp, double* __restrict__ value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && do_value[i]) value[i] = value_tmp[i];
}

}  // namespace

cudaError_t launch_update_select(sipp_ctx* ctx, const sipp_updater_params& u, int n, const double* d_end_xy, const double* d_gain,
                                 const uint8_t* d_terminal, double cur_x, double cur_y, int path_changed, uint8_t* d_do_gain, uint8_t* d_follow,
                                 int* d_sel, int* d_n_sel, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(d_n_sel, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  k_update_select<<<(n + 255) / 256, 256, 0, s>>>(u, n, d_end_xy, d_gain, d_terminal, cur_x, cur_y, path_changed, d_do_gain, d_follow, d_sel, d_n_sel);
  ctx->launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_update_apply(sipp_ctx* ctx, const sipp_updater_params& u, int n, int sets_terminal, const uint8_t* d_do_gain, const uint8_t* d_follow,
                                const double* d_gain_old, const double* d_gain_tmp, const uint8_t* d_term_tmp, double* d_gain_new,
                                uint8_t* d_terminal, uint8_t* d_do_value, cudaStream_t s) {
  k_update_apply<<<(n + 255) / 256, 256, 0, s>>>(u, n, sets_terminal, d_do_gain, d_follow, d_gain_old, d_gain_tmp, d_term_tmp, d_gain_new, d_terminal, d_do_value);
  ctx->launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_update_value(sipp_ctx* ctx, int n, const uint8_t* d_do_value, const double* d_value_tmp, double* d_value, cudaStream_t s) {
  k_update_value<<<(n + 255) / 256, 256, 0, s>>>(n, d_do_value, d_value_tmp, d_value);
  ctx->launches += 1;
  return cudaGetLastError();
}
