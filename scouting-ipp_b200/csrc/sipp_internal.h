// Internal (non-ABI) declarations shared by the translation units of libsipp.so.
#ifndef SIPP_INTERNAL_H_
#define SIPP_INTERNAL_H_

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <list>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sipp.h"

// ---- packed per-cell gain record (one byte per cell, plane d_rec) ------------------------------------
// Built by k_build_views from state / path / reachability and the evaluator's bounding volume; off-map
// cells are never stored — TMA out-of-bounds fill returns 0 for them, which contributes to no count.
#define REC_UNK 0x01u        // state == UNKNOWN                       (and voxel inside the bounding volume)
#define REC_UNK_PATH 0x02u   // UNKNOWN and on the current path mask   (")
#define REC_UNK_REACH 0x04u  // UNKNOWN and a 3x3 neighbour is MSR_REACHABLE (")
#define REC_OBS_FREE 0x08u   // state == FREE                          (")
// raw bits (no bounding-volume masking): the duplicated centre voxel has its own coordinates, hence its own
// bounding-volume / map-bounds tests (done by lane 0 of the gain kernel)
#define REC_RAW_UNK 0x10u    // state == UNKNOWN
#define REC_PATH_RAW 0x20u   // path mask bit
#define REC_REACH_RAW 0x40u  // a 3x3 neighbour is MSR_REACHABLE
#define REC_RAW_FREE 0x80u   // state == FREE

// planner cell codes (plane d_pcode, uint16): the follower planner's own copy of the map
// (PRMStarPlanner::map_ / explored_, prm_star_planner.h:52-53)
#define PCODE_UNEXPLORED 0xFFFFu  // cost = init_cost
#define PCODE_REMOVED 0xFFFEu     // first observation was an obstacle id: all edges deleted

constexpr int kRecipEntries = 32768;   // ExpandLowCost view: idx = closest-observed label + 1 in 15 bits (labels <= 32766)
#define SIPP_MAX_CLASSES 64  // distinct |dX|^2 (or |dY|^2) values of the node lattice (9..14 in practice)

struct FieldGeom {  // device-side view of the implicit per-pixel graph geometry
  int W, H;
  int Kx, Ky;              // number of edge-length classes per axis
  const uint8_t* xcls;     // [W-1] class of the edge between column i and i+1
  const uint8_t* ycls;     // [H-1]
  const double* hdist;     // [Kx] axial-x edge length   sqrt(dX^2 + 0)      (+inf when not < max_dist)
  const double* vdist;     // [Ky] axial-y edge length
  const double* ddist;     // [Kx*Ky] diagonal edge length sqrt(dX^2 + dY^2)  (+inf when not < max_dist)
  double init_half_cost;   // init_cost / 2
  double spacing, off_x, off_y, min_cost;
  int start_x, start_y, goal_x, goal_y;
};

struct GainGeom {  // per-candidate index math of MapPlanExp (map_planexp.cpp:72-101,134-144,222-224)
  int W, H;
  double vs;          // c_voxel_size_
  double inv_vs;      // 1 / c_voxel_size_ (the reference multiplies by (1/vs))
  double map_width, map_height;  // map server bounds test
  double goal_x, goal_y;
  double mean_cost;
  // bounding volume as inclusive cell-index ranges for the corner voxels (k * vs inside [min,max])
  int bv_kx_lo, bv_kx_hi, bv_ky_lo, bv_ky_hi;
};

struct BestRec {  // 16-byte record moved by the multi-GPU gather
  double value;
  long long index;
};

// ---- multi-GPU peer exchange (protocol: sipp_peer.cuh) -------------------------------------------------------------------
constexpr int kPeerSlots = 4;

// One record = three self-validating 8-byte words, each carrying a tag derived from the step number, so neither side needs a
// system-scope fence (measured: the release / acquire pair cost ~4 us per step, more than the NVLink hop itself): an aligned
// 8-byte store is atomic, the words may arrive in any order, the reader polls until all three carry the tag of the step.
struct PeerSlot {
  unsigned long long w0;        // tag32 << 32 | low  half of the value's bit pattern
  unsigned long long w1;        // tag32 << 32 | high half
  unsigned long long w2;        // tag16 << 48 | global candidate index as a 48-bit two's complement (< 0: the shard was empty)
  unsigned long long pad;
};
struct PeerBox {
  PeerSlot slots[kPeerSlots][SIPP_PEER_MAX_WORLD];   // [step % 4][writer rank]
  // device resident so a captured graph replays unchanged; only this rank's own kernels (one stream) touch them
  unsigned long long published;  // records this rank has published
  unsigned long long collected;  // steps this rank has collected (merged)
  unsigned long long status;     // != 0: a wait timed out (a peer died or the ranks are not calling in lockstep)
};
struct PeerDev {
  PeerBox* box[SIPP_PEER_MAX_WORLD];   // box[r]: rank r's mailbox as mapped into this rank's address space
  int rank, world;
  unsigned long long timeout_ns;
};

struct sipp_ctx {
  sipp_map_desc desc;
  int device;
  cudaStream_t stream;
  cudaStream_t s_in, s_out;          // copy streams of the chunked host-pointer pipeline (sipp_cycle)
  cudaEvent_t ev_in[8], ev_k[8];
  cudaEvent_t ev_plan[4];            // device timestamps of the last plan: relaxation start / end, plan start / end (sipp_plan_phases)
  int cycle_chunks;                  // 0 = automatic; tuning knob SIPP_CYCLE_CHUNKS
  // sipp_cycle's pipeline captured as a CUDA graph, replayed while the call's pointers / parameters stay the same
  bool cycle_use_graph;
  bool cycle_zero_copy;              // results go straight into the caller's pinned host arrays (opt-in: SIPP_CYCLE_ZEROCOPY=1)
  cudaGraphExec_t cycle_graph;
  std::vector<unsigned char> cycle_key;
  int cycle_graph_launches;
  std::mutex mu;
  std::string err;
  int64_t launches;
  int gain_warps, gain_stages;   // k_gain pipeline shape (0 = default); tuning knobs SIPP_GAIN_WARPS / SIPP_GAIN_STAGES

  int W, H;
  size_t pitch;        // bytes per row of the uint8 planes (multiple of 128)
  size_t pitch16;      // elements per row of the uint16 planes
  // host mirrors (bookkeeping that is sequential in the reference)
  std::vector<uint8_t> h_state;   // map_state_
  double mean_cost, max_cost;
  int mean_counter;
  double init_cost;               // current planner init cost (update_unknown_cost changes it)
  int start_loc[2], goal_loc[2];
  double mpp, vs;
  // planner geometry (host copies)
  double spacing, max_dist, off_x, off_y;
  int start_node[2], goal_node[2];
  std::vector<double> X, Y;
  // device planes
  uint8_t* d_state;
  uint16_t* d_label;     // map_cost_ (last label)
  uint16_t* d_pcode;     // planner code
  uint8_t* d_path;       // path mask 0/1
  uint8_t* d_rec;        // packed gain record (view)
  uint8_t* d_rec2[2];    // 2 bits per cell (4 cells per byte, row pitch = pitch / 4): bit 0 UNKNOWN, bit 1 UNKNOWN & path [0] / & reachable [1]
                         // (both inside the bounding volume) — what the count-type evaluators scan
  uint16_t* d_freelab;   // view: label where state == FREE (inside bv) else 0   (AverageViewCost)
  uint16_t* d_closest;   // ExpandLowCost view: closest-observed cost index + counted flag per cell (k_build_lowcost_view)
  uint16_t* d_cclab;     // map_cost_closest_ (label of the closest observed cell), only with compute_closest_observed
  float* d_closest_dist; // map_distance_closest_ as squared pixel distance
  bool lowcost_view_dirty;
  double recip_mean; int recip_augment; bool recip_valid;   // what d_recip was filled for
  double* d_recip;       // [kRecipEntries] 1/label lookup for ExpandLowCost, indexed by closest cost + 1
  double* d_field;       // W*H doubles
  double* d_tile_w;      // [ntiles][4][34][34] edge-weight cache of the relaxation tiles (k_weights)
  uint8_t* d_pred;       // canonical predecessor direction per cell (k_pred), [H][pitch]
  uint16_t* d_truth;     // the simulator's ground-truth labels (sipp_truth_upload), [H][pitch16]; nullptr until uploaded
  // geometry tables on device
  uint8_t *d_xcls, *d_ycls;
  double *d_hdist, *d_vdist, *d_ddist;
  int Kx, Ky;
  // relax work lists
  int n_tiles_x, n_tiles_y;
  int* d_tile_flags;       // [ntiles] tile is queued
  int* d_tile_list;        // [ntiles] ring queue of tile ids
  int* d_relax_ctl;        // small control block (counters, stats)
  // path
  int* d_path_nodes;       // reverse order (goal -> start), packed (x<<16 | y)... see sipp_field.cu
  double* d_path_xy;
  int path_cap_nodes;
  int* d_plan_out;         // {len, all_explored, status...}
  int path_generation;
  bool path_valid;
  bool field_valid;
  int64_t plan_stats[4];
  // incremental re-plan (sipp_field.cu): the converged field + predecessor plane of the last plan stay valid as a warm start while
  // the only changes since are ingested patches
  bool inc_enabled;            // SIPP_INCREMENTAL=0 turns it off
  bool path_mask_foreign;      // d_path was written by sipp_set_path_mask since the last plan: the next plan must rasterise again
  bool warm_valid;             // d_field / d_pred hold the fixed point of the planner map as it was at the last plan
  double warm_init_cost;       // init cost the last plan used
  std::vector<uint8_t> h_seed_stage;  // staging of the seed map + list of an incremental plan (one copy)
  std::vector<int> pending_rects;   // x0, y0, x1, y1 (cell ranges, inclusive) of the patches ingested since the last plan
  uint8_t* d_mark;             // [H][pitch] cells whose cost-to-go lost its support: 1 = first observed at a higher cost (k_ingest_patch), 2 = published
  uint8_t* d_tile_seed;        // [ntiles] tile holds invalidated cells / was touched by a patch
  int* d_seed_list;            // [ntiles] tile ids of the patches since the last plan
  int64_t plan_info[4];        // last plan: [0] 1 = incremental, [1] tiles seeded from patches, [2] tiles seeded for the relaxation, [3] cells invalidated
  // views
  bool views_dirty;
  sipp_eval_params view_params;   // bounding volume the views were built for
  bool view_params_valid;
  // TMA descriptors cached per half_fov
  struct TmaEntry { int half_fov; int kind; CUtensorMap map; };
  std::list<TmaEntry> tma_cache;   // list: handed-out pointers stay valid
  // staging buffers for the host-pointer entry points
  void* h_pinned; size_t h_pinned_bytes;
  void* d_stage; size_t d_stage_bytes;
  void* d_tree_ws; size_t d_tree_ws_bytes;
  BestRec* d_best; void* d_select_ws;
  double* d_plan_cost_dummy;
  // multi-GPU peer exchange (sipp_peer.cu): own mailbox, the peers' mailboxes as mapped here, device copy of the table
  PeerBox* d_peer_box;
  PeerDev peer;                             // the mailbox table, passed to the kernels by value
  void* peer_mapped[SIPP_PEER_MAX_WORLD];   // cudaIpcOpenMemHandle results to close again (nullptr: not IPC-mapped)
  int peer_rank, peer_world;                // peer_world == 0: not attached
};

// error helpers -----------------------------------------------------------------------------------------
int sipp_set_err(sipp_ctx* ctx, int code, const char* fmt, ...);
#define SIPP_CUDA_CHECK(ctx, expr)                                                                   \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return sipp_set_err((ctx), SIPP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// SIPP_DEBUG_SYNC=1 in the environment: synchronise after every kernel so an asynchronous fault is attributed to its kernel
int sipp_debug_sync(sipp_ctx* ctx, cudaStream_t s, const char* what);

// kernel launchers (each returns cudaError_t from the launch) ---------------------------------------------
// sipp_map.cu
cudaError_t launch_clear_planes(sipp_ctx* ctx, cudaStream_t s);
cudaError_t launch_ingest_patch(sipp_ctx* ctx, int x0, int y0, int rows, int cols, const int32_t* d_labels,
                                const uint8_t* d_observed, cudaStream_t s);
cudaError_t launch_unknown_cost_noop(sipp_ctx* ctx, cudaStream_t s);
cudaError_t launch_build_views(sipp_ctx* ctx, const GainGeom& g, bool use_bv, cudaStream_t s);
cudaError_t launch_closest_update(sipp_ctx* ctx, int x0, int y0, int rows, int cols, cudaStream_t s);
cudaError_t launch_closest_clear(sipp_ctx* ctx, cudaStream_t s);
cudaError_t launch_build_lowcost_view(sipp_ctx* ctx, const GainGeom& g, bool use_bv, cudaStream_t s);
cudaError_t launch_download_planes(sipp_ctx* ctx, uint8_t* d_state_out, int32_t* d_cost_out, uint8_t* d_reach_out,
                                   uint8_t* d_path_out, cudaStream_t s);
cudaError_t launch_upload_mask(sipp_ctx* ctx, const uint8_t* d_mask_packed, cudaStream_t s);
// sipp_gain.cu
struct FusedSelect {       // argmax over value fused into the tail of k_gain (flat tree): one launch per evaluation cycle
  BestRec* best_out;
  long long index_offset;
  const PeerDev* peer;     // non-null (host pointer, copied into the kernel parameters): exchange the winner with the peers, write the GLOBAL one
  int peer_mode;           // PEER_SYNC / PEER_PIPELINED (sipp_peer.cuh)
};
int gain_launch(sipp_ctx* ctx, const sipp_eval_params* p, const GainGeom& g, int n, const double* d_xy,
                const double* d_start_xy, const double* d_cost, double* d_gain, uint32_t* d_counts,
                uint8_t* d_terminal, double* d_value, cudaStream_t s, const int* d_sel = nullptr, const int* d_n = nullptr,
                double* d_value2 = nullptr, const FusedSelect* fs = nullptr, double* d_fsum = nullptr);
int gain_launch_views(sipp_ctx* ctx, const sipp_eval_params* p, const GainGeom& g, int n, const int* d_view_off, int n_views, const double* d_view_xy,
                      const double* d_start_xy, double* d_vgain, uint32_t* d_vcounts, double* d_vfsum, double* d_gain, uint32_t* d_counts,
                      uint8_t* d_terminal, cudaStream_t s);
// sipp_update.cu: evaluator-updater policy on the device (row A13)
cudaError_t launch_update_select(sipp_ctx* ctx, const sipp_updater_params& u, int n, const double* d_end_xy, const double* d_gain,
                                 const uint8_t* d_terminal, double cur_x, double cur_y, int path_changed, uint8_t* d_do_gain, uint8_t* d_follow,
                                 int* d_sel, int* d_n_sel, cudaStream_t s);
cudaError_t launch_update_apply(sipp_ctx* ctx, const sipp_updater_params& u, int n, int sets_terminal, const uint8_t* d_do_gain, const uint8_t* d_follow,
                                const double* d_gain_old, const double* d_gain_tmp, const uint8_t* d_term_tmp, double* d_gain_new,
                                uint8_t* d_terminal, uint8_t* d_do_value, cudaStream_t s);
cudaError_t launch_update_value(sipp_ctx* ctx, int n, const uint8_t* d_do_value, const double* d_value_tmp, double* d_value, cudaStream_t s);
// sipp_select.cu
cudaError_t launch_select_flat(sipp_ctx* ctx, int n, const double* d_value, long long index_offset, BestRec* d_best, double* d_root_value,
                               cudaStream_t s, const PeerDev* d_peer = nullptr, int peer_mode = 0);
// sipp_peer.cu: standalone exchange of one record (the fused form lives in k_select)
cudaError_t launch_peer_exchange(sipp_ctx* ctx, const BestRec* d_mine, BestRec* d_out, int mode, cudaStream_t s);
void peer_release(sipp_ctx* ctx);
cudaError_t launch_value_flat(sipp_ctx* ctx, int n, const double* d_gain, const double* d_cost, double* d_value, cudaStream_t s);
cudaError_t launch_value_tree(sipp_ctx* ctx, int n, const double* d_gain, const double* d_gain_below, const double* d_cost,
                              const int32_t* d_parent, double* d_value, BestRec* d_best, void* ws, cudaStream_t s, int root_live = 0,
                              const double* d_cost_below = nullptr);
cudaError_t launch_merge_best(sipp_ctx* ctx, const BestRec* d_recs, int world, long long stride, BestRec* d_out, cudaStream_t s);
size_t select_workspace_bytes();
// sipp_cost.cu
cudaError_t launch_cost_integral(sipp_ctx* ctx, const sipp_cost_params& p, int n, const int32_t* d_off, const double* d_pts, const double* d_parent_cost,
                                 double* d_cost, uint8_t* d_valid, cudaStream_t s);
cudaError_t launch_travel_cost(sipp_ctx* ctx, const sipp_travel_params& p, int n, const double* d_start, const double* d_goal, double* d_cost, cudaStream_t s);
// sipp_field.cu
int field_plan(sipp_ctx* ctx, cudaStream_t s);

// ---- batched forms (sipp_batch_*: an ensemble of maps advanced by one launch per kernel, gridDim.y = maps) ------------------------
size_t field_plan_batch_args_bytes();
int field_plan_batch(sipp_ctx* const* ctxs, int m, void* h_args, void* d_args, uint8_t* h_seed, uint8_t* d_seed, size_t seed_stride, int* d_out,
                     cudaStream_t s, int* launches, const std::function<void(int, const std::function<void(int)>&)>* runner,
                     int* roam_queue, int roam_cap, int* roam_ctl, int roam_share, int roam_kind);
// the shared relaxation queue of a batch: capacity in ints for m maps of ntiles tiles (0: not representable, use per-map workers); reset
// (all slots empty, counters zero) before the first plan and after a watchdog abort
size_t field_roam_queue_ints(int ntiles, int m);
int field_roam_reset(sipp_ctx* c0, int* roam_queue, int roam_cap, int* roam_ctl, cudaStream_t s);
size_t ingest_args_bytes();
size_t build_views_args_bytes();
void fill_ingest_args(sipp_ctx* ctx, int x0, int y0, int rows, int cols, const int32_t* d_labels, void* dst);
void fill_build_views_args(sipp_ctx* ctx, const GainGeom& g, bool use_bv, int on, void* dst);
cudaError_t launch_ingest_patch_batch(const void* d_args, int m, int rows, int cols, cudaStream_t s);
cudaError_t launch_build_views_batch(const void* d_args, int m, size_t pitch, int H, cudaStream_t s);
size_t gain_args_bytes();
int gain_fill_batch_args(sipp_ctx* ctx, const sipp_eval_params* p, const GainGeom& g, int n, const double* d_xy, const double* d_cost, double* d_gain,
                         uint8_t* d_terminal, double* d_value, const FusedSelect* fs, void* dst, int* kind_out, int* smem_out, int* wpc_out);
int gain_launch_batch(sipp_ctx* c0, const void* d_args, int m, int n, int kind, int smem, int wpc, cudaStream_t s);

#endif  // SIPP_INTERNAL_H_
