/*
 * sipp.h — C ABI of the B200-native scouting-ipp evaluation hot path (libsipp.so).
 *
 * This is the drop-in boundary. Every entry point replaces a C++ interface of the reference
 * (ethz-asl/scouting-ipp; citations are <pkg>/<path>:<line> relative to the reference root).
 * Plain pointers and sizes only; no torch / Eigen / ROS types. All entry points return an int status
 * (SIPP_OK == 0, negative = error) and never throw; sipp_last_error() gives the message.
 *
 * There is NO CPU fallback behind this ABI: every compute entry point launches hand-written sm_100a CUDA
 * kernels and fails with SIPP_ERR_CUDA when no usable device/driver is present.
 *
 * Conventions
 *   - maps are W x H cells (W = horizontal_units_, H = vertical_units_), host arrays are ROW-MAJOR
 *     [y * W + x] (the MapData wire layout, advanced_planner_msgs/msg/MapData.msg:1-3 and
 *     map_server_planexp/src/map_server_planexp.cpp:306-317); the reference's Eigen arrays are
 *     column-major (row=y, col=x) — only the host adapters care.
 *   - positions are metres in the map frame, doubles, like Eigen::Vector3d position_W.
 *   - "_dev" variants take DEVICE pointers and a cudaStream_t (passed as void*) and do not synchronise.
 */
#ifndef SIPP_H_
#define SIPP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIPP_ABI_VERSION 3

/* ---- status codes ---------------------------------------------------------------------------------- */
#define SIPP_OK 0
#define SIPP_ERR_INVALID_ARG (-1)
#define SIPP_ERR_CUDA (-2)            /* any CUDA failure incl. "no device": there is no CPU fallback   */
#define SIPP_ERR_UNSUPPORTED (-3)     /* geometry / parameter combination the kernels do not cover      */
#define SIPP_ERR_NO_PATH_MASK (-4)    /* path-dependent evaluator used before a plan / mask upload      */
#define SIPP_ERR_ALLOC (-5)

/* ---- cell state codes: map_server_planexp/include/map_server_planexp/map_server_planexp.h:37-39 ---- */
#define SIPP_STATE_OCCUPIED 0
#define SIPP_STATE_FREE 1
#define SIPP_STATE_UNKNOWN 2
/* reachability codes: map_server_planexp.h:45-46 */
#define SIPP_REACHABILITY_UNKNOWN 0
#define SIPP_REACHABLE 1

/* ---- planner result: planexp_base_planners/include/planexp_base_planners/base_planner.h:39-44 ------ */
#define SIPP_PLAN_VALID 0
#define SIPP_PLAN_INVALID 1
#define SIPP_PLAN_TERMINAL_VALID 2
#define SIPP_PLAN_TERMINAL_INVALID 3

#define SIPP_MAX_IDS 32

/*
 * Map + follower-planner description. Mirrors what MapServerPlanExp parses from the map cfg YAML
 * (map_server_planexp.cpp:73-140) and the PlannerParams it hands to the planner (:181-192,
 * base_planner.h:13-37).
 */
typedef struct sipp_map_desc {
  int32_t W;                    /* horizontal_units_ (cfg resolution[0])                                */
  int32_t H;                    /* vertical_units_   (cfg resolution[1])                                */
  double width;                 /* [m] */
  double height;                /* [m] */
  double start_x, start_y;      /* start_location [m] */
  double goal_x, goal_y;        /* goal_location  [m] */
  double min_rov_cost;          /* PlannerParams.min_cost  (map_server_planexp.cpp:124-125)             */
  double max_rov_cost;          /* PlannerParams.max_cost  (:126-127)                                   */
  double unknown_init_cost;     /* PlannerParams.init_cost (:165-179)                                   */
  int32_t unknown_id;           /* 3599 */
  int32_t n_obstacle_ids_mav;   /* ids that make a cell MSR_OCCUPIED in the map state (:252)            */
  int32_t obstacle_ids_mav[SIPP_MAX_IDS];
  int32_t n_obstacle_ids_rov;   /* ids whose graph edges the follower planner deletes                   */
  int32_t obstacle_ids_rov[SIPP_MAX_IDS]; /* (prm_star_planner.cpp:113; the eval binary appends unknown_id, utils.h:272) */
  int32_t compute_reachability; /* map/compute_reachability (map_planexp.cpp:27)                        */
  int32_t compute_closest_observed; /* map/compute_closest_observed (map_planexp.cpp:28)                */
} sipp_map_desc;

/* ---- gain evaluators (following_evaluator `type:` strings) ---------------------------------------- */
#define SIPP_EVAL_RRT_STAR 0          /* "RRTStarEvaluator"            rrt_star_evaluator.cpp:36-96                 */
#define SIPP_EVAL_GOAL_DISTANCE 1     /* "GoalDistanceEvaluator"       goal_distance_evaluator.cpp:33-66            */
#define SIPP_EVAL_EXPAND_REACHABLE 2  /* "ExpandReachableAreaEvaluator" expand_reachable_area_evaluator.cpp:33-67   */
#define SIPP_EVAL_AVERAGE_VIEW_COST 3 /* "AverageViewCostEvaluator"    average_view_cost_evaluator.cpp:31-61        */
#define SIPP_EVAL_EXPAND_LOW_COST 4   /* "ExpandLowCostAreaEvaluator"  expand_low_cost_area_evaluator.cpp:33-72     */
#define SIPP_EVAL_NAIVE 5             /* [upstream] "NaiveEvaluator": gain = number of unknown visible voxels       */
#define SIPP_EVAL_COUNT_ 6

/*
 * Evaluator + sensor-model parameters; same keys and defaults as the reference's setupFromParamMap
 * (rrt_star_evaluator.cpp:22-25, goal_distance_evaluator.cpp:22, expand_reachable_area_evaluator.cpp:22,
 *  expand_low_cost_area_evaluator.cpp:22, camera_model_planexp.cpp:20) plus the [upstream]
 *  SimulatedSensorEvaluator bounding volume (cfg/experiments/example.yaml:12-18).
 */
typedef struct sipp_eval_params {
  int32_t evaluator;             /* SIPP_EVAL_*                                               */
  int32_t half_fov;              /* sensor_model/half_fov, default 10, shipped cfg 32          */
  double non_path_voxel_gain;    /* RRTStarEvaluator, default 0.0                              */
  int32_t use_distance_discount; /* RRTStarEvaluator, default false                            */
  int32_t do_binary_check;       /* RRTStarEvaluator, default false                            */
  double discount_offset;        /* RRTStarEvaluator, default 0.1                              */
  double max_gain;               /* GoalDistanceEvaluator, default 100.0                       */
  double discount_factor;        /* ExpandReachableAreaEvaluator, default 0.1                  */
  int32_t augment_unknown_with_mean; /* ExpandLowCostAreaEvaluator, default false              */
  int32_t use_bounding_volume;   /* 0: every on-map voxel passes (the shipped cfgs)            */
  double bv_x_min, bv_x_max, bv_y_min, bv_y_max; /* inclusive, z is always inside (voxel z = 0)*/
} sipp_eval_params;

/* ---- evaluator-updater policy (A13): which tree nodes are re-scored in a cycle ------------------------------------- */
#define SIPP_UPD_RRT_STAR 0     /* "RRTStarUpdater"  advanced_planner_modules/src/evaluator_updater/rrt_star_updater.cpp:19-42 */
#define SIPP_UPD_CONSTRAINED 1  /* [upstream] "ConstrainedUpdater" (cfg/planners/example_config.yaml:109-117)               */
#define SIPP_UPD_DIRECT 2       /* the following updater IS the evaluator_updater (cfg/evaluators/expand_reachable_area.yaml:23-27) */
#define SIPP_FOLLOW_UPDATE_ALL 0              /* [upstream] "UpdateAll"                                                  */
#define SIPP_FOLLOW_UPDATE_ALL_NON_TERMINAL 1 /* "UpdateAllNonTerminal"  src/evaluator_updater/update_all_non_terminal.cpp:16-27 */
typedef struct sipp_updater_params {
  int32_t updater;        /* SIPP_UPD_*                                   */
  int32_t following;      /* SIPP_FOLLOW_*                                */
  double minimum_gain;    /* default -1e9 (rrt_star_updater.cpp:45)       */
  double update_range;    /* default  1e9 (:46)                           */
  int32_t update_gain, update_cost, update_value;   /* following updater flags, default true */
  int32_t pad_;
} sipp_updater_params;

/* counts[] layout written per candidate by sipp_gain_batch (SIPP_COUNTS_PER_CANDIDATE uint32 each). */
#define SIPP_COUNTS_PER_CANDIDATE 4
#define SIPP_CNT_VISIBLE 0  /* voxels emitted by the footprint after the bounding-volume filter (centre counted twice) */
#define SIPP_CNT_UNKNOWN 1  /* of those, state == UNKNOWN                                                               */
#define SIPP_CNT_AUX 2      /* RRTStar: unknown & on path mask; ExpandReachable: unknown & 3x3-reachable;               */
                            /* AverageViewCost: observed FREE voxels; others: 0                                         */
#define SIPP_CNT_AUX2 3     /* AverageViewCost: sum of map cost over observed FREE voxels (integer labels); else 0      */

typedef struct sipp_ctx sipp_ctx;

/* ---- lifecycle ------------------------------------------------------------------------------------- */
int sipp_abi_version(void);
/* Replaces MapServerPlanExp::MapServerPlanExp + PRMStarPlanner::setup (map_server_planexp.cpp:6-218,
 * prm_star_planner.cpp:11-36,132-156): allocates all map planes on `device`, builds the geometry tables
 * of the per-pixel grid graph (no explicit edge lists), clears the map. */
int sipp_create(const sipp_map_desc* desc, int device, sipp_ctx** out);
void sipp_destroy(sipp_ctx* ctx);
/* message of the last failing call on this ctx (or of the last failing sipp_create when ctx == NULL) */
const char* sipp_last_error(const sipp_ctx* ctx);
/* the stream every host-pointer entry point of this ctx runs on (cudaStream_t) */
void* sipp_stream(sipp_ctx* ctx);

/* ---- map store (A9/A10) ---------------------------------------------------------------------------- */
/* MapServerPlanExp::clearMap (map_server_planexp.cpp:461-487) + planner reset to init_cost. */
int sipp_map_clear(sipp_ctx* ctx);
/* top-left cell of a rows x cols patch centred on pose (px,py): map_server_planexp.cpp:222-228,340-342. */
int sipp_patch_origin(const sipp_ctx* ctx, double px, double py, int32_t rows, int32_t cols,
                      int32_t* x0, int32_t* y0);
/* MapServerPlanExp::mapDataCallback cell loop + PRMStarPlanner::update_states + reachability /
 * closest-observed bookkeeping (map_server_planexp.cpp:238-276, prm_star_planner.cpp:104-119).
 * labels: rows x cols int32 row-major (MapData.data). Off-map cells are skipped. */
int sipp_map_update_patch(sipp_ctx* ctx, int32_t x0, int32_t y0, int32_t rows, int32_t cols,
                          const int32_t* labels);
/* Whole-map ingest: equivalent to ONE mapDataCallback whose patch is the full W x H label image with
 * top-left (0,0), restricted to the cells where observed[y*W+x] != 0 (observed == NULL: all cells).
 * This is how planexp_eval_planner feeds a fully known map (planexp_eval_planner.cpp:263-285). */
int sipp_map_upload_full(sipp_ctx* ctx, const int32_t* labels, const uint8_t* observed);
/* PRMStarPlanner::update_unknown_cost (prm_star_planner.cpp:121-130). */
int sipp_update_unknown_cost(sipp_ctx* ctx, double new_cost);
/* MapServerPlanExp::getMeanCost / getMaxCost (map_server_planexp.h:91-92). */
int sipp_map_stats(const sipp_ctx* ctx, double* mean_cost, double* max_cost, int64_t* mean_counter);
/* Debug / test read-back of the per-cell planes (any pointer may be NULL). state: 0/1/2; cost: last
 * label; reach: 1 iff MSR_REACHABLE; path: current path mask (0/1). All W*H row-major. */
int sipp_map_download(sipp_ctx* ctx, uint8_t* state, int32_t* cost, uint8_t* reach, uint8_t* path);

/* ---- follower path cost (A11/A12) ------------------------------------------------------------------ */
/* PRMStarPlanner::plan + MapServerPlanExp::computeCurrentPathEstimate/computePathMap
 * (prm_star_planner.cpp:38-70, map_server_planexp.cpp:880-926): start-rooted fp64 cost field on the
 * implicit per-pixel graph (bit-exact fixed point), cost = d[goal] (-1.0 when unreachable), canonical
 * predecessor path, rasterised path mask swapped in as the current mask, path generation += 1.
 * path_xy: optional out, up to path_cap points (x,y metres, start -> goal); *path_len = number of path
 * nodes (may exceed path_cap; only path_cap are written). status: SIPP_PLAN_*. */
int sipp_plan(sipp_ctx* ctx, double* cost, int32_t* status, double* path_xy, int32_t path_cap,
              int32_t* path_len);
/* the path of the last sipp_plan again (start -> goal, x y metres): up to path_cap points are written, *path_len = its full length.
 * For callers whose buffer was too small for a long, winding path (PRMStarPlanner::plan hands back every node, :54-62). */
int sipp_path_download(sipp_ctx* ctx, double* path_xy, int32_t path_cap, int32_t* path_len);
/* the converged field d[] of the last sipp_plan, W*H doubles row-major, +inf = unreachable */
int sipp_field_download(sipp_ctx* ctx, double* field);
/* MapServerPlanExp::getCurrentPathMapId (map_server_planexp.h:103) */
int sipp_path_mask_generation(const sipp_ctx* ctx);
/* inject a path mask computed elsewhere (tests; MapPlanExp::updatePathMap map_planexp.cpp:240-242) */
int sipp_set_path_mask(sipp_ctx* ctx, const uint8_t* mask);
/* relaxation statistics of the last sipp_plan: [0]=tile queue pushes, [1]=tile activations, [2]=tile sweeps, [3]=aborted */
int sipp_plan_stats(const sipp_ctx* ctx, int64_t stats[4]);

/* device-timed phases of the last sipp_plan (CUDA events on the context's stream): ms[0] = the relaxation kernel alone, ms[1] = the
 * whole plan (field + predecessors + path + mask). Measurement aid for bench.py's cost-to-go roofline entry. */
int sipp_plan_phases(sipp_ctx* ctx, double ms[2]);

/* Incremental re-plan. Between two plans only ingested patches change the planner's map, so sipp_plan
 * reuses the previous fixed point: cells first observed at a HIGHER cost than the planner assumed (obstacles, labels above the
 * init cost) invalidate — along the predecessor plane of the last plan — exactly the cost-to-go values whose path ran through
 * them; those are reset to +inf and the relaxation restarts from the patch tiles and the rim of the invalidated region. The
 * result is bit-identical to a plan from scratch (the least fixed point is unique and every surviving value is an upper bound
 * of it). A plan falls back to a full one after sipp_map_clear / sipp_update_unknown_cost / a change of the init cost, or when
 * the patches cover more than half of the tiles. On by default; SIPP_INCREMENTAL=0 or sipp_set_incremental(ctx, 0) turn it off.
 * info: [0] 1 = the last plan was incremental, [1] tiles touched by patches, [2] tiles the relaxation started from,
 * [3] cells invalidated. */
int sipp_set_incremental(sipp_ctx* ctx, int32_t enable);
int sipp_plan_info(const sipp_ctx* ctx, int64_t info[4]);

/* ---- batched footprint gain (A1-A8) ---------------------------------------------------------------- */
/* Replaces, for n candidates at once, SimulatedSensorEvaluator::computeGain -> getVisibleVoxelsFromTrajectory
 * -> MapPlanExp::getVoxelArea -> <Evaluator>::computeGainFromVisibleVoxels
 * (sensor_model_planexp.cpp:18-63, camera_model_planexp.cpp:23-28, map_planexp.cpp:81-101,
 *  trajectory_evaluator/(all five evaluator sources)).
 * xy:        n x 2 doubles, view position (last trajectory point + mounting translation).
 * start_xy:  n x 2 doubles, trajectory[0].position_W of each segment; only read by RRTStarEvaluator in
 *            binary+discount mode (rrt_star_evaluator.cpp:65), may be NULL otherwise.
 * gain:      n doubles out.           counts: n x SIPP_COUNTS_PER_CANDIDATE uint32 out (may be NULL).
 * terminal:  n bytes out, 1 where the reference sets gain_terminal (may be NULL). */
int sipp_gain_batch(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* xy,
                    const double* start_xy, double* gain, uint32_t* counts, uint8_t* terminal);
int sipp_gain_batch_dev(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* d_xy,
                        const double* d_start_xy, double* d_gain, uint32_t* d_counts,
                        uint8_t* d_terminal, void* stream);

/* ---- value + selection (A14) ----------------------------------------------------------------------- */
/* [upstream] GlobalNormalizedGain::computeValue + SubsequentBest::selectNextBest on a parent-index tree.
 * Node 0 is the root (parent[0] = -1, its gain/cost are forced to 0 like OnlinePlanner does);
 * parent[i] < i for i > 0. parent == NULL means a flat tree (every node a child of the root).
 * value[i] = max over the subtree of i of (sum gain root..n)/(sum cost root..n) (0 where cost <= 0).
 * best_child = index of the root child with the largest value, lowest index on ties (the reference
 * breaks ties with rand()); -1 when the root has no children. Trees deeper than 512 levels are rejected
 * (SIPP_ERR_UNSUPPORTED; the _dev variant reports it in band as best_child == -2). */
int sipp_value_select(sipp_ctx* ctx, int32_t n, const double* gain, const double* cost,
                      const int32_t* parent, double* value, int32_t* best_child, double* best_value);
int sipp_value_select_dev(sipp_ctx* ctx, int32_t n, const double* d_gain, const double* d_cost,
                          const int32_t* d_parent, double* d_value, int32_t* d_best_child,
                          double* d_best_value, void* stream);

/* Values as the reference's pre-order re-evaluation pass leaves them ([upstream] OnlinePlanner::updateEvaluatorStep calls
 * updateSegment -> computeValue node by node, parents first): value[i] is computed with this cycle's gains (gain_new) on the
 * root..i chain and LAST cycle's gains (gain_old) on i's descendants. Same tree conventions as sipp_value_select. */
int sipp_value_preorder(sipp_ctx* ctx, int32_t n, const double* gain_new, const double* gain_old, const double* cost,
                        const int32_t* parent, double* value);

/* ---- tree re-evaluation pass (A13 + A1-A8 + A14 in one call) ---------------------------------------------------------- */
/* One pass of [upstream] OnlinePlanner::updateEvaluatorStep over a tree in pre-order (node 0 = root, parent[i] < i, children in
 * index order), i.e. what the patched planner triggers with trajectory_evaluator_->updateSegment(root)
 * (mav_active_3d_planning.patch:56): per node the updater chain RRTStarUpdater | ConstrainedUpdater | none ->
 * UpdateAll | UpdateAllNonTerminal (rrt_star_updater.cpp:19-42, update_all_non_terminal.cpp:16-27) decides whether the
 * node's gain (computeGain at end_xy) and value (GlobalNormalizedGain) are recomputed. The predicate runs on the device, only
 * the eligible nodes are scored (one batched gain launch), values get the pre-order semantics of sipp_value_preorder.
 * end_xy / start_xy: n x 2 (last / first trajectory point of each segment; start_xy may be NULL unless RRTStar binary+discount).
 * gain, value, terminal: in/out (terminal is only ever set). cost: in. cur_xy: current robot position (2 doubles).
 * The root (node 0) enters with gain = cost = 0 (it just became the current segment). The root-first call applies the chain to it:
 * an RRTStarUpdater only compares path ids there (:21,30-38) and leaves the root alone; a ConstrainedUpdater / a following updater
 * used directly re-score it like any node, and its new gain is part of every value below it (gain[0] / value[0] / terminal[0] are
 * then written too).
 * Costs are the caller's: the library never recomputes them inside the pass. With update_cost set ([upstream] UpdateAll /
 * UpdateAllNonTerminal call computeCost before computeValue, update_all_non_terminal.cpp:20-22) the caller recomputes the costs of the
 * nodes the chain follows first (sipp_cost_integral; SegmentTime costs do not depend on the map) and passes them in `cost`, and the
 * costs as they were before the pass in `cost_below` (what computeValue(i) still finds on the nodes below i; cost[0] is then the
 * root's cost of this pass: 0 unless the chain re-scores the root); cost_below == NULL: costs did not change.
 * path_changed: RRTStarUpdater's rrt_path_has_changed_ for this pass. n_rescored (optional out): nodes whose gain was recomputed. */
int sipp_update_tree(sipp_ctx* ctx, const sipp_eval_params* params, const sipp_updater_params* updater, int32_t n,
                     const int32_t* parent, const double* end_xy, const double* start_xy, double* gain, const double* cost,
                     double* value, uint8_t* terminal, const double* cur_xy, int32_t path_changed, int32_t* n_rescored,
                     const double* cost_below);

/* ---- fused evaluation cycle ------------------------------------------------------------------------ */
/* gain for all n candidates (flat tree, candidate i is node i+1 under a virtual root) + value + argmax
 * in one call; what the batched root-first updateSegment (mav_active_3d_planning.patch:56) needs.
 * best_index is a candidate index in [0,n) or -1. */
int sipp_cycle(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* xy,
               const double* start_xy, const double* cost, double* gain, uint8_t* terminal,
               double* value, int32_t* best_index, double* best_value);
/* Device variant; d_best is a 16-byte record {double value; int64 index} (what the multi-GPU gather moves). */
int sipp_cycle_dev(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* d_xy,
                   const double* d_start_xy, const double* d_cost, double* d_gain, uint8_t* d_terminal,
                   double* d_value, void* d_best, void* stream);

/* ---- multi-GPU: merge of the per-shard best records ---------------------------------------------------- */
/* d_recs: `world` 16-byte records {double value; int64 index} as gathered from the ranks (NCCL all_gather of the d_best
 * of sipp_cycle_dev); record r holds a shard-local index that becomes global by adding r * shard_stride (pass 0 when the
 * indices are already global). Picks what one rank scanning all candidates would select: the largest value, the lowest
 * global index on ties; NaN never beats a number; a record with index < 0 (empty shard) never wins; if no record is
 * comparable the first one is returned. Writes one 16-byte record with the GLOBAL index to d_out. One tiny kernel. */
int sipp_merge_best_dev(sipp_ctx* ctx, const void* d_recs, int32_t world, int64_t shard_stride, void* d_out, void* stream);

/* ---- trajectory cost: line integral of the map cost --------------------------------------------- */
/* "CostPlanExpMapIntegral" (advanced_planner_modules/src/cost_computer/cost_planexp_map_integral.cpp:15-73,
 * cfg/evaluators/baseline.yaml:14-15): parameters as in setupFromParamMap (:15-18). */
typedef struct sipp_cost_params {
  int32_t accumulate;            /* default false: add the parent's cost                         */
  int32_t pad_;
  double max_sample_distance;    /* default 0.05 m                                               */
} sipp_cost_params;
/* computeCost for n trajectory segments at once. Segment i owns the points [point_offset[i], point_offset[i+1]) of
 * points_xyz (x, y, z doubles: trajectory[k].position_W). cost[i] = sum over the consecutive point pairs of the sampled
 * integral (sub-samples every max_sample_distance along the pair, each charged distance * mean of the two map costs, or
 * distance * mean map cost when either sample lies on an unobserved cell; off-map samples read -1 like getCostAtLocation);
 * + parent_cost[i] when accumulate is set and parent_cost != NULL. valid[i] (optional) = computeCost's return value: 0 for
 * segments with fewer than two points (their cost is 0). Every segment is integrated sequentially in the reference's
 * operation order (one thread per segment), so the result is bit-identical to the C++ loop. */
int sipp_cost_integral(sipp_ctx* ctx, const sipp_cost_params* params, int32_t n, const int32_t* point_offset,
                       const double* points_xyz, const double* parent_cost, double* cost, uint8_t* valid);

/* ---- simulator travel cost (second half) -------------------------------------------------------------------- */
/* MavSimulator::computeCost (mav_simulator/src/mav_simulator.cpp:785-821): the cost the simulator books for every flown segment,
 * integrated over the GROUND-TRUTH label map (mav_simulator.h:112 map_), not the observed one. The truth map is uploaded once
 * (sipp_truth_upload; W*H int32 row-major like every host plane here) and stays resident; the episode / ensemble runners cut the
 * scout's observation patches out of it on the device as well.
 * Per (start, goal) pair: distance = |goal - start|, n = ceil(distance / max_step_size) equal steps, each charged
 *   MapCost:      mean of the two end cells' labels * step length   (:799,805)   pixel = (int)(p * cols / map_width) (:582-587)
 *   DistanceCost: step length                                        (:808)
 *   TimeCost:     step length / mav_speed                            (:811)
 * accumulated sequentially in the reference's order (one thread per pair): bit-identical to the C++ loop. A sample outside the
 * map (undefined behaviour in the reference: unchecked Eigen index) reads label 0. */
#define SIPP_TRAVEL_MAP_COST 0
#define SIPP_TRAVEL_DISTANCE_COST 1
#define SIPP_TRAVEL_TIME_COST 2
typedef struct sipp_travel_params {
  int32_t formulation;      /* SIPP_TRAVEL_* (mav_simulator/include/mav_simulator/types.h:7-11, mav_simulator.cpp:33-43) */
  int32_t pad_;
  double max_step_size;     /* default 0.05 m (mav_simulator.h:104) */
  double mav_speed;         /* default 1.0 m/s (mav_simulator.h:134) */
} sipp_travel_params;
int sipp_truth_upload(sipp_ctx* ctx, const int32_t* labels);
int sipp_travel_cost(sipp_ctx* ctx, const sipp_travel_params* params, int32_t n, const double* start_xyz, const double* goal_xyz,
                     double* cost);

/* ---- episode replay (offline evaluation replay / ensembles of independent maps) ------------------ */
/* The host loop of one exploration episode, run natively so that many episodes can be driven side by side from plain host
 * threads (one context each; contexts are independent and their kernels overlap on the GPU): per cycle c
 *   1. ingest the fov x fov patch of `truth` (W x H row-major labels) centred on the scout, as the simulator publishes it and
 *      mapDataCallback ingests it (mav_simulator.cpp:438-479, map_server_planexp.cpp:222-276); before the first cycle the
 *      init x init square at the map origin (mav_simulator.cpp:156-159; init == 0: none);
 *   2. re-plan the follower path (sipp_plan);
 *   3. score the n candidate views cand_xy[c][n][2] with travel costs cand_cost[c][n] and select (sipp_cycle);
 *   4. move the scout to the selected candidate (it stays where it is when none wins).
 * Outputs (any may be NULL), one entry per cycle: path cost, SIPP_PLAN_* status, path nodes, selected candidate, its value.
 * Replaces, for evaluation runs, the loop data_analysis/evaluation.py:39-62 + planexp_eval_planner.cpp:201-303 drive. */
int sipp_episode_run(sipp_ctx* ctx, const int32_t* truth, const sipp_eval_params* params, int32_t cycles, int32_t n,
                     const double* cand_xy, const double* cand_cost, int32_t fov, int32_t init, double start_x, double start_y,
                     double* path_cost, int32_t* plan_status, int32_t* path_len, int32_t* best_index, double* best_value);

/* ---- segments viewed from several sampled viewpoints (sensor_model/sampling_time > 0) ---------------------------------------
 * SensorModelPlanExp::getVisibleVoxelsFromTrajectory (sensor_model_planexp.cpp:18-63) views a segment from every trajectory point
 * sampleViewpoints (:65-90) selects — with the mounting translation applied (:28-31) — concatenates the views' voxel lists and
 * removes only adjacent duplicates; the evaluators then count the whole list. The host selects the viewpoints (the rule is host
 * logic over time stamps: plugin/sipp_modules.cpp), these calls score them: segment i has the views
 * view_xy[view_offset[i] .. view_offset[i+1]) (view_offset: n + 1 entries starting at 0; a segment without views scores 0 and,
 * for the evaluators that set it, is terminal — an empty voxel list). start_xy: trajectory[0] per segment (RRTStarEvaluator binary
 * + discount; may be NULL otherwise). One view per segment gives exactly sipp_gain_batch / sipp_update_tree. */
int sipp_gain_batch_views(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const int32_t* view_offset, const double* view_xy,
                          const double* start_xy, double* gain, uint32_t* counts, uint8_t* terminal);
int sipp_update_tree_views(sipp_ctx* ctx, const sipp_eval_params* params, const sipp_updater_params* updater, int32_t n, const int32_t* parent,
                           const double* end_xy, const double* start_xy, double* gain, const double* cost, double* value, uint8_t* terminal,
                           const double* cur_xy, int32_t path_changed, int32_t* n_rescored, const double* cost_below, const int32_t* view_offset,
                           const double* view_xy);

/* ---- batched ensembles: m maps of one size on one device advance together ---------------------------------------------------
 * Every kernel of a step (patch ingest, re-plan, view planes, gain + value + argmax) is launched ONCE for all maps (gridDim.y = m,
 * per-map argument blocks in device memory), with one host->device copy of the inputs, one device->host copy of the results and
 * one synchronisation per map-cycle — instead of m contexts x ~20 stream operations fed by m host threads. The contexts keep all
 * state (sipp_map_download, sipp_field_download, sipp_path_download ... work between batch calls); they must be idle when a batch
 * call starts and must not be driven from other threads while it runs. Results are bit-identical to the same calls made map by
 * map (tests/test_gpu_batch.py). The two persistent kernels of a re-plan (invalidation, relaxation) run on ONE work queue per batch:
 * their workers take tiles of whichever map has work, so a batch-cycle lasts as long as the batch's total work, not as its slowest
 * map (a watchdog abort of that queue fails every map of the step with SIPP_ERR_CUDA).
 * Limits: maps without compute_closest_observed; sipp_batch_cycle serves every evaluator except ExpandLowCostArea (its
 * closest-observed planes) and RRTStarEvaluator's binary + discount mode (per-candidate start points) and refuses those two.
 * Replaces, for evaluation runs over many maps, the one-planner-process-per-saved-map loop of data_analysis/evaluation.py:39-62,
 * 139-144 + planexp_base_planners/src/planexp_eval_planner.cpp:201-303 and the one-experiment-per-map loop of
 * advanced_planner_ros/scripts/run_experiment.sh:190-212. */
typedef struct sipp_batch sipp_batch;
int sipp_batch_create(sipp_ctx* const* ctxs, int32_t m, sipp_batch** out);   /* on failure the message is in sipp_last_error(ctxs[i]) */
void sipp_batch_destroy(sipp_batch* batch);                                   /* the contexts stay alive */
const char* sipp_batch_last_error(const sipp_batch* batch);
int64_t sipp_batch_launch_count(const sipp_batch* batch);
/* How many batches the caller drives side by side on this GPU (default 1, or SIPP_BATCH_SHARE): the batch's persistent relaxation /
 * invalidation workers then take 1/n of the SMs, so that the chain-bound phases of one batch run beside the busy phases of another. */
int sipp_batch_set_sharing(sipp_batch* batch, int32_t batches_on_this_gpu);
/* sipp_map_update_patch for every map: map i ingests the rows x cols patch labels[i] at (x0[i], y0[i]) */
int sipp_batch_ingest(sipp_batch* batch, const int32_t* x0, const int32_t* y0, int32_t rows, int32_t cols, const int32_t* const* labels);
/* sipp_plan for every map (from scratch or incremental, decided per map): cost / status / path_len are m entries (any may be NULL) */
int sipp_batch_plan(sipp_batch* batch, double* cost, int32_t* status, int32_t* path_len);
/* sipp_cycle for every map: xy[i] = n x 2, cost[i] = n candidates of map i; best_index / best_value: m entries; gain: optional m
 * pointers to n doubles each (entries may be NULL) */
int sipp_batch_cycle(sipp_batch* batch, const sipp_eval_params* params, int32_t n, const double* const* xy, const double* const* cost,
                     int64_t* best_index, double* best_value, double* const* gain);
/* sipp_episode_run for every map, in lockstep. truth[i]: W x H labels; cand_xy[i] / cand_cost[i]: cycles x n candidates of map i.
 * Outputs (any may be NULL): m x cycles, map-major. */
int sipp_batch_episode_run(sipp_batch* batch, const int32_t* const* truth, const sipp_eval_params* params, int32_t cycles, int32_t n,
                           const double* const* cand_xy, const double* const* cand_cost, int32_t fov, int32_t init, double start_x,
                           double start_y, double* path_cost, int32_t* plan_status, int32_t* path_len, int32_t* best_index, double* best_value);

/* ---- multi-GPU: exchange of the shard winners through peer memory (NVLink / NVSwitch), fused into the argmax kernel ----- */
/* One process per GPU (or several contexts in one process). Every rank exports a handle to its mailbox, the host program
 * moves the handles between the ranks with whatever it has (MPI, torch.distributed, a ROS topic), and every rank attaches
 * all of them. After that sipp_cycle_shard[_dev] scores this rank's shard and its argmax kernel (a) stores the shard's
 * 16-byte winner into every peer's mailbox over NVLink, (b) waits for the peers' records of the same step in its own
 * mailbox and (c) merges them: every rank ends the call with the GLOBAL winner — the candidate one GPU scanning all shards
 * in index order would select (largest value, lowest global index on ties; NaN / empty shards never win). No collective
 * library and no host round trip on the data path. All ranks must issue the same sequence of shard calls (lockstep); a rank
 * that waits longer than the timeout (SIPP_PEER_TIMEOUT_MS, default 20000) returns index -3 and sipp_peer_status != 0.
 * The blocking host-pointer form (sipp_cycle_shard) needs one rank per device: ranks that share a device share its driver
 * context, and a rank blocked in its call keeps the peer it waits for from launching — contexts that share a device use
 * the asynchronous _dev forms, enqueued by one thread.
 * Replaces, for N GPUs, [upstream] SubsequentBest::selectNextBest over the whole candidate set. */
#define SIPP_PEER_MAX_WORLD 16
#define SIPP_PEER_HANDLE_BYTES 128
int sipp_peer_export(sipp_ctx* ctx, void* handle /* SIPP_PEER_HANDLE_BYTES out */);
int sipp_peer_attach(sipp_ctx* ctx, int32_t rank, int32_t world, const void* handles /* world x SIPP_PEER_HANDLE_BYTES */);
int sipp_peer_detach(sipp_ctx* ctx);
/* 0: healthy; 1: an exchange timed out. Synchronises the ctx stream. */
int sipp_peer_status(sipp_ctx* ctx, int32_t* status, int64_t* steps_done);
/* sipp_cycle_dev over this rank's shard + the fused exchange. index_offset = global index of this shard's candidate 0.
 * d_best_global: 16-byte record {double value; int64 index} of the global winner (identical on every rank). */
int sipp_cycle_shard_dev(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* d_xy,
                         const double* d_start_xy, const double* d_cost, double* d_gain, uint8_t* d_terminal,
                         double* d_value, int64_t index_offset, void* d_best_global, void* stream);
/* Pipelined form for back-to-back cycles: publishes this call's shard winner but collects the records of the PREVIOUS call
 * (normally already there, so the exchange latency hides under this call's gain kernel). d_prev_best_global receives the
 * global winner of the previous pipelined call ({0.0, -1} on the first one); sipp_peer_flush_dev collects the last one.
 * Do not mix with the synchronous calls without a flush in between. */
int sipp_cycle_shard_pipelined_dev(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* d_xy,
                                   const double* d_start_xy, const double* d_cost, double* d_gain, uint8_t* d_terminal,
                                   double* d_value, int64_t index_offset, void* d_prev_best_global, void* stream);
int sipp_peer_flush_dev(sipp_ctx* ctx, void* d_best_global, void* stream);
/* host-pointer form (sipp_cycle + exchange): the call a host program makes per planning cycle on every rank */
int sipp_cycle_shard(sipp_ctx* ctx, const sipp_eval_params* params, int32_t n, const double* xy, const double* start_xy,
                     const double* cost, int64_t index_offset, double* gain, uint8_t* terminal, double* value,
                     int64_t* best_index_global, double* best_value);
/* exchange of an arbitrary 16-byte record (e.g. the tree-mode winner): d_mine in, merged global record out */
int sipp_peer_exchange_dev(sipp_ctx* ctx, const void* d_mine, void* d_out, void* stream);

/* number of kernel launches issued by this ctx since creation (bench.py's gpu_launches) */
int64_t sipp_launch_count(const sipp_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SIPP_H_ */
