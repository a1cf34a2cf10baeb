// =====================================================================================================
// sipp_modules.h — C++ host-side mirror of the scouting-ipp module interface for the evaluation hot path.
// Thin adapters (no arithmetic): they translate the reference's module calls into the C ABI of include/sipp.h.
//
//   type string (cfg `type:`)            replaces (reference file:line)
//   "GpuMapPlanExp"                      MapPlanExp + MapServerPlanExp + PRMStarPlanner ownership
//                                        (advanced_planner_modules/src/map/map_planexp.cpp:14-46,
//                                         map_server_planexp/src/map_server_planexp.cpp:6-282,880-926)
//   "GpuRRTStarEvaluator"                RRTStarEvaluator               (src/trajectory_evaluator/rrt_star_evaluator.cpp:12-96)
//   "GpuGoalDistanceEvaluator"           GoalDistanceEvaluator          (goal_distance_evaluator.cpp:12-66)
//   "GpuExpandReachableAreaEvaluator"    ExpandReachableAreaEvaluator   (expand_reachable_area_evaluator.cpp:12-67)
//   "GpuAverageViewCostEvaluator"        AverageViewCostEvaluator       (average_view_cost_evaluator.cpp:12-61)
//   "GpuExpandLowCostAreaEvaluator"      ExpandLowCostAreaEvaluator     (expand_low_cost_area_evaluator.cpp:12-72)
//   "GpuNaiveEvaluator"                  [upstream] NaiveEvaluator
// Each Gpu*Evaluator is a full TrajectoryEvaluator: computeGain (batch of one), computeCost ([upstream] SegmentTime),
// computeValue / selectNextBest ([upstream] GlobalNormalizedGain / SubsequentBest) and — the batching point — updateSegment:
// the patched planner calls trajectory_evaluator_->updateSegment(root) FIRST (mav_active_3d_planning.patch:56); on that call
// the module flattens the whole tree and makes ONE sipp_update_tree call: the updater policy (RRTStarUpdater / ConstrainedUpdater +
// UpdateAll / UpdateAllNonTerminal, rrt_star_updater.cpp:19-42, update_all_non_terminal.cpp:16-27) is evaluated on the DEVICE for
// all nodes at once, the eligible nodes are compacted and scored by one batched gain launch, the values get the pre-order semantics
// of the reference pass (updaterMasks() is the host statement of the same predicate, used for the batched re-costing when
// update_cost is set with the map-integral cost computer and by the tests); the per-node updateSegment calls that follow return
// true without work.
//   GpuPRMStarPlanner : planexp_base_planners::PlannerBase  replaces PRMStarPlanner (prm_star_planner.cpp:11-186).
// Error convention as in the reference: bool returns, configuration problems through checkParamsValid / printError,
// never an exception across the module boundary.
// =====================================================================================================
#ifndef SIPP_MODULES_H_
#define SIPP_MODULES_H_

#ifdef SIPP_WITH_UPSTREAM
#include "active_3d_planning_core/module/module_factory_registry.h"
#include "active_3d_planning_core/module/trajectory_evaluator.h"
#include "active_3d_planning_core/map/map.h"
#include "planexp_base_planners/base_planner.h"
#else
#include "shim/active_3d_planning_shim.h"
#endif

#include <memory>
#include <string>
#include <vector>

#include "../../include/sipp.h"

namespace active_3d_planning {
namespace map {

class GpuMapPlanExp : public Map {
 public:
  explicit GpuMapPlanExp(PlannerI& planner);
  ~GpuMapPlanExp() override;
  void setupFromParamMap(Module::ParamMap* param_map) override;
  bool checkParamsValid(std::string* error_message) override;

  // Map interface
  bool isObserved(const Eigen::Vector3d& point) override;      // map_planexp.cpp:59-61
  void targetReachedCallback() override;                       // map_planexp.h:87 -> map_server_->request_path()

  // MapServerPlanExp side
  bool mapDataCallback(const Eigen::Vector3d& pose, const Eigen::ArrayXXi& labels);   // map_server_planexp.cpp:220-282
  bool computeCurrentPathEstimate();                           // :880-900 (synchronous here)
  int getPathMapId() const;                                    // map_planexp.cpp:236-238
  void updatePathMap() {}                                      // :240-242 — the mask lives on the device, nothing to copy
  double getMeanWeight() const;
  Eigen::Vector3d getGoalPosition() const { return Eigen::Vector3d(desc_.goal_x, desc_.goal_y, 0.0); }
  double lastPathCost() const { return last_cost_; }
  int lastPathStatus() const { return last_status_; }
  const std::vector<Eigen::Vector2d>& lastPath() const { return last_path_; }

  sipp_ctx* ctx() { return ctx_; }
  const sipp_map_desc& desc() const { return desc_; }

 protected:
  static ModuleFactoryRegistry::Registration<GpuMapPlanExp> registration;
  sipp_map_desc desc_;
  sipp_ctx* ctx_ = nullptr;
  int device_ = 0;
  bool compute_path_ = false, path_requested_ = false;
  std::string error_;
  std::vector<uint8_t> state_cache_;
  bool state_dirty_ = true;
  double last_cost_ = -1.0;
  int last_status_ = SIPP_PLAN_TERMINAL_INVALID;
  std::vector<Eigen::Vector2d> last_path_;
};

}  // namespace map

namespace trajectory_evaluator {

// flattened pre-order view of a TrajectorySegment tree (node 0 = root)
struct FlatTree {
  std::vector<TrajectorySegment*> nodes;
  std::vector<int32_t> parent;
  std::vector<double> end_xy, start_xy, gain, cost, value;
  std::vector<uint8_t> terminal;
  // the viewpoints every node is scored from (SensorModelPlanExp::sampleViewpoints + mounting translation): node i has
  // view_xy[view_off[i] .. view_off[i+1]); filled when the sensor model samples along the segment, empty otherwise
  std::vector<int32_t> view_off;
  std::vector<double> view_xy;
  void build(TrajectorySegment* root, const struct ViewSampler* sampler = nullptr);
};

// SensorModelPlanExp::sampleViewpoints (sensor_model_planexp.cpp:65-90) + the mounting translation (:28-31) on the host: which
// trajectory points a segment is viewed from, and where the sensor is then
struct ViewSampler {
  double sampling_time = 0.0;                       // sensor_model/sampling_time; <= 0: the last point only
  Eigen::Vector3d mounting_translation;             // sensor_model/mounting_translation_{x,y,z}, rotated by the point's orientation
  bool single_view() const { return sampling_time <= 0.0; }
  bool identity_mount() const { return mounting_translation[0] == 0.0 && mounting_translation[1] == 0.0 && mounting_translation[2] == 0.0; }
  void indices(const EigenTrajectoryPointVector& trajectory, std::vector<int>* out) const;
  void appendViews(const EigenTrajectoryPointVector& trajectory, std::vector<double>* view_xy) const;     // x, y per sampled viewpoint
};

// the host half of row A13: which nodes get computeGain / computeValue in this pass (pre-order semantics)
void updaterMasks(const sipp_updater_params& u, const FlatTree& t, const double cur_xy[2], bool path_changed, std::vector<uint8_t>* do_gain,
                  std::vector<uint8_t>* maybe_value, bool followed_instead_of_value = false);

class GpuGainEvaluator : public TrajectoryEvaluator {
 public:
  GpuGainEvaluator(PlannerI& planner, int evaluator);
  void setupFromParamMap(Module::ParamMap* param_map) override;
  bool checkParamsValid(std::string* error_message) override;
  bool computeGain(TrajectorySegment* traj_in) override;
  bool computeCost(TrajectorySegment* traj_in) override;
  bool computeValue(TrajectorySegment* traj_in) override;
  int selectNextBest(TrajectorySegment* traj_in) override;
  bool updateSegment(TrajectorySegment* segment) override;

  const sipp_eval_params& evalParams() const { return params_; }
  const sipp_updater_params& updaterParams() const { return updater_; }
  int lastBatchSize() const { return last_batch_; }

 protected:
  map::GpuMapPlanExp* map_ = nullptr;
  sipp_eval_params params_;
  sipp_updater_params updater_;
  bool cost_is_integral_ = false;     // cost_computer/type: "SegmentTime" [upstream] or "(Gpu)CostPlanExpMapIntegral"
  sipp_cost_params cost_params_ = {0, 0, 0.05};
  int last_rrt_path_id_ = 0;          // rrt_star_updater.h:30-31
  bool rrt_path_has_changed_ = true;
  int last_batch_ = 0;
  ViewSampler sampler_;               // sensor_model/sampling_time (sensor_model_planexp.cpp:170) + sensor_model/mounting_translation_*
  bool multi_view() const { return !sampler_.single_view() || !sampler_.identity_mount(); }
  std::string config_error_;
};

template <int EVAL>
class GpuEvaluatorT : public GpuGainEvaluator {
 public:
  explicit GpuEvaluatorT(PlannerI& planner) : GpuGainEvaluator(planner, EVAL) {}

 protected:
  static ModuleFactoryRegistry::Registration<GpuEvaluatorT<EVAL>> registration;
};

}  // namespace trajectory_evaluator

namespace initialize {
void sipp_gpu_package();   // keeps the static registrations alive, like initialize::planexp_package() (planexp_package.cpp:9)
}

}  // namespace active_3d_planning

namespace planexp_base_planners {

// Drop-in for PRMStarPlanner behind PlannerBase (prm_star_planner.h:36-95). Owns its own context.
class GpuPRMStarPlanner : public PlannerBase {
 public:
  explicit GpuPRMStarPlanner(PlannerParams params, int device = 0);
  ~GpuPRMStarPlanner() override;
  void setup() override;
  ResultStatus plan(std::vector<Eigen::Vector2d>* path, double& cost) override;
  void store_path(const std::string file_name, const std::string path) override;
  void update_states(Eigen::Vector2i location, Eigen::ArrayXXi map_data) override;
  void update_unknown_cost(double new_cost) override;
  PlannerType get_planner_type() { return PlannerType::PRM_STAR_PLANNER; }
  const std::string& lastError() const { return error_; }

 private:
  sipp_ctx* ctx_ = nullptr;
  int device_;
  std::string error_;
  std::vector<Eigen::Vector2d> current_best_path_;
};

}  // namespace planexp_base_planners

#endif  // SIPP_MODULES_H_
