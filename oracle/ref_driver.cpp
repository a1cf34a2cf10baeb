// Test driver around the REFERENCE'S OWN evaluator sources (TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline leg may load the library this builds; the product never does).
//
// oracle/Makefile (target _ref) compiles, unmodified and where they lie under /root/reference/advanced_planner_modules/src:
//   map/map_planexp.cpp, sensor_model/{sensor_model_planexp,camera_model_planexp}.cpp,
//   trajectory_evaluator/{rrt_star,goal_distance,expand_reachable_area,average_view_cost,expand_low_cost_area}_evaluator.cpp,
//   evaluator_updater/{rrt_star_updater,update_all_non_terminal}.cpp, cost_computer/cost_planexp_map_integral.cpp
// against the stand-in headers in oracle/ref_shim/ (Eigen subset, the un-vendored mav_active_3d_planning interfaces, a MapServerPlanExp
// over flat planes) and links them with this file into oracle/_ref/libsipp_ref.so. This file adds what the reference gets from
// elsewhere: a PlannerI, the [upstream] helper modules its cfg files name (UpdateAll, ConstrainedUpdater, UpdateNothing,
// GlobalNormalizedGain, SubsequentBest, BoundingVolume) restated from the upstream semantics, and a C ABI for ctypes.
// The restated oracle (sipp_oracle.cpp) is checked against this library in tests/test_ref_pin.py.
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../include/sipp.h"
#include "active_3d_planning_core/shim_core.h"
#include "planner_modules_planexp/map/map_planexp.h"

namespace active_3d_planning {

// ---- [upstream] helper modules named by the reference's cfg files (advanced_planner_ros/cfg/planners/example_config.yaml:97-117) ----
namespace evaluator_updater {
class UpdateNothing : public EvaluatorUpdater {       // default following updater: end of every chain
 public:
  explicit UpdateNothing(PlannerI& planner) : EvaluatorUpdater(planner) {}
  bool updateSegment(TrajectorySegment*) override { return true; }
  void setupFromParamMap(Module::ParamMap*) override {}
  static ModuleFactoryRegistry::Registration<UpdateNothing> registration;
};
ModuleFactoryRegistry::Registration<UpdateNothing> UpdateNothing::registration("UpdateNothing");

class UpdateAll : public EvaluatorUpdater {           // [upstream] default_evaluator_updaters: gain, cost, value, then the follower
 public:
  explicit UpdateAll(PlannerI& planner) : EvaluatorUpdater(planner) {}
  bool updateSegment(TrajectorySegment* segment) override {
    if (update_gain_) planner_.getTrajectoryEvaluator().computeGain(segment);
    if (update_cost_) planner_.getTrajectoryEvaluator().computeCost(segment);
    if (update_value_) planner_.getTrajectoryEvaluator().computeValue(segment);
    return following_updater_->updateSegment(segment);
  }
  void setupFromParamMap(Module::ParamMap* param_map) override {
    setParam<bool>(param_map, "update_gain", &update_gain_, true);
    setParam<bool>(param_map, "update_cost", &update_cost_, true);
    setParam<bool>(param_map, "update_value", &update_value_, true);
    std::string args;
    const std::string ns = (*param_map)["param_namespace"];
    setParam<std::string>(param_map, "following_updater_args", &args, ns + "/following_updater");
    following_updater_ = planner_.getFactory().createModule<EvaluatorUpdater>(args, planner_, verbose_modules_);
  }
  static ModuleFactoryRegistry::Registration<UpdateAll> registration;

 private:
  bool update_gain_, update_cost_, update_value_;
  std::unique_ptr<EvaluatorUpdater> following_updater_;
};
ModuleFactoryRegistry::Registration<UpdateAll> UpdateAll::registration("UpdateAll");

class ConstrainedUpdater : public EvaluatorUpdater {  // [upstream]: RRTStarUpdater without the terminal / path-changed terms
 public:
  explicit ConstrainedUpdater(PlannerI& planner) : EvaluatorUpdater(planner) {}
  bool updateSegment(TrajectorySegment* segment) override {
    if (segment->gain > p_minimum_gain_) {
      if ((planner_.getCurrentPosition() - segment->trajectory.back().position_W).norm() <= p_update_range_) {
        return following_updater_->updateSegment(segment);
      }
    }
    return true;
  }
  void setupFromParamMap(Module::ParamMap* param_map) override {
    setParam<double>(param_map, "minimum_gain", &p_minimum_gain_, -1e9);
    setParam<double>(param_map, "update_range", &p_update_range_, 1e9);
    std::string args;
    const std::string ns = (*param_map)["param_namespace"];
    setParam<std::string>(param_map, "following_updater_args", &args, ns + "/following_updater");
    following_updater_ = planner_.getFactory().createModule<EvaluatorUpdater>(args, planner_, verbose_modules_);
  }
  static ModuleFactoryRegistry::Registration<ConstrainedUpdater> registration;

 private:
  double p_minimum_gain_, p_update_range_;
  std::unique_ptr<EvaluatorUpdater> following_updater_;
};
ModuleFactoryRegistry::Registration<ConstrainedUpdater> ConstrainedUpdater::registration("ConstrainedUpdater");
}  // namespace evaluator_updater

namespace value_computer {
class GlobalNormalizedGain : public ValueComputer {   // [upstream]: best (sum gain)/(sum cost) over the subtree, sums start at the root
 public:
  explicit GlobalNormalizedGain(PlannerI& planner) : ValueComputer(planner) {}
  void setupFromParamMap(Module::ParamMap*) override {}
  bool computeValue(TrajectorySegment* traj_in) override {
    double gain = 0.0, cost = 0.0;
    TrajectorySegment* current = traj_in->parent;
    while (current) {
      gain += current->gain;
      cost += current->cost;
      current = current->parent;
    }
    traj_in->value = findBest(traj_in, gain, cost);
    return true;
  }
  static ModuleFactoryRegistry::Registration<GlobalNormalizedGain> registration;

 private:
  double findBest(TrajectorySegment* current, double gain, double cost) {
    double value = 0.0;
    gain += current->gain;
    cost += current->cost;
    if (cost > 0) value = gain / cost;
    for (size_t i = 0; i < current->children.size(); ++i) value = std::max(value, findBest(current->children[i].get(), gain, cost));
    return value;
  }
};
ModuleFactoryRegistry::Registration<GlobalNormalizedGain> GlobalNormalizedGain::registration("GlobalNormalizedGain");
}  // namespace value_computer

class BoundingVolumeRegistration {
 public:
  static ModuleFactoryRegistry::Registration<BoundingVolume> registration;
};
ModuleFactoryRegistry::Registration<BoundingVolume> BoundingVolumeRegistration::registration("BoundingVolume");

}  // namespace active_3d_planning

namespace {
using namespace active_3d_planning;

// the planner the modules see: one map, one evaluator, a factory fed from a namespace table
class RefPlanner : public PlannerI {
 public:
  const Eigen::Vector3d& getCurrentPosition() const override { return position; }
  const Eigen::Quaterniond& getCurrentOrientation() const override { return orientation; }
  Map& getMap() override { return *map; }
  TrajectoryEvaluator& getTrajectoryEvaluator() override { return *evaluator; }
  ModuleFactory& getFactory() override { return factory; }
  SystemConstraints& getSystemConstraints() override { return constraints; }
  void printInfo(const std::string&) override {}
  void printWarning(const std::string&) override {}
  void printError(const std::string& text) override { errors.push_back(text); }

  Eigen::Vector3d position;
  Eigen::Quaterniond orientation;
  std::unique_ptr<Map> map;
  std::unique_ptr<TrajectoryEvaluator> evaluator;
  ModuleFactory factory;
  SystemConstraints constraints;
  std::vector<std::string> errors;
};

std::string num(double v) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.17g", v);
  return buf;
}

const char* evaluator_type(int e) {
  switch (e) {
    case SIPP_EVAL_RRT_STAR: return "RRTStarEvaluator";
    case SIPP_EVAL_GOAL_DISTANCE: return "GoalDistanceEvaluator";
    case SIPP_EVAL_EXPAND_REACHABLE: return "ExpandReachableAreaEvaluator";
    case SIPP_EVAL_AVERAGE_VIEW_COST: return "AverageViewCostEvaluator";
    case SIPP_EVAL_EXPAND_LOW_COST: return "ExpandLowCostAreaEvaluator";
    default: return nullptr;     // NaiveEvaluator is [upstream]: not part of the reference tree
  }
}

}  // namespace

struct ref_ctx {
  sipp_map_desc desc;
  sipp_ref::MapSpec spec;
};

// (re)creates planner + map + evaluator for one parameter set; the map module is the reference's MapPlanExp over ctx->spec
static std::unique_ptr<RefPlanner> make_planner(ref_ctx* c, const sipp_eval_params* p, const sipp_updater_params* u, const sipp_cost_params* cp,
                                                double sampling_time) {
  std::unique_ptr<RefPlanner> pl(new RefPlanner());
  auto& ns = pl->factory.namespaces;
  ns["map"] = {{"type", "MapPlanExp"}, {"compute_path", "true"}, {"compute_reachability", c->desc.compute_reachability ? "true" : "false"},
               {"compute_closest_observed", c->desc.compute_closest_observed ? "true" : "false"}};
  sipp_ref::current_map_spec() = &c->spec;
  pl->map = pl->factory.createModule<Map>("map", *pl, false);
  if (!pl->map) return nullptr;
  if (!p) return pl;
  const char* et = evaluator_type(p->evaluator);
  if (!et) return nullptr;
  ns["ev"] = {{"type", et},
              {"non_path_voxel_gain", num(p->non_path_voxel_gain)},
              {"use_distance_discount", p->use_distance_discount ? "true" : "false"},
              {"do_binary_check", p->do_binary_check ? "true" : "false"},
              {"discount_offset", num(p->discount_offset)},
              {"max_gain", num(p->max_gain)},
              {"discount_factor", num(p->discount_factor)},
              {"augment_unknown_with_mean", p->augment_unknown_with_mean ? "true" : "false"}};
  ns["ev/sensor_model"] = {{"type", "CameraModelPlanExp"}, {"half_fov", std::to_string(p->half_fov)}, {"sampling_time", num(sampling_time)}};
  if (p->use_bounding_volume)
    ns["ev/bounding_volume"] = {{"x_min", num(p->bv_x_min)}, {"x_max", num(p->bv_x_max)}, {"y_min", num(p->bv_y_min)}, {"y_max", num(p->bv_y_max)},
                                {"z_min", "-1.0"}, {"z_max", "1.0"}};
  ns["ev/value_computer"] = {{"type", "GlobalNormalizedGain"}};
  if (cp) ns["ev/cost_computer"] = {{"type", "CostPlanExpMapIntegral"}, {"accumulate", cp->accumulate ? "true" : "false"},
                                    {"max_sample_distance", num(cp->max_sample_distance)}};
  if (u) {
    const char* fol = u->following == SIPP_FOLLOW_UPDATE_ALL ? "UpdateAll" : "UpdateAllNonTerminal";
    Module::ParamMap folmap = {{"type", fol}, {"update_gain", u->update_gain ? "true" : "false"}, {"update_cost", u->update_cost ? "true" : "false"},
                               {"update_value", u->update_value ? "true" : "false"}};
    if (u->updater == SIPP_UPD_DIRECT) ns["ev/evaluator_updater"] = folmap;
    else {
      ns["ev/evaluator_updater"] = {{"type", u->updater == SIPP_UPD_RRT_STAR ? "RRTStarUpdater" : "ConstrainedUpdater"},
                                    {"minimum_gain", num(u->minimum_gain)}, {"update_range", num(u->update_range)}};
      ns["ev/evaluator_updater/following_updater"] = folmap;
    }
  }
  pl->evaluator = pl->factory.createModule<TrajectoryEvaluator>("ev", *pl, false);
  if (!pl->evaluator) return nullptr;
  return pl;
}

extern "C" {

ref_ctx* ref_create(const sipp_map_desc* d) {
  ref_ctx* c = new ref_ctx();
  c->desc = *d;
  sipp_ref::MapSpec& s = c->spec;
  s.horizontal_units = (unsigned)d->W;
  s.vertical_units = (unsigned)d->H;
  s.width = d->width;
  s.height = d->height;
  s.m_per_pixel = d->width / (unsigned)d->W;                         // map_server_planexp.cpp:116
  s.start_position = Eigen::Vector3d(d->start_x, d->start_y, 0.0);   // :120-122
  s.goal_position = Eigen::Vector3d(d->goal_x, d->goal_y, 0.0);
  s.compute_reachability = d->compute_reachability != 0;
  s.compute_closest_observed = d->compute_closest_observed != 0;
  s.compute_path = true;
  s.map_cost.reset(new CostMap_t(d->H, d->W));
  s.map_cost_closest.reset(new CostMap_t(d->H, d->W));
  s.map_state.reset(new OccupancyMap_t(d->H, d->W));
  s.map_reachability.reset(new ReachabilityMap_t(d->H, d->W));
  s.path_map.reset(new Eigen::ArrayXXi(d->H, d->W));
  for (auto& v : s.map_state->a) v = MSR_UNKNOWN;
  return c;
}
void ref_destroy(ref_ctx* c) { delete c; }

// planes are host row-major [y * W + x] (the layout of sipp_map_download); the reference's arrays are column-major (row = y, col = x)
int ref_set_planes(ref_ctx* c, const uint8_t* state, const int32_t* cost, const double* closest, const uint8_t* reach, const uint8_t* path,
                   double mean_cost, double max_cost) {
  const int W = c->desc.W, H = c->desc.H;
  sipp_ref::MapSpec& s = c->spec;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const size_t i = (size_t)y * W + x;
      if (state) (*s.map_state)(y, x) = state[i];
      if (cost) (*s.map_cost)(y, x) = (double)cost[i];
      if (closest) (*s.map_cost_closest)(y, x) = closest[i];
      if (reach) (*s.map_reachability)(y, x) = reach[i] ? MSR_REACHABLE : MSR_REACHABILITY_UNKNOWN;
      if (path) (*s.path_map)(y, x) = path[i];
    }
  s.mean_cost = mean_cost;
  s.max_cost = max_cost;
  return 0;
}

// MapPlanExp::getVoxelArea through CameraModelPlanExp (rows A1/A2): the voxel list of one view, x y z triples
int ref_visible_voxels(ref_ctx* c, int half_fov, double x, double y, double* out_xyz, int cap) {
  sipp_eval_params p;
  memset(&p, 0, sizeof(p));
  p.evaluator = SIPP_EVAL_GOAL_DISTANCE;
  p.half_fov = half_fov;
  p.max_gain = 100.0;
  auto pl = make_planner(c, &p, nullptr, nullptr, 0.0);
  if (!pl) return -1;
  std::vector<Eigen::Vector3d> v;
  dynamic_cast<map::MapPlanExp*>(pl->map.get())->getVoxelArea(&v, Eigen::Vector3d(x, y, 0.0), half_fov);
  for (size_t i = 0; i < v.size() && (int)i < cap; ++i) { out_xyz[3 * i] = v[i][0]; out_xyz[3 * i + 1] = v[i][1]; out_xyz[3 * i + 2] = v[i][2]; }
  return (int)v.size();
}

// SimulatedSensorEvaluator::computeGain for n independent segments (rows A1-A8). Segment i's trajectory = `npts` points: start_xy[i]
// first (trajectory[0], only read by RRTStarEvaluator binary + discount), then — when traj_xy is given — npts - 2 intermediate points,
// then xy[i] (the view). times_ns (optional, npts entries, shared by all segments) feed sampleViewpoints when sampling_time > 0.
// n_unknown[i] = voxels left after the evaluator erased the observed ones; terminal[i] = gain_terminal.
static int gain_range(RefPlanner* pl, int lo, int hi, const double* xy, const double* start_xy, double* gain, uint32_t* n_unknown, uint8_t* terminal,
                      int32_t npts, const double* traj_xy, const int64_t* times_ns);

// threads > 1: the reference itself is single-threaded (one planner loop); for the "all host cores" baseline every worker thread gets
// its own planner + map + evaluator instances (module state such as the local path-map copy is per instance) over the shared planes.
int ref_gain_batch_mt(ref_ctx* c, const sipp_eval_params* p, int32_t n, const double* xy, const double* start_xy, double* gain, uint32_t* n_unknown,
                      uint8_t* terminal, int threads) {
  if (threads < 1) threads = 1;
  std::vector<std::unique_ptr<RefPlanner>> pls;
  for (int t = 0; t < threads; ++t) {           // module construction is not thread safe (factory tables, the map-spec hand-over): do it here
    pls.push_back(make_planner(c, p, nullptr, nullptr, 0.0));
    if (!pls.back()) return SIPP_ERR_UNSUPPORTED;
  }
  std::vector<std::thread> th;
  const int per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const int lo = t * per, hi = std::min(n, lo + per);
    if (lo < hi) th.emplace_back(gain_range, pls[t].get(), lo, hi, xy, start_xy, gain, n_unknown, terminal, 0, nullptr, nullptr);
  }
  for (auto& t : th) t.join();
  return 0;
}

int ref_gain_batch(ref_ctx* c, const sipp_eval_params* p, int32_t n, const double* xy, const double* start_xy, double* gain, uint32_t* n_unknown,
                   uint8_t* terminal, double sampling_time, int32_t npts, const double* traj_xy, const int64_t* times_ns) {
  auto pl = make_planner(c, p, nullptr, nullptr, sampling_time);
  if (!pl) return SIPP_ERR_UNSUPPORTED;
  return gain_range(pl.get(), 0, n, xy, start_xy, gain, n_unknown, terminal, npts, traj_xy, times_ns);
}

static int gain_range(RefPlanner* pl, int lo, int hi, const double* xy, const double* start_xy, double* gain, uint32_t* n_unknown, uint8_t* terminal,
                      int32_t npts, const double* traj_xy, const int64_t* times_ns) {
  for (int i = lo; i < hi; ++i) {
    TrajectorySegment seg;
    if (traj_xy && npts >= 2) {
      seg.trajectory.resize(npts);
      for (int k = 0; k < npts; ++k) {
        seg.trajectory[k].position_W = Eigen::Vector3d(traj_xy[((size_t)i * npts + k) * 2], traj_xy[((size_t)i * npts + k) * 2 + 1], 0.0);
        seg.trajectory[k].time_from_start_ns = times_ns ? times_ns[k] : 0;
      }
    } else {
      seg.trajectory.resize(2);
      seg.trajectory[0].position_W = Eigen::Vector3d(start_xy ? start_xy[2 * i] : xy[2 * i], start_xy ? start_xy[2 * i + 1] : xy[2 * i + 1], 0.0);
      seg.trajectory[1].position_W = Eigen::Vector3d(xy[2 * i], xy[2 * i + 1], 0.0);
    }
    pl->evaluator->computeGain(&seg);
    gain[i] = seg.gain;
    if (terminal) terminal[i] = seg.gain_terminal ? 1 : 0;
    if (n_unknown) n_unknown[i] = seg.info ? (uint32_t) reinterpret_cast<SimulatedSensorInfo*>(seg.info.get())->visible_voxels.size() : 0u;
  }
  return 0;
}

// One re-evaluation pass over a tree given in pre-order (parent[i] < i, children in index order): what the patched
// OnlinePlanner::requestNextTrajectory does — trajectory_evaluator_->updateSegment(root) (mav_active_3d_planning.patch:56), then
// [upstream] updateEvaluatorStep: updateSegment(child) for every node below the root, parents first.
// path_changed selects whether RRTStarUpdater sees a new path id on its root call (rrt_star_updater.cpp:30-38).
int ref_update_tree(ref_ctx* c, const sipp_eval_params* p, const sipp_updater_params* u, int32_t n, const int32_t* parent, const double* end_xy,
                    const double* start_xy, double* gain, double* cost, double* value, uint8_t* terminal, const double* cur_xy,
                    int path_changed, const sipp_cost_params* cp) {
  c->spec.path_counter = path_changed ? 1 : 0;          // RRTStarUpdater starts with last_rrt_path_id_ = 0
  // cp != NULL: the evaluator's cost computer is the reference's CostPlanExpMapIntegral (update_cost then re-integrates every followed
  // segment's 2-point trajectory and `cost` comes back updated); NULL with update_cost set would need [upstream] SegmentTime
  auto pl = make_planner(c, p, u, cp, 0.0);
  if (!pl) return SIPP_ERR_UNSUPPORTED;
  pl->position = Eigen::Vector3d(cur_xy[0], cur_xy[1], 0.0);
  std::vector<TrajectorySegment*> node(n, nullptr);
  std::unique_ptr<TrajectorySegment> root(new TrajectorySegment());
  node[0] = root.get();
  for (int i = 1; i < n; ++i) node[i] = node[parent[i]]->spawnChild();
  for (int i = 0; i < n; ++i) {
    TrajectorySegment* s = node[i];
    s->trajectory.resize(2);
    s->trajectory[0].position_W = Eigen::Vector3d(start_xy ? start_xy[2 * i] : end_xy[2 * i], start_xy ? start_xy[2 * i + 1] : end_xy[2 * i + 1], 0.0);
    s->trajectory[1].position_W = Eigen::Vector3d(end_xy[2 * i], end_xy[2 * i + 1], 0.0);
    s->gain = gain[i];
    s->cost = cost[i];
    s->value = value[i];
    s->gain_terminal = terminal[i] != 0;
  }
  root->gain = 0.0; root->cost = 0.0; root->value = 0.0;     // [upstream] requestNextTrajectory zeroes the new current segment
  // RRTStarUpdater's root visit compares its last path id (0 in a fresh module) with the map's: path_counter 0 -> "unchanged"
  pl->evaluator->updateSegment(root.get());
  // updateEvaluatorStep(segment): for each child { updateSegment(child); updateEvaluatorStep(child); }  == pre-order below the root
  for (int i = 1; i < n; ++i) pl->evaluator->updateSegment(node[i]);
  for (int i = 0; i < n; ++i) {
    gain[i] = node[i]->gain;
    cost[i] = node[i]->cost;
    value[i] = node[i]->value;
    terminal[i] = node[i]->gain_terminal ? 1 : 0;
  }
  return 0;
}

// CostPlanExpMapIntegral::computeCost for n segments; segment i owns points [off[i], off[i+1]) of pts (x y z); parent_cost (optional)
// becomes the parent segment's cost (accumulate mode)
int ref_cost_integral(ref_ctx* c, const sipp_cost_params* cp, int32_t n, const int32_t* off, const double* pts, const double* parent_cost,
                      double* cost, uint8_t* valid) {
  sipp_eval_params p;
  memset(&p, 0, sizeof(p));
  p.evaluator = SIPP_EVAL_GOAL_DISTANCE;
  p.half_fov = 1;
  p.max_gain = 100.0;
  auto pl = make_planner(c, &p, nullptr, cp, 0.0);
  if (!pl) return SIPP_ERR_UNSUPPORTED;
  for (int i = 0; i < n; ++i) {
    TrajectorySegment par, seg;
    if (parent_cost) { par.cost = parent_cost[i]; seg.parent = &par; }
    const int np = off[i + 1] - off[i];
    seg.trajectory.resize(np);
    for (int k = 0; k < np; ++k) seg.trajectory[k].position_W = Eigen::Vector3d(pts[3 * (size_t)(off[i] + k)], pts[3 * (size_t)(off[i] + k) + 1], pts[3 * (size_t)(off[i] + k) + 2]);
    const bool ok = pl->evaluator->computeCost(&seg);
    cost[i] = seg.cost;
    if (valid) valid[i] = ok ? 1 : 0;
  }
  return 0;
}

// which of the reference's module type strings were registered by the compiled sources (registration objects ran)
int ref_has_module(const char* type) { return ModuleFactoryRegistry::Instance().has(type) ? 1 : 0; }

}  // extern "C"
