// Host-side C++ mirror of the scouting-ipp modules over the C ABI (see sipp_modules.h). No arithmetic of the hot path lives
// here: gains, fields, paths, masks, values and the argmax all come from libsipp.so.
#include "sipp_modules.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace {

std::vector<int> parseIds(const std::string& s) {   // "[0, 3599]" / "0,3599"
  std::vector<int> out;
  std::string t = s;
  for (char& c : t)
    if (c == '[' || c == ']' || c == ',') c = ' ';
  std::istringstream is(t);
  int v;
  while (is >> v) out.push_back(v);
  return out;
}

// patches arrive as Eigen::ArrayXXi (column-major in the reference); the C ABI takes the row-major MapData wire layout
std::vector<int32_t> rowMajor(const Eigen::ArrayXXi& a) {
  std::vector<int32_t> out((size_t)a.rows() * a.cols());
  for (int i = 0; i < a.rows(); ++i)
    for (int j = 0; j < a.cols(); ++j) out[(size_t)i * a.cols() + j] = a(i, j);
  return out;
}

}  // namespace

namespace active_3d_planning {
namespace map {

ModuleFactoryRegistry::Registration<GpuMapPlanExp> GpuMapPlanExp::registration("GpuMapPlanExp");

GpuMapPlanExp::GpuMapPlanExp(PlannerI& planner) : Map(planner) {}
GpuMapPlanExp::~GpuMapPlanExp() {
  if (ctx_) sipp_destroy(ctx_);
}

// keys: the map cfg YAML of MapServerPlanExp (map_server_planexp.cpp:76-91) flattened into the param map, plus the map/* keys
// MapPlanExp forwards (map_planexp.cpp:25-31)
void GpuMapPlanExp::setupFromParamMap(Module::ParamMap* param_map) {
  sipp_map_desc& d = desc_;
  d = sipp_map_desc();
  int W, H;
  setParam<int>(param_map, "horizontal_units", &W, 640);
  setParam<int>(param_map, "vertical_units", &H, 480);
  d.W = W; d.H = H;
  setParam<double>(param_map, "width", &d.width, 40.0);
  setParam<double>(param_map, "height", &d.height, 30.0);
  setParam<double>(param_map, "start_x", &d.start_x, 0.0);
  setParam<double>(param_map, "start_y", &d.start_y, 0.0);
  setParam<double>(param_map, "goal_x", &d.goal_x, 39.5);
  setParam<double>(param_map, "goal_y", &d.goal_y, 29.5);
  setParam<int>(param_map, "unknown_id", &d.unknown_id, 3599);
  std::string ids_mav, ids_rov, present, init_str;
  setParam<std::string>(param_map, "obstacle_ids_mav", &ids_mav, "0");
  setParam<std::string>(param_map, "obstacle_ids_rov", &ids_rov, "0");
  setParam<std::string>(param_map, "present_ids", &present, "");
  setParam<std::string>(param_map, "unknown_init_cost", &init_str, "min");
  bool ignore_mav = false;
  setParam<bool>(param_map, "ignore_mav_obstacles", &ignore_mav, false);
  std::vector<int> mav = parseIds(ids_mav), rov = parseIds(ids_rov), pres = parseIds(present);
  if (ignore_mav) mav.clear();                                            // map_server_planexp.cpp:159
  d.n_obstacle_ids_mav = (int)std::min<size_t>(mav.size(), SIPP_MAX_IDS);
  for (int i = 0; i < d.n_obstacle_ids_mav; ++i) d.obstacle_ids_mav[i] = mav[i];
  d.n_obstacle_ids_rov = (int)std::min<size_t>(rov.size(), SIPP_MAX_IDS);
  for (int i = 0; i < d.n_obstacle_ids_rov; ++i) d.obstacle_ids_rov[i] = rov[i];
  // min / max rover cost over the present non-obstacle ids (:122-127); defaults of the no-config branch (:149-150)
  double mn = 3599.0, mx = 0.0;
  bool any = false;
  for (int id : pres) {
    bool obst = false;
    for (int r : rov) obst |= (r == id);
    if (obst) continue;
    any = true;
    if (id < mn) mn = id;
    if (id > mx) mx = id;
  }
  d.min_rov_cost = any ? mn : 1.0;
  d.max_rov_cost = any ? mx : 3599.0;
  if (init_str == "min") d.unknown_init_cost = d.min_rov_cost;            // :165-179
  else if (init_str == "max") d.unknown_init_cost = d.max_rov_cost;
  else {
    char* end = nullptr;
    const double v = std::strtod(init_str.c_str(), &end);
    d.unknown_init_cost = (end && end != init_str.c_str()) ? v : d.min_rov_cost;
  }
  bool reach, closest;
  setParam<bool>(param_map, "compute_reachability", &reach, true);
  setParam<bool>(param_map, "compute_closest_observed", &closest, false);
  setParam<bool>(param_map, "compute_path", &compute_path_, false);
  setParam<int>(param_map, "device", &device_, 0);
  d.compute_reachability = reach;
  d.compute_closest_observed = closest;
  path_requested_ = compute_path_;                                          // :162
  if (ctx_) { sipp_destroy(ctx_); ctx_ = nullptr; }
  if (sipp_create(&d, device_, &ctx_) != SIPP_OK) {
    error_ = sipp_last_error(nullptr);
    planner_.printError("'GpuMapPlanExp': " + error_);
  }
}

bool GpuMapPlanExp::checkParamsValid(std::string* error_message) {
  if (!ctx_) {
    *error_message = error_.empty() ? "no sipp context" : error_;
    return false;
  }
  return true;
}

bool GpuMapPlanExp::isObserved(const Eigen::Vector3d& point) {
  if (!ctx_) return false;
  if (!(0 <= point[0] && 0 <= point[1] && desc_.width > point[0] && desc_.height > point[1])) return false;   // :764-769 -> UNKNOWN
  if (state_dirty_) {
    state_cache_.resize((size_t)desc_.W * desc_.H);
    if (sipp_map_download(ctx_, state_cache_.data(), nullptr, nullptr, nullptr) != SIPP_OK) return false;
    state_dirty_ = false;
  }
  const double mpp = desc_.width / (unsigned int)desc_.W;
  const int r = (int)(point[1] / mpp), c = (int)(point[0] / mpp);
  return state_cache_[(size_t)r * desc_.W + c] != SIPP_STATE_UNKNOWN;
}

void GpuMapPlanExp::targetReachedCallback() { path_requested_ = true; }

bool GpuMapPlanExp::mapDataCallback(const Eigen::Vector3d& pose, const Eigen::ArrayXXi& labels) {
  if (!ctx_) return false;
  int32_t x0, y0;
  if (sipp_patch_origin(ctx_, pose[0], pose[1], labels.rows(), labels.cols(), &x0, &y0) != SIPP_OK) return false;
  if (sipp_map_update_patch(ctx_, x0, y0, labels.rows(), labels.cols(), rowMajor(labels).data()) != SIPP_OK) {
    planner_.printError(std::string("'GpuMapPlanExp': ") + sipp_last_error(ctx_));
    return false;
  }
  state_dirty_ = true;
  if (compute_path_ && path_requested_) {                                   // :278-281 (no background thread: the GPU plan takes ms)
    path_requested_ = false;
    return computeCurrentPathEstimate();
  }
  return true;
}

bool GpuMapPlanExp::computeCurrentPathEstimate() {
  if (!ctx_) return false;
  int32_t status = 0, len = 0;
  std::vector<double> xy(2 * 4 * (size_t)(desc_.W + desc_.H));
  if (sipp_plan(ctx_, &last_cost_, &status, xy.data(), (int32_t)(xy.size() / 2), &len) != SIPP_OK) {
    planner_.printError(std::string("'GpuMapPlanExp': ") + sipp_last_error(ctx_));
    return false;
  }
  if ((size_t)len > xy.size() / 2) {     // longer than 4 (W + H) nodes (a path winding through a maze): fetch all of it
    xy.resize(2 * (size_t)len);
    if (sipp_path_download(ctx_, xy.data(), len, &len) != SIPP_OK) {
      planner_.printError(std::string("'GpuMapPlanExp': ") + sipp_last_error(ctx_));
      return false;
    }
  }
  last_status_ = status;
  last_path_.clear();
  for (int i = 0; i < len && (size_t)i < xy.size() / 2; ++i) last_path_.push_back(Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]));
  return true;
}

int GpuMapPlanExp::getPathMapId() const { return ctx_ ? sipp_path_mask_generation(ctx_) : 0; }

double GpuMapPlanExp::getMeanWeight() const {
  double m = 0.0;
  if (ctx_) sipp_map_stats(ctx_, &m, nullptr, nullptr);
  return m;
}

}  // namespace map

namespace trajectory_evaluator {

void ViewSampler::indices(const EigenTrajectoryPointVector& trajectory, std::vector<int>* out) const {
  out->clear();
  if (trajectory.empty()) return;                                            // sensor_model_planexp.cpp:67-70
  const int n = (int)trajectory.size();
  if (sampling_time <= 0.0) { out->push_back(n - 1); return; }               // :71-73: rate 0 means the last point only
  const int64_t sampling_time_ns = static_cast<int64_t>(sampling_time * 1.0e9);      // :75
  if (sampling_time_ns >= trajectory.back().time_from_start_ns) { out->push_back(n - 1); return; }   // :76-78
  int64_t current_time = sampling_time_ns;                                   // :81-87
  for (int i = 0; i < n; ++i)
    if (trajectory[i].time_from_start_ns >= current_time) {
      current_time += sampling_time_ns;
      out->push_back(i);
    }
}

void ViewSampler::appendViews(const EigenTrajectoryPointVector& trajectory, std::vector<double>* view_xy) const {
  std::vector<int> idx;
  indices(trajectory, &idx);
  for (int i : idx) {
    const auto& p = trajectory[i];
    // position = position + orientation * mounting_translation (:28-30); the mounting rotation only turns the sensor, and the planar
    // footprint of CameraModelPlanExp (camera_model_planexp.cpp:23-28) does not depend on where it looks
    const Eigen::Vector3d t = p.orientation_W_B * mounting_translation;
    view_xy->push_back(p.position_W[0] + t[0]);
    view_xy->push_back(p.position_W[1] + t[1]);
  }
}

void FlatTree::build(TrajectorySegment* root, const ViewSampler* sampler) {
  nodes.clear(); parent.clear(); end_xy.clear(); start_xy.clear(); gain.clear(); cost.clear(); value.clear(); terminal.clear();
  view_off.clear(); view_xy.clear();
  if (sampler) view_off.push_back(0);
  // pre-order, children in vector order — the order of [upstream] OnlinePlanner::updateEvaluatorStep
  std::vector<std::pair<TrajectorySegment*, int>> stack;
  stack.push_back({root, -1});
  while (!stack.empty()) {
    auto cur = stack.back();
    stack.pop_back();
    const int id = (int)nodes.size();
    TrajectorySegment* s = cur.first;
    nodes.push_back(s);
    parent.push_back(cur.second);
    const bool has = !s->trajectory.empty();
    end_xy.push_back(has ? s->trajectory.back().position_W[0] : 0.0);     // sampling_time <= 0: last point only (sensor_model_planexp.cpp:72-75)
    end_xy.push_back(has ? s->trajectory.back().position_W[1] : 0.0);
    start_xy.push_back(has ? s->trajectory.front().position_W[0] : 0.0);  // trajectory[0] (rrt_star_evaluator.cpp:65)
    start_xy.push_back(has ? s->trajectory.front().position_W[1] : 0.0);
    gain.push_back(s->gain);
    cost.push_back(s->cost);
    value.push_back(s->value);
    terminal.push_back(s->gain_terminal ? 1 : 0);
    if (sampler) {
      sampler->appendViews(s->trajectory, &view_xy);
      view_off.push_back((int32_t)(view_xy.size() / 2));
    }
    for (int k = (int)s->children.size() - 1; k >= 0; --k) stack.push_back({s->children[k].get(), id});
  }
}

void updaterMasks(const sipp_updater_params& u, const FlatTree& t, const double cur_xy[2], bool path_changed, std::vector<uint8_t>* do_gain,
                  std::vector<uint8_t>* maybe_value, bool followed_instead_of_value) {
  const int n = (int)t.nodes.size();
  do_gain->assign(n, 0);
  maybe_value->assign(n, 0);
  // node 0 is the root: the patched planner's root-first call (mav_active_3d_planning.patch:56) runs the chain on it as well. An
  // RRTStarUpdater only compares path ids there (rrt_star_updater.cpp:21,30-38); the other chains treat it like any node, with the
  // gain = 0 the planner gave it when it became the current segment.
  for (int i = (u.updater == SIPP_UPD_RRT_STAR ? 1 : 0); i < n; ++i) {
    const double dx = cur_xy[0] - t.end_xy[2 * i], dy = cur_xy[1] - t.end_xy[2 * i + 1];
    const double dist = std::sqrt(dx * dx + dy * dy + 0.0 * 0.0);
    const double gi = i == 0 ? 0.0 : t.gain[i];
    bool follow;
    if (u.updater == SIPP_UPD_RRT_STAR) {
      if (t.terminal[i]) continue;                                        // :22
      follow = (gi > u.minimum_gain || path_changed) && (dist <= u.update_range || path_changed);   // :24-28
    } else if (u.updater == SIPP_UPD_CONSTRAINED) {
      follow = gi > u.minimum_gain && dist <= u.update_range;             // [upstream] ConstrainedUpdater
    } else {
      follow = true;                                                      // following updater used directly
    }
    if (!follow) continue;
    if (u.following == SIPP_FOLLOW_UPDATE_ALL) (*do_gain)[i] = u.update_gain ? 1 : 0;
    else (*do_gain)[i] = (u.update_gain && !t.terminal[i]) ? 1 : 0;       // update_all_non_terminal.cpp:17-19
    (*maybe_value)[i] = (followed_instead_of_value || u.update_value) ? 1 : 0;   // NonTerminal re-checks gain_terminal after the gain update (:23)
  }
}

GpuGainEvaluator::GpuGainEvaluator(PlannerI& planner, int evaluator) : TrajectoryEvaluator(planner) {
  params_ = sipp_eval_params();
  params_.evaluator = evaluator;
  updater_ = sipp_updater_params();
}

void GpuGainEvaluator::setupFromParamMap(Module::ParamMap* param_map) {
  // evaluator keys and defaults of the reference (rrt_star_evaluator.cpp:22-25, goal_distance_evaluator.cpp:22,
  // expand_reachable_area_evaluator.cpp:22, expand_low_cost_area_evaluator.cpp:22) and camera_model_planexp.cpp:20
  bool b;
  setParam<double>(param_map, "non_path_voxel_gain", &params_.non_path_voxel_gain, 0.0);
  setParam<bool>(param_map, "use_distance_discount", &b, false);
  params_.use_distance_discount = b;
  setParam<double>(param_map, "discount_offset", &params_.discount_offset, 0.1);
  setParam<bool>(param_map, "do_binary_check", &b, false);
  params_.do_binary_check = b;
  setParam<double>(param_map, "max_gain", &params_.max_gain, 100.0);
  setParam<double>(param_map, "discount_factor", &params_.discount_factor, 0.1);
  setParam<bool>(param_map, "augment_unknown_with_mean", &b, false);
  params_.augment_unknown_with_mean = b;
  int hf;
  setParam<int>(param_map, "sensor_model/half_fov", &hf, 10);
  params_.half_fov = hf;
  setParam<double>(param_map, "sensor_model/sampling_time", &sampler_.sampling_time, 0.0);          // sensor_model_planexp.cpp:170
  {
    double m[3];
    const char* keys[3] = {"mounting_translation_x", "mounting_translation_y", "mounting_translation_z"};   // [upstream] SensorModel::setupFromParamMap
    for (int k = 0; k < 3; ++k) setParam<double>(param_map, std::string("sensor_model/") + keys[k], &m[k], 0.0);
    sampler_.mounting_translation = Eigen::Vector3d(m[0], m[1], m[2]);
    // sensor_model/mounting_rotation_*: turns the sensor about its own position; the planar footprint does not depend on it
  }
  // [upstream] bounding volume of SimulatedSensorEvaluator (cfg/experiments/example.yaml:12-18)
  setParam<bool>(param_map, "bounding_volume/enabled", &b, false);
  params_.use_bounding_volume = b;
  setParam<double>(param_map, "bounding_volume/x_min", &params_.bv_x_min, 0.0);
  setParam<double>(param_map, "bounding_volume/x_max", &params_.bv_x_max, 0.0);
  setParam<double>(param_map, "bounding_volume/y_min", &params_.bv_y_min, 0.0);
  setParam<double>(param_map, "bounding_volume/y_max", &params_.bv_y_max, 0.0);
  // cost_computer block (cfg/evaluators/baseline.yaml:14-15): [upstream] "SegmentTime" or the map line integral
  std::string ct;
  setParam<std::string>(param_map, "cost_computer/type", &ct, "SegmentTime");
  cost_is_integral_ = (ct == "GpuCostPlanExpMapIntegral" || ct == "CostPlanExpMapIntegral");
  if (!cost_is_integral_ && ct != "SegmentTime") config_error_ = "unknown cost_computer type '" + ct + "'";
  setParam<bool>(param_map, "cost_computer/accumulate", &b, false);                                    // cost_planexp_map_integral.cpp:16-17
  cost_params_.accumulate = b;
  setParam<double>(param_map, "cost_computer/max_sample_distance", &cost_params_.max_sample_distance, 0.05);
  // evaluator_updater block (cfg/evaluators/rrt_star.yaml:24-32, example_config.yaml:109-117)
  std::string ut, ft;
  setParam<std::string>(param_map, "evaluator_updater/type", &ut, "RRTStarUpdater");
  setParam<std::string>(param_map, "evaluator_updater/following_updater/type", &ft, "UpdateAll");
  std::string fns = "evaluator_updater/following_updater/";
  if (ut == "UpdateAll" || ut == "UpdateAllNonTerminal") {                // cfg/evaluators/expand_reachable_area.yaml:23-27
    updater_.updater = SIPP_UPD_DIRECT;
    ft = ut;
    fns = "evaluator_updater/";
  } else {
    updater_.updater = (ut == "ConstrainedUpdater") ? SIPP_UPD_CONSTRAINED : SIPP_UPD_RRT_STAR;
    if (ut != "ConstrainedUpdater" && ut != "RRTStarUpdater") config_error_ = "unknown evaluator_updater type '" + ut + "'";
  }
  if (ft != "UpdateAll" && ft != "UpdateAllNonTerminal") config_error_ = "unknown following_updater type '" + ft + "'";
  updater_.following = (ft == "UpdateAllNonTerminal") ? SIPP_FOLLOW_UPDATE_ALL_NON_TERMINAL : SIPP_FOLLOW_UPDATE_ALL;
  setParam<double>(param_map, "evaluator_updater/minimum_gain", &updater_.minimum_gain, -1e9);   // rrt_star_updater.cpp:45-46
  setParam<double>(param_map, "evaluator_updater/update_range", &updater_.update_range, 1e9);
  setParam<bool>(param_map, fns + "update_gain", &b, true);               // update_all_non_terminal.cpp:30-32
  updater_.update_gain = b;
  setParam<bool>(param_map, fns + "update_cost", &b, true);
  updater_.update_cost = b;
  setParam<bool>(param_map, fns + "update_value", &b, true);
  updater_.update_value = b;
  map_ = dynamic_cast<map::GpuMapPlanExp*>(&(planner_.getMap()));
  if (!map_) planner_.printError("'Gpu*Evaluator' requires a map of type 'GpuMapPlanExp'!");   // rrt_star_evaluator.cpp:28-32
}

bool GpuGainEvaluator::checkParamsValid(std::string* error_message) {
  if (!map_ || !map_->ctx()) { *error_message = "map of type 'GpuMapPlanExp' with a live context expected"; return false; }
  if (!config_error_.empty()) { *error_message = config_error_; return false; }
  if (params_.half_fov < 0 || params_.half_fov > 112) { *error_message = "sensor_model/half_fov expected in [0, 112]"; return false; }
  // sensor_model/sampling_time > 0 (several views per segment, sensor_model_planexp.cpp:65-90) and a mounting translation (:28-31) are
  // served: the host samples the viewpoints, sipp_gain_batch_views / sipp_update_tree_views score them
  return true;
}

bool GpuGainEvaluator::computeGain(TrajectorySegment* traj_in) {
  if (!map_ || traj_in->trajectory.empty()) { traj_in->gain = 0.0; return false; }
  const double xy[2] = {traj_in->trajectory.back().position_W[0], traj_in->trajectory.back().position_W[1]};
  const double sxy[2] = {traj_in->trajectory.front().position_W[0], traj_in->trajectory.front().position_W[1]};
  double gain = 0.0;
  uint8_t term = 0;
  int rc;
  if (multi_view()) {
    std::vector<double> views;
    sampler_.appendViews(traj_in->trajectory, &views);
    const int32_t off[2] = {0, (int32_t)(views.size() / 2)};
    rc = sipp_gain_batch_views(map_->ctx(), &params_, 1, off, views.data(), sxy, &gain, nullptr, &term);
  } else {
    rc = sipp_gain_batch(map_->ctx(), &params_, 1, xy, sxy, &gain, nullptr, &term);
  }
  if (rc != SIPP_OK) {
    planner_.printError(std::string("'Gpu*Evaluator': ") + sipp_last_error(map_->ctx()));
    traj_in->gain = 0.0;
    return false;
  }
  traj_in->gain = gain;
  if (term) traj_in->gain_terminal = true;
  return true;
}

bool GpuGainEvaluator::computeCost(TrajectorySegment* traj_in) {
  if (cost_is_integral_) {   // CostPlanExpMapIntegral::computeCost (cost_planexp_map_integral.cpp:21-73) as a batch of one segment
    if (!map_) return false;
    std::vector<double> pts;
    pts.reserve(traj_in->trajectory.size() * 3);
    for (const auto& p : traj_in->trajectory) { pts.push_back(p.position_W.x()); pts.push_back(p.position_W.y()); pts.push_back(p.position_W.z()); }
    const int32_t off[2] = {0, (int32_t)traj_in->trajectory.size()};
    const double parent_cost = traj_in->parent ? traj_in->parent->cost : 0.0;
    double cost = 0.0;
    uint8_t valid = 0;
    if (sipp_cost_integral(map_->ctx(), &cost_params_, 1, off, pts.data(), traj_in->parent ? &parent_cost : nullptr, &cost, &valid) != SIPP_OK) return false;
    traj_in->cost = cost;
    return valid != 0;
  }
  // [upstream] SegmentTime: duration of the segment in seconds
  if (traj_in->trajectory.empty()) { traj_in->cost = 0.0; return false; }
  traj_in->cost = static_cast<double>(traj_in->trajectory.back().time_from_start_ns) * 1.0e-9;
  return true;
}

bool GpuGainEvaluator::computeValue(TrajectorySegment* traj_in) {
  if (!map_) return false;
  TrajectorySegment* root = traj_in;
  while (root->parent) root = root->parent;
  FlatTree t;
  t.build(root);
  std::vector<double> value(t.nodes.size());
  if (sipp_value_select(map_->ctx(), (int32_t)t.nodes.size(), t.gain.data(), t.cost.data(), t.parent.data(), value.data(), nullptr, nullptr) != SIPP_OK)
    return false;
  for (size_t i = 0; i < t.nodes.size(); ++i)
    if (t.nodes[i] == traj_in) traj_in->value = value[i];
  return true;
}

int GpuGainEvaluator::selectNextBest(TrajectorySegment* traj_in) {
  // [upstream] SubsequentBest: argmax over the children of the subtree-max of the STORED values (ties: lowest index here)
  if (traj_in->children.empty()) return -1;
  FlatTree t;
  t.build(traj_in);
  const int n = (int)t.nodes.size();
  std::vector<double> sub(t.value);
  for (int i = n - 1; i >= 1; --i) sub[t.parent[i]] = (i == 0 || t.parent[i] == 0) ? sub[t.parent[i]] : std::max(sub[t.parent[i]], sub[i]);
  int best = -1, k = 0;
  double bv = 0.0;
  for (int i = 1; i < n; ++i)
    if (t.parent[i] == 0) {
      if (best < 0 || sub[i] > bv) { best = k; bv = sub[i]; }
      ++k;
    }
  return best;
}

bool GpuGainEvaluator::updateSegment(TrajectorySegment* segment) {
  if (segment->parent) return true;          // done in the root-first call of this pass
  if (!map_) return false;
  // root branch of RRTStarUpdater (rrt_star_updater.cpp:31-40)
  if (last_rrt_path_id_ != map_->getPathMapId()) {
    rrt_path_has_changed_ = true;
    map_->updatePathMap();
  } else {
    rrt_path_has_changed_ = false;
  }
  last_rrt_path_id_ = map_->getPathMapId();

  FlatTree t;
  t.build(segment, multi_view() ? &sampler_ : nullptr);
  const int n = (int)t.nodes.size();
  const Eigen::Vector3d cur = planner_.getCurrentPosition();
  const double cur_xy[2] = {cur[0], cur[1]};
  // ONE call: updater predicate on the device, one batched gain launch over the eligible nodes, values with the pre-order
  // semantics of the reference pass (sipp.h: sipp_update_tree). updaterMasks() above is the host statement of the same predicate.
  std::vector<double> gain(t.gain), value(t.value);
  std::vector<uint8_t> terminal(t.terminal);
  int32_t rescored = 0;
  const int path_changed = (updater_.updater == SIPP_UPD_RRT_STAR && rrt_path_has_changed_) ? 1 : 0;
  // update_cost ([upstream] UpdateAll / UpdateAllNonTerminal: computeCost before computeValue, update_all_non_terminal.cpp:20-22).
  // SegmentTime costs do not depend on the map, so recomputing them changes nothing; the map line integral does: the segments the
  // chain follows are re-integrated in ONE batched call, parents first for the accumulate mode, and handed to the pass together with
  // the costs the tree carried before it (what computeValue still finds below the node it visits).
  std::vector<double> cost_new;
  const bool recost = updater_.update_cost && cost_is_integral_;
  if (recost) {
    std::vector<uint8_t> do_gain, followed;
    updaterMasks(updater_, t, cur_xy, path_changed != 0, &do_gain, &followed, true);
    std::vector<int32_t> off(1, 0), ids;
    std::vector<double> pts;
    for (int i = 0; i < n; ++i) {
      if (!followed[i]) continue;
      ids.push_back(i);
      for (const auto& p : t.nodes[i]->trajectory) { pts.push_back(p.position_W.x()); pts.push_back(p.position_W.y()); pts.push_back(p.position_W.z()); }
      off.push_back((int32_t)(pts.size() / 3));
    }
    cost_new = t.cost;
    cost_new[0] = 0.0;
    if (!ids.empty()) {
      std::vector<double> c(ids.size());
      sipp_cost_params cp = cost_params_;
      cp.accumulate = 0;
      if (sipp_cost_integral(map_->ctx(), &cp, (int32_t)ids.size(), off.data(), pts.data(), nullptr, c.data(), nullptr) != SIPP_OK) {
        planner_.printError(std::string("'Gpu*Evaluator': ") + sipp_last_error(map_->ctx()));
        return false;
      }
      for (size_t k = 0; k < ids.size(); ++k) {       // pre-order: a parent's cost of this pass is final before its children are visited
        const int i = ids[k];
        cost_new[i] = c[k];
        if (cost_params_.accumulate && t.parent[i] >= 0) cost_new[i] += cost_new[t.parent[i]];      // cost_planexp_map_integral.cpp:68-70
      }
    }
  }
  const int urc = multi_view()
      ? sipp_update_tree_views(map_->ctx(), &params_, &updater_, n, t.parent.data(), t.end_xy.data(), t.start_xy.data(), gain.data(),
                               recost ? cost_new.data() : t.cost.data(), value.data(), terminal.data(), cur_xy, path_changed, &rescored,
                               recost ? t.cost.data() : nullptr, t.view_off.data(), t.view_xy.data())
      : sipp_update_tree(map_->ctx(), &params_, &updater_, n, t.parent.data(), t.end_xy.data(), t.start_xy.data(), gain.data(),
                         recost ? cost_new.data() : t.cost.data(), value.data(), terminal.data(), cur_xy, path_changed, &rescored,
                         recost ? t.cost.data() : nullptr);
  if (urc != SIPP_OK) {
    planner_.printError(std::string("'Gpu*Evaluator': ") + sipp_last_error(map_->ctx()));
    return false;
  }
  last_batch_ = rescored;
  for (int i = (updater_.updater == SIPP_UPD_RRT_STAR ? 1 : 0); i < n; ++i) {
    t.nodes[i]->gain = gain[i];
    t.nodes[i]->value = value[i];
    if (recost) t.nodes[i]->cost = cost_new[i];
    if (terminal[i]) t.nodes[i]->gain_terminal = true;
  }
  return true;
}

template <> ModuleFactoryRegistry::Registration<GpuEvaluatorT<SIPP_EVAL_RRT_STAR>> GpuEvaluatorT<SIPP_EVAL_RRT_STAR>::registration("GpuRRTStarEvaluator");
template <> ModuleFactoryRegistry::Registration<GpuEvaluatorT<SIPP_EVAL_GOAL_DISTANCE>> GpuEvaluatorT<SIPP_EVAL_GOAL_DISTANCE>::registration("GpuGoalDistanceEvaluator");
template <> ModuleFactoryRegistry::Registration<GpuEvaluatorT<SIPP_EVAL_EXPAND_REACHABLE>> GpuEvaluatorT<SIPP_EVAL_EXPAND_REACHABLE>::registration("GpuExpandReachableAreaEvaluator");
template <> ModuleFactoryRegistry::Registration<GpuEvaluatorT<SIPP_EVAL_AVERAGE_VIEW_COST>> GpuEvaluatorT<SIPP_EVAL_AVERAGE_VIEW_COST>::registration("GpuAverageViewCostEvaluator");
template <> ModuleFactoryRegistry::Registration<GpuEvaluatorT<SIPP_EVAL_EXPAND_LOW_COST>> GpuEvaluatorT<SIPP_EVAL_EXPAND_LOW_COST>::registration("GpuExpandLowCostAreaEvaluator");
template <> ModuleFactoryRegistry::Registration<GpuEvaluatorT<SIPP_EVAL_NAIVE>> GpuEvaluatorT<SIPP_EVAL_NAIVE>::registration("GpuNaiveEvaluator");

}  // namespace trajectory_evaluator

namespace initialize {
void sipp_gpu_package() {}
}  // namespace initialize

}  // namespace active_3d_planning

// ---------------------------------------------------------------------------------------------------------------------
namespace planexp_base_planners {

GpuPRMStarPlanner::GpuPRMStarPlanner(PlannerParams params, int device) : PlannerBase(params), device_(device) {}
GpuPRMStarPlanner::~GpuPRMStarPlanner() {
  if (ctx_) sipp_destroy(ctx_);
}

void GpuPRMStarPlanner::setup() {   // PRMStarPlanner::setup = set_derived_params + build_graph (prm_star_planner.cpp:11-14)
  sipp_map_desc d = sipp_map_desc();
  d.W = params_.px_x; d.H = params_.px_y;
  d.width = params_.width; d.height = params_.height;
  d.start_x = params_.start.x(); d.start_y = params_.start.y();
  d.goal_x = params_.goal.x(); d.goal_y = params_.goal.y();
  d.min_rov_cost = params_.min_cost; d.max_rov_cost = params_.max_cost; d.unknown_init_cost = params_.init_cost;
  d.unknown_id = params_.unknown_id;
  d.n_obstacle_ids_rov = (int)std::min<size_t>(params_.obstacle_ids.size(), SIPP_MAX_IDS);
  for (int i = 0; i < d.n_obstacle_ids_rov; ++i) d.obstacle_ids_rov[i] = params_.obstacle_ids[i];
  d.n_obstacle_ids_mav = 0;
  if (ctx_) { sipp_destroy(ctx_); ctx_ = nullptr; }
  if (sipp_create(&d, device_, &ctx_) != SIPP_OK) error_ = sipp_last_error(nullptr);
}

ResultStatus GpuPRMStarPlanner::plan(std::vector<Eigen::Vector2d>* path, double& cost) {
  cost = -1.0;                                                             // prm_star_planner.cpp:39
  if (!ctx_) return ResultStatus::INVALID;                                 // the reference swallows every failure into INVALID (:66-68)
  int32_t status = 0, len = 0;
  std::vector<double> xy(2 * 4 * (size_t)(params_.px_x + params_.px_y));
  if (sipp_plan(ctx_, &cost, &status, xy.data(), (int32_t)(xy.size() / 2), &len) != SIPP_OK) {
    error_ = sipp_last_error(ctx_);
    cost = -1.0;
    return ResultStatus::INVALID;
  }
  if ((size_t)len > xy.size() / 2) {     // a path that winds through a maze can be longer than 4 (W + H) nodes: fetch all of it
    xy.resize(2 * (size_t)len);
    if (sipp_path_download(ctx_, xy.data(), len, &len) != SIPP_OK) {
      error_ = sipp_last_error(ctx_);
      cost = -1.0;
      return ResultStatus::INVALID;
    }
  }
  current_best_path_.clear();
  for (int i = 0; i < len && (size_t)i < xy.size() / 2; ++i) current_best_path_.push_back(Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]));
  if (path) for (auto& p : current_best_path_) path->push_back(p);
  switch (status) {
    case SIPP_PLAN_VALID: return ResultStatus::VALID;
    case SIPP_PLAN_TERMINAL_VALID: return ResultStatus::TERMINAL_VALID;
    case SIPP_PLAN_INVALID: return ResultStatus::INVALID;
    default: return ResultStatus::TERMINAL_INVALID;
  }
}

void GpuPRMStarPlanner::store_path(const std::string file_name, const std::string path) {   // prm_star_planner.cpp:72-102
  if (path.empty()) {
    std::printf("[GpuPRMStarPlanner] store_path was requested but not setup properly! (%s/%s)\n", path.c_str(), file_name.c_str());
    return;
  }
  std::ofstream f(path + "/" + file_name + ".csv", std::ios::out | std::ios::trunc);
  if (!f) {
    std::printf("[GpuPRMStarPlanner] store_path was requested but not setup properly! (%s/%s)\n", path.c_str(), file_name.c_str());
    return;
  }
  f << "PositionX" << "," << "PositionY" << std::endl;
  f << "m" << "," << "m" << std::endl;
  for (auto& p : current_best_path_) f << p[0] << "," << p[1] << std::endl;
}

void GpuPRMStarPlanner::update_states(Eigen::Vector2i location, Eigen::ArrayXXi map_data) {
  if (!ctx_) return;
  if (sipp_map_update_patch(ctx_, location.x(), location.y(), map_data.rows(), map_data.cols(), rowMajor(map_data).data()) != SIPP_OK)
    error_ = sipp_last_error(ctx_);
}

void GpuPRMStarPlanner::update_unknown_cost(double new_cost) {
  std::printf("[GpuPRMStarPlanner] Updating unknown cost to: %f\n", new_cost);   // prm_star_planner.cpp:122
  if (ctx_) sipp_update_unknown_cost(ctx_, new_cost);
}

}  // namespace planexp_base_planners
