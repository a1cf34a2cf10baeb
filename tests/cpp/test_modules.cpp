// Test of the C++ module mirror (scouting-ipp_b200/plugin) — reads like a test of the reference's own modules would:
// modules are created by `type:` string through the factory registry, fed MapData patches, asked to plan, and the
// root-first updateSegment pass is compared node by node with the oracle's literal pre-order emulation of the reference's
// updater chain (oracle/sipp_oracle.cpp: orc_update_tree). The oracle is linked here as the CHECKER only.
//
//   test_modules cpu   registry / param parsing / tree flattening / updater predicate / error path without a device
//   test_modules gpu   full parity on cuda:0 (map ingest, plan, per-node gain, batched update pass, selection)
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../scouting-ipp_b200/plugin/sipp_modules.h"

extern "C" {
struct orc_ctx;
orc_ctx* orc_create(const sipp_map_desc* d);
void orc_destroy(orc_ctx* c);
int orc_patch_origin(orc_ctx* c, double px, double py, int rows, int cols, int* x0, int* y0);
int orc_map_update_patch(orc_ctx* c, int x0, int y0, int rows, int cols, const int32_t* labels);
int orc_map_stats(orc_ctx* c, double* mean_cost, double* max_cost, int64_t* mean_counter);
int orc_plan(orc_ctx* c, double* cost, int32_t* status, double* path_xy, int32_t path_cap, int32_t* path_len, int update_mask);
int orc_path_mask_generation(orc_ctx* c);
int orc_gain_batch(orc_ctx* c, const sipp_eval_params* p, int32_t n, const double* xy, const double* start_xy, double* gain,
                   uint32_t* counts, uint8_t* terminal, int variant, int threads);
int orc_cost_integral(orc_ctx* c, const sipp_cost_params* cp, int32_t n, const int32_t* off, const double* pts, const double* parent_cost,
                      double* cost_out, uint8_t* valid);
int orc_value_select(int32_t n, const double* gain_in, const double* cost_in, const int32_t* parent, double* value,
                     int32_t* best_child, double* best_value);
int orc_update_tree(orc_ctx* c, const sipp_eval_params* p, const sipp_updater_params* u, int32_t n, const int32_t* parent,
                    const double* end_xy, const double* start_xy, double* gain, const double* cost, double* value,
                    uint8_t* terminal, const double* cur_xy, int path_changed, int variant, const double* cost_below);
int orc_sample_viewpoints(const int64_t* times_ns, int32_t npts, double sampling_time, int32_t* out_idx, int32_t cap);
int orc_gain_batch_views(orc_ctx* c, const sipp_eval_params* p, int32_t n, const int32_t* view_off, const double* view_xy, const double* start_xy,
                         double* gain, uint32_t* counts, uint8_t* terminal, int variant);
int orc_update_tree_views(orc_ctx* c, const sipp_eval_params* p, const sipp_updater_params* u, int32_t n, const int32_t* parent,
                          const double* end_xy, const double* start_xy, double* gain, const double* cost, double* value, uint8_t* terminal,
                          const double* cur_xy, int path_changed, int variant, const double* cost_below, const int32_t* view_off, const double* view_xy);
}

using namespace active_3d_planning;

static int g_fail = 0;
#define CHECK(cond, ...)                                                      \
  do {                                                                        \
    if (!(cond)) {                                                            \
      std::printf("FAIL %s:%d  %s  ", __FILE__, __LINE__, #cond);             \
      std::printf(__VA_ARGS__);                                               \
      std::printf("\n");                                                      \
      ++g_fail;                                                               \
    }                                                                         \
  } while (0)

struct FakePlanner : public PlannerI {
  std::unique_ptr<Map> map;
  std::unique_ptr<TrajectoryEvaluator> evaluator;
  Eigen::Vector3d position;
  std::vector<std::string> errors;
  Map& getMap() override { return *map; }
  TrajectoryEvaluator& getTrajectoryEvaluator() override { return *evaluator; }
  const Eigen::Vector3d& getCurrentPosition() const override { return position; }
  void printError(const std::string& text) override { errors.push_back(text); }
};

static Module::ParamMap mapParams(int W, int H) {
  Module::ParamMap p;
  p["type"] = "GpuMapPlanExp";
  p["horizontal_units"] = std::to_string(W);
  p["vertical_units"] = std::to_string(H);
  p["width"] = std::to_string(W * 0.0625);
  p["height"] = std::to_string(H * 0.0625);
  p["start_x"] = "0.53125"; p["start_y"] = "0.53125";
  p["goal_x"] = std::to_string((W - 8) * 0.0625 + 0.03125);
  p["goal_y"] = std::to_string((H - 8) * 0.0625 + 0.03125);
  p["unknown_id"] = "3599";
  p["obstacle_ids_mav"] = "[0]";
  p["obstacle_ids_rov"] = "[0]";
  p["present_ids"] = "[0, 30, 45, 60, 90, 120]";
  p["unknown_init_cost"] = "max";
  p["compute_path"] = "true";
  p["compute_closest_observed"] = "true";
  return p;
}

static Module::ParamMap evalParams(const char* type, const char* updater, const char* following) {
  Module::ParamMap p;
  p["type"] = type;
  p["sensor_model/half_fov"] = "32";
  p["non_path_voxel_gain"] = "0.001";
  p["evaluator_updater/type"] = updater;
  if (std::strcmp(updater, "UpdateAll") && std::strcmp(updater, "UpdateAllNonTerminal")) {
    p["evaluator_updater/minimum_gain"] = "0.001";
    p["evaluator_updater/update_range"] = "4.0";
    p["evaluator_updater/following_updater/type"] = following;
    p["evaluator_updater/following_updater/update_gain"] = "true";
    p["evaluator_updater/following_updater/update_cost"] = "false";
    p["evaluator_updater/following_updater/update_value"] = "true";
  } else {
    p["evaluator_updater/update_gain"] = "true";
    p["evaluator_updater/update_cost"] = "false";
    p["evaluator_updater/update_value"] = "true";
  }
  return p;
}

// label image in the style of helper_scripts/create_artificial_maps.py:70-86 (rectangles over a default label)
static std::vector<int32_t> makeLabels(int W, int H, unsigned seed) {
  std::mt19937 rng(seed);
  std::vector<int32_t> L((size_t)W * H, 30);
  const int labels[5] = {45, 60, 90, 120, 0};
  for (int r = 0; r < 24; ++r) {
    const int lab = labels[r % 5];
    const int w = (lab ? W / 10 : W / 24) + (int)(rng() % (W / 8)), h = (lab ? H / 10 : H / 24) + (int)(rng() % (H / 8));
    const int x = (int)(rng() % W), y = (int)(rng() % H);
    for (int j = y; j < std::min(H, y + h); ++j)
      for (int i = x; i < std::min(W, x + w); ++i) L[(size_t)j * W + i] = lab;
  }
  for (int j = 0; j < 24; ++j)
    for (int i = 0; i < 24; ++i) { L[(size_t)j * W + i] = 30; L[(size_t)(H - 1 - j) * W + (W - 1 - i)] = 30; }
  return L;
}

static Eigen::ArrayXXi cutPatch(const std::vector<int32_t>& L, int W, int H, int x0, int y0, int rows, int cols) {
  Eigen::ArrayXXi a(rows, cols);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      const int x = x0 + c, y = y0 + r;
      a(r, c) = (x >= 0 && x < W && y >= 0 && y < H) ? L[(size_t)y * W + x] : 3599;
    }
  return a;
}

static std::vector<int32_t> rowMajorOf(const Eigen::ArrayXXi& a) {
  std::vector<int32_t> out((size_t)a.rows() * a.cols());
  for (int i = 0; i < a.rows(); ++i)
    for (int j = 0; j < a.cols(); ++j) out[(size_t)i * a.cols() + j] = a(i, j);
  return out;
}

// random RRT-like tree below `root`: each new node hangs under a random earlier node; 2 trajectory points per segment
static void growTree(TrajectorySegment* root, int n, double width, double height, std::mt19937& rng) {
  std::vector<TrajectorySegment*> all{root};
  std::uniform_real_distribution<double> ux(-0.2, width + 0.2), uy(-0.2, height + 0.2), ut(0.2, 5.0);
  for (int i = 1; i < n; ++i) {
    TrajectorySegment* par = all[rng() % all.size()];
    TrajectorySegment* s = par->spawnChild();
    EigenTrajectoryPoint a, b;
    a.position_W = par->trajectory.empty() ? Eigen::Vector3d(1.0, 1.0, 0.0) : par->trajectory.back().position_W;
    b.position_W = Eigen::Vector3d(ux(rng), uy(rng), 0.0);
    b.time_from_start_ns = (int64_t)(ut(rng) * 1e9);
    s->trajectory.push_back(a);
    s->trajectory.push_back(b);
    all.push_back(s);
  }
}

// like growTree, but every segment is a polyline of `npts` points with increasing time stamps and a heading per point: what a sensor
// model with sampling_time > 0 samples its viewpoints from
static void growTreeSampled(TrajectorySegment* root, int n, int npts, double width, double height, std::mt19937& rng) {
  std::vector<TrajectorySegment*> all{root};
  std::uniform_real_distribution<double> ux(-0.2, width + 0.2), uy(-0.2, height + 0.2), ut(0.15, 0.6), uyaw(-3.1, 3.1);
  for (int i = 1; i < n; ++i) {
    TrajectorySegment* par = all[rng() % all.size()];
    TrajectorySegment* s = par->spawnChild();
    const Eigen::Vector3d a = par->trajectory.empty() ? Eigen::Vector3d(1.0, 1.0, 0.0) : par->trajectory.back().position_W;
    const Eigen::Vector3d b(ux(rng), uy(rng), 0.0);
    int64_t t = 0;
    for (int k = 0; k < npts; ++k) {
      EigenTrajectoryPoint p;
      const double f = (double)k / (npts - 1);
      p.position_W = Eigen::Vector3d(a[0] + f * (b[0] - a[0]), a[1] + f * (b[1] - a[1]), 0.0);
      const double yaw = uyaw(rng);
      p.orientation_W_B = Eigen::Quaterniond(std::cos(yaw / 2), 0.0, 0.0, std::sin(yaw / 2));
      p.time_from_start_ns = t;
      t += (int64_t)(ut(rng) * 1e9);
      s->trajectory.push_back(p);
    }
    all.push_back(s);
  }
}

static int testCpu() {
  // every type string of the GPU package is registered by static initialisers, like the reference modules
  initialize::sipp_gpu_package();
  const char* types[] = {"GpuMapPlanExp", "GpuRRTStarEvaluator", "GpuGoalDistanceEvaluator", "GpuExpandReachableAreaEvaluator",
                         "GpuAverageViewCostEvaluator", "GpuExpandLowCostAreaEvaluator", "GpuNaiveEvaluator"};
  for (const char* t : types) CHECK(ModuleFactoryRegistry::table().count(t) == 1, "type '%s' not registered", t);

  // FlatTree: pre-order with children in vector order, parent[i] < i
  TrajectorySegment root;
  std::mt19937 rng(5);
  growTree(&root, 200, 16.0, 12.0, rng);
  trajectory_evaluator::FlatTree t;
  t.build(&root);
  CHECK(t.nodes.size() == 200, "flattened %zu nodes", t.nodes.size());
  CHECK(t.parent[0] == -1, "root parent");
  for (size_t i = 1; i < t.nodes.size(); ++i) {
    CHECK(t.parent[i] >= 0 && t.parent[i] < (int)i, "parent[%zu] = %d", i, t.parent[i]);
    CHECK(t.nodes[i]->parent == t.nodes[t.parent[i]], "parent pointer mismatch at %zu", i);
  }
  {  // pre-order: first child directly follows its parent
    for (size_t i = 0; i + 1 < t.nodes.size(); ++i)
      if (!t.nodes[i]->children.empty()) CHECK(t.nodes[i + 1] == t.nodes[i]->children[0].get(), "pre-order broken at %zu", i);
  }

  // updater predicate (rrt_star_updater.cpp:19-42)
  sipp_updater_params u = sipp_updater_params();
  u.updater = SIPP_UPD_RRT_STAR; u.following = SIPP_FOLLOW_UPDATE_ALL;
  u.minimum_gain = 0.001; u.update_range = 4.0; u.update_gain = 1; u.update_cost = 0; u.update_value = 1;
  for (size_t i = 1; i < t.nodes.size(); ++i) { t.gain[i] = (i % 3 == 0) ? 0.0 : 1.0; t.terminal[i] = (i % 7 == 0); }
  const double cur[2] = {8.0, 6.0};
  std::vector<uint8_t> dg, mv;
  trajectory_evaluator::updaterMasks(u, t, cur, false, &dg, &mv);
  for (size_t i = 1; i < t.nodes.size(); ++i) {
    const double dx = cur[0] - t.end_xy[2 * i], dy = cur[1] - t.end_xy[2 * i + 1];
    const bool expect = !t.terminal[i] && t.gain[i] > 0.001 && std::sqrt(dx * dx + dy * dy) <= 4.0;
    CHECK((dg[i] != 0) == expect && (mv[i] != 0) == expect, "RRTStarUpdater predicate at %zu", i);
  }
  trajectory_evaluator::updaterMasks(u, t, cur, true, &dg, &mv);   // path changed: everything non-terminal is re-scored
  for (size_t i = 1; i < t.nodes.size(); ++i) CHECK((dg[i] != 0) == !t.terminal[i], "path_changed predicate at %zu", i);
  u.updater = SIPP_UPD_CONSTRAINED;
  trajectory_evaluator::updaterMasks(u, t, cur, true, &dg, &mv);   // ConstrainedUpdater ignores terminal / path change
  for (size_t i = 1; i < t.nodes.size(); ++i) {
    const double dx = cur[0] - t.end_xy[2 * i], dy = cur[1] - t.end_xy[2 * i + 1];
    CHECK((dg[i] != 0) == (t.gain[i] > 0.001 && std::sqrt(dx * dx + dy * dy) <= 4.0), "ConstrainedUpdater predicate at %zu", i);
  }
  u.updater = SIPP_UPD_DIRECT; u.following = SIPP_FOLLOW_UPDATE_ALL_NON_TERMINAL;
  trajectory_evaluator::updaterMasks(u, t, cur, false, &dg, &mv);
  for (size_t i = 1; i < t.nodes.size(); ++i) CHECK((dg[i] != 0) == !t.terminal[i] && mv[i] == 1, "UpdateAllNonTerminal at %zu", i);

  // viewpoint sampling (SensorModelPlanExp::sampleViewpoints, sensor_model_planexp.cpp:65-90) against the checker's statement of it,
  // and the mounting translation rotated by the point's orientation
  {
    TrajectorySegment sroot;
    std::mt19937 r2(9);
    growTreeSampled(&sroot, 60, 7, 16.0, 12.0, r2);
    trajectory_evaluator::FlatTree st;
    for (double stime : {0.0, -1.0, 0.2, 0.5, 0.9, 1.5, 100.0}) {
      trajectory_evaluator::ViewSampler vs;
      vs.sampling_time = stime;
      vs.mounting_translation = Eigen::Vector3d(0.25, -0.1, 0.3);
      st.build(&sroot, &vs);
      CHECK(st.view_off.size() == st.nodes.size() + 1 && st.view_off[0] == 0, "view_off shape");
      for (size_t i = 0; i < st.nodes.size(); ++i) {
        const auto& tr = st.nodes[i]->trajectory;
        std::vector<int64_t> times;
        for (const auto& p : tr) times.push_back(p.time_from_start_ns);
        std::vector<int32_t> want(tr.size() + 1);
        const int k = orc_sample_viewpoints(times.data(), (int32_t)times.size(), stime, want.data(), (int32_t)want.size());
        std::vector<int> got;
        vs.indices(tr, &got);
        CHECK((int)got.size() == k && st.view_off[i + 1] - st.view_off[i] == k, "node %zu: %zu viewpoints, checker %d (sampling_time %g)", i, got.size(), k, stime);
        for (int j = 0; j < k && j < (int)got.size(); ++j) {
          CHECK(got[j] == want[j], "node %zu viewpoint %d: index %d vs %d", i, j, got[j], want[j]);
          const auto& p = tr[got[j]];
          const double yaw = 2.0 * std::atan2(p.orientation_W_B.z(), p.orientation_W_B.w());
          const double ex = p.position_W[0] + std::cos(yaw) * 0.25 - std::sin(yaw) * -0.1, ey = p.position_W[1] + std::sin(yaw) * 0.25 + std::cos(yaw) * -0.1;
          const double gx = st.view_xy[2 * (st.view_off[i] + j)], gy = st.view_xy[2 * (st.view_off[i] + j) + 1];
          CHECK(std::fabs(gx - ex) < 1e-12 && std::fabs(gy - ey) < 1e-12, "node %zu viewpoint %d: (%.15g, %.15g) vs (%.15g, %.15g)", i, j, gx, gy, ex, ey);
        }
      }
    }
  }

  // without a CUDA device the map module reports the failure through printError / checkParamsValid: no CPU fallback
  FakePlanner planner;
  Module::ParamMap mp = mapParams(128, 96);
  std::string err;
  planner.map = ModuleFactoryRegistry::createModule<Map>(&mp, planner, &err);
  if (!planner.map) {
    CHECK(!planner.errors.empty() && planner.errors[0].find("GpuMapPlanExp") != std::string::npos, "error not reported");
    CHECK(err.find("no CPU fallback") != std::string::npos || !err.empty(), "empty error");
    std::printf("no device: createModule refused with: %s\n", err.c_str());
  } else {
    std::printf("device present: map module created\n");
  }
  // unknown updater type is a configuration error, not a silent default
  Module::ParamMap bad = evalParams("GpuRRTStarEvaluator", "NoSuchUpdater", "UpdateAll");
  if (planner.map) {
    auto ev = ModuleFactoryRegistry::createModule<TrajectoryEvaluator>(&bad, planner, &err);
    CHECK(!ev && err.find("NoSuchUpdater") != std::string::npos, "bad updater accepted");
    // ... and so is an unknown cost computer (cost_computer/type: SegmentTime [upstream] or (Gpu)CostPlanExpMapIntegral)
    Module::ParamMap badc = evalParams("GpuRRTStarEvaluator", "RRTStarUpdater", "UpdateAll");
    badc["cost_computer/type"] = "NoSuchCost";
    auto ev2 = ModuleFactoryRegistry::createModule<TrajectoryEvaluator>(&badc, planner, &err);
    CHECK(!ev2 && err.find("NoSuchCost") != std::string::npos, "bad cost computer accepted");
  }
  // GpuPRMStarPlanner: INVALID + cost -1 when it has no context (prm_star_planner.cpp:39,66-68)
  planexp_base_planners::PlannerParams pp;
  pp.px_x = 0; pp.px_y = 0; pp.width = 0; pp.height = 0; pp.min_cost = 30; pp.max_cost = 120; pp.init_cost = 30;
  pp.unknown_id = 3599;
  planexp_base_planners::GpuPRMStarPlanner bp(pp);
  bp.setup();
  double cost = 0;
  std::vector<Eigen::Vector2d> path;
  CHECK(bp.plan(&path, cost) == planexp_base_planners::ResultStatus::INVALID && cost == -1.0, "plan without ctx");
  CHECK(!bp.lastError().empty(), "setup error not kept");
  return g_fail;
}

struct Config {
  const char *type, *updater, *following;
  int evaluator;
};

static int testGpu() {
  const int W = 256, H = 192;
  const std::vector<int32_t> L = makeLabels(W, H, 11);
  const Config cfgs[] = {
      {"GpuRRTStarEvaluator", "RRTStarUpdater", "UpdateAll", SIPP_EVAL_RRT_STAR},
      {"GpuRRTStarEvaluator", "RRTStarUpdater", "UpdateAllNonTerminal", SIPP_EVAL_RRT_STAR},
      {"GpuGoalDistanceEvaluator", "ConstrainedUpdater", "UpdateAll", SIPP_EVAL_GOAL_DISTANCE},
      {"GpuExpandReachableAreaEvaluator", "UpdateAllNonTerminal", "", SIPP_EVAL_EXPAND_REACHABLE},
      {"GpuAverageViewCostEvaluator", "ConstrainedUpdater", "UpdateAll", SIPP_EVAL_AVERAGE_VIEW_COST},
      {"GpuNaiveEvaluator", "RRTStarUpdater", "UpdateAll", SIPP_EVAL_NAIVE},
      {"GpuExpandLowCostAreaEvaluator", "ConstrainedUpdater", "UpdateAll", SIPP_EVAL_EXPAND_LOW_COST},
  };
  for (const Config& cfg : cfgs) {
    std::printf("--- %s / %s / %s\n", cfg.type, cfg.updater, cfg.following);
    FakePlanner planner;
    Module::ParamMap mp = mapParams(W, H);
    std::string err;
    planner.map = ModuleFactoryRegistry::createModule<Map>(&mp, planner, &err);
    CHECK(planner.map != nullptr, "map module: %s", err.c_str());
    if (!planner.map) return g_fail;
    auto* gmap = dynamic_cast<map::GpuMapPlanExp*>(planner.map.get());
    Module::ParamMap ep = evalParams(cfg.type, cfg.updater, cfg.following);
    planner.evaluator = ModuleFactoryRegistry::createModule<TrajectoryEvaluator>(&ep, planner, &err);
    CHECK(planner.evaluator != nullptr, "evaluator module: %s", err.c_str());
    if (!planner.evaluator) return g_fail;
    auto* gev = dynamic_cast<trajectory_evaluator::GpuGainEvaluator*>(planner.evaluator.get());
    CHECK(gev->evalParams().evaluator == cfg.evaluator && gev->evalParams().half_fov == 32, "params not parsed");

    orc_ctx* orc = orc_create(&gmap->desc());
    // MapData stream: an init area at the start and a handful of 65x65 footprints (mav_simulator.cpp:438-495)
    std::mt19937 rng(3);
    auto ingest = [&](double px, double py, int fov) {
      int x0, y0;
      orc_patch_origin(orc, px, py, fov, fov, &x0, &y0);
      Eigen::ArrayXXi patch = cutPatch(L, W, H, x0, y0, fov, fov);
      CHECK(gmap->mapDataCallback(Eigen::Vector3d(px, py, 0.0), patch), "mapDataCallback");
      orc_map_update_patch(orc, x0, y0, fov, fov, rowMajorOf(patch).data());
    };
    ingest(0.53125, 0.53125, 129);
    double px = 2.0, py = 2.0;
    for (int k = 0; k < 14; ++k) {
      ingest(px, py, 65);
      px = std::min(W * 0.0625 - 0.1, px + 0.9 + (rng() % 100) * 0.01);
      py = std::min(H * 0.0625 - 0.1, py + 0.6 + (rng() % 100) * 0.01);
    }
    double gm = gmap->getMeanWeight(), om, omax;
    int64_t ocnt;
    orc_map_stats(orc, &om, &omax, &ocnt);
    CHECK(gm == om, "mean cost %.17g vs %.17g", gm, om);

    // follower path (compute_path: the first mapDataCallback already planned once; plan again on the final map)
    std::vector<double> oxy(2 * 4 * (size_t)(W + H));
    double ocost; int32_t ostatus, olen;
    orc_plan(orc, &ocost, &ostatus, oxy.data(), (int32_t)(oxy.size() / 2), &olen, 1);
    gmap->targetReachedCallback();
    CHECK(gmap->computeCurrentPathEstimate(), "plan failed");
    CHECK(gmap->lastPathCost() == ocost, "path cost %.17g vs %.17g", gmap->lastPathCost(), ocost);
    CHECK(gmap->lastPathStatus() == ostatus, "status %d vs %d", gmap->lastPathStatus(), ostatus);
    CHECK((int)gmap->lastPath().size() == olen, "path length %zu vs %d", gmap->lastPath().size(), olen);
    for (int i = 0; i < olen && i < (int)gmap->lastPath().size(); ++i)
      if (gmap->lastPath()[i][0] != oxy[2 * i] || gmap->lastPath()[i][1] != oxy[2 * i + 1]) { CHECK(false, "path node %d differs", i); break; }
    // isObserved (map_planexp.cpp:59-61) — probe a few points
    CHECK(gmap->isObserved(Eigen::Vector3d(0.6, 0.6, 0.0)), "start area must be observed");
    CHECK(!gmap->isObserved(Eigen::Vector3d(-0.1, 0.6, 0.0)) && !gmap->isObserved(Eigen::Vector3d(W * 0.0625, 1.0, 0.0)), "off-map must be unobserved");

    // tree: per-node computeGain / computeCost (what expandTrajectories does), then computeValue
    TrajectorySegment root;
    { EigenTrajectoryPoint p0; p0.position_W = Eigen::Vector3d(2.0, 2.0, 0.0); root.trajectory.push_back(p0); }
    std::mt19937 trng(17);
    const int N = 400;
    growTree(&root, N, W * 0.0625, H * 0.0625, trng);
    trajectory_evaluator::FlatTree t;
    t.build(&root);
    std::vector<double> og(N), oc(N), ov(N);
    std::vector<uint8_t> ot(N);
    const sipp_eval_params P = gev->evalParams();
    orc_gain_batch(orc, &P, N, t.end_xy.data(), t.start_xy.data(), og.data(), nullptr, ot.data(), 0, 1);
    const bool sets_terminal = cfg.evaluator == SIPP_EVAL_RRT_STAR || cfg.evaluator == SIPP_EVAL_EXPAND_REACHABLE;
    int bad_gain = 0;
    for (int i = 1; i < N; ++i) {
      CHECK(gev->computeGain(t.nodes[i]) && gev->computeCost(t.nodes[i]), "computeGain/Cost");
      const double g = t.nodes[i]->gain, e = og[i];
      const bool ok = (std::isnan(g) && std::isnan(e)) || g == e || std::fabs(g - e) <= 1e-5 * std::fabs(e);
      bad_gain += !ok;
      if (sets_terminal) bad_gain += (t.nodes[i]->gain_terminal != (ot[i] != 0));
      // feed the oracle the module's own gains so that the later comparisons isolate the update pass
      og[i] = g; oc[i] = t.nodes[i]->cost; ot[i] = t.nodes[i]->gain_terminal;
    }
    CHECK(bad_gain == 0, "%d per-node gains differ", bad_gain);
    root.gain = 0; root.cost = 0; root.value = 0; og[0] = oc[0] = 0.0; ot[0] = 0;
    orc_value_select(N, og.data(), oc.data(), t.parent.data(), ov.data(), nullptr, nullptr);
    for (int i : {1, 7, 133, N - 1}) {
      CHECK(gev->computeValue(t.nodes[i]), "computeValue");
      CHECK(t.nodes[i]->value == ov[i], "computeValue node %d: %.17g vs %.17g", i, t.nodes[i]->value, ov[i]);
    }
    if (&cfg == &cfgs[0]) {
      // cost_computer/type: CostPlanExpMapIntegral (cfg/evaluators/baseline.yaml:14-15) — computeCost node by node, parents first,
      // accumulate on; against the oracle's statement of cost_planexp_map_integral.cpp:21-73
      Module::ParamMap cp2 = evalParams(cfg.type, cfg.updater, cfg.following);
      cp2["cost_computer/type"] = "CostPlanExpMapIntegral";
      cp2["cost_computer/accumulate"] = "true";
      cp2["cost_computer/max_sample_distance"] = "0.05";
      auto cev = ModuleFactoryRegistry::createModule<TrajectoryEvaluator>(&cp2, planner, &err);
      CHECK(cev != nullptr, "evaluator with the map-integral cost: %s", err.c_str());
      const sipp_cost_params CP = {1, 0, 0.05};
      std::vector<double> saved(N);
      for (int i = 0; i < N; ++i) saved[i] = t.nodes[i]->cost;
      int bad_cost = 0;
      for (int i = 1; i < N && cev; ++i) {
        std::vector<double> pts;
        for (const auto& p : t.nodes[i]->trajectory) { pts.push_back(p.position_W.x()); pts.push_back(p.position_W.y()); pts.push_back(p.position_W.z()); }
        const int32_t off[2] = {0, (int32_t)t.nodes[i]->trajectory.size()};
        const double pc = t.nodes[i]->parent->cost;      // already the integral cost (parents come first in the flat order)
        double want; uint8_t wv;
        orc_cost_integral(orc, &CP, 1, off, pts.data(), &pc, &want, &wv);
        const bool rv = cev->computeCost(t.nodes[i]);
        bad_cost += (rv != (wv != 0)) || t.nodes[i]->cost != want;
      }
      CHECK(bad_cost == 0, "%d map-integral costs differ", bad_cost);
      for (int i = 0; i < N; ++i) t.nodes[i]->cost = saved[i];
    }
    for (int i = 1; i < N; ++i) t.nodes[i]->value = ov[i];

    // the map changes (new patch + re-plan -> path id changes), then the planner's re-evaluation pass
    ingest(W * 0.0625 * 0.5, H * 0.0625 * 0.5, 65);
    orc_plan(orc, &ocost, &ostatus, oxy.data(), (int32_t)(oxy.size() / 2), &olen, 1);
    gmap->targetReachedCallback();
    CHECK(gmap->computeCurrentPathEstimate() && gmap->lastPathCost() == ocost, "re-plan cost");
    for (int pass = 0; pass < 2; ++pass) {   // pass 0: path changed; pass 1: same path id, range/min-gain gating active
      planner.position = Eigen::Vector3d(pass ? 6.0 : 3.0, pass ? 5.0 : 2.5, 0.0);
      const double cur_xy[2] = {planner.position[0], planner.position[1]};
      std::vector<double> g2(N), v2(N);
      std::vector<uint8_t> t2(N);
      for (int i = 0; i < N; ++i) { g2[i] = t.nodes[i]->gain; v2[i] = t.nodes[i]->value; t2[i] = t.nodes[i]->gain_terminal; }
      const sipp_updater_params U = gev->updaterParams();
      const int path_changed = (U.updater == SIPP_UPD_RRT_STAR && pass == 0) ? 1 : 0;
      orc_update_tree(orc, &P, &U, N, t.parent.data(), t.end_xy.data(), t.start_xy.data(), g2.data(), oc.data(), v2.data(), t2.data(),
                      cur_xy, path_changed, 0, nullptr);
      // the planner calls updateSegment(root) first (mav_active_3d_planning.patch:56), then once per node
      const auto tm0 = std::chrono::steady_clock::now();
      CHECK(gev->updateSegment(&root), "updateSegment(root)");
      const auto tm1 = std::chrono::steady_clock::now();
      for (int i = 1; i < N; ++i) CHECK(gev->updateSegment(t.nodes[i]), "updateSegment(node)");
      const auto tm2 = std::chrono::steady_clock::now();
      // what the planner pays per re-evaluation of the tree through the module interface: the root call flattens the tree, runs
      // the whole pass on the device (sipp_update_tree: copies in, predicate + compaction, gain, values, copies out) and writes
      // the nodes back; the per-node calls that follow find their results already there
      std::printf("    timing %s pass %d: updateSegment(root) %.3f ms for %d nodes (%d re-scored), the %d per-node calls after it %.3f ms\n", cfg.type, pass,
                  std::chrono::duration<double, std::milli>(tm1 - tm0).count(), N, gev->lastBatchSize(), N - 1,
                  std::chrono::duration<double, std::milli>(tm2 - tm1).count());
      int dg = 0, dv = 0, dt = 0;
      // the root-first call re-scores the root too unless the chain starts with an RRTStarUpdater (pinned by tests/test_ref_pin.py)
      for (int i = (U.updater == SIPP_UPD_RRT_STAR ? 1 : 0); i < N; ++i) {
        const double g = t.nodes[i]->gain, v = t.nodes[i]->value;
        dg += !((std::isnan(g) && std::isnan(g2[i])) || g == g2[i] || std::fabs(g - g2[i]) <= 1e-5 * std::fabs(g2[i]));
        dv += !((std::isnan(v) && std::isnan(v2[i])) || v == v2[i] || std::fabs(v - v2[i]) <= 1e-5 * std::fabs(v2[i]));
        dt += (t.nodes[i]->gain_terminal != (t2[i] != 0));
      }
      CHECK(dg == 0 && dv == 0 && dt == 0, "pass %d: %d gains, %d values, %d terminal flags differ (batch %d)", pass, dg, dv, dt, gev->lastBatchSize());
      std::printf("    pass %d: batch of %d nodes re-scored in one launch, gains/values/terminal match the pre-order emulation\n", pass,
                  gev->lastBatchSize());
      if (pass == 0 && U.updater != SIPP_UPD_CONSTRAINED) CHECK(gev->lastBatchSize() > N / 2, "path change must re-score the tree");
    }
    // selection: [upstream] SubsequentBest on the stored values (lowest index on ties)
    {
      std::vector<double> sub(N);
      for (int i = 0; i < N; ++i) sub[i] = t.nodes[i]->value;
      for (int i = N - 1; i >= 1; --i)
        if (t.parent[i] != 0) sub[t.parent[i]] = std::max(sub[t.parent[i]], sub[i]);
      int best = -1, k = 0; double bv = 0;
      for (int i = 1; i < N; ++i)
        if (t.parent[i] == 0) { if (best < 0 || sub[i] > bv) { best = k; bv = sub[i]; } ++k; }
      CHECK(gev->selectNextBest(&root) == best, "selectNextBest");
    }
    // sensor_model/sampling_time > 0 + a mounting translation (sensor_model_planexp.cpp:18-90): per-node computeGain and the batched
    // root-first pass over a tree of sampled polylines, against the checker scoring the same viewpoints
    {
      Module::ParamMap sp = evalParams(cfg.type, cfg.updater, cfg.following);
      sp["sensor_model/sampling_time"] = "0.7";
      sp["sensor_model/mounting_translation_x"] = "0.2";
      sp["sensor_model/mounting_translation_y"] = "-0.15";
      auto sev_owner = ModuleFactoryRegistry::createModule<TrajectoryEvaluator>(&sp, planner, &err);
      CHECK(sev_owner != nullptr, "evaluator with sampled viewpoints: %s", err.c_str());
      auto* sev = dynamic_cast<trajectory_evaluator::GpuGainEvaluator*>(sev_owner.get());
      TrajectorySegment sroot;
      std::mt19937 r3(21);
      growTreeSampled(&sroot, 300, 6, W * 0.0625, H * 0.0625, r3);
      trajectory_evaluator::ViewSampler vs;
      vs.sampling_time = 0.7;
      vs.mounting_translation = Eigen::Vector3d(0.2, -0.15, 0.0);
      trajectory_evaluator::FlatTree st;
      st.build(&sroot, &vs);
      const int SN = (int)st.nodes.size();
      std::vector<double> sg(SN), scost(SN), sval(SN, 0.0);
      std::vector<uint8_t> stt(SN);
      orc_gain_batch_views(orc, &P, SN, st.view_off.data(), st.view_xy.data(), st.start_xy.data(), sg.data(), nullptr, stt.data(), 1);
      int bad = 0, multi = 0;
      for (int i = 1; i < SN && sev; ++i) {
        multi += st.view_off[i + 1] - st.view_off[i] > 1;
        CHECK(sev->computeGain(st.nodes[i]), "computeGain (sampled viewpoints)");
        CHECK(sev->computeCost(st.nodes[i]), "computeCost");
        scost[i] = st.nodes[i]->cost;
        const double a = st.nodes[i]->gain, b = sg[i];
        bad += !(a == b || std::fabs(a - b) <= 1e-5 * std::fabs(b)) || (st.nodes[i]->gain_terminal != (stt[i] != 0));
      }
      CHECK(bad == 0, "%d gains / terminal flags differ with sampled viewpoints", bad);
      CHECK(multi > SN / 3, "only %d of %d segments have several viewpoints", multi, SN);
      // the batched pass
      planner.position = Eigen::Vector3d(4.0, 3.0, 0.0);
      const double cur_xy[2] = {4.0, 3.0};
      std::vector<double> g2(SN), v2(SN);
      std::vector<uint8_t> t2(SN);
      for (int i = 0; i < SN; ++i) { g2[i] = st.nodes[i]->gain; v2[i] = st.nodes[i]->value; t2[i] = st.nodes[i]->gain_terminal; }
      const sipp_updater_params U = sev ? sev->updaterParams() : sipp_updater_params();
      orc_update_tree_views(orc, &P, &U, SN, st.parent.data(), st.end_xy.data(), st.start_xy.data(), g2.data(), scost.data(), v2.data(), t2.data(), cur_xy,
                            U.updater == SIPP_UPD_RRT_STAR ? 1 : 0, 1, nullptr, st.view_off.data(), st.view_xy.data());
      if (sev) CHECK(sev->updateSegment(&sroot), "updateSegment(root) with sampled viewpoints");
      int dg = 0, dv = 0, dt = 0;
      for (int i = (U.updater == SIPP_UPD_RRT_STAR ? 1 : 0); i < SN; ++i) {
        const double g = st.nodes[i]->gain, v = st.nodes[i]->value;
        dg += !(g == g2[i] || std::fabs(g - g2[i]) <= 1e-5 * std::fabs(g2[i]));
        dv += !(v == v2[i] || std::fabs(v - v2[i]) <= 1e-5 * std::fabs(v2[i]));
        dt += (st.nodes[i]->gain_terminal != (t2[i] != 0));
      }
      CHECK(dg == 0 && dv == 0 && dt == 0, "sampled viewpoints pass: %d gains, %d values, %d terminal flags differ", dg, dv, dt);
      std::printf("    sampled viewpoints (sampling_time 0.7 s, mounting translation): %d of %d segments viewed from several points, gains / pass match\n", multi, SN);
    }
    orc_destroy(orc);
    for (auto& e : planner.errors) std::printf("    planner error: %s\n", e.c_str());
    CHECK(planner.errors.empty(), "modules reported errors");
  }

  // GpuPRMStarPlanner behind PlannerBase, driven like planexp_eval_planner.cpp:263-285 (fully known map, obstacles U {unknown})
  {
    planexp_base_planners::PlannerParams pp;
    pp.px_x = W; pp.px_y = H; pp.width = W * 0.0625; pp.height = H * 0.0625;
    pp.min_cost = 30; pp.max_cost = 120; pp.init_cost = 30;
    pp.start = Eigen::Vector2d(0.53125, 0.53125);
    pp.goal = Eigen::Vector2d((W - 8) * 0.0625 + 0.03125, (H - 8) * 0.0625 + 0.03125);
    pp.obstacle_ids = {0, 3599};
    pp.unknown_id = 3599;
    planexp_base_planners::GpuPRMStarPlanner bp(pp);
    bp.setup();
    CHECK(bp.lastError().empty(), "setup: %s", bp.lastError().c_str());
    Eigen::ArrayXXi whole(H, W);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) whole(y, x) = L[(size_t)y * W + x];
    bp.update_states(Eigen::Vector2i(0, 0), whole);
    std::vector<Eigen::Vector2d> path;
    double cost = 0;
    const auto st = bp.plan(&path, cost);
    sipp_map_desc d = sipp_map_desc();
    d.W = W; d.H = H; d.width = pp.width; d.height = pp.height; d.start_x = pp.start.x(); d.start_y = pp.start.y();
    d.goal_x = pp.goal.x(); d.goal_y = pp.goal.y(); d.min_rov_cost = 30; d.max_rov_cost = 120; d.unknown_init_cost = 30; d.unknown_id = 3599;
    d.n_obstacle_ids_rov = 2; d.obstacle_ids_rov[0] = 0; d.obstacle_ids_rov[1] = 3599;
    orc_ctx* orc = orc_create(&d);
    orc_map_update_patch(orc, 0, 0, H, W, L.data());
    std::vector<double> oxy(2 * 4 * (size_t)(W + H));
    double ocost; int32_t ostatus, olen;
    orc_plan(orc, &ocost, &ostatus, oxy.data(), (int32_t)(oxy.size() / 2), &olen, 1);
    CHECK(cost == ocost && (int)path.size() == olen, "GpuPRMStarPlanner cost %.17g vs %.17g, %zu vs %d nodes", cost, ocost, path.size(), olen);
    CHECK((int)st == (ostatus == SIPP_PLAN_VALID ? 0 : ostatus == SIPP_PLAN_INVALID ? 1 : ostatus == SIPP_PLAN_TERMINAL_VALID ? 2 : 3), "status");
    std::printf("--- GpuPRMStarPlanner: cost %.6f, %zu nodes, status %d\n", cost, path.size(), (int)st);
    orc_destroy(orc);
  }
  return g_fail;
}

int main(int argc, char** argv) {
  const bool gpu = argc > 1 && !std::strcmp(argv[1], "gpu");
  const int f = gpu ? testGpu() : testCpu();
  std::printf("%s: %d failure(s)\n", gpu ? "gpu" : "cpu", f);
  return f ? 1 : 0;
}
