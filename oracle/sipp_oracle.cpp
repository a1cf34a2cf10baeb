// =====================================================================================================
// sipp_oracle.cpp — CPU restatement of scouting-ipp's per-planning-cycle evaluation hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it. libsipp.so never links or calls anything in here.
//
// The reference (ethz-asl/scouting-ipp) cannot be compiled in this environment: its hot-path translation
// units need ROS, Eigen, Boost.Graph, OMPL, OpenCV, yaml-cpp, glog and the un-vendored
// ethz-asl/mav_active_3d_planning @ 4c139dcb (docker/scouting_ipp.Dockerfile:41-44). This file restates the
// algorithms on flat arrays with the reference's exact double-precision operation order; every function
// cites the reference file:line it follows (paths relative to the reference root).
//
// Parity pinning:
//   * follower path cost (A11): PINNED by the ten golden `optimal_prm_rov_path_cost` values shipped in
//     docker/data/maps/cfg/map_*.yaml:15 (tests/test_oracle_golden.py, fixtures tests/golden/ref_maps.*).
//   * footprint/gain evaluators (A1-A8), map update (A10), path mask (A12): parity UNPINNED by the
//     reference (it ships no test or fixture for them); defined by this restatement + the hand-checkable
//     cases derived from the cited sources (tests/test_oracle_gain.py).
//   * value / selection (A14): [upstream] GlobalNormalizedGain / SubsequentBest, restated from the public
//     upstream sources' semantics; parity UNPINNED; ties broken by lowest index (the reference uses rand()).
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared (see oracle/Makefile). -ffp-contract=off because
// the reference was built for baseline x86-64 (no FMA), and the goldens depend on ulp-level rounding.
// =====================================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <queue>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../include/sipp.h"

namespace {

struct Vec3 {  // stands in for Eigen::Vector3d (24 bytes — matters for the cost of the reference's erase loop)
  double x, y, z;
  bool operator==(const Vec3& o) const { return x == o.x && y == o.y && z == o.z; }
};

struct Oracle {
  sipp_map_desc desc;
  int W, H;
  // ---- MapServerPlanExp state (map_server_planexp.h:150-190), flat row-major [y*W+x]
  double mpp;                 // m_per_pixel_ = width_/horizontal_units_            map_server_planexp.cpp:116
  std::vector<int> state;     // map_state_
  std::vector<double> cost;   // map_cost_
  std::vector<int> reach;     // map_reachability_
  std::vector<double> cost_closest, dist_closest;
  std::vector<int> path_mask;  // current_path_map_estimate_
  int path_counter = 0;
  bool path_valid = false;
  double mean_cost = 0.0, max_cost = 0.0;
  int mean_counter = 0;
  int max_area_id = 1;
  int start_loc[2], goal_loc[2];
  int oob_events = 0;  // places where the reference would read/write out of bounds (we clamp/skip and count)
  // ---- MapPlanExp cached constants (map_planexp.cpp:35-44)
  double vs;  // c_voxel_size_ = width_/horizontal_units_
  // ---- PRMStarPlanner state (prm_star_planner.h:52-75)
  double spacing, max_dist, off_x, off_y;
  double init_cost;
  std::vector<double> X, Y;     // node poses per column / row
  std::vector<double> pmap;     // planner map_ (double)
  std::vector<uint8_t> explored, removed;
  int start_node[2], goal_node[2];
  std::vector<double> field;    // d[] of the last plan
  std::vector<int> last_path;   // node ids (x*H+y) start -> goal
  std::string err;
};

inline bool in_ids(const int32_t* ids, int n, int v) {
  for (int i = 0; i < n; ++i)
    if (ids[i] == v) return true;
  return false;
}

// ----------------------------------------------------------------------------------------------------
// PRMStarPlanner::set_derived_params  — planexp_base_planners/src/prm_star_planner.cpp:16-36
// node_id2pose / closest_xy          — include/planexp_base_planners/prm_star_planner.h:78-81
// ----------------------------------------------------------------------------------------------------
void planner_set_derived(Oracle& o) {
  const int W = o.W, H = o.H;
  o.spacing = o.desc.width / (W + 1);                       // :29  NOT width / W
  o.max_dist = o.spacing * std::sqrt(2.0);                  // :30
  o.off_x = (o.desc.width - (W + 1) * o.spacing) / 2.0;     // :31
  o.off_y = (o.desc.height - (H + 1) * o.spacing) / 2.0;    // :32  (x spacing used for y as well)
  o.X.resize(W);
  o.Y.resize(H);
  for (int i = 0; i < W; ++i) o.X[i] = o.off_x + o.spacing * i;   // .h:80
  for (int j = 0; j < H; ++j) o.Y[j] = o.off_y + o.spacing * j;
  // closest_xy: std::round (half away from zero) — .h:81
  o.start_node[0] = (int)std::round((o.desc.start_x - o.off_x) / o.spacing);
  o.start_node[1] = (int)std::round((o.desc.start_y - o.off_y) / o.spacing);
  o.goal_node[0] = (int)std::round((o.desc.goal_x - o.off_x) / o.spacing);
  o.goal_node[1] = (int)std::round((o.desc.goal_y - o.off_y) / o.spacing);
  o.init_cost = o.desc.unknown_init_cost;
  o.pmap.assign((size_t)W * H, o.init_cost);                 // :26-27
  o.explored.assign((size_t)W * H, 0);
  o.removed.assign((size_t)W * H, 0);
  o.field.clear();
  o.last_path.clear();
}

// (node_id2pose(a)-node_id2pose(b)).norm() — prm_star_planner.cpp:138,148,179
inline double edge_dist(const Oracle& o, int ax, int ay, int bx, int by) {
  const double dX = o.X[ax] - o.X[bx];
  const double dY = o.Y[ay] - o.Y[by];
  return std::sqrt(dX * dX + dY * dY);
}

// edge (a,b) exists in the graph: build_graph admission test :138 and no delete_connections on either end :158-168
inline bool edge_present(const Oracle& o, int ax, int ay, int bx, int by, double* dist_out) {
  const double dist = edge_dist(o, ax, ay, bx, by);
  *dist_out = dist;
  if (!(dist < o.max_dist)) return false;
  if (o.removed[(size_t)ay * o.W + ax] || o.removed[(size_t)by * o.W + bx]) return false;
  return true;
}

// compute_edge_cost — prm_star_planner.cpp:178-181 : dist * (cost_a + cost_b) / 2.0
inline double edge_weight(const Oracle& o, int ax, int ay, int bx, int by, double dist) {
  return dist * (o.pmap[(size_t)ay * o.W + ax] + o.pmap[(size_t)by * o.W + bx]) / 2.0;
}

// PRMStarPlanner::update_states — prm_star_planner.cpp:104-119. `observed` (optional, full-map upload only)
// restricts the update to flagged cells.
void planner_update_states(Oracle& o, int x0, int y0, int rows, int cols, const int32_t* labels,
                           const uint8_t* observed) {
  for (int i = 0; i < cols; ++i) {
    for (int j = 0; j < rows; ++j) {
      const int x = x0 + i, y = y0 + j;
      if (x < 0 || x >= o.W || y < 0 || y >= o.H) continue;
      if (observed && !observed[(size_t)j * cols + i]) continue;
      const size_t c = (size_t)y * o.W + x;
      if (!o.explored[c]) {                                          // is_new :111
        const int lab = labels[(size_t)j * cols + i];
        o.pmap[c] = lab;                                             // :112
        if (in_ids(o.desc.obstacle_ids_rov, o.desc.n_obstacle_ids_rov, lab)) o.removed[c] = 1;  // :113
        o.explored[c] = 1;                                           // :115
      }
    }
  }
}

// PRMStarPlanner::update_unknown_cost — prm_star_planner.cpp:121-130
void planner_update_unknown_cost(Oracle& o, double new_cost) {
  for (size_t c = 0; c < o.explored.size(); ++c)
    if (!o.explored[c]) o.pmap[c] = new_cost;
  o.init_cost = new_cost;
}

// distance_heuristic::operator() — prm_star_planner.h:108-112
inline double heuristic(const Oracle& o, int x, int y) {
  const int dx = o.goal_node[0] - x, dy = o.goal_node[1] - y;
  return ::sqrt(static_cast<double>(dx * dx + dy * dy)) * o.desc.min_rov_cost * o.spacing;
}

const int NBR_DX[8] = {-1, -1, -1, 0, 0, 1, 1, 1};  // ascending node id = x*H+y  (out-edge order of the reference)
const int NBR_DY[8] = {-1, 0, 1, -1, 1, -1, 0, 1};

// Start-rooted Dijkstra to the fixed point d[v] = min_u d[u] + w(u,v). In IEEE double the least fixed point is
// independent of relaxation order (rounding is monotone), so this equals what boost::astar_search_tree
// (prm_star_planner.cpp:46-52) leaves in d[goal]; pinned by the golden costs.
void planner_field(Oracle& o) {
  const int W = o.W, H = o.H;
  const double INF = std::numeric_limits<double>::infinity();
  o.field.assign((size_t)W * H, INF);
  const int sx = o.start_node[0], sy = o.start_node[1];
  if (sx < 0 || sx >= W || sy < 0 || sy >= H) return;
  typedef std::pair<double, uint32_t> QE;
  std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
  o.field[(size_t)sy * W + sx] = 0.0;
  pq.push(QE(0.0, (uint32_t)(sy * W + sx)));
  while (!pq.empty()) {
    const QE top = pq.top();
    pq.pop();
    const uint32_t c = top.second;
    if (top.first > o.field[c]) continue;
    const int x = c % W, y = c / W;
    for (int k = 0; k < 8; ++k) {
      const int nx = x + NBR_DX[k], ny = y + NBR_DY[k];
      if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
      double dist;
      if (!edge_present(o, x, y, nx, ny, &dist)) continue;
      const double nd = top.first + edge_weight(o, x, y, nx, ny, dist);
      const size_t nc = (size_t)ny * W + nx;
      if (nd < o.field[nc]) {
        o.field[nc] = nd;
        pq.push(QE(nd, (uint32_t)nc));
      }
    }
  }
}

// Canonical predecessor path: backtrack from the goal; among predecessors u with
// d[u] + w(u,v) == d[v] bitwise pick the smallest (d[u] + h(u), id(u)), id = x*H + y. This reproduces the order
// in which boost::astar_search_tree's strict-< relaxation would have fixed p[v] up to heap-order effects.
bool planner_backtrack(Oracle& o) {
  const int W = o.W, H = o.H;
  o.last_path.clear();
  int x = o.goal_node[0], y = o.goal_node[1];
  if (x < 0 || x >= W || y < 0 || y >= H) return false;
  if (!std::isfinite(o.field[(size_t)y * W + x])) return false;
  std::vector<int> rev;
  rev.push_back(x * H + y);
  size_t guard = 0;
  while (!(x == o.start_node[0] && y == o.start_node[1])) {
    const double dv = o.field[(size_t)y * W + x];
    int bx = -1, by = -1;
    double bf = 0.0;
    for (int k = 0; k < 8; ++k) {
      const int ux = x + NBR_DX[k], uy = y + NBR_DY[k];
      if (ux < 0 || ux >= W || uy < 0 || uy >= H) continue;
      double dist;
      if (!edge_present(o, ux, uy, x, y, &dist)) continue;
      const double du = o.field[(size_t)uy * W + ux];
      if (!(du + edge_weight(o, ux, uy, x, y, dist) == dv)) continue;
      const double f = du + heuristic(o, ux, uy);
      // ascending id order + strict '<' keeps the smallest id among equal f
      if (bx < 0 || f < bf) {
        bx = ux; by = uy; bf = f;
      }
    }
    if (bx < 0 || ++guard > (size_t)W * H) return false;
    x = bx; y = by;
    rev.push_back(x * H + y);
  }
  o.last_path.assign(rev.rbegin(), rev.rend());
  return true;
}

// MapServerPlanExp::position2pixelLocation — map_server_planexp.cpp:340-342
inline void position2pixel(const Oracle& o, double px, double py, int* ix, int* iy) {
  *ix = (int)(px * (long)o.W / o.desc.width);
  *iy = (int)(py * (long)o.H / o.desc.height);
}

// MapServerPlanExp::computePathMap — map_server_planexp.cpp:902-926
void compute_path_map(Oracle& o, const std::vector<double>& path_xy, std::vector<int>* mask) {
  const int W = o.W, H = o.H;
  mask->assign((size_t)W * H, 0);
  const size_t n = path_xy.size() / 2;
  if (n == 0) return;
  std::vector<std::pair<int, int>> discrete;
  int ix, iy;
  position2pixel(o, path_xy[0], path_xy[1], &ix, &iy);
  discrete.push_back({ix, iy});
  for (size_t j = 1; j < n; ++j) {
    const double dx = path_xy[2 * j] - path_xy[2 * (j - 1)];
    const double dy = path_xy[2 * j + 1] - path_xy[2 * (j - 1) + 1];
    const double distance = std::sqrt(dx * dx + dy * dy);
    const int samples = (int)(distance / (o.mpp / 3.0));
    if (samples > 0) {
      const double dpx = dx / samples, dpy = dy / samples;
      for (int i = 0; i < samples; i++) {
        const double qx = path_xy[2 * (j - 1)] + i * dpx;
        const double qy = path_xy[2 * (j - 1) + 1] + i * dpy;
        position2pixel(o, qx, qy, &ix, &iy);
        if (discrete.back() != std::make_pair(ix, iy)) discrete.push_back({ix, iy});
      }
    } else {
      position2pixel(o, path_xy[2 * j], path_xy[2 * j + 1], &ix, &iy);
      if (discrete.back() != std::make_pair(ix, iy)) discrete.push_back({ix, iy});
    }
  }
  for (auto& p : discrete) {
    if (p.first < 0 || p.first >= W || p.second < 0 || p.second >= H) {  // Eigen would be out of bounds here
      ++o.oob_events;
      continue;
    }
    (*mask)[(size_t)p.second * W + p.first] = 1;
  }
}

// ----------------------------------------------------------------------------------------------------
// MapServerPlanExp::clearMap — map_server_planexp.cpp:461-487
// ----------------------------------------------------------------------------------------------------
void map_clear(Oracle& o) {
  const size_t N = (size_t)o.W * o.H;
  o.state.assign(N, SIPP_STATE_UNKNOWN);
  o.cost.assign(N, 0.0);
  if (o.desc.compute_closest_observed) {
    o.cost_closest.assign(N, 0.0);
    o.dist_closest.assign(N, std::sqrt(o.desc.width * o.desc.width + o.desc.height * o.desc.height));
  }
  o.reach.assign(N, SIPP_REACHABILITY_UNKNOWN);
  if (o.desc.compute_reachability) {
    o.reach[(size_t)o.start_loc[1] * o.W + o.start_loc[0]] = SIPP_REACHABLE;   // :476
    o.max_area_id = 1;
  }
  o.mean_cost = 0;
  o.mean_counter = 0;
  o.path_mask.assign(N, 0);   // current_path_map_estimate_ = Zero  :161
}

// MapServerPlanExp::setReachabilityMap — map_server_planexp.cpp:825-870. The connected_areas_ vectors only matter
// for their side effect on max_area_id_ (:853-857); area ids >= 2 are never relabelled to MSR_REACHABLE.
void set_reachability(Oracle& o, int sx, int sy) {
  const int W = o.W, H = o.H;
  if (o.reach[(size_t)sy * W + sx] != SIPP_REACHABILITY_UNKNOWN || o.state[(size_t)sy * W + sx] == SIPP_STATE_OCCUPIED) return;
  std::vector<std::pair<int, int>> stack;
  stack.push_back({sx, sy});
  bool found_other = false;
  const int current_area_id = o.max_area_id + 1;
  while (!stack.empty()) {
    const auto p = stack.back();
    stack.pop_back();
    for (int i = -1; i <= 1; ++i) {
      for (int j = -1; j <= 1; ++j) {
        const int cx = p.first + i, cy = p.second + j;
        if (!(0 <= cy && 0 <= cx && cy < H && cx < W)) continue;
        const size_t c = (size_t)cy * W + cx;
        if (o.state[c] == SIPP_STATE_FREE) {
          const int rs = o.reach[c];
          if (rs == current_area_id) continue;
          if (rs != SIPP_REACHABILITY_UNKNOWN) {
            found_other = true;
            continue;
          }
          o.reach[c] = current_area_id;
          stack.push_back({cx, cy});
        }
      }
    }
  }
  if (!found_other) ++o.max_area_id;
}

// MapServerPlanExp::setClosestObservedMaps — map_server_planexp.cpp:803-823. The reference admits i == W / j == H
// (off-by-one at :809,:811) and may read map_cost_ at closest_x == W; we skip those cells and count them.
void set_closest_observed(Oracle& o, int tlx, int tly, int brx, int bry) {
  const int W = o.W, H = o.H;
  const int off_x = 15, off_y = 10;
  for (int i = tlx - off_x; i <= brx + off_x; ++i) {
    if (i < 0 || W < i) continue;
    for (int j = tly - off_y; j <= bry + off_y; ++j) {
      if (j < 0 || H < j) continue;
      if (i >= W || j >= H) { ++o.oob_events; continue; }
      if (o.state[(size_t)j * W + i] != SIPP_STATE_UNKNOWN) continue;
      int cx = std::max(tlx, i);
      cx = std::min(brx, cx);
      int cy = std::max(tly, j);
      cy = std::min(bry, cy);
      const double distance = std::sqrt(std::pow(i - cx, 2) + std::pow(j - cy, 2));
      if (o.dist_closest[(size_t)j * W + i] <= distance) continue;
      if (cx < 0 || cx >= W || cy < 0 || cy >= H) { ++o.oob_events; continue; }
      o.dist_closest[(size_t)j * W + i] = distance;
      o.cost_closest[(size_t)j * W + i] = o.cost[(size_t)cy * W + cx];
    }
  }
}

// MapServerPlanExp::mapDataCallback — map_server_planexp.cpp:220-282 (everything but ROS / visualisation / thread).
void map_update(Oracle& o, int x0, int y0, int rows, int cols, const int32_t* labels, const uint8_t* observed) {
  const int W = o.W, H = o.H;
  const bool pre_goal = o.state[(size_t)o.goal_loc[1] * W + o.goal_loc[0]] != SIPP_STATE_UNKNOWN;  // goalExplored() :221
  for (int i = 0; i < rows; ++i) {
    for (int j = 0; j < cols; ++j) {
      const int y = y0 + i, x = x0 + j;
      if (!(0 <= y && 0 <= x && y < H && x < W)) continue;                       // :240
      if (observed && !observed[(size_t)i * cols + j]) continue;
      const size_t c = (size_t)y * W + x;
      const int data_point = labels[(size_t)i * cols + j];
      o.cost[c] = (double)data_point;                                             // :242
      if (o.state[c] == SIPP_STATE_UNKNOWN) {                                     // :246-248 (int / int !)
        o.mean_cost = o.mean_cost * ((double)o.mean_counter / (o.mean_counter + 1)) + data_point / (o.mean_counter + 1);
        ++o.mean_counter;
      } else if (o.state[c] == SIPP_STATE_FREE) {                                 // :249-251 (always + 0)
        o.mean_cost = o.mean_cost + (data_point - o.cost[c]) / o.mean_counter;
      }
      o.state[c] = in_ids(o.desc.obstacle_ids_mav, o.desc.n_obstacle_ids_mav, data_point) ? SIPP_STATE_OCCUPIED
                                                                                             : SIPP_STATE_FREE;  // :252-256
      if (data_point > o.max_cost) o.max_cost = data_point;                       // :257
    }
  }
  planner_update_states(o, x0, y0, rows, cols, labels, observed);                 // :262
  if (o.desc.compute_closest_observed && !observed) set_closest_observed(o, x0, y0, x0 + cols, y0 + rows);  // :264
  if (o.desc.compute_reachability) {                                              // :266-275
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < cols; ++j) {
        const int y = y0 + i, x = x0 + j;
        if (!(0 <= y && 0 <= x && y < H && x < W)) continue;
        if (observed && !observed[(size_t)i * cols + j]) continue;
        set_reachability(o, x, y);
      }
  }
  const bool post_goal = o.state[(size_t)o.goal_loc[1] * W + o.goal_loc[0]] != SIPP_STATE_UNKNOWN;
  if (post_goal && !pre_goal && o.init_cost != o.desc.min_rov_cost)               // :276 (unknown_init_cost_ is the cfg value)
    planner_update_unknown_cost(o, o.desc.min_rov_cost);
}

// ----------------------------------------------------------------------------------------------------
// Cell lookups — MapServerPlanExp::get*AtLocation map_server_planexp.cpp:757-801,872-878 and the MapPlanExp
// wrappers map_planexp.cpp:59-66,110-117,210-229.
// ----------------------------------------------------------------------------------------------------
inline bool loc_in_bounds(const Oracle& o, double x, double y) {
  return 0 <= x && 0 <= y && o.desc.width > x && o.desc.height > y;
}
inline size_t loc_index(Oracle& o, double x, double y) {   // (int)(y/mpp), (int)(x/mpp)
  int r = (int)(y / o.mpp), c = (int)(x / o.mpp);
  if (r >= o.H) { r = o.H - 1; ++o.oob_events; }
  if (c >= o.W) { c = o.W - 1; ++o.oob_events; }
  return (size_t)r * o.W + c;
}
inline int status_at(Oracle& o, const Vec3& p) {            // getStatusAtLocation :764-769
  if (loc_in_bounds(o, p.x, p.y)) return o.state[loc_index(o, p.x, p.y)];
  return SIPP_STATE_UNKNOWN;
}
inline bool is_observed(Oracle& o, const Vec3& p) { return SIPP_STATE_UNKNOWN != status_at(o, p); }  // map_planexp.cpp:59-61
inline double cost_at(Oracle& o, const Vec3& p) {           // getCostAtLocation :757-762
  if (loc_in_bounds(o, p.x, p.y)) return o.cost[loc_index(o, p.x, p.y)];
  return -1.0;
}
inline double closest_cost_at(Oracle& o, const Vec3& p) {   // getClosestCostAtLocation :795-801
  if (!o.desc.compute_closest_observed) return -1.0;
  if (loc_in_bounds(o, p.x, p.y)) return o.cost_closest[loc_index(o, p.x, p.y)];
  return -1.0;
}
inline bool reachable_at(Oracle& o, const Vec3& p) {        // getReachabilityAtLocation :872-878 + isReachable map_planexp.cpp:214-216
  if (!o.desc.compute_reachability) return false;
  if (loc_in_bounds(o, p.x, p.y)) return SIPP_REACHABLE == o.reach[loc_index(o, p.x, p.y)];
  return false;
}
inline void pos2pix_voxel(const Oracle& o, double x, double y, int* ix, int* iy) {  // position32pixelLocation map_planexp.cpp:136-138
  *ix = (int)(x * (1 / o.vs));
  *iy = (int)(y * (1 / o.vs));
}
inline bool is_on_map2(const Oracle& o, double lx, double ly) {  // isOnMap2 map_planexp.cpp:222-224
  return (0 <= (int)lx) && (0 <= (int)ly) && (o.W > (int)lx) && (o.H > (int)ly);
}
inline int path_status_at(Oracle& o, const Vec3& p) {       // getPathMapStatusAtPosition map_planexp.cpp:226-229 (no bounds test there)
  int ix, iy;
  pos2pix_voxel(o, p.x, p.y, &ix, &iy);
  if (ix < 0 || ix >= o.W || iy < 0 || iy >= o.H) { ++o.oob_events; return 0; }
  return o.path_mask[(size_t)iy * o.W + ix];
}

// MapPlanExp::getVoxelArea (+ getVoxelCenter) — map_planexp.cpp:72-101
bool get_voxel_area(const Oracle& o, std::vector<Vec3>* out, const Vec3& point, int half_size) {
  int lx, ly;
  pos2pix_voxel(o, point.x, point.y, &lx, &ly);
  const double clx = lx + 0.5, cly = ly + 0.5;                       // :74
  Vec3 centre{clx * o.vs, cly * o.vs, 0.0};                          // pixelLocation2position :142-144
  if (!is_on_map2(o, clx, cly)) return false;                        // :78,:87
  out->push_back(centre);                                            // :88
  int cx, cy;
  pos2pix_voxel(o, centre.x, centre.y, &cx, &cy);                    // :89
  for (int i = -half_size; i <= half_size; ++i) {
    for (int j = -half_size; j <= half_size; ++j) {
      const int nx = cx + i, ny = cy + j;
      Vec3 v{(double)nx * o.vs, (double)ny * o.vs, 0.0};             // pixelLocation2position(Vector2i) :139-141
      if (is_on_map2(o, (double)nx, (double)ny)) out->push_back(v);  // :97
    }
  }
  return true;
}

// [upstream] BoundingVolume::contains — inclusive box test; z (0.0) is inside [z_min,z_max] = [0,0.01] in every
// shipped cfg (cfg/experiments/example.yaml:12-18).
inline bool bv_contains(const sipp_eval_params& p, const Vec3& v) {
  if (!p.use_bounding_volume) return true;
  return v.x >= p.bv_x_min && v.x <= p.bv_x_max && v.y >= p.bv_y_min && v.y <= p.bv_y_max;
}

inline double norm3(double ax, double ay, double az) { return std::sqrt(ax * ax + ay * ay + az * az); }  // Eigen .norm()

// the reference's accidental O(n^2) filter (rrt_star_evaluator.cpp:43-51 and the four copies of it) when
// `faithful`, or the same result through an O(n) pass.
template <class Pred>
inline void erase_observed(std::vector<Vec3>& v, bool faithful, Pred observed) {
  if (faithful) {
    size_t i = 0;
    while (i < v.size()) {
      if (observed(v[i])) v.erase(v.begin() + i);
      else ++i;
    }
  } else {
    v.erase(std::remove_if(v.begin(), v.end(), observed), v.end());
  }
}

// One candidate: [upstream] SimulatedSensorEvaluator::computeGain -> SensorModelPlanExp::getVisibleVoxelsFromTrajectory
// (sensor_model_planexp.cpp:18-63) -> CameraModelPlanExp::getVisibleVoxels (camera_model_planexp.cpp:23-28) ->
// bounding-volume filter -> <Evaluator>::computeGainFromVisibleVoxels.
// `views`: the nv sampled viewpoints of the segment (sampleViewpoints, sensor_model_planexp.cpp:65-90, with the mounting translation
// applied, :28-31) in trajectory order; their voxel lists are concatenated (:34) and only ADJACENT duplicates are removed (:61).
void eval_views(Oracle& o, const sipp_eval_params& p, const double* views, int nv, const double* start_xy, bool faithful,
                double* gain_out, uint32_t* cnt, uint8_t* terminal_out) {
  std::vector<Vec3> raw;
  raw.reserve(((size_t)(2 * p.half_fov + 1) * (2 * p.half_fov + 1) + 1) * (size_t)(nv > 0 ? nv : 1));
  for (int k = 0; k < nv; ++k) get_voxel_area(o, &raw, Vec3{views[2 * k], views[2 * k + 1], 0.0}, p.half_fov);
  raw.erase(std::unique(raw.begin(), raw.end()), raw.end());         // sensor_model_planexp.cpp:61
  std::vector<Vec3> vox;
  vox.reserve(raw.size());
  for (const Vec3& v : raw)
    if (bv_contains(p, v)) vox.push_back(v);
  uint32_t c_vis = (uint32_t)vox.size(), c_unk = 0, c_aux = 0, c_aux2 = 0;
  double gain = 0.0;
  bool terminal = false;
  const Vec3 goal{o.desc.goal_x, o.desc.goal_y, 0.0};                // getGoalPosition map_planexp.cpp:127

  switch (p.evaluator) {
    case SIPP_EVAL_RRT_STAR: {                                       // rrt_star_evaluator.cpp:36-96
      erase_observed(vox, faithful, [&](const Vec3& v) { return is_observed(o, v); });
      c_unk = (uint32_t)vox.size();
      for (const Vec3& v : vox) c_aux += (path_status_at(o, v) == 1);
      if (vox.empty()) { gain = 0.0; terminal = true; break; }       // :53-57
      if (p.do_binary_check) {
        if (p.use_distance_discount) {                               // :61-68 (distance from trajectory[0])
          for (const Vec3& v : vox)
            if (path_status_at(o, v) == 1) {
              gain = 1.0 / std::sqrt(norm3(goal.x - start_xy[0], goal.y - start_xy[1], goal.z - 0.0) + p.discount_offset);
              break;
            }
        } else {                                                     // :69-75
          for (const Vec3& v : vox)
            if (path_status_at(o, v) == 1) { gain = 1.0; break; }
        }
      } else {
        if (p.use_distance_discount) {                               // :78-84
          for (const Vec3& v : vox) {
            if (path_status_at(o, v) == 1) gain += 1.0 / std::sqrt(norm3(goal.x - v.x, goal.y - v.y, goal.z - v.z) + p.discount_offset);
            else gain += p.non_path_voxel_gain;
          }
        } else {                                                     // :85-89
          for (const Vec3& v : vox) {
            if (path_status_at(o, v) == 1) gain += 1.0;
            else gain += p.non_path_voxel_gain;
          }
        }
      }
      break;
    }
    case SIPP_EVAL_GOAL_DISTANCE: {                                  // goal_distance_evaluator.cpp:33-66
      erase_observed(vox, faithful, [&](const Vec3& v) { return is_observed(o, v); });
      c_unk = (uint32_t)vox.size();
      for (const Vec3& v : vox) {
        const double distance = norm3(goal.x - v.x, goal.y - v.y, goal.z - v.z);
        if (distance > 1.0 / p.max_gain) gain += 1.0 / distance;
        else gain += p.max_gain;
      }
      break;
    }
    case SIPP_EVAL_EXPAND_REACHABLE: {                               // expand_reachable_area_evaluator.cpp:33-67
      erase_observed(vox, faithful, [&](const Vec3& v) { return is_observed(o, v); });
      c_unk = (uint32_t)vox.size();
      if (vox.empty()) { gain = 0.0; terminal = true; break; }
      size_t n_reachable = 0;
      for (const Vec3& v : vox) {
        bool is_reachable = false;
        std::vector<Vec3> nb;
        get_voxel_area(o, &nb, v, 1);
        for (const Vec3& q : nb) is_reachable |= reachable_at(o, q);
        if (is_reachable) ++n_reachable;
      }
      c_aux = (uint32_t)n_reachable;
      gain = n_reachable + (vox.size() - n_reachable) * p.discount_factor;   // :65
      break;
    }
    case SIPP_EVAL_AVERAGE_VIEW_COST: {                              // average_view_cost_evaluator.cpp:31-61
      size_t n_observed = 0;
      double mean_view_cost = 0.0;
      auto pred = [&](const Vec3& v) {
        if (is_observed(o, v)) {
          if (status_at(o, v) == SIPP_STATE_FREE) { ++n_observed; mean_view_cost += cost_at(o, v); }
          return true;
        }
        return false;
      };
      if (faithful) {
        size_t i = 0;
        while (i < vox.size()) {
          if (pred(vox[i])) vox.erase(vox.begin() + i);
          else ++i;
        }
      } else {
        std::vector<Vec3> keep;
        for (const Vec3& v : vox)
          if (!pred(v)) keep.push_back(v);
        vox.swap(keep);
      }
      c_unk = (uint32_t)vox.size();
      c_aux = (uint32_t)n_observed;
      c_aux2 = (uint32_t)mean_view_cost;
      if (n_observed > 0) {
        mean_view_cost /= n_observed;
        gain = vox.size() * (1.0 / mean_view_cost);                  // :54-56
      } else {
        gain = vox.size() * (1.0 / o.mean_cost);                     // :57-59 (0 * inf = NaN when mean_cost_ == 0)
      }
      break;
    }
    case SIPP_EVAL_EXPAND_LOW_COST: {                                // expand_low_cost_area_evaluator.cpp:33-72
      erase_observed(vox, faithful, [&](const Vec3& v) { return is_observed(o, v); });
      c_unk = (uint32_t)vox.size();
      for (const Vec3& v : vox) {
        const double c = closest_cost_at(o, v);
        if (c > 0.0) gain += 1.0 / c;
        else gain += p.augment_unknown_with_mean ? 1.0 / o.mean_cost : 0.0;
      }
      break;
    }
    case SIPP_EVAL_NAIVE: {                                          // [upstream] NaiveEvaluator: count of unobserved voxels
      erase_observed(vox, faithful, [&](const Vec3& v) { return is_observed(o, v); });
      c_unk = (uint32_t)vox.size();
      gain = (double)vox.size();
      break;
    }
    default: break;
  }
  *gain_out = gain;
  if (cnt) { cnt[SIPP_CNT_VISIBLE] = c_vis; cnt[SIPP_CNT_UNKNOWN] = c_unk; cnt[SIPP_CNT_AUX] = c_aux; cnt[SIPP_CNT_AUX2] = c_aux2; }
  if (terminal_out) *terminal_out = terminal ? 1 : 0;
}
// one view at the segment's last trajectory point: sampling_time <= 0, every shipped cfg (sensor_model_planexp.cpp:72-75)
void eval_one(Oracle& o, const sipp_eval_params& p, double px, double py, const double* start_xy, bool faithful,
              double* gain_out, uint32_t* cnt, uint8_t* terminal_out) {
  const double v[2] = {px, py};
  eval_views(o, p, v, 1, start_xy, faithful, gain_out, cnt, terminal_out);
}

// ----------------------------------------------------------------------------------------------------
// [upstream] GlobalNormalizedGain::computeValue / findBest and SubsequentBest::selectNextBest / evaluateSingle
// (call sites: advanced_planner_ros/cfg/planners/example_config.yaml:100-106).
// ----------------------------------------------------------------------------------------------------
double find_best(const std::vector<std::vector<int>>& children, const double* gain, const double* cost, int cur, double g, double c) {
  double value = 0.0;
  g += gain[cur];
  c += cost[cur];
  if (c > 0) value = g / c;
  for (int ch : children[cur]) value = std::max(value, find_best(children, gain, cost, ch, g, c));
  return value;
}

}  // namespace

// =====================================================================================================
// extern "C" surface (ctypes)
// =====================================================================================================
extern "C" {

struct orc_ctx { Oracle o; };

orc_ctx* orc_create(const sipp_map_desc* d) {
  if (!d || d->W <= 0 || d->H <= 0 || d->width <= 0 || d->height <= 0) return nullptr;
  orc_ctx* c = new orc_ctx();
  Oracle& o = c->o;
  o.desc = *d;
  o.W = d->W;
  o.H = d->H;
  o.mpp = d->width / (unsigned int)d->W;   // m_per_pixel_  map_server_planexp.cpp:116
  o.vs = d->width / (unsigned int)d->W;    // c_voxel_size_ map_planexp.cpp:42
  o.goal_loc[0] = (int)(d->goal_x / o.mpp);    // :117
  o.goal_loc[1] = (int)(d->goal_y / o.mpp);
  o.start_loc[0] = (int)(d->start_x / o.mpp);  // :119
  o.start_loc[1] = (int)(d->start_y / o.mpp);
  planner_set_derived(o);
  map_clear(o);
  return c;
}
void orc_destroy(orc_ctx* c) { delete c; }

int orc_map_clear(orc_ctx* c) {
  map_clear(c->o);
  planner_set_derived(c->o);
  c->o.path_counter = 0;
  c->o.path_valid = false;
  return 0;
}
int orc_patch_origin(orc_ctx* c, double px, double py, int rows, int cols, int* x0, int* y0) {
  int ix, iy;
  position2pixel(c->o, px, py, &ix, &iy);   // map_server_planexp.cpp:224
  *x0 = ix - (int)(cols / 2);               // :227
  *y0 = iy - (int)(rows / 2);               // :228
  return 0;
}
int orc_map_update_patch(orc_ctx* c, int x0, int y0, int rows, int cols, const int32_t* labels) {
  map_update(c->o, x0, y0, rows, cols, labels, nullptr);
  return 0;
}
int orc_map_upload_full(orc_ctx* c, const int32_t* labels, const uint8_t* observed) {
  map_update(c->o, 0, 0, c->o.H, c->o.W, labels, observed);
  return 0;
}
int orc_update_unknown_cost(orc_ctx* c, double nc) { planner_update_unknown_cost(c->o, nc); return 0; }
int orc_map_stats(orc_ctx* c, double* mean_cost, double* max_cost, int64_t* mean_counter) {
  if (mean_cost) *mean_cost = c->o.mean_cost;
  if (max_cost) *max_cost = c->o.max_cost;
  if (mean_counter) *mean_counter = c->o.mean_counter;
  return 0;
}
int orc_map_download(orc_ctx* c, uint8_t* state, int32_t* cost, uint8_t* reach, uint8_t* path) {
  const Oracle& o = c->o;
  const size_t N = (size_t)o.W * o.H;
  for (size_t i = 0; i < N; ++i) {
    if (state) state[i] = (uint8_t)o.state[i];
    if (cost) cost[i] = (int32_t)o.cost[i];
    if (reach) reach[i] = o.reach[i] == SIPP_REACHABLE;
    if (path) path[i] = (uint8_t)o.path_mask[i];
  }
  return 0;
}
// raw reachability ids (to show that only the start cell ever equals MSR_REACHABLE)
int orc_reach_ids(orc_ctx* c, int32_t* out) {
  for (size_t i = 0; i < c->o.reach.size(); ++i) out[i] = c->o.reach[i];
  return 0;
}
int orc_closest_download(orc_ctx* c, double* cost_closest, double* dist_closest) {
  if (!c->o.desc.compute_closest_observed) return -1;
  for (size_t i = 0; i < c->o.cost_closest.size(); ++i) {
    if (cost_closest) cost_closest[i] = c->o.cost_closest[i];
    if (dist_closest) dist_closest[i] = c->o.dist_closest[i];
  }
  return 0;
}
int orc_planner_download(orc_ctx* c, double* pmap, uint8_t* explored, uint8_t* removed) {
  const size_t N = c->o.pmap.size();
  for (size_t i = 0; i < N; ++i) {
    if (pmap) pmap[i] = c->o.pmap[i];
    if (explored) explored[i] = c->o.explored[i];
    if (removed) removed[i] = c->o.removed[i];
  }
  return 0;
}
int orc_oob_events(orc_ctx* c) { return c->o.oob_events; }

// geometry probes (unit tests of the diagonal predicate)
int orc_geometry(orc_ctx* c, double* spacing, double* off_x, double* off_y, double* max_dist, int32_t nodes[4]) {
  const Oracle& o = c->o;
  if (spacing) *spacing = o.spacing;
  if (off_x) *off_x = o.off_x;
  if (off_y) *off_y = o.off_y;
  if (max_dist) *max_dist = o.max_dist;
  if (nodes) { nodes[0] = o.start_node[0]; nodes[1] = o.start_node[1]; nodes[2] = o.goal_node[0]; nodes[3] = o.goal_node[1]; }
  return 0;
}
int64_t orc_diag_admitted(orc_ctx* c) {
  const Oracle& o = c->o;
  int64_t n = 0;
  for (int i = 0; i + 1 < o.W; ++i)
    for (int j = 0; j + 1 < o.H; ++j)
      n += edge_dist(o, i, j, i + 1, j + 1) < o.max_dist;
  return n;
}

// PRMStarPlanner::plan (prm_star_planner.cpp:38-70) + computeCurrentPathEstimate / computePathMap
// (map_server_planexp.cpp:880-926). update_mask = 0 leaves the current path mask untouched (eval-binary mode).
int orc_plan(orc_ctx* c, double* cost, int32_t* status, double* path_xy, int32_t path_cap, int32_t* path_len, int update_mask) {
  Oracle& o = c->o;
  planner_field(o);
  double cst = -1.0;                                                  // :39
  int st = SIPP_PLAN_TERMINAL_INVALID;                                // :69 (queue exhausted without reaching the goal)
  std::vector<double> pxy;
  const int gx = o.goal_node[0], gy = o.goal_node[1];
  const bool goal_ok = gx >= 0 && gx < o.W && gy >= 0 && gy < o.H && std::isfinite(o.field[(size_t)gy * o.W + gx]);
  const bool walked = goal_ok && planner_backtrack(o);
  if (goal_ok && !walked) st = SIPP_PLAN_INVALID;   // reachable goal, non-terminating predecessor walk (zero-weight plateau): catch-all of :66-68
  if (walked) {
    cst = o.field[(size_t)gy * o.W + gx];                             // :63
    bool all_explored = true;                                         // check_termination :183-186
    for (int id : o.last_path) {
      const int x = id / o.H, y = id % o.H;
      pxy.push_back(o.X[x]);                                          // node_id2pose :62
      pxy.push_back(o.Y[y]);
      if (!o.explored[(size_t)y * o.W + x]) all_explored = false;
    }
    st = all_explored ? SIPP_PLAN_TERMINAL_VALID : SIPP_PLAN_VALID;   // :64-65
  }
  if (cost) *cost = cst;
  if (status) *status = st;
  const int n = (int)(pxy.size() / 2);
  if (path_len) *path_len = n;
  if (path_xy) for (int i = 0; i < std::min(n, path_cap); ++i) { path_xy[2 * i] = pxy[2 * i]; path_xy[2 * i + 1] = pxy[2 * i + 1]; }
  if (update_mask) {
    // computeCurrentPathEstimate :880-900. An empty path would crash the reference at path[0] (:905); we leave a zero mask.
    if (n > 0) compute_path_map(o, pxy, &o.path_mask);
    else o.path_mask.assign((size_t)o.W * o.H, 0);
    o.path_counter += 1;
    o.path_valid = true;
  }
  return 0;
}
int orc_field_download(orc_ctx* c, double* field) {
  if (c->o.field.empty()) return -1;
  std::memcpy(field, c->o.field.data(), c->o.field.size() * sizeof(double));
  return 0;
}
int orc_path_nodes(orc_ctx* c, int32_t* ids, int cap) {
  const int n = (int)c->o.last_path.size();
  for (int i = 0; i < std::min(n, cap); ++i) ids[i] = c->o.last_path[i];
  return n;
}
int orc_path_mask_generation(orc_ctx* c) { return c->o.path_counter; }
int orc_set_path_mask(orc_ctx* c, const uint8_t* mask) {
  for (size_t i = 0; i < c->o.path_mask.size(); ++i) c->o.path_mask[i] = mask[i];
  c->o.path_valid = true;
  return 0;
}

// variant: 0 = faithful (quadratic vector::erase loop), 1 = O(n) filter (B2).
// threads: candidates split in contiguous blocks over std::thread (B3); 1 = the reference's single thread.
int orc_gain_batch(orc_ctx* c, const sipp_eval_params* p, int32_t n, const double* xy, const double* start_xy,
                   double* gain, uint32_t* counts, uint8_t* terminal, int variant, int threads) {
  Oracle& o = c->o;
  if (threads < 1) threads = 1;
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i)
      eval_one(o, *p, xy[2 * i], xy[2 * i + 1], start_xy ? start_xy + 2 * i : xy + 2 * i, variant == 0, gain + i,
               counts ? counts + (size_t)SIPP_COUNTS_PER_CANDIDATE * i : nullptr, terminal ? terminal + i : nullptr);
  };
  if (threads == 1) { work(0, n); return 0; }
  std::vector<std::thread> th;
  const int per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const int lo = t * per, hi = std::min(n, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& t : th) t.join();
  return 0;
}

// SensorModelPlanExp::sampleViewpoints (sensor_model_planexp.cpp:65-90): indices of the trajectory points a segment is viewed from.
// times_ns = time_from_start_ns of its npts points. Returns the number of indices (0 for an empty trajectory).
int orc_sample_viewpoints(const int64_t* times_ns, int32_t npts, double sampling_time, int32_t* out_idx, int32_t cap) {
  if (npts < 1) return 0;                                            // :67-70
  int k = 0;
  auto push = [&](int i) { if (k < cap) out_idx[k] = i; ++k; };
  if (sampling_time <= 0.0) {                                        // :71-73: only the last point
    push(npts - 1);
  } else {
    const int64_t sampling_time_ns = static_cast<int64_t>(sampling_time * 1.0e9);      // :75
    if (sampling_time_ns >= times_ns[npts - 1]) {                    // :76-78: no point within one interval: the last point
      push(npts - 1);
    } else {
      int64_t current_time = sampling_time_ns;                       // :81-87
      for (int i = 0; i < npts; ++i)
        if (times_ns[i] >= current_time) {
          current_time += sampling_time_ns;
          push(i);
        }
    }
  }
  return k;
}

// Segments with several sampled viewpoints (sensor_model/sampling_time > 0): segment i has the views view_xy[view_off[i] .. view_off[i+1]).
int orc_gain_batch_views(orc_ctx* c, const sipp_eval_params* p, int32_t n, const int32_t* view_off, const double* view_xy, const double* start_xy,
                         double* gain, uint32_t* counts, uint8_t* terminal, int variant) {
  Oracle& o = c->o;
  for (int i = 0; i < n; ++i) {
    const int a = view_off[i], nv = view_off[i + 1] - view_off[i];
    const double none[2] = {0.0, 0.0};
    eval_views(o, *p, view_xy + 2 * (size_t)a, nv, start_xy ? start_xy + 2 * i : (nv > 0 ? view_xy + 2 * (size_t)a : none), variant == 0, gain + i,
               counts ? counts + (size_t)SIPP_COUNTS_PER_CANDIDATE * i : nullptr, terminal ? terminal + i : nullptr);
  }
  return 0;
}

// [upstream] GlobalNormalizedGain + SubsequentBest on a parent-index tree (node 0 = root, parent[i] < i).
int orc_value_select(int32_t n, const double* gain_in, const double* cost_in, const int32_t* parent, double* value,
                     int32_t* best_child, double* best_value) {
  if (n < 1) return -1;
  std::vector<double> gain(gain_in, gain_in + n), cost(cost_in, cost_in + n);
  gain[0] = 0.0; cost[0] = 0.0;                       // OnlinePlanner zeroes the current segment [upstream]
  std::vector<std::vector<int>> children(n);
  for (int i = 1; i < n; ++i) children[parent ? parent[i] : 0].push_back(i);
  for (int i = 0; i < n; ++i) {
    double g = 0.0, cst = 0.0;
    int cur = (i == 0) ? -1 : (parent ? parent[i] : 0);
    while (cur >= 0) {                                // computeValue: walk the parents up to the root
      g += gain[cur];
      cst += cost[cur];
      cur = (cur == 0) ? -1 : (parent ? parent[cur] : 0);
    }
    value[i] = find_best(children, gain.data(), cost.data(), i, g, cst);
  }
  int best = -1;
  double bv = 0.0;
  for (int ch : children[0]) {                        // selectNextBest over the root's children; evaluateSingle = subtree max of value
    double v = value[ch];
    std::vector<int> st(children[ch].begin(), children[ch].end());
    while (!st.empty()) { int q = st.back(); st.pop_back(); v = std::max(v, value[q]); for (int r : children[q]) st.push_back(r); }
    if (best < 0 || v > bv) { best = ch; bv = v; }
  }
  if (best_child) *best_child = best;
  if (best_value) *best_value = bv;
  return 0;
}

// Literal emulation of one tree re-evaluation pass of the reference: [upstream] OnlinePlanner::updateEvaluatorStep visits the
// tree in pre-order and calls trajectory_evaluator_->updateSegment(node) -> evaluator_updater chain:
//   RRTStarUpdater::updateSegment            advanced_planner_modules/src/evaluator_updater/rrt_star_updater.cpp:19-42
//   [upstream] ConstrainedUpdater::updateSegment (same predicate without the terminal / path-changed terms)
//   [upstream] UpdateAll::updateSegment / UpdateAllNonTerminal::updateSegment  (update_all_non_terminal.cpp:16-27)
// Nodes are given in pre-order (parent[i] < i, children in index order), node 0 is the root (its gain / cost are zeroed on entry: it just
// became the current segment; whether the root-first call of the patched planner re-scores it depends on the chain, see below).
// gain / value / terminal are in-out; computeValue sees whatever gains the array holds at that moment.
int orc_update_tree_views(orc_ctx* c, const sipp_eval_params* p, const sipp_updater_params* u, int32_t n, const int32_t* parent,
                          const double* end_xy, const double* start_xy, double* gain, const double* cost, double* value, uint8_t* terminal,
                          const double* cur_xy, int path_changed, int variant, const double* cost_below, const int32_t* view_off, const double* view_xy);
int orc_update_tree(orc_ctx* c, const sipp_eval_params* p, const sipp_updater_params* u, int32_t n, const int32_t* parent,
                    const double* end_xy, const double* start_xy, double* gain, const double* cost, double* value, uint8_t* terminal,
                    const double* cur_xy, int path_changed, int variant, const double* cost_below) {
  return orc_update_tree_views(c, p, u, n, parent, end_xy, start_xy, gain, cost, value, terminal, cur_xy, path_changed, variant, cost_below, nullptr, nullptr);
}
// view_off / view_xy (optional): node i is scored from its sampled viewpoints instead of its end point alone
int orc_update_tree_views(orc_ctx* c, const sipp_eval_params* p, const sipp_updater_params* u, int32_t n, const int32_t* parent,
                          const double* end_xy, const double* start_xy, double* gain, const double* cost, double* value, uint8_t* terminal,
                          const double* cur_xy, int path_changed, int variant, const double* cost_below, const int32_t* view_off, const double* view_xy) {
  Oracle& o = c->o;
  std::vector<std::vector<int>> children(n);
  for (int i = 1; i < n; ++i) children[parent[i]].push_back(i);
  // cost_below != NULL: update_cost — computeCost(i) runs before computeValue(i) (update_all_non_terminal.cpp:20-22): `cost` holds the
  // costs the pass computes, cost_below the costs the tree carried before it; node i switches from the old to the new cost when visited
  std::vector<double> g(gain, gain + n), cst(cost_below ? cost_below : cost, (cost_below ? cost_below : cost) + n);
  g[0] = 0.0; cst[0] = 0.0;
  if (u->updater != SIPP_UPD_RRT_STAR) value[0] = 0.0;     // the new current segment enters with gain = cost = value = 0 [upstream requestNextTrajectory]
  const bool setsTerminal = p->evaluator == SIPP_EVAL_RRT_STAR || p->evaluator == SIPP_EVAL_EXPAND_REACHABLE;
  auto compute_gain = [&](int i) {
    double gi; uint8_t ti;
    if (view_off)
      eval_views(o, *p, view_xy + 2 * (size_t)view_off[i], view_off[i + 1] - view_off[i], start_xy ? start_xy + 2 * i : end_xy + 2 * i, variant == 0, &gi, nullptr, &ti);
    else
      eval_one(o, *p, end_xy[2 * i], end_xy[2 * i + 1], start_xy ? start_xy + 2 * i : end_xy + 2 * i, variant == 0, &gi, nullptr, &ti);
    g[i] = gi;
    if (setsTerminal && ti) terminal[i] = 1;      // the evaluators only ever set gain_terminal, never clear it
  };
  auto compute_value = [&](int i) {
    double gs = 0.0, cs = 0.0;
    for (int cur = parent[i]; cur >= 0; cur = (cur == 0) ? -1 : parent[cur]) { gs += g[cur]; cs += cst[cur]; }
    value[i] = find_best(children, g.data(), cst.data(), i, gs, cs);
  };
  auto following = [&](int i) {
    if (u->following == SIPP_FOLLOW_UPDATE_ALL) {
      if (u->update_gain) compute_gain(i);
      if (u->update_cost && cost_below) cst[i] = cost[i];
      if (u->update_value) compute_value(i);
    } else {
      if (u->update_gain && !terminal[i]) compute_gain(i);
      if (u->update_cost && cost_below) cst[i] = cost[i];
      if (u->update_value && !(terminal[i] && !u->update_cost)) compute_value(i);
    }
  };
  // The patched planner calls trajectory_evaluator_->updateSegment(root) first (mav_active_3d_planning.patch:56), then [upstream]
  // updateEvaluatorStep visits every node BELOW the root in pre-order. RRTStarUpdater only inspects the path id on its root visit
  // (rrt_star_updater.cpp:21,30-38: "cannot update the root segment"); a ConstrainedUpdater / a following updater used directly
  // treats the root like any node, so the root's gain (zeroed when it became the current segment) is recomputed and enters every
  // value below it. Pinned against the reference's own sources by tests/test_ref_pin.py.
  for (int i = 0; i < n; ++i) {
    if (i == 0 && u->updater == SIPP_UPD_RRT_STAR) continue;
    const double dx = cur_xy[0] - end_xy[2 * i], dy = cur_xy[1] - end_xy[2 * i + 1];
    const double dist = std::sqrt(dx * dx + dy * dy + 0.0 * 0.0);
    if (u->updater == SIPP_UPD_RRT_STAR) {
      if (terminal[i]) continue;
      if (g[i] > u->minimum_gain || path_changed)
        if (dist <= u->update_range || path_changed) following(i);
    } else if (u->updater == SIPP_UPD_CONSTRAINED) {
      if (g[i] > u->minimum_gain)
        if (dist <= u->update_range) following(i);
    } else {
      following(i);                 // cfg/evaluators/expand_reachable_area.yaml:23-27: no gating updater in front
    }
  }
  for (int i = (u->updater == SIPP_UPD_RRT_STAR ? 1 : 0); i < n; ++i) gain[i] = g[i];
  return 0;
}

// flat-tree cycle: candidate i is node i+1 under a virtual root.
// CostPlanExpMapIntegral::computeCost (advanced_planner_modules/src/cost_computer/cost_planexp_map_integral.cpp:21-73): line integral of
// the map cost along the piecewise-linear trajectory, sampled every max_sample_distance; a sample pair touching an unobserved cell
// (weight 0) is charged the mean map cost; off-map samples read -1 (getCostAtLocation :757-762) and are used as they are.
// Segment i has points pts[off[i] .. off[i+1]) (x, y, z); valid[i] = the reference's return value (false for < 2 points).
int orc_cost_integral(orc_ctx* c, const sipp_cost_params* cp, int32_t n, const int32_t* off, const double* pts, const double* parent_cost,
                      double* cost_out, uint8_t* valid) {
  Oracle& o = c->o;
  const double msd = cp->max_sample_distance;
  auto weight = [&](const Vec3& p) { return cost_at(o, p); };                       // MapPlanExp::getVoxelWeight map_planexp.cpp:110-112
  auto norm3 = [](const Vec3& a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); };
  for (int s = 0; s < n; ++s) {
    const int np = off[s + 1] - off[s];
    const double* P = pts + 3 * (size_t)off[s];
    if (np < 2) {                                                                    // :22-25
      cost_out[s] = 0.0;
      if (valid) valid[s] = 0;
      continue;
    }
    double cost = 0.0, segment_cost = 0.0;
    Vec3 last{P[0], P[1], P[2]};
    for (int i = 1; i < np; ++i) {
      const Vec3 pi{P[3 * i], P[3 * i + 1], P[3 * i + 2]};
      const Vec3 diff{pi.x - last.x, pi.y - last.y, pi.z - last.z};
      const double segment_distance = norm3(diff);                                   // :37
      if (segment_distance > 0.005) {
        const Vec3 dir{diff.x / segment_distance, diff.y / segment_distance, diff.z / segment_distance};   // :39
        const int n_sub = (int)(segment_distance / msd);                             // :40
        segment_cost = 0;
        Vec3 start = last, end = last;
        for (int j = 0; j <= n_sub; ++j) {
          double sub = msd;
          if (j == n_sub) sub = segment_distance - n_sub * msd;                      // :46
          end.x += dir.x * sub; end.y += dir.y * sub; end.z += dir.z * sub;          // :47
          const double ws = weight(start), we = weight(end);
          double sub_cost = sub * (ws + we) / 2.0;                                   // :48-49
          if (ws == 0 || we == 0) sub_cost = sub * o.mean_cost;                      // :50-51
          segment_cost += sub_cost;
          start = end;
        }
      } else {                                                                       // :55-61
        const double w = weight(pi);
        segment_cost = (w == 0) ? segment_distance * o.mean_cost : segment_distance * w;
      }
      cost += segment_cost;
      last = pi;
    }
    if (cp->accumulate && parent_cost) cost += parent_cost[s];                       // :68-70
    cost_out[s] = cost;
    if (valid) valid[s] = 1;
  }
  return 0;
}

// MavSimulator::computeCost (mav_simulator/src/mav_simulator.cpp:785-821) on the simulator's ground-truth label map `truth`
// (W x H int32 row-major; the reference's map_ is an Eigen::ArrayXXi indexed (row = y, col = x) through AM(), mav_simulator.h:15,112).
// pixel = ((int)(x * cols / map_width), (int)(y * rows / map_height)) (:582-587). A pixel outside the array is undefined behaviour in
// the reference (unchecked Eigen index): here it reads 0 and is counted in oob_events.
int orc_travel_cost(orc_ctx* c, const int32_t* truth, const sipp_travel_params* tp, int32_t n, const double* start_xyz, const double* goal_xyz,
                    double* cost_out) {
  Oracle& o = c->o;
  auto label_at = [&](double x, double y) -> int {
    const double fx = x * (long)o.W / o.desc.width, fy = y * (long)o.H / o.desc.height;
    if (!(fx > -2.0e9 && fx < 2.0e9 && fy > -2.0e9 && fy < 2.0e9)) { ++o.oob_events; return 0; }
    const int ix = (int)fx, iy = (int)fy;
    if (ix < 0 || ix >= o.W || iy < 0 || iy >= o.H) { ++o.oob_events; return 0; }
    return truth[(size_t)iy * o.W + ix];
  };
  for (int s = 0; s < n; ++s) {
    const double sx = start_xyz[3 * s], sy = start_xyz[3 * s + 1], sz = start_xyz[3 * s + 2];
    const double dx = goal_xyz[3 * s] - sx, dy = goal_xyz[3 * s + 1] - sy, dz = goal_xyz[3 * s + 2] - sz;
    double cost = 0.0;
    const double distance = std::sqrt(dx * dx + dy * dy + dz * dz);                  // :788 (goal - start).norm()
    const double q = distance / tp->max_step_size;
    const int n_elements = (q < 2.0e9) ? (int)std::ceil(q) : 0;                       // :789
    if (n_elements > 0) {
      const double step_distance = distance / n_elements;                             // :791
      const double stx = dx / n_elements, sty = dy / n_elements, stz = dz / n_elements;   // :792
      double cx = sx, cy = sy, cz = sz;
      for (int i = 0; i < n_elements; ++i) {
        const double tx = cx + stx, ty = cy + sty, tz = cz + stz;                     // :796
        const int ls = label_at(cx, cy), le = label_at(tx, ty);                       // :797-798
        const double average_segment_cost = (ls + le) / 2.0;                          // :799
        switch (tp->formulation) {
          case SIPP_TRAVEL_MAP_COST: cost += average_segment_cost * step_distance; break;     // :805
          case SIPP_TRAVEL_DISTANCE_COST: cost += step_distance; break;                       // :808
          case SIPP_TRAVEL_TIME_COST: cost += (step_distance / tp->mav_speed); break;         // :811
          default: break;
        }
        cx = tx; cy = ty; cz = tz;                                                    // :817
      }
    }
    cost_out[s] = cost;
  }
  return 0;
}

int orc_cycle(orc_ctx* c, const sipp_eval_params* p, int32_t n, const double* xy, const double* start_xy, const double* cost,
              double* gain, uint8_t* terminal, double* value, int32_t* best_index, double* best_value, int variant, int threads) {
  orc_gain_batch(c, p, n, xy, start_xy, gain, nullptr, terminal, variant, threads);
  int best = -1;
  double bv = 0.0;
  for (int i = 0; i < n; ++i) {
    double g = 0.0, cst = 0.0;   // root gain/cost = 0
    g += gain[i];
    cst += cost[i];
    const double v = cst > 0 ? g / cst : 0.0;
    if (value) value[i] = v;
    if (best < 0 || v > bv) { best = i; bv = v; }
  }
  if (best_index) *best_index = best;
  if (best_value) *best_value = bv;
  return 0;
}

}  // extern "C"
