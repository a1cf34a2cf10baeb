// Stand-in for map_server_planexp/include/map_server_planexp/map_server_planexp.h (TEST INFRASTRUCTURE, see ../README.md).
// The real MapServerPlanExp is a ROS node (subscribers, OpenCV, grid_map, yaml-cpp, the Boost.Graph planner) and cannot be built
// in this image. This class keeps its query interface — the only part the advanced_planner_modules sources call — over planes the
// test driver fills in. The lookups restate map_server_planexp/src/map_server_planexp.cpp:757-769 (cost / status), :795-801
// (closest cost), :872-878 (reachability), :928-936 (path map); constants from map_server_planexp.h:36-46.
#ifndef SIPP_SHIM_MAP_SERVER_PLANEXP_H_
#define SIPP_SHIM_MAP_SERVER_PLANEXP_H_
#include <memory>
#include <string>

#include "../sipp_eigen_lite.h"
#include "../ros/ros.h"

#define AM(x_) (x_)[1], (x_)[0]
#define MSR_OCCUPIED 0
#define MSR_FREE 1
#define MSR_UNKNOWN 2
#define MSR_REACHABILITY_UNKNOWN 0
#define MSR_REACHABLE 1

typedef Eigen::ArrayXXd CostMap_t;
typedef Eigen::ArrayXXi OccupancyMap_t;
typedef Eigen::ArrayXXd DistanceMap_t;
typedef Eigen::ArrayXXi ReachabilityMap_t;

namespace sipp_ref {
// what the driver hands to the next MapServerPlanExp that is constructed (the real one reads its map cfg from ROS parameters)
struct MapSpec {
  unsigned int horizontal_units = 0, vertical_units = 0;
  double width = 0, height = 0, m_per_pixel = 0;
  double mean_cost = 0, max_cost = 0;
  Eigen::Vector3d start_position, goal_position;
  bool compute_reachability = true, compute_closest_observed = false, compute_path = true;
  std::shared_ptr<CostMap_t> map_cost, map_cost_closest;
  std::shared_ptr<OccupancyMap_t> map_state;
  std::shared_ptr<ReachabilityMap_t> map_reachability;
  std::shared_ptr<Eigen::ArrayXXi> path_map;
  int path_counter = 0;
};
inline MapSpec*& current_map_spec() {
  static MapSpec* p = nullptr;
  return p;
}
}  // namespace sipp_ref

class MapServerPlanExp {
 public:
  MapServerPlanExp(const ros::NodeHandle&, const ros::NodeHandle&) : s_(sipp_ref::current_map_spec()) {}

  double getMapWidth() const { return s_->width; }
  double getMapHeight() const { return s_->height; }
  double getMapResolutionH() const { return s_->horizontal_units; }
  double getMapResolutionV() const { return s_->vertical_units; }
  double getCellSize() const { return s_->m_per_pixel; }
  double getMeanCost() const { return s_->mean_cost; }
  double getMaxCost() const { return s_->max_cost; }
  Eigen::Vector3d getGoalPosition() const { return s_->goal_position; }
  Eigen::Vector3d getStartPosition() const { return s_->start_position; }

  double getCostAtLocation(Eigen::Vector2d position) {                                  // :757-762
    if (0 <= position[0] && 0 <= position[1] && s_->width > position[0] && s_->height > position[1])
      return (*s_->map_cost)((int)(position[1] / s_->m_per_pixel), (int)(position[0] / s_->m_per_pixel));
    return -1.0;
  }
  int getStatusAtLocation(Eigen::Vector2d position) {                                   // :764-769
    if (0 <= position[0] && 0 <= position[1] && s_->width > position[0] && s_->height > position[1])
      return (*s_->map_state)((int)(position[1] / s_->m_per_pixel), (int)(position[0] / s_->m_per_pixel));
    return MSR_UNKNOWN;
  }
  double getClosestCostAtLocation(Eigen::Vector2d position) {                           // :795-801
    if (!s_->compute_closest_observed) return -1.0;
    if (0 <= position[0] && 0 <= position[1] && s_->width > position[0] && s_->height > position[1])
      return (*s_->map_cost_closest)((int)(position[1] / s_->m_per_pixel), (int)(position[0] / s_->m_per_pixel));
    return -1.0;
  }
  bool getReachabilityAtLocation(Eigen::Vector2d position) {                            // :872-878
    if (!s_->compute_reachability) return false;
    if (0 <= position[0] && 0 <= position[1] && s_->width > position[0] && s_->height > position[1])
      return MSR_REACHABLE == ((*s_->map_reachability)((int)(position[1] / s_->m_per_pixel), (int)(position[0] / s_->m_per_pixel)));
    return false;
  }
  Eigen::ArrayXXi getCurrentPathMap() {                                                 // :928-936
    if (s_->compute_path && s_->path_map) return *s_->path_map;
    return Eigen::ArrayXXi();
  }
  int getCurrentPathMapId() const { return s_->path_counter; }                          // map_server_planexp.h:103
  void request_path() {}

 private:
  sipp_ref::MapSpec* s_;
};
#endif
