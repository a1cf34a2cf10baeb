// =====================================================================================================
// active_3d_planning_shim.h — header-only STAND-IN for the interfaces of ethz-asl/mav_active_3d_planning
// (pinned 4c139dcb + mav_active_3d_planning.patch) and planexp_base_planners::PlannerBase that the scouting-ipp
// modules derive from. The upstream package, Eigen and ROS are not available in this build environment, so the GPU
// modules in ../sipp_modules.{h,cpp} are compiled and tested against this shim; in a catkin workspace the same
// sources build against the real headers by defining SIPP_WITH_UPSTREAM.
//
// Only what the modules touch is declared, with the upstream names and signatures:
//   Module / Module::ParamMap / setParam<T>            active_3d_planning_core/module/module.h           [upstream]
//   ModuleFactoryRegistry::Registration<T>("Name")     active_3d_planning_core/module/module_factory_registry.h [upstream]
//   TrajectorySegment (+ gain_terminal)                data/trajectory_segment.h + mav_active_3d_planning.patch:9,36,44
//   TrajectoryEvaluator / EvaluatorUpdater / Map       module/trajectory_evaluator.h, map/map.h (+ patch :23)
//   planexp_base_planners::PlannerBase / PlannerParams  planexp_base_planners/include/planexp_base_planners/base_planner.h:13-71
// Eigen::Vector3d / Vector2d / ArrayXXi are replaced by minimal PODs with the same member access used by the modules.
// =====================================================================================================
#ifndef SIPP_ACTIVE_3D_PLANNING_SHIM_H_
#define SIPP_ACTIVE_3D_PLANNING_SHIM_H_

#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace Eigen {   // the two fixed-size vectors the module interfaces mention
struct Vector3d {
  double v[3];
  Vector3d() : v{0, 0, 0} {}
  Vector3d(double x, double y, double z) : v{x, y, z} {}
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  Vector3d operator-(const Vector3d& o) const { return Vector3d(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
// unit quaternion (w, x, y, z): the orientation of a trajectory point; operator* rotates a vector like Eigen::Quaterniond does
struct Quaterniond {
  double qw, qx, qy, qz;
  Quaterniond() : qw(1), qx(0), qy(0), qz(0) {}
  Quaterniond(double w, double x, double y, double z) : qw(w), qx(x), qy(y), qz(z) {}
  double w() const { return qw; }
  double x() const { return qx; }
  double y() const { return qy; }
  double z() const { return qz; }
  Vector3d operator*(const Vector3d& p) const {     // v + 2 w (u x v) + 2 u x (u x v), u = (x, y, z)
    const double ux = qx, uy = qy, uz = qz;
    const double cx = uy * p.v[2] - uz * p.v[1], cy = uz * p.v[0] - ux * p.v[2], cz = ux * p.v[1] - uy * p.v[0];
    const double dx = uy * cz - uz * cy, dy = uz * cx - ux * cz, dz = ux * cy - uy * cx;
    return Vector3d(p.v[0] + 2.0 * (qw * cx + dx), p.v[1] + 2.0 * (qw * cy + dy), p.v[2] + 2.0 * (qw * cz + dz));
  }
};
struct Vector2d {
  double v[2];
  Vector2d() : v{0, 0} {}
  Vector2d(double x, double y) : v{x, y} {}
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
};
struct Vector2i {
  int v[2];
  Vector2i() : v{0, 0} {}
  Vector2i(int x, int y) : v{x, y} {}
  int x() const { return v[0]; }
  int y() const { return v[1]; }
};
// row-major int matrix with (row, col) access — stands in for the Eigen::ArrayXXi patches of update_states
struct ArrayXXi {
  int r = 0, c = 0;
  std::vector<int> d;
  ArrayXXi() {}
  ArrayXXi(int rows_, int cols_) : r(rows_), c(cols_), d((size_t)rows_ * cols_, 0) {}
  int rows() const { return r; }
  int cols() const { return c; }
  int& operator()(int i, int j) { return d[(size_t)i * c + j]; }
  int operator()(int i, int j) const { return d[(size_t)i * c + j]; }
};
}  // namespace Eigen

namespace active_3d_planning {

struct EigenTrajectoryPoint {           // mav_msgs::EigenTrajectoryPoint subset
  Eigen::Vector3d position_W;
  Eigen::Quaterniond orientation_W_B;
  int64_t time_from_start_ns = 0;
};
typedef std::vector<EigenTrajectoryPoint> EigenTrajectoryPointVector;

struct TrajectoryInfo {
  virtual ~TrajectoryInfo() = default;
};

struct TrajectorySegment {
  EigenTrajectoryPointVector trajectory;
  double gain = 0.0, cost = 0.0, value = 0.0;
  bool tg_visited = false;
  bool gain_terminal = false;           // mav_active_3d_planning.patch:9
  TrajectorySegment* parent = nullptr;
  std::vector<std::unique_ptr<TrajectorySegment>> children;
  std::unique_ptr<TrajectoryInfo> info;
  TrajectorySegment* spawnChild() {
    children.push_back(std::unique_ptr<TrajectorySegment>(new TrajectorySegment()));
    children.back()->parent = this;
    return children.back().get();
  }
};

class PlannerI;

class Module {
 public:
  typedef std::map<std::string, std::string> ParamMap;
  explicit Module(PlannerI& planner) : planner_(planner) {}
  virtual ~Module() = default;
  virtual void setupFromParamMap(ParamMap* param_map) = 0;
  virtual bool checkParamsValid(std::string* error_message) { return true; }

 protected:
  PlannerI& planner_;
  bool verbose_modules_ = false;
  template <typename T>
  void setParam(ParamMap* map, const std::string& name, T* param, const T& default_value) {
    auto it = map->find(name);
    if (it == map->end()) { *param = default_value; return; }
    std::istringstream is(it->second);
    if (std::is_same<T, bool>::value) is >> std::boolalpha;
    T v = default_value;
    if (!(is >> v)) {   // "1"/"0" for bools
      std::istringstream is2(it->second);
      is2 >> v;
    }
    *param = v;
  }
};
template <>
inline void Module::setParam<std::string>(ParamMap* map, const std::string& name, std::string* param, const std::string& default_value) {
  auto it = map->find(name);
  *param = it == map->end() ? default_value : it->second;
}

namespace map_i {}
class Map : public Module {
 public:
  explicit Map(PlannerI& planner) : Module(planner) {}
  virtual bool isObserved(const Eigen::Vector3d& point) = 0;
  virtual void targetReachedCallback() = 0;   // mav_active_3d_planning.patch:23
};

class TrajectoryEvaluator : public Module {
 public:
  explicit TrajectoryEvaluator(PlannerI& planner) : Module(planner) {}
  virtual bool computeGain(TrajectorySegment* traj_in) = 0;
  virtual bool computeCost(TrajectorySegment* traj_in) = 0;
  virtual bool computeValue(TrajectorySegment* traj_in) = 0;
  virtual int selectNextBest(TrajectorySegment* traj_in) = 0;
  virtual bool updateSegment(TrajectorySegment* segment) = 0;
};

class EvaluatorUpdater : public Module {
 public:
  explicit EvaluatorUpdater(PlannerI& planner) : Module(planner) {}
  virtual bool updateSegment(TrajectorySegment* segment) = 0;
};

class PlannerI {
 public:
  virtual ~PlannerI() = default;
  virtual Map& getMap() = 0;
  virtual TrajectoryEvaluator& getTrajectoryEvaluator() = 0;
  virtual const Eigen::Vector3d& getCurrentPosition() const = 0;
  virtual void printError(const std::string& text) = 0;
};

// type-string registry: static `Registration<T> T::registration("Name")` objects, exactly like the reference modules
class ModuleFactoryRegistry {
 public:
  typedef std::function<Module*(PlannerI&)> Creator;
  static std::map<std::string, Creator>& table() {
    static std::map<std::string, Creator> t;
    return t;
  }
  template <class T>
  struct Registration {
    explicit Registration(const std::string& type) {
      table()[type] = [](PlannerI& planner) -> Module* { return new T(planner); };
    }
  };
  // createModule: looks up (*param_map)["type"], constructs, calls setupFromParamMap
  template <class Base>
  static std::unique_ptr<Base> createModule(Module::ParamMap* param_map, PlannerI& planner, std::string* error) {
    auto ti = param_map->find("type");
    if (ti == param_map->end()) { if (error) *error = "no 'type' in param map"; return nullptr; }
    auto it = table().find(ti->second);
    if (it == table().end()) { if (error) *error = "unknown module type '" + ti->second + "'"; return nullptr; }
    std::unique_ptr<Module> m(it->second(planner));
    Base* b = dynamic_cast<Base*>(m.get());
    if (!b) { if (error) *error = "module '" + ti->second + "' has the wrong base type"; return nullptr; }
    m.release();
    std::unique_ptr<Base> out(b);
    out->setupFromParamMap(param_map);
    std::string msg;
    if (!out->checkParamsValid(&msg)) { if (error) *error = msg; return nullptr; }
    return out;
  }
};

}  // namespace active_3d_planning

namespace planexp_base_planners {   // base_planner.h:13-71

struct PlannerParams {
  int px_x;
  int px_y;
  double width;
  double height;
  double min_cost;
  double max_cost;
  double init_cost;
  Eigen::Vector2d start;
  Eigen::Vector2d goal;
  std::vector<int> obstacle_ids;
  int unknown_id;
  double time_limit = 10.0;
  double goal_bias = 0.05;
  double step_range = 10.0;
};

enum ResultStatus { VALID, INVALID, TERMINAL_VALID, TERMINAL_INVALID };
enum PlannerType { UNKNOWN_PLANNER, RRT_STAR_PLANNER, PRM_STAR_PLANNER };

class PlannerBase {
 public:
  explicit PlannerBase(PlannerParams params) : params_(params) {}
  virtual ~PlannerBase() = default;
  virtual void setup() = 0;
  virtual ResultStatus plan(std::vector<Eigen::Vector2d>* path, double& cost) = 0;
  virtual void store_path(const std::string file_name, const std::string path) = 0;
  virtual void update_states(Eigen::Vector2i location, Eigen::ArrayXXi map_data) = 0;
  virtual void update_unknown_cost(double new_cost) = 0;
  PlannerType get_planner_type() { return PlannerType::UNKNOWN_PLANNER; }

 protected:
  PlannerParams params_;
};

}  // namespace planexp_base_planners

#endif  // SIPP_ACTIVE_3D_PLANNING_SHIM_H_
