// Stand-in for the interfaces of ethz-asl/mav_active_3d_planning (pinned 4c139dcb, docker/scouting_ipp.Dockerfile:41-44, patched by
// mav_active_3d_planning.patch) that the reference's advanced_planner_modules sources are written against. That repository is NOT
// vendored under /root/reference, so these declarations are RESTATED from its public interface — only as far as the reference's
// hot-path sources use them (call sites cited per class). TEST INFRASTRUCTURE: nothing under scouting-ipp_b200/ includes this;
// it exists so that oracle/_ref can compile the reference's own .cpp files unmodified (oracle/Makefile, target _ref).
#ifndef SIPP_SHIM_ACTIVE_3D_PLANNING_CORE_H_
#define SIPP_SHIM_ACTIVE_3D_PLANNING_CORE_H_
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../sipp_eigen_lite.h"

namespace active_3d_planning {

// ---- data ------------------------------------------------------------------------------------------------------------------------
// mav_msgs::EigenTrajectoryPoint as used through TrajectorySegment::trajectory (rrt_star_evaluator.cpp:65, rrt_star_updater.cpp:26,
// sensor_model_planexp.cpp:27-31,75-86, cost_planexp_map_integral.cpp:35-63)
struct EigenTrajectoryPoint {
  Eigen::Vector3d position_W, velocity_W, acceleration_W;
  Eigen::Quaterniond orientation_W_B;
  Eigen::Vector3d angular_velocity_W;
  int64_t time_from_start_ns = 0;
};
typedef std::vector<EigenTrajectoryPoint> EigenTrajectoryPointVector;

struct TrajectoryInfo {
  virtual ~TrajectoryInfo() = default;
};
struct SimulatedSensorInfo : public TrajectoryInfo {      // simulated_sensor_evaluator.h [upstream]
  std::vector<Eigen::Vector3d> visible_voxels;
};

// data/trajectory_segment.h [upstream] + gain_terminal (mav_active_3d_planning.patch:9,36,44)
struct TrajectorySegment {
  TrajectorySegment() : gain(0.0), cost(0.0), value(0.0), tg_visited(false), gain_terminal(false), parent(nullptr) {}
  EigenTrajectoryPointVector trajectory;
  double gain, cost, value;
  std::unique_ptr<TrajectoryInfo> info;
  bool tg_visited;
  bool gain_terminal;
  TrajectorySegment* parent;
  std::vector<std::unique_ptr<TrajectorySegment>> children;
  TrajectorySegment* spawnChild() {
    children.push_back(std::unique_ptr<TrajectorySegment>(new TrajectorySegment()));
    children.back()->parent = this;
    return children.back().get();
  }
};

struct SystemConstraints {
  double v_max = 1.0, a_max = 1.0, yaw_rate_max = 1.57, yaw_accel_max = 1.57, collision_radius = 0.35;
};

struct Color { double r = 0, g = 0, b = 0, a = 1; };
struct VisualizationMarker {       // data/visualization_markers.h [upstream]; sensor_model_planexp.cpp:108-157
  enum { POINTS = 0, LINE_LIST = 1, CUBE_LIST = 2 };
  int type = 0, action = 0, id = 0;
  Eigen::Vector3d position, scale;
  Eigen::Quaterniond orientation;
  Color color;
  std::vector<Eigen::Vector3d> points;
  std::vector<Color> colors;
};
class VisualizationMarkers {
 public:
  void addMarker(const VisualizationMarker& m) { markers_.push_back(m); }
  const std::vector<VisualizationMarker>& getMarkers() const { return markers_; }
 private:
  std::vector<VisualizationMarker> markers_;
};

class PlannerI;
class ModuleFactory;

// ---- module.h [upstream]: base of everything the factory creates -------------------------------------------------------------------
class Module {
 public:
  typedef std::map<std::string, std::string> ParamMap;
  explicit Module(PlannerI& planner) : planner_(planner), verbose_modules_(false) {}
  virtual ~Module() = default;
  virtual void setupFromParamMap(ParamMap* param_map) = 0;
  virtual bool checkParamsValid(std::string* error_message) { (void)error_message; return true; }
  void setVerbose(bool verbose) { verbose_modules_ = verbose; }

 protected:
  PlannerI& planner_;
  bool verbose_modules_;

  // look the key up in the map, else take the default (module.h [upstream]; every setupFromParamMap of the reference)
  template <typename T>
  void setParam(ParamMap* map, const std::string& param_name, T* param, const T& default_value) {
    auto it = map->find(param_name);
    if (it == map->end()) {
      *param = default_value;
      return;
    }
    parse(it->second, param);
  }

 private:
  template <typename T>
  static void parse(const std::string& s, T* out) {
    std::istringstream is(s);
    is >> *out;
  }
  static void parse(const std::string& s, bool* out) { *out = (s == "true" || s == "True" || s == "1"); }
  static void parse(const std::string& s, std::string* out) { *out = s; }
};

// ---- module_factory_registry.h [upstream]: static Registration<T> objects register a creator under the module's type string ------
class ModuleFactoryRegistry {
 public:
  typedef std::function<Module*(PlannerI&)> Creator;
  static ModuleFactoryRegistry& Instance() {
    static ModuleFactoryRegistry r;
    return r;
  }
  template <class T>
  struct Registration {
    explicit Registration(const std::string& type) {
      ModuleFactoryRegistry::Instance().creators_[type] = [](PlannerI& planner) -> Module* { return new T(planner); };
    }
  };
  Module* create(const std::string& type, PlannerI& planner) const {
    auto it = creators_.find(type);
    return it == creators_.end() ? nullptr : it->second(planner);
  }
  bool has(const std::string& type) const { return creators_.count(type) != 0; }

 private:
  std::map<std::string, Creator> creators_;
};

// ---- module_factory.h [upstream]: createModule<T>(param namespace, planner, verbose) ------------------------------------------------
// The real factory reads the namespace from the ROS parameter server; here the driver fills `namespaces` (namespace -> ParamMap).
class ModuleFactory {
 public:
  std::map<std::string, Module::ParamMap> namespaces;
  template <class T>
  std::unique_ptr<T> createModule(const std::string& args, PlannerI& planner, bool verbose) {
    Module::ParamMap pm;
    auto it = namespaces.find(args);
    if (it != namespaces.end()) pm = it->second;
    pm["param_namespace"] = args;
    std::string type = pm.count("type") ? pm["type"] : default_type<T>();
    Module* m = ModuleFactoryRegistry::Instance().create(type, planner);
    T* t = dynamic_cast<T*>(m);
    if (!t) {
      delete m;
      std::cerr << "[ref shim] cannot create module '" << type << "' for namespace '" << args << "'" << std::endl;
      return std::unique_ptr<T>();
    }
    t->setVerbose(verbose);
    t->setupFromParamMap(&pm);
    return std::unique_ptr<T>(t);
  }

 private:
  template <class T>
  static std::string default_type();
};

// ---- bounding volume [upstream] (cfg/experiments/example.yaml:12-18): inclusive box, active only when set up -------------------------
class BoundingVolume : public Module {
 public:
  explicit BoundingVolume(PlannerI& planner) : Module(planner), is_setup(false) {}
  bool is_setup;
  double x_min = 0, x_max = 0, y_min = 0, y_max = 0, z_min = 0, z_max = 0;
  bool contains(const Eigen::Vector3d& point) {
    if (!is_setup) return true;   // uninitialised boxes always return true [upstream]
    if (point.x() < x_min) return false;
    if (point.x() > x_max) return false;
    if (point.y() < y_min) return false;
    if (point.y() > y_max) return false;
    if (point.z() < z_min) return false;
    if (point.z() > z_max) return false;
    return true;
  }
  void setupFromParamMap(ParamMap* param_map) override {
    setParam<double>(param_map, "x_min", &x_min, 0.0);
    setParam<double>(param_map, "x_max", &x_max, 0.0);
    setParam<double>(param_map, "y_min", &y_min, 0.0);
    setParam<double>(param_map, "y_max", &y_max, 0.0);
    setParam<double>(param_map, "z_min", &z_min, 0.0);
    setParam<double>(param_map, "z_max", &z_max, 0.0);
    is_setup = x_max != x_min && y_max != y_min && z_max != z_min;
  }
};

// ---- maps: map/map.h (+ targetReachedCallback, mav_active_3d_planning.patch:20-29), occupancy_map.h, tsdf_map.h [upstream] --------
class Map : public Module {
 public:
  explicit Map(PlannerI& planner) : Module(planner) {}
  virtual bool isTraversable(const Eigen::Vector3d& position, const Eigen::Quaterniond& orientation = Eigen::Quaterniond(1, 0, 0, 0)) = 0;
  virtual bool isObserved(const Eigen::Vector3d& point) = 0;
  virtual void targetReachedCallback() = 0;
};
class OccupancyMap : public Map {
 public:
  explicit OccupancyMap(PlannerI& planner) : Map(planner) {}
  const static unsigned char OCCUPIED = 0, FREE = 1, UNKNOWN = 2;      // map_planexp.cpp:60 compares against UNKNOWN
  virtual unsigned char getVoxelState(const Eigen::Vector3d& point) = 0;
  virtual double getVoxelSize() = 0;
  virtual bool getVoxelCenter(Eigen::Vector3d* center, const Eigen::Vector3d& point) = 0;
};
class TSDFMap : public OccupancyMap {
 public:
  explicit TSDFMap(PlannerI& planner) : OccupancyMap(planner) {}
  virtual double getVoxelDistance(const Eigen::Vector3d& point) = 0;
  virtual double getVoxelWeight(const Eigen::Vector3d& point) = 0;
  virtual double getMaximumWeight() = 0;
};

// ---- trajectory_evaluator.h [upstream]: evaluator + the four helper module kinds -------------------------------------------------
class CostComputer : public Module {
 public:
  explicit CostComputer(PlannerI& planner) : Module(planner) {}
  virtual bool computeCost(TrajectorySegment* traj_in) = 0;
};
class ValueComputer : public Module {
 public:
  explicit ValueComputer(PlannerI& planner) : Module(planner) {}
  virtual bool computeValue(TrajectorySegment* traj_in) = 0;
};
class NextSelector : public Module {
 public:
  explicit NextSelector(PlannerI& planner) : Module(planner) {}
  virtual int selectNextBest(TrajectorySegment* traj_in) = 0;
};
class EvaluatorUpdater : public Module {
 public:
  explicit EvaluatorUpdater(PlannerI& planner) : Module(planner) {}
  virtual bool updateSegment(TrajectorySegment* segment) = 0;
};

class TrajectoryEvaluator : public Module {
 public:
  explicit TrajectoryEvaluator(PlannerI& planner) : Module(planner) {}
  virtual bool computeGain(TrajectorySegment* traj_in) = 0;
  virtual bool computeCost(TrajectorySegment* traj_in);
  virtual bool computeValue(TrajectorySegment* traj_in);
  virtual int selectNextBest(TrajectorySegment* traj_in);
  virtual bool updateSegment(TrajectorySegment* segment);
  virtual void visualizeTrajectoryValue(VisualizationMarkers* markers, const TrajectorySegment& trajectory) { (void)markers; (void)trajectory; }
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  std::unique_ptr<BoundingVolume> bounding_volume_;
  std::unique_ptr<CostComputer> cost_computer_;
  std::unique_ptr<ValueComputer> value_computer_;
  std::unique_ptr<NextSelector> next_selector_;
  std::unique_ptr<EvaluatorUpdater> evaluator_updater_;
  std::string p_cost_args_, p_value_args_, p_next_args_, p_updater_args_;
};

// ---- sensor_model.h [upstream] (sensor_model_planexp.h:18-83 derives from it) ------------------------------------------------------
class SensorModel : public Module {
 public:
  explicit SensorModel(PlannerI& planner) : Module(planner), map_(nullptr) {}
  virtual bool getVisibleVoxelsFromTrajectory(std::vector<Eigen::Vector3d>* result, const TrajectorySegment& traj_in) = 0;
  virtual void visualizeSensorView(VisualizationMarkers* markers, const TrajectorySegment& trajectory) = 0;
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  Map* map_;
  Eigen::Vector3d mounting_translation_;
  Eigen::Quaterniond mounting_rotation_;
};

// ---- simulated_sensor_evaluator.h [upstream] (row A3): sensor model -> bounding volume -> info -> computeGainFromVisibleVoxels ---
class SimulatedSensorEvaluator : public TrajectoryEvaluator {
 public:
  explicit SimulatedSensorEvaluator(PlannerI& planner) : TrajectoryEvaluator(planner), p_clear_from_parents_(false), p_visualize_sensor_view_(false) {}
  bool computeGain(TrajectorySegment* traj_in) override;
  void setupFromParamMap(Module::ParamMap* param_map) override;

 protected:
  std::unique_ptr<SensorModel> sensor_model_;
  bool p_clear_from_parents_, p_visualize_sensor_view_;
  virtual bool storeTrajectoryInformation(TrajectorySegment* traj_in, const std::vector<Eigen::Vector3d>& new_voxels);
  virtual bool computeGainFromVisibleVoxels(TrajectorySegment* traj_in) = 0;
};

// ---- planner_I.h [upstream]: what the modules reach the rest of the planner through ---------------------------------------------
class PlannerI {
 public:
  virtual ~PlannerI() = default;
  virtual const Eigen::Vector3d& getCurrentPosition() const = 0;
  virtual const Eigen::Quaterniond& getCurrentOrientation() const = 0;
  virtual Map& getMap() = 0;
  virtual TrajectoryEvaluator& getTrajectoryEvaluator() = 0;
  virtual ModuleFactory& getFactory() = 0;
  virtual SystemConstraints& getSystemConstraints() = 0;
  virtual void printInfo(const std::string& text) = 0;
  virtual void printWarning(const std::string& text) = 0;
  virtual void printError(const std::string& text) = 0;
};

// ---- inline definitions ----------------------------------------------------------------------------------------------------------
template <class T> inline std::string ModuleFactory::default_type() { return ""; }
template <> inline std::string ModuleFactory::default_type<EvaluatorUpdater>() { return "UpdateNothing"; }
template <> inline std::string ModuleFactory::default_type<BoundingVolume>() { return "BoundingVolume"; }

inline void TrajectoryEvaluator::setupFromParamMap(Module::ParamMap* param_map) {
  const std::string ns = (*param_map)["param_namespace"];
  std::string bv_args;
  setParam<std::string>(param_map, "bounding_volume_args", &bv_args, ns + "/bounding_volume");
  bounding_volume_ = planner_.getFactory().createModule<BoundingVolume>(bv_args, planner_, verbose_modules_);
  setParam<std::string>(param_map, "cost_computer_args", &p_cost_args_, ns + "/cost_computer");
  setParam<std::string>(param_map, "value_computer_args", &p_value_args_, ns + "/value_computer");
  setParam<std::string>(param_map, "next_selector_args", &p_next_args_, ns + "/next_selector");
  setParam<std::string>(param_map, "evaluator_updater_args", &p_updater_args_, ns + "/evaluator_updater");
}
inline bool TrajectoryEvaluator::computeCost(TrajectorySegment* traj_in) {
  if (!cost_computer_) cost_computer_ = planner_.getFactory().createModule<CostComputer>(p_cost_args_, planner_, verbose_modules_);
  return cost_computer_->computeCost(traj_in);
}
inline bool TrajectoryEvaluator::computeValue(TrajectorySegment* traj_in) {
  if (!value_computer_) value_computer_ = planner_.getFactory().createModule<ValueComputer>(p_value_args_, planner_, verbose_modules_);
  return value_computer_->computeValue(traj_in);
}
inline int TrajectoryEvaluator::selectNextBest(TrajectorySegment* traj_in) {
  if (!next_selector_) next_selector_ = planner_.getFactory().createModule<NextSelector>(p_next_args_, planner_, verbose_modules_);
  return next_selector_->selectNextBest(traj_in);
}
inline bool TrajectoryEvaluator::updateSegment(TrajectorySegment* segment) {
  if (!evaluator_updater_) evaluator_updater_ = planner_.getFactory().createModule<EvaluatorUpdater>(p_updater_args_, planner_, verbose_modules_);
  return evaluator_updater_->updateSegment(segment);
}

inline void SensorModel::setupFromParamMap(Module::ParamMap* param_map) {
  map_ = &(planner_.getMap());
  double tx, ty, tz, rx, ry, rz, rw;
  setParam<double>(param_map, "mounting_translation_x", &tx, 0.0);
  setParam<double>(param_map, "mounting_translation_y", &ty, 0.0);
  setParam<double>(param_map, "mounting_translation_z", &tz, 0.0);
  setParam<double>(param_map, "mounting_rotation_x", &rx, 0.0);
  setParam<double>(param_map, "mounting_rotation_y", &ry, 0.0);
  setParam<double>(param_map, "mounting_rotation_z", &rz, 0.0);
  setParam<double>(param_map, "mounting_rotation_w", &rw, 1.0);
  mounting_translation_ = Eigen::Vector3d(tx, ty, tz);
  mounting_rotation_ = Eigen::Quaterniond(rw, rx, ry, rz);
  mounting_rotation_.normalize();
}

inline void SimulatedSensorEvaluator::setupFromParamMap(Module::ParamMap* param_map) {
  TrajectoryEvaluator::setupFromParamMap(param_map);
  setParam<bool>(param_map, "clear_from_parents", &p_clear_from_parents_, false);
  setParam<bool>(param_map, "visualize_sensor_view", &p_visualize_sensor_view_, false);
  std::string args;
  const std::string ns = (*param_map)["param_namespace"];
  setParam<std::string>(param_map, "sensor_model_args", &args, ns + "/sensor_model");
  sensor_model_ = planner_.getFactory().createModule<SensorModel>(args, planner_, verbose_modules_);
}
inline bool SimulatedSensorEvaluator::computeGain(TrajectorySegment* traj_in) {
  std::vector<Eigen::Vector3d> new_voxels;
  sensor_model_->getVisibleVoxelsFromTrajectory(&new_voxels, *traj_in);
  if (bounding_volume_ && bounding_volume_->is_setup) {     // only the voxels inside the target bounding volume count
    new_voxels.erase(std::remove_if(new_voxels.begin(), new_voxels.end(),
                                    [this](const Eigen::Vector3d& voxel) { return !bounding_volume_->contains(voxel); }),
                     new_voxels.end());
  }
  if (p_clear_from_parents_) {                              // drop what a parent segment already sees (off in every shipped cfg)
    TrajectorySegment* previous = traj_in->parent;
    while (previous) {
      if (previous->info) {
        const std::vector<Eigen::Vector3d>& old = reinterpret_cast<SimulatedSensorInfo*>(previous->info.get())->visible_voxels;
        new_voxels.erase(std::remove_if(new_voxels.begin(), new_voxels.end(),
                                        [&old](const Eigen::Vector3d& voxel) { return std::find(old.begin(), old.end(), voxel) != old.end(); }),
                         new_voxels.end());
      }
      previous = previous->parent;
    }
  }
  storeTrajectoryInformation(traj_in, new_voxels);
  computeGainFromVisibleVoxels(traj_in);
  return true;
}
inline bool SimulatedSensorEvaluator::storeTrajectoryInformation(TrajectorySegment* traj_in, const std::vector<Eigen::Vector3d>& new_voxels) {
  SimulatedSensorInfo* new_info = new SimulatedSensorInfo();
  new_info->visible_voxels.assign(new_voxels.begin(), new_voxels.end());
  traj_in->info.reset(new_info);
  return true;
}

}  // namespace active_3d_planning
#endif  // SIPP_SHIM_ACTIVE_3D_PLANNING_CORE_H_
