// stand-in for the slice of roscpp that MapPlanExp::setupFromParamMap touches (map_planexp.cpp:22-33): a NodeHandle whose
// setParam() stores strings in a process-wide table. Test infrastructure only (see ../README.md).
#ifndef SIPP_SHIM_ROS_H_
#define SIPP_SHIM_ROS_H_
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#define ROS_ERROR(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_INFO(...) ((void)0)
namespace ros {
inline std::map<std::string, std::string>& shim_param_table() {
  static std::map<std::string, std::string> t;
  return t;
}
class NodeHandle {
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string& ns) : ns_(ns) {}
  template <class T>
  void setParam(const std::string& name, const T& value) const {
    std::ostringstream os;
    os << std::boolalpha << value;
    shim_param_table()[ns_ + "/" + name] = os.str();
  }
 private:
  std::string ns_;
};
}  // namespace ros
#endif
