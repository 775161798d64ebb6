// Stand-in for ROS <filters/filter_base.h>: parameter access + the two virtuals.
#pragma once
#include <map>
#include <string>
namespace filters {
struct ParamValue {
  enum Kind { Double, Int, String } kind = Double;
  double d = 0.0;
  int i = 0;
  std::string s;
};
template <typename T>
class FilterBase {
 public:
  virtual ~FilterBase() {}
  // FilterChain::configure: store the `params` block, then call the filter's own configure()
  bool configure(const std::string& name, const std::map<std::string, ParamValue>& params) {
    filter_name_ = name;
    params_ = params;
    configured_ = configure();
    return configured_;
  }
  virtual bool update(const T& data_in, T& data_out) = 0;
  const std::string& getName() const { return filter_name_; }

 protected:
  virtual bool configure() = 0;
  bool getParam(const std::string& name, double& value) const {
    auto it = params_.find(name);
    if (it == params_.end()) return false;
    if (it->second.kind == ParamValue::Double) { value = it->second.d; return true; }
    if (it->second.kind == ParamValue::Int) { value = it->second.i; return true; }
    return false;
  }
  bool getParam(const std::string& name, int& value) const {  // an int given as double in YAML is rejected
    auto it = params_.find(name);
    if (it == params_.end() || it->second.kind != ParamValue::Int) return false;
    value = it->second.i;
    return true;
  }
  bool getParam(const std::string& name, std::string& value) const {
    auto it = params_.find(name);
    if (it == params_.end() || it->second.kind != ParamValue::String) return false;
    value = it->second.s;
    return true;
  }
  std::string filter_name_;
  bool configured_ = false;
  std::map<std::string, ParamValue> params_;
};
}  // namespace filters
