// Stand-in for <pluginlib/class_list_macros.h>: records "derived -> factory" in a process-wide table so
// that the test harness can instantiate plugins by their filter_plugins.xml type string.
#pragma once
#include <functional>
#include <map>
#include <string>
namespace pluginlib_stub {
inline std::map<std::string, std::function<void*()>>& registry() {
  static std::map<std::string, std::function<void*()>> r;
  return r;
}
struct Registrar {
  Registrar(const char* derived, std::function<void*()> f) { registry()[derived] = std::move(f); }
};
}  // namespace pluginlib_stub
#define TE_PL_CAT2(a, b) a##b
#define TE_PL_CAT(a, b) TE_PL_CAT2(a, b)
#define PLUGINLIB_EXPORT_CLASS(Derived, Base) \
  static pluginlib_stub::Registrar TE_PL_CAT(te_pl_registrar_, __COUNTER__)(#Derived, []() -> void* { return static_cast<Base*>(new Derived()); });
