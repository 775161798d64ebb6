/*
 * traversability_filters.hpp — drop-in filters::FilterBase<grid_map::GridMap> plugins backed by libte_b200.
 *
 * Class names, parameter names, defaults, validation and error behaviour mirror the reference's
 * traversability_estimation_filters package so that filter_plugins.xml and the YAML chain
 * (traversability_estimation/config/robot_filter_parameter.yaml:10-28) load them unchanged:
 *   filters::SlopeFilter<T>      replaces  traversability_estimation_filters/src/SlopeFilter.cpp:59-89
 *   filters::StepFilter<T>       replaces  .../src/StepFilter.cpp:102-182
 *   filters::RoughnessFilter<T>  replaces  .../src/RoughnessFilter.cpp:73-132
 * and one extra plugin for users who can edit the YAML:
 *   filters::FusedTraversabilityFilter<T>  = the whole chain of robot_filter_parameter.yaml:2-37 in one launch.
 * The shells hold no arithmetic: they marshal grid_map layers into the C ABI of include/te_b200.h.
 */
#pragma once

#if __has_include(<filters/filter_base.hpp>)
#include <filters/filter_base.hpp>
#else
#include <filters/filter_base.h>
#endif
#include <grid_map_core/GridMap.hpp>

#include <memory>
#include <string>

struct te_ctx;

namespace te_plugin {

// One te_ctx per plugin instance, created lazily on the first update() (so that configure() works
// on machines without a GPU, like the reference's) and destroyed with the filter.
class Device {
 public:
  Device();
  ~Device();
  Device(const Device&) = delete;
  Device& operator=(const Device&) = delete;
  te_ctx* get();  // nullptr + ROS_ERROR when no CUDA device can be opened

 private:
  te_ctx* ctx_ = nullptr;
  bool failed_ = false;
};

}  // namespace te_plugin

namespace filters {

template <typename T>
class SlopeFilter : public FilterBase<T> {
 public:
  SlopeFilter();
  ~SlopeFilter() override;
  bool configure() override;
  bool update(const T& mapIn, T& mapOut) override;

 private:
  double criticalValue_;  // `critical_value` [rad], must lie in [0, pi/2]
  std::string type_;      // `map_type`: name of the output layer
  te_plugin::Device device_;
};

template <typename T>
class StepFilter : public FilterBase<T> {
 public:
  StepFilter();
  ~StepFilter() override;
  bool configure() override;
  bool update(const T& mapIn, T& mapOut) override;

 private:
  double criticalValue_;                           // `critical_value` [m]
  double firstWindowRadius_, secondWindowRadius_;  // `first_window_radius`, `second_window_radius` [m]
  int nCellCritical_;                              // `critical_cell_number`
  std::string type_;                               // `map_type`
  te_plugin::Device device_;
};

template <typename T>
class RoughnessFilter : public FilterBase<T> {
 public:
  RoughnessFilter();
  ~RoughnessFilter() override;
  bool configure() override;
  bool update(const T& mapIn, T& mapOut) override;

 private:
  double criticalValue_;     // `critical_value` [m]
  double estimationRadius_;  // `estimation_radius` [m]
  std::string type_;         // `map_type`
  te_plugin::Device device_;
};

// elevation -> traversability_slope / _step / _roughness / traversability in one fused launch.
// Parameters (all optional, defaulting to the shipped YAML values): normals_radius, slope_critical_value,
// step_critical_value, step_first_window_radius, step_second_window_radius, step_critical_cell_number,
// roughness_critical_value, roughness_estimation_radius, keep_surface_normals (int, 0/1).
template <typename T>
class FusedTraversabilityFilter : public FilterBase<T> {
 public:
  FusedTraversabilityFilter();
  ~FusedTraversabilityFilter() override;
  bool configure() override;
  bool update(const T& mapIn, T& mapOut) override;

 private:
  double normalsRadius_, slopeCritical_, stepCritical_, stepR1_, stepR2_, roughCritical_, roughRadius_;
  int stepCells_;
  bool keepNormals_;
  te_plugin::Device device_;
};

}  // namespace filters
