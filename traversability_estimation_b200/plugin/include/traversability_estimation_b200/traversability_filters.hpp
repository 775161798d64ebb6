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
 *
 * Fusion under the UNCHANGED YAML.  filters::FilterChain configures every filter before the first
 * update(), so the three plugins register their parameters in a process-wide te_plugin::ChainRegistry at
 * configure().  When exactly one SlopeFilter, one StepFilter and one RoughnessFilter are registered, the first of them whose
 * update() sees a new `elevation` layer (keyed by geometry and a checksum of the layer) launches the whole fused chain once —
 * te_chain(), normals included, regardless of what gridMapFilters/NormalVectorsFilter already wrote — and caches the three
 * output layers in page-locked memory; the other two filters only emit their cached layer.  The normals radius is not visible to
 * these plugins (it belongs to the third-party NormalVectorsFilter, robot_filter_parameter.yaml:8): the registry takes the
 * RoughnessFilter's `estimation_radius` for it (equal in the shipped YAML; TE_B200_NORMALS_RADIUS overrides) and verifies the
 * assumption on the first map by comparing its own surface_normal_z against the layer the upstream filter wrote; on a mismatch
 * — or with TE_B200_FUSE_CHAIN=0 — every filter falls back to its stand-alone kernel, which reads the upstream normals.
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
  te_ctx* get();  // nullptr + ROS_ERROR when no CUDA device can be opened (device index: TE_B200_DEVICE, default 0)

 private:
  te_ctx* ctx_ = nullptr;
  bool failed_ = false;
};

// Process-wide registry behind the cross-plugin fusion described at the top of this file.
class ChainRegistry {
 public:
  enum Kind { kSlope = 0, kStep = 1, kRoughness = 2 };
  static ChainRegistry& instance();
  // configure() / destructor of a plugin instance
  void enroll(Kind kind, const void* owner);
  void retire(Kind kind, const void* owner);
  void setSlope(const void* owner, double critical);
  void setStep(const void* owner, double critical, double r1, double r2, int cells);
  void setRoughness(const void* owner, double critical, double radius);
  // The cached output layer of `kind` for this map (unwrapped to the default start index, rows*cols floats, column-major), running
  // the fused chain first when `elevation` is new.  nullptr: fusion is not available (not all three filters registered, disabled,
  // no device, parameters the fused chain cannot take, or the upstream normals do not match) — the caller runs its own kernel.
  const float* layer(Kind kind, const grid_map::GridMap& unwrapped_map);
  // diagnostics for the tests: fused launches so far / layers served from the cache
  long launches() const { return launches_; }
  long served() const { return served_; }

 private:
  ChainRegistry();
  ~ChainRegistry();
  struct Impl;
  Impl* impl_;
  long launches_ = 0, served_ = 0;
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
