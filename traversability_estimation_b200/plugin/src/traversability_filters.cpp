// traversability_filters.cpp — see include/traversability_estimation_b200/traversability_filters.hpp
#include "traversability_estimation_b200/traversability_filters.hpp"

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <set>

#if __has_include(<ros/console.h>)
#include <ros/console.h>
#endif
#include <pluginlib/class_list_macros.h>

#include "te_b200.h"

namespace te_plugin {

Device::Device() {}
Device::~Device() {
  if (ctx_) te_destroy(ctx_);
}
te_ctx* Device::get() {
  if (!ctx_ && !failed_) {
    const char* dev = std::getenv("TE_B200_DEVICE");
    if (te_create(&ctx_, dev ? std::atoi(dev) : 0) != TE_OK) {
      ROS_ERROR("libte_b200: %s", te_last_error());
      failed_ = true;
      ctx_ = nullptr;
    }
  }
  return ctx_;
}

namespace {

te_geometry geometry_of(const grid_map::GridMap& m) {
  te_geometry g;
  g.rows = m.getSize()(0);
  g.cols = m.getSize()(1);
  g.resolution = m.getResolution();
  g.length_x = m.getLength()(0);
  g.length_y = m.getLength()(1);
  g.position_x = m.getPosition()(0);
  g.position_y = m.getPosition()(1);
  g.start_row = m.getStartIndex()(0);
  g.start_col = m.getStartIndex()(1);
  return g;
}

// The kernels want the default start index.  For a circular-buffer map compute on an unwrapped copy
// and scatter the new layers back into the buffer order of `out`.
struct Unwrapped {
  const grid_map::GridMap* in;
  grid_map::GridMap copy;
  bool wrapped;
  explicit Unwrapped(const grid_map::GridMap& m) : in(&m), wrapped(m.getStartIndex()(0) != 0 || m.getStartIndex()(1) != 0) {
    if (wrapped) {
      copy = m;
      copy.convertToDefaultStartIndex();
    }
  }
  const grid_map::GridMap& map() const { return wrapped ? copy : *in; }
};

void scatter_back(const grid_map::GridMap& like, const float* unwrapped, float* dst) {
  const int rows = like.getSize()(0), cols = like.getSize()(1);
  const int s0 = like.getStartIndex()(0), s1 = like.getStartIndex()(1);
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i < rows; ++i) dst[(size_t)((j + s1) % cols) * rows + (i + s0) % rows] = unwrapped[(size_t)j * rows + i];
}

te_chain_params yaml_defaults() {
  te_chain_params p{};
  p.normals_radius = 0.05;
  p.normals_algorithm = TE_NORMALS_FIXTURE;
  p.normals_positive_axis = 2;
  p.slope_critical = 1.0;
  p.step_critical = 0.12;
  p.step_first_radius = 0.04;
  p.step_second_radius = 0.04;
  p.step_critical_cells = 4;
  p.roughness_critical = 0.05;
  p.roughness_radius = 0.05;
  p.fuse_weight = 1.0f / 3.0f;
  return p;
}

}  // namespace

// ------------------------------------------------------------------ ChainRegistry
struct ChainRegistry::Impl {
  std::mutex mu;
  std::set<const void*> owners[3];   // configured instances per filter type: fusion needs exactly one of each
  bool have[3] = {false, false, false};
  te_chain_params params = yaml_defaults();
  bool disabled = false;        // TE_B200_FUSE_CHAIN=0, or the upstream normals did not match ours
  bool normals_checked = false;
  Device device;
  // cache: key + three layers in page-locked memory
  te_geometry key_geo{};
  uint64_t key_sum = 0;
  bool valid = false;
  size_t cap = 0;
  float* out[4] = {nullptr, nullptr, nullptr, nullptr};   // slope, step, roughness, traversability (scratch)
  float* nz = nullptr;                                    // first map only: our surface_normal_z (+ x, y scratch)
  float* nxy[2] = {nullptr, nullptr};
  ~Impl() {
    for (float*& q : out) { te_host_free(q); q = nullptr; }
    te_host_free(nz);
    te_host_free(nxy[0]);
    te_host_free(nxy[1]);
  }
};

ChainRegistry::ChainRegistry() : impl_(new Impl) {
  const char* e = std::getenv("TE_B200_FUSE_CHAIN");
  if (e && std::atoi(e) == 0) impl_->disabled = true;
}
ChainRegistry::~ChainRegistry() { delete impl_; }
ChainRegistry& ChainRegistry::instance() {
  static ChainRegistry r;
  return r;
}
void ChainRegistry::enroll(Kind k, const void* owner) {
  std::lock_guard<std::mutex> lock(impl_->mu);
  impl_->owners[k].insert(owner);   // a second instance of a type makes the chain ambiguous: no fusion while it lives
  impl_->valid = false;
}
void ChainRegistry::retire(Kind k, const void* owner) {
  std::lock_guard<std::mutex> lock(impl_->mu);
  if (impl_->owners[k].erase(owner) && impl_->owners[k].empty()) impl_->have[k] = false;
  impl_->valid = false;
}
void ChainRegistry::setSlope(const void* owner, double critical) {
  enroll(kSlope, owner);
  std::lock_guard<std::mutex> lock(impl_->mu);
  impl_->params.slope_critical = critical;
  impl_->have[kSlope] = true;
}
void ChainRegistry::setStep(const void* owner, double critical, double r1, double r2, int cells) {
  enroll(kStep, owner);
  std::lock_guard<std::mutex> lock(impl_->mu);
  impl_->params.step_critical = critical;
  impl_->params.step_first_radius = r1;
  impl_->params.step_second_radius = r2;
  impl_->params.step_critical_cells = cells;
  impl_->have[kStep] = true;
}
void ChainRegistry::setRoughness(const void* owner, double critical, double radius) {
  enroll(kRoughness, owner);
  std::lock_guard<std::mutex> lock(impl_->mu);
  impl_->params.roughness_critical = critical;
  impl_->params.roughness_radius = radius;
  const char* e = std::getenv("TE_B200_NORMALS_RADIUS");
  impl_->params.normals_radius = e ? std::atof(e) : radius;   // see the header: not visible to these plugins
  impl_->have[kRoughness] = true;
}

namespace {
// Order-sensitive 64-bit checksum of a float layer (four independent lanes so that it runs at memory speed).
uint64_t layer_checksum(const float* p, size_t n) {
  uint64_t a = 0x9e3779b97f4a7c15ull, b = 0xc2b2ae3d27d4eb4full, c = 0x165667b19e3779f9ull, d = 0x27d4eb2f165667c5ull;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
  size_t i = 0;
  for (; i + 4 <= n; i += 4) {
    a = (a ^ w[i]) * 0x100000001b3ull;
    b = (b ^ w[i + 1]) * 0x100000001b3ull;
    c = (c ^ w[i + 2]) * 0x100000001b3ull;
    d = (d ^ w[i + 3]) * 0x100000001b3ull;
  }
  for (; i < n; ++i) a = (a ^ w[i]) * 0x100000001b3ull;
  return a ^ (b << 1 | b >> 63) ^ (c << 2 | c >> 62) ^ (d << 3 | d >> 61) ^ (uint64_t)n;
}
}  // namespace

const float* ChainRegistry::layer(Kind kind, const grid_map::GridMap& m) {
  Impl& I = *impl_;
  std::lock_guard<std::mutex> lock(I.mu);
  if (I.disabled) return nullptr;
  for (int k = 0; k < 3; ++k)
    if (I.owners[k].size() != 1 || !I.have[k]) return nullptr;
  if (!m.exists("elevation")) return nullptr;   // the stand-alone path reports the missing layer the reference's way
  const te_geometry g = geometry_of(m);
  const size_t n = (size_t)g.rows * g.cols;
  const float* elevation = m.get("elevation").data();
  const uint64_t sum = layer_checksum(elevation, n);
  if (I.valid && std::memcmp(&I.key_geo, &g, sizeof(g)) == 0 && I.key_sum == sum) {
    ++served_;
    return I.out[kind];
  }
  te_ctx* ctx = I.device.get();
  if (!ctx) return nullptr;   // the caller's own kernel reports the failure
  if (I.cap < n) {
    for (float*& q : I.out) { te_host_free(q); q = nullptr; }
    I.cap = 0;
    for (float*& q : I.out)
      if (te_host_alloc((void**)&q, n * sizeof(float)) != TE_OK) return nullptr;
    I.cap = n;
  }
  I.valid = false;
  const bool check = !I.normals_checked && m.exists("surface_normal_z");
  if (check) {
    te_host_free(I.nz); te_host_free(I.nxy[0]); te_host_free(I.nxy[1]);
    I.nz = I.nxy[0] = I.nxy[1] = nullptr;
    if (te_host_alloc((void**)&I.nz, n * sizeof(float)) != TE_OK || te_host_alloc((void**)&I.nxy[0], n * sizeof(float)) != TE_OK ||
        te_host_alloc((void**)&I.nxy[1], n * sizeof(float)) != TE_OK)
      return nullptr;
  }
  if (te_chain(ctx, &g, nullptr, &I.params, elevation, I.out[0], I.out[1], I.out[2], I.out[3], check ? I.nxy[0] : nullptr,
               check ? I.nxy[1] : nullptr, check ? I.nz : nullptr, TE_MEM_HOST) != TE_OK) {
    ROS_ERROR("fused chain: %s", te_last_error());
    I.disabled = true;   // parameters the fused entry point rejects: the stand-alone kernels take over
    return nullptr;
  }
  ++launches_;
  if (check) {
    // the upstream NormalVectorsFilter must have computed what the fused chain computes itself (same radius, same algorithm)
    const float* theirs = m.get("surface_normal_z").data();
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
      const float a = I.nz[i], b = theirs[i];
      if (std::isnan(a) != std::isnan(b) || (!std::isnan(a) && std::fabs(a - b) > 1e-5f)) ++bad;
    }
    te_host_free(I.nz); te_host_free(I.nxy[0]); te_host_free(I.nxy[1]);
    I.nz = I.nxy[0] = I.nxy[1] = nullptr;
    I.normals_checked = true;
    if (bad * 1000 > n) {   // more than 0.1 % of the cells: not rounding noise on degenerate windows
      ROS_ERROR("libte_b200: surface normals of the upstream filter differ from the fused chain's in %zu of %zu cells "
                "(radius %.4f assumed): cross-plugin fusion disabled, using the stand-alone kernels", bad, n, I.params.normals_radius);
      I.disabled = true;
      return nullptr;
    }
  }
  I.key_geo = g;
  I.key_sum = sum;
  I.valid = true;
  ++served_;
  return I.out[kind];
}

}  // namespace te_plugin

namespace filters {

using te_plugin::geometry_of;
using te_plugin::ChainRegistry;

namespace {
// Emit a cached layer into `dst`, the layer's storage in the output map.  The registry runs te_chain on the map AS STORED
// (circular-buffer start index included: the C ABI unwraps and re-wraps in its copies), so the cached layer already has the
// buffer order of the map it was computed from — no unwrapped host copy of a moving map, no scatter.
void emit(const grid_map::GridMap& like, const float* cached, float* dst) {
  std::memcpy(dst, cached, sizeof(float) * (size_t)like.getSize()(0) * like.getSize()(1));
}
}  // namespace

// ------------------------------------------------------------------ SlopeFilter
template <typename T>
SlopeFilter<T>::SlopeFilter() : criticalValue_(M_PI_4), type_("traversability_slope") {}
template <typename T>
SlopeFilter<T>::~SlopeFilter() { ChainRegistry::instance().retire(ChainRegistry::kSlope, this); }

template <typename T>
bool SlopeFilter<T>::configure() {
  if (!FilterBase<T>::getParam(std::string("critical_value"), criticalValue_)) {
    ROS_ERROR("SlopeFilter did not find param critical_value");
    return false;
  }
  if (criticalValue_ > M_PI_2 || criticalValue_ < 0.0) {
    ROS_ERROR("Critical slope must be in the interval [0, PI/2]");
    return false;
  }
  if (!FilterBase<T>::getParam(std::string("map_type"), type_)) {
    ROS_ERROR("SlopeFilter did not find param map_type");
    return false;
  }
  ChainRegistry::instance().setSlope(this, criticalValue_);
  return true;
}

template <typename T>
bool SlopeFilter<T>::update(const T& mapIn, T& mapOut) {
  mapOut = mapIn;
  mapOut.add(type_);  // NaN everywhere; cells without a surface normal keep it
  (void)mapOut.get("surface_normal_z");  // throws std::out_of_range like GridMap::at
  if (const float* cached = ChainRegistry::instance().layer(ChainRegistry::kSlope, mapOut)) {
    emit(mapOut, cached, mapOut.get(type_).data());
    return true;
  }
  const te_plugin::Unwrapped u(mapOut);
  const float* nz = u.map().get("surface_normal_z").data();
  te_ctx* ctx = device_.get();
  if (!ctx) return false;
  const te_geometry g = geometry_of(u.map());
  grid_map::Matrix tmp;
  float* dst = mapOut.get(type_).data();
  if (u.wrapped) { tmp = grid_map::Matrix(g.rows, g.cols, 0.f); dst = tmp.data(); }
  if (te_slope(ctx, &g, criticalValue_, nz, dst, TE_MEM_HOST) != TE_OK) {
    ROS_ERROR("SlopeFilter: %s", te_last_error());
    return false;
  }
  if (u.wrapped) te_plugin::scatter_back(mapOut, tmp.data(), mapOut.get(type_).data());
  return true;
}

// ------------------------------------------------------------------ StepFilter
template <typename T>
StepFilter<T>::StepFilter()
    : criticalValue_(0.3), firstWindowRadius_(0.08), secondWindowRadius_(0.08), nCellCritical_(5), type_("traversability_step") {}
template <typename T>
StepFilter<T>::~StepFilter() { ChainRegistry::instance().retire(ChainRegistry::kStep, this); }

template <typename T>
bool StepFilter<T>::configure() {
  if (!FilterBase<T>::getParam(std::string("critical_value"), criticalValue_)) {
    ROS_ERROR("Step filter did not find param critical_value");
    return false;
  }
  if (criticalValue_ < 0.0) {
    ROS_ERROR("Critical step height must be greater than zero");
    return false;
  }
  if (!FilterBase<T>::getParam(std::string("first_window_radius"), firstWindowRadius_)) {
    ROS_ERROR("Step filter did not find param 'first_window_radius'");
    return false;
  }
  if (firstWindowRadius_ < 0.0) {
    ROS_ERROR("'first_window_radius' must be greater than zero");
    return false;
  }
  if (!FilterBase<T>::getParam(std::string("second_window_radius"), secondWindowRadius_)) {
    ROS_ERROR("Step filter did not find param 'second_window_radius'");
    return false;
  }
  if (secondWindowRadius_ < 0.0) {
    ROS_ERROR("'second_window_radius' must be greater than zero");
    return false;
  }
  if (!FilterBase<T>::getParam(std::string("critical_cell_number"), nCellCritical_)) {
    ROS_ERROR("Step filter did not find param 'critical_cell_number'");
    return false;
  }
  if (nCellCritical_ <= 0) {
    ROS_ERROR("'critical_cell_number' must be greater than zero");
    return false;
  }
  if (!FilterBase<T>::getParam(std::string("map_type"), type_)) {
    ROS_ERROR("Step filter did not find param map_type");
    return false;
  }
  ChainRegistry::instance().setStep(this, criticalValue_, firstWindowRadius_, secondWindowRadius_, nCellCritical_);
  return true;
}

template <typename T>
bool StepFilter<T>::update(const T& mapIn, T& mapOut) {
  mapOut = mapIn;
  mapOut.add(type_);
  (void)mapOut.get("elevation");
  if (const float* cached = ChainRegistry::instance().layer(ChainRegistry::kStep, mapOut)) {
    emit(mapOut, cached, mapOut.get(type_).data());
    return true;
  }
  const te_plugin::Unwrapped u(mapOut);
  const float* elevation = u.map().get("elevation").data();
  te_ctx* ctx = device_.get();
  if (!ctx) return false;
  const te_geometry g = geometry_of(u.map());
  te_chain_params p = te_plugin::yaml_defaults();
  p.step_critical = criticalValue_;
  p.step_first_radius = firstWindowRadius_;
  p.step_second_radius = secondWindowRadius_;
  p.step_critical_cells = nCellCritical_;
  grid_map::Matrix tmp;
  float* dst = mapOut.get(type_).data();
  if (u.wrapped) { tmp = grid_map::Matrix(g.rows, g.cols, 0.f); dst = tmp.data(); }
  if (te_step(ctx, &g, &p, elevation, dst, TE_MEM_HOST) != TE_OK) {  // the step_height layer never leaves the device
    ROS_ERROR("StepFilter: %s", te_last_error());
    return false;
  }
  if (u.wrapped) te_plugin::scatter_back(mapOut, tmp.data(), mapOut.get(type_).data());
  return true;
}

// ------------------------------------------------------------------ RoughnessFilter
template <typename T>
RoughnessFilter<T>::RoughnessFilter() : criticalValue_(0.3), estimationRadius_(0.3), type_("traversability_roughness") {}
template <typename T>
RoughnessFilter<T>::~RoughnessFilter() { ChainRegistry::instance().retire(ChainRegistry::kRoughness, this); }

template <typename T>
bool RoughnessFilter<T>::configure() {
  if (!FilterBase<T>::getParam(std::string("critical_value"), criticalValue_)) {
    ROS_ERROR("RoughnessFilter did not find param critical_value");
    return false;
  }
  if (criticalValue_ < 0.0) {
    ROS_ERROR("Critical roughness must be greater than zero");
    return false;
  }
  if (!FilterBase<T>::getParam(std::string("estimation_radius"), estimationRadius_)) {
    ROS_ERROR("RoughnessFilter did not find param estimation_radius");
    return false;
  }
  if (estimationRadius_ < 0.0) {
    ROS_ERROR("Roughness estimation radius must be greater than zero");
    return false;
  }
  if (!FilterBase<T>::getParam(std::string("map_type"), type_)) {
    ROS_ERROR("RoughnessFilter did not find param map_type");
    return false;
  }
  ChainRegistry::instance().setRoughness(this, criticalValue_, estimationRadius_);
  return true;
}

template <typename T>
bool RoughnessFilter<T>::update(const T& mapIn, T& mapOut) {
  mapOut = mapIn;
  mapOut.add(type_);
  (void)mapOut.get("surface_normal_x");  // same order of look-ups (and of the exception for a missing layer) as before
  (void)mapOut.get("elevation");
  (void)mapOut.get("surface_normal_y");
  (void)mapOut.get("surface_normal_z");
  if (const float* cached = ChainRegistry::instance().layer(ChainRegistry::kRoughness, mapOut)) {
    emit(mapOut, cached, mapOut.get(type_).data());
    return true;
  }
  const te_plugin::Unwrapped u(mapOut);
  const float* nx = u.map().get("surface_normal_x").data();
  const float* elevation = u.map().get("elevation").data();
  const float* ny = u.map().get("surface_normal_y").data();
  const float* nz = u.map().get("surface_normal_z").data();
  te_ctx* ctx = device_.get();
  if (!ctx) return false;
  const te_geometry g = geometry_of(u.map());
  te_chain_params p = te_plugin::yaml_defaults();
  p.roughness_critical = criticalValue_;
  p.roughness_radius = estimationRadius_;
  grid_map::Matrix tmp;
  float* dst = mapOut.get(type_).data();
  if (u.wrapped) { tmp = grid_map::Matrix(g.rows, g.cols, 0.f); dst = tmp.data(); }
  if (te_roughness(ctx, &g, &p, elevation, nx, ny, nz, dst, TE_MEM_HOST) != TE_OK) {
    ROS_ERROR("RoughnessFilter: %s", te_last_error());
    return false;
  }
  if (u.wrapped) te_plugin::scatter_back(mapOut, tmp.data(), mapOut.get(type_).data());
  return true;
}

// ------------------------------------------------------------------ FusedTraversabilityFilter
template <typename T>
FusedTraversabilityFilter<T>::FusedTraversabilityFilter()
    : normalsRadius_(0.05), slopeCritical_(1.0), stepCritical_(0.12), stepR1_(0.04), stepR2_(0.04), roughCritical_(0.05),
      roughRadius_(0.05), stepCells_(4), keepNormals_(false) {}
template <typename T>
FusedTraversabilityFilter<T>::~FusedTraversabilityFilter() {}

template <typename T>
bool FusedTraversabilityFilter<T>::configure() {
  FilterBase<T>::getParam(std::string("normals_radius"), normalsRadius_);
  FilterBase<T>::getParam(std::string("slope_critical_value"), slopeCritical_);
  FilterBase<T>::getParam(std::string("step_critical_value"), stepCritical_);
  FilterBase<T>::getParam(std::string("step_first_window_radius"), stepR1_);
  FilterBase<T>::getParam(std::string("step_second_window_radius"), stepR2_);
  FilterBase<T>::getParam(std::string("step_critical_cell_number"), stepCells_);
  FilterBase<T>::getParam(std::string("roughness_critical_value"), roughCritical_);
  FilterBase<T>::getParam(std::string("roughness_estimation_radius"), roughRadius_);
  int keep = 0;
  if (FilterBase<T>::getParam(std::string("keep_surface_normals"), keep)) keepNormals_ = keep != 0;
  if (slopeCritical_ > M_PI_2 || slopeCritical_ < 0.0) {
    ROS_ERROR("Critical slope must be in the interval [0, PI/2]");
    return false;
  }
  if (stepCritical_ < 0.0 || stepR1_ < 0.0 || stepR2_ < 0.0 || stepCells_ <= 0 || roughCritical_ < 0.0 || roughRadius_ < 0.0 ||
      normalsRadius_ < 0.0) {
    ROS_ERROR("FusedTraversabilityFilter: critical values and radii must not be negative, critical_cell_number must be positive");
    return false;
  }
  return true;
}

template <typename T>
bool FusedTraversabilityFilter<T>::update(const T& mapIn, T& mapOut) {
  mapOut = mapIn;
  static const char* kOut[4] = {"traversability_slope", "traversability_step", "traversability_roughness", "traversability"};
  static const char* kNrm[3] = {"surface_normal_x", "surface_normal_y", "surface_normal_z"};
  for (const char* l : kOut) mapOut.add(l);
  if (keepNormals_)
    for (const char* l : kNrm) mapOut.add(l);
  // The map is passed AS STORED: te_chain(TE_MEM_HOST) honours the circular-buffer start index in its copies to and from the
  // device, so a moving (robot-centric) map needs neither an unwrapped host copy nor a scatter of the results.
  const float* elevation = mapOut.get("elevation").data();  // throws std::out_of_range like GridMap::at
  te_ctx* ctx = device_.get();
  if (!ctx) return false;
  const te_geometry g = geometry_of(mapOut);
  te_chain_params p = te_plugin::yaml_defaults();
  p.normals_radius = normalsRadius_;
  p.slope_critical = slopeCritical_;
  p.step_critical = stepCritical_;
  p.step_first_radius = stepR1_;
  p.step_second_radius = stepR2_;
  p.step_critical_cells = stepCells_;
  p.roughness_critical = roughCritical_;
  p.roughness_radius = roughRadius_;
  float* dst[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int k = 0; k < 4; ++k) dst[k] = mapOut.get(kOut[k]).data();
  if (keepNormals_)
    for (int k = 0; k < 3; ++k) dst[4 + k] = mapOut.get(kNrm[k]).data();
  if (te_chain(ctx, &g, nullptr, &p, elevation, dst[0], dst[1], dst[2], dst[3], dst[4], dst[5], dst[6], TE_MEM_HOST) != TE_OK) {
    ROS_ERROR("FusedTraversabilityFilter: %s", te_last_error());
    return false;
  }
  return true;
}

}  // namespace filters

PLUGINLIB_EXPORT_CLASS(filters::SlopeFilter<grid_map::GridMap>, filters::FilterBase<grid_map::GridMap>)
PLUGINLIB_EXPORT_CLASS(filters::StepFilter<grid_map::GridMap>, filters::FilterBase<grid_map::GridMap>)
PLUGINLIB_EXPORT_CLASS(filters::RoughnessFilter<grid_map::GridMap>, filters::FilterBase<grid_map::GridMap>)
PLUGINLIB_EXPORT_CLASS(filters::FusedTraversabilityFilter<grid_map::GridMap>, filters::FilterBase<grid_map::GridMap>)
