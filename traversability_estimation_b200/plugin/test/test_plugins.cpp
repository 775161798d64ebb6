// test_plugins.cpp — harness for the plugin shells, built against the stub ROS/grid_map headers.
//   test_plugins configure                       parameter handling of configure(), no GPU needed
//   test_plugins nogpu                           update() must fail loudly (return false) without a CUDA device
//   test_plugins chain R C res px py in.bin out  run the YAML chain through the plugins (needs a GPU)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <filters/filter_base.h>
#include <grid_map_core/GridMap.hpp>
#include <pluginlib/class_list_macros.h>

#include "te_b200.h"
#include "traversability_estimation_b200/traversability_filters.hpp"

#include <chrono>

using Base = filters::FilterBase<grid_map::GridMap>;
using Params = std::map<std::string, filters::ParamValue>;

static std::unique_ptr<Base> create(const std::string& type) {
  auto& r = pluginlib_stub::registry();
  auto it = r.find(type);
  if (it == r.end()) { std::fprintf(stderr, "no plugin of type %s\n", type.c_str()); std::exit(2); }
  return std::unique_ptr<Base>(static_cast<Base*>(it->second()));
}
static filters::ParamValue D(double v) { filters::ParamValue p; p.kind = filters::ParamValue::Double; p.d = v; return p; }
static filters::ParamValue I(int v) { filters::ParamValue p; p.kind = filters::ParamValue::Int; p.i = v; return p; }
static filters::ParamValue S(const char* v) { filters::ParamValue p; p.kind = filters::ParamValue::String; p.s = v; return p; }

static int g_fail = 0;
static void expect(bool ok, const char* what) {
  std::printf("%s %s\n", ok ? "PASS" : "FAIL", what);
  if (!ok) ++g_fail;
}

static const char* kSlope = "filters::SlopeFilter<grid_map::GridMap>";
static const char* kStep = "filters::StepFilter<grid_map::GridMap>";
static const char* kRough = "filters::RoughnessFilter<grid_map::GridMap>";
static const char* kFused = "filters::FusedTraversabilityFilter<grid_map::GridMap>";

static Params slope_yaml() { return {{"map_type", S("traversability_slope")}, {"critical_value", D(1.0)}}; }
static Params step_yaml() {
  return {{"map_type", S("traversability_step")}, {"critical_value", D(0.12)}, {"first_window_radius", D(0.04)},
          {"second_window_radius", D(0.04)}, {"critical_cell_number", I(4)}};
}
static Params rough_yaml() { return {{"map_type", S("traversability_roughness")}, {"critical_value", D(0.05)}, {"estimation_radius", D(0.05)}}; }

static int test_configure() {
  expect(create(kSlope)->configure("slopeFilter", slope_yaml()), "SlopeFilter accepts the YAML block");
  { Params p = slope_yaml(); p.erase("critical_value"); expect(!create(kSlope)->configure("s", p), "SlopeFilter rejects missing critical_value"); }
  { Params p = slope_yaml(); p["critical_value"] = D(1.6); expect(!create(kSlope)->configure("s", p), "SlopeFilter rejects critical_value > pi/2"); }
  { Params p = slope_yaml(); p["critical_value"] = D(-0.1); expect(!create(kSlope)->configure("s", p), "SlopeFilter rejects negative critical_value"); }
  { Params p = slope_yaml(); p.erase("map_type"); expect(!create(kSlope)->configure("s", p), "SlopeFilter rejects missing map_type"); }
  expect(create(kStep)->configure("stepFilter", step_yaml()), "StepFilter accepts the YAML block");
  { Params p = step_yaml(); p["critical_cell_number"] = D(4.0); expect(!create(kStep)->configure("s", p), "StepFilter rejects critical_cell_number given as double"); }
  { Params p = step_yaml(); p["critical_cell_number"] = I(0); expect(!create(kStep)->configure("s", p), "StepFilter rejects critical_cell_number <= 0"); }
  { Params p = step_yaml(); p["first_window_radius"] = D(-1.0); expect(!create(kStep)->configure("s", p), "StepFilter rejects negative first_window_radius"); }
  { Params p = step_yaml(); p.erase("second_window_radius"); expect(!create(kStep)->configure("s", p), "StepFilter rejects missing second_window_radius"); }
  expect(create(kRough)->configure("roughnessFilter", rough_yaml()), "RoughnessFilter accepts the YAML block");
  { Params p = rough_yaml(); p["critical_value"] = D(-0.01); expect(!create(kRough)->configure("s", p), "RoughnessFilter rejects negative critical_value"); }
  { Params p = rough_yaml(); p.erase("estimation_radius"); expect(!create(kRough)->configure("s", p), "RoughnessFilter rejects missing estimation_radius"); }
  expect(create(kFused)->configure("fused", Params{}), "FusedTraversabilityFilter configures with defaults");
  { Params p; p["slope_critical_value"] = D(2.0); expect(!create(kFused)->configure("f", p), "FusedTraversabilityFilter rejects slope critical > pi/2"); }
  // missing input layer: GridMap::get throws std::out_of_range exactly like the reference's GridMap::at
  {
    grid_map::GridMap m;
    m.setGeometry(grid_map::Length{{0.2, 0.2}}, 0.02, grid_map::Position{{0, 0}});
    m.add("elevation", 0.f);
    auto f = create(kSlope);
    f->configure("s", slope_yaml());
    grid_map::GridMap out;
    bool threw = false;
    try { f->update(m, out); } catch (const std::out_of_range&) { threw = true; }
    expect(threw, "SlopeFilter::update throws std::out_of_range when surface_normal_z is missing");
  }
  return g_fail;
}

static grid_map::GridMap make_map(int rows, int cols, double res, double px, double py, const std::vector<float>& elev) {
  grid_map::GridMap m;
  m.setGeometry(grid_map::Length{{rows * res, cols * res}}, res, grid_map::Position{{px, py}});
  m.add("elevation");
  std::memcpy(m.get("elevation").data(), elev.data(), elev.size() * sizeof(float));
  return m;
}

static int test_nogpu() {
  std::vector<float> e(100, 0.f);
  grid_map::GridMap m = make_map(10, 10, 0.02, 0, 0, e);
  m.add("surface_normal_z", 1.f);
  auto f = create(kSlope);
  f->configure("s", slope_yaml());
  grid_map::GridMap out;
  const bool ok = f->update(m, out);
  te_ctx* c = nullptr;
  const int rc = te_create(&c, 0);
  if (rc == TE_OK) {
    te_destroy(c);
    expect(ok, "a CUDA device is present: update() succeeds");
  } else {
    expect(!ok, "no CUDA device: update() returns false (no CPU fallback)");
    expect(std::strlen(te_last_error()) > 0, "te_last_error() explains why");
  }
  return g_fail;
}

static void dump(const std::string& path, const grid_map::Matrix& m) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) { std::perror(path.c_str()); std::exit(3); }
  std::fwrite(m.data(), sizeof(float), (size_t)m.rows() * m.cols(), f);
  std::fclose(f);
}

static int test_chain(int argc, char** argv) {
  if (argc < 9) { std::fprintf(stderr, "usage: chain R C res px py in.bin outprefix [start_row start_col]\n"); return 2; }
  const int rows = std::atoi(argv[2]), cols = std::atoi(argv[3]);
  const double res = std::atof(argv[4]), px = std::atof(argv[5]), py = std::atof(argv[6]);
  const std::string out = argv[8];
  std::vector<float> elev((size_t)rows * cols);
  FILE* f = std::fopen(argv[7], "rb");
  if (!f || std::fread(elev.data(), sizeof(float), elev.size(), f) != elev.size()) { std::fprintf(stderr, "cannot read %s\n", argv[7]); return 3; }
  std::fclose(f);
  grid_map::GridMap m0 = make_map(rows, cols, res, px, py, elev);
  if (argc >= 11) {
    // present the same map as a circular buffer with a non-zero start index
    const int s0 = std::atoi(argv[9]), s1 = std::atoi(argv[10]);
    grid_map::Matrix w(rows, cols, 0.f);
    for (int j = 0; j < cols; ++j)
      for (int i = 0; i < rows; ++i) w((i + s0) % rows, (j + s1) % cols) = m0.get("elevation")(i, j);
    m0.get("elevation") = w;
    m0.setStartIndex(grid_map::Index{{s0, s1}});
  }
  // gridMapFilters/NormalVectorsFilter is third-party and not replaced by this package: emulate its place
  // in the chain with the C ABI's stand-alone normals entry point (default start index only).
  grid_map::GridMap m1 = m0;
  const bool wrapped = m0.getStartIndex()(0) != 0 || m0.getStartIndex()(1) != 0;
  if (!wrapped) {
    for (const char* l : {"surface_normal_x", "surface_normal_y", "surface_normal_z"}) m1.add(l);
    te_ctx* ctx = nullptr;
    if (te_create(&ctx, 0) != TE_OK) { std::fprintf(stderr, "%s\n", te_last_error()); return 4; }
    te_geometry g{rows, cols, res, rows * res, cols * res, px, py, 0, 0};
    te_chain_params p{};
    p.normals_radius = 0.05; p.normals_algorithm = TE_NORMALS_FIXTURE; p.normals_positive_axis = 2; p.slope_critical = 1.0;
    p.step_critical = 0.12; p.step_first_radius = 0.04; p.step_second_radius = 0.04; p.step_critical_cells = 4;
    p.roughness_critical = 0.05; p.roughness_radius = 0.05; p.fuse_weight = 1.0f / 3.0f;
    if (te_normals(ctx, &g, &p, m1.get("elevation").data(), m1.get("surface_normal_x").data(), m1.get("surface_normal_y").data(),
                   m1.get("surface_normal_z").data(), TE_MEM_HOST) != TE_OK) { std::fprintf(stderr, "%s\n", te_last_error()); return 4; }
    te_destroy(ctx);
    grid_map::GridMap a, b, c;
    auto fs = create(kSlope); fs->configure("slopeFilter", slope_yaml());
    auto ft = create(kStep); ft->configure("stepFilter", step_yaml());
    auto fr = create(kRough); fr->configure("roughnessFilter", rough_yaml());
    if (!fs->update(m1, a) || !ft->update(a, b) || !fr->update(b, c)) { std::fprintf(stderr, "chain failed\n"); return 5; }
    // cross-plugin fusion under the unchanged YAML: one fused launch, three layers served (0/0 with TE_B200_FUSE_CHAIN=0)
    std::printf("REGISTRY launches=%ld served=%ld\n", te_plugin::ChainRegistry::instance().launches(),
                te_plugin::ChainRegistry::instance().served());
    expect(!c.exists("step_height"), "StepFilter leaves no step_height layer behind");
    expect(c.exists("elevation") && c.exists("surface_normal_z"), "filters keep the other layers");
    dump(out + "_slope.bin", c.get("traversability_slope"));
    dump(out + "_step.bin", c.get("traversability_step"));
    dump(out + "_roughness.bin", c.get("traversability_roughness"));
  }
  grid_map::GridMap d;
  auto ff = create(kFused);
  ff->configure("fused", Params{});
  if (!ff->update(m0, d)) { std::fprintf(stderr, "fused filter failed\n"); return 5; }
  expect(!d.exists("surface_normal_x"), "fused filter deletes the surface normals like the YAML DeletionFilter");
  d.convertToDefaultStartIndex();
  dump(out + "_fused_slope.bin", d.get("traversability_slope"));
  dump(out + "_fused_step.bin", d.get("traversability_step"));
  dump(out + "_fused_roughness.bin", d.get("traversability_roughness"));
  dump(out + "_fused_traversability.bin", d.get("traversability"));
  return g_fail;
}

// End-to-end time of the reference-named plugins exactly as filters::FilterChain drives them (host GridMaps in and out,
// `mapOut = mapIn` deep copies included): slopeFilter -> stepFilter -> roughnessFilter on a map that already carries the
// surface normals of the upstream NormalVectorsFilter.  One cell of the elevation changes per pass so every pass is a new map.
static int test_bench(int argc, char** argv) {
  if (argc < 6) { std::fprintf(stderr, "usage: bench R C res in.bin passes\n"); return 2; }
  const int rows = std::atoi(argv[2]), cols = std::atoi(argv[3]);
  const double res = std::atof(argv[4]);
  const int passes = argc >= 7 ? std::atoi(argv[6]) : 5;
  std::vector<float> elev((size_t)rows * cols);
  FILE* f = std::fopen(argv[5], "rb");
  if (!f || std::fread(elev.data(), sizeof(float), elev.size(), f) != elev.size()) { std::fprintf(stderr, "cannot read %s\n", argv[5]); return 3; }
  std::fclose(f);
  grid_map::GridMap m = make_map(rows, cols, res, 0.0, 0.0, elev);
  for (const char* l : {"surface_normal_x", "surface_normal_y", "surface_normal_z"}) m.add(l);
  te_ctx* ctx = nullptr;
  if (te_create(&ctx, 0) != TE_OK) { std::fprintf(stderr, "%s\n", te_last_error()); return 4; }
  te_geometry g{rows, cols, res, rows * res, cols * res, 0.0, 0.0, 0, 0};
  te_chain_params p{};
  p.normals_radius = 0.05; p.normals_algorithm = TE_NORMALS_FIXTURE; p.normals_positive_axis = 2; p.slope_critical = 1.0;
  p.step_critical = 0.12; p.step_first_radius = 0.04; p.step_second_radius = 0.04; p.step_critical_cells = 4;
  p.roughness_critical = 0.05; p.roughness_radius = 0.05; p.fuse_weight = 1.0f / 3.0f;
  if (te_normals(ctx, &g, &p, m.get("elevation").data(), m.get("surface_normal_x").data(), m.get("surface_normal_y").data(),
                 m.get("surface_normal_z").data(), TE_MEM_HOST) != TE_OK) { std::fprintf(stderr, "%s\n", te_last_error()); return 4; }
  te_destroy(ctx);
  auto fs = create(kSlope); fs->configure("slopeFilter", slope_yaml());
  auto ft = create(kStep); ft->configure("stepFilter", step_yaml());
  auto fr = create(kRough); fr->configure("roughnessFilter", rough_yaml());
  double best = 1e30, total = 0.0;
  for (int it = -1; it < passes; ++it) {   // pass -1 warms up (context creation, page-locked cache, normals check)
    m.get("elevation")(0, 0) = elev[0] + 1e-4f * (float)(it + 2);
    grid_map::GridMap a, b, c;
    const auto t0 = std::chrono::steady_clock::now();
    if (!fs->update(m, a) || !ft->update(a, b) || !fr->update(b, c)) { std::fprintf(stderr, "chain failed\n"); return 5; }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (it >= 0) { best = std::min(best, ms); total += ms; }
  }
  std::printf("PLUGIN_CHAIN rows=%d cols=%d passes=%d mean_ms=%.3f best_ms=%.3f launches=%ld served=%ld\n", rows, cols, passes,
              total / passes, best, te_plugin::ChainRegistry::instance().launches(), te_plugin::ChainRegistry::instance().served());
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: test_plugins configure|nogpu|chain|bench ...\n"); return 2; }
  const std::string mode = argv[1];
  if (mode == "configure") return test_configure();
  if (mode == "nogpu") return test_nogpu();
  if (mode == "chain") return test_chain(argc, argv);
  if (mode == "bench") return test_bench(argc, argv);
  return 2;
}
