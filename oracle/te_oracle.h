/*
 * te_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A from-scratch restatement, in plain C++17 with a C ABI, of the arithmetic of the reference's
 * filter chain and footprint sweep.  It exists only so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / `--impl reference` leg have something to check against and to time.
 * Nothing in traversability_estimation_b200/ may include, link or call it.
 *
 * The reference itself (leggedrobotics/traversability_estimation) cannot be compiled in this
 * image: every translation unit needs ROS `filters`, `pluginlib`, `grid_map_*` and Eigen, none of
 * which are installed.  The functions below therefore follow
 *   - in-tree sources, cited per function as path:line relative to /root/reference, and
 *   - the recalled behaviour of the un-vendored third-party `grid_map_core` / `grid_map_filters`
 *     (ros-noetic-grid-map 1.6.x; not pinned by the reference, package.xml:14-21) as written down
 *     in SURVEY.md Appendix A.
 * Pinning: the chain part is pinned by the reference's bag fixture (tests/golden/fixture_gridmap.npz,
 * 13 300/13 300 cells bit-exact for all four output layers, see tests/test_oracle_fixture.py).
 * The footprint sweep, NaN-hole handling and on-circle window membership are NOT pinned by any
 * reference data ("parity unpinned" for those; decisions are listed in oracle/README.md).
 */
#ifndef TE_ORACLE_H
#define TE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* grid_map::GridMap geometry, doubles exactly as the container holds them. */
typedef struct teo_geometry {
  int32_t rows, cols;        /* getSize(): rows = x direction, cols = y direction */
  double resolution;         /* getResolution() */
  double length_x, length_y; /* getLength() */
  double position_x, position_y; /* getPosition(): map centre */
} teo_geometry;

enum { TEO_NORMALS_FIXTURE = 0, /* mean-subtracted scatter + rank test (Appendix A.4b) - pinned by fixture */
       TEO_NORMALS_RAW_MOMENT = 1 /* raw second moment + 1e-8 eigenvalue test (Appendix A.4a) */ };

/* Parameters of the YAML chain, traversability_estimation/config/robot_filter_parameter.yaml:2-37 */
typedef struct teo_chain_params {
  double normals_radius;          /* NormalVectorsFilter radius */
  int32_t normals_algorithm;      /* TEO_NORMALS_* */
  int32_t normals_positive_axis;  /* 0=x 1=y 2=z (normal_vector_positive_axis) */
  double slope_critical;          /* SlopeFilter critical_value */
  double step_critical;           /* StepFilter critical_value */
  double step_first_radius;       /* first_window_radius */
  double step_second_radius;      /* second_window_radius */
  int32_t step_critical_cells;    /* critical_cell_number */
  int32_t reserved0;
  double roughness_critical;      /* RoughnessFilter critical_value */
  double roughness_radius;        /* estimation_radius */
  float fuse_weight;              /* MathExpressionFilter: weight * ((slope + step) + roughness), float32 */
  int32_t reserved1;
} teo_chain_params;

/* Parameters of TraversabilityMap::traversabilityFootprint(radius, offset) and the members it reads. */
typedef struct teo_footprint_params {
  double radius;                  /* footprint radius = radiusMin, TraversabilityMap.cpp:313 */
  double offset;                  /* radiusMax = radius + offset */
  double traversability_default;  /* traversabilityDefault_, robot_footprint_parameter.yaml:8 */
  double max_gap_width;           /* maxGapWidth_, robot.yaml:10 */
  double critical_step_height;    /* criticalStepHeight_ = stepFilter critical_value, TraversabilityMap.cpp:117-126 */
  int32_t radius_is_integer_norm; /* 1: SpiralIterator::getCurrentRadius uses Eigen's integer norm (floor), 0: exact */
  int32_t verify_roughness;       /* checkForRoughness_ (robot_footprint_parameter.yaml:9 verify_roughness_footprint) */
} teo_footprint_params;

/* All layers: float32, column-major, value(i,j) = data[j*rows + i]; NaN/Inf = invalid cell.
 * nthreads <= 0 → all host threads (OpenMP).  Return 0 on success, non-zero on bad arguments. */

int teo_normals(const teo_geometry* g, const teo_chain_params* p, const float* elevation,
                float* nx, float* ny, float* nz, int nthreads);
int teo_slope(const teo_geometry* g, double critical, const float* nz, float* out, int nthreads);
int teo_step(const teo_geometry* g, const teo_chain_params* p, const float* elevation, float* out,
             float* step_height_or_null, int nthreads);
int teo_roughness(const teo_geometry* g, const teo_chain_params* p, const float* elevation,
                  const float* nx, const float* ny, const float* nz, float* out, int nthreads);
int teo_fuse(int64_t n, float weight, const float* slope, const float* step, const float* roughness, float* out);
/* Whole YAML chain; normals outputs may be NULL (DeletionFilter semantics). */
int teo_chain(const teo_geometry* g, const teo_chain_params* p, const float* elevation,
              float* slope, float* step, float* roughness, float* traversability,
              float* nx_or_null, float* ny_or_null, float* nz_or_null, int nthreads);

/* Footprint sweep: TraversabilityMap::traversabilityFootprint(radius, offset), TraversabilityMap.cpp:307-318.
 * slope_fp / step_fp receive the memoisation layers as the reference leaves them (may be NULL). */
int teo_footprint(const teo_geometry* g, const teo_footprint_params* p, const float* traversability,
                  const float* slope, const float* step, const float* elevation,
                  float* out_footprint, float* slope_fp_or_null, float* step_fp_or_null, int nthreads);

/* The same with checkForRoughness (TraversabilityMap.cpp:779-783, 895-921) when p->verify_roughness is set: `roughness` is the
 * traversability_roughness layer, rough_fp receives the roughness_footprint memoisation layer (may be NULL). */
int teo_footprint2(const teo_geometry* g, const teo_footprint_params* p, const float* traversability,
                   const float* slope, const float* step, const float* roughness_or_null, const float* elevation,
                   float* out_footprint, float* slope_fp_or_null, float* step_fp_or_null, float* rough_fp_or_null, int nthreads);

/* TraversabilityMap::traversabilityFootprint(double footprintYaw) (TraversabilityMap.cpp:239-305) with the polygon
 * isTraversable (:592-645): layers traversability_x (footprint polygon at every cell centre, unrotated) and traversability_rot
 * (rotated by yaw about z).  pts_xy: npts footprint vertices (x, y) in the footprint frame (robot_footprint_parameter.yaml:3). */
int teo_footprint_polygon(const teo_geometry* g, const teo_footprint_params* p, int npts, const double* pts_xy, double yaw,
                          const float* traversability, const float* slope, const float* step, const float* roughness_or_null,
                          const float* elevation, float* out_traversability_x, float* out_traversability_rot, int nthreads);

/* TraversabilityMap::checkCircularFootprintPath (TraversabilityMap.cpp:345-462) for a batch of paths on a
 * traversability_footprint layer that is valid everywhere (memoised isTraversable branch :667-673): path q is the poses
 * poses_xy[2*path_begin[q] .. 2*path_begin[q+1]).  Outputs TraversabilityResult.is_safe / .traversability per path. */
int teo_check_circular_paths(const teo_geometry* g, const float* traversability_footprint, double traversability_default,
                             int npaths, const int32_t* path_begin, const double* poses_xy, uint8_t* is_safe,
                             double* traversability);
/* ... with checkRobotInclination_ (checkInclination, TraversabilityMap.cpp:748-762) on the `robot_slope` layer (NULL: off). */
int teo_check_circular_paths2(const teo_geometry* g, const float* traversability_footprint, const float* robot_slope_or_null,
                              double traversability_default, int npaths, const int32_t* path_begin, const double* poses_xy,
                              uint8_t* is_safe, double* traversability);

/* Visit order of grid_map::SpiralIterator for a centre far from the map border: writes up to `cap`
 * (di,dj) pairs, returns the number of cells visited.  radius/resolution in metres. */
int teo_spiral_offsets(double radius, double resolution, int32_t* di, int32_t* dj, int cap);

/* Window membership table of grid_map::CircleIterator for centre cell (i,j): returns count and writes
 * the (a,b) absolute indices in iteration order. */
int teo_circle_cells(const teo_geometry* g, int i, int j, double radius, int32_t* a, int32_t* b, int cap);

int teo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
