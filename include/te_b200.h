/*
 * te_b200.h — C ABI of the B200-native traversability filter chain and footprint sweep.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry point
 * names the reference interface it replaces (path:line relative to the reference repository
 * leggedrobotics/traversability_estimation).  The reference-side binding a maintainer would add
 * (the filters::FilterBase<grid_map::GridMap> plugin shells) is in
 * traversability_estimation_b200/plugin/ and described in INTEGRATION.md.
 *
 * Layers are float32, column-major exactly like grid_map::Matrix (Eigen::MatrixXf):
 * value(i, j) = data[j * rows + i]; NaN/Inf = invalid cell (GridMap::isValid == std::isfinite).
 *
 * Every function returns TE_OK (0) or a negative te_status; it never throws.  The message of the
 * last failure on the calling thread is available from te_last_error().  There is NO CPU fallback:
 * without a CUDA device every compute entry point fails with TE_ERR_CUDA.
 */
#ifndef TE_B200_H
#define TE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TE_B200_ABI_VERSION 1

typedef enum te_status {
  TE_OK = 0,
  TE_ERR_BAD_ARG = -1,       /* null pointer, non-positive size, invalid parameter value */
  TE_ERR_MISSING_LAYER = -2, /* a required input layer pointer is null (reference: GridMap::at throws) */
  TE_ERR_CUDA = -3,          /* CUDA runtime/driver failure, or no device */
  TE_ERR_UNSUPPORTED = -4,   /* e.g. non-zero circular-buffer start index */
  TE_ERR_NCCL = -5           /* reserved (the halo exchange of this library is peer-mapped, see te_halo_pull) */
} te_status;

/* Where the layer pointers passed to a call live. */
typedef enum te_memory { TE_MEM_HOST = 0, TE_MEM_DEVICE = 1 } te_memory;

/* grid_map::GridMap geometry, doubles exactly as the container holds them; needed bit-for-bit
 * because CircleIterator decides window membership on absolute double positions.
 * (grid_map_core GridMap::getSize/getResolution/getLength/getPosition/getStartIndex) */
typedef struct te_geometry {
  int32_t rows, cols;
  double resolution;
  double length_x, length_y;
  double position_x, position_y;
  int32_t start_row, start_col; /* grid_map circular-buffer start index (GridMap::getStartIndex): cell (i, j) is stored at buffer
                                   index ((i + start_row) % rows, (j + start_col) % cols).  te_chain, te_footprint(2) and
                                   te_footprint_polygon with TE_MEM_HOST on a whole map (no slab) take and return layers in
                                   that order (the copies to and from the device unwrap / re-wrap); everything else wants 0
                                   (TE_ERR_UNSUPPORTED otherwise: convertToDefaultStartIndex) */
} te_geometry;

/* Column slab of a larger map (multi-GPU tiling, one slab per rank).  Inputs cover columns
 * [col_begin - halo_left, col_begin + col_count + halo_right) of the global map; outputs cover
 * [col_begin, col_begin + col_count).  Columns outside the global map need not exist.
 * Pass NULL for "the whole map". */
typedef struct te_slab {
  int32_t col_begin, col_count;
  int32_t halo_left, halo_right;
} te_slab;

enum { TE_NORMALS_FIXTURE = 0, TE_NORMALS_RAW_MOMENT = 1 };

/* Parameters of the YAML chain (traversability_estimation/config/robot_filter_parameter.yaml:2-37);
 * names follow the filters' own parameter names. */
typedef struct te_chain_params {
  double normals_radius;          /* NormalVectorsFilter `radius` (:8) */
  int32_t normals_algorithm;      /* TE_NORMALS_* */
  int32_t normals_positive_axis;  /* `normal_vector_positive_axis` 0=x 1=y 2=z (:9) */
  double slope_critical;          /* SlopeFilter `critical_value` (:14), SlopeFilter.cpp:36-46 */
  double step_critical;           /* StepFilter `critical_value` (:18), StepFilter.cpp:40-50 */
  double step_first_radius;       /* `first_window_radius` (:19) */
  double step_second_radius;      /* `second_window_radius` (:20) */
  int32_t step_critical_cells;    /* `critical_cell_number` (:21) */
  int32_t reserved0;
  double roughness_critical;      /* RoughnessFilter `critical_value` (:26), RoughnessFilter.cpp:38-48 */
  double roughness_radius;        /* `estimation_radius` (:27) */
  float fuse_weight;              /* MathExpressionFilter (:29-33): weight*((slope+step)+roughness) in float32 */
  int32_t reserved1;
} te_chain_params;

/* TraversabilityMap::traversabilityFootprint(radius, offset) and the members it reads. */
typedef struct te_footprint_params {
  double radius;                  /* = radiusMin; TraversabilityMap.cpp:313 */
  double offset;                  /* radiusMax = radius + offset */
  double traversability_default;  /* traversabilityDefault_ (robot_footprint_parameter.yaml:8) */
  double max_gap_width;           /* maxGapWidth_ (robot.yaml:10) */
  double critical_step_height;    /* criticalStepHeight_ (TraversabilityMap.cpp:117-126) */
  int32_t radius_is_integer_norm; /* SpiralIterator::getCurrentRadius via Eigen integer norm (1) or exact (0) */
  int32_t verify_roughness;       /* checkForRoughness_ (robot_footprint_parameter.yaml:9 `verify_roughness_footprint`):
                                     isTraversableForFilters also runs checkForRoughness (TraversabilityMap.cpp:779-783, 895-921);
                                     needs the traversability_roughness layer, i.e. te_footprint2 */
} te_footprint_params;

/* Which implementation te_chain uses. AUTO picks the fused stencil when the window shapes have a
 * specialised instantiation, else the generic kernel; both compute the same layers. */
typedef enum te_kernel { TE_KERNEL_AUTO = 0, TE_KERNEL_GENERIC = 1, TE_KERNEL_FUSED = 2 } te_kernel;

typedef struct te_ctx te_ctx;

/* Lifetime: one context per plugin instance / per rank.  Owns a non-blocking CUDA stream, the
 * per-geometry position tables and staging buffers.  Thread-safe: one in-flight call per context. */
int te_create(te_ctx** out, int device);
int te_destroy(te_ctx* ctx);
const char* te_last_error(void);
int te_abi_version(void);

/* Page-locked host memory for layers handed to TE_MEM_HOST calls (cudaHostAlloc): pageable layers work too, but their
 * transfers are staged by the driver at roughly half the PCIe rate.  The plugin shells keep their cached layers here. */
int te_host_alloc(void** out, size_t bytes);
int te_host_free(void* p);

/* Run subsequent calls on an external stream (cudaStream_t passed as void*; NULL restores the
 * context's own stream).  Device-memory calls are asynchronous on that stream; te_synchronize waits. */
int te_set_stream(te_ctx* ctx, void* cuda_stream);
int te_synchronize(te_ctx* ctx);
int te_set_kernel(te_ctx* ctx, int te_kernel_choice);
/* Counters since creation: kernels launched, cells that took the certified slow path in the fused kernel. */
int te_get_stats(te_ctx* ctx, int64_t* kernel_launches, int64_t* slow_path_cells);

/* Device-side timing of the kernels this context launches (CUDA events on the launching stream).
 * te_get_timing waits for the stream, returns the accumulated milliseconds of the main chain kernel
 * and of the fix-up kernel and the number of timed launches, then resets the accumulators. */
int te_enable_timing(te_ctx* ctx, int on);
int te_get_timing(te_ctx* ctx, double* main_ms, double* fixup_ms, int64_t* samples);
/* Work-list counters of the last fused launch: [0] cells the fp32 stencil could not certify (tier 2,
 * fp64 on centred coordinates); [1] work-list entries reserved for them (warp-private chunks, padding included);
 * [2] non-zero if a work list overflowed (cannot happen: the lists are sized for every cell of the launch);
 * [4] cells tier 2 passed on to the literal kernel (tier 3); [3] reserved (0). */
int te_get_flag_counters(te_ctx* ctx, uint32_t out[5]);
/* Why tier 2 passed cells of the last fused launch on to the literal kernel: reasons[r] = cells escalated for reason r
 * (1 degenerate / 2 rank-deficient / 3 small eigen-gap full window, 4 no convergence / 5 rank / 6 conditioning of a partial
 * window, 7 null eigenvector, 8 n_z == 0, 9 n_z on a float32 rounding boundary); valid_cells[n] = escalated cells whose
 * normals window held n valid cells.  Diagnostics only; no reference counterpart. */
int te_get_escalation_stats(te_ctx* ctx, uint32_t reasons[16], uint32_t valid_cells[26]);
/* Work decomposition the fused kernel uses for `nmaps` maps of rows x out_ncols output cells on a GPU with `sms`
 * multiprocessors (host arithmetic only, needs no GPU; no reference counterpart — the reference iterates cell by cell).
 * Units are (level, map, column segment, 60-row strip), popped from a queue in that order; a level is a run of columns cut
 * into segments of one length.  out[0] = strips per map, out[1] = levels, then per level {first unit, first column,
 * segment length, segments per map}, then the total number of units: 2 + 4*levels + 1 values (levels <= 4, so 19). */
int te_fused_plan(int rows, int out_ncols, int nmaps, int sms, int32_t out[19]);

/* filters::SlopeFilter<grid_map::GridMap>::update — traversability_estimation_filters/src/SlopeFilter.cpp:59-89.
 * in: surface_normal_z, out: the `map_type` layer. */
int te_slope(te_ctx* ctx, const te_geometry* g, double critical_value, const float* surface_normal_z,
             float* out, int memory);

/* grid_map::NormalVectorsFilter::update (area method) — third-party, configured at
 * robot_filter_parameter.yaml:3-9.  out: surface_normal_{x,y,z}. */
int te_normals(te_ctx* ctx, const te_geometry* g, const te_chain_params* p, const float* elevation,
               float* nx, float* ny, float* nz, int memory);

/* filters::StepFilter<grid_map::GridMap>::update — StepFilter.cpp:102-182 (both passes; the
 * temporary step_height layer never leaves the device). */
int te_step(te_ctx* ctx, const te_geometry* g, const te_chain_params* p, const float* elevation,
            float* out, int memory);

/* filters::RoughnessFilter<grid_map::GridMap>::update — RoughnessFilter.cpp:73-132. */
int te_roughness(te_ctx* ctx, const te_geometry* g, const te_chain_params* p, const float* elevation,
                 const float* nx, const float* ny, const float* nz, float* out, int memory);

/* The whole chain filters::FilterChain<grid_map::GridMap>::update runs at TraversabilityMap.cpp:214:
 * normals -> slope, step, roughness -> weighted sum; normals are deleted unless pointers are given
 * (DeletionFilter, robot_filter_parameter.yaml:34-37).  `slab` NULL = whole map. */
int te_chain(te_ctx* ctx, const te_geometry* g, const te_slab* slab, const te_chain_params* p,
             const float* elevation, float* slope, float* step, float* roughness, float* traversability,
             float* nx_or_null, float* ny_or_null, float* nz_or_null, int memory);

/* nmaps independent maps of identical geometry, stored back to back (map m at offset m*rows*cols). */
int te_chain_batched(te_ctx* ctx, const te_geometry* g, const te_chain_params* p, int32_t nmaps,
                     const float* elevation, float* slope, float* step, float* roughness,
                     float* traversability, int memory);

/* TraversabilityMap::traversabilityFootprint(const double& radius, const double& offset) —
 * traversability_estimation/src/TraversabilityMap.cpp:307-318 with isTraversable (:654-746),
 * isTraversableForFilters (:774-792), checkForStep (:794-865), checkForSlope (:867-893).
 * out: traversability_footprint; slope_footprint/step_footprint memoisation layers optional. */
int te_footprint(te_ctx* ctx, const te_geometry* g, const te_slab* slab, const te_footprint_params* p,
                 const float* traversability, const float* slope, const float* step, const float* elevation,
                 float* traversability_footprint, float* slope_footprint_or_null, float* step_footprint_or_null,
                 int memory);

/* te_footprint with the traversability_roughness layer: when p->verify_roughness is set, a visited cell is also blocked by
 * checkForRoughness (TraversabilityMap.cpp:895-921: more than floor(1.5 * 3 res * max_gap_width / 3 / res^2) cells of zero
 * roughness traversability within 3 res).  `roughness` may be NULL when the flag is clear; roughness_footprint receives the
 * memoisation layer the reference keeps (optional). */
int te_footprint2(te_ctx* ctx, const te_geometry* g, const te_slab* slab, const te_footprint_params* p,
                  const float* traversability, const float* slope, const float* step, const float* roughness_or_null,
                  const float* elevation, float* traversability_footprint, float* slope_footprint_or_null,
                  float* step_footprint_or_null, float* roughness_footprint_or_null, int memory);

/* TraversabilityMap::traversabilityFootprint(double footprintYaw), TraversabilityMap.cpp:239-305 — what the reference's
 * `traversability_footprint` service runs (TraversabilityEstimation.cpp:272-276): every cell gets the footprint polygon
 * (footprint/footprint_polygon, robot_footprint_parameter.yaml:3; `polygon_xy` = npts vertices (x, y) in the footprint frame,
 * 3..16) placed at its centre, unrotated -> layer traversability_x, rotated by footprint_yaw about z -> traversability_rot.  Each
 * value is isTraversable(polygon, traversability) (:592-645): 0 when a cell inside the polygon (grid_map::PolygonIterator) fails
 * isTraversableForFilters, else the mean of the traversability layer over those cells (traversability_default for invalid
 * cells, and when the polygon covers no cell).  Uses p->traversability_default, max_gap_width, critical_step_height,
 * verify_roughness (radius / offset are ignored).  Slab halo: te_footprint's predicate halo + the polygon's reach.
 * TE_ERR_UNSUPPORTED for a polygon that reaches further than 31 cells from its centre. */
int te_footprint_polygon(te_ctx* ctx, const te_geometry* g, const te_slab* slab, const te_footprint_params* p, int32_t npts,
                         const double* polygon_xy, double footprint_yaw, const float* traversability, const float* slope,
                         const float* step, const float* roughness_or_null, const float* elevation, float* traversability_x,
                         float* traversability_rot, int memory);

/* TraversabilityMap::checkFootprintPath for circular footprints — checkCircularFootprintPath, TraversabilityMap.cpp:345-462 —
 * for a BATCH of paths in one launch (one thread per path: the service callback of the reference checks one path per call;
 * planners and MPC roll-outs ask for hundreds).  It is evaluated on a complete traversability_footprint layer, i.e. the output
 * of te_footprint(radius, offset) for the radius of the paths: every isTraversable(center, radius + offset, ...) of the reference
 * then takes its memoised branch (:667-673: traversability = layer value, traversable iff value != 0), a centre outside the map the
 * default branch (:660-666).  Path q is the poses poses_xy[2*path_begin[q] .. 2*path_begin[q+1]) (x, y in the map frame); for a
 * path of one pose the circle at the pose is checked (:365-390), otherwise every fourth cell (nSkip = 3, :401) of the grid line
 * between consecutive poses, and the segment means are combined weighted by segment length (:437-449).  Outputs per path:
 * TraversabilityResult.is_safe and .traversability (0 when unsafe); .area is 0 for circular footprints.  Not covered:
 * the untraversable polygon, publishing.  Poses of a multi-pose path must lie inside the map
 * (the reference does not check getIndex's return value there): such a path is reported unsafe. */
int te_check_footprint_paths(te_ctx* ctx, const te_geometry* g, const float* traversability_footprint,
                             double traversability_default, int32_t npaths, const int32_t* path_begin, const double* poses_xy,
                             uint8_t* is_safe, double* traversability, int memory);

/* te_check_footprint_paths with checkRobotInclination_ set (TraversabilityMap.cpp:359-363, :386-390): before the circles of a
 * pose / segment are looked at, TraversabilityMap::checkInclination (:748-762) reads the `robot_slope` layer (robotSlopeType_,
 * config/robot.yaml:1; column-major like every layer) — at the pose for a single pose, along LineIterator(start, end) otherwise,
 * skipping invalid cells — and the path is unsafe as soon as a cell is exactly 0.0.  robot_slope_or_null == NULL is the call above.
 * A single pose outside the map (the reference's atPosition throws) is reported unsafe. */
int te_check_footprint_paths2(te_ctx* ctx, const te_geometry* g, const float* traversability_footprint,
                              const float* robot_slope_or_null, double traversability_default, int32_t npaths,
                              const int32_t* path_begin, const double* poses_xy, uint8_t* is_safe, double* traversability, int memory);

/* ---- Multi-GPU: one map tiled into column slabs, one process (rank) per GPU -------------------------------
 * The chain and the footprint sweep are stencils of fixed radius, so the only exchange step is a one-shot copy of the
 * neighbours' boundary columns of the INPUT layer(s) into this rank's halo.  The reference has no counterpart (it is a
 * single-process CPU node; TraversabilityMap.cpp:202-237 runs the chain on one whole map); these entry points are what a
 * sharded caller of te_chain / te_footprint needs around the `te_slab` argument:
 *   1. every rank exports its slab buffer (te_ipc_export) and a "layer ready" event (te_event_create_ipc), ships the two
 *      handles (TE_IPC_HANDLE_BYTES and 64 bytes) to its neighbours over any channel (MPI, a socket, torch.distributed), and opens theirs
 *      (te_ipc_open / te_event_open_ipc) — once;
 *   2. per map: te_event_record(ready) after the rank's producer wrote its owned columns; te_halo_pull() then waits for the
 *      neighbours' ready events on the context stream and copies their boundary columns straight out of their buffers over
 *      NVLink (peer-mapped cudaMemcpyAsync, no staging, no collective); te_chain(..., slab, ..., TE_MEM_DEVICE) follows on the
 *      same stream.  Outputs stay sharded.
 * A cross-process event wait sees the most recent te_event_record that the RECORDING process had issued when the waiting
 * process called te_halo_pull; callers that rewrite a layer per frame order the two host-side (a message after the record)
 * and must not overwrite boundary columns a neighbour may still be pulling (double-buffer, or wait for the neighbour's own
 * event recorded after its pull). */
#define TE_IPC_HANDLE_BYTES 80 /* CUDA IPC handle of the containing allocation + offset of the pointer inside it */
int te_ipc_export(const void* device_ptr, void* handle_TE_IPC_HANDLE_BYTES);
int te_ipc_open(const void* handle_TE_IPC_HANDLE_BYTES, void** device_ptr_out);
int te_ipc_close(void* device_ptr);
/* Interprocess events (cudaEventInterprocess | cudaEventDisableTiming). */
int te_event_create_ipc(te_ctx* ctx, void** event_out, void* handle_64_bytes_out);
int te_event_open_ipc(const void* handle_64_bytes, void** event_out);
int te_event_record(te_ctx* ctx, void* event);   /* on the context stream */
int te_event_destroy(void* event);

/* A neighbour's slab buffer as mapped into this process: `layer` holds the global columns
 * [slab.col_begin - slab.halo_left, slab.col_begin + slab.col_count + slab.halo_right) of a rows x cols layer. */
typedef struct te_halo_peer {
  const float* layer;   /* te_ipc_open()ed pointer (or a plain device pointer of a peer-accessible GPU in this process) */
  te_slab slab;         /* the neighbour's slab */
  void* ready_event;    /* te_event_open_ipc()ed event, or NULL: no wait */
} te_halo_peer;

/* Fill this rank's halo columns of `layer` (a buffer laid out like the te_chain input for `slab`) from the neighbours' OWNED
 * columns, asynchronously on the context stream.  The exchange is one hop: TE_ERR_BAD_ARG if a neighbour owns fewer columns
 * than the halo needs.  Pass NULL for a side without a neighbour (map edge). */
int te_halo_pull(te_ctx* ctx, const te_geometry* g, const te_slab* slab, float* layer,
                 const te_halo_peer* left_or_null, const te_halo_peer* right_or_null);

#ifdef __cplusplus
}
#endif
#endif
