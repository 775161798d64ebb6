// te_capi.cu — the extern "C" boundary of libte_b200 (see include/te_b200.h).
// Host-side responsibilities only: argument validation with the reference's conventions, the
// per-geometry position tables, host<->device staging for TE_MEM_HOST callers, kernel selection.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/te_b200.h"
#include "te_kernels.h"
#include "te_fused.h"
#include "te_footprint.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define TE_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) return fail(TE_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// grid_map::getPositionFromIndex operand order.
inline double cell_coord(double map_pos, double length, double res, int idx) {
  const double offset = 0.5 * length - 0.5 * res;
  return (map_pos + offset) + res * (-(double)idx);
}

// Largest index offset that can satisfy the circle test.  A cell at offset R + 1 lies (R + 1) * res from the centre up to a few
// ulps of the absolute coordinates (~1e-14 m); it can pass `d^2 <= r^2` only if radius / res is within rounding of R + 1, and
// only then is the guard cell scanned (the literal kernels test every candidate with the reference's own arithmetic anyway).
inline int reach_generous(double radius, double res) {
  const double q = radius / res;
  const int R = (int)std::floor(q);
  return (q - (double)R > 1.0 - 1e-6) ? R + 1 : R;
}
// Dependency radius in cells (what a slab halo must provide).
inline int reach_true(double radius, double res) { return (int)std::floor(radius / res + 1e-9); }

}  // namespace

struct te_ctx {
  int device = 0;
  int sms = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  int kernel_choice = TE_KERNEL_AUTO;
  int64_t launches = 0;
  int64_t slow_cells = 0;

  te_geometry geo{};
  bool have_geo = false;
  DevBuf dX, dY;
  std::vector<double> hX, hY;

  DevBuf stage[12];          // TE_MEM_HOST staging: 0..3 inputs, 4..11 outputs
  DevBuf worklist, worklist3, counter;  // fused-kernel fix-up lists (tier 2, tier 3) and their counters
  // The counters are two 512-byte blocks used alternately: the last kernel of a chain call (k_fixup_cells) zeroes the block of the
  // NEXT call, so a call needs no cudaMemsetAsync of its own (one stream operation and one launch gap less per map).
  int counter_phase = 0;    // block of the most recent fused launch (what the statistics entry points read)
  int counter_clean = -1;   // block known to be zero on the stream (-1: none)
  cudaStream_t counter_stream = nullptr;  // the stream whose order that knowledge belongs to
  te::FusedState fused;      // tensor maps / tables of the fused stencil
  te::FootprintState fp;

  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;  // te_chain(TE_MEM_HOST) pipeline
  bool timing = false;
  struct Ev3 { cudaEvent_t a, b, c; };
  std::vector<Ev3> events;   // timing events, created once and reused (te_get_timing rewinds events_used)
  size_t events_used = 0;
};

namespace {

int check_geometry(const te_geometry* g, bool allow_start_index = false) {
  if (!g) return fail(TE_ERR_BAD_ARG, "geometry is null");
  if (g->rows <= 0 || g->cols <= 0) return fail(TE_ERR_BAD_ARG, "map size must be positive (rows=%d cols=%d)", g->rows, g->cols);
  if (!(g->resolution > 0.0) || !std::isfinite(g->resolution)) return fail(TE_ERR_BAD_ARG, "resolution must be positive");
  if (g->start_row != 0 || g->start_col != 0) {
    if (!allow_start_index)
      return fail(TE_ERR_UNSUPPORTED, "circular-buffer start index (%d,%d) != (0,0): call convertToDefaultStartIndex() first",
                  g->start_row, g->start_col);
    if (g->start_row < 0 || g->start_row >= g->rows || g->start_col < 0 || g->start_col >= g->cols)
      return fail(TE_ERR_BAD_ARG, "circular-buffer start index (%d,%d) outside the map", g->start_row, g->start_col);
  }
  if ((long long)g->rows * g->cols > 0x7fffffffLL * 2) return fail(TE_ERR_UNSUPPORTED, "map has more than 2^32 cells");
  return TE_OK;
}

// Same validation as the filters' configure(): SlopeFilter.cpp:41, StepFilter.cpp:45,58,71,84,
// RoughnessFilter.cpp:43,55.
int check_params(const te_chain_params* p) {
  if (!p) return fail(TE_ERR_BAD_ARG, "chain parameters are null");
  if (!(p->normals_radius >= 0.0)) return fail(TE_ERR_BAD_ARG, "normals radius must be >= 0");
  if (p->normals_algorithm != TE_NORMALS_FIXTURE && p->normals_algorithm != TE_NORMALS_RAW_MOMENT)
    return fail(TE_ERR_BAD_ARG, "unknown normals algorithm %d", p->normals_algorithm);
  if (p->normals_positive_axis < 0 || p->normals_positive_axis > 2) return fail(TE_ERR_BAD_ARG, "positive axis must be 0, 1 or 2");
  if (p->slope_critical > M_PI_2 || p->slope_critical < 0.0 || std::isnan(p->slope_critical))
    return fail(TE_ERR_BAD_ARG, "Critical slope must be in the interval [0, PI/2]");
  if (!(p->step_critical >= 0.0)) return fail(TE_ERR_BAD_ARG, "Critical step height must be greater than zero");
  if (!(p->step_first_radius >= 0.0) || !(p->step_second_radius >= 0.0))
    return fail(TE_ERR_BAD_ARG, "step window radii must be greater than zero");
  if (p->step_critical_cells <= 0) return fail(TE_ERR_BAD_ARG, "Number of critical cells must be greater than zero");
  if (!(p->roughness_critical >= 0.0)) return fail(TE_ERR_BAD_ARG, "Critical roughness must be greater than zero");
  if (!(p->roughness_radius >= 0.0)) return fail(TE_ERR_BAD_ARG, "Roughness estimation radius must be greater than zero");
  return TE_OK;
}

int ensure_geometry(te_ctx* c, const te_geometry* g) {
  if (c->have_geo && std::memcmp(&c->geo, g, sizeof(te_geometry)) == 0) return TE_OK;
  c->hX.resize(g->rows);
  c->hY.resize(g->cols);
  for (int i = 0; i < g->rows; ++i) c->hX[i] = cell_coord(g->position_x, g->length_x, g->resolution, i);
  for (int j = 0; j < g->cols; ++j) c->hY[j] = cell_coord(g->position_y, g->length_y, g->resolution, j);
  TE_CUDA(c->dX.reserve(sizeof(double) * g->rows));
  TE_CUDA(c->dY.reserve(sizeof(double) * g->cols));
  // The tables may still be in use by kernels queued on the stream: order the overwrite after them.
  TE_CUDA(cudaMemcpyAsync(c->dX.p, c->hX.data(), sizeof(double) * g->rows, cudaMemcpyHostToDevice, c->stream));
  TE_CUDA(cudaMemcpyAsync(c->dY.p, c->hY.data(), sizeof(double) * g->cols, cudaMemcpyHostToDevice, c->stream));
  TE_CUDA(cudaStreamSynchronize(c->stream));  // hX/hY are reused
  c->geo = *g;
  c->have_geo = true;
  c->fused.invalidate();
  c->fp.invalidate();
  return TE_OK;
}

te::ChainDev make_chain_dev(const te_geometry* g, const te_chain_params* p) {
  te::ChainDev d{};
  const double res = g->resolution;
  d.rn = p->normals_radius; d.rn2 = d.rn * d.rn; d.Rn = reach_generous(d.rn, res);
  d.alg = p->normals_algorithm; d.axis = p->normals_positive_axis;
  d.slope_crit = p->slope_critical;
  d.step_crit = p->step_critical;
  d.r1 = p->step_first_radius; d.r1sq = d.r1 * d.r1; d.R1 = reach_generous(d.r1, res);
  d.r2 = p->step_second_radius; d.r2sq = d.r2 * d.r2; d.R2 = reach_generous(d.r2, res);
  d.ncrit = p->step_critical_cells;
  d.rough_crit = p->roughness_critical;
  d.rr = p->roughness_radius; d.rr2 = d.rr * d.rr; d.Rr = reach_generous(d.rr, res);
  d.fuse_w = p->fuse_weight;
  return d;
}

int chain_halo(const te_geometry* g, const te_chain_params* p) {
  const double res = g->resolution;
  const int hn = reach_true(p->normals_radius, res), hr = reach_true(p->roughness_radius, res);
  const int hs = reach_true(p->step_first_radius, res) + reach_true(p->step_second_radius, res);
  return std::max(hn, std::max(hr, hs));
}

int resolve_slab(const te_geometry* g, const te_slab* s, int need_halo, te_slab* out) {
  if (!s) {
    *out = te_slab{0, g->cols, 0, 0};
    return TE_OK;
  }
  if (s->col_begin < 0 || s->col_count <= 0 || s->col_begin + s->col_count > g->cols)
    return fail(TE_ERR_BAD_ARG, "slab columns [%d,%d) outside map of %d columns", s->col_begin, s->col_begin + s->col_count, g->cols);
  if (s->halo_left < 0 || s->halo_right < 0 || s->halo_left > s->col_begin || s->halo_right > g->cols - (s->col_begin + s->col_count))
    return fail(TE_ERR_BAD_ARG, "slab halo (%d,%d) reaches outside the map", s->halo_left, s->halo_right);
  const int need_l = std::min(need_halo, s->col_begin);
  const int need_r = std::min(need_halo, g->cols - (s->col_begin + s->col_count));
  if (s->halo_left < need_l || s->halo_right < need_r)
    return fail(TE_ERR_BAD_ARG, "slab halo (%d,%d) smaller than the dependency radius %d of these parameters", s->halo_left,
                s->halo_right, need_halo);
  *out = *s;
  return TE_OK;
}

te::SlabView make_view(te_ctx* c, const te_geometry* g, const te_slab& s) {
  te::SlabView v{};
  v.rows = g->rows;
  v.cols_total = g->cols;
  v.in_col0 = s.col_begin - s.halo_left;
  v.in_ncols = s.halo_left + s.col_count + s.halo_right;
  v.out_col0 = s.col_begin;
  v.out_ncols = s.col_count;
  v.X = (const double*)c->dX.p;
  v.Y = (const double*)c->dY.p;
  v.res = g->resolution;
  v.coord_max = std::max(std::fabs(g->position_x) + 0.5 * g->length_x, std::fabs(g->position_y) + 0.5 * g->length_y);
  return v;
}

struct Guard {
  te_ctx* c;
  int prev = -1;
  bool ok = false;
  explicit Guard(te_ctx* ctx) : c(ctx) {
    c->mu.lock();
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(c->device) == cudaSuccess;
  }
  ~Guard() {
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    c->mu.unlock();
  }
};

#define TE_ENTER(ctx)                                                   \
  if (!(ctx)) return fail(TE_ERR_BAD_ARG, "context is null");           \
  Guard guard__(ctx);                                                   \
  if (!guard__.ok) return fail(TE_ERR_CUDA, "cudaSetDevice(%d) failed", (ctx)->device)

int launch_check(te_ctx* c, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TE_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  ++c->launches;
  return TE_OK;
}

// The next triple of timing events; the pool grows on demand and is reused after te_get_timing.
int next_timing_slot(te_ctx* c, te_ctx::Ev3** out) {
  if (c->events_used == c->events.size()) {
    te_ctx::Ev3 ev{};
    TE_CUDA(cudaEventCreate(&ev.a)); TE_CUDA(cudaEventCreate(&ev.b)); TE_CUDA(cudaEventCreate(&ev.c));
    c->events.push_back(ev);
  }
  *out = &c->events[c->events_used];
  return TE_OK;
}

// Device-memory chain on one slab: picks the kernel.
int run_chain_device(te_ctx* c, const te_geometry* g, const te::SlabView& v, const te_chain_params* p, const float* elev,
                     const te::ChainOut& o, int nmaps) {
  const te::ChainDev d = make_chain_dev(g, p);
  bool use_fused = false;
  if (c->kernel_choice != TE_KERNEL_GENERIC) {
    use_fused = te::fused_eligible(c->fused, c->hX, c->hY, g, p, c->stream);
    if (!use_fused && c->kernel_choice == TE_KERNEL_FUSED)
      return fail(TE_ERR_UNSUPPORTED, "fused stencil has no instantiation for these window shapes: %s", c->fused.why.c_str());
    if (use_fused) {
      if (const char* why = te::fused_launch_obstacle(v, nmaps, elev, o)) {
        if (c->kernel_choice == TE_KERNEL_FUSED) return fail(TE_ERR_UNSUPPORTED, "fused stencil cannot run this launch: %s", why);
        use_fused = false;  // TE_KERNEL_AUTO: the generic kernel computes the same layers
      }
    }
  }
  const size_t in_stride = (size_t)g->rows * v.in_ncols, out_stride = (size_t)g->rows * v.out_ncols;
  if (use_fused) {
    // The work lists hold every cell of the launch (plus chunk padding), so they cannot overflow: a degenerate map (exact planes,
    // holes everywhere) sends all of its cells down the certified slow path instead of returning uncertified fp32 values.
    const size_t cells = out_stride * (size_t)nmaps;
    const size_t cap = te::fused_list_capacity(cells, c->sms);
    if (cap >= ((size_t)1 << 32)) return fail(TE_ERR_UNSUPPORTED, "launch of %zu cells exceeds the work-list index range", cells);
    TE_CUDA(c->worklist.reserve(sizeof(unsigned) * cap));
    TE_CUDA(c->worklist3.reserve(sizeof(unsigned) * cells));
    {
      const void* before = c->counter.p;
      TE_CUDA(c->counter.reserve(sizeof(unsigned) * 256));
      if (c->counter.p != before) c->counter_clean = -1;
    }
    {  // one launch covers every map of the batch
      const te::ChainOut& om = o;
      const int phase = c->counter_phase ^ 1;
      unsigned* const cnt = (unsigned*)c->counter.p + 128 * phase;
      unsigned* const cnt_next = (unsigned*)c->counter.p + 128 * (phase ^ 1);
      if (c->counter_stream != c->stream) c->counter_clean = -1;  // zeroed in another stream's order: not ordered before this launch
      c->counter_stream = c->stream;
      if (c->counter_clean != phase) TE_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned) * 128, c->stream));
      c->counter_clean = -1;
      c->counter_phase = phase;
      te_ctx::Ev3* ev = nullptr;
      if (c->timing) {
        if (int rc = next_timing_slot(c, &ev)) return rc;
        TE_CUDA(cudaEventRecord(ev->a, c->stream));
      }
      int rc = te::launch_chain_fused(c->fused, v, d, nmaps, elev, om, (unsigned*)c->worklist.p, cnt, (unsigned)cap, c->sms, c->stream);
      if (rc != 0) return fail(TE_ERR_CUDA, "fused chain launch failed: %s", c->fused.why.c_str());
      if (int r2 = launch_check(c, "k_chain_fused")) return r2;
      if (c->timing) TE_CUDA(cudaEventRecord(ev->b, c->stream));
      te::FixupArgs fa;
      te::make_fixup_args(c->fused, v, d, &fa);
      // tiers 2 and 3 are programmatic dependent launches unless events are recorded in between (timing): their grids are set
      // up while the predecessor drains and wait on griddepcontrol.wait before they read the lists
      te::launch_fixup_t2(fa, elev, om, (const unsigned*)c->worklist.p, cnt, (unsigned)cap, (unsigned*)c->worklist3.p, cnt + 4,
                          (unsigned)cells, c->sms, c->stream, !c->timing);
      if (int r2 = launch_check(c, "k_fixup_t2")) return r2;
      te::launch_fixup(v, d, elev, om, (const unsigned*)c->worklist3.p, cnt + 4, (unsigned)cells, cnt_next, c->sms, c->stream, true);
      if (int r2 = launch_check(c, "k_fixup_cells")) return r2;
      c->counter_clean = phase ^ 1;  // k_fixup_cells zeroes the other block
      if (c->timing) {
        TE_CUDA(cudaEventRecord(ev->c, c->stream));
        ++c->events_used;
      }
    }
  } else {
    for (int m = 0; m < nmaps; ++m) {
      te::ChainOut om = o;
      om.slope += m * out_stride; om.step += m * out_stride; om.rough += m * out_stride; om.trav += m * out_stride;
      if (om.nx) om.nx += m * out_stride;
      if (om.ny) om.ny += m * out_stride;
      if (om.nz) om.nz += m * out_stride;
      te_ctx::Ev3* ev = nullptr;
      if (c->timing) {
        if (int rc = next_timing_slot(c, &ev)) return rc;
        TE_CUDA(cudaEventRecord(ev->a, c->stream));
      }
      te::launch_chain_generic(v, d, elev + m * in_stride, om, c->sms, c->stream);
      if (int r2 = launch_check(c, "k_chain_generic")) return r2;
      if (c->timing) {
        TE_CUDA(cudaEventRecord(ev->b, c->stream));
        TE_CUDA(cudaEventRecord(ev->c, c->stream));
        ++c->events_used;
      }
    }
  }
  return TE_OK;
}

}  // namespace

extern "C" {

int te_abi_version(void) { return TE_B200_ABI_VERSION; }

const char* te_last_error(void) { return g_last_error.c_str(); }

int te_create(te_ctx** out, int device) {
  if (!out) return fail(TE_ERR_BAD_ARG, "out pointer is null");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(TE_ERR_CUDA, "no CUDA device available (%s); libte_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(TE_ERR_BAD_ARG, "device %d out of range [0,%d)", device, n);
  int prev = 0;
  cudaGetDevice(&prev);
  TE_CUDA(cudaSetDevice(device));
  te_ctx* c = new te_ctx();
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sms = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    cudaSetDevice(prev);
    return fail(TE_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
  }
  c->stream = c->own_stream;
  cudaSetDevice(prev);
  *out = c;
  return TE_OK;
}

int te_destroy(te_ctx* c) {
  if (!c) return TE_OK;
  {
    Guard g(c);
    if (g.ok) {
      cudaStreamSynchronize(c->stream);
      c->dX.release();
      c->dY.release();
      for (auto& b : c->stage) b.release();
      c->worklist.release();
      c->worklist3.release();
      c->counter.release();
      c->fused.release();
      c->fp.release();
      for (auto& e : c->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); cudaEventDestroy(e.c); }
      if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
      if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
      if (c->own_stream) cudaStreamDestroy(c->own_stream);
    }
  }
  delete c;
  return TE_OK;
}

int te_host_alloc(void** out, size_t bytes) {
  if (!out) return fail(TE_ERR_BAD_ARG, "out pointer is null");
  *out = nullptr;
  TE_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return TE_OK;
}

int te_host_free(void* p) {
  if (!p) return TE_OK;
  TE_CUDA(cudaFreeHost(p));
  return TE_OK;
}

int te_set_stream(te_ctx* c, void* s) {
  TE_ENTER(c);
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return TE_OK;
}

int te_synchronize(te_ctx* c) {
  TE_ENTER(c);
  TE_CUDA(cudaStreamSynchronize(c->stream));
  return TE_OK;
}

int te_set_kernel(te_ctx* c, int choice) {
  TE_ENTER(c);
  if (choice < TE_KERNEL_AUTO || choice > TE_KERNEL_FUSED) return fail(TE_ERR_BAD_ARG, "unknown kernel choice %d", choice);
  c->kernel_choice = choice;
  return TE_OK;
}

int te_get_stats(te_ctx* c, int64_t* launches, int64_t* slow) {
  TE_ENTER(c);
  if (launches) *launches = c->launches;
  if (slow) {
    // the fix-up counter of the last fused launch (device word 0); cumulative host tally otherwise
    unsigned last[4] = {0, 0, 0, 0};
    if (c->counter.p) {
      TE_CUDA(cudaStreamSynchronize(c->stream));
      TE_CUDA(cudaMemcpy(last, (unsigned*)c->counter.p + 128 * c->counter_phase, sizeof(last), cudaMemcpyDeviceToHost));
    }
    *slow = (int64_t)last[1];  // cells flagged (word 0 counts reserved list entries, chunk padding included)
  }
  return TE_OK;
}

int te_enable_timing(te_ctx* c, int on) {
  TE_ENTER(c);
  c->timing = on != 0;
  return TE_OK;
}

int te_get_timing(te_ctx* c, double* main_ms, double* fixup_ms, int64_t* samples) {
  TE_ENTER(c);
  TE_CUDA(cudaStreamSynchronize(c->stream));
  double m = 0.0, f = 0.0;
  for (size_t k = 0; k < c->events_used; ++k) {
    const auto& e = c->events[k];
    float t1 = 0.f, t2 = 0.f;
    TE_CUDA(cudaEventElapsedTime(&t1, e.a, e.b));
    TE_CUDA(cudaEventElapsedTime(&t2, e.b, e.c));
    m += t1;
    f += t2;
  }
  if (main_ms) *main_ms = m;
  if (fixup_ms) *fixup_ms = f;
  if (samples) *samples = (int64_t)c->events_used;
  c->events_used = 0;
  return TE_OK;
}

int te_fused_plan(int rows, int out_ncols, int nmaps, int sms, int32_t out[19]) {
  if (rows <= 0 || out_ncols <= 0 || nmaps <= 0 || sms <= 0 || !out) return TE_ERR_BAD_ARG;
  int tmp[19];
  te::fused_plan(rows, out_ncols, nmaps, sms, tmp);
  for (int i = 0; i < 19; ++i) out[i] = tmp[i];
  return TE_OK;
}

int te_get_flag_counters(te_ctx* c, uint32_t out[5]) {
  TE_ENTER(c);
  if (!out) return fail(TE_ERR_BAD_ARG, "null argument");
  for (int k = 0; k < 5; ++k) out[k] = 0;
  if (c->counter.p) {
    unsigned raw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    TE_CUDA(cudaStreamSynchronize(c->stream));
    TE_CUDA(cudaMemcpy(raw, (unsigned*)c->counter.p + 128 * c->counter_phase, sizeof(raw), cudaMemcpyDeviceToHost));
    out[0] = raw[1];            // cells flagged by the fp32 stencil
    out[1] = raw[0];            // list entries reserved (warp-private chunks, padding included)
    out[2] = raw[2] | raw[5];   // a work list overflowed (never: the lists hold every cell of the launch)
    out[4] = raw[4];            // cells tier 2 passed on to the literal kernel
  }
  return TE_OK;
}

int te_get_escalation_stats(te_ctx* c, uint32_t reasons[16], uint32_t valid_cells[26]) {
  TE_ENTER(c);
  if (!reasons || !valid_cells) return fail(TE_ERR_BAD_ARG, "null argument");
  unsigned raw[64];
  std::memset(raw, 0, sizeof(raw));
  if (c->counter.p) {
    TE_CUDA(cudaStreamSynchronize(c->stream));
    TE_CUDA(cudaMemcpy(raw, (unsigned*)c->counter.p + 128 * c->counter_phase, sizeof(raw), cudaMemcpyDeviceToHost));
  }
  for (int k = 0; k < 16; ++k) reasons[k] = raw[8 + k];
  for (int k = 0; k < 26; ++k) valid_cells[k] = raw[24 + k];
  return TE_OK;
}

int te_slope(te_ctx* c, const te_geometry* g, double crit, const float* nz, float* out, int memory) {
  TE_ENTER(c);
  if (int rc = check_geometry(g)) return rc;
  if (crit > M_PI_2 || crit < 0.0 || std::isnan(crit)) return fail(TE_ERR_BAD_ARG, "Critical slope must be in the interval [0, PI/2]");
  if (!nz) return fail(TE_ERR_MISSING_LAYER, "layer surface_normal_z is missing");
  if (!out) return fail(TE_ERR_BAD_ARG, "output layer is null");
  const long long n = (long long)g->rows * g->cols;
  const size_t bytes = sizeof(float) * (size_t)n;
  const float* din = nz;
  float* dout = out;
  if (memory == TE_MEM_HOST) {
    TE_CUDA(c->stage[0].reserve(bytes));
    TE_CUDA(c->stage[4].reserve(bytes));
    TE_CUDA(cudaMemcpyAsync(c->stage[0].p, nz, bytes, cudaMemcpyHostToDevice, c->stream));
    din = (const float*)c->stage[0].p;
    dout = (float*)c->stage[4].p;
  }
  te::launch_slope(n, crit, din, dout, c->sms, c->stream);
  if (int rc = launch_check(c, "k_slope")) return rc;
  if (memory == TE_MEM_HOST) {
    TE_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, c->stream));
    TE_CUDA(cudaStreamSynchronize(c->stream));
  }
  return TE_OK;
}

int te_normals(te_ctx* c, const te_geometry* g, const te_chain_params* p, const float* elev, float* nx, float* ny, float* nz, int memory) {
  TE_ENTER(c);
  if (int rc = check_geometry(g)) return rc;
  if (int rc = check_params(p)) return rc;
  if (!elev) return fail(TE_ERR_MISSING_LAYER, "layer elevation is missing");
  if (!nx || !ny || !nz) return fail(TE_ERR_BAD_ARG, "output layer is null");
  if (int rc = ensure_geometry(c, g)) return rc;
  const size_t bytes = sizeof(float) * (size_t)g->rows * g->cols;
  const float* din = elev;
  float* o[3] = {nx, ny, nz};
  if (memory == TE_MEM_HOST) {
    TE_CUDA(c->stage[0].reserve(bytes));
    TE_CUDA(cudaMemcpyAsync(c->stage[0].p, elev, bytes, cudaMemcpyHostToDevice, c->stream));
    din = (const float*)c->stage[0].p;
    for (int k = 0; k < 3; ++k) {
      TE_CUDA(c->stage[4 + k].reserve(bytes));
      o[k] = (float*)c->stage[4 + k].p;
    }
  }
  te_slab s{0, g->cols, 0, 0};
  te::launch_normals(make_view(c, g, s), make_chain_dev(g, p), din, o[0], o[1], o[2], c->sms, c->stream);
  if (int rc = launch_check(c, "k_normals")) return rc;
  if (memory == TE_MEM_HOST) {
    float* h[3] = {nx, ny, nz};
    for (int k = 0; k < 3; ++k) TE_CUDA(cudaMemcpyAsync(h[k], o[k], bytes, cudaMemcpyDeviceToHost, c->stream));
    TE_CUDA(cudaStreamSynchronize(c->stream));
  }
  return TE_OK;
}

int te_step(te_ctx* c, const te_geometry* g, const te_chain_params* p, const float* elev, float* out, int memory) {
  TE_ENTER(c);
  if (int rc = check_geometry(g)) return rc;
  if (int rc = check_params(p)) return rc;
  if (!elev) return fail(TE_ERR_MISSING_LAYER, "layer elevation is missing");
  if (!out) return fail(TE_ERR_BAD_ARG, "output layer is null");
  if (int rc = ensure_geometry(c, g)) return rc;
  const size_t bytes = sizeof(float) * (size_t)g->rows * g->cols;
  const float* din = elev;
  float* dout = out;
  if (memory == TE_MEM_HOST) {
    TE_CUDA(c->stage[0].reserve(bytes));
    TE_CUDA(c->stage[4].reserve(bytes));
    TE_CUDA(cudaMemcpyAsync(c->stage[0].p, elev, bytes, cudaMemcpyHostToDevice, c->stream));
    din = (const float*)c->stage[0].p;
    dout = (float*)c->stage[4].p;
  }
  te_slab s{0, g->cols, 0, 0};
  te::launch_step(make_view(c, g, s), make_chain_dev(g, p), din, dout, c->sms, c->stream);
  if (int rc = launch_check(c, "k_step")) return rc;
  if (memory == TE_MEM_HOST) {
    TE_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, c->stream));
    TE_CUDA(cudaStreamSynchronize(c->stream));
  }
  return TE_OK;
}

int te_roughness(te_ctx* c, const te_geometry* g, const te_chain_params* p, const float* elev, const float* nx, const float* ny,
                 const float* nz, float* out, int memory) {
  TE_ENTER(c);
  if (int rc = check_geometry(g)) return rc;
  if (int rc = check_params(p)) return rc;
  if (!elev) return fail(TE_ERR_MISSING_LAYER, "layer elevation is missing");
  if (!nx || !ny || !nz) return fail(TE_ERR_MISSING_LAYER, "layer surface_normal_{x,y,z} is missing");
  if (!out) return fail(TE_ERR_BAD_ARG, "output layer is null");
  if (int rc = ensure_geometry(c, g)) return rc;
  const size_t bytes = sizeof(float) * (size_t)g->rows * g->cols;
  const float* in[4] = {elev, nx, ny, nz};
  float* dout = out;
  if (memory == TE_MEM_HOST) {
    for (int k = 0; k < 4; ++k) {
      TE_CUDA(c->stage[k].reserve(bytes));
      TE_CUDA(cudaMemcpyAsync(c->stage[k].p, in[k], bytes, cudaMemcpyHostToDevice, c->stream));
      in[k] = (const float*)c->stage[k].p;
    }
    TE_CUDA(c->stage[4].reserve(bytes));
    dout = (float*)c->stage[4].p;
  }
  te_slab s{0, g->cols, 0, 0};
  te::launch_roughness(make_view(c, g, s), make_chain_dev(g, p), in[0], in[1], in[2], in[3], dout, c->sms, c->stream);
  if (int rc = launch_check(c, "k_roughness")) return rc;
  if (memory == TE_MEM_HOST) {
    TE_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, c->stream));
    TE_CUDA(cudaStreamSynchronize(c->stream));
  }
  return TE_OK;
}

// Host-memory chain on a large map: column chunks flow H2D -> kernels -> D2H on three streams so that the two PCIe
// directions and the compute overlap (the transfers dominate: 20 B/cell over PCIe against 20 B/cell over HBM).
// Columns [c0, c0 + n) of the map in DEFAULT order <-> a host layer stored as a grid_map circular buffer with start index
// (sr, sc): cell (i, j) of the map lives at stored[((j + sc) % cols) * rows + (i + sr) % rows] (GridMap::getIndex... /
// convertToDefaultStartIndex).  `dev` holds map column c at dev + (c - dev_col0) * rows.  Up to four 2-D copies;
// one plain copy when nothing wraps.  This is what lets te_chain(TE_MEM_HOST) take the message / GridMap buffers of a moving
// (robot-centric) map as they are, without an unwrapped host copy.
static cudaError_t copy_wrapped(float* dev, int dev_col0, float* host, int rows, int cols, int sr, int sc, int c0, int n, bool to_device,
                                cudaStream_t s) {
  const size_t pitch = sizeof(float) * (size_t)rows;
  int done = 0;
  while (done < n) {
    const int c = c0 + done, js = (c + sc) % cols;     // stored column of map column c
    const int m = std::min(n - done, cols - js);       // columns until the stored index wraps
    float* d = dev + (size_t)(c - dev_col0) * rows;
    float* h = host + (size_t)js * rows;
    // rows [0, rows - sr) of the map are stored rows [sr, rows); rows [rows - sr, rows) are stored rows [0, sr)
    const int r1 = rows - sr;
    cudaError_t e;
    if (to_device) e = cudaMemcpy2DAsync(d, pitch, h + sr, pitch, sizeof(float) * (size_t)r1, m, cudaMemcpyHostToDevice, s);
    else e = cudaMemcpy2DAsync(h + sr, pitch, d, pitch, sizeof(float) * (size_t)r1, m, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    if (sr > 0) {
      if (to_device) e = cudaMemcpy2DAsync(d + r1, pitch, h, pitch, sizeof(float) * (size_t)sr, m, cudaMemcpyHostToDevice, s);
      else e = cudaMemcpy2DAsync(h, pitch, d + r1, pitch, sizeof(float) * (size_t)sr, m, cudaMemcpyDeviceToHost, s);
      if (e != cudaSuccess) return e;
    }
    done += m;
  }
  return cudaSuccess;
}

static int chain_host_pipelined(te_ctx* c, const te_geometry* g, const te_slab& s, const te_chain_params* p, const float* elev,
                                float* const host_out[7], int sr = 0, int sc = 0) {
  const int rows = g->rows, need = chain_halo(g, p);
  const bool wrapped = sr != 0 || sc != 0;  // only with the whole map (no slab): buffer column == map column
  const int in_cols = s.halo_left + s.col_count + s.halo_right;
  const size_t col_bytes = sizeof(float) * (size_t)rows;
  TE_CUDA(c->stage[0].reserve(col_bytes * in_cols));
  float* dev_out[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int k = 0; k < 7; ++k)
    if (host_out[k]) {
      TE_CUDA(c->stage[4 + k].reserve(col_bytes * s.col_count));
      dev_out[k] = (float*)c->stage[4 + k].p;
    }
  if (!c->s_h2d) TE_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  if (!c->s_d2h) TE_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  const int nchunk = std::min(16, std::max(2, s.col_count / 512));
  const int chunk = (s.col_count + nchunk - 1) / nchunk;
  // events and the drain of the three streams are released on every exit path
  struct Pipeline {
    te_ctx* c;
    std::vector<cudaEvent_t> ev;
    cudaError_t make(cudaEvent_t* out) {
      cudaError_t e = cudaEventCreateWithFlags(out, cudaEventDisableTiming);
      if (e == cudaSuccess) ev.push_back(*out);
      return e;
    }
    ~Pipeline() {
      cudaStreamSynchronize(c->s_h2d);
      cudaStreamSynchronize(c->stream);
      cudaStreamSynchronize(c->s_d2h);
      for (cudaEvent_t e : ev) cudaEventDestroy(e);
    }
  } pipe{c, {}};
  std::vector<cudaEvent_t> up(nchunk), done(nchunk);
  for (int k = 0; k < nchunk; ++k) {
    TE_CUDA(pipe.make(&up[k]));
    TE_CUDA(pipe.make(&done[k]));
  }
  cudaEvent_t start;
  TE_CUDA(pipe.make(&start));
  TE_CUDA(cudaEventRecord(start, c->stream));  // order after whatever the caller queued on the context stream
  TE_CUDA(cudaStreamWaitEvent(c->s_h2d, start, 0));
  TE_CUDA(cudaStreamWaitEvent(c->s_d2h, start, 0));
  const float* din = (const float*)c->stage[0].p;
  te::SlabView v = make_view(c, g, s);
  int uploaded = 0;  // input-buffer columns already queued for upload
  int rc = TE_OK;
  for (int k = 0; k < nchunk && rc == TE_OK; ++k) {
    const int off = k * chunk, cnt = std::min(chunk, s.col_count - off);
    if (cnt <= 0) break;
    // chunk k reads input-buffer columns [halo_left + off - need, halo_left + off + cnt + need)
    const int upto = std::min(in_cols, s.halo_left + off + cnt + need);
    if (upto > uploaded) {
      cudaError_t e = wrapped ? copy_wrapped((float*)c->stage[0].p, 0, const_cast<float*>(elev), rows, g->cols, sr, sc, uploaded,
                                             upto - uploaded, true, c->s_h2d)
                              : cudaMemcpyAsync((char*)c->stage[0].p + col_bytes * uploaded, (const char*)elev + col_bytes * uploaded,
                                                col_bytes * (upto - uploaded), cudaMemcpyHostToDevice, c->s_h2d);
      if (e != cudaSuccess) { rc = fail(TE_ERR_CUDA, "H2D chunk copy failed: %s", cudaGetErrorString(e)); break; }
      uploaded = upto;
    }
    cudaEventRecord(up[k], c->s_h2d);
    cudaStreamWaitEvent(c->stream, up[k], 0);
    te::SlabView vk = v;
    vk.out_col0 = s.col_begin + off;
    vk.out_ncols = cnt;
    te::ChainOut ok{dev_out[0] + (size_t)off * rows, dev_out[1] + (size_t)off * rows, dev_out[2] + (size_t)off * rows,
                    dev_out[3] + (size_t)off * rows, dev_out[4] ? dev_out[4] + (size_t)off * rows : nullptr,
                    dev_out[5] ? dev_out[5] + (size_t)off * rows : nullptr, dev_out[6] ? dev_out[6] + (size_t)off * rows : nullptr};
    rc = run_chain_device(c, g, vk, p, din, ok, 1);
    if (rc != TE_OK) break;
    cudaEventRecord(done[k], c->stream);
    cudaStreamWaitEvent(c->s_d2h, done[k], 0);
    for (int l = 0; l < 7; ++l)
      if (host_out[l]) {
        cudaError_t e = wrapped ? copy_wrapped(dev_out[l], 0, host_out[l], rows, g->cols, sr, sc, off, cnt, false, c->s_d2h)
                                : cudaMemcpyAsync(host_out[l] + (size_t)off * rows, dev_out[l] + (size_t)off * rows, col_bytes * cnt,
                                                  cudaMemcpyDeviceToHost, c->s_d2h);
        if (e != cudaSuccess) { rc = fail(TE_ERR_CUDA, "D2H chunk copy failed: %s", cudaGetErrorString(e)); break; }
      }
  }
  cudaStreamSynchronize(c->s_h2d);
  cudaStreamSynchronize(c->stream);
  cudaError_t e = cudaStreamSynchronize(c->s_d2h);
  if (rc == TE_OK && e != cudaSuccess) rc = fail(TE_ERR_CUDA, "pipelined chain failed: %s", cudaGetErrorString(e));
  return rc;
}

static int chain_common(te_ctx* c, const te_geometry* g_in, const te_slab* slab, const te_chain_params* p, int nmaps, const float* elev,
                        float* slope, float* step, float* rough, float* trav, float* nx, float* ny, float* nz, int memory) {
  if (int rc = check_geometry(g_in, true)) return rc;
  // A circular-buffer start index is honoured for whole host maps: the copies to and from the device unwrap / re-wrap the
  // layers, the kernels always see the default order (positions depend on the unwrapped index only).
  const int sr = g_in->start_row, sc = g_in->start_col;
  const bool wrapped = sr != 0 || sc != 0;
  if (wrapped && (memory != TE_MEM_HOST || slab != nullptr || nmaps != 1))
    return fail(TE_ERR_UNSUPPORTED, "circular-buffer start index (%d,%d) != (0,0) is supported for whole maps in host memory only", sr, sc);
  te_geometry g0 = *g_in;
  g0.start_row = g0.start_col = 0;
  const te_geometry* g = &g0;
  if (int rc = check_params(p)) return rc;
  if (nmaps <= 0) return fail(TE_ERR_BAD_ARG, "number of maps must be positive");
  if (!elev) return fail(TE_ERR_MISSING_LAYER, "layer elevation is missing");
  if (!slope || !step || !rough || !trav) return fail(TE_ERR_BAD_ARG, "output layer is null");
  te_slab s;
  if (int rc = resolve_slab(g, slab, chain_halo(g, p), &s)) return rc;
  if (int rc = ensure_geometry(c, g)) return rc;
  const size_t in_bytes = sizeof(float) * (size_t)g->rows * (s.halo_left + s.col_count + s.halo_right) * nmaps;
  const size_t out_bytes = sizeof(float) * (size_t)g->rows * s.col_count * nmaps;
  te::ChainOut o{slope, step, rough, trav, nx, ny, nz};
  const float* din = elev;
  float* host_out[7] = {slope, step, rough, trav, nx, ny, nz};
  if (memory == TE_MEM_HOST && nmaps == 1 && s.col_count >= 1024 && (size_t)g->rows * s.col_count >= ((size_t)1 << 22))
    return chain_host_pipelined(c, g, s, p, elev, host_out, sr, sc);
  if (memory == TE_MEM_HOST) {
    TE_CUDA(c->stage[0].reserve(in_bytes));
    if (wrapped) TE_CUDA(copy_wrapped((float*)c->stage[0].p, 0, const_cast<float*>(elev), g->rows, g->cols, sr, sc, 0, g->cols, true, c->stream));
    else TE_CUDA(cudaMemcpyAsync(c->stage[0].p, elev, in_bytes, cudaMemcpyHostToDevice, c->stream));
    din = (const float*)c->stage[0].p;
    float** dev_out[7] = {&o.slope, &o.step, &o.rough, &o.trav, &o.nx, &o.ny, &o.nz};
    for (int k = 0; k < 7; ++k) {
      if (!host_out[k]) continue;
      TE_CUDA(c->stage[4 + k].reserve(out_bytes));
      *dev_out[k] = (float*)c->stage[4 + k].p;
    }
  }
  if (int rc = run_chain_device(c, g, make_view(c, g, s), p, din, o, nmaps)) return rc;
  if (memory == TE_MEM_HOST) {
    float* dev_out[7] = {o.slope, o.step, o.rough, o.trav, o.nx, o.ny, o.nz};
    for (int k = 0; k < 7; ++k)
      if (host_out[k]) {
        if (wrapped) TE_CUDA(copy_wrapped(dev_out[k], 0, host_out[k], g->rows, g->cols, sr, sc, 0, g->cols, false, c->stream));
        else TE_CUDA(cudaMemcpyAsync(host_out[k], dev_out[k], out_bytes, cudaMemcpyDeviceToHost, c->stream));
      }
    TE_CUDA(cudaStreamSynchronize(c->stream));
  }
  return TE_OK;
}

int te_chain(te_ctx* c, const te_geometry* g, const te_slab* slab, const te_chain_params* p, const float* elev, float* slope,
             float* step, float* rough, float* trav, float* nx, float* ny, float* nz, int memory) {
  TE_ENTER(c);
  return chain_common(c, g, slab, p, 1, elev, slope, step, rough, trav, nx, ny, nz, memory);
}

int te_chain_batched(te_ctx* c, const te_geometry* g, const te_chain_params* p, int32_t nmaps, const float* elev, float* slope,
                     float* step, float* rough, float* trav, int memory) {
  TE_ENTER(c);
  return chain_common(c, g, nullptr, p, nmaps, elev, slope, step, rough, trav, nullptr, nullptr, nullptr, memory);
}

int te_footprint2(te_ctx* c, const te_geometry* g_in, const te_slab* slab, const te_footprint_params* p, const float* trav,
                  const float* slope, const float* step, const float* rough, const float* elev, float* out, float* slope_fp,
                  float* step_fp, float* rough_fp, int memory) {
  TE_ENTER(c);
  if (int rc = check_geometry(g_in, true)) return rc;
  const int sr = g_in->start_row, sc = g_in->start_col;  // circular-buffer maps: whole host maps only, like te_chain
  const bool wrapped = sr != 0 || sc != 0;
  if (wrapped && (memory != TE_MEM_HOST || slab != nullptr))
    return fail(TE_ERR_UNSUPPORTED, "circular-buffer start index (%d,%d) != (0,0) is supported for whole maps in host memory only", sr, sc);
  te_geometry g0 = *g_in;
  g0.start_row = g0.start_col = 0;
  const te_geometry* g = &g0;
  if (!p) return fail(TE_ERR_BAD_ARG, "footprint parameters are null");
  if (!(p->radius >= 0.0) || !(p->offset >= 0.0)) return fail(TE_ERR_BAD_ARG, "footprint radius/offset must be >= 0");
  if (!trav) return fail(TE_ERR_MISSING_LAYER, "layer traversability is missing");
  if (!slope) return fail(TE_ERR_MISSING_LAYER, "layer traversability_slope is missing");
  if (!step) return fail(TE_ERR_MISSING_LAYER, "layer traversability_step is missing");
  if (!elev) return fail(TE_ERR_MISSING_LAYER, "layer elevation is missing");
  if (p->verify_roughness && !rough) return fail(TE_ERR_MISSING_LAYER, "layer traversability_roughness is missing (verify_roughness is set)");
  if (!out) return fail(TE_ERR_BAD_ARG, "output layer is null");
  const bool use_rough = p->verify_roughness != 0;
  te_slab s;
  const int need = te::footprint_halo(g, p);
  if (int rc = resolve_slab(g, slab, need, &s)) return rc;
  if (int rc = ensure_geometry(c, g)) return rc;
  const size_t in_bytes = sizeof(float) * (size_t)g->rows * (s.halo_left + s.col_count + s.halo_right);
  const size_t out_bytes = sizeof(float) * (size_t)g->rows * s.col_count;
  const float* in[5] = {trav, slope, step, elev, use_rough ? rough : nullptr};
  float* o[4] = {out, slope_fp, step_fp, use_rough ? rough_fp : nullptr};
  float* host_o[4] = {out, slope_fp, step_fp, use_rough ? rough_fp : nullptr};
  if (memory == TE_MEM_HOST) {
    // staging: inputs 0..3 (+ roughness in slot 11), outputs 4..7
    const int slot_in[5] = {0, 1, 2, 3, 11};
    for (int k = 0; k < 5; ++k) {
      if (!in[k]) continue;
      TE_CUDA(c->stage[slot_in[k]].reserve(in_bytes));
      if (wrapped) TE_CUDA(copy_wrapped((float*)c->stage[slot_in[k]].p, 0, const_cast<float*>(in[k]), g->rows, g->cols, sr, sc, 0, g->cols, true, c->stream));
      else TE_CUDA(cudaMemcpyAsync(c->stage[slot_in[k]].p, in[k], in_bytes, cudaMemcpyHostToDevice, c->stream));
      in[k] = (const float*)c->stage[slot_in[k]].p;
    }
    for (int k = 0; k < 4; ++k) {
      if (!host_o[k]) continue;
      TE_CUDA(c->stage[4 + k].reserve(out_bytes));
      o[k] = (float*)c->stage[4 + k].p;
    }
  }
  const te::SlabView v = make_view(c, g, s);
  int nl = 0;
  int rc = te::launch_footprint(c->fp, v, g, p, c->hX, c->hY, in[0], in[1], in[2], in[4], in[3], o[0], o[1], o[2], o[3], c->sms,
                                c->stream, &nl);
  if (rc != 0) return fail(rc, "footprint sweep failed: %s", c->fp.why.c_str());
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TE_ERR_CUDA, "footprint launch failed: %s", cudaGetErrorString(e));
  c->launches += nl;
  if (memory == TE_MEM_HOST) {
    for (int k = 0; k < 4; ++k)
      if (host_o[k]) {
        if (wrapped) TE_CUDA(copy_wrapped(o[k], 0, host_o[k], g->rows, g->cols, sr, sc, 0, g->cols, false, c->stream));
        else TE_CUDA(cudaMemcpyAsync(host_o[k], o[k], out_bytes, cudaMemcpyDeviceToHost, c->stream));
      }
    TE_CUDA(cudaStreamSynchronize(c->stream));
  }
  return TE_OK;
}

int te_footprint_polygon(te_ctx* c, const te_geometry* g, const te_slab* slab, const te_footprint_params* p, int32_t npts,
                         const double* pts_xy, double yaw, const float* trav, const float* slope, const float* step, const float* rough,
                         const float* elev, float* out_x, float* out_rot, int memory) {
  TE_ENTER(c);
  if (int rc = check_geometry(g, true)) return rc;
  const int sr = g->start_row, sc = g->start_col;  // circular-buffer maps: whole host maps only, like te_chain
  const bool wrapped = sr != 0 || sc != 0;
  if (wrapped && (memory != TE_MEM_HOST || slab != nullptr))
    return fail(TE_ERR_UNSUPPORTED, "circular-buffer start index (%d,%d) != (0,0) is supported for whole maps in host memory only", sr, sc);
  te_geometry g0 = *g;
  g0.start_row = g0.start_col = 0;
  g = &g0;
  if (!p) return fail(TE_ERR_BAD_ARG, "footprint parameters are null");
  if (npts < 3 || npts > 16 || !pts_xy) return fail(TE_ERR_BAD_ARG, "footprint polygon needs 3 to 16 vertices");
  if (!std::isfinite(yaw)) return fail(TE_ERR_BAD_ARG, "footprint yaw is not finite");
  for (int k = 0; k < 2 * npts; ++k)
    if (!std::isfinite(pts_xy[k])) return fail(TE_ERR_BAD_ARG, "footprint polygon vertex is not finite");
  if (!trav) return fail(TE_ERR_MISSING_LAYER, "layer traversability is missing");
  if (!slope) return fail(TE_ERR_MISSING_LAYER, "layer traversability_slope is missing");
  if (!step) return fail(TE_ERR_MISSING_LAYER, "layer traversability_step is missing");
  if (!elev) return fail(TE_ERR_MISSING_LAYER, "layer elevation is missing");
  if (p->verify_roughness && !rough) return fail(TE_ERR_MISSING_LAYER, "layer traversability_roughness is missing (verify_roughness is set)");
  if (!out_x || !out_rot) return fail(TE_ERR_BAD_ARG, "output layer is null");
  const bool use_rough = p->verify_roughness != 0;
  te_slab s;
  if (int rc = resolve_slab(g, slab, te::footprint_polygon_halo(g, p, npts, pts_xy), &s)) return rc;
  if (int rc = ensure_geometry(c, g)) return rc;
  const size_t in_bytes = sizeof(float) * (size_t)g->rows * (s.halo_left + s.col_count + s.halo_right);
  const size_t out_bytes = sizeof(float) * (size_t)g->rows * s.col_count;
  const float* in[5] = {trav, slope, step, elev, use_rough ? rough : nullptr};
  float* o[2] = {out_x, out_rot};
  if (memory == TE_MEM_HOST) {
    const int slot_in[5] = {0, 1, 2, 3, 11};
    for (int k = 0; k < 5; ++k) {
      if (!in[k]) continue;
      TE_CUDA(c->stage[slot_in[k]].reserve(in_bytes));
      if (wrapped) TE_CUDA(copy_wrapped((float*)c->stage[slot_in[k]].p, 0, const_cast<float*>(in[k]), g->rows, g->cols, sr, sc, 0, g->cols, true, c->stream));
      else TE_CUDA(cudaMemcpyAsync(c->stage[slot_in[k]].p, in[k], in_bytes, cudaMemcpyHostToDevice, c->stream));
      in[k] = (const float*)c->stage[slot_in[k]].p;
    }
    for (int k = 0; k < 2; ++k) {
      TE_CUDA(c->stage[4 + k].reserve(out_bytes));
      o[k] = (float*)c->stage[4 + k].p;
    }
  }
  const te::SlabView v = make_view(c, g, s);
  int nl = 0;
  int rc = te::launch_footprint_polygon(c->fp, v, g, p, npts, pts_xy, yaw, in[0], in[1], in[2], in[4], in[3], o[0], o[1], c->sms, c->stream, &nl);
  if (rc != 0) return fail(rc, "polygon footprint sweep failed: %s", c->fp.why.c_str());
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TE_ERR_CUDA, "polygon footprint launch failed: %s", cudaGetErrorString(e));
  c->launches += nl;
  if (memory == TE_MEM_HOST) {
    if (wrapped) {
      TE_CUDA(copy_wrapped(o[0], 0, out_x, g->rows, g->cols, sr, sc, 0, g->cols, false, c->stream));
      TE_CUDA(copy_wrapped(o[1], 0, out_rot, g->rows, g->cols, sr, sc, 0, g->cols, false, c->stream));
    } else {
      TE_CUDA(cudaMemcpyAsync(out_x, o[0], out_bytes, cudaMemcpyDeviceToHost, c->stream));
      TE_CUDA(cudaMemcpyAsync(out_rot, o[1], out_bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    TE_CUDA(cudaStreamSynchronize(c->stream));
  }
  return TE_OK;
}

int te_footprint(te_ctx* c, const te_geometry* g, const te_slab* slab, const te_footprint_params* p, const float* trav,
                 const float* slope, const float* step, const float* elev, float* out, float* slope_fp, float* step_fp, int memory) {
  if (p && p->verify_roughness) return fail(TE_ERR_MISSING_LAYER, "verify_roughness is set: call te_footprint2 with the traversability_roughness layer");
  return te_footprint2(c, g, slab, p, trav, slope, step, nullptr, elev, out, slope_fp, step_fp, nullptr, memory);
}

int te_check_footprint_paths(te_ctx* c, const te_geometry* g, const float* footprint, double traversability_default, int32_t npaths,
                             const int32_t* path_begin, const double* poses_xy, uint8_t* is_safe, double* traversability, int memory) {
  return te_check_footprint_paths2(c, g, footprint, nullptr, traversability_default, npaths, path_begin, poses_xy, is_safe, traversability, memory);
}

int te_check_footprint_paths2(te_ctx* c, const te_geometry* g, const float* footprint, const float* robot_slope, double traversability_default,
                              int32_t npaths, const int32_t* path_begin, const double* poses_xy, uint8_t* is_safe, double* traversability,
                              int memory) {
  TE_ENTER(c);
  if (int rc = check_geometry(g)) return rc;
  if (!footprint) return fail(TE_ERR_MISSING_LAYER, "layer traversability_footprint is missing");
  if (npaths < 0 || !path_begin || !poses_xy || !is_safe || !traversability) return fail(TE_ERR_BAD_ARG, "null argument or negative path count");
  if (npaths == 0) return TE_OK;
  if (int rc = ensure_geometry(c, g)) return rc;
  const te_slab s{0, g->cols, 0, 0};
  const te::SlabView v = make_view(c, g, s);
  if (memory == TE_MEM_DEVICE) {
    te::launch_check_paths(v, g, traversability_default, footprint, robot_slope, npaths, path_begin, poses_xy, is_safe, traversability, c->stream);
    return launch_check(c, "k_check_paths");
  }
  // host arguments: path_begin is readable here, so the pose count is known
  const int32_t nposes = path_begin[npaths];
  if (nposes < 0) return fail(TE_ERR_BAD_ARG, "path_begin must be non-decreasing");
  const size_t lbytes = sizeof(float) * (size_t)g->rows * g->cols;
  TE_CUDA(c->stage[0].reserve(lbytes));
  TE_CUDA(c->stage[1].reserve(sizeof(int32_t) * (size_t)(npaths + 1)));
  TE_CUDA(c->stage[2].reserve(sizeof(double) * 2 * (size_t)std::max(nposes, 1)));
  TE_CUDA(c->stage[4].reserve((size_t)npaths));
  TE_CUDA(c->stage[5].reserve(sizeof(double) * (size_t)npaths));
  TE_CUDA(cudaMemcpyAsync(c->stage[0].p, footprint, lbytes, cudaMemcpyHostToDevice, c->stream));
  if (robot_slope) {
    TE_CUDA(c->stage[3].reserve(lbytes));
    TE_CUDA(cudaMemcpyAsync(c->stage[3].p, robot_slope, lbytes, cudaMemcpyHostToDevice, c->stream));
  }
  TE_CUDA(cudaMemcpyAsync(c->stage[1].p, path_begin, sizeof(int32_t) * (size_t)(npaths + 1), cudaMemcpyHostToDevice, c->stream));
  TE_CUDA(cudaMemcpyAsync(c->stage[2].p, poses_xy, sizeof(double) * 2 * (size_t)nposes, cudaMemcpyHostToDevice, c->stream));
  te::launch_check_paths(v, g, traversability_default, (const float*)c->stage[0].p, robot_slope ? (const float*)c->stage[3].p : nullptr, npaths, (const int*)c->stage[1].p,
                         (const double*)c->stage[2].p, (unsigned char*)c->stage[4].p, (double*)c->stage[5].p, c->stream);
  if (int rc = launch_check(c, "k_check_paths")) return rc;
  TE_CUDA(cudaMemcpyAsync(is_safe, c->stage[4].p, (size_t)npaths, cudaMemcpyDeviceToHost, c->stream));
  TE_CUDA(cudaMemcpyAsync(traversability, c->stage[5].p, sizeof(double) * (size_t)npaths, cudaMemcpyDeviceToHost, c->stream));
  TE_CUDA(cudaStreamSynchronize(c->stream));
  return TE_OK;
}

// A te IPC handle is the CUDA handle of the ALLOCATION that contains the pointer (cudaIpcGetMemHandle always describes the whole
// allocation; sub-allocating pools such as torch's caching allocator hand out interior pointers) plus the offset into it.
struct TeIpcHandle {
  cudaIpcMemHandle_t mem;
  uint64_t offset;
  uint64_t reserved;
};
static_assert(sizeof(TeIpcHandle) == TE_IPC_HANDLE_BYTES, "te IPC handle layout");

namespace {
std::mutex g_ipc_mu;
std::map<void*, void*> g_ipc_opened;  // pointer handed out by te_ipc_open -> base of the mapping to close
}  // namespace

int te_ipc_export(const void* device_ptr, void* handle) {
  if (!device_ptr || !handle) return fail(TE_ERR_BAD_ARG, "null argument");
  typedef CUresult (*PFN_range)(CUdeviceptr*, size_t*, CUdeviceptr);
  static PFN_range range = nullptr;
  if (!range) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return fail(TE_ERR_CUDA, "cuMemGetAddressRange entry point unavailable");
    range = reinterpret_cast<PFN_range>(p);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (range(&base, &size, (CUdeviceptr)device_ptr) != CUDA_SUCCESS) return fail(TE_ERR_CUDA, "cuMemGetAddressRange failed (not a device allocation?)");
  TeIpcHandle h{};
  TE_CUDA(cudaIpcGetMemHandle(&h.mem, reinterpret_cast<void*>(base)));
  h.offset = (uint64_t)((CUdeviceptr)device_ptr - base);
  std::memcpy(handle, &h, sizeof(h));
  return TE_OK;
}

int te_ipc_open(const void* handle, void** out) {
  if (!handle || !out) return fail(TE_ERR_BAD_ARG, "null argument");
  TeIpcHandle h;
  std::memcpy(&h, handle, sizeof(h));
  void* base = nullptr;
  TE_CUDA(cudaIpcOpenMemHandle(&base, h.mem, cudaIpcMemLazyEnablePeerAccess));
  *out = (char*)base + h.offset;
  std::lock_guard<std::mutex> lock(g_ipc_mu);
  g_ipc_opened[*out] = base;
  return TE_OK;
}

int te_ipc_close(void* p) {
  if (!p) return TE_OK;
  void* base = p;
  {
    std::lock_guard<std::mutex> lock(g_ipc_mu);
    auto it = g_ipc_opened.find(p);
    if (it != g_ipc_opened.end()) { base = it->second; g_ipc_opened.erase(it); }
  }
  TE_CUDA(cudaIpcCloseMemHandle(base));
  return TE_OK;
}

int te_event_create_ipc(te_ctx* c, void** event_out, void* handle) {
  TE_ENTER(c);
  if (!event_out || !handle) return fail(TE_ERR_BAD_ARG, "null argument");
  cudaEvent_t ev;
  TE_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventInterprocess));
  cudaIpcEventHandle_t h;
  cudaError_t e = cudaIpcGetEventHandle(&h, ev);
  if (e != cudaSuccess) {
    cudaEventDestroy(ev);
    return fail(TE_ERR_CUDA, "cudaIpcGetEventHandle failed: %s", cudaGetErrorString(e));
  }
  static_assert(sizeof(h) == 64, "cudaIpcEventHandle_t is 64 bytes");
  std::memcpy(handle, &h, sizeof(h));
  *event_out = (void*)ev;
  return TE_OK;
}

int te_event_open_ipc(const void* handle, void** event_out) {
  if (!handle || !event_out) return fail(TE_ERR_BAD_ARG, "null argument");
  cudaIpcEventHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  cudaEvent_t ev;
  TE_CUDA(cudaIpcOpenEventHandle(&ev, h));
  *event_out = (void*)ev;
  return TE_OK;
}

int te_event_record(te_ctx* c, void* event) {
  TE_ENTER(c);
  if (!event) return fail(TE_ERR_BAD_ARG, "event is null");
  TE_CUDA(cudaEventRecord((cudaEvent_t)event, c->stream));
  return TE_OK;
}

int te_event_destroy(void* event) {
  if (!event) return TE_OK;
  TE_CUDA(cudaEventDestroy((cudaEvent_t)event));
  return TE_OK;
}

int te_halo_pull(te_ctx* c, const te_geometry* g, const te_slab* slab, float* layer, const te_halo_peer* left,
                 const te_halo_peer* right) {
  TE_ENTER(c);
  if (int rc = check_geometry(g)) return rc;
  if (!slab || !layer) return fail(TE_ERR_BAD_ARG, "null argument");
  te_slab s;
  if (int rc = resolve_slab(g, slab, 0, &s)) return rc;
  const size_t col_bytes = sizeof(float) * (size_t)g->rows;
  // Global columns [first, first + count) into this rank's buffer, from the OWNED columns of `p`.
  auto pull = [&](const te_halo_peer* p, int first, int count, const char* side) -> int {
    if (count <= 0) return TE_OK;
    if (!p || !p->layer) return fail(TE_ERR_BAD_ARG, "slab has a %s halo of %d columns but no %s neighbour was given", side, count, side);
    const te_slab& q = p->slab;
    if (q.col_begin < 0 || q.col_count <= 0 || q.halo_left < 0 || q.halo_right < 0 || q.col_begin + q.col_count > g->cols)
      return fail(TE_ERR_BAD_ARG, "%s neighbour's slab is malformed", side);
    if (first < q.col_begin || first + count > q.col_begin + q.col_count)
      return fail(TE_ERR_BAD_ARG, "%s halo columns [%d,%d) are not owned by the %s neighbour [%d,%d): the exchange is one hop", side,
                  first, first + count, side, q.col_begin, q.col_begin + q.col_count);
    if (p->ready_event) TE_CUDA(cudaStreamWaitEvent(c->stream, (cudaEvent_t)p->ready_event, 0));
    const char* src = (const char*)p->layer + col_bytes * (size_t)(first - (q.col_begin - q.halo_left));
    char* dst = (char*)layer + col_bytes * (size_t)(first - (s.col_begin - s.halo_left));
    TE_CUDA(cudaMemcpyAsync(dst, src, col_bytes * (size_t)count, cudaMemcpyDefault, c->stream));
    return TE_OK;
  };
  if (int rc = pull(left, s.col_begin - s.halo_left, s.halo_left, "left")) return rc;
  if (int rc = pull(right, s.col_begin + s.col_count, s.halo_right, "right")) return rc;
  return TE_OK;
}

}  // extern "C"
