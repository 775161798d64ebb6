/*
 * te_oracle.cpp — CPU ORACLE of the filter chain (test infrastructure, NOT product code).
 * See te_oracle.h for what this is and is not.  Build: `make -C oracle` (g++ -O2 -ffp-contract=off).
 *
 * Every function cites the reference lines (relative to /root/reference) whose arithmetic it
 * restates.  Where the arithmetic lives in un-vendored grid_map the
 * citation is the reference's call site plus the appendix paragraph.
 *
 * Floating-point discipline: all position / membership / moment arithmetic is IEEE double with
 * no contraction (-ffp-contract=off) and no reassociation, in the operand order written here;
 * every layer is narrowed to float32 on store exactly where the reference stores into a
 * grid_map::Matrix (Eigen::MatrixXf).
 */
#include "te_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr float kNaNf = std::numeric_limits<float>::quiet_NaN();

struct Geo {
  int rows, cols;
  double res, lenx, leny, posx, posy;
  std::vector<double> X, Y;  // cell-centre coordinates per row index / column index
};

// grid_map::getPositionFromIndex: p = mapPosition + (0.5*length - 0.5*resolution)
// + resolution * (-index), evaluated per axis as (mapPosition + offset) + resolution*(-idx).
// startIndex is (0,0) for every map this oracle sees (the shell normalises circular buffers).
inline double cell_coord(double map_pos, double length, double res, int idx) {
  const double offset = 0.5 * length - 0.5 * res;
  return (map_pos + offset) + res * (-(double)idx);
}

Geo make_geo(const teo_geometry* g) {
  Geo o{g->rows, g->cols, g->resolution, g->length_x, g->length_y, g->position_x, g->position_y, {}, {}};
  o.X.resize(o.rows);
  o.Y.resize(o.cols);
  for (int i = 0; i < o.rows; ++i) o.X[i] = cell_coord(o.posx, o.lenx, o.res, i);
  for (int j = 0; j < o.cols; ++j) o.Y[j] = cell_coord(o.posy, o.leny, o.res, j);
  return o;
}

inline bool valid(float v) { return std::isfinite(v); }  // GridMap::isValid == std::isfinite (A.1)

// grid_map::CircleIterator: cells of the clipped bounding box, row index outer and
// column index inner, that satisfy ||p(a,b) - p(i,j)||^2 <= r^2 with both sides in double.
// The bounding box never excludes a cell that passes the test (it covers whole cells around
// centre +/- radius), so a generous integer box is equivalent.
template <class F>
inline void for_circle(const Geo& g, int i, int j, double radius, F&& f) {
  const int R = (int)std::floor(radius / g.res) + 1;
  const double r2 = radius * radius;  // pow(radius, 2)
  const int a0 = std::max(0, i - R), a1 = std::min(g.rows - 1, i + R);
  const int b0 = std::max(0, j - R), b1 = std::min(g.cols - 1, j + R);
  const double cx = g.X[i], cy = g.Y[j];
  for (int a = a0; a <= a1; ++a) {
    const double dx = g.X[a] - cx;
    for (int b = b0; b <= b1; ++b) {
      const double dy = g.Y[b] - cy;
      if (dx * dx + dy * dy <= r2) f(a, b);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Small dense linear algebra in double (stand-ins for the Eigen routines the reference's
// dependency calls; Eigen is not installed here).
// ---------------------------------------------------------------------------------------------

// Rank as Eigen::FullPivHouseholderQR<Matrix3d>::rank() reports it with the default threshold
// (epsilon * diagonalSize relative to the largest pivot).  Appendix A.4b.
int full_piv_householder_rank3(const double cov[3][3]) {
  double m[3][3];
  std::memcpy(m, cov, sizeof(m));
  const double eps = std::numeric_limits<double>::epsilon();
  const double precision = eps * 3.0;
  double biggest = 0.0, maxpivot = 0.0;
  int nonzero_pivots = 3;
  double diag[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    int pr = k, pc = k;
    double big = -1.0;
    // Eigen's maxCoeff visitor walks a column-major block: column outer, row inner, strict '>'.
    for (int c = k; c < 3; ++c)
      for (int r = k; r < 3; ++r) {
        const double v = std::fabs(m[r][c]);
        if (v > big) { big = v; pr = r; pc = c; }
      }
    if (k == 0) biggest = big;
    if (std::fabs(big) <= std::fabs(biggest) * precision) {  // isMuchSmallerThan
      nonzero_pivots = k;
      break;
    }
    if (pr != k) for (int c = 0; c < 3; ++c) std::swap(m[k][c], m[pr][c]);
    if (pc != k) for (int r = 0; r < 3; ++r) std::swap(m[r][k], m[r][pc]);
    // makeHouseholderInPlace on column k, rows k..2
    double tail2 = 0.0;
    for (int r = k + 1; r < 3; ++r) tail2 += m[r][k] * m[r][k];
    const double c0 = m[k][k];
    double beta, tau;
    double ess[3] = {0, 0, 0};
    if (tail2 <= std::numeric_limits<double>::min()) {
      tau = 0.0;
      beta = c0;
    } else {
      beta = std::sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0) beta = -beta;
      for (int r = k + 1; r < 3; ++r) ess[r] = m[r][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    m[k][k] = beta;
    diag[k] = beta;
    if (std::fabs(beta) > maxpivot) maxpivot = std::fabs(beta);
    // applyHouseholderOnTheLeft to the trailing columns
    for (int c = k + 1; c < 3; ++c) {
      double tmp = m[k][c];
      for (int r = k + 1; r < 3; ++r) tmp += ess[r] * m[r][c];
      m[k][c] -= tau * tmp;
      for (int r = k + 1; r < 3; ++r) m[r][c] -= tau * ess[r] * tmp;
    }
  }
  const double threshold = maxpivot * (eps * 3.0);
  int rank = 0;
  for (int k = 0; k < nonzero_pivots; ++k)
    if (std::fabs(diag[k]) > threshold) ++rank;
  return rank;
}

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3 (double).  evec columns are unit vectors.
void jacobi_eig3(const double a_in[3][3], double eval[3], double evec[3][3]) {
  double a[3][3];
  std::memcpy(a, a_in, sizeof(a));
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) evec[r][c] = (r == c) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double g = 100.0 * std::fabs(apq);
        if (sweep > 3 && std::fabs(a[p][p]) + g == std::fabs(a[p][p]) &&
            std::fabs(a[q][q]) + g == std::fabs(a[q][q])) {
          a[p][q] = a[q][p] = 0.0;
          continue;
        }
        const double h = a[q][q] - a[p][p];
        double t;
        if (std::fabs(h) + g == std::fabs(h)) {
          t = apq / h;
        } else {
          const double theta = 0.5 * h / apq;
          t = 1.0 / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
          if (theta < 0.0) t = -t;
        }
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        // A <- J^T A J
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        a[p][q] = a[q][p] = 0.0;
        for (int k = 0; k < 3; ++k) {
          const double vkp = evec[k][p], vkq = evec[k][q];
          evec[k][p] = c * vkp - s * vkq;
          evec[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int k = 0; k < 3; ++k) eval[k] = a[k][k];
}

// ---------------------------------------------------------------------------------------------
// a1: grid_map::NormalVectorsFilter, area method.  Third-party; call site
// traversability_estimation/config/robot_filter_parameter.yaml:3-9; consumers SlopeFilter.cpp:71,74 and
// RoughnessFilter.cpp:84,108-110.  Semantics: SURVEY.md Appendix A.4 (a)/(b).
// Decisions not pinned by any reference data (oracle/README.md): cells with invalid elevation get
// no normal (the consumers' "empty cell" guards, SlopeFilter.cpp:70-71, RoughnessFilter.cpp:83-84).
// ---------------------------------------------------------------------------------------------
void normal_at(const Geo& g, const teo_chain_params& p, const float* elev, int i, int j, double n_out[3]) {
  double pts[3][64];
  std::vector<double> big;  // only for windows > 64 cells
  double* px = pts[0];
  double* py = pts[1];
  double* pz = pts[2];
  const int R = (int)std::floor(p.normals_radius / g.res) + 1;
  const int cap = (2 * R + 1) * (2 * R + 1);
  if (cap > 64) {
    big.resize(3 * (size_t)cap);
    px = big.data();
    py = px + cap;
    pz = py + cap;
  }
  int n = 0;
  for_circle(g, i, j, p.normals_radius, [&](int a, int b) {
    const float z = elev[(size_t)b * g.rows + a];
    if (!valid(z)) return;
    px[n] = g.X[a];  // getPosition3: (x, y, (double)value)
    py[n] = g.Y[b];
    pz[n] = (double)z;
    ++n;
  });
  double nrm[3] = {0.0, 0.0, 1.0};
  if (p.normals_algorithm == TEO_NORMALS_FIXTURE) {
    // mean = points.rowwise().sum() / nPoints; NN = points.colwise() - mean; cov = NN * NN^T
    double sx = 0, sy = 0, sz = 0;
    for (int k = 0; k < n; ++k) { sx += px[k]; sy += py[k]; sz += pz[k]; }
    const double mx = sx / (double)n, my = sy / (double)n, mz = sz / (double)n;
    double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < n; ++k) {
      const double d[3] = {px[k] - mx, py[k] - my, pz[k] - mz};
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) cov[r][c] += d[r] * d[c];
    }
    if (full_piv_householder_rank3(cov) >= 3) {
      double eval[3], evec[3][3];
      jacobi_eig3(cov, eval, evec);
      int s = 0;
      double sv = std::numeric_limits<double>::max();
      for (int k = 0; k < 3; ++k)
        if (eval[k] < sv) { sv = eval[k]; s = k; }
      nrm[0] = evec[0][s]; nrm[1] = evec[1][s]; nrm[2] = evec[2][s];
    }  // else: eigenvalues (1,1,0), eigenvectors identity -> (0,0,1)
  } else {
    if (n >= 3) {
      double s[3] = {0, 0, 0}, ss[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (int k = 0; k < n; ++k) {
        const double d[3] = {px[k], py[k], pz[k]};
        for (int r = 0; r < 3; ++r) {
          s[r] += d[r];
          for (int c = 0; c < 3; ++c) ss[r][c] += d[r] * d[c];
        }
      }
      double cov[3][3];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) cov[r][c] = ss[r][c] / (double)n - (s[r] / (double)n) * (s[c] / (double)n);
      double eval[3], evec[3][3];
      jacobi_eig3(cov, eval, evec);
      int order[3] = {0, 1, 2};
      std::sort(order, order + 3, [&](int x, int y) { return eval[x] < eval[y]; });
      if (eval[order[1]] > 1e-8) {
        nrm[0] = evec[0][order[0]]; nrm[1] = evec[1][order[0]]; nrm[2] = evec[2][order[0]];
      }
    }
  }
  const int ax = p.normals_positive_axis;
  if (nrm[ax] < 0.0) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; nrm[2] = -nrm[2]; }
  n_out[0] = nrm[0]; n_out[1] = nrm[1]; n_out[2] = nrm[2];
}

// a2: SlopeFilter<T>::update, traversability_estimation_filters/src/SlopeFilter.cpp:59-89.
inline float slope_value(float nz, double crit) {
  if (!valid(nz)) return kNaNf;                       // :71  no surface normal -> layer stays NaN
  const double slope = std::acos((double)nz);         // :74  double acos of the float32 layer value
  return (float)(slope < crit ? 1.0 - slope / crit : 0.0);  // :76-81
}

// a3 pass 1: StepFilter.cpp:112-144.
inline float step_height_at(const Geo& g, const float* elev, double r1, int i, int j) {
  if (!valid(elev[(size_t)j * g.rows + i])) return kNaNf;  // :113
  bool init = false;
  double hmax = 0, hmin = 0;
  for_circle(g, i, j, r1, [&](int a, int b) {
    const float z = elev[(size_t)b * g.rows + a];
    if (!valid(z)) return;                                 // :126
    const double h = (double)z;
    if (!init) { hmax = hmin = h; init = true; return; }   // :130-135
    if (h > hmax) hmax = h;
    if (h < hmin) hmin = h;
  });
  return init ? (float)(hmax - hmin) : kNaNf;              // :142-143  stored into a float layer
}

// a3 pass 2: StepFilter.cpp:147-178.
inline float step_value(const Geo& g, const teo_chain_params& p, const float* sh, int i, int j) {
  int nCells = 0;
  double stepMax = 0.0;
  bool isValid = false;
  for_circle(g, i, j, p.step_second_radius, [&](int a, int b) {
    const float v = sh[(size_t)b * g.rows + a];
    if (!valid(v)) return;                                 // :159
    isValid = true;
    if ((double)v > stepMax) stepMax = (double)v;          // :162-164
    if ((double)v > p.step_critical) ++nCells;             // :165-166
  });
  if (!isValid) return kNaNf;
  const double step = std::min(stepMax, (double)nCells / (double)p.step_critical_cells * stepMax);  // :170-171
  return (float)(step < p.step_critical ? 1.0 - step / p.step_critical : 0.0);                       // :172-176
}

// a4: RoughnessFilter<T>::update, traversability_estimation_filters/src/RoughnessFilter.cpp:73-132.
inline float roughness_value(const Geo& g, const teo_chain_params& p, const float* elev, const float* nx,
                             const float* ny, const float* nz, int i, int j) {
  const size_t c = (size_t)j * g.rows + i;
  if (!valid(nx[c])) return kNaNf;  // :84
  double sx = 0, sy = 0, sz = 0;
  size_t n = 0;
  for_circle(g, i, j, p.roughness_radius, [&](int a, int b) {
    const float z = elev[(size_t)b * g.rows + a];
    if (!valid(z)) return;
    sx += g.X[a]; sy += g.Y[b]; sz += (double)z;  // rowwise().sum() (:105)
    ++n;
  });
  const double mx = sx / (double)n, my = sy / (double)n, mz = sz / (double)n;
  const double normalX = nx[c], normalY = ny[c], normalZ = nz[c];           // :108-110 float layers
  const double plane = mx * normalX + my * normalY + mz * normalZ;          // :111
  double sum = 0.0;
  for_circle(g, i, j, p.roughness_radius, [&](int a, int b) {
    const float z = elev[(size_t)b * g.rows + a];
    if (!valid(z)) return;
    const double dist = normalX * g.X[a] + normalY * g.Y[b] + normalZ * (double)z - plane;  // :113
    sum += dist * dist;                                                                     // :114 pow(dist,2)
  });
  // :117  nPoints is size_t: n == 1 -> 0/0 = NaN -> comparison false -> 0.0
  const double roughness = std::sqrt(sum / (double)(n - 1));
  return (float)(roughness < p.roughness_critical ? 1.0 - roughness / p.roughness_critical : 0.0);  // :119-124
}

int resolve_threads(int nthreads) {
#ifdef _OPENMP
  return nthreads > 0 ? nthreads : omp_get_max_threads();
#else
  (void)nthreads;
  return 1;
#endif
}

bool bad_geo(const teo_geometry* g) {
  return !g || g->rows <= 0 || g->cols <= 0 || !(g->resolution > 0.0);
}

}  // namespace

extern "C" {

int teo_max_threads(void) { return resolve_threads(0); }

int teo_normals(const teo_geometry* gg, const teo_chain_params* p, const float* elev, float* nx, float* ny,
                float* nz, int nthreads) {
  if (bad_geo(gg) || !p || !elev || !nx || !ny || !nz) return 1;
  const Geo g = make_geo(gg);
  const int nt = resolve_threads(nthreads);
  (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int j = 0; j < g.cols; ++j)
    for (int i = 0; i < g.rows; ++i) {
      const size_t c = (size_t)j * g.rows + i;
      if (!valid(elev[c])) { nx[c] = ny[c] = nz[c] = kNaNf; continue; }
      double n[3];
      normal_at(g, *p, elev, i, j, n);
      nx[c] = (float)n[0]; ny[c] = (float)n[1]; nz[c] = (float)n[2];
    }
  return 0;
}

int teo_slope(const teo_geometry* gg, double critical, const float* nz, float* out, int nthreads) {
  if (bad_geo(gg) || !nz || !out) return 1;
  const int64_t n = (int64_t)gg->rows * gg->cols;
  const int nt = resolve_threads(nthreads);
  (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int64_t c = 0; c < n; ++c) out[c] = slope_value(nz[c], critical);
  return 0;
}

int teo_step(const teo_geometry* gg, const teo_chain_params* p, const float* elev, float* out, float* sh_out,
             int nthreads) {
  if (bad_geo(gg) || !p || !elev || !out) return 1;
  const Geo g = make_geo(gg);
  const int nt = resolve_threads(nthreads);
  (void)nt;
  std::vector<float> tmp;
  float* sh = sh_out;
  if (!sh) { tmp.resize((size_t)g.rows * g.cols); sh = tmp.data(); }
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int j = 0; j < g.cols; ++j)
    for (int i = 0; i < g.rows; ++i) sh[(size_t)j * g.rows + i] = step_height_at(g, elev, p->step_first_radius, i, j);
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int j = 0; j < g.cols; ++j)
    for (int i = 0; i < g.rows; ++i) out[(size_t)j * g.rows + i] = step_value(g, *p, sh, i, j);
  return 0;
}

int teo_roughness(const teo_geometry* gg, const teo_chain_params* p, const float* elev, const float* nx,
                  const float* ny, const float* nz, float* out, int nthreads) {
  if (bad_geo(gg) || !p || !elev || !nx || !ny || !nz || !out) return 1;
  const Geo g = make_geo(gg);
  const int nt = resolve_threads(nthreads);
  (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int j = 0; j < g.cols; ++j)
    for (int i = 0; i < g.rows; ++i) out[(size_t)j * g.rows + i] = roughness_value(g, *p, elev, nx, ny, nz, i, j);
  return 0;
}

// a5: grid_map::MathExpressionFilter with the YAML expression
// "(1.0 / 3.0) * (traversability_slope + traversability_step + traversability_roughness)"
// (robot_filter_parameter.yaml:29-33): EigenLab evaluates on MatrixXf, i.e. in float32, left to right.
int teo_fuse(int64_t n, float weight, const float* slope, const float* step, const float* rough, float* out) {
  if (n < 0 || !slope || !step || !rough || !out) return 1;
  for (int64_t c = 0; c < n; ++c) {
    const float st = slope[c] + step[c];
    const float sum = st + rough[c];
    out[c] = weight * sum;
  }
  return 0;
}

// a1..a6 in YAML order (filters::FilterChain<grid_map::GridMap>::update, TraversabilityMap.cpp:214).
int teo_chain(const teo_geometry* gg, const teo_chain_params* p, const float* elev, float* slope, float* step,
              float* rough, float* trav, float* nx_o, float* ny_o, float* nz_o, int nthreads) {
  if (bad_geo(gg) || !p || !elev || !slope || !step || !rough || !trav) return 1;
  const size_t n = (size_t)gg->rows * gg->cols;
  std::vector<float> bx, by, bz;
  float* nx = nx_o; float* ny = ny_o; float* nz = nz_o;
  if (!nx) { bx.resize(n); nx = bx.data(); }
  if (!ny) { by.resize(n); ny = by.data(); }
  if (!nz) { bz.resize(n); nz = bz.data(); }
  int rc = teo_normals(gg, p, elev, nx, ny, nz, nthreads);
  if (!rc) rc = teo_slope(gg, p->slope_critical, nz, slope, nthreads);
  if (!rc) rc = teo_step(gg, p, elev, step, nullptr, nthreads);
  if (!rc) rc = teo_roughness(gg, p, elev, nx, ny, nz, rough, nthreads);
  if (!rc) rc = teo_fuse((int64_t)n, p->fuse_weight, slope, step, rough, trav);
  return rc;
}

int teo_circle_cells(const teo_geometry* gg, int i, int j, double radius, int32_t* a_out, int32_t* b_out, int cap) {
  if (bad_geo(gg) || i < 0 || j < 0 || i >= gg->rows || j >= gg->cols) return -1;
  const Geo g = make_geo(gg);
  int n = 0;
  for_circle(g, i, j, radius, [&](int a, int b) {
    if (n < cap) { a_out[n] = a; b_out[n] = b; }
    ++n;
  });
  return n;
}

}  // extern "C"
