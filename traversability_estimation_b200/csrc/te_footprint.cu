// te_footprint.cu — TraversabilityMap::traversabilityFootprint(radius, offset) as two kernels.
//
//   k_predicates   isTraversableForFilters (TraversabilityMap.cpp:774-792) for every cell of the slab
//                  and its halo: checkForSlope (:867-893) and checkForStep (:794-865).  Both are pure
//                  functions of the layers (their slope_footprint / step_footprint layers are only
//                  memoisation), so they are evaluated once per cell instead of once per visit.  The
//                  gap walk of checkForStep decides on absolute double positions (submap geometry,
//                  dot products of perpendicular vectors, `norm < max_gap_width`), so this translation
//                  unit is compiled with --fmad=false and replays the reference's operand order.
//   k_sweep        isTraversable(center, radius + offset, ..., radius) (:654-746) for every cell:
//                  walk the SpiralIterator visit order (table built on the host by executing the
//                  iterator's own ring walk) until the first blocked cell.
//
// Layers are column-major float32; the slab/halo conventions are those of te_chain.
#include "te_footprint.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace te {
namespace {

struct Layers {
  const float* __restrict__ trav;
  const float* __restrict__ slope;
  const float* __restrict__ step;
  const float* __restrict__ elev;
  const float* __restrict__ rough;  // traversability_roughness, only read when verify_rough is set
};

struct FpArgs {
  int rows, cols_total, in_col0, in_ncols, out_col0, out_ncols;
  double res, lenx, leny, posx, posy;
  const double* X;
  const double* Y;
  double rmin, rmax, rmax2, tdefault, maxgap, crit;
  int int_norm;
  int verify_rough;  // checkForRoughness_ (TraversabilityMap.cpp:779)
  int n_spiral;
  const int* spiral;  // di & 0xff | (dj & 0xff) << 8 | edge << 16
  int slope_R, step_R;
  // prefix-sum sweep (k_fp_prepare / k_sweep_fast)
  int L;                      // max |column offset| of the certain-in disk
  int nrings;                 // SpiralIterator nRings
  const signed char* halfw;   // [2L+1]: half-width of the certain-in disk per column offset (-1: none)
  const signed char* inner;   // [(nrings+2) x (2L+1)]: half-width of {k^2+l^2 < d^2} intersected with the disk, per ring d
  const int* ring_start;      // [nrings+2]: first index of ring d in `spiral`
  int n_fuzzy;                // offsets lying exactly on the circle: decided per cell on absolute positions
  const int* fuzzy;           // same packing as `spiral`
  const double* P;            // per input-buffer column: rows+1 prefix sums of t' along the row index
  const unsigned* bits;       // per input-buffer column: (rows+31)/32 words of blocked flags
  int words;                  // words per column
  const unsigned char* near;  // per input-buffer cell: distance (rows) to the nearest blocked cell of its column, 255 = none within 31
  signed char halfw_c[64];    // copy of `halfw` in the kernel parameters (constant bank)
  // full-disk sum without a blocked cell in sight (the common case): element offsets into the prefix sums, relative to the
  // centre's own entry P[column j][row i], of the two ends of disk column l (index l + L), and the running cell count
  int off_hi[64], off_lo[64];
  short cntp[65];
  // the same for k_sweep_fast as non-negative BYTE offsets from the prefix entry L columns and L rows before the centre's own
  // (one 32-bit add to a 64-bit base per load instead of a sign-extended 64-bit index computation)
  unsigned off8_hi[64], off8_lo[64];
};

__device__ __forceinline__ float lay(const FpArgs& A, const float* l, int i, int j) {  // caller guarantees (i,j) is in the map
  const int lb = j - A.in_col0;
  if (lb < 0 || lb >= A.in_ncols) return nanf_();
  return __ldg(l + (size_t)lb * A.rows + i);
}

__device__ __forceinline__ double cell_coord_d(double map_pos, double length, double res, int idx) {
  const double offset = 0.5 * length - 0.5 * res;
  return (map_pos + offset) + res * (-(double)idx);
}

__device__ __forceinline__ bool is_inside_d(const FpArgs& A, double px, double py) {
  const double tx = -((px - A.posx) - 0.5 * A.lenx);
  const double ty = -((py - A.posy) - 0.5 * A.leny);
  return tx >= 0.0 && ty >= 0.0 && tx < A.lenx && ty < A.leny;
}

__device__ __forceinline__ bool get_index_d(const FpArgs& A, double px, double py, int& i, int& j) {
  const double vx = ((px - 0.5 * A.lenx) - A.posx) / A.res;
  const double vy = ((py - 0.5 * A.leny) - A.posy) / A.res;
  i = (int)(-vx);
  j = (int)(-vy);
  return is_inside_d(A, px, py) && i >= 0 && j >= 0 && i < A.rows && j < A.cols_total;
}

__device__ __forceinline__ void bound_position_d(const FpArgs& A, double& px, double& py) {
  double sx = (px - A.posx) + 0.5 * A.lenx, sy = (py - A.posy) + 0.5 * A.leny;
  double ex = 10.0 * 2.220446049250313e-16, ey = ex;
  if (fabs(px) > 1.0) ex *= fabs(px);
  if (fabs(py) > 1.0) ey *= fabs(py);
  if (sx <= 0.0) sx = ex; else if (sx >= A.lenx) sx = A.lenx - ex;
  if (sy <= 0.0) sy = ey; else if (sy >= A.leny) sy = A.leny - ey;
  px = (sx + A.posx) - 0.5 * A.lenx;
  py = (sy + A.posy) - 0.5 * A.leny;
}

template <class F>
__device__ __forceinline__ void for_circle_d(const FpArgs& A, int i, int j, double r2, int R, F&& f) {
  const int a0 = max(0, i - R), a1 = min(A.rows - 1, i + R);
  const int b0 = max(0, j - R), b1 = min(A.cols_total - 1, j + R);
  const double cx = A.X[i], cy = A.Y[j];
  for (int a = a0; a <= a1; ++a) {
    const double dx = A.X[a] - cx;
    for (int b = b0; b <= b1; ++b) {
      const double dy = A.Y[b] - cy;
      if (dx * dx + dy * dy <= r2) f(a, b);
    }
  }
}

// TraversabilityMap::checkForSlope, TraversabilityMap.cpp:867-893.
__device__ bool check_slope_d(const FpArgs& A, const Layers& L, int i, int j) {
  if (!(lay(A, L.slope, i, j) == 0.0f)) return true;
  const double windowRadius = 3.0 * A.res;
  const double criticalLength = A.maxgap / 3.0;
  const int nCrit = (int)floor(2 * windowRadius * criticalLength / (A.res * A.res));
  int n = 0;
  for_circle_d(A, i, j, windowRadius * windowRadius, A.slope_R, [&](int a, int b) {
    if (lay(A, L.slope, a, b) == 0.0f) ++n;
  });
  return !(n > nCrit);
}

// TraversabilityMap::checkForRoughness, TraversabilityMap.cpp:895-921 (the slope check on another layer with another count).
__device__ bool check_rough_d(const FpArgs& A, const Layers& L, int i, int j) {
  if (!(lay(A, L.rough, i, j) == 0.0f)) return true;
  const double windowRadius = 3.0 * A.res;
  const double criticalLength = A.maxgap / 3.0;
  const int nCrit = (int)floor(1.5 * windowRadius * criticalLength / (A.res * A.res));
  int n = 0;
  for_circle_d(A, i, j, windowRadius * windowRadius, A.slope_R, [&](int a, int b) {
    if (lay(A, L.rough, a, b) == 0.0f) ++n;
  });
  return !(n > nCrit);
}

// TraversabilityMap::checkForStep, TraversabilityMap.cpp:794-865.
__device__ bool check_step_d(const FpArgs& A, const Layers& L, int i, int j) {
  if (!(lay(A, L.step, i, j) == 0.0f)) return true;
  const double crit = A.crit;
  const double wr = 2.5 * A.res;
  const double cx = A.X[i], cy = A.Y[j];
  const double h0 = (double)lay(A, L.elev, i, j);
  // candidate cells of the 2.5*res circle (at most 21): a bit per box cell, row index outer
  const int R = A.step_R;
  unsigned long long cand = 0ull;
  {
    const int a0 = max(0, i - R), a1 = min(A.rows - 1, i + R);
    const int b0 = max(0, j - R), b1 = min(A.cols_total - 1, j + R);
    for (int a = a0; a <= a1; ++a) {
      const double dx = A.X[a] - cx;
      for (int b = b0; b <= b1; ++b) {
        const double dy = A.Y[b] - cy;
        if (!(dx * dx + dy * dy <= wr * wr)) continue;
        if ((double)lay(A, L.elev, a, b) > crit + h0 && lay(A, L.step, a, b) == 0.0f)
          cand |= 1ull << ((a - (i - R)) * (2 * R + 1) + (b - (j - R)));
      }
    }
  }
  const bool self_only = cand == 0ull;
  if (self_only) cand = 1ull << (R * (2 * R + 1) + R);
  for (int bit = 0; bit < (2 * R + 1) * (2 * R + 1); ++bit) {
    if (!((cand >> bit) & 1ull)) continue;
    const int a = i - R + bit / (2 * R + 1), b = j - R + bit % (2 * R + 1);
    const double sx = A.X[a], sy = A.Y[b];      // subMapPos
    const double tcx = cx - sx, tcy = cy - sy;  // toCenter
    const double half = 0.5 * (2.5 * A.res);
    double tlx = sx + half, tly = sy + half;
    bound_position_d(A, tlx, tly);
    int ti, tj, bi, bj;
    if (!get_index_d(A, tlx, tly, ti, tj)) return false;
    double brx = sx - half, bry = sy - half;
    bound_position_d(A, brx, bry);
    if (!get_index_d(A, brx, bry, bi, bj)) return false;
    const double cornx = A.X[ti] + 0.5 * A.res, corny = A.Y[tj] + 0.5 * A.res;
    const int srows = bi - ti + 1, scols = bj - tj + 1;
    const double slx = (double)srows * A.res, sly = (double)scols * A.res;
    const double spx = cornx - 0.5 * slx, spy = corny - 0.5 * sly;
    const double height = (double)lay(A, L.elev, a, b);
    for (int k = 0; k < srows * scols; ++k) {
      const int si = k % srows, sj = k / srows;
      const int pi = ti + si, pj = tj + sj;
      if (!(lay(A, L.step, pi, pj) == 0.0f && (double)lay(A, L.elev, pi, pj) < height - crit)) continue;
      double px = cell_coord_d(spx, slx, A.res, si), py = cell_coord_d(spy, sly, A.res, sj);
      const double vx = px - sx, vy = py - sy;
      if (sqrt(vx * vx + vy * vy) < 0.025) continue;
      if (sqrt(tcx * tcx + tcy * tcy) > 0.025) {
        if (tcx * vx + tcy * vy < 0.0) continue;
      }
      px = sx + vx;
      py = sy + vy;
      for (;;) {
        const double ex = (px - sx) + vx, ey = (py - sy) + vy;
        if (!(sqrt(ex * ex + ey * ey) < A.maxgap && is_inside_d(A, px + vx, py + vy))) break;
        px = px + vx;
        py = py + vy;
      }
      int ei, ej;
      get_index_d(A, px, py, ei, ej);
      // LineIterator (Bresenham) from (a,b) to (ei,ej)
      const int dx = abs(ei - a), dy = abs(ej - b);
      int i1x = (ei >= a) ? 1 : -1, i2x = i1x, i1y = (ej >= b) ? 1 : -1, i2y = i1y;
      int den, num, numAdd, nCells;
      if (dx >= dy) { i1x = 0; i2y = 0; den = dx; num = dx / 2; numAdd = dy; nCells = dx + 1; }
      else { i2x = 0; i1y = 0; den = dy; num = dy / 2; numAdd = dx; nCells = dy + 1; }
      int li = a, lj = b;
      bool gapStart = false, gapEnd = false;
      for (int c = 0; c < nCells; ++c) {
        if (li < 0 || lj < 0 || li >= A.rows || lj >= A.cols_total) break;
        const float ef = lay(A, L.elev, li, lj);
        const double e = (double)ef;
        if (e > height + crit) return false;
        if (e < height - crit || !finitef(ef)) {
          gapStart = true;
        } else if (gapStart) {
          gapEnd = true;
          break;
        }
        num += numAdd;
        if (num >= den) { num -= den; li += i1x; lj += i1y; }
        li += i2x; lj += i2y;
      }
      if (gapStart && !gapEnd) return false;
    }
  }
  return true;
}

// isTraversableForFilters for every cell of the input buffer (columns in_col0 .. in_col0+in_ncols), in two steps.  Almost every
// cell passes trivially — checkForSlope / checkForStep / checkForRoughness return true at once unless the cell's own layer
// value is exactly 0 (TraversabilityMap.cpp:869, :796, :897) — so k_pred_classify settles those with two or three coalesced
// loads and collects the others on a work list, which k_pred_heavy walks with one thread per listed cell (the window count,
// the submap / gap walk): the heavy threads are no longer scattered one or two per warp over the whole map.
__global__ void __launch_bounds__(256) k_pred_classify(FpArgs A, Layers L, unsigned char* __restrict__ blocked, float* slope_fp,
                                                       float* step_fp, float* rough_fp, unsigned* __restrict__ list,
                                                       unsigned* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lb = blockIdx.y;
  const bool in = i < A.rows;
  bool heavy = false;
  size_t c = 0;
  if (in) {
    c = (size_t)lb * A.rows + i;
    heavy = __ldg(L.slope + c) == 0.0f || __ldg(L.step + c) == 0.0f || (A.verify_rough && __ldg(L.rough + c) == 0.0f);
    if (!heavy) {
      blocked[c] = 0;
      const int oj = lb + A.in_col0 - A.out_col0;
      if (oj >= 0 && oj < A.out_ncols) {
        const size_t oc = (size_t)oj * A.rows + i;
        if (slope_fp) slope_fp[oc] = nanf_();
        if (step_fp) step_fp[oc] = nanf_();
        if (rough_fp) rough_fp[oc] = nanf_();
      }
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, heavy);
  if (m) {
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (heavy) list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned)c;
  }
}

__global__ void __launch_bounds__(128) k_pred_heavy(FpArgs A, Layers L, unsigned char* __restrict__ blocked, float* slope_fp, float* step_fp,
                                                    float* rough_fp, const unsigned* __restrict__ list, const unsigned* __restrict__ count) {
  const unsigned n = *count;
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const unsigned c = list[k];
    const int i = (int)(c % (unsigned)A.rows);
    const int j = A.in_col0 + (int)(c / (unsigned)A.rows);
    const bool s_ok = check_slope_d(A, L, i, j);
    bool t_ok = true;
    float sfp = nanf_(), tfp = nanf_();
    if (lay(A, L.slope, i, j) == 0.0f) sfp = s_ok ? 1.0f : 0.0f;
    if (s_ok) {
      t_ok = check_step_d(A, L, i, j);
      if (lay(A, L.step, i, j) == 0.0f) tfp = t_ok ? 1.0f : 0.0f;
    }
    bool r_ok = true;
    float rfp = nanf_();
    if (A.verify_rough && s_ok && t_ok) {  // TraversabilityMap.cpp:779-783: only after slope and step passed
      r_ok = check_rough_d(A, L, i, j);
      if (lay(A, L.rough, i, j) == 0.0f) rfp = r_ok ? 1.0f : 0.0f;
    }
    blocked[c] = (s_ok && t_ok && r_ok) ? 0 : 1;
    const int oj = j - A.out_col0;
    if (oj >= 0 && oj < A.out_ncols) {
      const size_t oc = (size_t)oj * A.rows + i;
      if (slope_fp) slope_fp[oc] = sfp;
      if (step_fp) step_fp[oc] = tfp;
      if (rough_fp) rough_fp[oc] = rfp;
    }
  }
}

__global__ void __launch_bounds__(256) k_sweep(FpArgs A, Layers L, const unsigned char* __restrict__ blocked, float* __restrict__ out) {
  const long long total = (long long)A.rows * A.out_ncols;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(c % A.rows);
    const int j = A.out_col0 + (int)(c / A.rows);
    const double cx = A.X[i], cy = A.Y[j];
    int n = 0;
    double t = 0.0;
    float result = nanf_();
    bool done = false;
    for (int k = 0; k < A.n_spiral; ++k) {
      const int w = __ldg(A.spiral + k);
      const int di = (int)(signed char)(w & 0xff), dj = (int)(signed char)((w >> 8) & 0xff);
      const int a = i + di, b = j + dj;
      if (a < 0 || b < 0 || a >= A.rows || b >= A.cols_total) continue;
      if (w & 0x10000) {
        const double dx = A.X[a] - cx, dy = A.Y[b] - cy;
        if (!(dx * dx + dy * dy <= A.rmax2)) continue;
      }
      const int lb = b - A.in_col0;
      if (lb < 0 || lb >= A.in_ncols) continue;  // cannot happen with the halo te_footprint demands
      const size_t cc = (size_t)lb * A.rows + a;
      if (blocked[cc]) {
        const int d2 = di * di + dj * dj;
        const double nr = A.int_norm ? (double)(int)sqrt((double)d2) : sqrt((double)d2);
        const double uR = nr * A.res;
        if (A.rmin == 0.0 || uR <= A.rmin) {
          result = 0.0f;
        } else {
          const double factor = ((uR - A.rmin) / (A.rmax - A.rmin) + 1.0) / 2.0;
          t *= factor / (double)n;
          result = (float)t;
        }
        done = true;
        break;
      }
      ++n;
      const float v = __ldg(L.trav + cc);
      t += finitef(v) ? (double)v : A.tdefault;
    }
    if (!done) {
      t /= (double)n;
      result = (float)t;
    }
    out[c] = result;
  }
}

// One warp per input-buffer column: the blocked flags packed 32 rows per word (the tiled sweep's early outs and k_fp_nearest
// read them; the prefix sums of t' are built per tile in shared memory by k_sweep_tile).
__global__ void __launch_bounds__(256) k_fp_prepare(FpArgs A, const unsigned char* __restrict__ blocked, unsigned* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int lb = warp; lb < A.in_ncols; lb += nwarps) {
    const unsigned char* bcol = blocked + (size_t)lb * A.rows;
    unsigned* wcol = bits + (size_t)lb * A.words;
    for (int base = 0; base < A.rows; base += 32) {
      const int i = base + lane;
      const unsigned w = __ballot_sync(0xffffffffu, i < A.rows && bcol[i] != 0);
      if (lane == 0) wcol[base >> 5] = w;
    }
  }
}

// Distance along the row index from every cell to the nearest blocked cell of its own column (from the packed flags):
// the sweep then needs ONE byte per disk column to know the nearest blocked cell of that column.
__global__ void __launch_bounds__(256) k_fp_nearest(FpArgs A, unsigned char* __restrict__ near) {
  const long long total = (long long)A.rows * A.in_ncols;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(c % A.rows);
    const int lb = (int)(c / A.rows);
    const unsigned* wcol = A.bits + (size_t)lb * A.words;
    const int r0 = i - 31, w0 = r0 >> 5, sh = r0 & 31;
    const unsigned a0 = (w0 >= 0 && w0 < A.words) ? __ldg(wcol + w0) : 0u;
    const unsigned a1 = (w0 + 1 >= 0 && w0 + 1 < A.words) ? __ldg(wcol + w0 + 1) : 0u;
    const unsigned a2 = (w0 + 2 >= 0 && w0 + 2 < A.words) ? __ldg(wcol + w0 + 2) : 0u;
    const unsigned long long lo = ((unsigned long long)a1 << 32) | a0;
    unsigned long long f = (lo >> sh) | (sh ? ((unsigned long long)a2 << (64 - sh)) : 0ull);  // bit t <-> row r0 + t, centre at bit 31
    f &= ~(1ull << 63);                                                                        // rows i-31 .. i+31
    int kmin = 255;
    const unsigned long long up = f >> 31, dn = f & ((1ull << 31) - 1ull);
    if (up) kmin = __ffsll((long long)up) - 1;
    if (dn) kmin = min(kmin, 31 - (63 - __clzll((long long)dn)));
    near[c] = (unsigned char)kmin;
  }
}

// One warp per input-buffer column: prefix sums of t' = finite(traversability) ? value : default along the row
// index (double; exact for float32 terms, so the order of summation does not matter) and packed blocked flags.
__global__ void __launch_bounds__(256) k_fp_prepare_p(FpArgs A, Layers L, const unsigned char* __restrict__ blocked, double* __restrict__ P,
                                                    unsigned* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int lb = warp; lb < A.in_ncols; lb += nwarps) {
    const float* tcol = L.trav + (size_t)lb * A.rows;
    const unsigned char* bcol = blocked + (size_t)lb * A.rows;
    double* pcol = P + (size_t)lb * (A.rows + 1);
    unsigned* wcol = bits + (size_t)lb * A.words;
    double carry = 0.0;
    if (lane == 0) pcol[0] = 0.0;
    for (int base = 0; base < A.rows; base += 32) {
      const int i = base + lane;
      double v = 0.0;
      bool b = false;
      if (i < A.rows) {
        const float t = __ldg(tcol + i);
        v = finitef(t) ? (double)t : A.tdefault;
        b = bcol[i] != 0;
      }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
      }
      if (i < A.rows) pcol[i + 1] = carry + v;
      const unsigned w = __ballot_sync(0xffffffffu, b);
      if (lane == 0) wcol[base >> 5] = w;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
}

// isTraversable for every cell on prefix sums: the visited set is a lattice disk, so "is anything blocked in it"
// and "sum / count of the visited cells" are 2L+1 column queries instead of ~pi r^2 visits; only when a blocker
// exists is the ring that holds the first one walked in SpiralIterator order.
__global__ void __launch_bounds__(256) k_sweep_fast(FpArgs A, Layers L, const unsigned char* __restrict__ blocked, float* __restrict__ out) {
  const int W = 2 * A.L + 1;
  // one block = 256 consecutive rows of ONE output column (blockIdx.y): a warp's centres share their column, so everything that
  // depends only on the column — which disk columns exist, whether any blocked cell lies near the warp at all — is warp-uniform
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = A.out_col0 + (int)blockIdx.y;
  {
    const int lane = threadIdx.x & 31, i0 = i - lane;
    if (i0 >= A.rows) return;  // whole warp beyond the last row
    const bool active = i < A.rows;
    const size_t c = (size_t)blockIdx.y * A.rows + (active ? i : 0);
    const double cx = A.X[active ? i : 0], cy = A.Y[j];
    // disk columns that exist in the map and in this slab's buffer
    const int l_lo = max(-A.L, max(-j, A.in_col0 - j)), l_hi = min(A.L, min(A.cols_total - 1 - j, A.in_col0 + A.in_ncols - 1 - j));
    // ---- warp-wide early out: is any cell blocked in the box of rows [i0 - L, i0 + 31 + L] x disk columns?  (packed flags,
    //      a few words per column, shared by the 32 centres)  Mostly not: then no lane has anything to look for.
    bool warp_any;
    {
      const int w0 = max(i0 - A.L, 0) >> 5, w1 = min(i0 + 31 + A.L, A.rows - 1) >> 5;
      unsigned acc = 0;
      for (int l = l_lo + lane; l <= l_hi; l += 32) {  // a lane per disk column, at most four words each
        const unsigned* wc = A.bits + (size_t)(j + l - A.in_col0) * A.words;
        for (int w = w0; w <= w1; ++w) acc |= __ldg(wc + w);
      }
      warp_any = __any_sync(0xffffffffu, acc != 0u);
    }
    if (!active) return;
    // ---- nearest blocked cell of the visited set, as a squared index distance ------------------------
    int best = 0x7fffffff;
    if (warp_any) {
      const unsigned char* nr = A.near + (size_t)(j + l_lo - A.in_col0) * A.rows + i;
      for (int l = l_lo; l <= l_hi; ++l, nr += A.rows) {
        const int g = (int)__ldg(nr);              // nearest blocked row offset in this column
        const int hw = A.halfw_c[l + A.L];
        if (g <= hw) best = min(best, g * g + l * l);
      }
      for (int q = 0; q < A.n_fuzzy; ++q) {
        const int w = A.fuzzy[q];
        const int di = (int)(signed char)(w & 0xff), dj = (int)(signed char)((w >> 8) & 0xff);
        const int a = i + di, b = j + dj, lb = b - A.in_col0;
        if (a < 0 || b < 0 || a >= A.rows || b >= A.cols_total || lb < 0 || lb >= A.in_ncols) continue;
        const double dx = A.X[a] - cx, dy = A.Y[b] - cy;
        if (!(dx * dx + dy * dy <= A.rmax2)) continue;
        if (blocked[(size_t)lb * A.rows + a]) best = min(best, di * di + dj * dj);
      }
    }
    // ---- sums over the visited cells before the first blocked one --------------------------------------
    const bool any = best != 0x7fffffff;
    const int dstar = any ? (int)sqrt((double)best) : A.nrings + 1;  // ring of the first blocked cell
    // The first blocked cell in visit order lies in ring dstar: its index-space radius is in [dstar, dstar + 1) (exactly dstar with
    // the integer norm).  Within the inner radius the result is 0 (TraversabilityMap.cpp:694-704) whatever the sums are: most
    // centres near an obstacle end here, without prefix sums or a ring walk.
    if (any && (A.rmin == 0.0 || (A.int_norm ? (double)dstar : (double)(dstar + 1)) * A.res <= A.rmin)) {
      out[c] = 0.0f;
      return;
    }
    const signed char* hwt = any ? (A.inner + (size_t)dstar * W) : A.halfw;
    double t = 0.0, t_b = 0.0;
    int n = 0;
    if (!warp_any && i - A.L >= 0 && i + A.L < A.rows) {
      // nothing blocked near this warp and no clipping along the rows: 2 loads and 2 additions per disk column, offsets from
      // the constant bank
      // biased base: the entry L columns and L rows before the centre's own, so that every table offset is a non-negative byte count
      const char* pb = reinterpret_cast<const char*>(A.P + ((size_t)(j - A.in_col0) * ((size_t)A.rows + 1) + i)) -
                       8 * ((ptrdiff_t)A.L * ((ptrdiff_t)A.rows + 1) + A.L);
      auto P8 = [&](unsigned off) { return __ldg(reinterpret_cast<const double*>(pb + off)); };
      int l = l_lo + A.L;
      const int l_end = l_hi + A.L;
      for (; l + 1 <= l_end; l += 2) {
        t += P8(A.off8_hi[l]) - P8(A.off8_lo[l]);
        t_b += P8(A.off8_hi[l + 1]) - P8(A.off8_lo[l + 1]);
      }
      if (l <= l_end) t += P8(A.off8_hi[l]) - P8(A.off8_lo[l]);
      t += t_b;
      n = (int)A.cntp[l_end + 1] - (int)A.cntp[l_lo + A.L];
    } else {
      const double* pc = A.P + (size_t)(j + l_lo - A.in_col0) * (A.rows + 1);
      const size_t pstride = (size_t)A.rows + 1;
      if (i - A.L >= 0 && i + A.L < A.rows) {  // no clipping along the rows: two independent accumulators
        int l = l_lo;
        for (; l + 1 <= l_hi; l += 2, pc += 2 * pstride) {
          const int h0 = hwt[l + A.L], h1 = hwt[l + 1 + A.L];
          if (h0 >= 0) { t += pc[i + h0 + 1] - pc[i - h0]; n += 2 * h0 + 1; }
          if (h1 >= 0) { t_b += pc[pstride + i + h1 + 1] - pc[pstride + i - h1]; n += 2 * h1 + 1; }
        }
        if (l <= l_hi) {
          const int h0 = hwt[l + A.L];
          if (h0 >= 0) { t += pc[i + h0 + 1] - pc[i - h0]; n += 2 * h0 + 1; }
        }
      } else {
        for (int l = l_lo; l <= l_hi; ++l, pc += pstride) {
          const int hw = hwt[l + A.L];
          if (hw < 0) continue;
          const int a0 = max(i - hw, 0), a1 = min(i + hw, A.rows - 1);
          t += pc[a1 + 1] - pc[a0];
          n += a1 - a0 + 1;
        }
      }
      t += t_b;
    }
    float result;
    if (!any) {
      for (int q = 0; q < A.n_fuzzy; ++q) {  // on-circle cells belong to the last ring: they are visited last
        const int w = A.fuzzy[q];
        const int di = (int)(signed char)(w & 0xff), dj = (int)(signed char)((w >> 8) & 0xff);
        const int a = i + di, b = j + dj, lb = b - A.in_col0;
        if (a < 0 || b < 0 || a >= A.rows || b >= A.cols_total || lb < 0 || lb >= A.in_ncols) continue;
        const double dx = A.X[a] - cx, dy = A.Y[b] - cy;
        if (!(dx * dx + dy * dy <= A.rmax2)) continue;
        const float v = __ldg(L.trav + (size_t)lb * A.rows + a);
        t += finitef(v) ? (double)v : A.tdefault;
        ++n;
      }
      t /= (double)n;
      result = (float)t;
    } else {
      // walk ring dstar in visit order up to its first blocked cell
      int di = 0, dj = 0;
      for (int k = A.ring_start[dstar]; k < A.ring_start[dstar + 1]; ++k) {
        const int w = __ldg(A.spiral + k);
        di = (int)(signed char)(w & 0xff);
        dj = (int)(signed char)((w >> 8) & 0xff);
        const int a = i + di, b = j + dj, lb = b - A.in_col0;
        if (a < 0 || b < 0 || a >= A.rows || b >= A.cols_total || lb < 0 || lb >= A.in_ncols) continue;
        if (w & 0x10000) {
          const double dx = A.X[a] - cx, dy = A.Y[b] - cy;
          if (!(dx * dx + dy * dy <= A.rmax2)) continue;
        }
        const size_t cc = (size_t)lb * A.rows + a;
        if (blocked[cc]) break;
        const float v = __ldg(L.trav + cc);
        t += finitef(v) ? (double)v : A.tdefault;
        ++n;
      }
      const int d2 = di * di + dj * dj;
      const double nr = A.int_norm ? (double)(int)sqrt((double)d2) : sqrt((double)d2);
      const double uR = nr * A.res;
      if (A.rmin == 0.0 || uR <= A.rmin) {
        result = 0.0f;
      } else {
        const double factor = ((uR - A.rmin) / (A.rmax - A.rmin) + 1.0) / 2.0;
        t *= factor / (double)n;
        result = (float)t;
      }
    }
    out[c] = result;
  }
}

// The sweep, tiled: a CTA owns TR x TC centres and stages what their disks touch — the column prefix sums of t' over the tile's
// rows plus L rows of halo, for the tile's columns plus L columns of halo — in shared memory, computing them in place from the
// traversability layer (prefix sums local to the tile: small magnitudes, no 8-byte-per-cell array in HBM).  k_sweep_fast sends
// 2(2L+1) 8-byte loads per centre to L1/L2 (12 GB of L2 traffic at 4096^2, r = 0.45 m: that, not HBM, bounds it); here they are
// conflict-free LDS.64 (lanes = consecutive rows).  Cells outside the map or the slab's buffer contribute 0 to a prefix sum, so
// sums need no clipping — only the cell count of a centre near the border does.
constexpr int TR = 64, TC = 32;
static_assert(TR + 2 * 31 <= 128, "k_sweep_tile scans a staged column (TR + 2L rows, L <= 31) in one pass of 4 rows per lane");
__global__ void __launch_bounds__(256) k_sweep_tile(FpArgs A, Layers L, const unsigned char* __restrict__ blocked, float* __restrict__ out) {
  extern __shared__ double sP[];  // [NC][PS]: sP[c][k] = sum of t' over staged rows [0, k) of staged column c
  const int Lr = A.L, NR = TR + 2 * Lr, NC = TC + 2 * Lr, PS = NR + 1, W = 2 * Lr + 1;
  unsigned char* sNear = reinterpret_cast<unsigned char*>(sP + (size_t)NC * PS);  // [NC][TR]: nearest-blocked bytes of the centre rows
  unsigned char* sBlk = sNear + (size_t)NC * TR;                                   // [NC][128]: blocked flag of every staged cell
  const int r0 = (int)blockIdx.x * TR, c0 = A.out_col0 + (int)blockIdx.y * TC;
  const int rb = r0 - Lr, cb = c0 - Lr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto tprime = [&](int a, int b) { const double* q = sP + (size_t)(b - cb) * PS + (a - rb); return q[1] - q[0]; };  // one float32 term: exact
  // ---- phase 1: prefix sums (one warp per staged column, lanes along the rows) and "is anything blocked near this tile"
  int anyb = 0;
  for (int cc = warp; cc < NC; cc += 8) {
    const int gcol = cb + cc, lb = gcol - A.in_col0;
    const bool col_ok = gcol >= 0 && gcol < A.cols_total && lb >= 0 && lb < A.in_ncols;
    const float* tcol = L.trav + (size_t)(col_ok ? lb : 0) * A.rows;
    const unsigned char* bcol = blocked + (size_t)(col_ok ? lb : 0) * A.rows;
    double* pcol = sP + (size_t)cc * PS;
    // NR <= 4 * 32 (L <= 31): a lane takes four consecutive rows, scans them in registers, the warp scans the lane totals
    double v[4];
    unsigned bw = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = 4 * lane + m, row = rb + k;
      v[m] = 0.0;
      if (k < NR && col_ok && row >= 0 && row < A.rows) {
        const float t = __ldg(tcol + row);
        v[m] = finitef(t) ? (double)t : A.tdefault;
        bw |= (unsigned)(bcol[row] != 0) << (8 * m);
      }
    }
    anyb |= (int)bw;
    reinterpret_cast<unsigned*>(sBlk)[cc * 32 + lane] = bw;
    v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
    double tot = v[3];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double o = __shfl_up_sync(0xffffffffu, tot, d);
      if (lane >= d) tot += o;
    }
    double before = __shfl_up_sync(0xffffffffu, tot, 1);  // sum of the rows held by the lanes below
    if (lane == 0) before = 0.0;
    if (lane == 0) pcol[0] = 0.0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = 4 * lane + m;
      if (k < NR) pcol[k + 1] = before + v[m];
    }
  }
  const bool tile_any = __syncthreads_or(anyb) != 0;
  if (tile_any) {  // the tile's rows of the nearest-blocked bytes, every staged column (TR = 64 bytes per column: 16 words)
    for (int e = threadIdx.x; e < NC * (TR / 4); e += blockDim.x) {
      const int cc = e / (TR / 4), w = e % (TR / 4);
      const int gcol = cb + cc, lb = gcol - A.in_col0, row = r0 + 4 * w;
      unsigned v = 0xffffffffu;  // 255: nothing blocked within reach
      if (gcol >= 0 && gcol < A.cols_total && lb >= 0 && lb < A.in_ncols) {
        const unsigned char* src = A.near + (size_t)lb * A.rows + row;
        if (row + 3 < A.rows && ((size_t)src & 3u) == 0) v = __ldg(reinterpret_cast<const unsigned*>(src));
        else {
          v = 0;
          for (int m = 0; m < 4; ++m) v |= (unsigned)((row + m < A.rows) ? __ldg(src + m) : 255u) << (8 * m);
        }
      }
      reinterpret_cast<unsigned*>(sNear)[cc * (TR / 4) + w] = v;
    }
    __syncthreads();
  }
  // ---- phase 2: centres.  A warp is 32 consecutive rows of one column at a time (8 columns per warp).
  const int i = r0 + (warp & 1) * 32 + lane;
  const int i0 = i - lane;
  if (i0 >= A.rows) return;
  const bool active = i < A.rows;
  const double cx = A.X[active ? i : 0];
  const int k0 = i - rb;  // centre's staged row
  // One bit per buffer column the warp's centres can reach: "a cell of rows [i0 - L, i0 + 31 + L] of this column is blocked" (from
  // the packed flags; one pass for all TC/4 centres of the warp).  A centre whose disk columns are all clear has nothing to look for.
  unsigned long long cm0 = 0ull, cm1 = 0ull;
  int ca = 0;
  if (tile_any) {
    const int ja = c0 + (warp >> 1) * (TC / 4), jb = min(ja + TC / 4 - 1, A.out_col0 + A.out_ncols - 1);
    ca = max(ja - Lr, max(0, A.in_col0));
    const int cz = min(jb + Lr, min(A.cols_total - 1, A.in_col0 + A.in_ncols - 1));
    const int w0 = max(i0 - Lr, 0) >> 5, w1 = min(i0 + 31 + Lr, A.rows - 1) >> 5;
    for (int e0 = 0; e0 <= cz - ca; e0 += 32) {
      unsigned acc = 0;
      if (e0 + lane <= cz - ca) {
        const unsigned* wc = A.bits + (size_t)(ca + e0 + lane - A.in_col0) * A.words;
        for (int w = w0; w <= w1; ++w) acc |= __ldg(wc + w);
      }
      const unsigned long long bal = (unsigned long long)__ballot_sync(0xffffffffu, acc != 0u);
      if (e0 < 64) cm0 |= bal << e0; else cm1 |= bal << (e0 - 64);
    }
  }
  for (int q = 0; q < TC / 4; ++q) {
    const int j = c0 + (warp >> 1) * (TC / 4) + q;
    if (j >= A.out_col0 + A.out_ncols) break;
    const size_t c = (size_t)(j - A.out_col0) * A.rows + (active ? i : 0);
    const double cy = A.Y[j];
    const int l_lo = max(-Lr, max(-j, A.in_col0 - j)), l_hi = min(Lr, min(A.cols_total - 1 - j, A.in_col0 + A.in_ncols - 1 - j));
    bool warp_any = false;
    if (tile_any) {
      const int s0 = j + l_lo - ca, cnt = l_hi - l_lo + 1;  // cnt <= 63, bits [s0, s0 + cnt) of the 128-bit mask
      unsigned long long m = s0 < 64 ? (cm0 >> s0) : (cm1 >> (s0 - 64));
      if (s0 > 0 && s0 < 64) m |= cm1 << (64 - s0);
      warp_any = (m & ((1ull << cnt) - 1ull)) != 0ull;
    }
    if (!active) continue;
    int best = 0x7fffffff;
    if (warp_any) {
      const unsigned char* nr = sNear + (size_t)(j + l_lo - cb) * TR + (i - r0);
      for (int l = l_lo; l <= l_hi; ++l, nr += TR) {
        const int g = (int)*nr;
        const int hw = A.halfw_c[l + Lr];
        if (g <= hw) best = min(best, g * g + l * l);
      }
      for (int f = 0; f < A.n_fuzzy; ++f) {
        const int w = A.fuzzy[f];
        const int di = (int)(signed char)(w & 0xff), dj = (int)(signed char)((w >> 8) & 0xff);
        const int a = i + di, b = j + dj, lb = b - A.in_col0;
        if (a < 0 || b < 0 || a >= A.rows || b >= A.cols_total || lb < 0 || lb >= A.in_ncols) continue;
        const double dx = A.X[a] - cx, dy = A.Y[b] - cy;
        if (!(dx * dx + dy * dy <= A.rmax2)) continue;
        if (sBlk[(b - cb) * 128 + (a - rb)]) best = min(best, di * di + dj * dj);
      }
    }
    const bool any = best != 0x7fffffff;
    const int dstar = any ? (int)sqrt((double)best) : A.nrings + 1;
    // The first blocked cell in visit order lies in ring dstar, so its index-space radius is in [dstar, dstar + 1) (exactly dstar
    // with the integer norm): when that is within the inner radius the result is 0 (TraversabilityMap.cpp:694-704) whatever the
    // sums are — most centres near an obstacle end here, without prefix sums or a ring walk.
    if (any && (A.rmin == 0.0 || (A.int_norm ? (double)dstar : (double)(dstar + 1)) * A.res <= A.rmin)) {
      out[c] = 0.0f;
      continue;
    }
    const double* pc0 = sP + (size_t)(j - cb) * PS + k0;  // the centre's own prefix entry
    const bool interior = i - Lr >= 0 && i + Lr < A.rows;
    double t = 0.0, t_b = 0.0;
    int n = 0;
    if (!any) {
      // whole disk: offsets from the constant bank; columns that do not exist are all-zero prefix columns
      int l = 0;
      for (; l + 1 < W; l += 2) {
        t += pc0[A.off_hi[l]] - pc0[A.off_lo[l]];
        t_b += pc0[A.off_hi[l + 1]] - pc0[A.off_lo[l + 1]];
      }
      t += pc0[A.off_hi[l]] - pc0[A.off_lo[l]];
      t += t_b;
      if (interior) {
        n = (int)A.cntp[l_hi + Lr + 1] - (int)A.cntp[l_lo + Lr];
      } else {
        for (int l2 = l_lo; l2 <= l_hi; ++l2) {
          const int hw = A.halfw_c[l2 + Lr];
          if (hw >= 0) n += min(i + hw, A.rows - 1) - max(i - hw, 0) + 1;
        }
      }
    } else {
      const signed char* hwt = A.inner + (size_t)dstar * W;
      for (int l = l_lo; l <= l_hi; ++l) {
        const int hw = hwt[l + Lr];
        if (hw < 0) continue;
        t += pc0[l * PS + hw + 1] - pc0[l * PS - hw];
        n += min(i + hw, A.rows - 1) - max(i - hw, 0) + 1;
      }
    }
    float result;
    if (!any) {
      for (int f = 0; f < A.n_fuzzy; ++f) {  // on-circle cells belong to the last ring: they are visited last
        const int w = A.fuzzy[f];
        const int di = (int)(signed char)(w & 0xff), dj = (int)(signed char)((w >> 8) & 0xff);
        const int a = i + di, b = j + dj, lb = b - A.in_col0;
        if (a < 0 || b < 0 || a >= A.rows || b >= A.cols_total || lb < 0 || lb >= A.in_ncols) continue;
        const double dx = A.X[a] - cx, dy = A.Y[b] - cy;
        if (!(dx * dx + dy * dy <= A.rmax2)) continue;
        t += tprime(a, b);
        ++n;
      }
      t /= (double)n;
      result = (float)t;
    } else {
      int di = 0, dj = 0;
      for (int k = A.ring_start[dstar]; k < A.ring_start[dstar + 1]; ++k) {  // ring dstar in visit order up to its first blocked cell
        const int w = __ldg(A.spiral + k);
        di = (int)(signed char)(w & 0xff);
        dj = (int)(signed char)((w >> 8) & 0xff);
        const int a = i + di, b = j + dj, lb = b - A.in_col0;
        if (a < 0 || b < 0 || a >= A.rows || b >= A.cols_total || lb < 0 || lb >= A.in_ncols) continue;
        if (w & 0x10000) {
          const double dx = A.X[a] - cx, dy = A.Y[b] - cy;
          if (!(dx * dx + dy * dy <= A.rmax2)) continue;
        }
        if (sBlk[(b - cb) * 128 + (a - rb)]) break;
        t += tprime(a, b);
        ++n;
      }
      const int d2 = di * di + dj * dj;
      const double nr = A.int_norm ? (double)(int)sqrt((double)d2) : sqrt((double)d2);
      const double uR = nr * A.res;
      if (A.rmin == 0.0 || uR <= A.rmin) {
        result = 0.0f;
      } else {
        const double factor = ((uR - A.rmin) / (A.rmax - A.rmin) + 1.0) / 2.0;
        t *= factor / (double)n;
        result = (float)t;
      }
    }
    out[c] = result;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Polygonal footprint sweep: TraversabilityMap::traversabilityFootprint(double footprintYaw), TraversabilityMap.cpp:239-305, with the
// polygon isTraversable (:592-645).  Every cell gets the footprint polygon placed at its centre; the value is 0 when a cell of the
// polygon fails isTraversableForFilters, otherwise the mean of t' over the polygon's cells (traversabilityDefault_ when it covers
// none).  The set of cells inside the polygon is the same offset pattern for every centre EXCEPT for offsets whose cell centre lies
// on (within rounding of) an edge: grid_map::Polygon::isInside decides those on absolute double coordinates, differently from
// centre to centre.  The host classifies the offsets once (launch_footprint_polygon): certain-in cells become per-column runs that
// are summed from the tile's prefix sums; the few uncertain offsets are decided per centre with the reference's own arithmetic —
// cooperatively when the decision does not depend on the centre's row (an edge parallel to the x axis through cell centres: the
// YAML footprint at 0.02 m has 92 such offsets), 32 offsets per warp pass.
constexpr int PTR = 64, PTC = 16;  // tile of centres: rows x columns
constexpr int PMAXV = 16;          // polygon vertices
constexpr int PB = 130;            // pitch of a staged blocked-count column (uint16)
struct PolyArgs {
  int Lp;             // reach of the polygon in cells (<= 31)
  int nruns, nfz, npts;
  const int* runs;    // (dj & 0xff) | (lo & 0xff) << 8 | (hi & 0xff) << 16: rows i+lo .. i+hi of column j+dj are certainly inside
  const int* fz;      // (di & 0xff) | (dj & 0xff) << 8 | flags << 16: uncertain offsets, sorted by (dj, di); flag bit 0: the decision
                      // depends on the centre's row; bit 1: same column as the previous entry, next row, neither depends on the row
  int ncert;          // number of certain cells (sum of the run lengths)
  double r00, r01, r10, r11;  // Eigen::Quaternion::toRotationMatrix of (cos(yaw/2), 0, 0, sin(yaw/2)), upper-left 2 x 2
  double px[PMAXV], py[PMAXV];
};

// grid_map::Polygon::isInside for the polygon placed at (cx, cy): vertices = R * p + centre in the operand order of Eigen's
// Transform * vector (oracle: teo_footprint_polygon), crossing-number test over (v[i], v[i-1]).
__device__ bool poly_inside_d(const PolyArgs& Q, double cx, double cy, double ptx, double pty) {
  int cross = 0;
  const int last = Q.npts - 1;
  double jx = cx + ((Q.r00 * Q.px[last] + Q.r01 * Q.py[last]) + 0.0);
  double jy = cy + ((Q.r10 * Q.px[last] + Q.r11 * Q.py[last]) + 0.0);
  for (int k = 0; k < Q.npts; ++k) {
    const double ix = cx + ((Q.r00 * Q.px[k] + Q.r01 * Q.py[k]) + 0.0);
    const double iy = cy + ((Q.r10 * Q.px[k] + Q.r11 * Q.py[k]) + 0.0);
    if (((iy > pty) != (jy > pty)) && (ptx < (jx - ix) * (pty - iy) / (jy - iy) + ix)) ++cross;
    jx = ix;
    jy = iy;
  }
  return (cross & 1) != 0;
}

__global__ void __launch_bounds__(256) k_poly_tile(FpArgs A, PolyArgs Q, const float* __restrict__ trav, const unsigned char* __restrict__ blocked,
                                                   float* __restrict__ out) {
  extern __shared__ double sP[];  // [NC][PS] prefix sums of t'; then [NC][PB] uint16 prefix counts of blocked cells
  const int Lr = Q.Lp, NR = PTR + 2 * Lr, NC = PTC + 2 * Lr, PS = NR + 1;
  unsigned short* sB = reinterpret_cast<unsigned short*>(sP + (size_t)NC * PS);
  const int r0 = (int)blockIdx.x * PTR, c0 = A.out_col0 + (int)blockIdx.y * PTC;
  const int rb = r0 - Lr, cb = c0 - Lr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- phase 1 (as in k_sweep_tile, plus the blocked counts): one warp per staged column, four rows per lane
  int anyb = 0;
  for (int cc = warp; cc < NC; cc += 8) {
    const int gcol = cb + cc, lb = gcol - A.in_col0;
    const bool col_ok = gcol >= 0 && gcol < A.cols_total && lb >= 0 && lb < A.in_ncols;
    const float* tcol = trav + (size_t)(col_ok ? lb : 0) * A.rows;
    const unsigned char* bcol = blocked + (size_t)(col_ok ? lb : 0) * A.rows;
    double v[4];
    int cb4[4];
    unsigned bw = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = 4 * lane + m, row = rb + k;
      v[m] = 0.0;
      cb4[m] = 0;
      if (k < NR && col_ok && row >= 0 && row < A.rows) {
        const float t = __ldg(tcol + row);
        v[m] = finitef(t) ? (double)t : A.tdefault;
        cb4[m] = bcol[row] != 0 ? 1 : 0;
        bw |= (unsigned)cb4[m] << (8 * m);
      }
    }
    anyb |= (int)bw;
    v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
    cb4[1] += cb4[0]; cb4[2] += cb4[1]; cb4[3] += cb4[2];
    double tot = v[3];
    int ctot = cb4[3];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double o = __shfl_up_sync(0xffffffffu, tot, d);
      const int oc = __shfl_up_sync(0xffffffffu, ctot, d);
      if (lane >= d) { tot += o; ctot += oc; }
    }
    double before = __shfl_up_sync(0xffffffffu, tot, 1);
    int cbefore = __shfl_up_sync(0xffffffffu, ctot, 1);
    if (lane == 0) { before = 0.0; cbefore = 0; }
    double* pcol = sP + (size_t)cc * PS;
    unsigned short* ccol = sB + (size_t)cc * PB;
    if (lane == 0) { pcol[0] = 0.0; ccol[0] = 0; }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = 4 * lane + m;
      if (k < NR) { pcol[k + 1] = before + v[m]; ccol[k + 1] = (unsigned short)(cbefore + cb4[m]); }
    }
  }
  const bool tile_any = __syncthreads_or(anyb) != 0;
  // ---- phase 2: a warp is 32 consecutive rows of one column at a time (4 columns per warp)
  const int i = r0 + (warp & 1) * 32 + lane;
  const int i0 = i - lane;
  if (i0 >= A.rows) return;
  const bool active = i < A.rows;
  const int ic = active ? i : A.rows - 1;
  const double cx = A.X[ic];
  const int k0 = ic - rb;
  for (int q = 0; q < PTC / 4; ++q) {
    const int j = c0 + (warp >> 1) * (PTC / 4) + q;
    if (j >= A.out_col0 + A.out_ncols) break;
    const double cy = A.Y[j];
    double t = 0.0;
    int n = 0, nblk = 0;
    // certain cells: per-column runs from the prefix sums (cells outside the map were staged as t' = 0, not blocked)
    const double* pbase = sP + (size_t)(j - cb) * PS + k0;
    const unsigned short* bbase = sB + (size_t)(j - cb) * PB + k0;
    const bool interior = ic - Lr >= 0 && ic + Lr < A.rows && j - Lr >= max(0, A.in_col0) && j + Lr <= min(A.cols_total, A.in_col0 + A.in_ncols) - 1;
    if (interior) {  // no clipping anywhere: the cell count is the table's
      for (int r = 0; r < Q.nruns; ++r) {
        const int w = __ldg(Q.runs + r);
        const int dj = (int)(signed char)(w & 0xff), lo = (int)(signed char)((w >> 8) & 0xff), hi = (int)(signed char)((w >> 16) & 0xff);
        const double* pc = pbase + dj * PS;
        t += pc[hi + 1] - pc[lo];
        if (tile_any) {
          const unsigned short* bc = bbase + dj * PB;
          nblk += (int)bc[hi + 1] - (int)bc[lo];
        }
      }
      n = Q.ncert;
    } else {
      for (int r = 0; r < Q.nruns; ++r) {
        const int w = __ldg(Q.runs + r);
        const int dj = (int)(signed char)(w & 0xff), lo = (int)(signed char)((w >> 8) & 0xff), hi = (int)(signed char)((w >> 16) & 0xff);
        const int b = j + dj, lb = b - A.in_col0;
        if (b < 0 || b >= A.cols_total || lb < 0 || lb >= A.in_ncols) continue;
        const int a0 = max(ic + lo, 0), a1 = min(ic + hi, A.rows - 1);
        if (a0 > a1) continue;
        const double* pc = pbase + dj * PS;
        t += pc[hi + 1] - pc[lo];
        n += a1 - a0 + 1;
        if (tile_any) {
          const unsigned short* bc = bbase + dj * PB;
          nblk += (int)bc[hi + 1] - (int)bc[lo];
        }
      }
    }
    // uncertain offsets, 32 per pass: lane l decides offset base + l when the decision is the same for every row of the column; the
    // offsets that came out inside and follow each other down a column are then summed as ONE run from the prefix sums
    for (int base = 0; base < Q.nfz; base += 32) {
      const int idx = base + lane;
      int w = 0;
      bool cand = false;
      if (idx < Q.nfz) {
        w = __ldg(Q.fz + idx);
        const int di = (int)(signed char)(w & 0xff), dj = (int)(signed char)((w >> 8) & 0xff);
        if ((w >> 16) & 1) {
          cand = true;  // depends on the row: every lane decides for itself below
        } else {
          const int a = ic + di, b = j + dj;
          // the decision does not depend on the row, so any row's coordinates will do — but they must exist
          const int ar = min(max(a, 0), A.rows - 1), icr = ar - di;
          if (b >= 0 && b < A.cols_total && icr >= 0 && icr < A.rows) cand = poly_inside_d(Q, A.X[icr], cy, A.X[ar], A.Y[b]);
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, cand);
      const unsigned ext = m & __ballot_sync(0xffffffffu, ((w >> 17) & 1) != 0);  // inside AND continues its predecessor down the column
      while (m) {
        const int src = __ffs((int)m) - 1;
        const unsigned tail = src == 31 ? 0u : (ext >> (src + 1));
        const int len = __ffs((int)~tail) - 1;  // further entries of the run (0..31 - src)
        m &= ~((len >= 31 ? 0xffffffffu : ((2u << len) - 1u)) << src);
        const int wv = __shfl_sync(0xffffffffu, w, src);
        const int di = (int)(signed char)(wv & 0xff), dj = (int)(signed char)((wv >> 8) & 0xff);
        const int b = j + dj, lb = b - A.in_col0;
        if (b < 0 || b >= A.cols_total || lb < 0 || lb >= A.in_ncols) continue;
        const int a0 = ic + di, a1 = a0 + len;
        const int a0c = max(a0, 0), a1c = min(a1, A.rows - 1);
        if (a0c > a1c) continue;
        if (((wv >> 16) & 1) && !poly_inside_d(Q, cx, cy, A.X[a0], A.Y[b])) continue;  // row-dependent entries never chain (len == 0)
        const int cc = b - cb;
        const double* pc = sP + (size_t)cc * PS + (a0 - rb);
        t += pc[len + 1] - pc[0];
        n += a1c - a0c + 1;
        if (tile_any) {
          const unsigned short* bc = sB + (size_t)cc * PB + (a0 - rb);
          nblk += (int)bc[len + 1] - (int)bc[0];
        }
      }
    }
    if (!active) continue;
    float result;
    if (nblk > 0) result = 0.0f;                       // :297 / :301
    else if (n == 0) result = (float)A.tdefault;       // :625-628
    else result = (float)(t / (double)n);              // :630
    out[(size_t)(j - A.out_col0) * A.rows + i] = result;
  }
}

// TraversabilityMap::checkCircularFootprintPath (TraversabilityMap.cpp:345-462) for a batch of paths — one thread per path — on a
// traversability_footprint layer that is valid everywhere: every isTraversable(center, ...) takes the memoised branch
// (:667-673), centres outside the map the default branch (:660-666).  No inclination check, no polygons.
__global__ void __launch_bounds__(128) k_check_paths(FpArgs A, const float* __restrict__ fp, const float* __restrict__ rslope, int npaths, const int* __restrict__ path_begin,
                                                     const double* __restrict__ xy, unsigned char* __restrict__ is_safe,
                                                     double* __restrict__ trav_out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npaths) return;
  const int b = path_begin[q], n = path_begin[q + 1] - b;
  is_safe[q] = 0;
  trav_out[q] = 0.0;
  if (n <= 0) return;
  auto circle = [&](double cx, double cy, double& t) -> bool {
    int i, j;
    if (!is_inside_d(A, cx, cy) || !get_index_d(A, cx, cy, i, j)) {
      t = A.tdefault;
      return A.tdefault != 0.0;
    }
    t = (double)lay(A, fp, i, j);
    return t != 0.0;
  };
  // TraversabilityMap::checkInclination (TraversabilityMap.cpp:748-762) on the robot_slope layer; rslope == nullptr: check off
  auto inclination_ok = [&](double ax, double ay, double bx, double by) -> bool {
    if (!rslope) return true;
    int si, sj, ei, ej;
    if (bx == ax && by == ay) {
      if (!is_inside_d(A, ax, ay) || !get_index_d(A, ax, ay, si, sj)) return false;
      return !(lay(A, rslope, si, sj) == 0.0f);
    }
    if (!get_index_d(A, ax, ay, si, sj) || !get_index_d(A, bx, by, ei, ej)) return false;
    const int dx = abs(ei - si), dy = abs(ej - sj);
    int i1x = (ei >= si) ? 1 : -1, i2x = i1x, i1y = (ej >= sj) ? 1 : -1, i2y = i1y;
    int den, num, numAdd, nCells;
    if (dx >= dy) { i1x = 0; i2y = 0; den = dx; num = dx / 2; numAdd = dy; nCells = dx + 1; }
    else { i2x = 0; i1y = 0; den = dy; num = dy / 2; numAdd = dx; nCells = dy + 1; }
    int li = si, lj = sj;
    for (int c = 0; c < nCells; ++c) {
      const float v = lay(A, rslope, li, lj);
      if (finitef(v) && v == 0.0f) return false;
      num += numAdd;
      if (num >= den) { num -= den; li += i1x; lj += i1y; }
      li += i2x; lj += i2y;
    }
    return true;
  };
  double result = 0.0, lengthPath = 0.0;
  double sx = 0.0, sy = 0.0, ex = 0.0, ey = 0.0;
  for (int k = 0; k < n; ++k) {
    sx = ex; sy = ey;
    ex = xy[2 * (b + k)]; ey = xy[2 * (b + k) + 1];
    if (n == 1) {
      if (!inclination_ok(ex, ey, ex, ey)) return;
      double t;
      if (!circle(ex, ey, t)) return;
      result = t;
    }
    if (n > 1 && k > 0) {
      if (!inclination_ok(sx, sy, ex, ey)) return;
      int si, sj, ei, ej;
      if (!get_index_d(A, sx, sy, si, sj) || !get_index_d(A, ex, ey, ei, ej)) return;
      // LineIterator (Bresenham) from the end index to the start index, every fourth cell checked
      const int dx = abs(si - ei), dy = abs(sj - ej);
      int i1x = (si >= ei) ? 1 : -1, i2x = i1x, i1y = (sj >= ej) ? 1 : -1, i2y = i1y;
      int den, num, numAdd, nCells;
      if (dx >= dy) { i1x = 0; i2y = 0; den = dx; num = dx / 2; numAdd = dy; nCells = dx + 1; }
      else { i2x = 0; i1y = 0; den = dy; num = dy / 2; numAdd = dx; nCells = dy + 1; }
      int li = ei, lj = ej, nLine = 0;
      double sum = 0.0;
      for (int c = 0; c < nCells; ++c) {
        if ((c & 3) == 0) {
          double t;
          if (!circle(A.X[li], A.Y[lj], t)) return;
          sum += t;
          ++nLine;
        }
        num += numAdd;
        if (num >= den) { num -= den; li += i1x; lj += i1y; }
        li += i2x; lj += i2y;
      }
      const double t = sum / (double)nLine;
      const double lx = ex - sx, ly = ey - sy;
      const double lengthSegment = sqrt(lx * lx + ly * ly);
      if (k > 1) {
        const double lengthPreviousPath = lengthPath;
        lengthPath += lengthSegment;
        result = (lengthSegment * t + lengthPreviousPath * result) / lengthPath;
      } else {
        lengthPath = lengthSegment;
        result = t;
      }
    }
  }
  is_safe[q] = 1;
  trav_out[q] = result;
}

inline int signum(int v) { return (0 < v) - (v < 0); }

// grid_map::SpiralIterator::generateRing, executed literally.
std::vector<int> build_spiral(double radius, double res) {
  std::vector<int> s;
  const int nRings = (int)std::ceil(radius / res);
  s.push_back(0);
  for (int d = 1; d <= nRings; ++d) {
    std::vector<std::pair<int, int>> ring;
    int px = d, py = 0;
    do {
      ring.emplace_back(px, py);
      const int nx = -signum(py), ny = signum(px);
      if (nx != 0 && (unsigned)std::sqrt((double)((px + nx) * (px + nx) + py * py)) == (unsigned)d) px += nx;
      else if (ny != 0 && (unsigned)std::sqrt((double)(px * px + (py + ny) * (py + ny))) == (unsigned)d) py += ny;
      else { px += nx; py += ny; }
    } while (px != d || py != 0);
    const int edge = (d == nRings || d == nRings - 1) ? 0x10000 : 0;
    for (auto it = ring.rbegin(); it != ring.rend(); ++it) s.push_back((it->first & 0xff) | ((it->second & 0xff) << 8) | edge);
  }
  return s;
}

}  // namespace

void FootprintState::release() {
  if (d_spiral) cudaFree(d_spiral);
  if (d_block) cudaFree(d_block);
  if (d_tables) cudaFree(d_tables);
  if (d_prefix) cudaFree(d_prefix);
  if (d_list) cudaFree(d_list);
  for (int k = 0; k < 2; ++k) {
    if (d_poly[k]) cudaFree(d_poly[k]);
    d_poly[k] = nullptr;
    poly_cap[k] = 0;
  }
  d_spiral = d_block = d_tables = d_prefix = d_list = nullptr;
  spiral_cap = block_cap = tables_cap = prefix_cap = list_cap = 0;
  tables_valid = false;
  valid = false;
}

void launch_check_paths(const SlabView& v, const te_geometry* g, double traversability_default, const float* footprint, const float* robot_slope, int npaths,
                        const int* path_begin, const double* xy, unsigned char* is_safe, double* trav, cudaStream_t s) {
  FpArgs a{};
  a.rows = v.rows; a.cols_total = v.cols_total; a.in_col0 = v.in_col0; a.in_ncols = v.in_ncols;
  a.out_col0 = v.out_col0; a.out_ncols = v.out_ncols;
  a.res = g->resolution; a.lenx = g->length_x; a.leny = g->length_y; a.posx = g->position_x; a.posy = g->position_y;
  a.X = v.X; a.Y = v.Y;
  a.tdefault = traversability_default;
  k_check_paths<<<(npaths + 127) / 128, 128, 0, s>>>(a, footprint, robot_slope, npaths, path_begin, xy, is_safe, trav);
}

int footprint_halo(const te_geometry* g, const te_footprint_params* p) {
  const double res = g->resolution;
  const int spiral = (int)std::ceil((p->radius + p->offset) / res);
  // predicates of a visited cell: slope window 3 cells; step: 2.5-cell circle + 3x3 submap + gap walk
  const int walk = (int)std::ceil(p->max_gap_width / res) + 1;
  return spiral + std::max(4, 3 + 1 + walk);
}

// isTraversableForFilters for every cell of the slab + halo into st.d_block (k_pred_classify + k_pred_heavy); fills the geometry /
// parameter part of the kernel arguments.  Shared by the circular and the polygonal sweep.
int run_predicates(FootprintState& st, const SlabView& v, const te_geometry* g, const te_footprint_params* p, const float* trav,
                   const float* slope, const float* step, const float* rough, const float* elev, float* slope_fp, float* step_fp,
                   float* rough_fp, int sms, cudaStream_t s, FpArgs* out_args) {
  const double rmax = p->radius + p->offset;
  const size_t ncell_in = (size_t)v.rows * v.in_ncols;
  if (st.block_cap < ncell_in) {
    if (st.d_block) cudaFree(st.d_block);
    st.d_block = nullptr;
    st.block_cap = 0;
    if (cudaMalloc(&st.d_block, ncell_in) != cudaSuccess) { st.why = "cudaMalloc(predicate bytes) failed"; return TE_ERR_CUDA; }
    st.block_cap = ncell_in;
  }
  FpArgs a{};
  a.rows = v.rows; a.cols_total = v.cols_total; a.in_col0 = v.in_col0; a.in_ncols = v.in_ncols;
  a.out_col0 = v.out_col0; a.out_ncols = v.out_ncols;
  a.res = g->resolution; a.lenx = g->length_x; a.leny = g->length_y; a.posx = g->position_x; a.posy = g->position_y;
  a.X = v.X; a.Y = v.Y;
  a.rmin = p->radius; a.rmax = rmax; a.rmax2 = rmax * rmax; a.tdefault = p->traversability_default;
  a.maxgap = p->max_gap_width; a.crit = p->critical_step_height; a.int_norm = p->radius_is_integer_norm;
  a.verify_rough = (p->verify_roughness != 0 && rough != nullptr) ? 1 : 0;
  a.n_spiral = st.n_spiral; a.spiral = (const int*)st.d_spiral;
  a.slope_R = (int)std::floor(3.0 * g->resolution / g->resolution) + 1;
  a.step_R = (int)std::floor(2.5 * g->resolution / g->resolution) + 1;
  const Layers L{trav, slope, step, elev, rough};
  const long long t1 = (long long)ncell_in;
  const int g1 = (int)std::min<long long>((t1 + 127) / 128, (long long)sms * 16);
  {
    if (ncell_in >= ((size_t)1 << 32)) { st.why = "slab of 2^32 or more cells"; return TE_ERR_UNSUPPORTED; }
    if (st.list_cap < ncell_in + 1) {
      if (st.d_list) cudaFree(st.d_list);
      st.d_list = nullptr; st.list_cap = 0;
      if (cudaMalloc(&st.d_list, sizeof(unsigned) * (ncell_in + 1)) != cudaSuccess) { st.why = "cudaMalloc(predicate work list) failed"; return TE_ERR_CUDA; }
      st.list_cap = ncell_in + 1;
    }
    unsigned* cnt = (unsigned*)st.d_list;          // word 0: list length; entries follow
    unsigned* lst = cnt + 1;
    cudaMemsetAsync(cnt, 0, sizeof(unsigned), s);
    const dim3 gc((unsigned)((v.rows + 255) / 256), (unsigned)v.in_ncols);
    k_pred_classify<<<gc, 256, 0, s>>>(a, L, (unsigned char*)st.d_block, slope_fp, step_fp, rough_fp, lst, cnt);
    k_pred_heavy<<<std::max(g1, 1), 128, 0, s>>>(a, L, (unsigned char*)st.d_block, slope_fp, step_fp, rough_fp, lst, cnt);
  }
  *out_args = a;
  return 0;
}

int launch_footprint(FootprintState& st, const SlabView& v, const te_geometry* g, const te_footprint_params* p,
                     const std::vector<double>& X, const std::vector<double>& Y, const float* trav, const float* slope,
                     const float* step, const float* rough, const float* elev, float* out, float* slope_fp, float* step_fp,
                     float* rough_fp, int sms, cudaStream_t s, int* launches) {
  (void)X; (void)Y;
  const double rmax = p->radius + p->offset;
  if (std::ceil(rmax / g->resolution) > 120.0) { st.why = "footprint radius exceeds 120 cells"; return TE_ERR_UNSUPPORTED; }
  if (!st.valid || std::memcmp(&st.key_geo, g, sizeof(*g)) != 0 || std::memcmp(&st.key_par, p, sizeof(*p)) != 0) {
    const std::vector<int> sp = build_spiral(rmax, g->resolution);
    if (st.spiral_cap < sp.size() * sizeof(int)) {
      if (st.d_spiral) cudaFree(st.d_spiral);
      st.d_spiral = nullptr;
      st.spiral_cap = 0;
      if (cudaMalloc(&st.d_spiral, sp.size() * sizeof(int)) != cudaSuccess) { st.why = "cudaMalloc(spiral table) failed"; return TE_ERR_CUDA; }
      st.spiral_cap = sp.size() * sizeof(int);
    }
    if (cudaMemcpyAsync(st.d_spiral, sp.data(), sp.size() * sizeof(int), cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { st.why = "spiral table upload failed"; return TE_ERR_CUDA; }
    st.n_spiral = (int)sp.size();
    st.key_geo = *g;
    st.key_par = *p;
    st.valid = true;
    st.tables_valid = false;
  }
  FpArgs a{};
  if (int rc = run_predicates(st, v, g, p, trav, slope, step, rough, elev, slope_fp, step_fp, rough_fp, sms, s, &a)) return rc;
  const size_t ncell_in = (size_t)v.rows * v.in_ncols;
  const Layers L{trav, slope, step, elev, rough};
  const long long t1 = (long long)ncell_in, t2 = (long long)v.rows * v.out_ncols;
  const int g2 = (int)std::min<long long>((t2 + 255) / 256, (long long)sms * 8);
  const int Lmax = (int)std::floor(rmax / g->resolution + 1e-9);
  const bool fast = Lmax <= 31 && v.out_ncols <= 65535 && std::getenv("TE_FOOTPRINT_BRUTE") == nullptr;
  if (!fast) {
    k_sweep<<<std::max(g2, 1), 256, 0, s>>>(a, L, (const unsigned char*)st.d_block, out);
    if (launches) *launches = 3;
    return 0;
  }
  // ---- prefix-sum sweep: tables (cached with the spiral) + per-call prefix/bit arrays -----------------
  if (!st.tables_valid) {
    const double res = g->resolution, r2 = rmax * rmax, tol = 1e-9 * r2 + 1e-300;
    const int nR = (int)std::ceil(rmax / res), Wd = 2 * Lmax + 1;
    std::vector<signed char> halfw(Wd, -1), inner((size_t)(nR + 2) * Wd, -1);
    std::vector<int> fuzzy, ring_start(nR + 2, 0);
    for (int l = -Lmax; l <= Lmax; ++l)
      for (int k = -Lmax - 1; k <= Lmax + 1; ++k) {
        const double d2 = (double)(k * k + l * l) * res * res;
        if (d2 < r2 - tol) halfw[l + Lmax] = (signed char)std::max<int>(halfw[l + Lmax], std::abs(k));
        else if (std::fabs(d2 - r2) <= tol && k >= -Lmax && k <= Lmax) fuzzy.push_back((k & 0xff) | ((l & 0xff) << 8) | 0x10000);
      }
    for (int d = 0; d <= nR + 1; ++d)
      for (int l = -Lmax; l <= Lmax; ++l) {
        int u = -1;
        for (int k = 0; k <= halfw[l + Lmax]; ++k)
          if (k * k + l * l < d * d) u = k;
        inner[(size_t)d * Wd + l + Lmax] = (signed char)u;
      }
    {  // ring boundaries of the spiral table: ring d = entries with floor(|offset|) == d, in table order
      const std::vector<int> sp = build_spiral(rmax, res);
      int idx = 0;
      for (int d = 0; d <= nR; ++d) {
        ring_start[d] = idx;
        while (idx < (int)sp.size()) {
          const int di = (int)(signed char)(sp[idx] & 0xff), dj = (int)(signed char)((sp[idx] >> 8) & 0xff);
          if ((int)std::sqrt((double)(di * di + dj * dj)) != d) break;
          ++idx;
        }
      }
      ring_start[nR + 1] = idx;
    }
    const size_t bytes = halfw.size() + inner.size() + 4 * (fuzzy.size() + 1) + 4 * ring_start.size() + 64;
    if (st.tables_cap < bytes) {
      if (st.d_tables) cudaFree(st.d_tables);
      st.d_tables = nullptr; st.tables_cap = 0;
      if (cudaMalloc(&st.d_tables, bytes) != cudaSuccess) { st.why = "cudaMalloc(footprint tables) failed"; return TE_ERR_CUDA; }
      st.tables_cap = bytes;
    }
    char* base = (char*)st.d_tables;
    size_t off = 0;
    auto put = [&](const void* src, size_t n, size_t align) {
      off = (off + align - 1) / align * align;
      cudaMemcpyAsync(base + off, src, n, cudaMemcpyHostToDevice, s);
      const size_t at = off;
      off += n;
      return at;
    };
    st.off_ring = put(ring_start.data(), 4 * ring_start.size(), 4);
    st.off_fuzzy = put(fuzzy.empty() ? (const void*)ring_start.data() : (const void*)fuzzy.data(), 4 * std::max<size_t>(fuzzy.size(), 1), 4);
    st.off_halfw = put(halfw.data(), halfw.size(), 1);
    st.off_inner = put(inner.data(), inner.size(), 1);
    if (cudaStreamSynchronize(s) != cudaSuccess) { st.why = "footprint table upload failed"; return TE_ERR_CUDA; }
    std::memset(st.h_halfw, -1, sizeof(st.h_halfw));
    std::memcpy(st.h_halfw, halfw.data(), std::min(halfw.size(), sizeof(st.h_halfw)));
    st.n_fuzzy = (int)fuzzy.size();
    st.L = Lmax;
    st.nrings = nR;
    st.tables_valid = true;
  }
  const int words = (v.rows + 31) / 32;
#ifdef TE_CALIBRATION  // calibration builds only: the tiled variant for A/B timing (tools/gpu_r2_*.sh)
  const bool tiled = std::getenv("TE_FOOTPRINT_TILE") != nullptr;
#else
  const bool tiled = false;
#endif
  // global prefix sums (k_sweep_fast) or none (k_sweep_tile builds them per tile in shared memory)
  const size_t pbytes = tiled ? 0 : sizeof(double) * (size_t)(v.rows + 1) * v.in_ncols;
  const size_t wbytes = (sizeof(unsigned) * (size_t)words * v.in_ncols + 15) / 16 * 16;
  const size_t gbytes = ncell_in;
  if (st.prefix_cap < pbytes + wbytes + gbytes) {
    if (st.d_prefix) cudaFree(st.d_prefix);
    st.d_prefix = nullptr; st.prefix_cap = 0;
    if (cudaMalloc(&st.d_prefix, pbytes + wbytes + gbytes) != cudaSuccess) { st.why = "cudaMalloc(footprint prefix sums) failed"; return TE_ERR_CUDA; }
    st.prefix_cap = pbytes + wbytes + gbytes;
  }
  a.L = st.L; a.nrings = st.nrings; a.n_fuzzy = st.n_fuzzy; a.words = words;
  a.ring_start = (const int*)((char*)st.d_tables + st.off_ring);
  a.fuzzy = (const int*)((char*)st.d_tables + st.off_fuzzy);
  a.halfw = (const signed char*)((char*)st.d_tables + st.off_halfw);
  a.inner = (const signed char*)((char*)st.d_tables + st.off_inner);
  a.P = tiled ? nullptr : (const double*)st.d_prefix;
  a.bits = (const unsigned*)((char*)st.d_prefix + pbytes);
  a.near = (const unsigned char*)st.d_prefix + pbytes + wbytes;
  std::memcpy(a.halfw_c, st.h_halfw, sizeof(a.halfw_c));
  const int pstride = tiled ? TR + 2 * st.L + 1 : v.rows + 1;  // pitch of a prefix-sum column
  {
    a.cntp[0] = 0;
    for (int k = 0; k < 64; ++k) {
      const int l = k - st.L, h = (k <= 2 * st.L) ? (int)st.h_halfw[k] : -1;
      a.off_hi[k] = h >= 0 ? l * pstride + h + 1 : 0;
      a.off_lo[k] = h >= 0 ? l * pstride - h : 0;   // a column outside the disk contributes P[0] - P[0]
      const long long bias = (long long)st.L * pstride + st.L;
      a.off8_hi[k] = (unsigned)(8 * (bias + a.off_hi[k]));
      a.off8_lo[k] = (unsigned)(8 * (bias + a.off_lo[k]));
      a.cntp[k + 1] = (short)(a.cntp[k] + (h >= 0 ? 2 * h + 1 : 0));
    }
  }
  const int g3 = std::min(sms * 8, (v.in_ncols + 7) / 8);
  const int g1b = (int)std::min<long long>((t1 + 255) / 256, (long long)sms * 8);
  if (tiled) k_fp_prepare<<<std::max(g3, 1), 256, 0, s>>>(a, (const unsigned char*)st.d_block, (unsigned*)((char*)st.d_prefix + pbytes));
  else k_fp_prepare_p<<<std::max(g3, 1), 256, 0, s>>>(a, L, (const unsigned char*)st.d_block, (double*)st.d_prefix, (unsigned*)((char*)st.d_prefix + pbytes));
  k_fp_nearest<<<std::max(g1b, 1), 256, 0, s>>>(a, (unsigned char*)st.d_prefix + pbytes + wbytes);
  if (tiled) {
    const size_t smem = sizeof(double) * (size_t)(TC + 2 * st.L) * pstride + (size_t)(TC + 2 * st.L) * (TR + 128);
    if (!st.tile_attr) {
      if (cudaFuncSetAttribute(k_sweep_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (TC + 62) * (TR + 63) + (TC + 62) * (TR + 128))) != cudaSuccess) {
        st.why = "cudaFuncSetAttribute(max dynamic shared memory) failed"; return TE_ERR_CUDA;
      }
      st.tile_attr = true;
    }
    k_sweep_tile<<<dim3((unsigned)((v.rows + TR - 1) / TR), (unsigned)((v.out_ncols + TC - 1) / TC)), 256, smem, s>>>(a, L, (const unsigned char*)st.d_block, out);
  } else {
    k_sweep_fast<<<dim3((unsigned)((v.rows + 255) / 256), (unsigned)v.out_ncols), 256, 0, s>>>(a, L, (const unsigned char*)st.d_block, out);
  }
  if (launches) *launches = 5;
  return 0;
}

int polygon_reach(const te_geometry* g, int npts, const double* pts_xy) {
  double r = 0.0;
  for (int k = 0; k < npts; ++k) r = std::max(r, std::hypot(pts_xy[2 * k], pts_xy[2 * k + 1]));
  return (int)std::ceil(r / g->resolution) + 1;
}

int footprint_polygon_halo(const te_geometry* g, const te_footprint_params* p, int npts, const double* pts_xy) {
  const int walk = (int)std::ceil(p->max_gap_width / g->resolution) + 1;
  return polygon_reach(g, npts, pts_xy) + std::max(4, 3 + 1 + walk);
}

namespace {
// Offsets (di, dj) of the cells inside the polygon placed at a cell centre, relative to that centre: certain-in cells as runs per
// column offset, cells within `tol` of a comparison of grid_map::Polygon::isInside as the uncertain list.
struct PolyTables {
  std::vector<int> runs, fz;
};
bool classify_polygon(double res, int Lp, int npts, const double* px, const double* py, const double R[4], PolyTables* out, std::string* why) {
  std::vector<double> vx(npts), vy(npts);
  for (int k = 0; k < npts; ++k) {
    vx[k] = (R[0] * px[k] + R[1] * py[k]) + 0.0;
    vy[k] = (R[2] * px[k] + R[3] * py[k]) + 0.0;
  }
  const double tol = 1e-7 * res;  // rounding moves a comparison by ~1e-13 m at most; anything closer than this is decided per centre
  out->runs.clear();
  out->fz.clear();
  for (int dj = -Lp; dj <= Lp; ++dj) {
    int run_lo = 0;
    bool in_run = false;
    for (int di = -Lp; di <= Lp + 1; ++di) {
      bool inside = false, fuzzy = false, row_dep = false;
      if (di <= Lp) {
        // cell centre relative to the polygon's centre: X decreases with the row index, Y with the column index
        const double ptx = -res * (double)di, pty = -res * (double)dj;
        int cross = 0;
        for (int i = 0, j = npts - 1; i < npts; j = i++) {
          // equal y (bitwise; per centre both get the same centre coordinate added): (v[i].y > pt.y) == (v[j].y > pt.y) always
          if (vy[i] == vy[j]) continue;
          const bool ui = std::fabs(vy[i] - pty) < tol, uj = std::fabs(vy[j] - pty) < tol;
          if (ui || uj) fuzzy = true;  // (v.y > pt.y) may fall either way: depends on the column pair only
          const bool ci = vy[i] > pty, cj = vy[j] > pty;
          if ((ci != cj) || ui || uj) {
            const double thr = (vx[j] - vx[i]) * (pty - vy[i]) / (vy[j] - vy[i]) + vx[i];
            // the x comparison involves the centre's row; so does a threshold that is a quotient of two rounding-sized numbers
            if (std::fabs(ptx - thr) < tol || std::fabs(vy[j] - vy[i]) < 1e3 * tol) { fuzzy = true; row_dep = true; }
            if ((ci != cj) && ptx < thr) ++cross;
          }
        }
        inside = (cross & 1) != 0;
      }
      if (fuzzy) {
        if (out->fz.size() >= 4096) { *why = "footprint polygon has more than 4096 cells on its outline"; return false; }
        int chain = 0;
        if (!row_dep && !out->fz.empty()) {
          const int pw = out->fz.back();
          if ((int)(signed char)((pw >> 8) & 0xff) == dj && (int)(signed char)(pw & 0xff) == di - 1 && ((pw >> 16) & 1) == 0) chain = 1;
        }
        out->fz.push_back((di & 0xff) | ((dj & 0xff) << 8) | ((row_dep ? 1 : 0) << 16) | (chain << 17));
        inside = false;
      }
      if (inside && !in_run) { in_run = true; run_lo = di; }
      if (!inside && in_run) {
        in_run = false;
        out->runs.push_back((dj & 0xff) | ((run_lo & 0xff) << 8) | (((di - 1) & 0xff) << 16));
      }
    }
  }
  return true;
}
}  // namespace

// traversabilityFootprint(yaw): predicates once, then one tiled launch per polygon (unrotated -> out_x, rotated -> out_rot).
int launch_footprint_polygon(FootprintState& st, const SlabView& v, const te_geometry* g, const te_footprint_params* p, int npts,
                             const double* pts_xy, double yaw, const float* trav, const float* slope, const float* step,
                             const float* rough, const float* elev, float* out_x, float* out_rot, int sms, cudaStream_t s, int* launches) {
  if (npts < 3 || npts > PMAXV) { st.why = "footprint polygon needs 3 to 16 vertices"; return TE_ERR_UNSUPPORTED; }
  const int Lp = polygon_reach(g, npts, pts_xy);
  if (Lp > 31) { st.why = "footprint polygon reaches further than 31 cells from its centre"; return TE_ERR_UNSUPPORTED; }
  FpArgs a{};
  if (int rc = run_predicates(st, v, g, p, trav, slope, step, rough, elev, nullptr, nullptr, nullptr, sms, s, &a)) return rc;
  const size_t smem = sizeof(double) * (size_t)(PTC + 2 * Lp) * (PTR + 2 * Lp + 1) + (size_t)(PTC + 2 * Lp) * (PB * 2);
  if (!st.poly_attr) {
    const size_t smax = sizeof(double) * (size_t)(PTC + 62) * (PTR + 63) + (size_t)(PTC + 62) * (PB * 2);
    if (cudaFuncSetAttribute(k_poly_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax) != cudaSuccess) {
      st.why = "cudaFuncSetAttribute(max dynamic shared memory) failed"; return TE_ERR_CUDA;
    }
    st.poly_attr = true;
  }
  int nl = 2;
  for (int which = 0; which < 2; ++which) {
    PolyArgs q{};
    q.Lp = Lp; q.npts = npts;
    const double w = which ? std::cos(0.5 * yaw) : 1.0, z = which ? std::sin(0.5 * yaw) : 0.0;
    const double tz = 2.0 * z, twz = tz * w, tzz = tz * z;
    q.r00 = 1.0 - (0.0 + tzz); q.r01 = 0.0 - twz; q.r10 = 0.0 + twz; q.r11 = 1.0 - (0.0 + tzz);
    for (int k = 0; k < npts; ++k) { q.px[k] = pts_xy[2 * k]; q.py[k] = pts_xy[2 * k + 1]; }
    const double R[4] = {q.r00, q.r01, q.r10, q.r11};
    PolyTables tb;
    if (!classify_polygon(g->resolution, Lp, npts, q.px, q.py, R, &tb, &st.why)) return TE_ERR_UNSUPPORTED;
    const size_t bytes = sizeof(int) * (tb.runs.size() + tb.fz.size() + 2);
    void*& d = st.d_poly[which];
    if (st.poly_cap[which] < bytes) {
      if (d) { cudaStreamSynchronize(s); cudaFree(d); }
      d = nullptr; st.poly_cap[which] = 0;
      if (cudaMalloc(&d, bytes) != cudaSuccess) { st.why = "cudaMalloc(polygon tables) failed"; return TE_ERR_CUDA; }
      st.poly_cap[which] = bytes;
    }
    int* dr = (int*)d;
    int* df = dr + tb.runs.size() + 1;
    if (!tb.runs.empty() && cudaMemcpyAsync(dr, tb.runs.data(), sizeof(int) * tb.runs.size(), cudaMemcpyHostToDevice, s) != cudaSuccess) { st.why = "polygon table upload failed"; return TE_ERR_CUDA; }
    if (!tb.fz.empty() && cudaMemcpyAsync(df, tb.fz.data(), sizeof(int) * tb.fz.size(), cudaMemcpyHostToDevice, s) != cudaSuccess) { st.why = "polygon table upload failed"; return TE_ERR_CUDA; }
    q.runs = dr; q.fz = df; q.nruns = (int)tb.runs.size(); q.nfz = (int)tb.fz.size();
    q.ncert = 0;
    for (int rw : tb.runs) q.ncert += (int)(signed char)((rw >> 16) & 0xff) - (int)(signed char)((rw >> 8) & 0xff) + 1;
    k_poly_tile<<<dim3((unsigned)((v.rows + PTR - 1) / PTR), (unsigned)((v.out_ncols + PTC - 1) / PTC)), 256, smem, s>>>(
        a, q, trav, (const unsigned char*)st.d_block, which ? out_rot : out_x);
  }
  if (launches) *launches = 2 + nl;
  return 0;
}

}  // namespace te
