// te_fused.cu — the fused chain stencil for sm_100a.
//
// One launch computes, for every cell of a column slab, what the reference's six-filter chain
// (robot_filter_parameter.yaml:2-37) computes — normals -> slope, step (both passes), roughness,
// weighted sum — reading `elevation` once and writing the four output layers once; no
// intermediate layer (surface normals, step_height) ever reaches HBM.
//
// Execution model (DESIGN.md §"fused stencil"):
//   * The layer is column-major, row index contiguous.  A WARP owns a strip of 64 rows (two adjacent
//     rows per lane, so all arithmetic is issued as packed f32x2 FFMA2/FADD2) and marches along the
//     column index.  Warps are autonomous: each has its own TMA ring (4 stages x 5 columns x 72 rows,
//     NaN out-of-bounds fill so map borders look like invalid cells), its own mbarriers and a tiny
//     step_height exchange buffer; there is no __syncthreads in the kernel.
//   * Everything a lane needs from columns other than the arriving one lives in registers as
//     five-deep rings indexed at compile time (the march is unrolled by 5 = ring depth = TMA chunk).
//   * Surface normals use the closed form of the 3x3 eigen-problem that holds for a full disk window
//     (scatter matrix [[a,0,p],[0,a,q],[p,q,c]]); moments are accumulated from per-column run sums
//     centred on the run's own middle cell and re-centred on the output cell, so fp32 never sees a
//     cancellation larger than the window's own elevation spread.
//   * fp32 results are CERTIFIED, not trusted: a cell whose window holds an invalid cell (NaN
//     poisoning of the moments / of the NaN-propagating min-max), whose n_z lies too close to a
//     float32 rounding boundary where acos amplifies it, whose scatter matrix is numerically
//     rank-deficient, or whose roughness cancels too far, is appended to a work list and recomputed
//     by the literal double-precision kernel (te_generic.cu: k_fixup_cells).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>

#include "te_fused.h"

namespace te {
namespace {

constexpr int STRIP = 64;          // rows per warp strip
constexpr int EROWS = STRIP + 8;   // staged elevation rows (halo 4 each side)
constexpr int CH = 5;              // columns per TMA chunk = unroll factor = ring depth
constexpr int NST = 4;             // TMA ring stages per warp
constexpr int STAGE_FLOATS = 384;  // EROWS*CH = 360 floats, padded so every stage is 128-byte aligned
constexpr int SHROWS = STRIP + 4;  // step_height rows exchanged per column (halo 2)
constexpr int SHBUF_FLOATS = 80;   // SHROWS padded
constexpr int WARP_SMEM_BYTES = NST * STAGE_FLOATS * 4 + 2 * SHBUF_FLOATS * 4 + 128;  // 6912
constexpr int WARPS_PER_CTA = 12;
constexpr unsigned FULL = 0xffffffffu;

struct FusedArgs {
  int rows, cols_total;
  int in_col0, in_ncols, out_col0, out_ncols;
  int nstrips, nseg, seg_len;
  float a_cov;      // res^2 * K2 / N   (Cxx = Cyy of a full window)
  float kp;         // -res / N         (Cxz = kp * sum k*w)
  float invN;       // 1 / N
  float n_over_nm1; // N / (N-1)
  float slope_crit, inv_slope_crit;
  float step_crit, inv_step_crit, inv_ncrit;
  float rough_crit, inv_rough_crit;
  float fuse_w;
  const unsigned char* rowmask;  // per global row: bit0/1 pass-1 tips (-2,0)/(+2,0); bit2/3 pass-2 tips
  const unsigned char* colmask;  // per global column, same bits for (0,-2)/(0,+2)
  float* slope;
  float* step;
  float* rough;
  float* trav;
  float* nx;
  float* ny;
  float* nz;
  unsigned* list;
  unsigned* count;
  unsigned cap;
};

// ---------------------------------------------------------------------------------------------
// packed f32x2 arithmetic (Blackwell FFMA2/FADD2/FMUL2): .x = row i, .y = row i+1 of the lane
// ---------------------------------------------------------------------------------------------
struct f2 {
  float x, y;
};
__device__ __forceinline__ unsigned long long pk(f2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ f2 up(unsigned long long v) {
  f2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
  return up(r);
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
  return up(r);
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
  return up(r);
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
  return up(r);
}
__device__ __forceinline__ f2 bc(float v) { return f2{v, v}; }

// NaN-propagating three-input min/max (FMNMX3.NAN): an invalid cell poisons the window result.
__device__ __forceinline__ float max3n(float a, float b, float c) {
  float r;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float min3n(float a, float b, float c) {
  float r;
  asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA (per-warp pipelines)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// Window shapes within reach 2 (half-width per |column offset|; -1 = column not in the window).
// TIP = the four offsets (+-2,0),(0,+-2) lie exactly ON the circle and are decided per row/column
// from the double-precision tables (rowmask/colmask), SURVEY.md Appendix C.1.
// ---------------------------------------------------------------------------------------------
template <int WN0_, int WN1_, int WN2_, int W10_, int W11_, int W12_, bool TIP1_, int W20_, int W21_, int W22_, bool TIP2_>
struct Shape {
  static constexpr int WN0 = WN0_, WN1 = WN1_, WN2 = WN2_;
  static constexpr int W10 = W10_, W11 = W11_, W12 = W12_;
  static constexpr bool TIP1 = TIP1_;
  static constexpr int W20 = W20_, W21 = W21_, W22 = W22_;
  static constexpr bool TIP2 = TIP2_;
  static constexpr bool NEED_N1 = (WN0_ == 1 || WN1_ == 1 || WN2_ == 1);
  static constexpr bool NEED_N2 = (WN0_ == 2 || WN1_ == 2 || WN2_ == 2);
  static constexpr bool MASKS = TIP1_ || TIP2_;
};

template <int W>
__device__ __forceinline__ float colmin_w(const float* z, int r) {  // z[r+2] is the centre row
  if constexpr (W == 0) return z[r + 2];
  else if constexpr (W == 1) return min3n(z[r + 1], z[r + 2], z[r + 3]);
  else return min3n(min3n(z[r + 1], z[r + 2], z[r + 3]), z[r], z[r + 4]);
}
template <int W>
__device__ __forceinline__ float colmax_w(const float* z, int r) {
  if constexpr (W == 0) return z[r + 2];
  else if constexpr (W == 1) return max3n(z[r + 1], z[r + 2], z[r + 3]);
  else return max3n(max3n(z[r + 1], z[r + 2], z[r + 3]), z[r], z[r + 4]);
}
template <int W>
__device__ __forceinline__ int colcnt_w(const float* z, int r, float crit) {
  int c = z[r + 2] > crit;
  if constexpr (W >= 1) c += (z[r + 1] > crit) + (z[r + 3] > crit);
  if constexpr (W >= 2) c += (z[r] > crit) + (z[r + 4] > crit);
  return c;
}

struct i2 {
  int x, y;
};

// Per-lane register state.  Rings are indexed by the arrival phase of the column (0..4).
template <class S>
struct Lane {
  f2 e[5];                           // own-row elevation
  f2 a1[5], b1[5], q1[5];            // three-row run sums centred on the run's middle cell
  f2 a2[5], b2[5], q2[5];            // five-row run sums
  f2 c1mn[5], c1mx[5];               // pass-1 column min/max of width W11 (columns l = +-1)
  f2 p1mn[5], p1mx[5];               // pass-1 centre column (width W10 + masked tips)
  f2 sh[5];                          // own-row step_height
  f2 s3mx[5];                        // pass-2 column max of width W21
  f2 pcmx[5];                        // pass-2 centre column
  i2 s3c[5], pcc[5];                 // matching counts of step_height > critical
  f2 dslope[5], drough[5];           // slope / roughness layers waiting for the step layer
  i2 dflag[5];                       // certification flags of the normals stage
};

template <int W, class S>
__device__ __forceinline__ f2 runA(const Lane<S>& L, int s) {
  if constexpr (W == 2) return L.a2[s];
  else if constexpr (W == 1) return L.a1[s];
  else return f2{0.f, 0.f};
}
template <int W, class S>
__device__ __forceinline__ f2 runB(const Lane<S>& L, int s) {
  if constexpr (W == 2) return L.b2[s];
  else if constexpr (W == 1) return L.b1[s];
  else return f2{0.f, 0.f};
}
template <int W, class S>
__device__ __forceinline__ f2 runQ(const Lane<S>& L, int s) {
  if constexpr (W == 2) return L.q2[s];
  else if constexpr (W == 1) return L.q1[s];
  else return f2{0.f, 0.f};
}

__device__ __forceinline__ float rsqrt_nr(float x) {  // rsqrt with one Newton step (rel. error ~1e-7)
  float r = rsqrtf(x);
  return r * fmaf(-0.5f * x, r * r, 1.5f);
}

struct NormalOut {
  float nx, ny, nz, slope, rough;
  int flag;
};

// Closed-form smallest eigenpair of [[a,0,p],[0,a,q],[p,q,c]] + certification (one row).
__device__ __forceinline__ NormalOut finish_normal(const FusedArgs& A, float Sw, float Sk, float Sl, float Sww) {
  NormalOut o;
  const float mw = Sw * A.invN;
  const float c = fmaf(-mw, mw, Sww * A.invN);  // Czz
  const float p = A.kp * Sk, q = A.kp * Sl;       // Cxz, Cyz
  const float g2 = fmaf(p, p, q * q);
  const float h = 0.5f * (A.a_cov - c);
  const float hh = fmaf(h, h, g2);
  float D = hh * rsqrtf(hh);
  D = (hh > 0.f) ? fmaf(0.5f * (hh - D * D), __frcp_rn(D), D) : 0.f;  // sqrt with one correction
  const float dph = D + fabsf(h);
  const float qq = (dph > 0.f) ? __fdividef(g2, dph) : 0.f;
  const float qq2 = (dph > 0.f) ? fmaf(fmaf(-qq, dph, g2), __frcp_rn(dph), qq) : 0.f;  // refined g2/dph
  const bool hpos = h >= 0.f;
  const float m = hpos ? dph : qq2;                       // a - lambda0
  float lam0 = hpos ? (c - qq2) : (A.a_cov - qq2);       // smallest eigenvalue
  const float nn = fmaf(m, m, g2);
  const float rn = rsqrt_nr(nn);
  o.nx = -p * rn;
  o.ny = -q * rn;
  float nz = m * rn;
  int flag = 0;
  // invalid window (NaN/Inf poisoning, overflow) or degenerate pencil
  if (!(fabsf(Sww) < 3.0e38f) || !(nn > 0.f)) flag = 1;
  // small inclination: n_z = 1 - s with s from tan^2; certify the float32 rounding of n_z where
  // acos amplifies one ulp beyond the tolerance (theta < ~0.012 rad)
  const float t = g2 * __frcp_rn(m * m);
  if (hpos && t < 2.5e-3f) {
    const float s = t * fmaf(-t, fmaf(-0.3125f, t, 0.375f), 0.5f);
    nz = 1.0f - s;
    if (s < 7.2e-5f) {
      const float qv = s * 16777216.0f;
      const float fr = qv - floorf(qv);
      // relative error bound of s: first moments carry <= 2.5e-6*sqrt(Sww) absolute error
      const float gm = sqrtf(fmaf(Sk, Sk, Sl * Sl));
      const float eps = fmaf(7.1e-6f, sqrtf(Sww) * __frcp_rn(fmaxf(gm, 1e-30f)), 2e-6f);
      if (fabsf(fr - 0.5f) <= fmaf(qv, eps, 1e-3f) && Sww > 0.f) flag = 1;
    }
  }
  // numerically rank-deficient scatter (exactly planar data): the reference's rank test decides.
  // lambda0 is the difference of two terms of size cmag, each carrying ~1e-6 relative error.
  const float cmag = hpos ? c : A.a_cov;
  if (Sww > 0.f && !(lam0 > 1e-5f * cmag && lam0 > 1e-10f * A.a_cov)) flag = 1;
  lam0 = fmaxf(lam0, 0.f);
  const float r = sqrtf(lam0 * A.n_over_nm1);
  // roughness layer error = d(lambda0) * N/(N-1) / (2 r crit) with d(lambda0) ~ 1.1e-6 cmag; keep it < 3e-6
  if (Sww > 0.f && r * A.rough_crit < 0.2f * cmag) flag = 1;
  // eigenvector conditioning: the gap to the other two eigenvalues is min(2D, m); when it is small
  // against the matrix scale the fp32 moment errors are amplified beyond the tolerance
  if (Sww > 0.f && fminf(2.0f * D, m) < 0.25f * fmaxf(A.a_cov, c)) flag = 1;
  nz = fminf(nz, 1.0f);
  const float theta = acosf(nz);
  o.nz = nz;
  o.slope = theta < A.slope_crit ? fmaf(-theta, A.inv_slope_crit, 1.0f) : 0.0f;
  o.rough = r < A.rough_crit ? fmaf(-r, A.inv_rough_crit, 1.0f) : 0.0f;
  o.flag = flag;
  return o;
}

template <class S>
struct StepCtx {
  const FusedArgs& A;
  const CUtensorMap* map;
  float* ering;     // NST stages
  float* shbuf;     // 2 x SHBUF_FLOATS
  unsigned bar0;    // shared address of the first stage barrier
  int lane;
  int s0;           // first row of the strip
  int q0, q1;       // output columns of the unit
  unsigned rm0, rm1;  // row masks of the lane's two rows
  unsigned rmx;       // row mask of the extra row handled by lanes 0..3
  int xrow;           // strip-relative row (-2,-1,64,65) of the extra row
};

// step_height of one arbitrary strip row at column `js`, straight from the staged elevation
// (used by lanes 0..3 for the four halo rows of the exchange buffer).
template <class S>
__device__ __noinline__ float sh_direct(const StepCtx<S>& C, int t, unsigned kc_stage_base, unsigned cmask) {
  // columns js-2..js+2 are the ones that arrived at steps t-4..t
  float col[5][5];
  const int er = C.xrow + 4;  // row inside the staged window
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    const int tt = t - a;  // arrival step of column js+2-a
    const int chunk = tt / CH, slot = tt - chunk * CH;
    const float* base = C.ering + ((kc_stage_base + chunk) % NST) * STAGE_FLOATS + slot * EROWS + er;
#pragma unroll
    for (int k = 0; k < 5; ++k) col[a][k] = base[k - 2];
  }
  // col[a] : a=0 -> js+2, a=2 -> js, a=4 -> js-2 ; col[a][2] centre row
  const float z0 = col[2][2];
  float mn, mx;
  {
    const float* c = col[2];
    float lo = min3n(c[1], c[2], c[3]), hi = max3n(c[1], c[2], c[3]);
    if constexpr (S::W10 == 2) {
      lo = min3n(lo, c[0], c[4]);
      hi = max3n(hi, c[0], c[4]);
    } else if constexpr (S::TIP1) {
      const float tu = (C.rmx & 1u) ? c[0] : z0, td = (C.rmx & 2u) ? c[4] : z0;
      lo = min3n(lo, tu, td);
      hi = max3n(hi, tu, td);
    }
    mn = lo;
    mx = hi;
  }
  if constexpr (S::W11 >= 0) {
    const float* cl = col[3];
    const float* cr = col[1];
    float l0, l1, h0, h1;
    if constexpr (S::W11 == 0) { l0 = h0 = cl[2]; l1 = h1 = cr[2]; }
    else if constexpr (S::W11 == 1) {
      l0 = min3n(cl[1], cl[2], cl[3]); h0 = max3n(cl[1], cl[2], cl[3]);
      l1 = min3n(cr[1], cr[2], cr[3]); h1 = max3n(cr[1], cr[2], cr[3]);
    } else {
      l0 = min3n(min3n(cl[1], cl[2], cl[3]), cl[0], cl[4]); h0 = max3n(max3n(cl[1], cl[2], cl[3]), cl[0], cl[4]);
      l1 = min3n(min3n(cr[1], cr[2], cr[3]), cr[0], cr[4]); h1 = max3n(max3n(cr[1], cr[2], cr[3]), cr[0], cr[4]);
    }
    mn = min3n(mn, l0, l1);
    mx = max3n(mx, h0, h1);
  }
  if constexpr (S::W12 == 0 || S::TIP1) {
    float tl = col[4][2], tr = col[0][2];
    if constexpr (S::TIP1) {
      tl = (cmask & 1u) ? tl : z0;
      tr = (cmask & 2u) ? tr : z0;
    }
    mn = min3n(mn, tl, tr);
    mx = max3n(mx, tl, tr);
  }
  return mx - mn;
}

// One march step: column ce = q0 - 4 + t arrives.  PH = t % 5.
template <class S, int PH>
__device__ __forceinline__ void march_step(const StepCtx<S>& C, Lane<S>& L, int t, unsigned stage_of_chunk, unsigned kc_stage_base,
                                           unsigned cm_js, unsigned cm_jo) {
  const FusedArgs& A = C.A;
  constexpr int S0 = PH, S1 = (PH + 4) % 5, S2 = (PH + 3) % 5, S3 = (PH + 2) % 5, S4 = (PH + 1) % 5;  // slot of age 0..4
  const int ce = C.q0 - 4 + t;
  // ---- stage A: the arriving elevation column -------------------------------------------------
  float z[6];
  {
    const float* col = C.ering + stage_of_chunk * STAGE_FLOATS + PH * EROWS + 2 + 2 * C.lane;  // staged row 0 is strip row -4
    const float2 v0 = *reinterpret_cast<const float2*>(col);
    const float2 v1 = *reinterpret_cast<const float2*>(col + 2);
    const float2 v2 = *reinterpret_cast<const float2*>(col + 4);
    z[0] = v0.x; z[1] = v0.y; z[2] = v1.x; z[3] = v1.y; z[4] = v2.x; z[5] = v2.y;
  }
  const f2 Z0{z[2], z[3]}, ZM1{z[1], z[2]}, ZP1{z[3], z[4]}, ZM2{z[0], z[1]}, ZP2{z[4], z[5]};
  L.e[S0] = Z0;
  if constexpr (S::NEED_N1 || S::NEED_N2) {
    const f2 D1 = sub2(ZP1, Z0), Dm1 = sub2(ZM1, Z0);
    const f2 A1 = add2(D1, Dm1);
    const f2 B1 = sub2(ZP1, ZM1);
    const f2 Q1 = fma2(D1, D1, mul2(Dm1, Dm1));
    if constexpr (S::NEED_N1) { L.a1[S0] = A1; L.b1[S0] = B1; L.q1[S0] = Q1; }
    if constexpr (S::NEED_N2) {
      const f2 D2 = sub2(ZP2, Z0), Dm2 = sub2(ZM2, Z0);
      L.a2[S0] = add2(A1, add2(D2, Dm2));
      L.b2[S0] = fma2(sub2(ZP2, ZM2), bc(2.0f), B1);
      L.q2[S0] = fma2(D2, D2, fma2(Dm2, Dm2, Q1));
    }
  }
  {
    float cmn[2], cmx[2], pmn[2], pmx[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if constexpr (S::W11 >= 0) { cmn[r] = colmin_w<(S::W11 < 0 ? 0 : S::W11)>(z, r); cmx[r] = colmax_w<(S::W11 < 0 ? 0 : S::W11)>(z, r); }
      else { cmn[r] = cmx[r] = 0.f; }
      float lo = min3n(z[r + 1], z[r + 2], z[r + 3]), hi = max3n(z[r + 1], z[r + 2], z[r + 3]);
      if constexpr (S::W10 == 2) {
        lo = min3n(lo, z[r], z[r + 4]);
        hi = max3n(hi, z[r], z[r + 4]);
      } else if constexpr (S::TIP1) {
        const unsigned rm = r ? C.rm1 : C.rm0;
        const float tu = (rm & 1u) ? z[r] : z[r + 2], td = (rm & 2u) ? z[r + 4] : z[r + 2];
        lo = min3n(lo, tu, td);
        hi = max3n(hi, tu, td);
      }
      pmn[r] = lo;
      pmx[r] = hi;
    }
    L.c1mn[S0] = f2{cmn[0], cmn[1]}; L.c1mx[S0] = f2{cmx[0], cmx[1]};
    L.p1mn[S0] = f2{pmn[0], pmn[1]}; L.p1mx[S0] = f2{pmx[0], pmx[1]};
  }
  if (t < 4) return;  // rings not primed yet
  // ---- stage B: step_height of column js = ce - 2 (ages: js+1 -> 1, js -> 2, js-1 -> 3) ---------
  float* shcol = C.shbuf + (t & 1) * SHBUF_FLOATS;
  {
    float shv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      auto R = [&](const f2& v) { return r ? v.y : v.x; };
      float mn = R(L.p1mn[S2]), mx = R(L.p1mx[S2]);
      if constexpr (S::W11 >= 0) {
        mn = min3n(mn, R(L.c1mn[S1]), R(L.c1mn[S3]));
        mx = max3n(mx, R(L.c1mx[S1]), R(L.c1mx[S3]));
      }
      if constexpr (S::W12 == 0 || S::TIP1) {
        float tl = R(L.e[S4]), tr = R(L.e[S0]);
        if constexpr (S::TIP1) {
          const float z0 = R(L.e[S2]);
          tl = (cm_js & 1u) ? tl : z0;
          tr = (cm_js & 2u) ? tr : z0;
        }
        mn = min3n(mn, tl, tr);
        mx = max3n(mx, tl, tr);
      }
      shv[r] = mx - mn;
    }
    *reinterpret_cast<float2*>(shcol + 2 + 2 * C.lane) = make_float2(shv[0], shv[1]);
    if (C.lane < 4) {
      const float x = sh_direct<S>(C, t, kc_stage_base, cm_js);
      shcol[C.xrow + 2] = x;
    }
  }
  __syncwarp();
  float v[6];
  {
    const float* p = shcol + 2 * C.lane;
    const float2 v0 = *reinterpret_cast<const float2*>(p);
    const float2 v1 = *reinterpret_cast<const float2*>(p + 2);
    const float2 v2 = *reinterpret_cast<const float2*>(p + 4);
    v[0] = v0.x; v[1] = v0.y; v[2] = v1.x; v[3] = v1.y; v[4] = v2.x; v[5] = v2.y;
  }
  {
    float smx[2], pmx[2];
    int sc[2], pc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if constexpr (S::W21 >= 0) {
        smx[r] = colmax_w<(S::W21 < 0 ? 0 : S::W21)>(v, r);
        sc[r] = colcnt_w<(S::W21 < 0 ? 0 : S::W21)>(v, r, A.step_crit);
      } else { smx[r] = 0.f; sc[r] = 0; }
      float hi = max3n(v[r + 1], v[r + 2], v[r + 3]);
      int cnt = (v[r + 1] > A.step_crit) + (v[r + 2] > A.step_crit) + (v[r + 3] > A.step_crit);
      if constexpr (S::W20 == 2) {
        hi = max3n(hi, v[r], v[r + 4]);
        cnt += (v[r] > A.step_crit) + (v[r + 4] > A.step_crit);
      } else if constexpr (S::TIP2) {
        const unsigned rm = r ? C.rm1 : C.rm0;
        const bool iu = rm & 4u, id = rm & 8u;
        hi = max3n(hi, iu ? v[r] : v[r + 2], id ? v[r + 4] : v[r + 2]);
        cnt += (iu && v[r] > A.step_crit) + (id && v[r + 4] > A.step_crit);
      }
      pmx[r] = hi;
      pc[r] = cnt;
    }
    L.sh[S0] = f2{v[2], v[3]};
    L.s3mx[S0] = f2{smx[0], smx[1]}; L.s3c[S0] = i2{sc[0], sc[1]};
    L.pcmx[S0] = f2{pmx[0], pmx[1]}; L.pcc[S0] = i2{pc[0], pc[1]};
  }
  // ---- normals / slope / roughness of column jn = ce - 2 (ages: l = 2 - age) -------------------
  const int jn = ce - 2;
  const int row0 = C.s0 + 2 * C.lane;
  const bool rows_ok = row0 < A.rows;
  if (jn >= C.q0 && jn < C.q1) {
    constexpr int m0 = S::WN2 >= 0 ? 2 * S::WN2 + 1 : 0;  // cells in the columns at l = +-2
    constexpr int m1 = S::WN1 >= 0 ? 2 * S::WN1 + 1 : 0;
    const f2 ec = L.e[S2];
    f2 Sw = runA<S::WN0>(L, S2), Sk = runB<S::WN0>(L, S2), Sww = runQ<S::WN0>(L, S2);
    f2 Sl{0.f, 0.f};
    if constexpr (S::WN1 >= 0) {
      const f2 dR = sub2(L.e[S1], ec), dL = sub2(L.e[S3], ec);
      const f2 aR = runA<S::WN1>(L, S1), aL = runA<S::WN1>(L, S3);
      const f2 tR = fma2(bc((float)m1), dR, aR), tL = fma2(bc((float)m1), dL, aL);
      Sw = add2(Sw, add2(tR, tL));
      Sl = sub2(tR, tL);
      Sk = add2(Sk, add2(runB<S::WN1>(L, S1), runB<S::WN1>(L, S3)));
      Sww = add2(Sww, fma2(dR, add2(aR, tR), runQ<S::WN1>(L, S1)));
      Sww = add2(Sww, fma2(dL, add2(aL, tL), runQ<S::WN1>(L, S3)));
    }
    if constexpr (S::WN2 >= 0) {
      const f2 dR = sub2(L.e[S0], ec), dL = sub2(L.e[S4], ec);
      const f2 aR = runA<S::WN2>(L, S0), aL = runA<S::WN2>(L, S4);
      const f2 tR = fma2(bc((float)m0), dR, aR), tL = fma2(bc((float)m0), dL, aL);
      Sw = add2(Sw, add2(tR, tL));
      Sl = fma2(bc(2.0f), sub2(tR, tL), Sl);
      Sk = add2(Sk, add2(runB<S::WN2>(L, S0), runB<S::WN2>(L, S4)));
      Sww = add2(Sww, fma2(dR, add2(aR, tR), runQ<S::WN2>(L, S0)));
      Sww = add2(Sww, fma2(dL, add2(aL, tL), runQ<S::WN2>(L, S4)));
    }
    // column index grows toward -y and row index toward -x; kp carries the sign and 1/N
    const NormalOut n0 = finish_normal(A, Sw.x, Sk.x, Sl.x, Sww.x);
    const NormalOut n1 = finish_normal(A, Sw.y, Sk.y, Sl.y, Sww.y);
    L.dslope[S0] = f2{n0.slope, n1.slope};
    L.drough[S0] = f2{n0.rough, n1.rough};
    L.dflag[S0] = i2{n0.flag, n1.flag};
    if (A.nx != nullptr && rows_ok) {
      const size_t oc = (size_t)(jn - A.out_col0) * A.rows + row0;
      *reinterpret_cast<float2*>(A.nx + oc) = make_float2(n0.nx, n1.nx);
      *reinterpret_cast<float2*>(A.ny + oc) = make_float2(n0.ny, n1.ny);
      *reinterpret_cast<float2*>(A.nz + oc) = make_float2(n0.nz, n1.nz);
    }
  }
  if (t < 8) return;
  // ---- stage C: step layer of column jo = ce - 4 and the fuse ----------------------------------
  const int jo = ce - 4;
  if (jo >= C.q1) return;
  {
    float outv[2];
    int sflag[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      auto R = [&](const f2& x) { return r ? x.y : x.x; };
      auto RI = [&](const i2& x) { return r ? x.y : x.x; };
      float mx = R(L.pcmx[S2]);
      int cnt = RI(L.pcc[S2]);
      if constexpr (S::W21 >= 0) {
        mx = max3n(mx, R(L.s3mx[S1]), R(L.s3mx[S3]));
        cnt += RI(L.s3c[S1]) + RI(L.s3c[S3]);
      }
      if constexpr (S::W22 == 0 || S::TIP2) {
        float tl = R(L.sh[S4]), tr = R(L.sh[S0]);
        bool il = true, ir = true;
        if constexpr (S::TIP2) {
          il = cm_jo & 4u;
          ir = cm_jo & 8u;
        }
        const float c0 = R(L.sh[S2]);
        mx = max3n(mx, il ? tl : c0, ir ? tr : c0);
        cnt += (il && tl > A.step_crit) + (ir && tr > A.step_crit);
      }
      sflag[r] = !(mx < 3.0e38f);  // NaN/Inf: some window cell was invalid -> literal path decides
      const float stepMax = fmaxf(mx, 0.0f);
      const float st = fminf(stepMax, (float)cnt * A.inv_ncrit * stepMax);
      outv[r] = st < A.step_crit ? fmaf(-st, A.inv_step_crit, 1.0f) : 0.0f;
    }
    const f2 sl = L.dslope[S2], ro = L.drough[S2];
    const i2 nf = L.dflag[S2];
    if (rows_ok) {
      const size_t oc = (size_t)(jo - A.out_col0) * A.rows + row0;
      *reinterpret_cast<float2*>(A.slope + oc) = make_float2(sl.x, sl.y);
      *reinterpret_cast<float2*>(A.rough + oc) = make_float2(ro.x, ro.y);
      *reinterpret_cast<float2*>(A.step + oc) = make_float2(outv[0], outv[1]);
      const float t0 = __fmul_rn(A.fuse_w, __fadd_rn(__fadd_rn(sl.x, outv[0]), ro.x));
      const float t1 = __fmul_rn(A.fuse_w, __fadd_rn(__fadd_rn(sl.y, outv[1]), ro.y));
      *reinterpret_cast<float2*>(A.trav + oc) = make_float2(t0, t1);
    }
    // certified slow path: append flagged cells (bit 30: normals part, bit 31: step part)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int nfl = r ? nf.y : nf.x;
      const unsigned fl = (rows_ok ? ((nfl ? 1u : 0u) | (sflag[r] ? 2u : 0u)) : 0u);
      const unsigned ballot = __ballot_sync(FULL, fl != 0u);
      if (ballot) {
        unsigned base = 0;
        if (C.lane == 0) base = atomicAdd(A.count, (unsigned)__popc(ballot));
        base = __shfl_sync(FULL, base, 0);
        if (fl) {
          const unsigned idx = base + __popc(ballot & ((1u << C.lane) - 1u));
          if (idx < A.cap) A.list[idx] = ((unsigned)(jo - A.out_col0) * (unsigned)A.rows + (unsigned)(row0 + r)) | (fl << 30);
        }
      }
    }
  }
}

template <class S>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 1) k_chain_fused(const __grid_constant__ CUtensorMap map, FusedArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wbase = smem_raw + warp * WARP_SMEM_BYTES;
  float* ering = reinterpret_cast<float*>(wbase);
  float* shbuf = reinterpret_cast<float*>(wbase + NST * STAGE_FLOATS * 4);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(wbase + NST * STAGE_FLOATS * 4 + 2 * SHBUF_FLOATS * 4);
  const unsigned bar0 = smem_u32(bars);
  if (lane == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const int total_warps = gridDim.x * WARPS_PER_CTA;
  const int gwarp = blockIdx.x * WARPS_PER_CTA + warp;
  const int nunits = A.nstrips * A.nseg;
  unsigned kglob = 0;  // chunks consumed so far by this warp (stage = kglob % NST, parity = (kglob / NST) & 1)

  Lane<S> L;
  StepCtx<S> C{A, &map, ering, shbuf, bar0, lane, 0, 0, 0, 0u, 0u, 0u, 0};
  C.xrow = lane < 2 ? lane - 2 : STRIP + lane - 2;

  for (int unit = gwarp; unit < nunits; unit += total_warps) {
    const int strip = unit % A.nstrips, seg = unit / A.nstrips;
    C.s0 = strip * STRIP;
    C.q0 = A.out_col0 + seg * A.seg_len;
    C.q1 = min(C.q0 + A.seg_len, A.out_col0 + A.out_ncols);
    const int nsteps = (C.q1 - C.q0) + 8;
    const int nchunks = (nsteps + CH - 1) / CH;
    if constexpr (S::MASKS) {
      const int r0 = C.s0 + 2 * lane;
      C.rm0 = r0 < A.rows ? A.rowmask[r0] : 0u;
      C.rm1 = r0 + 1 < A.rows ? A.rowmask[r0 + 1] : 0u;
      const int xr = C.s0 + C.xrow;
      C.rmx = (lane < 4 && xr >= 0 && xr < A.rows) ? A.rowmask[xr] : 0u;
    }
    __syncwarp();  // every lane is done with the previous unit's smem
    const unsigned kbase = kglob;
    if (lane == 0) {
      for (int k = 0; k < NST - 2 && k < nchunks; ++k) {
        const unsigned st = (kbase + k) % NST;
        mbar_expect_tx(bar0 + 8 * st, EROWS * CH * 4);
        tma_load_2d(smem_u32(ering + st * STAGE_FLOATS), &map, C.s0 - 4, (C.q0 - 4 + CH * k) - A.in_col0, bar0 + 8 * st);
      }
    }
    // column masks of the first chunk (lane l holds column cbase + l - 8)
    unsigned cm_cur = 0, cm_next = 0;
    if constexpr (S::MASKS) {
      const int c = C.q0 - 4 + lane - 8;
      cm_cur = (lane < 16 && c >= 0 && c < A.cols_total) ? A.colmask[c] : 0u;
    }
    for (int kc = 0; kc < nchunks; ++kc) {
      const unsigned st = (kbase + kc) % NST;
      __syncwarp();  // all lanes finished the previous chunk: its predecessor's stage may be refilled
      if (lane == 0 && kc + NST - 2 < nchunks) {
        const unsigned sn = (kbase + kc + NST - 2) % NST;
        mbar_expect_tx(bar0 + 8 * sn, EROWS * CH * 4);
        tma_load_2d(smem_u32(ering + sn * STAGE_FLOATS), &map, C.s0 - 4, (C.q0 - 4 + CH * (kc + NST - 2)) - A.in_col0,
                    bar0 + 8 * sn);
      }
      if constexpr (S::MASKS) {
        const int c = C.q0 - 4 + CH * (kc + 1) + lane - 8;
        cm_next = (lane < 16 && c >= 0 && c < A.cols_total) ? A.colmask[c] : 0u;
      }
      mbar_wait(bar0 + 8 * st, ((kbase + kc) / NST) & 1u);
      const int t0 = kc * CH;
      unsigned mjs[CH], mjo[CH];
#pragma unroll
      for (int ph = 0; ph < CH; ++ph) {
        if constexpr (S::MASKS) {
          mjs[ph] = __shfl_sync(FULL, cm_cur, ph + 6);  // js = ce - 2
          mjo[ph] = __shfl_sync(FULL, cm_cur, ph + 4);  // jo = ce - 4
        } else {
          mjs[ph] = mjo[ph] = 0u;
        }
      }
      march_step<S, 0>(C, L, t0 + 0, st, kbase, mjs[0], mjo[0]);
      march_step<S, 1>(C, L, t0 + 1, st, kbase, mjs[1], mjo[1]);
      march_step<S, 2>(C, L, t0 + 2, st, kbase, mjs[2], mjo[2]);
      march_step<S, 3>(C, L, t0 + 3, st, kbase, mjs[3], mjo[3]);
      march_step<S, 4>(C, L, t0 + 4, st, kbase, mjs[4], mjo[4]);
      cm_cur = cm_next;
    }
    kglob += nchunks;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct WindowClass {
  int w[3] = {-1, -1, -1};
  bool tips_on_circle = false;  // exactly {(+-2,0),(0,+-2)} undecidable
  bool ok = true;               // representable within reach 2
  int n = 0;                    // certain cells
  int k2 = 0;                   // sum of k^2 over certain cells
};

WindowClass classify(double radius, double res) {
  WindowClass c;
  const double r2 = radius * radius;
  const double tol = 1e-9 * r2 + 1e-300;
  for (int l = -4; l <= 4; ++l)
    for (int k = -4; k <= 4; ++k) {
      const double d2 = (double)(k * k + l * l) * res * res;
      const bool in = d2 < r2 - tol, on = std::fabs(d2 - r2) <= tol;
      if (!in && !on) continue;
      if (std::abs(k) > 2 || std::abs(l) > 2) { c.ok = false; continue; }
      if (on) {
        if ((k == 0 && std::abs(l) == 2) || (l == 0 && std::abs(k) == 2)) c.tips_on_circle = true;
        else c.ok = false;
        continue;
      }
      c.w[std::abs(l)] = std::max(c.w[std::abs(l)], std::abs(k));
      ++c.n;
      c.k2 += k * k;
    }
  return c;
}

using ShapeA = Shape<2, 2, 1, 1, 1, -1, true, 1, 1, -1, true>;    // YAML radii at 0.02 m
using ShapeB = Shape<1, 1, -1, 1, 0, -1, false, 1, 0, -1, false>; // YAML radii at 0.03 m (fixture)

int match_shape(const WindowClass& n, const WindowClass& s1, const WindowClass& s2) {
  auto is = [](const WindowClass& c, int a, int b, int d, bool tip) {
    return c.ok && c.w[0] == a && c.w[1] == b && c.w[2] == d && c.tips_on_circle == tip;
  };
  if (is(n, 2, 2, 1, false) && is(s1, 1, 1, -1, true) && is(s2, 1, 1, -1, true)) return 0;
  if (is(n, 1, 1, -1, false) && is(s1, 1, 0, -1, false) && is(s2, 1, 0, -1, false)) return 1;
  return -1;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

template <class S>
int launch_shape(FusedState& st, const CUtensorMap& map, const FusedArgs& a, int sms, cudaStream_t s) {
  static bool attr_set = false;
  const int smem = WARPS_PER_CTA * WARP_SMEM_BYTES;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_chain_fused<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      st.why = "cudaFuncSetAttribute(max dynamic shared memory) failed";
      return 1;
    }
    attr_set = true;
  }
  const int nunits = a.nstrips * a.nseg;
  int grid = std::min(sms, (nunits + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
  if (grid < 1) grid = 1;
  k_chain_fused<S><<<grid, WARPS_PER_CTA * 32, smem, s>>>(map, a);
  return 0;
}

}  // namespace

void FusedState::release() {
  if (d_rowmask) cudaFree(d_rowmask);
  if (d_colmask) cudaFree(d_colmask);
  d_rowmask = d_colmask = nullptr;
  rowmask_cap = colmask_cap = 0;
  valid = false;
}

bool fused_eligible(FusedState& st, const std::vector<double>& X, const std::vector<double>& Y, const te_geometry* g,
                    const te_chain_params* p) {
  if (st.valid && std::memcmp(&st.key_geo, g, sizeof(*g)) == 0 && std::memcmp(&st.key_par, p, sizeof(*p)) == 0) return st.shape_id >= 0;
  st.valid = false;
  st.shape_id = -1;
  st.key_geo = *g;
  st.key_par = *p;
  auto no = [&](const char* why) {
    st.why = why;
    st.valid = true;
    return false;
  };
  if (p->normals_algorithm != TE_NORMALS_FIXTURE) return no("normals algorithm is not the fixture-pinned one");
  if (p->normals_positive_axis != 2) return no("normal_vector_positive_axis is not z");
  if (g->rows % 4 != 0) return no("row count is not a multiple of 4 (TMA needs 16-byte column pitch)");
  if (!(p->slope_critical > 0.0) || !(p->step_critical > 0.0) || !(p->roughness_critical > 0.0)) return no("a critical value is zero");
  const double res = g->resolution;
  const WindowClass wn = classify(p->normals_radius, res), wr = classify(p->roughness_radius, res);
  const WindowClass w1 = classify(p->step_first_radius, res), w2 = classify(p->step_second_radius, res);
  if (!wn.ok || wn.tips_on_circle) return no("normals window not representable (reach > 2 cells or on-circle offsets)");
  if (!wr.ok || wr.tips_on_circle || std::memcmp(wn.w, wr.w, sizeof(wn.w)) != 0) return no("roughness window differs from the normals window");
  const int id = match_shape(wn, w1, w2);
  if (id < 0) return no("no fused instantiation for these window shapes");
  // on-circle tip membership, decided exactly like CircleIterator does (double, absolute positions)
  const double r1sq = p->step_first_radius * p->step_first_radius, r2sq = p->step_second_radius * p->step_second_radius;
  std::vector<unsigned char> rm(g->rows, 0), cm(g->cols, 0);
  auto bits = [&](const std::vector<double>& P, int i, int n) {
    unsigned b = 0;
    if (i - 2 >= 0) {
      const double d = P[i - 2] - P[i];
      if (d * d + 0.0 <= r1sq) b |= 1u;
      if (d * d + 0.0 <= r2sq) b |= 4u;
    }
    if (i + 2 < n) {
      const double d = P[i + 2] - P[i];
      if (d * d + 0.0 <= r1sq) b |= 2u;
      if (d * d + 0.0 <= r2sq) b |= 8u;
    }
    return (unsigned char)b;
  };
  for (int i = 0; i < g->rows; ++i) rm[i] = bits(X, i, g->rows);
  for (int j = 0; j < g->cols; ++j) cm[j] = bits(Y, j, g->cols);
  auto upload = [&](void*& d, size_t& cap, const std::vector<unsigned char>& h) {
    if (cap < h.size()) {
      if (d) cudaFree(d);
      d = nullptr;
      cap = 0;
      if (cudaMalloc(&d, h.size()) != cudaSuccess) return false;
      cap = h.size();
    }
    return cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice) == cudaSuccess;
  };
  if (!upload(st.d_rowmask, st.rowmask_cap, rm) || !upload(st.d_colmask, st.colmask_cap, cm)) return no("mask table upload failed");
  st.shape_id = id;
  st.valid = true;
  st.why.clear();
  return true;
}

int launch_chain_fused(FusedState& st, const SlabView& v, const ChainDev& p, const float* elev, const ChainOut& o, unsigned* list,
                       unsigned* count, unsigned cap, int sms, cudaStream_t s) {
  if (st.shape_id < 0) { st.why = "fused stencil not eligible"; return 1; }
  if ((reinterpret_cast<uintptr_t>(elev) & 15u) != 0) { st.why = "elevation pointer is not 16-byte aligned"; return 1; }
  if ((size_t)v.rows * v.out_ncols >= ((size_t)1 << 30)) { st.why = "slab has 2^30 or more cells"; return 1; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) { st.why = "cuTensorMapEncodeTiled entry point unavailable"; return 1; }
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)v.rows, (cuuint64_t)v.in_ncols};
  const cuuint64_t strides[1] = {(cuuint64_t)v.rows * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)EROWS, (cuuint32_t)CH};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(elev), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
  if (cr != CUDA_SUCCESS) { st.why = "cuTensorMapEncodeTiled failed (" + std::to_string((int)cr) + ")"; return 1; }

  const double res = v.res;
  const WindowClass wn = classify(p.rn, res);
  FusedArgs a{};
  a.rows = v.rows; a.cols_total = v.cols_total;
  a.in_col0 = v.in_col0; a.in_ncols = v.in_ncols; a.out_col0 = v.out_col0; a.out_ncols = v.out_ncols;
  a.nstrips = (v.rows + STRIP - 1) / STRIP;
  {
    const int total_warps = sms * WARPS_PER_CTA;
    const int nseg0 = std::max(1, (v.out_ncols + 199) / 200);
    const long long units0 = (long long)a.nstrips * nseg0;
    const long long waves = (units0 + total_warps - 1) / total_warps;
    long long nseg = std::max<long long>(1, waves * total_warps / a.nstrips);
    int seg_len = (int)((v.out_ncols + nseg - 1) / nseg);
    if (seg_len < 16) seg_len = std::min(16, v.out_ncols);
    a.seg_len = seg_len;
    a.nseg = (v.out_ncols + seg_len - 1) / seg_len;
  }
  const double N = wn.n, K2 = wn.k2;
  a.a_cov = (float)(res * res * K2 / N);
  a.kp = (float)(-res / N);
  a.invN = (float)(1.0 / N);
  a.n_over_nm1 = (float)(N / (N - 1.0));
  a.slope_crit = (float)p.slope_crit; a.inv_slope_crit = (float)(1.0 / p.slope_crit);
  a.step_crit = (float)p.step_crit; a.inv_step_crit = (float)(1.0 / p.step_crit);
  a.inv_ncrit = (float)(1.0 / (double)p.ncrit);
  a.rough_crit = (float)p.rough_crit; a.inv_rough_crit = (float)(1.0 / p.rough_crit);
  a.fuse_w = p.fuse_w;
  a.rowmask = (const unsigned char*)st.d_rowmask;
  a.colmask = (const unsigned char*)st.d_colmask;
  a.slope = o.slope; a.step = o.step; a.rough = o.rough; a.trav = o.trav;
  a.nx = (o.nx && o.ny && o.nz) ? o.nx : nullptr; a.ny = o.ny; a.nz = o.nz;
  a.list = list; a.count = count; a.cap = cap;
  switch (st.shape_id) {
    case 0: return launch_shape<ShapeA>(st, map, a, sms, s);
    case 1: return launch_shape<ShapeB>(st, map, a, sms, s);
  }
  st.why = "unknown shape id";
  return 1;
}

}  // namespace te
