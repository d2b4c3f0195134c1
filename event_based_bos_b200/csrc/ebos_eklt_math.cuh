// Per-pixel / per-cell arithmetic of the EKLT inner loop, shared by the kernels of ebos_eklt.cu.
//
// Everything here is `__host__ __device__` and free of CUDA-only constructs, so that the very same functions can be
// compiled by g++ into a serial checker (tests/eklt_host_check.cpp) and compared with oracle/spec_eklt.py in the
// GPU-less build container.  The checker is TEST INFRASTRUCTURE: libebos.so contains no host path.
//
// Reference (file:line into tub-rip/event_based_bos):
//   poisson_to_flow                      src/solver/patch_eklt_dependent.py:259-281, src/utils/stat_utils.py:90-92
//   interpolate_dense_flow_from_patch_tensor   src/solver/patch_eklt.py:173-204
//   warp_image_forward                   src/utils/frame_utils.py:56-89
//   _make_prediction_torch / _objective_scipy  src/solver/patch_eklt_pyramid2.py:345-392
//   DifferenceNorm (matrix 1-norm)       src/costs/diff_norm.py:52
//   FlowNorm (pxy)                       src/costs/flow_norm.py:52
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define EK_HD __host__ __device__ __forceinline__
#else
#define EK_HD inline
#endif

namespace ebos {
namespace eklt {

constexpr double kNormEps = 1e-4;   // src/solver/patch_eklt_pyramid2.py:364

// switches of generative_ml.* (mirror EBOS_EKLT_* in include/ebos.h)
constexpr int kPoisson = 1;      // theta[0] is an intensity, patch flow = Sobel(theta[0])/8; else theta[0:2] is the patch flow
constexpr int kWarp = 2;         // the last two channels translate the frame gradients (and carry the pxy term)
constexpr int kNoPolarity = 4;   // q = |q|
EK_HD int flow_channels(int flags) { return (flags & kPoisson) ? 1 : 2; }
EK_HD int theta_channels(int flags) { return flow_channels(flags) + ((flags & kWarp) ? 2 : 0); }

// Separately rounded IEEE operations (never contracted into FMA): used wherever a rounding decides a floor() or the
// sign of a difference of equal-looking numbers (the reference's result depends on both).
template <typename T> struct Ar;
template <> struct Ar<float> {
#if defined(__CUDA_ARCH__)
  static EK_HD float add(float a, float b) { return __fadd_rn(a, b); }
  static EK_HD float sub(float a, float b) { return __fsub_rn(a, b); }
  static EK_HD float mul(float a, float b) { return __fmul_rn(a, b); }
  static EK_HD float div(float a, float b) { return __fdiv_rn(a, b); }
#else
  static EK_HD float add(float a, float b) { volatile float r = a + b; return r; }
  static EK_HD float sub(float a, float b) { volatile float r = a - b; return r; }
  static EK_HD float mul(float a, float b) { volatile float r = a * b; return r; }
  static EK_HD float div(float a, float b) { volatile float r = a / b; return r; }
#endif
};
template <> struct Ar<double> {
#if defined(__CUDA_ARCH__)
  static EK_HD double add(double a, double b) { return __dadd_rn(a, b); }
  static EK_HD double sub(double a, double b) { return __dsub_rn(a, b); }
  static EK_HD double mul(double a, double b) { return __dmul_rn(a, b); }
  static EK_HD double div(double a, double b) { return __ddiv_rn(a, b); }
#else
  static EK_HD double add(double a, double b) { volatile double r = a + b; return r; }
  static EK_HD double sub(double a, double b) { volatile double r = a - b; return r; }
  static EK_HD double mul(double a, double b) { volatile double r = a * b; return r; }
  static EK_HD double div(double a, double b) { volatile double r = a / b; return r; }
#endif
};
// the fp32 base grid of warp_image_forward
EK_HD float f32_div(float a, float b) { return Ar<float>::div(a, b); }
EK_HD float f32_sub(float a, float b) { return Ar<float>::sub(a, b); }

// One pyramid level.  Patch grid [ph,pw] of square patches of `patch` pixels; replicate pad `pad`; the dense image
// ((ph+2pad)*patch rows) is centre-cropped to [H,W] starting at (h1,w1); ROI rows [x0,x1) x cols [y0,y1).
struct Geom {
  int H, W, ph, pw, patch, pad, h1, w1, x0, x1, y0, y1;
  double inv_patch;   // 1/patch when patch is a power of two (x * inv_patch == x / patch bit for bit), else 0
};
EK_HD Geom make_geom(int H, int W, int ph, int pw, int patch, int x0, int x1, int y0, int y1) {
  Geom g;
  g.H = H; g.W = W; g.ph = ph; g.pw = pw; g.patch = patch;
  g.pad = (patch / 2) / patch + 1;                                   // int(patch/2 // sliding) + 1, sliding == patch
  g.h1 = ((ph + 2 * g.pad) * patch) / 2 - H / 2;
  g.w1 = ((pw + 2 * g.pad) * patch) / 2 - W / 2;
  g.x0 = x0; g.x1 = x1; g.y0 = y0; g.y1 = y1;
  g.inv_patch = (patch > 0 && (patch & (patch - 1)) == 0) ? 1.0 / (double)patch : 0.0;
  return g;
}
EK_HD bool in_roi(const Geom& g, int i, int j) { return i >= g.x0 && i < g.x1 && j >= g.y0 && j < g.y1; }
EK_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Bilinear up-sampling taps of one output row/column (align_corners=False):  u = (I + 0.5)/patch - 0.5.
// `a` is the cell of the PADDED axis; lo/hi index the unpadded one (replicate pad = clamp).
template <typename T> struct AxisTap { int a, lo, hi; T fr; };
template <typename T>
EK_HD AxisTap<T> axis_tap(int i, int offset, int patch, int n_patch, int pad, double inv_patch = 0.0) {
  AxisTap<T> t;
  const T num = Ar<T>::add((T)(i + offset), (T)0.5);
  // a power-of-two patch (every level of the pyramid) divides exactly by multiplying with its reciprocal
  const T u = Ar<T>::sub(inv_patch != 0.0 ? Ar<T>::mul(num, (T)inv_patch) : Ar<T>::div(num, (T)patch), (T)0.5);
  const T fl = floor(u);
  t.fr = Ar<T>::sub(u, fl);
  t.a = (int)fl;
  t.lo = clampi(t.a - pad, 0, n_patch - 1);
  t.hi = clampi(t.a + 1 - pad, 0, n_patch - 1);
  return t;
}
// columns first, then rows; each product and sum separately rounded (see the note on sign noise in DESIGN.md)
template <typename T>
EK_HD T upsample_at(const T* P, int pw, const AxisTap<T>& r, const AxisTap<T>& c) {
  const T ncf = Ar<T>::sub((T)1, c.fr), nrf = Ar<T>::sub((T)1, r.fr);
  const T top = Ar<T>::add(Ar<T>::mul(P[r.lo * pw + c.lo], ncf), Ar<T>::mul(P[r.lo * pw + c.hi], c.fr));
  const T bot = Ar<T>::add(Ar<T>::mul(P[r.hi * pw + c.lo], ncf), Ar<T>::mul(P[r.hi * pw + c.hi], c.fr));
  return Ar<T>::add(Ar<T>::mul(top, nrf), Ar<T>::mul(bot, r.fr));
}

// Sobel/8 of the intensity patch grid with replicate padding: out0 differentiates along rows, out1 along columns.
template <typename T>
EK_HD void sobel_over_8_at(const T* P, int ph, int pw, int a, int b, T& out0, T& out1) {
  const int am = clampi(a - 1, 0, ph - 1), ap = clampi(a + 1, 0, ph - 1);
  const int bm = clampi(b - 1, 0, pw - 1), bp = clampi(b + 1, 0, pw - 1);
  const T e00 = P[am * pw + bm], e01 = P[am * pw + b], e02 = P[am * pw + bp];
  const T e10 = P[a * pw + bm], e12 = P[a * pw + bp];
  const T e20 = P[ap * pw + bm], e21 = P[ap * pw + b], e22 = P[ap * pw + bp];
  const T two = (T)2;
  const T dx = Ar<T>::sub(Ar<T>::add(Ar<T>::add(e20, Ar<T>::mul(two, e21)), e22),
                          Ar<T>::add(Ar<T>::add(e00, Ar<T>::mul(two, e01)), e02));
  const T dy = Ar<T>::sub(Ar<T>::add(Ar<T>::add(e02, Ar<T>::mul(two, e12)), e22),
                          Ar<T>::add(Ar<T>::add(e00, Ar<T>::mul(two, e10)), e20));
  out0 = Ar<T>::div(dx, (T)8);
  out1 = Ar<T>::div(dy, (T)8);
}

// warp_image_forward sample position of pixel index `i` along an axis of `size` pixels displaced by `t`:
//   base = float32(i) / float32((size-1)/2) - 1f      (the reference builds the base grid in FLOAT32)
//   g    = base - t / ((size-1)/2) ;  pos = ((g + 1) / 2) * (size - 1)          (float64 from here on upstream)
// Always evaluated in double, also by the float32 instantiation: at zero translation the sample sits ~1e-5 px beside the
// pixel centre and float32 position arithmetic would round it onto the other side, i.e. into another bilinear cell (a
// different one-sided difference in the translation gradient: 3-35 % off in the serial check).
template <typename T>
EK_HD double sample_pos(int i, T t, int size) {
  const double half = (double)(size - 1) / 2.0;
  const float base = f32_sub(f32_div((float)i, (float)half), 1.0f);
  const double g = Ar<double>::sub((double)base, Ar<double>::div((double)t, half));
  return Ar<double>::mul(Ar<double>::mul(Ar<double>::add(g, 1.0), 0.5), (double)(size - 1));   // x / 2 == x * 0.5
}

// Bilinear sample with zeros outside (grid_sample, align_corners=True once positions are in pixels) and the
// derivatives of the sample w.r.t. the sample row / column.  Cell and fractions from the double position; the
// interpolation itself in T.
template <typename T> struct Sample { T v, d_r, d_c; };
template <typename T>
EK_HD Sample<T> bilinear_sample(const T* img, int H, int W, double pr, double pc) {
  Sample<T> s;
  s.v = s.d_r = s.d_c = (T)0;
  // outside (-1, size) every tap is padding; this also keeps the float->int conversions in range (NaN fails both)
  if (!(pr > -1.0 && pr < (double)H && pc > -1.0 && pc < (double)W)) return s;
  const double fr = floor(pr), fc = floor(pc);
  const int r0 = (int)fr, c0 = (int)fc;
  const T a = (T)(pr - fr), b = (T)(pc - fc);
  const bool r0ok = r0 >= 0, r1ok = r0 + 1 < H, c0ok = c0 >= 0, c1ok = c0 + 1 < W;
  const T v00 = (r0ok && c0ok) ? img[(int64_t)r0 * W + c0] : (T)0;
  const T v01 = (r0ok && c1ok) ? img[(int64_t)r0 * W + c0 + 1] : (T)0;
  const T v10 = (r1ok && c0ok) ? img[(int64_t)(r0 + 1) * W + c0] : (T)0;
  const T v11 = (r1ok && c1ok) ? img[(int64_t)(r0 + 1) * W + c0 + 1] : (T)0;
  const T nb = (T)1 - b, na = (T)1 - a;
  s.v = (v00 * nb + v01 * b) * na + (v10 * nb + v11 * b) * a;
  s.d_r = (v10 - v00) * nb + (v11 - v01) * b;
  s.d_c = (v01 - v00) * na + (v11 - v10) * a;
  return s;
}

// Everything the forward needs at one pixel.
template <typename T> struct Pixel {
  T f0, f1, t0, t1;        // up-sampled flow and translation (0 without kWarp)
  Sample<T> sx, sy;        // (warped) frame gradients (row derivative, column derivative) and their position derivatives
  T q0;                    // f0*sx + f1*sy
  T wgt;                   // histogram weight (1 without weights)
  T q;                     // predicted increment before normalisation: (|q0| or q0) * wgt
  bool m;                  // inside the ROI
};
// pf: [2,ph,pw] patch flow (Sobel/8 of the intensity, or theta[0:2]);  tr: [2,ph,pw] patch translation or NULL;
// gx, gy: [H,W];  weights: [H,W] or NULL
template <typename T>
EK_HD Pixel<T> eval_pixel(const Geom& g, int flags, const T* pf, const T* tr, const T* gx, const T* gy, const T* weights,
                          int i, int j) {
  Pixel<T> p;
  const AxisTap<T> r = axis_tap<T>(i, g.h1, g.patch, g.ph, g.pad, g.inv_patch);
  const AxisTap<T> c = axis_tap<T>(j, g.w1, g.patch, g.pw, g.pad, g.inv_patch);
  const int np = g.ph * g.pw;
  const int64_t k = (int64_t)i * g.W + j;
  p.f0 = upsample_at(pf, g.pw, r, c);
  p.f1 = upsample_at(pf + np, g.pw, r, c);
  if (flags & kWarp) {
    p.t0 = upsample_at(tr, g.pw, r, c);
    p.t1 = upsample_at(tr + np, g.pw, r, c);
    const double pr = sample_pos<T>(i, p.t0, g.H), pc = sample_pos<T>(j, p.t1, g.W);
    p.sx = bilinear_sample(gx, g.H, g.W, pr, pc);
    p.sy = bilinear_sample(gy, g.H, g.W, pr, pc);
  } else {                                  // no grid_sample at all upstream: the gradients as they are
    p.t0 = p.t1 = (T)0;
    p.sx.v = gx[k]; p.sx.d_r = p.sx.d_c = (T)0;
    p.sy.v = gy[k]; p.sy.d_r = p.sy.d_c = (T)0;
  }
  p.q0 = Ar<T>::add(Ar<T>::mul(p.f0, p.sx.v), Ar<T>::mul(p.f1, p.sy.v));
  p.q = (flags & kNoPolarity) ? (p.q0 < 0 ? -p.q0 : p.q0) : p.q0;
  p.wgt = weights ? weights[k] : (T)1;
  if (weights) p.q = Ar<T>::mul(p.q, p.wgt);
  p.m = in_roi(g, i, j);
  return p;
}

// |pred - meas| at one pixel; pred = q / (n + eps) inside the ROI, 0 outside.
template <typename T>
EK_HD T residual(T q, bool m, T meas, T inv_norm) { return (m ? q * inv_norm : (T)0) - meas; }

EK_HD double sgn(double v) { return (double)((v > 0) - (v < 0)); }

// Scalars of the backward, produced by the column pass.
struct BackScalars {
  double n;          // ||q||_F
  double mx;         // max column sum (the data term)
  double tie_w;      // w_data / (number of columns attaining the max)   (amax backward spreads evenly)
  double S;          // sum over pixels of g_pred * M * q
};

// Dense gradients at one pixel: d/d f0, f1, t0, t1.  dF0/dF1: TV gradient w.r.t. the masked flow, already scaled by
// w_tv;  w_pxy_hw = w_pxy / (H*W).
template <typename T>
EK_HD void backward_pixel(const Pixel<T>& p, int flags, T meas, bool col_is_max, const BackScalars& s, T dF0, T dF1,
                          double w_pxy_hw, T out[4]) {
  const double inv = 1.0 / (s.n + kNormEps);
  const double D = (double)residual<T>(p.q, p.m, meas, (T)inv);
  const double gm = (p.m && col_is_max) ? sgn(D) * s.tie_w : 0.0;
  double dq = gm * inv;
  if (s.n > 0.0) dq -= ((double)p.q / s.n) * (s.S * inv * inv);
  dq *= (double)p.wgt;                                         // q = (|q0| or q0) * wgt
  if (flags & kNoPolarity) dq *= sgn((double)p.q0);
  double d0 = dq * (double)p.sx.v, d1 = dq * (double)p.sy.v;
  double e0 = -dq * ((double)p.f0 * (double)p.sx.d_r + (double)p.f1 * (double)p.sy.d_r);
  double e1 = -dq * ((double)p.f0 * (double)p.sx.d_c + (double)p.f1 * (double)p.sy.d_c);
  if (p.m) {
    d0 += (double)dF0;
    d1 += (double)dF1;
    if (flags & kWarp) {
      const double tn = sqrt((double)p.t0 * (double)p.t0 + (double)p.t1 * (double)p.t1);
      if (tn > 0.0) {
        e0 += w_pxy_hw * (double)p.t0 / tn;
        e1 += w_pxy_hw * (double)p.t1 / tn;
      }
    }
  }
  out[0] = (T)d0; out[1] = (T)d1; out[2] = (T)e0; out[3] = (T)e1;
}

// ---- stored-planes variant of the backward (EBOS_EKLT_STORED=1, experimental) ---------------------------------------
// The forward keeps what the backward needs of a pixel in six planes instead of having it re-evaluate the up-sampling and
// the eight gathered taps:  sx.v, sy.v, A = f0*sx.d_r + f1*sy.d_r, B = f0*sx.d_c + f1*sy.d_c, t0, t1.
template <typename T>
EK_HD void pack_pixel(const Pixel<T>& p, T out[6]) {
  out[0] = p.sx.v;
  out[1] = p.sy.v;
  out[2] = (T)((double)p.f0 * (double)p.sx.d_r + (double)p.f1 * (double)p.sy.d_r);
  out[3] = (T)((double)p.f0 * (double)p.sx.d_c + (double)p.f1 * (double)p.sy.d_c);
  out[4] = p.t0;
  out[5] = p.t1;
}
// A Pixel that makes backward_pixel reproduce the same expressions from the stored planes (f0 = 1, f1 = 0 turn
// f0*d + f1*d' into the stored combination).  Not for kNoPolarity (the sign of q0 is not stored).
template <typename T>
EK_HD Pixel<T> unpack_pixel(const T s[6], T q, T wgt, bool m) {
  Pixel<T> p;
  p.f0 = (T)1; p.f1 = (T)0;
  p.t0 = s[4]; p.t1 = s[5];
  p.sx.v = s[0]; p.sx.d_r = s[2]; p.sx.d_c = s[3];
  p.sy.v = s[1]; p.sy.d_r = (T)0; p.sy.d_c = (T)0;
  p.q0 = q; p.q = q; p.wgt = wgt; p.m = m;
  return p;
}

// Transposed up-sampling, gather form.  Padded cell A of an axis receives from the dense rows
//   I in [(A-1)*patch + patch/2, (A+1)*patch + patch/2)   with the triangle weight 1 - |u(I) - A|.
EK_HD void cell_support(int A, int patch, int offset, int size, int& i_begin, int& i_end) {
  // ceil/floor for odd patch sizes: u >= A-1  <=>  I >= (A-1)*patch + patch/2 - 0.5
  const int lo = (A - 1) * patch + (patch + 1) / 2 - ((patch & 1) ? 1 : 0);
  const int hi = lo + 2 * patch;
  i_begin = lo - offset < 0 ? 0 : lo - offset;
  i_end = hi - offset > size ? size : hi - offset;
  if (i_end < i_begin) i_end = i_begin;
}
template <typename T>
EK_HD T cell_weight(int A, int i, int offset, int patch) {
  const T u = ((T)(i + offset) + (T)0.5) / (T)patch - (T)0.5;
  const T d = u - (T)A;
  const T w = (T)1 - (d < 0 ? -d : d);
  return w > 0 ? w : (T)0;
}

// ---- segment form of the column pass (EBOS_EKLT_GATHER_SEG=1, experimental; even patch sizes dividing 32) ----------
// Columns whose dense index I = j + w1 satisfies (I - patch/2) mod patch == 0 start a run of `patch` columns with the same
// floor cell A = floor((I - patch/2) / patch) and the up-sampling fractions (r + 0.5) / patch, r = 0 .. patch-1 (both exact
// for even patch sizes).  Warps start at `region_offset - patch` so that lane groups of `patch` lanes ARE those runs.
EK_HD int region_offset(int offset, int patch) {
  const int o = (patch / 2 - offset) % patch;
  return o < 0 ? o + patch : o;
}
EK_HD int floor_div(int n, int d) { return n >= 0 ? n / d : -((-n + d - 1) / d); }
// padded floor cell and hi-weight of column j (lo-weight = 1 - hi)
template <typename T>
EK_HD void segment_tap(int j, int offset, int patch, int& A, T& hi) {
  const int n = j + offset - patch / 2;
  A = floor_div(n, patch);
  hi = ((T)(n - A * patch) + (T)0.5) / (T)patch;
}

// Fold the replicate padding: sum of the padded cells that clamp to unpadded cell a.  [begin,end) in padded indices.
EK_HD void fold_range(int a, int n_patch, int pad, int& begin, int& end) {
  begin = (a == 0) ? 0 : a + pad;
  end = (a == n_patch - 1) ? n_patch + 2 * pad : a + pad + 1;
}

// Adjoint of sobel_over_8_at, gather form: d/dP[r,c] given dF0, dF1 on the patch grid.
template <typename T>
EK_HD T sobel_over_8_adjoint_at(const T* d0, const T* d1, int ph, int pw, int r, int c) {
  const double kx[3][3] = {{-1, -2, -1}, {0, 0, 0}, {1, 2, 1}};
  double acc = 0.0;
  for (int a = r - 1; a <= r + 1; ++a) {
    if (a < 0 || a >= ph) continue;
    for (int b = c - 1; b <= c + 1; ++b) {
      if (b < 0 || b >= pw) continue;
      for (int u = 0; u < 3; ++u) {
        if (clampi(a + u - 1, 0, ph - 1) != r) continue;
        for (int v = 0; v < 3; ++v) {
          if (clampi(b + v - 1, 0, pw - 1) != c) continue;
          acc += kx[u][v] * (double)d0[a * pw + b] + kx[v][u] * (double)d1[a * pw + b];
        }
      }
    }
  }
  return (T)(acc / 8.0);
}


// ---- Adam (torch.optim.Adam defaults; the arithmetic of k_adam in ebos_costs.cu) -----------------------------------
// step_size = lr / (1 - b1^step), inv_bc2_sqrt = 1 / sqrt(1 - b2^step)
template <typename T>
EK_HD void adam_one(T& p, T g, T& m, T& v, T b1, T b2, T eps, T step_size, T inv_bc2_sqrt) {
  m = m * b1 + ((T)1 - b1) * g;
  v = v * b2 + ((T)1 - b2) * g * g;
  const T denom = (T)sqrt((double)v) * inv_bc2_sqrt + eps;
  p -= step_size * (m / denom);
}
// sum of the padded cells that clamp to unpadded cell (a,b) of channel c
template <typename T>
EK_HD T fold_at(const Geom& g, const T* dPad, int c, int a, int b) {
  const int PW = g.pw + 2 * g.pad, PH = g.ph + 2 * g.pad;
  int a0, a1, b0, b1;
  fold_range(a, g.ph, g.pad, a0, a1);
  fold_range(b, g.pw, g.pad, b0, b1);
  double s = 0.0;
  for (int A = a0; A < a1; ++A)
    for (int B = b0; B < b1; ++B) s += (double)dPad[((int64_t)c * PH + A) * PW + B];
  return (T)s;
}

// ---- TV of the masked flow (ImageGradient.calculate_torch, src/costs/image_gradient.py:60-75), gather form --------
// torch.gradient along one axis of n samples: one-sided at both ends, (f[m+1] - f[m-1]) / 2 inside.
template <typename T>
EK_HD T tv_g(const T* line, int64_t stride, int m, int n) {
  if (m == 0) return line[stride] - line[0];
  if (m == n - 1) return line[(int64_t)(n - 1) * stride] - line[(int64_t)(n - 2) * stride];
  return (line[(int64_t)(m + 1) * stride] - line[(int64_t)(m - 1) * stride]) / (T)2;
}
// |g(m) w(m)| and s(m) = sign(g(m) w(m)) w(m)
template <typename T>
EK_HD double tv_abs(const T* line, const T* wline, int64_t stride, int m, int n) {
  const double gw = (double)(tv_g(line, stride, m, n) * wline[(int64_t)m * stride]);
  return gw < 0 ? -gw : gw;
}
template <typename T>
EK_HD double tv_s(const T* line, const T* wline, int64_t stride, int m, int n) {
  const T w = wline[(int64_t)m * stride];
  return sgn((double)(tv_g(line, stride, m, n) * w)) * (double)w;
}
// d/d f[k] of sum_m |g(m) w(m)| : the adjoint of torch.gradient applied to s
template <typename T>
EK_HD double tv_adjoint(const T* line, const T* wline, int64_t stride, int k, int n) {
  double a = 0.0;
  if (k - 1 >= 1 && k - 1 <= n - 2) a += 0.5 * tv_s(line, wline, stride, k - 1, n);
  if (k + 1 >= 1 && k + 1 <= n - 2) a -= 0.5 * tv_s(line, wline, stride, k + 1, n);
  if (k == 1) a += tv_s(line, wline, stride, 0, n);
  if (k == 0) a -= tv_s(line, wline, stride, 0, n);
  if (k == n - 1) a += tv_s(line, wline, stride, n - 1, n);
  if (k == n - 2) a -= tv_s(line, wline, stride, n - 1, n);
  return a;
}
// Rows/columns outside [lo, hi) cannot see the ROI through the 5-point support of the adjoint: F = f*M vanishes there
// together with every difference that touches it.
EK_HD void tv_box(int a0, int a1, int n, int& lo, int& hi) {
  lo = a0 - 2 < 0 ? 0 : a0 - 2;
  hi = a1 + 2 > n ? n : a1 + 2;
}
// value and gradient contributions of pixel (i,j), channel plane F (one of the two), weights winv
template <typename T>
EK_HD void tv_pixel(const T* F, const T* winv, int H, int W, int i, int j, double& value, double& adjoint) {
  const T* row = F + (int64_t)i * W;      // along columns: stride 1
  const T* wrow = winv + (int64_t)i * W;
  const T* col = F + j;                   // along rows: stride W
  const T* wcol = winv + j;
  value = tv_abs(col, wcol, (int64_t)W, i, H) + tv_abs(row, wrow, (int64_t)1, j, W);
  adjoint = tv_adjoint(col, wcol, (int64_t)W, i, H) + tv_adjoint(row, wrow, (int64_t)1, j, W);
}

// ---- separable correlation with mirrored borders (per-window preprocessing) --------------------------------------
// border 0: reflect-101  (cv2.BORDER_REFLECT_101, the default of cv2.Sobel / cv2.GaussianBlur:  c b | a b c | b a)
// border 1: reflect      (scipy.ndimage mode='reflect':                                         b a | a b c | c b)
constexpr int kMaxTaps = 127;
struct ConvTaps {
  double w[kMaxTaps];
  int n;
};
EK_HD int border_index(int i, int n, int border) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (border == 0) i = i < 0 ? -i : 2 * (n - 1) - i;
    else i = i < 0 ? -i - 1 : 2 * n - 1 - i;
  }
  return i;
}
// out = sum_k w[k] * in[border(i + k - (n-1)/2)] along one axis (stride between consecutive samples of that axis)
template <typename T>
EK_HD T correlate_at(const T* line, int64_t stride, int i, int n, const ConvTaps& t, int border) {
  double acc = 0.0;
  const int r = (t.n - 1) / 2;
  for (int k = 0; k < t.n; ++k) acc += t.w[k] * (double)line[(int64_t)border_index(i + k - r, n, border) * stride];
  return (T)acc;
}

}  // namespace eklt
}  // namespace ebos
