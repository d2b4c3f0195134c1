// TEST INFRASTRUCTURE (never linked into libebos.so): a serial g++ build of the per-pixel / per-cell functions of
// event_based_bos_b200/csrc/ebos_eklt_math.cuh, walked in the same order as the kernels of ebos_eklt.cu, so that the
// arithmetic the GPU executes can be compared with oracle/spec_eklt.py in the GPU-less build container
// (tests/test_eklt_host_math.py).  The TV term between the forward and the backward is supplied by the caller.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../event_based_bos_b200/csrc/ebos_eklt_math.cuh"

using namespace ebos::eklt;

template <typename T>
static void forward_t(const Geom& g, int flags, const T* theta, const T* gx, const T* gy, const T* weights, T* pf, T* q,
                      T* F, T* trans, double* sums) {
  const int np = g.ph * g.pw;
  for (int k = 0; k < np; ++k) {
    if (flags & kPoisson) sobel_over_8_at(theta, g.ph, g.pw, k / g.pw, k % g.pw, pf[k], pf[np + k]);
    else { pf[k] = theta[k]; pf[np + k] = theta[np + k]; }
  }
  const T* tr = (flags & kWarp) ? theta + flow_channels(flags) * np : nullptr;
  double sq = 0.0, sp = 0.0;
  const int64_t plane = (int64_t)g.H * g.W;
  for (int i = 0; i < g.H; ++i)
    for (int j = 0; j < g.W; ++j) {
      const Pixel<T> p = eval_pixel<T>(g, flags, pf, tr, gx, gy, weights, i, j);
      const int64_t k = (int64_t)i * g.W + j;
      q[k] = p.q;
      F[k] = p.m ? p.f0 : (T)0;
      F[plane + k] = p.m ? p.f1 : (T)0;
      trans[k] = p.t0;
      trans[plane + k] = p.t1;
      sq += (double)p.q * (double)p.q;
      if (p.m) sp += sqrt((double)p.t0 * (double)p.t0 + (double)p.t1 * (double)p.t1);
    }
  sums[0] = sq;
  sums[1] = sp;
}

template <typename T>
static void columns_t(const Geom& g, const T* q, const T* meas, double q2, double w_data, double* colsum, double* scal) {
  const double n = sqrt(q2);
  const T inv = (T)(1.0 / (n + kNormEps));
  for (int j = 0; j < g.W; ++j) colsum[j] = 0.0;
  for (int i = 0; i < g.H; ++i)
    for (int j = 0; j < g.W; ++j) {
      const int64_t k = (int64_t)i * g.W + j;
      const double D = (double)residual<T>(q[k], in_roi(g, i, j), meas[k], inv);
      colsum[j] += D < 0 ? -D : D;
    }
  double mx = -1.0, cnt = 0.0;
  for (int j = 0; j < g.W; ++j) mx = colsum[j] > mx ? colsum[j] : mx;
  for (int j = 0; j < g.W; ++j) cnt += colsum[j] == mx ? 1.0 : 0.0;
  const double tie_w = w_data / cnt;
  double S = 0.0;
  for (int j = g.y0; j < g.y1; ++j) {
    if (colsum[j] != mx) continue;
    for (int i = g.x0; i < g.x1; ++i) {
      const int64_t k = (int64_t)i * g.W + j;
      S += sgn((double)residual<T>(q[k], true, meas[k], inv)) * tie_w * (double)q[k];
    }
  }
  scal[0] = n; scal[1] = mx; scal[2] = tie_w; scal[3] = S;
}

template <typename T>
static void backward_t(const Geom& g, int flags, const T* theta, const T* pf, const T* gx, const T* gy, const T* weights,
                       const T* meas, const T* dF, const double* colsum, const double* scal, double w_pxy, T* dU, T* dPad,
                       T* dP, T* grad, bool stored = false) {
  const T* tr = (flags & kWarp) ? theta + flow_channels(flags) * g.ph * g.pw : nullptr;
  if (!(flags & kWarp)) w_pxy = 0.0;
  BackScalars s;
  s.n = scal[0]; s.mx = scal[1]; s.tie_w = scal[2]; s.S = scal[3];
  const int64_t plane = (int64_t)g.H * g.W;
  const double w_pxy_hw = w_pxy / ((double)g.H * g.W);
  for (int i = 0; i < g.H; ++i)
    for (int j = 0; j < g.W; ++j) {
      Pixel<T> p = eval_pixel<T>(g, flags, pf, tr, gx, gy, weights, i, j);
      const int64_t k = (int64_t)i * g.W + j;
      if (stored) {                    // k_forward packs six planes, k_backward_stored rebuilds the pixel from them
        T pk[6];
        pack_pixel<T>(p, pk);
        p = unpack_pixel<T>(pk, p.q, weights ? weights[k] : (T)1, p.m);
      }
      T out[4];
      backward_pixel<T>(p, flags, meas[k], colsum[j] == s.mx, s, p.m ? dF[k] : (T)0, p.m ? dF[plane + k] : (T)0, w_pxy_hw,
                        out);
      for (int c = 0; c < 4; ++c) dU[c * plane + k] = out[c];
    }
  const int PW = g.pw + 2 * g.pad, PH = g.ph + 2 * g.pad;
  for (int A = 0; A < PH; ++A)
    for (int B = 0; B < PW; ++B) {
      int i0, i1, j0, j1;
      cell_support(A, g.patch, g.h1, g.H, i0, i1);
      cell_support(B, g.patch, g.w1, g.W, j0, j1);
      double acc[4] = {0, 0, 0, 0};
      for (int i = i0; i < i1; ++i)
        for (int j = j0; j < j1; ++j) {
          const double w = (double)cell_weight<T>(A, i, g.h1, g.patch) * (double)cell_weight<T>(B, j, g.w1, g.patch);
          for (int c = 0; c < 4; ++c) acc[c] += w * (double)dU[c * plane + (int64_t)i * g.W + j];
        }
      for (int c = 0; c < 4; ++c) dPad[((int64_t)c * PH + A) * PW + B] = (T)acc[c];
    }
  const int np = g.ph * g.pw;
  for (int k = 0; k < 4 * np; ++k) {
    const int c = k / np, a = (k % np) / g.pw, b = k % g.pw;
    int a0, a1, b0, b1;
    fold_range(a, g.ph, g.pad, a0, a1);
    fold_range(b, g.pw, g.pad, b0, b1);
    double sacc = 0.0;
    for (int A = a0; A < a1; ++A)
      for (int B = b0; B < b1; ++B) sacc += (double)dPad[((int64_t)c * PH + A) * PW + B];
    dP[k] = (T)sacc;
  }
  const int nf = flow_channels(flags);
  for (int k = 0; k < np; ++k) {
    if (flags & kPoisson) {
      grad[k] = sobel_over_8_adjoint_at(dP, dP + np, g.ph, g.pw, k / g.pw, k % g.pw);
    } else {
      grad[k] = dP[k];
      grad[np + k] = dP[np + k];
    }
    if (flags & kWarp) {
      grad[nf * np + k] = dP[2 * np + k];
      grad[(nf + 1) * np + k] = dP[3 * np + k];
    }
  }
}

extern "C" {

// dims = {H, W, ph, pw, patch, x0, x1, y0, y1}; is_f64 selects the element type of every array argument.
int eklt_host_forward(const int* dims, int is_f64, int flags, const void* theta, const void* gx, const void* gy,
                      const void* weights, void* pf, void* q, void* F, void* trans, double* sums) {
  const Geom g = make_geom(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5], dims[6], dims[7], dims[8]);
  if (is_f64)
    forward_t<double>(g, flags, (const double*)theta, (const double*)gx, (const double*)gy, (const double*)weights,
                      (double*)pf, (double*)q, (double*)F, (double*)trans, sums);
  else
    forward_t<float>(g, flags, (const float*)theta, (const float*)gx, (const float*)gy, (const float*)weights, (float*)pf,
                     (float*)q, (float*)F, (float*)trans, sums);
  return 0;
}

int eklt_host_columns(const int* dims, int is_f64, const void* q, const void* meas, double q2, double w_data, double* colsum,
                      double* scal) {
  const Geom g = make_geom(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5], dims[6], dims[7], dims[8]);
  if (is_f64) columns_t<double>(g, (const double*)q, (const double*)meas, q2, w_data, colsum, scal);
  else columns_t<float>(g, (const float*)q, (const float*)meas, q2, w_data, colsum, scal);
  return 0;
}

int eklt_host_backward(const int* dims, int is_f64, int flags, const void* theta, const void* pf, const void* gx,
                       const void* gy, const void* weights, const void* meas, const void* dF, const double* colsum,
                       const double* scal, double w_pxy, void* dU, void* dPad, void* dP, void* grad) {
  const Geom g = make_geom(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5], dims[6], dims[7], dims[8]);
  if (is_f64)
    backward_t<double>(g, flags, (const double*)theta, (const double*)pf, (const double*)gx, (const double*)gy,
                       (const double*)weights, (const double*)meas, (const double*)dF, colsum, scal, w_pxy, (double*)dU,
                       (double*)dPad, (double*)dP, (double*)grad);
  else
    backward_t<float>(g, flags, (const float*)theta, (const float*)pf, (const float*)gx, (const float*)gy,
                      (const float*)weights, (const float*)meas, (const float*)dF, colsum, scal, w_pxy, (float*)dU,
                      (float*)dPad, (float*)dP, (float*)grad);
  return 0;
}

// ebos_sepconv2d walked serially: columns pass into tmp, rows pass into out (is_f64 selects the element type).
int eklt_host_sepconv(const void* image, int H, int W, const double* taps_rows, int n_rows, const double* taps_cols,
                      int n_cols, int border, int is_f64, void* tmp, void* out) {
  ConvTaps tr, tc;
  tr.n = n_rows;
  tc.n = n_cols;
  for (int k = 0; k < kMaxTaps; ++k) {
    tr.w[k] = k < n_rows ? taps_rows[k] : 0.0;
    tc.w[k] = k < n_cols ? taps_cols[k] : 0.0;
  }
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < H; ++i)
      for (int j = 0; j < W; ++j) {
        if (is_f64) {
          const double* in = (const double*)(pass == 0 ? image : tmp);
          double* o = (double*)(pass == 0 ? tmp : out);
          o[(int64_t)i * W + j] = pass == 0 ? correlate_at<double>(in + (int64_t)i * W, 1, j, W, tc, border)
                                            : correlate_at<double>(in + j, W, i, H, tr, border);
        } else {
          const float* in = (const float*)(pass == 0 ? image : tmp);
          float* o = (float*)(pass == 0 ? tmp : out);
          o[(int64_t)i * W + j] = pass == 0 ? correlate_at<float>(in + (int64_t)i * W, 1, j, W, tc, border)
                                            : correlate_at<float>(in + j, W, i, H, tr, border);
        }
      }
  return 0;
}

// k_tv_roi walked serially over the ROI box: value (un-normalised sum of |g w|) and dF = coef * adjoint inside the box
// (dF outside the box is left untouched, as on the GPU).
int eklt_host_tv(const int* dims, int is_f64, const void* F, const void* winv, double coef, void* dF, double* value) {
  const Geom g = make_geom(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5], dims[6], dims[7], dims[8]);
  int r0, r1, c0, c1;
  tv_box(g.x0, g.x1, g.H, r0, r1);
  tv_box(g.y0, g.y1, g.W, c0, c1);
  const int64_t plane = (int64_t)g.H * g.W;
  double total = 0.0;
  for (int c = 0; c < 2; ++c)
    for (int i = r0; i < r1; ++i)
      for (int j = c0; j < c1; ++j) {
        double v, a;
        if (is_f64) {
          tv_pixel<double>((const double*)F + c * plane, (const double*)winv, g.H, g.W, i, j, v, a);
          ((double*)dF)[c * plane + (int64_t)i * g.W + j] = coef * a;
        } else {
          tv_pixel<float>((const float*)F + c * plane, (const float*)winv, g.H, g.W, i, j, v, a);
          ((float*)dF)[c * plane + (int64_t)i * g.W + j] = (float)(coef * a);
        }
        total += v;
      }
  *value = total;
  return 0;
}

// the same through the stored-planes backward (EBOS_EKLT_STORED=1), float64
int eklt_host_backward_stored(const int* dims, int flags, const double* theta, const double* pf, const double* gx,
                              const double* gy, const double* weights, const double* meas, const double* dF,
                              const double* colsum, const double* scal, double w_pxy, double* dU, double* dPad, double* dP,
                              double* grad) {
  const Geom g = make_geom(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5], dims[6], dims[7], dims[8]);
  backward_t<double>(g, flags, theta, pf, gx, gy, weights, meas, dF, colsum, scal, w_pxy, dU, dPad, dP, grad, true);
  return 0;
}

// k_gather_cols_seg + k_gather_rows_thread walked serially, warp by warp and lane group by lane group (float64): the
// shuffle reduction of a group becomes a plain sum over its lanes.  T1 must be zero on entry.
int eklt_host_gather_seg(const int* dims, int nch, const double* dU, double* T1, double* dPad) {
  const Geom g = make_geom(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5], dims[6], dims[7], dims[8]);
  const int PW = g.pw + 2 * g.pad, PH = g.ph + 2 * g.pad, p = g.patch;
  if (p < 2 || p > 32 || 32 % p != 0) return -1;
  const int64_t plane = (int64_t)g.H * g.W;
  const int first = region_offset(g.w1, p) - p;
  for (int i = 0; i < g.H; ++i)
    for (int start = first; start < g.W; start += 32)
      for (int grp = 0; grp < 32 / p; ++grp) {
        double lo_s[4] = {0, 0, 0, 0}, hi_s[4] = {0, 0, 0, 0};
        int A_leader = 0;
        for (int l = 0; l < p; ++l) {
          const int lane = grp * p + l, j = start + lane;
          int A;
          double hi;
          segment_tap<double>(j, g.w1, p, A, hi);
          if (l == 0) A_leader = A;
          else if (A != A_leader) return -2;          // the alignment claim: one floor cell per lane group
          const bool valid = j >= 0 && j < g.W;
          for (int c = 0; c < nch; ++c) {
            const double v = valid ? dU[c * plane + (int64_t)i * g.W + j] : 0.0;
            hi_s[c] += v * hi;
            lo_s[c] += v - v * hi;
          }
        }
        for (int c = 0; c < nch; ++c) {
          double* row = T1 + ((int64_t)c * g.H + i) * PW;
          if (A_leader >= 0 && A_leader < PW) row[A_leader] += lo_s[c];
          if (A_leader + 1 >= 0 && A_leader + 1 < PW) row[A_leader + 1] += hi_s[c];
        }
      }
  for (int o = 0; o < nch * PH * PW; ++o) {
    const int c = o / (PH * PW), A = (o / PW) % PH, B = o % PW;
    int i0, i1;
    cell_support(A, g.patch, g.h1, g.H, i0, i1);
    double s = 0.0;
    for (int i = i0; i < i1; ++i) s += cell_weight<double>(A, i, g.h1, g.patch) * T1[((int64_t)c * g.H + i) * PW + B];
    dPad[((int64_t)c * PH + A) * PW + B] = s;
  }
  return 0;
}

}  // extern "C"
