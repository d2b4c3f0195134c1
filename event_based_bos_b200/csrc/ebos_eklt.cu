// EKLT inner loop of PatchEkltPyramid2: value and gradient of the per-level objective
//
//   L(theta) = w_d * ||pred - meas||_1(matrix) + w_tv * TV(flow*M; w_inv) + w_p * mean ||pxy*M||_2
//
// for theta [3,ph,pw] = (intensity, p_row, p_col) on the coarse patch grid  (src/solver/patch_eklt_pyramid2.py:345-392).
// Per-pixel arithmetic lives in ebos_eklt_math.cuh (host+device, checked on the CPU against oracle/spec_eklt.py);
// this file holds the parallel structure.  One evaluation =
//
//   memset(acc, colsum, colS)
//   k_patch_flow     pf = Sobel(theta[0]) / 8                              [2,ph,pw]       (tiny; poisson model only)
//   k_forward        per pixel: up-sample pf and theta[1:3], warp the frame gradients, q; writes q and the masked
//                    flow F = f*M; block-reduces sum q^2 and sum ||t*M|| (16 spread accumulator slots)
//   k_tv_roi         TV(F; w_inv) value and w_tv * dTV/dF in gather form over the ROI box +-2
//                    (legacy chain: ebos_flow_tv over the whole image -- 39 us in fp64 on B200, r01h launch list)
//   k_column_sums    colsum[j] = sum_i |pred - meas|_ij, colS[j] = sum_i sign(.) q_ij     (needs ||q||)
//   k_column_max     data term = max_j colsum, tie count, S = sum of colS over the arg-max columns   (one CTA)
//   k_backward       per pixel: re-evaluates the forward, writes d/d(f0,f1,t0,t1)  [4,H,W]
//   k_gather_cols/rows  transposed up-sampling as two separable passes (legacy: one CTA / warp per padded cell over its
//                    2-D support, which reads every gradient plane four times: 33-55 us)
//   k_fold           folds the replicate padding                                    [4,ph,pw]
//   k_param_grad     Sobel adjoint for the intensity channel, copies the translation channels, writes the loss
//
// All plane passes are HBM/L2 streaming work (no dense contraction: tensor cores unused).  Algorithmic bytes per
// evaluation (P = H*W*sizeof(T)): forward 2P (gradients) + 3P (q, F) ; TV 3P + 2P ; columns 2P ; backward 2P + 2P + P
// + 4P ; gather 4P  =>  25P (the figure bench.py credits; the ROI-restricted TV moves less).
#include <algorithm>
#include <cstdlib>

#include "ebos_common.cuh"
#include "ebos_eklt_math.cuh"

// Build-time knob (python -m event_based_bos_b200._build with EBOS_BUILD_EKLT_MINB=n): minimum resident CTAs per SM of the
// three per-pixel plane kernels.  They are latency-bound at 33-46 % of the warp slots (62-80 registers; ncu r01i);
// 4 caps them at 64 registers.  Default 1 = no cap (the measured configuration).
#ifndef EBOS_EKLT_MINB
#define EBOS_EKLT_MINB 1
#endif


// Programmatic dependent launch along the evaluation chain: ten short kernels, each waiting for its predecessor at its
// very top, but scheduled (and through its prologue) while the predecessor drains -- the launch gaps between one-wave
// kernels were a fifth of an evaluation.  EBOS_NO_PDL=1 disables the launch attribute (the prologue is then a no-op).
#define EKLT_PDL_PROLOGUE() do { ebos::pdl_launch_dependents(); ebos::pdl_wait(); } while (0)
#define EKLT_LAUNCH(kern, grid, block, st, ...)                                                       \
  do {                                                                                                \
    cudaError_t le__ = ebos::launch_pdl(kern, dim3(grid), dim3(block), st, __VA_ARGS__);              \
    if (le__ != cudaSuccess) return ebos::cuda_fail(le__, "ebos_eklt launch");                        \
  } while (0)

namespace ebos {
namespace eklt {

// acc (double) layout behind the TV accumulators of ebos_flow_tv.  Sums fed by one atomic per CTA are spread over 16
// slots each (same-address double atomics from ~1000 CTAs serialise in the L2 atomic unit, ~6 ns apiece).
constexpr int kAccMax = 0, kAccTieW = 1, kAccS = 2, kAccLoss = 3, kAccData = 4, kAccTv = 5, kAccPxyMean = 6;
constexpr int kSpread = 16, kAccQ2 = 16, kAccPxy = 32, kAccTvSum = 48, kAccN = 64;
constexpr int kAccTicketCols = 8, kAccTicketTail = 9;   // free slots reinterpreted as unsigned "CTAs done" counters
__device__ __forceinline__ double acc_sum(const double* __restrict__ acc, int base) {
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < kSpread; ++i) t += acc[base + i];
  return t;
}
__device__ __forceinline__ int acc_slot() { return (blockIdx.x + 7 * blockIdx.y) & (kSpread - 1); }

// EBOS_EKLT_LEGACY=1: the first (B200-validated) chain -- TV through ebos_flow_tv over the whole image, CTA/warp-per-cell
// gather -- for A/B runs.
static bool legacy_chain() {
  const char* v = getenv("EBOS_EKLT_LEGACY");      // read per call: bench.py times both chains in one process
  return v != nullptr && v[0] != '\0' && v[0] != '0';
}

struct Workspace {
  double* tv_acc;    // [EBOS_ACC_DOUBLES]
  double* acc;       // [kAccN]
  double* colsum;    // [W]
  double* colS;      // [W]  sum_i sign(D) * q over the ROI rows of each column
  char* T1;          // [4,H,pw+2pad] T: column pass of the transposed up-sampling
  char* St;          // [6,H,W] T: stored forward planes (stored-planes backward)
  char* pf;          // [2,ph,pw] T
  char* q;           // [H,W] T
  char* F;           // [2,H,W] T
  char* dF;          // [2,H,W] T
  char* dU;          // [4,H,W] T
  char* dPad;        // [4,ph+2pad,pw+2pad] T
  char* dP;          // [4,ph,pw] T
  size_t total;
};
static Workspace carve(void* base, int H, int W, int ph, int pw, int pad, size_t elem) {
  Workspace w;
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align256(bytes); return r; };
  w.tv_acc = reinterpret_cast<double*>(take(EBOS_ACC_DOUBLES * sizeof(double)));
  // acc and colsum are adjacent: one memset clears both
  w.acc = reinterpret_cast<double*>(p ? p + off : nullptr);
  off += kAccN * sizeof(double);
  w.colsum = reinterpret_cast<double*>(p ? p + off : nullptr);
  off += (size_t)W * sizeof(double);
  w.colS = reinterpret_cast<double*>(p ? p + off : nullptr);
  off += (size_t)W * sizeof(double);
  off = align256(off);
  const size_t plane = (size_t)H * W * elem, cells = (size_t)ph * pw * elem;
  w.pf = take(2 * cells);
  w.q = take(plane);
  w.F = take(2 * plane);
  w.dF = take(2 * plane);
  w.dU = take(4 * plane);
  w.dPad = take((size_t)4 * (ph + 2 * pad) * (pw + 2 * pad) * elem);
  w.dP = take(4 * cells);
  w.T1 = take((size_t)4 * H * (pw + 2 * pad) * elem);
  w.St = take(6 * plane);
  w.total = off;
  return w;
}

template <typename T>
__global__ void k_patch_flow(const T* __restrict__ theta, int ph, int pw, T* __restrict__ pf) {
  EKLT_PDL_PROLOGUE();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ph * pw) return;
  T o0, o1;
  sobel_over_8_at(theta, ph, pw, k / pw, k % pw, o0, o1);
  pf[k] = o0;
  pf[ph * pw + k] = o1;
}

// 2-D tiles of 32 x 8 pixels; gridDim.y strides the rows.
template <typename T>
__global__ void __launch_bounds__(256, EBOS_EKLT_MINB) k_forward(Geom g, int flags, const T* __restrict__ pf, const T* __restrict__ tr,
                                                 const T* __restrict__ gx, const T* __restrict__ gy,
                                                 const T* __restrict__ weights, T* __restrict__ q, T* __restrict__ F,
                                                 double* __restrict__ acc, T* __restrict__ St) {
  EKLT_PDL_PROLOGUE();
  __shared__ double red[32];
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  double sq = 0.0, sp = 0.0;
  if (j < g.W) {
    for (int i = blockIdx.y * 8 + (threadIdx.x >> 5); i < g.H; i += gridDim.y * 8) {
      const Pixel<T> p = eval_pixel<T>(g, flags, pf, tr, gx, gy, weights, i, j);
      const int64_t k = (int64_t)i * g.W + j;
      q[k] = p.q;
      F[k] = p.m ? p.f0 : (T)0;
      F[(int64_t)g.H * g.W + k] = p.m ? p.f1 : (T)0;
      if (St) {                                   // uniform: stored-planes backward
        T pk[6];
        pack_pixel<T>(p, pk);
#pragma unroll
        for (int c = 0; c < 6; ++c) St[c * ((int64_t)g.H * g.W) + k] = pk[c];
      }
      sq += (double)p.q * (double)p.q;
      if (p.m) sp += sqrt((double)p.t0 * (double)p.t0 + (double)p.t1 * (double)p.t1);
    }
  }
  sq = block_sum(sq, red);
  sp = block_sum(sp, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc + kAccQ2 + acc_slot(), sq);
    atomicAdd(acc + kAccPxy + acc_slot(), sp);
  }
}

__device__ __forceinline__ void column_max_body(int W, int y0, int y1, const double* __restrict__ colsum,
                                                const double* __restrict__ colS, double* __restrict__ acc,
                                                double w_data, double* red, double* s_mx);

template <typename T>
__global__ void __launch_bounds__(256) k_column_sums(Geom g, const T* __restrict__ q, const T* __restrict__ meas,
                                                     double* __restrict__ acc, double* __restrict__ colsum,
                                                     double* __restrict__ colS, int fuse_max, double w_data) {
  EKLT_PDL_PROLOGUE();
  __shared__ double part[8][33];
  __shared__ double partS[8][33];
  __shared__ double red[32];
  __shared__ double s_mx;
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const T inv = (T)(1.0 / (sqrt(acc_sum(acc, kAccQ2)) + kNormEps));
  double s = 0.0, sS = 0.0;
  if (j < g.W) {
    for (int i = blockIdx.y * 8 + row; i < g.H; i += gridDim.y * 8) {
      const int64_t k = (int64_t)i * g.W + j;
      const bool m = in_roi(g, i, j);
      const T qk = q[k];
      const double D = (double)residual<T>(qk, m, meas[k], inv);
      s += fabs(D);
      if (m) sS += sgn(D) * (double)qk;
    }
  }
  part[row][lane] = s;
  partS[row][lane] = sS;
  __syncthreads();
  if (row == 0 && j < g.W) {
    double t = 0.0, tS = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) { t += part[r][lane]; tS += partS[r][lane]; }
    atomicAdd(colsum + j, t);
    atomicAdd(colS + j, tS);
  }
  if (fuse_max) {
    // the CTA that finishes last (ticket in acc[kAccTicketCols], zeroed with the accumulators) also takes the maximum
    // over the columns: one launch and one single-CTA kernel less on the critical path of every evaluation
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(acc + kAccTicketCols), 1u);
      s_last = t == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      column_max_body(g.W, g.y0, g.y1, colsum, colS, acc, w_data, red, &s_mx);
    }
  }
}

// One CTA: max column sum (the data term), number of ties, S = tie_w * sum over the maximal ROI columns of colS.
// (colsum / colS are read through L2: the kernel that accumulated them with atomics may be this very launch)
__device__ __forceinline__ void column_max_body(int W, int y0, int y1, const double* __restrict__ colsum,
                                                const double* __restrict__ colS, double* __restrict__ acc,
                                                double w_data, double* red, double* s_mx) {
  double mx = -1.0;
  int nan = 0;
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    const double c = __ldcg(colsum + j);
    nan |= (c != c);
    mx = fmax(mx, c);                            // fmax drops NaN: tracked separately
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  const int any_nan = __syncthreads_or(nan);
  if (threadIdx.x == 0) {
    double m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
    *s_mx = any_nan ? __longlong_as_double(0x7ff8000000000000LL) : m;     // a NaN objective stays NaN, like torch's amax
  }
  __syncthreads();
  mx = *s_mx;
  double cnt = 0.0, S = 0.0;
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    if (__ldcg(colsum + j) == mx) {
      cnt += 1.0;
      if (j >= y0 && j < y1) S += __ldcg(colS + j);
    }
  }
  cnt = block_sum(cnt, red);
  S = block_sum(S, red);
  if (threadIdx.x == 0) {
    const double tie_w = w_data / cnt;
    acc[kAccMax] = mx;
    acc[kAccTieW] = tie_w;
    acc[kAccS] = S * tie_w;
  }
}

__global__ void __launch_bounds__(256) k_column_max(int W, int y0, int y1, const double* __restrict__ colsum,
                                                    const double* __restrict__ colS, double* __restrict__ acc,
                                                    double w_data) {
  EKLT_PDL_PROLOGUE();
  __shared__ double red[32];
  __shared__ double s_mx;
  column_max_body(W, y0, y1, colsum, colS, acc, w_data, red, &s_mx);
}

template <typename T>
__global__ void __launch_bounds__(256, EBOS_EKLT_MINB) k_backward(Geom g, int flags, const T* __restrict__ pf, const T* __restrict__ tr,
                                                  const T* __restrict__ gx, const T* __restrict__ gy,
                                                  const T* __restrict__ weights, const T* __restrict__ meas,
                                                  const T* __restrict__ dF,
                                                  const double* __restrict__ colsum, const double* __restrict__ acc,
                                                  double w_pxy_hw, T* __restrict__ dU) {
  EKLT_PDL_PROLOGUE();
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  if (j >= g.W) return;
  BackScalars s;
  s.n = sqrt(acc_sum(acc, kAccQ2));
  s.mx = acc[kAccMax];
  s.tie_w = acc[kAccTieW];
  s.S = acc[kAccS];
  const bool col_is_max = colsum[j] == s.mx;
  const int64_t plane = (int64_t)g.H * g.W;
  for (int i = blockIdx.y * 8 + (threadIdx.x >> 5); i < g.H; i += gridDim.y * 8) {
    const Pixel<T> p = eval_pixel<T>(g, flags, pf, tr, gx, gy, weights, i, j);
    const int64_t k = (int64_t)i * g.W + j;
    T out[4];
    backward_pixel<T>(p, flags, meas[k], col_is_max, s, p.m ? dF[k] : (T)0, p.m ? dF[plane + k] : (T)0, w_pxy_hw, out);
    dU[k] = out[0];
    dU[plane + k] = out[1];
    if (flags & kWarp) {
      dU[2 * plane + k] = out[2];
      dU[3 * plane + k] = out[3];
    }
  }
}

// Default since round 2 (EBOS_EKLT_STORED=0 selects the re-evaluating k_backward): the backward from six planes stored by the
// forward instead of re-evaluating up-sampling, sample positions and the eight gathered taps (530 instructions per pixel in
// k_backward, ncu r01i; DRAM at 9 % -- instructions are the scarce resource here, bytes are not).
template <typename T>
__global__ void __launch_bounds__(256) k_backward_stored(Geom g, int flags, const T* __restrict__ St, const T* __restrict__ q,
                                                         const T* __restrict__ weights, const T* __restrict__ meas,
                                                         const T* __restrict__ dF, const double* __restrict__ colsum,
                                                         const double* __restrict__ acc, double w_pxy_hw,
                                                         T* __restrict__ dU) {
  EKLT_PDL_PROLOGUE();
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  if (j >= g.W) return;
  BackScalars s;
  s.n = sqrt(acc_sum(acc, kAccQ2));
  s.mx = acc[kAccMax];
  s.tie_w = acc[kAccTieW];
  s.S = acc[kAccS];
  const bool col_is_max = colsum[j] == s.mx;
  const int64_t plane = (int64_t)g.H * g.W;
  for (int i = blockIdx.y * 8 + (threadIdx.x >> 5); i < g.H; i += gridDim.y * 8) {
    const int64_t k = (int64_t)i * g.W + j;
    T pk[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) pk[c] = St[c * plane + k];
    const bool m = in_roi(g, i, j);
    const Pixel<T> p = unpack_pixel<T>(pk, q[k], weights ? weights[k] : (T)1, m);
    T out[4];
    backward_pixel<T>(p, flags, meas[k], col_is_max, s, m ? dF[k] : (T)0, m ? dF[plane + k] : (T)0, w_pxy_hw, out);
    dU[k] = out[0];
    dU[plane + k] = out[1];
    if (flags & kWarp) {
      dU[2 * plane + k] = out[2];
      dU[3 * plane + k] = out[3];
    }
  }
}
static bool fused_tail() {   // default ON; EBOS_EKLT_FUSED_TAIL=0: separate k_column_max / k_adam / k_adam_bump launches (A/B)
  static const bool on = !(getenv("EBOS_EKLT_FUSED_TAIL") && atoi(getenv("EBOS_EKLT_FUSED_TAIL")) == 0);
  return on;
}
static bool stored_planes() {   // default ON (measured on B200: -5 % per evaluation at every level); EBOS_EKLT_STORED=0 disables
  const char* v = getenv("EBOS_EKLT_STORED");
  return v == nullptr || v[0] != '0';
}

// One CTA per padded cell (A,B): sums the four dense gradient planes over the cell's 2*patch x 2*patch support with
// the separable triangle weights.  Deterministic (no atomics).
template <typename T>
__global__ void __launch_bounds__(256) k_cell_gather(Geom g, int nch, const T* __restrict__ dU, T* __restrict__ dPad) {
  EKLT_PDL_PROLOGUE();
  __shared__ double red[32];
  const int PW = g.pw + 2 * g.pad, PH = g.ph + 2 * g.pad;
  const int A = blockIdx.y, B = blockIdx.x;
  int i0, i1, j0, j1;
  cell_support(A, g.patch, g.h1, g.H, i0, i1);
  cell_support(B, g.patch, g.w1, g.W, j0, j1);
  const int nj = j1 - j0, ni = i1 - i0;
  const int64_t plane = (int64_t)g.H * g.W;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  const int total = ni * nj;
  for (int k = threadIdx.x; k < total; k += blockDim.x) {
    const int i = i0 + k / nj, j = j0 + k % nj;
    const double w = (double)cell_weight<T>(A, i, g.h1, g.patch) * (double)cell_weight<T>(B, j, g.w1, g.patch);
    const int64_t idx = (int64_t)i * g.W + j;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < nch) s[c] += w * (double)dU[c * plane + idx];
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c >= nch) break;                      // uniform over the CTA
    const double t = block_sum(s[c], red);
    if (threadIdx.x == 0) dPad[((int64_t)c * PH + A) * PW + B] = (T)t;
  }
}

// TV(F; w_inv) value and w_tv * dTV/dF in gather form over the ROI box +-2 (F = f*M vanishes outside the ROI, and the
// backward reads dF only inside it).  One thread per pixel of the box, both channels.
template <typename T>
__global__ void __launch_bounds__(256, EBOS_EKLT_MINB) k_tv_roi(Geom g, const T* __restrict__ F, const T* __restrict__ winv, double coef,
                                                double* __restrict__ acc, T* __restrict__ dF) {
  EKLT_PDL_PROLOGUE();
  __shared__ double red[32];
  int r0, r1, c0, c1;
  tv_box(g.x0, g.x1, g.H, r0, r1);
  tv_box(g.y0, g.y1, g.W, c0, c1);
  const int j = c0 + blockIdx.x * 32 + (threadIdx.x & 31);
  const int64_t plane = (int64_t)g.H * g.W;
  double val = 0.0;
  if (j < c1) {
    for (int i = r0 + blockIdx.y * 8 + (threadIdx.x >> 5); i < r1; i += gridDim.y * 8) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double v, a;
        tv_pixel<T>(F + c * plane, winv, g.H, g.W, i, j, v, a);
        val += v;
        dF[c * plane + (int64_t)i * g.W + j] = (T)(coef * a);
      }
    }
  }
  val = block_sum(val, red);
  if (threadIdx.x == 0) atomicAdd(acc + kAccTvSum + acc_slot(), val);
}

// Transposed up-sampling in two separable passes.  Pass 1 (along columns): one warp per (row i, padded column cell B),
// lanes over the cell's 2*patch support columns (coalesced), four channels.  T1: [nch, H, PW].
template <typename T>
__global__ void __launch_bounds__(256) k_gather_cols(Geom g, int nch, const T* __restrict__ dU, T* __restrict__ T1) {
  EKLT_PDL_PROLOGUE();
  const int PW = g.pw + 2 * g.pad;
  const int B = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, i = blockIdx.y;
  if (B >= PW) return;
  int j0, j1;
  cell_support(B, g.patch, g.w1, g.W, j0, j1);
  const int64_t plane = (int64_t)g.H * g.W;
  const T* row = dU + (int64_t)i * g.W;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int j = j0 + lane; j < j1; j += 32) {
    const double w = (double)cell_weight<T>(B, j, g.w1, g.patch);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < nch) s[c] += w * (double)row[c * plane + j];
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c >= nch) break;
    const double t = warp_sum(s[c]);
    if (lane == 0) T1[((int64_t)c * g.H + i) * PW + B] = (T)t;
  }
}
// Pass 2 (along rows): one warp per (channel, padded cell), lanes over the cell's 2*patch support rows of T1.
template <typename T>
__global__ void __launch_bounds__(256) k_gather_rows(Geom g, int nch, const T* __restrict__ T1, T* __restrict__ dPad) {
  EKLT_PDL_PROLOGUE();
  const int PW = g.pw + 2 * g.pad, PH = g.ph + 2 * g.pad;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= nch * PH * PW) return;
  const int c = o / (PH * PW), A = (o / PW) % PH, B = o % PW;
  int i0, i1;
  cell_support(A, g.patch, g.h1, g.H, i0, i1);
  double s = 0.0;
  for (int i = i0 + lane; i < i1; i += 32)
    s += (double)cell_weight<T>(A, i, g.h1, g.patch) * (double)T1[((int64_t)c * g.H + i) * PW + B];
  s = warp_sum(s);
  if (lane == 0) dPad[((int64_t)c * PH + A) * PW + B] = (T)s;
}

// Default since round 2 for patch <= 32 (EBOS_EKLT_GATHER_SEG=0 disables): column pass for small patches.  A warp
// reads 32 consecutive pixels of a row ONCE (coalesced, all lanes busy), lane groups of `patch` lanes -- aligned with the
// runs of equal floor cell -- reduce their lo / hi weighted sums with shuffles, and the group leaders add them to the two
// column cells of T1 (zeroed first; <= 2 contributions per address).  k_gather_cols leaves half a warp idle at 8-px patches
// and reads every pixel twice (51 us); the warp-per-cell 2-D form reads it four times (36 us).
template <typename T>
__global__ void __launch_bounds__(256) k_gather_cols_seg(Geom g, int nch, const T* __restrict__ dU, T* __restrict__ T1) {
  EKLT_PDL_PROLOGUE();
  const int PW = g.pw + 2 * g.pad, p = g.patch;
  const int start = region_offset(g.w1, p) - p + (blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
  if (start >= g.W) return;                      // whole warps leave together
  const int lane = threadIdx.x & 31, i = blockIdx.y, j = start + lane;
  const bool valid = j >= 0 && j < g.W;
  int A;
  T hi;
  segment_tap<T>(j, g.w1, p, A, hi);
  const int64_t plane = (int64_t)g.H * g.W;
  double lo_s[4], hi_s[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double v = (valid && c < nch) ? (double)dU[c * plane + (int64_t)i * g.W + j] : 0.0;
    hi_s[c] = v * (double)hi;
    lo_s[c] = v - hi_s[c];                        // v * (1 - hi)
  }
  for (int o = p >> 1; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      lo_s[c] += __shfl_xor_sync(0xffffffffu, lo_s[c], o);
      hi_s[c] += __shfl_xor_sync(0xffffffffu, hi_s[c], o);
    }
  }
  if ((lane & (p - 1)) == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c >= nch) break;
      T* row = T1 + ((int64_t)c * g.H + i) * PW;
      if (A >= 0 && A < PW) atomicAdd(row + A, (T)lo_s[c]);
      if (A + 1 >= 0 && A + 1 < PW) atomicAdd(row + A + 1, (T)hi_s[c]);
    }
  }
}
// Row pass with one THREAD per (channel, padded cell), consecutive lanes on consecutive column cells (coalesced rows of
// T1): for small patches, where a warp per output would have 2*patch <= 32 rows to share.
template <typename T>
__global__ void __launch_bounds__(256) k_gather_rows_thread(Geom g, int nch, const T* __restrict__ T1, T* __restrict__ dPad) {
  EKLT_PDL_PROLOGUE();
  const int PW = g.pw + 2 * g.pad, PH = g.ph + 2 * g.pad;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= nch * PH * PW) return;
  const int c = o / (PH * PW), A = (o / PW) % PH, B = o % PW;
  int i0, i1;
  cell_support(A, g.patch, g.h1, g.H, i0, i1);
  double s = 0.0;
  for (int i = i0; i < i1; ++i)
    s += (double)cell_weight<T>(A, i, g.h1, g.patch) * (double)T1[((int64_t)c * g.H + i) * PW + B];
  dPad[((int64_t)c * PH + A) * PW + B] = (T)s;
}
static bool gather_segments() {   // default ON (measured on B200: -6 % per evaluation at patch 16 / 8); EBOS_EKLT_GATHER_SEG=0 disables
  const char* v = getenv("EBOS_EKLT_GATHER_SEG");
  return v == nullptr || v[0] != '0';
}

// Same for small supports (patch <= 16: at most 1024 pixels per cell): one WARP per padded cell, eight cells per CTA,
// shuffle reductions only.  At the finest level (8-px patches, 92 x 162 cells at 1280x720) the CTA-per-cell form
// would spend its time in four block reductions over one pixel per thread.
template <typename T>
__global__ void __launch_bounds__(256) k_cell_gather_warp(Geom g, int nch, const T* __restrict__ dU, T* __restrict__ dPad) {
  EKLT_PDL_PROLOGUE();
  const int PW = g.pw + 2 * g.pad, PH = g.ph + 2 * g.pad;
  const int cell = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (cell >= PW * PH) return;                  // whole warps leave together
  const int A = cell / PW, B = cell % PW;
  int i0, i1, j0, j1;
  cell_support(A, g.patch, g.h1, g.H, i0, i1);
  cell_support(B, g.patch, g.w1, g.W, j0, j1);
  const int nj = j1 - j0, total = (i1 - i0) * nj;
  const int64_t plane = (int64_t)g.H * g.W;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k = lane; k < total; k += 32) {
    const int i = i0 + k / nj, j = j0 + k % nj;
    const double w = (double)cell_weight<T>(A, i, g.h1, g.patch) * (double)cell_weight<T>(B, j, g.w1, g.patch);
    const int64_t idx = (int64_t)i * g.W + j;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < nch) s[c] += w * (double)dU[c * plane + idx];
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c >= nch) break;
    const double t = warp_sum(s[c]);
    if (lane == 0) dPad[((int64_t)c * PH + A) * PW + B] = (T)t;
  }
}

template <typename T>
__global__ void k_fold(Geom g, int nch, const T* __restrict__ dPad, T* __restrict__ dP) {
  EKLT_PDL_PROLOGUE();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = g.ph * g.pw;
  if (k >= nch * np) return;
  const int c = k / np, a = (k % np) / g.pw, b = k % g.pw;
  const int PW = g.pw + 2 * g.pad, PH = g.ph + 2 * g.pad;
  int a0, a1, b0, b1;
  fold_range(a, g.ph, g.pad, a0, a1);
  fold_range(b, g.pw, g.pad, b0, b1);
  double s = 0.0;
  for (int A = a0; A < a1; ++A)
    for (int B = b0; B < b1; ++B) s += (double)dPad[((int64_t)c * PH + A) * PW + B];
  dP[k] = (T)s;
}

// Adam update fused into the last kernel of the evaluation (ebos_eklt_adam_iteration): theta / moments of the cell this
// thread just computed the gradient of; the step counter is advanced by whichever CTA finishes last.
template <typename T> struct AdamFuse {
  T* theta;
  T* m;
  T* v;
  double lr, b1, b2, eps;
  int32_t* step_dev;     // nullptr: plain value_and_grad, no update
};
template <typename T>
__device__ __forceinline__ void eklt_adam_one(T& p, T g, T& m, T& v, T b1, T b2, T eps, T step_size, T inv_bc2_sqrt) {
  m = m * b1 + ((T)1 - b1) * g;                       // same expressions as ebos_costs.cu: adam_one
  v = v * b2 + ((T)1 - b2) * g * g;
  const T denom = sqrt(v) * inv_bc2_sqrt + eps;
  p -= step_size * (m / denom);
}

template <typename T>
__global__ void k_param_grad(Geom g, int flags, const T* __restrict__ dP, const double* __restrict__ tv_acc,
                             double* __restrict__ acc, double w_data, double w_tv, double w_pxy, T* __restrict__ grad,
                             T* __restrict__ loss, AdamFuse<T> ad) {
  EKLT_PDL_PROLOGUE();
  __shared__ T s_coef[2];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = g.ph * g.pw;
  if (ad.step_dev) {
    if (threadIdx.x == 0) {
      const int step = *ad.step_dev + 1;
      s_coef[0] = (T)(ad.lr / (1.0 - pow(ad.b1, (double)step)));
      s_coef[1] = (T)(1.0 / sqrt(1.0 - pow(ad.b2, (double)step)));
    }
    __syncthreads();
  }
  auto put = [&](int idx, T gv) {
    grad[idx] = gv;
    if (ad.step_dev) {
      T p = ad.theta[idx], mm = ad.m[idx], vv = ad.v[idx];
      eklt_adam_one<T>(p, gv, mm, vv, (T)ad.b1, (T)ad.b2, (T)ad.eps, s_coef[0], s_coef[1]);
      ad.theta[idx] = p; ad.m[idx] = mm; ad.v[idx] = vv;
    }
  };
  if (k < np) {
    const int nf = flow_channels(flags);
    if (flags & kPoisson) {
      put(k, sobel_over_8_adjoint_at(dP, dP + np, g.ph, g.pw, k / g.pw, k % g.pw));
    } else {
      put(k, dP[k]);
      put(np + k, dP[np + k]);
    }
    if (flags & kWarp) {
      put(nf * np + k, dP[2 * np + k]);
      put((nf + 1) * np + k, dP[3 * np + k]);
    }
  }
  if (ad.step_dev) {
    __syncthreads();                 // every thread of this CTA has read the step counter's factors
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(acc + kAccTicketTail), 1u);
      if (t == gridDim.x - 1) *ad.step_dev += 1;
    }
  }
  if (k == 0) {
    double tv = acc_sum(acc, kAccTvSum);                             // k_tv_roi
    if (tv_acc) {                                                    // legacy chain: ebos_flow_tv's accumulators
      tv = tv_acc[3];
      for (int i = 24; i < EBOS_ACC_DOUBLES; ++i) tv += tv_acc[i];
    }
    const double hw = (double)g.H * (double)g.W;
    const double tv_mean = tv / (2.0 * hw), pxy_mean = (flags & kWarp) ? acc_sum(acc, kAccPxy) / hw : 0.0;
    const double total = w_data * acc[kAccMax] + w_tv * tv_mean + w_pxy * pxy_mean;
    acc[kAccData] = acc[kAccMax];          // un-weighted terms, for diagnostics
    acc[kAccTv] = tv_mean;
    acc[kAccPxyMean] = pxy_mean;
    acc[kAccLoss] = total;
    *loss = (T)total;
  }
}

// up-sampling on its own: the dense fields returned to the caller (flow from the intensity, translation)
template <typename T>
__global__ void __launch_bounds__(256) k_upsample(Geom g, const T* __restrict__ P, int channels, T* __restrict__ out) {
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  if (j >= g.W) return;
  const AxisTap<T> c = axis_tap<T>(j, g.w1, g.patch, g.pw, g.pad, g.inv_patch);
  for (int i = blockIdx.y * 8 + (threadIdx.x >> 5); i < g.H; i += gridDim.y * 8) {
    const AxisTap<T> r = axis_tap<T>(i, g.h1, g.patch, g.ph, g.pad, g.inv_patch);
    for (int ch = 0; ch < channels; ++ch)
      out[((int64_t)ch * g.H + i) * g.W + j] = upsample_at(P + (int64_t)ch * g.ph * g.pw, g.pw, r, c);
  }
}


// separable correlation, one pass per axis (axis 1 = along columns first, like cv2's row filter, then axis 0)
template <typename T>
__global__ void __launch_bounds__(256) k_correlate(const T* __restrict__ in, int H, int W, ConvTaps taps, int axis, int border,
                                                   T* __restrict__ out) {
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  if (j >= W) return;
  for (int i = blockIdx.y * 8 + (threadIdx.x >> 5); i < H; i += gridDim.y * 8) {
    out[(int64_t)i * W + j] = axis == 1 ? correlate_at<T>(in + (int64_t)i * W, 1, j, W, taps, border)
                                        : correlate_at<T>(in + j, W, i, H, taps, border);
  }
}

static dim3 plane_grid(int H, int W) {
  const int gx = (W + 31) / 32;
  // enough row groups to fill the machine about four times over, at most one group per 8 rows
  const int want = std::max(1, (sm_count() * 8 + gx - 1) / gx);
  return dim3(gx, std::min((H + 7) / 8, want));
}

static int check_geometry(const Geom& g) {
  EBOS_REQUIRE(g.H >= 2 && g.W >= 2 && g.ph >= 1 && g.pw >= 1 && g.patch >= 1, "ebos_eklt: bad size");
  EBOS_REQUIRE(g.ph * g.patch >= g.H && g.pw * g.patch >= g.W &&
                   (g.ph - 1) * g.patch < g.H && (g.pw - 1) * g.patch < g.W,
               "ebos_eklt: patch grid must be ceil(H/patch) x ceil(W/patch)");
  EBOS_REQUIRE(g.x0 >= 0 && g.x0 <= g.x1 && g.x1 <= g.H && g.y0 >= 0 && g.y0 <= g.y1 && g.y1 <= g.W,
               "ebos_eklt: ROI outside the image");
  return EBOS_OK;
}

template <typename T>
static int value_and_grad_t(const Geom& g, int flags, const T* theta, const T* gx, const T* gy, const T* meas,
                            const T* winv, const T* weights, double w_data, double w_tv, double w_pxy, void* workspace,
                            T* loss, T* grad, cudaStream_t st, const AdamFuse<T>* adam = nullptr) {
  const Workspace w = carve(workspace, g.H, g.W, g.ph, g.pw, g.pad, sizeof(T));
  cudaError_t e = cudaMemsetAsync(w.acc, 0, (kAccN + 2 * (size_t)g.W) * sizeof(double), st);
  if (e != cudaSuccess) return cuda_fail(e, "ebos_eklt memset");
  const int np = g.ph * g.pw;
  const dim3 pg = plane_grid(g.H, g.W);
  T* q = reinterpret_cast<T*>(w.q);
  T* F = reinterpret_cast<T*>(w.F);
  T* dF = reinterpret_cast<T*>(w.dF);
  T* dU = reinterpret_cast<T*>(w.dU);
  T* dPad = reinterpret_cast<T*>(w.dPad);
  T* dP = reinterpret_cast<T*>(w.dP);
  const bool warp = flags & kWarp;
  const int nch = warp ? 4 : 2;
  if (!warp) w_pxy = 0.0;
  const T* pf = theta;                                         // flow model: theta[0:2] is the patch flow
  if (flags & kPoisson) {
    EKLT_LAUNCH(k_patch_flow<T>, (np + 127) / 128, 128, st, theta, g.ph, g.pw, reinterpret_cast<T*>(w.pf));
    pf = reinterpret_cast<const T*>(w.pf);
  }
  const T* tr = warp ? theta + (size_t)flow_channels(flags) * np : nullptr;
  // stored-planes backward: not with no_polarity (the sign of q0 is not kept) and not without the warp (nothing to store)
  const bool stored = stored_planes() && warp && !(flags & kNoPolarity);
  T* St = stored ? reinterpret_cast<T*>(w.St) : nullptr;
  EKLT_LAUNCH(k_forward<T>, pg, 256, st, g, flags, pf, tr, gx, gy, weights, q, F, w.acc, St);
  EBOS_LAUNCH_CHECK("ebos_eklt forward");
  const bool legacy = legacy_chain();
  if (legacy || w_tv == 0.0) {
    const int rc = ebos_flow_tv(F, winv, g.H, g.W, w_tv, sizeof(T) == 8 ? EBOS_F64 : EBOS_F32, w.tv_acc, dF, st);
    if (rc != EBOS_OK) return rc;
  } else {
    int r0, r1, c0, c1;
    tv_box(g.x0, g.x1, g.H, r0, r1);
    tv_box(g.y0, g.y1, g.W, c0, c1);
    if (r1 > r0 && c1 > c0) {
      const dim3 tg = plane_grid(r1 - r0, c1 - c0);
      EKLT_LAUNCH(k_tv_roi<T>, tg, 256, st, g, F, winv, w_tv / (2.0 * (double)g.H * (double)g.W), w.acc, dF);
    }
  }
  const bool tv_from_flow_tv = legacy || w_tv == 0.0;
  const bool fuse = fused_tail();
  EKLT_LAUNCH(k_column_sums<T>, pg, 256, st, g, q, meas, w.acc, w.colsum, w.colS, fuse ? 1 : 0, w_data);
  if (!fuse) EKLT_LAUNCH(k_column_max, 1, 256, st, g.W, g.y0, g.y1, w.colsum, w.colS, w.acc, w_data);
  if (stored)
    EKLT_LAUNCH(k_backward_stored<T>, pg, 256, st, g, flags, St, q, weights, meas, dF, w.colsum, w.acc,
                                             w_pxy / ((double)g.H * g.W), dU);
  else
    EKLT_LAUNCH(k_backward<T>, pg, 256, st, g, flags, pf, tr, gx, gy, weights, meas, dF, w.colsum, w.acc,
                                      w_pxy / ((double)g.H * g.W), dU);
  EBOS_LAUNCH_CHECK("ebos_eklt backward");
  const int PW = g.pw + 2 * g.pad, PH = g.ph + 2 * g.pad;
  const int n_cells = PW * PH;
  // B200, 1280x720 fp64 (profiles/eklt/r01i launch list): two separable passes 20 / 25 / 36 / 71 us at patch 64 / 32 /
  // 16 / 8 against 54 / 33 us (CTA per cell) and 36 / 36 us (warp per cell): a warp of pass 1 covers one 2*patch-pixel
  // support, which leaves half of it idle at patch 8.
  if (!legacy && gather_segments() && g.patch >= 2 && g.patch <= 32 && 32 % g.patch == 0) {
    T* T1 = reinterpret_cast<T*>(w.T1);
    e = cudaMemsetAsync(T1, 0, (size_t)nch * g.H * PW * sizeof(T), st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_eklt memset T1");
    const int n_seg = (g.W - (region_offset(g.w1, g.patch) - g.patch) + 31) / 32;
    EKLT_LAUNCH(k_gather_cols_seg<T>, dim3((n_seg + 7) / 8, g.H), 256, st, g, nch, dU, T1);
    if (g.patch <= 16) EKLT_LAUNCH(k_gather_rows_thread<T>, (nch * n_cells + 255) / 256, 256, st, g, nch, T1, dPad);
    else EKLT_LAUNCH(k_gather_rows<T>, (nch * n_cells + 7) / 8, 256, st, g, nch, T1, dPad);
  } else if (g.patch <= 16) {
    EKLT_LAUNCH(k_cell_gather_warp<T>, (n_cells + 7) / 8, 256, st, g, nch, dU, dPad);
  } else if (legacy) {
    EKLT_LAUNCH(k_cell_gather<T>, dim3(PW, PH), 256, st, g, nch, dU, dPad);
  } else {
    T* T1 = reinterpret_cast<T*>(w.T1);
    EKLT_LAUNCH(k_gather_cols<T>, dim3((PW + 7) / 8, g.H), 256, st, g, nch, dU, T1);
    EKLT_LAUNCH(k_gather_rows<T>, (nch * n_cells + 7) / 8, 256, st, g, nch, T1, dPad);
  }
  EKLT_LAUNCH(k_fold<T>, (nch * np + 127) / 128, 128, st, g, nch, dPad, dP);
  AdamFuse<T> ad{};
  if (adam && fuse) ad = *adam;
  EKLT_LAUNCH(k_param_grad<T>, (np + 127) / 128, 128, st, g, flags, dP, tv_from_flow_tv ? w.tv_acc : nullptr, w.acc, w_data,
                                                    w_tv, w_pxy, grad, loss, ad);
  EBOS_LAUNCH_CHECK("ebos_eklt gradient");
  return EBOS_OK;
}

}  // namespace eklt
}  // namespace ebos

using namespace ebos;
using namespace ebos::eklt;

extern "C" {

size_t ebos_eklt_workspace_bytes(int H, int W, int ph, int pw, int patch, int dtype) {
  if (H < 1 || W < 1 || ph < 1 || pw < 1 || patch < 1) return 0;
  const Geom g = make_geom(H, W, ph, pw, patch, 0, H, 0, W);
  return carve(nullptr, H, W, ph, pw, g.pad, dtype_size(dtype)).total;
}

int ebos_eklt_value_and_grad(const void* theta, int flags, const void* grad_x, const void* grad_y, const void* measured,
                             const void* weight_inverse, const void* weights, int H, int W, int ph, int pw, int patch,
                             int roi_x0, int roi_x1, int roi_y0, int roi_y1, double w_data, double w_tv, double w_pxy,
                             int dtype, void* workspace, size_t workspace_bytes, void* loss, void* grad, void* stream) {
  EBOS_REQUIRE(theta && grad_x && grad_y && measured && weight_inverse && workspace && loss && grad,
               "ebos_eklt_value_and_grad: null argument");
  EBOS_REQUIRE((flags & ~(kPoisson | kWarp | kNoPolarity)) == 0, "ebos_eklt_value_and_grad: unknown flag");
  if (dtype != EBOS_F32 && dtype != EBOS_F64) {
    set_error("ebos_eklt_value_and_grad: unsupported dtype");
    return EBOS_ERR_UNSUPPORTED;
  }
  const Geom g = make_geom(H, W, ph, pw, patch, roi_x0, roi_x1, roi_y0, roi_y1);
  const int rc = check_geometry(g);
  if (rc != EBOS_OK) return rc;
  if (workspace_bytes < ebos_eklt_workspace_bytes(H, W, ph, pw, patch, dtype)) {
    set_error("ebos_eklt_value_and_grad: workspace too small");
    return EBOS_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  if (dtype == EBOS_F64)
    return value_and_grad_t<double>(g, flags, (const double*)theta, (const double*)grad_x, (const double*)grad_y,
                                    (const double*)measured, (const double*)weight_inverse, (const double*)weights,
                                    w_data, w_tv, w_pxy, workspace, (double*)loss, (double*)grad, st);
  return value_and_grad_t<float>(g, flags, (const float*)theta, (const float*)grad_x, (const float*)grad_y,
                                 (const float*)measured, (const float*)weight_inverse, (const float*)weights, w_data,
                                 w_tv, w_pxy, workspace, (float*)loss, (float*)grad, st);
}

int ebos_eklt_adam_iteration(void* theta, int flags, const void* grad_x, const void* grad_y, const void* measured,
                             const void* weight_inverse, const void* weights, int H, int W, int ph, int pw, int patch,
                             int roi_x0, int roi_x1, int roi_y0, int roi_y1, double w_data, double w_tv, double w_pxy,
                             int dtype, void* workspace, size_t workspace_bytes, void* loss, void* grad, void* exp_avg,
                             void* exp_avg_sq, double lr, double beta1, double beta2, double eps, int32_t* step_dev,
                             void* stream) {
  EBOS_REQUIRE(exp_avg && exp_avg_sq && step_dev, "ebos_eklt_adam_iteration: null argument");
  if (fused_tail()) {
    // same checks as ebos_eklt_value_and_grad, then the evaluation with the Adam update inside its last kernel
    EBOS_REQUIRE(theta && grad_x && grad_y && measured && weight_inverse && workspace && loss && grad,
                 "ebos_eklt_adam_iteration: null argument");
    EBOS_REQUIRE((flags & ~(kPoisson | kWarp | kNoPolarity)) == 0, "ebos_eklt_adam_iteration: unknown flag");
    if (dtype != EBOS_F32 && dtype != EBOS_F64) { set_error("ebos_eklt_adam_iteration: unsupported dtype"); return EBOS_ERR_UNSUPPORTED; }
    const Geom g = make_geom(H, W, ph, pw, patch, roi_x0, roi_x1, roi_y0, roi_y1);
    const int grc = check_geometry(g);
    if (grc != EBOS_OK) return grc;
    if (workspace_bytes < ebos_eklt_workspace_bytes(H, W, ph, pw, patch, dtype)) {
      set_error("ebos_eklt_adam_iteration: workspace too small");
      return EBOS_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    if (dtype == EBOS_F64) {
      const AdamFuse<double> ad{(double*)theta, (double*)exp_avg, (double*)exp_avg_sq, lr, beta1, beta2, eps, step_dev};
      return value_and_grad_t<double>(g, flags, (const double*)theta, (const double*)grad_x, (const double*)grad_y,
                                      (const double*)measured, (const double*)weight_inverse, (const double*)weights,
                                      w_data, w_tv, w_pxy, workspace, (double*)loss, (double*)grad, st, &ad);
    }
    const AdamFuse<float> ad{(float*)theta, (float*)exp_avg, (float*)exp_avg_sq, lr, beta1, beta2, eps, step_dev};
    return value_and_grad_t<float>(g, flags, (const float*)theta, (const float*)grad_x, (const float*)grad_y,
                                   (const float*)measured, (const float*)weight_inverse, (const float*)weights, w_data,
                                   w_tv, w_pxy, workspace, (float*)loss, (float*)grad, st, &ad);
  }
  const int rc = ebos_eklt_value_and_grad(theta, flags, grad_x, grad_y, measured, weight_inverse, weights, H, W, ph, pw,
                                          patch, roi_x0, roi_x1, roi_y0, roi_y1, w_data, w_tv, w_pxy, dtype, workspace,
                                          workspace_bytes, loss, grad, stream);
  if (rc != EBOS_OK) return rc;
  return ebos_adam_step_graph(theta, grad, exp_avg, exp_avg_sq, (int64_t)theta_channels(flags) * ph * pw, lr, beta1,
                              beta2, eps, step_dev, dtype, stream);
}

int ebos_eklt_upsample(const void* patch_values, int channels, int H, int W, int ph, int pw, int patch, int dtype,
                       void* dense, void* stream) {
  EBOS_REQUIRE(patch_values && dense && channels >= 1, "ebos_eklt_upsample: bad argument");
  if (dtype != EBOS_F32 && dtype != EBOS_F64) {
    set_error("ebos_eklt_upsample: unsupported dtype");
    return EBOS_ERR_UNSUPPORTED;
  }
  const Geom g = make_geom(H, W, ph, pw, patch, 0, H, 0, W);
  const int rc = check_geometry(g);
  if (rc != EBOS_OK) return rc;
  const dim3 pg = plane_grid(H, W);
  if (dtype == EBOS_F64)
    k_upsample<double><<<pg, 256, 0, as_stream(stream)>>>(g, (const double*)patch_values, channels, (double*)dense);
  else
    k_upsample<float><<<pg, 256, 0, as_stream(stream)>>>(g, (const float*)patch_values, channels, (float*)dense);
  EBOS_LAUNCH_CHECK("ebos_eklt_upsample");
  return EBOS_OK;
}

int ebos_eklt_patch_flow(const void* intensity, int ph, int pw, int dtype, void* patch_flow, void* stream) {
  EBOS_REQUIRE(intensity && patch_flow && ph >= 1 && pw >= 1, "ebos_eklt_patch_flow: bad argument");
  if (dtype != EBOS_F32 && dtype != EBOS_F64) {
    set_error("ebos_eklt_patch_flow: unsupported dtype");
    return EBOS_ERR_UNSUPPORTED;
  }
  const int np = ph * pw;
  if (dtype == EBOS_F64)
    k_patch_flow<double><<<(np + 127) / 128, 128, 0, as_stream(stream)>>>((const double*)intensity, ph, pw, (double*)patch_flow);
  else
    k_patch_flow<float><<<(np + 127) / 128, 128, 0, as_stream(stream)>>>((const float*)intensity, ph, pw, (float*)patch_flow);
  EBOS_LAUNCH_CHECK("ebos_eklt_patch_flow");
  return EBOS_OK;
}

int ebos_sepconv2d(const void* image, int H, int W, const double* taps_rows, int n_rows, const double* taps_cols,
                   int n_cols, int border, int dtype, void* tmp, void* out, void* stream) {
  EBOS_REQUIRE(image && tmp && out && taps_rows && taps_cols && H >= 1 && W >= 1, "ebos_sepconv2d: bad argument");
  EBOS_REQUIRE(n_rows >= 1 && n_rows <= kMaxTaps && (n_rows & 1) && n_cols >= 1 && n_cols <= kMaxTaps && (n_cols & 1),
               "ebos_sepconv2d: tap counts must be odd and at most 127");
  EBOS_REQUIRE(border == 0 || border == 1, "ebos_sepconv2d: border must be 0 (reflect-101) or 1 (reflect)");
  if (dtype != EBOS_F32 && dtype != EBOS_F64) {
    set_error("ebos_sepconv2d: unsupported dtype");
    return EBOS_ERR_UNSUPPORTED;
  }
  ConvTaps tr, tc;
  tr.n = n_rows;
  tc.n = n_cols;
  for (int k = 0; k < kMaxTaps; ++k) {
    tr.w[k] = k < n_rows ? taps_rows[k] : 0.0;
    tc.w[k] = k < n_cols ? taps_cols[k] : 0.0;
  }
  const dim3 pg = plane_grid(H, W);
  cudaStream_t st = as_stream(stream);
  if (dtype == EBOS_F64) {
    k_correlate<double><<<pg, 256, 0, st>>>((const double*)image, H, W, tc, 1, border, (double*)tmp);
    k_correlate<double><<<pg, 256, 0, st>>>((const double*)tmp, H, W, tr, 0, border, (double*)out);
  } else {
    k_correlate<float><<<pg, 256, 0, st>>>((const float*)image, H, W, tc, 1, border, (float*)tmp);
    k_correlate<float><<<pg, 256, 0, st>>>((const float*)tmp, H, W, tr, 0, border, (float*)out);
  }
  EBOS_LAUNCH_CHECK("ebos_sepconv2d");
  return EBOS_OK;
}

}  // extern "C"
