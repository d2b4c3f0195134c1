// Fused contrast-maximisation path on a prepared window (fp32 fast path, fp64 parity path).
//
//   prepare (once per window):  time stats -> dt -> origin pixel key -> stable sort by key -> SoA
//   splat   (every iteration):  stream the pixel-sorted SoA, gather flow, warp, bilinear vote.
//                               Consecutive events of one origin pixel share the flow vector and
//                               are time ordered, so they fall into the same IWE cell in runs; each
//                               thread combines a run in registers and issues one REDG per tap per
//                               run instead of one per event (measured 4.5x fewer atomics at 16 Mi
//                               events, profiles/microbench/).
//   backward(every iteration):  same stream, re-warp, gather dL/dIWE at the taps, combine all events
//                               of an origin pixel in registers, one REDG pair per pixel run.
//
// Reference semantics: src/warp.py:283-287,333-337 and src/event_image_converter.py:586-619
// (SURVEY.md A.1-A.3).  Coordinate arithmetic is unfused so cells/masks match the reference bit
// for bit; only the floating-point ADD ORDER into a pixel differs (atomic mode, 1e-5 relative).
// The reference's solvers run in float64 (src/solver/patch_eklt_pyramid2.py:253); the fp64
// instantiation is the dtype-faithful path used for solve-level parity.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>

#include "ebos_common.cuh"

namespace ebos {


// events per thread (consecutive, so that runs can be combined in registers)
template <typename T> struct Ept;
template <> struct Ept<float> { static constexpr int splat = 8, bwd = 4; };
template <> struct Ept<double> { static constexpr int splat = 4, bwd = 4; };
// experiment knob (EBOS_SPLAT_EPT / EBOS_BWD_EPT environment variables; 0 = default)
static int env_int(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}

// ---- prepare --------------------------------------------------------------------------------
__global__ void k_win_hdr_init(WindowHeader* h) {
  h->enc_min = ~0ull;
  h->enc_max = 0ull;
  h->not_packable = 0;
  h->packed = 0;
}

template <typename T>
__global__ void __launch_bounds__(256) k_win_minmax(const T* __restrict__ ev, int64_t n, WindowHeader* h) {
  using U = typename Enc<T>::U;
  U lo = ~(U)0, hi = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    U u = Enc<T>::enc(__ldg(ev + 4 * i + 2));
    lo = u < lo ? u : lo;
    hi = u > hi ? u : hi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    U l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    // both encodings are order preserving, so the 64-bit slots can hold either width
    atomicMin(&h->enc_min, (unsigned long long)lo);
    atomicMax(&h->enc_max, (unsigned long long)hi);
  }
}

template <typename T>
__global__ void k_win_hdr_final(WindowHeader* h, int direction, double frac, const T* __restrict__ tminmax) {
  using U = typename Enc<T>::U;
  T tmin = tminmax ? tminmax[0] : Enc<T>::dec((U)h->enc_min);
  T tmax = tminmax ? tminmax[1] : Enc<T>::dec((U)h->enc_max);
  TimeRef<T> tr = make_time_ref<T>(tmin, tmax, direction, frac);
  h->t_min = (double)tmin;
  h->t_max = (double)tmax;
  h->t_ref = (double)tr.t_ref;
  h->period = (double)tr.period;
}

// key = origin pixel k (src/warp.py:334); an event whose k is outside the grid is flagged and
// parked behind all valid pixels.  Also records whether any coordinate is NOT a small non-negative
// integer (then the packed (row,col) layout cannot be used).
template <typename T>
__global__ void __launch_bounds__(256) k_win_keys(const T* __restrict__ ev, int64_t n, int H, int W,
                                                  unsigned int* __restrict__ keys, int* __restrict__ idx,
                                                  int32_t* __restrict__ status, WindowHeader* __restrict__ h) {
  const int64_t hw = (int64_t)H * W;
  bool bad = false, frac = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T x = __ldg(ev + 4 * i), y = __ldg(ev + 4 * i + 1);
    int64_t k = (int64_t)x * W + (int64_t)y;
    bool ok = k >= 0 && k < hw && Rn<T>::finite(x) && Rn<T>::finite(y);
    bad |= !ok;
    frac |= !(ok && x >= (T)0 && y >= (T)0 && x < (T)65536 && y < (T)65536 && x == (T)(int)x && y == (T)(int)y);
    keys[i] = ok ? (unsigned int)k : (unsigned int)hw;
    idx[i] = (int)i;
  }
  if (bad) atomicOr(status, EBOS_STATUS_PIXEL_OOB);
  if (frac) atomicOr(&h->not_packable, 1);
}

__global__ void k_win_layout(WindowHeader* h, int allow_packed, int32_t* __restrict__ status) {
  h->packed = (allow_packed && !h->not_packable) ? 1 : 0;
  if (h->packed) atomicOr(status, EBOS_STATUS_PACKED);
}

template <typename T>
__global__ void __launch_bounds__(256) k_win_gather(const T* __restrict__ ev, const T* __restrict__ weight, int64_t n,
                                                    int H, int W, const int* __restrict__ perm,
                                                    const WindowHeader* __restrict__ h, int normalize_t,
                                                    T* __restrict__ sx, T* __restrict__ sy, T* __restrict__ sd,
                                                    T* __restrict__ sw) {
  TimeRef<T> tr{(T)h->t_ref, (T)h->period};
  const int64_t hw = (int64_t)H * W;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = __ldg(perm + j);
    const T x = __ldg(ev + 4 * i), y = __ldg(ev + 4 * i + 1), t = __ldg(ev + 4 * i + 2);
    int64_t k = (int64_t)x * W + (int64_t)y;
    bool ok = k >= 0 && k < hw && Rn<T>::finite(x) && Rn<T>::finite(y);
    const T dt = event_dt<T>(t, tr, normalize_t);
    if (h->packed) {
      // packed layout (all events valid, integer coordinates < 65536): (row << 16 | col) in the x slot
      reinterpret_cast<unsigned int*>(sx)[j] = ((unsigned int)(int)x << 16) | (unsigned int)(int)y;
    } else {
      // an invalid event is parked as x = NaN, y = 0: it gathers flow[0] harmlessly, its warped
      // coordinate is NaN and the exact path drops it (a NaN ORIGIN coordinate has no pixel).
      sx[j] = ok ? x : (T)NAN;
      sy[j] = ok ? y : (T)0;
    }
    sd[j] = dt;
    if (sw) sw[j] = __ldg(weight + i);
  }
}

// ---- per-thread event block loads ---------------------------------------------------------------
__device__ __forceinline__ void load4(const float* p, float* o) {
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void load4(const double* p, double* o) {
  double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <typename T, int EPT>
__device__ __forceinline__ void load_block(const T* __restrict__ p, int64_t base, int64_t n, T fill, T (&out)[EPT]) {
  if (base + EPT <= n) {
#pragma unroll
    for (int j = 0; j < EPT; j += 4) load4(p + base + j, out + j);
  } else {
#pragma unroll
    for (int j = 0; j < EPT; ++j) out[j] = (base + j < n) ? p[base + j] : fill;
  }
}

// The EPT consecutive events of one thread: origin coordinates, dt, weight, and the flow at the origin
// pixel.  Every load is issued before any of them is used: the first profile (r01) showed each thread
// doing EPT serial gather -> compute -> RED round trips (41 % of stall samples on the long scoreboard).
// Two storage layouts: generic (x, y as T) and packed ((row << 16 | col) as uint32; fp32 windows whose
// coordinates are all integers), which saves a third of the stream and all float->int conversions.
// Events that must be skipped (tail of the last thread, invalid events parked by prepare) carry x = NaN:
// they gather flow[0], produce a NaN warped coordinate and are dropped by the exact path.
template <typename T, int EPT, bool HAS_W, bool PACKED>
struct EventBlock {
  T x[EPT], y[EPT], d[EPT], wt[EPT], f0[EPT], f1[EPT];
  int k[EPT];
  __device__ __forceinline__ void load(const T* __restrict__ sx, const T* __restrict__ sy, const T* __restrict__ sd,
                                       const T* __restrict__ sw, int64_t base, int64_t n, const T* __restrict__ flow,
                                       int W, int hw) {
    load_block<T, EPT>(sd, base, n, (T)0, d);
    if (HAS_W) load_block<T, EPT>(sw, base, n, (T)0, wt);
    if constexpr (PACKED) {
      float rcf[EPT];
      load_block<float, EPT>(reinterpret_cast<const float*>(sx), base, n, __uint_as_float(0xffffffffu), rcf);
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const unsigned rc = __float_as_uint(rcf[j]);
        const bool skip = rc == 0xffffffffu;  // only the tail of the very last thread
        const int r = rc >> 16, c = rc & 0xffff;
        k[j] = skip ? 0 : r * W + c;
        // exact int -> float for values below 2^22 on the FP32 pipe
        x[j] = skip ? (T)NAN : (T)(__int_as_float(0x4B400000 + r) - 12582912.0f);
        y[j] = (T)(__int_as_float(0x4B400000 + c) - 12582912.0f);
      }
    } else {
      load_block<T, EPT>(sx, base, n, (T)NAN, x);
      load_block<T, EPT>(sy, base, n, (T)0, y);
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int kk = (int)x[j] * W + (int)y[j];   // NaN converts to 0: parked events gather flow[0]
        k[j] = (unsigned)kk < (unsigned)hw ? kk : 0;
      }
    }
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      f0[j] = __ldg(flow + k[j]);
      f1[j] = __ldg(flow + hw + k[j]);
    }
  }
};

// ---- forward: fused warp + bilinear vote -----------------------------------------------------
// The per-event arithmetic is that of make_taps<T> (ebos_common.cuh, i.e. src/event_image_converter.py:
// 586-614), arranged for latency and instruction issue:
//  * floor() runs on the FP32 pipe: for |v| < 2^22, (v + 1.5*2^23) - 1.5*2^23 is the round-to-nearest
//    integer and one compare turns it into floor (FRND/F2I run at quarter rate);
//  * the current cell is tracked as the two floor VALUES (floats); the integer cell is only formed when
//    a run is flushed;
//  * anything unusual (|coordinate| >= 2^22, Inf, NaN, parked events) leaves through one out-of-line
//    exact path.
template <typename T> struct FastFloor;
template <> struct FastFloor<float> {
  static __device__ __forceinline__ bool in_range(float a, float b) { return fabsf(a) < 4194304.0f && fabsf(b) < 4194304.0f; }
  static __device__ __forceinline__ float flr(float v) {
    const float M = 12582912.0f;
    float f = __fsub_rn(__fadd_rn(v, M), M);
    return (f > v) ? __fsub_rn(f, 1.0f) : f;
  }
  // exact for integer-valued |f| < 2^22
  static __device__ __forceinline__ int to_int(float f) { return __float_as_int(__fadd_rn(f, 12582912.0f)) - 0x4B400000; }
};
template <> struct FastFloor<double> {
  static __device__ __forceinline__ bool in_range(double a, double b) { return fabs(a) < 1073741824.0 && fabs(b) < 1073741824.0; }
  static __device__ __forceinline__ double flr(double v) { return floor(v); }
  static __device__ __forceinline__ int to_int(double f) { return (int)f; }
};

// Exact (reference-order) handling of one event whose warped coordinate is outside the fast range.
// x0 is the ORIGIN row coordinate: NaN marks an event that must be skipped.
template <typename T>
__device__ __noinline__ void splat_event_exact(T* __restrict__ iwe, int Hp, int Wp, int pad_h, int pad_w, T x0, T xw,
                                               T yw, T wt) {
  if (x0 != x0) return;
  const Taps<T> t = make_taps<T>(xw, yw, pad_h, pad_w);
  const bool fin = Rn<T>::finite(xw) && Rn<T>::finite(yw);
  const int rr[4] = {t.r, t.r + 1, t.r, t.r + 1}, cc[4] = {t.c, t.c, t.c + 1, t.c + 1};
  const T ww[4] = {t.w0, t.w1, t.w2, t.w3};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const T v = Rn<T>::mul(ww[k], wt);
    const bool m = fin && (unsigned)rr[k] < (unsigned)Hp && (unsigned)cc[k] < (unsigned)Wp &&
                   !(k & 1 && t.r == INT_MAX) && !(k & 2 && t.c == INT_MAX);
    if (m) red_add_nc(iwe + ((int64_t)rr[k] * Wp + cc[k]), v);
    else if (!Rn<T>::finite(v)) red_add_nc(iwe, Rn<T>::mul(v, (T)0));  // vals*mask = NaN lands on pixel 0
  }
}

// Flush one run: add the four tap sums of cell (r,c).  When the row stride allows 16-byte alignment (VEC)
// the two column-adjacent taps of a row go out as ONE aligned red.v4 (zeros in the unused lanes), i.e.
// 2 LSU lane-ops per cell instead of 4, unless the pair straddles a 16-byte boundary (c % 4 == 3).
template <typename T, bool VEC>
__device__ __forceinline__ void flush_cell(T* __restrict__ iwe, int Hp, int Wp, int Hm1, int Wm1, int r, int c, T a0,
                                           T a1, T a2, T a3) {
  if ((unsigned)r < (unsigned)Hm1 && (unsigned)c < (unsigned)Wm1) {
    // all four taps inside: the common case
    T* p = iwe + (r * Wp + c);
    if constexpr (VEC) {
      const int j = c & 3;
      if (j != 3) {
        float* q = reinterpret_cast<float*>(p) - j;
        const float z = 0.0f;
        const bool j0 = j == 0, j1 = j == 1, j2 = j == 2;
        red_add_v4(q, j0 ? a0 : z, j0 ? a2 : (j1 ? a0 : z), j1 ? a2 : (j2 ? a0 : z), j2 ? a2 : z);
        red_add_v4(q + Wp, j0 ? a1 : z, j0 ? a3 : (j1 ? a1 : z), j1 ? a3 : (j2 ? a1 : z), j2 ? a3 : z);
        return;
      }
    }
    red_add_nc(p, a0);
    red_add_nc(p + 1, a2);
    p += Wp;
    red_add_nc(p, a1);
    red_add_nc(p + 1, a3);
  } else {
    const bool r0 = (unsigned)r < (unsigned)Hp, r1 = (unsigned)(r + 1) < (unsigned)Hp;
    const bool c0 = (unsigned)c < (unsigned)Wp, c1 = (unsigned)(c + 1) < (unsigned)Wp;
    T* p = iwe + ((int64_t)r * Wp + c);
    if (r0 & c0) red_add_nc(p, a0);
    if (r1 & c0) red_add_nc(p + Wp, a1);
    if (r0 & c1) red_add_nc(p + 1, a2);
    if (r1 & c1) red_add_nc(p + Wp + 1, a3);
  }
}

template <typename T, bool HAS_W, int EPT, bool VEC, bool PACKED>
__global__ void __launch_bounds__(256, (sizeof(T) == 4 && EPT <= 8) ? 4 : 1)
k_win_splat(const T* __restrict__ sx, const T* __restrict__ sy, const T* __restrict__ sd, const T* __restrict__ sw,
            int64_t n, const T* __restrict__ flow, int H, int W, int pad_h, int pad_w, T* __restrict__ iwe) {
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * EPT;
  if (base >= n) return;
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w, Hm1 = Hp - 1, Wm1 = Wp - 1;
  EventBlock<T, EPT, HAS_W, PACKED> e;
  e.load(sx, sy, sd, sw, base, n, flow, W, H * W);
  // current run: floor values of the cell (NaN = none) and the four tap sums
  T cfr = (T)NAN, cfc = (T)0;
  T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    const T xw = Rn<T>::sub(e.x[j], Rn<T>::mul(e.d[j], e.f0[j]));
    const T yw = Rn<T>::sub(e.y[j], Rn<T>::mul(e.d[j], e.f1[j]));
    const T xb = Rn<T>::add(xw, Rn<T>::bias()), yb = Rn<T>::add(yw, Rn<T>::bias());
    if (!FastFloor<T>::in_range(xb, yb)) {
      splat_event_exact<T>(iwe, Hp, Wp, pad_h, pad_w, e.x[j], xw, yw, HAS_W ? e.wt[j] : (T)1);
      continue;
    }
    const T fr = FastFloor<T>::flr(xb), fc = FastFloor<T>::flr(yb);
    const T a = Rn<T>::sub(xw, fr), b = Rn<T>::sub(yw, fc);
    const T na = Rn<T>::sub((T)1, a), nb = Rn<T>::sub((T)1, b);
    T w0 = Rn<T>::mul(na, nb), w1 = Rn<T>::mul(a, nb), w2 = Rn<T>::mul(na, b), w3 = Rn<T>::mul(a, b);
    if (HAS_W) {
      w0 = Rn<T>::mul(w0, e.wt[j]); w1 = Rn<T>::mul(w1, e.wt[j]);
      w2 = Rn<T>::mul(w2, e.wt[j]); w3 = Rn<T>::mul(w3, e.wt[j]);
    }
    const bool same = (fr == cfr) & (fc == cfc);
    if (!same) {
      if (cfr == cfr)
        flush_cell<T, VEC>(iwe, Hp, Wp, Hm1, Wm1, FastFloor<T>::to_int(cfr) + pad_h, FastFloor<T>::to_int(cfc) + pad_w,
                           a0, a1, a2, a3);
      cfr = fr; cfc = fc;
    }
    a0 = same ? a0 + w0 : w0;
    a1 = same ? a1 + w1 : w1;
    a2 = same ? a2 + w2 : w2;
    a3 = same ? a3 + w3 : w3;
  }
  if (cfr == cfr)
    flush_cell<T, VEC>(iwe, Hp, Wp, Hm1, Wm1, FastFloor<T>::to_int(cfr) + pad_h, FastFloor<T>::to_int(cfc) + pad_w, a0, a1,
                       a2, a3);
}

// ---- backward ------------------------------------------------------------------------------------
// GSRC 0: dL/dIWE read from a plane.  GSRC 1: variance objective, dL/dIWE = cv * (IWE - mean)
// derived on the fly from the IWE itself (saves writing and re-reading a gradient plane).
template <typename T> struct VarCoef { T mean, cv; int omit; };

template <typename T, int GSRC>
__device__ __forceinline__ T fetch_g(const T* __restrict__ g, int Hp, int Wp, int r, int c, const VarCoef<T>& vc) {
  if ((unsigned)r >= (unsigned)Hp || (unsigned)c >= (unsigned)Wp) return (T)0;
  T v = __ldg(g + (int64_t)r * Wp + c);
  if (GSRC == 1) {
    if (vc.omit && (r == 0 || c == 0 || r == Hp - 1 || c == Wp - 1)) return (T)0;
    v = vc.cv * (v - vc.mean);
  }
  return v;
}

// Exact handling of an event outside the fast range / on the image border: masked gathers.
template <typename T, int GSRC>
__device__ __noinline__ void bwd_event_exact(const T* __restrict__ g, int Hp, int Wp, int pad_h, int pad_w, T xw, T yw,
                                             VarCoef<T> vc, T& dx, T& dy) {
  dx = 0; dy = 0;
  if (!(Rn<T>::finite(xw) && Rn<T>::finite(yw))) return;  // all taps masked (or a skipped event): zero gradient
  const Taps<T> t = make_taps<T>(xw, yw, pad_h, pad_w);
  const bool r1ok = t.r != INT_MAX, c1ok = t.c != INT_MAX;
  const T g00 = fetch_g<T, GSRC>(g, Hp, Wp, t.r, t.c, vc);
  const T g10 = r1ok ? fetch_g<T, GSRC>(g, Hp, Wp, t.r + 1, t.c, vc) : (T)0;
  const T g01 = c1ok ? fetch_g<T, GSRC>(g, Hp, Wp, t.r, t.c + 1, vc) : (T)0;
  const T g11 = (r1ok && c1ok) ? fetch_g<T, GSRC>(g, Hp, Wp, t.r + 1, t.c + 1, vc) : (T)0;
  dx = ((T)1 - t.b) * (g10 - g00) + t.b * (g11 - g01);
  dy = ((T)1 - t.a) * (g01 - g00) + t.a * (g11 - g10);
}

// With a 16-byte aligned plane (VEC) the two column-adjacent taps of a row come from ONE aligned float4
// load (unless c % 4 == 3): 2 lane-loads per event instead of 4.
template <bool VEC>
__device__ __forceinline__ void load_pair(const float* __restrict__ p, int j, float& lo, float& hi) {
  if (VEC && j != 3) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p - j));
    lo = j == 0 ? v.x : (j == 1 ? v.y : v.z);
    hi = j == 0 ? v.y : (j == 1 ? v.z : v.w);
  } else {
    lo = __ldg(p);
    hi = __ldg(p + 1);
  }
}
template <bool VEC>
__device__ __forceinline__ void load_pair(const double* __restrict__ p, int, double& lo, double& hi) {
  lo = __ldg(p);
  hi = __ldg(p + 1);
}

template <typename T, int GSRC, bool HAS_W, int EPT, bool VEC, bool PACKED>
__global__ void __launch_bounds__(256, (sizeof(T) == 4 && EPT <= 4) ? 4 : 1)
k_win_bwd(const T* __restrict__ sx, const T* __restrict__ sy, const T* __restrict__ sd, const T* __restrict__ sw,
          int64_t n, const T* __restrict__ flow, int H, int W, int pad_h, int pad_w, const T* __restrict__ g,
          const double* __restrict__ acc, int omit, double scale, T* __restrict__ dflow) {
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * EPT;
  if (base >= n) return;
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  const int hw = H * W;
  VarCoef<T> vc{(T)0, (T)0, omit};
  if (GSRC == 1) {
    const double cnt = omit ? (double)(Hp - 2) * (double)(Wp - 2) : (double)Hp * (double)Wp;
    vc.mean = (T)(acc[0] / cnt);
    vc.cv = (T)(-2.0 * scale / (cnt - 1.0));
  }
  // fast range of cells whose four taps are all inside (and, for the cropped variance, all counted)
  const int lo = (GSRC == 1 && omit) ? 1 : 0;
  const unsigned r_span = (unsigned)max(Hp - 1 - 2 * lo, 0), c_span = (unsigned)max(Wp - 1 - 2 * lo, 0);
  EventBlock<T, EPT, HAS_W, PACKED> e;
  e.load(sx, sy, sd, sw, base, n, flow, W, hw);
  // phase 1: cells and fractions of all events, then all gathers of dL/dIWE in flight together
  T a[EPT], b[EPT], g00[EPT], g01[EPT], g10[EPT], g11[EPT];
  bool fast[EPT];
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    const T xw = Rn<T>::sub(e.x[j], Rn<T>::mul(e.d[j], e.f0[j]));
    const T yw = Rn<T>::sub(e.y[j], Rn<T>::mul(e.d[j], e.f1[j]));
    const T xb = Rn<T>::add(xw, Rn<T>::bias()), yb = Rn<T>::add(yw, Rn<T>::bias());
    fast[j] = FastFloor<T>::in_range(xb, yb);
    int r = 0, c = 0;
    if (fast[j]) {
      const T fr = FastFloor<T>::flr(xb), fc = FastFloor<T>::flr(yb);
      a[j] = Rn<T>::sub(xw, fr);
      b[j] = Rn<T>::sub(yw, fc);
      r = FastFloor<T>::to_int(fr) + pad_h;
      c = FastFloor<T>::to_int(fc) + pad_w;
      fast[j] = (unsigned)(r - lo) < r_span && (unsigned)(c - lo) < c_span;
    }
    if (fast[j]) {
      const T* p = g + (r * Wp + c);
      load_pair<VEC>(p, c & 3, g00[j], g01[j]);
      load_pair<VEC>(p + Wp, c & 3, g10[j], g11[j]);
    } else {
      // rare: border cell, out-of-range or skipped event -- exact masked gathers, result kept in g00/g01
      T dx, dy;
      bwd_event_exact<T, GSRC>(g, Hp, Wp, pad_h, pad_w, e.x[j] == e.x[j] ? xw : (T)NAN, yw, vc, dx, dy);
      g00[j] = dx; g01[j] = dy; g10[j] = 0; g11[j] = 0; a[j] = 0; b[j] = 0;
    }
  }
  // phase 2: gradients, combined over runs of the same origin pixel
  int ck = -1;
  T s0 = 0, s1 = 0;
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    T dx, dy;
    if (fast[j]) {
      dx = ((T)1 - b[j]) * (g10[j] - g00[j]) + b[j] * (g11[j] - g01[j]);
      dy = ((T)1 - a[j]) * (g01[j] - g00[j]) + a[j] * (g11[j] - g10[j]);
      if (GSRC == 1) { dx *= vc.cv; dy *= vc.cv; }  // differences: the mean cancels
    } else {
      dx = g00[j]; dy = g01[j];
    }
    if (HAS_W) { dx *= e.wt[j]; dy *= e.wt[j]; }
    if (e.x[j] != e.x[j]) continue;  // skipped event
    if (e.k[j] != ck) {
      if (ck >= 0) { red_add_nc(dflow + ck, s0); red_add_nc(dflow + hw + ck, s1); }
      ck = e.k[j]; s0 = 0; s1 = 0;
    }
    s0 -= e.d[j] * dx;
    s1 -= e.d[j] * dy;
  }
  if (ck >= 0) { red_add_nc(dflow + ck, s0); red_add_nc(dflow + hw + ck, s1); }
}

// ---- host side ------------------------------------------------------------------------------------
static int key_bits_for(int64_t hw) {
  int bits = 1;
  while (((int64_t)1 << bits) < hw + 1 && bits < 32) ++bits;
  return bits;
}

struct PrepWs { size_t off_kin, off_kout, off_iin, off_cub, cub_bytes, total; };
static PrepWs prep_ws(int64_t n) {
  PrepWs w;
  size_t a = align256((size_t)std::max<int64_t>(n, 1) * 4);
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned int*)nullptr, (unsigned int*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)std::max<int64_t>(n, 1));
  w.off_kin = 0; w.off_kout = a; w.off_iin = 2 * a; w.off_cub = 3 * a;
  w.cub_bytes = cub_bytes;
  w.total = 3 * a + align256(cub_bytes) + 256;
  return w;
}

template <typename T>
int window_prepare_impl(const T* events, int64_t n, int H, int W, int direction, double direction_frac, int normalize_t,
                        const T* weight, const T* tminmax, int allow_packed, void* window, void* workspace,
                        size_t workspace_bytes, int32_t* status, cudaStream_t st) {
  WindowHeader* hdr = reinterpret_cast<WindowHeader*>(window);
  k_win_hdr_init<<<1, 1, 0, st>>>(hdr);
  if (n == 0) {
    k_win_hdr_final<T><<<1, 1, 0, st>>>(hdr, direction, direction_frac, tminmax);
    EBOS_LAUNCH_CHECK("ebos_window_prepare");
    return EBOS_OK;
  }
  PrepWs ws = prep_ws(n);
  if (!workspace || workspace_bytes < ws.total) { set_error("ebos_window_prepare: workspace too small"); return EBOS_ERR_WORKSPACE; }
  char* wp = reinterpret_cast<char*>(align256(reinterpret_cast<size_t>(workspace)));
  unsigned int* k_in = reinterpret_cast<unsigned int*>(wp + ws.off_kin);
  unsigned int* k_out = reinterpret_cast<unsigned int*>(wp + ws.off_kout);
  int* i_in = reinterpret_cast<int*>(wp + ws.off_iin);
  void* cub_tmp = wp + ws.off_cub;
  WindowLayout L = window_layout(n, sizeof(T));
  char* b = reinterpret_cast<char*>(window);
  int* perm = reinterpret_cast<int*>(b + L.off_perm);
  int bx = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
  if (!tminmax) k_win_minmax<T><<<bx, 256, 0, st>>>(events, n, hdr);
  k_win_hdr_final<T><<<1, 1, 0, st>>>(hdr, direction, direction_frac, tminmax);
  k_win_keys<T><<<bx, 256, 0, st>>>(events, n, H, W, k_in, i_in, status, hdr);
  k_win_layout<<<1, 1, 0, st>>>(hdr, allow_packed && sizeof(T) == 4 && H <= 65536 && W <= 65536, status);
  size_t cub_bytes = ws.cub_bytes;
  cudaError_t ce = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, k_in, k_out, i_in, perm, (int)n, 0,
                                                   key_bits_for((int64_t)H * W), st);
  if (ce != cudaSuccess) return cuda_fail(ce, "ebos_window_prepare(sort)");
  k_win_gather<T><<<bx, 256, 0, st>>>(events, weight, n, H, W, perm, hdr, normalize_t,
                                      reinterpret_cast<T*>(b + L.off_x), reinterpret_cast<T*>(b + L.off_y),
                                      reinterpret_cast<T*>(b + L.off_d),
                                      weight ? reinterpret_cast<T*>(b + L.off_w) : nullptr);
  EBOS_LAUNCH_CHECK("ebos_window_prepare");
  return EBOS_OK;
}

template <typename T>
int window_splat_t(const void* window, int64_t n, int flags, const T* flow, int H, int W, int pad_h, int pad_w, T* iwe,
                   cudaStream_t st) {
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  cudaError_t e = cudaMemsetAsync(iwe, 0, (size_t)Hp * Wp * sizeof(T), st);
  if (e != cudaSuccess) return cuda_fail(e, "ebos_window_splat memset");
  if (n == 0) return EBOS_OK;
  const bool has_weight = flags & EBOS_WIN_HAS_WEIGHT, packed = flags & EBOS_WIN_PACKED;
  if (packed && sizeof(T) != 4) { set_error("ebos_window_splat: the packed layout exists for fp32 windows only"); return EBOS_ERR_BAD_ARG; }
  WindowLayout L = window_layout(n, sizeof(T));
  const char* b = reinterpret_cast<const char*>(window);
  const T* sx = reinterpret_cast<const T*>(b + L.off_x);
  const T* sy = reinterpret_cast<const T*>(b + L.off_y);
  const T* sd = reinterpret_cast<const T*>(b + L.off_d);
  const T* sw = reinterpret_cast<const T*>(b + L.off_w);
  static const int ept_env = env_int("EBOS_SPLAT_EPT");
  const int ept = (ept_env == 4 || ept_env == 8) ? ept_env : Ept<T>::splat;
  int64_t threads = (n + ept - 1) / ept;
  unsigned grid = (unsigned)((threads + 255) / 256);
  // vector REDs need fp32, a 16-byte aligned plane and a row stride that keeps the alignment
  static const int novec_env = env_int("EBOS_NO_VEC");
  const bool vec = sizeof(T) == 4 && !novec_env && (Wp % 4 == 0) && ((reinterpret_cast<size_t>(iwe) & 15) == 0);
#define EBOS_SPLAT_K(WGT, E, V, P) k_win_splat<T, WGT, E, V, P><<<grid, 256, 0, st>>>(sx, sy, sd, sw, n, flow, H, W, pad_h, pad_w, iwe)
#define EBOS_SPLAT_E(WGT, V, P) do { if (ept == 4) EBOS_SPLAT_K(WGT, 4, V, P); else EBOS_SPLAT_K(WGT, 8, V, P); } while (0)
  if constexpr (sizeof(T) == 4) {
    if (has_weight) {   // weighted windows: generic code path only
      if (packed) { if (vec) EBOS_SPLAT_E(true, true, true); else EBOS_SPLAT_E(true, false, true); }
      else { if (vec) EBOS_SPLAT_E(true, true, false); else EBOS_SPLAT_E(true, false, false); }
    } else {
      if (packed) { if (vec) EBOS_SPLAT_E(false, true, true); else EBOS_SPLAT_E(false, false, true); }
      else { if (vec) EBOS_SPLAT_E(false, true, false); else EBOS_SPLAT_E(false, false, false); }
    }
  } else {
    if (has_weight) EBOS_SPLAT_K(true, 4, false, false); else EBOS_SPLAT_K(false, 4, false, false);
  }
#undef EBOS_SPLAT_E
#undef EBOS_SPLAT_K
  EBOS_LAUNCH_CHECK("ebos_window_splat");
  return EBOS_OK;
}

template <typename T>
int window_backward_t(const void* window, int64_t n, int flags, const T* flow, int H, int W, int pad_h, int pad_w,
                      const T* grad_iwe, int kind, const T* iwe, const double* acc, int omit_boundary, double scale,
                      T* dflow, cudaStream_t st) {
  if (n == 0) return EBOS_OK;
  const bool has_weight = flags & EBOS_WIN_HAS_WEIGHT, packed = flags & EBOS_WIN_PACKED;
  if (packed && sizeof(T) != 4) { set_error("ebos_window_backward: the packed layout exists for fp32 windows only"); return EBOS_ERR_BAD_ARG; }
  WindowLayout L = window_layout(n, sizeof(T));
  const char* b = reinterpret_cast<const char*>(window);
  const T* sx = reinterpret_cast<const T*>(b + L.off_x);
  const T* sy = reinterpret_cast<const T*>(b + L.off_y);
  const T* sd = reinterpret_cast<const T*>(b + L.off_d);
  const T* sw = reinterpret_cast<const T*>(b + L.off_w);
  const int ept = Ept<T>::bwd;
  int64_t threads = (n + ept - 1) / ept;
  unsigned grid = (unsigned)((threads + 255) / 256);
  const bool affine = grad_iwe == nullptr;
  if (affine && (kind != EBOS_COST_VARIANCE || !iwe || !acc)) {
    set_error("ebos_window_backward: grad_iwe == NULL needs kind == VARIANCE with iwe and acc");
    return EBOS_ERR_BAD_ARG;
  }
  const T* gsrc = affine ? iwe : grad_iwe;
  static const int novec_env = env_int("EBOS_NO_VEC_LOAD");
  const int Wp = W + 2 * pad_w;
  const bool vec = sizeof(T) == 4 && !novec_env && (Wp % 4 == 0) && ((reinterpret_cast<size_t>(gsrc) & 15) == 0);
#define EBOS_BWD_K(G, WGT, V, P) k_win_bwd<T, G, WGT, Ept<T>::bwd, V, P><<<grid, 256, 0, st>>>(sx, sy, sd, sw, n, flow, H, W, pad_h, pad_w, gsrc, acc, omit_boundary, scale, dflow)
#define EBOS_BWD_G(G)                                                                                              \
  do {                                                                                                             \
    if constexpr (sizeof(T) == 4) {                                                                                \
      if (has_weight) { if (packed) { if (vec) EBOS_BWD_K(G, true, true, true); else EBOS_BWD_K(G, true, false, true); }          \
                        else { if (vec) EBOS_BWD_K(G, true, true, false); else EBOS_BWD_K(G, true, false, false); } }             \
      else { if (packed) { if (vec) EBOS_BWD_K(G, false, true, true); else EBOS_BWD_K(G, false, false, true); }                   \
             else { if (vec) EBOS_BWD_K(G, false, true, false); else EBOS_BWD_K(G, false, false, false); } }                      \
    } else {                                                                                                       \
      if (has_weight) EBOS_BWD_K(G, true, false, false); else EBOS_BWD_K(G, false, false, false);                  \
    }                                                                                                              \
  } while (0)
  if (affine) EBOS_BWD_G(1); else EBOS_BWD_G(0);
#undef EBOS_BWD_G
#undef EBOS_BWD_K
  EBOS_LAUNCH_CHECK("ebos_window_backward");
  return EBOS_OK;
}

// type-erased entry points used by ebos_costs.cu (fused iteration)
int window_splat_launch(const void* window, int64_t n, int flags, const void* flow, int H, int W, int pad_h,
                        int pad_w, int dtype, void* iwe, cudaStream_t st) {
  if (dtype == EBOS_F64)
    return window_splat_t<double>(window, n, flags, (const double*)flow, H, W, pad_h, pad_w, (double*)iwe, st);
  return window_splat_t<float>(window, n, flags, (const float*)flow, H, W, pad_h, pad_w, (float*)iwe, st);
}
int window_backward_launch(const void* window, int64_t n, int flags, const void* flow, int H, int W, int pad_h,
                           int pad_w, int dtype, const void* grad_iwe, int kind, const void* iwe, const double* acc,
                           int omit_boundary, double scale, void* dflow, cudaStream_t st) {
  if (dtype == EBOS_F64)
    return window_backward_t<double>(window, n, flags, (const double*)flow, H, W, pad_h, pad_w,
                                     (const double*)grad_iwe, kind, (const double*)iwe, acc, omit_boundary, scale,
                                     (double*)dflow, st);
  return window_backward_t<float>(window, n, flags, (const float*)flow, H, W, pad_h, pad_w, (const float*)grad_iwe,
                                  kind, (const float*)iwe, acc, omit_boundary, scale, (float*)dflow, st);
}

}  // namespace ebos

using namespace ebos;

#define EBOS_CHECK_DTYPE(dtype, who)                                              \
  do {                                                                            \
    if ((dtype) != EBOS_F32 && (dtype) != EBOS_F64) {                             \
      ebos::set_error(who ": unsupported dtype");                                 \
      return EBOS_ERR_UNSUPPORTED;                                                \
    }                                                                             \
  } while (0)

extern "C" {

size_t ebos_window_bytes(int64_t n, int dtype) { return n < 0 ? 0 : window_layout(n, dtype_size(dtype)).total; }

size_t ebos_window_workspace_bytes(int64_t n, int H, int W) {
  (void)H; (void)W;
  return n < 0 ? 0 : prep_ws(n).total;
}

int ebos_window_prepare(const void* events, int64_t n, int H, int W, int direction, double direction_frac,
                        int normalize_t, const void* weight, const void* tminmax, int allow_packed, int dtype,
                        void* window, void* workspace, size_t workspace_bytes, int32_t* status, void* stream) {
  EBOS_REQUIRE(n >= 0 && n < (int64_t)INT_MAX && H > 0 && W > 0 && window && status && (n == 0 || events),
               "ebos_window_prepare: bad argument");
  EBOS_REQUIRE((int64_t)H * W < ((int64_t)1 << 31) - 1, "ebos_window_prepare: grid too large");
  EBOS_REQUIRE(direction >= EBOS_DIR_FIRST && direction <= EBOS_DIR_FRAC, "ebos_window_prepare: bad direction");
  EBOS_REQUIRE((reinterpret_cast<size_t>(window) & 255) == 0, "ebos_window_prepare: window buffer must be 256-byte aligned");
  EBOS_CHECK_DTYPE(dtype, "ebos_window_prepare");
  if (dtype == EBOS_F64)
    return window_prepare_impl<double>((const double*)events, n, H, W, direction, direction_frac, normalize_t,
                                       (const double*)weight, (const double*)tminmax, 0, window, workspace,
                                       workspace_bytes, status, as_stream(stream));
  return window_prepare_impl<float>((const float*)events, n, H, W, direction, direction_frac, normalize_t,
                                    (const float*)weight, (const float*)tminmax, allow_packed, window, workspace,
                                    workspace_bytes, status, as_stream(stream));
}

int ebos_window_info(const void* window, int64_t n, int dtype, int32_t* perm_out, double* tinfo_out, void* stream) {
  EBOS_REQUIRE(window && n >= 0, "ebos_window_info: bad argument");
  EBOS_CHECK_DTYPE(dtype, "ebos_window_info");
  cudaStream_t st = as_stream(stream);
  const char* b = reinterpret_cast<const char*>(window);
  if (perm_out && n > 0) {
    cudaError_t e = cudaMemcpyAsync(perm_out, b + window_layout(n, dtype_size(dtype)).off_perm, (size_t)n * 4,
                                    cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_window_info(perm)");
  }
  if (tinfo_out) {
    cudaError_t e = cudaMemcpyAsync(tinfo_out, b, 32, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_window_info(tinfo)");
  }
  return EBOS_OK;
}

int ebos_window_splat(const void* window, int64_t n, int flags, const void* flow, int H, int W, int pad_h,
                      int pad_w, int dtype, void* iwe, void* stream) {
  EBOS_REQUIRE(window && flow && iwe && n >= 0 && H > 0 && W > 0 && pad_h >= 0 && pad_w >= 0, "ebos_window_splat: bad argument");
  EBOS_CHECK_DTYPE(dtype, "ebos_window_splat");
  return window_splat_launch(window, n, flags, flow, H, W, pad_h, pad_w, dtype, iwe, as_stream(stream));
}

int ebos_window_backward(const void* window, int64_t n, int flags, const void* flow, int H, int W, int pad_h,
                         int pad_w, int dtype, const void* grad_iwe, int kind, const void* iwe, const double* acc,
                         int omit_boundary, double scale, void* dflow, void* stream) {
  EBOS_REQUIRE(window && flow && dflow && n >= 0 && H > 0 && W > 0 && pad_h >= 0 && pad_w >= 0, "ebos_window_backward: bad argument");
  EBOS_CHECK_DTYPE(dtype, "ebos_window_backward");
  return window_backward_launch(window, n, flags, flow, H, W, pad_h, pad_w, dtype, grad_iwe, kind, iwe, acc,
                                omit_boundary, scale, dflow, as_stream(stream));
}

}  // extern "C"
