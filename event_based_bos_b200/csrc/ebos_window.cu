// Fused contrast-maximisation path on a prepared window (fp32).
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
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

#include "ebos_common.cuh"

namespace ebos {

constexpr int kSplatEpt = 8;  // events per thread, forward
constexpr int kBwdEpt = 4;    // events per thread, backward

// ---- prepare --------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int enc_f32(float f) {
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_win_hdr_init(WindowHeader* h) {
  h->enc_min = 0xffffffffu;
  h->enc_max = 0u;
}

__global__ void __launch_bounds__(256) k_win_minmax(const float* __restrict__ ev, int64_t n, WindowHeader* h) {
  unsigned int lo = 0xffffffffu, hi = 0u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned int u = enc_f32(__ldg(ev + 4 * i + 2));
    lo = min(lo, u);
    hi = max(hi, u);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&h->enc_min, lo);
    atomicMax(&h->enc_max, hi);
  }
}

__global__ void k_win_hdr_final(WindowHeader* h, int direction, double frac, const float* __restrict__ tminmax) {
  float tmin = tminmax ? tminmax[0] : dec_f32(h->enc_min), tmax = tminmax ? tminmax[1] : dec_f32(h->enc_max);
  TimeRef<float> tr = make_time_ref<float>(tmin, tmax, direction, frac);
  h->t_min = tmin;
  h->t_max = tmax;
  h->t_ref = tr.t_ref;
  h->period = tr.period;
}

// key = origin pixel k (src/warp.py:334); an event whose k is outside the grid is flagged and
// parked behind all valid pixels.
__global__ void __launch_bounds__(256) k_win_keys(const float* __restrict__ ev, int64_t n, int H, int W,
                                                  unsigned int* __restrict__ keys, int* __restrict__ idx,
                                                  int32_t* __restrict__ status) {
  const int64_t hw = (int64_t)H * W;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float2 xy = __ldg(reinterpret_cast<const float2*>(ev) + 2 * i);
    int64_t k = (int64_t)xy.x * W + (int64_t)xy.y;
    bool ok = k >= 0 && k < hw && fabsf(xy.x) <= FLT_MAX && fabsf(xy.y) <= FLT_MAX;
    bad |= !ok;
    keys[i] = ok ? (unsigned int)k : (unsigned int)hw;
    idx[i] = (int)i;
  }
  if (bad) atomicOr(status, EBOS_STATUS_PIXEL_OOB);
}

__global__ void __launch_bounds__(256) k_win_gather(const float* __restrict__ ev, const float* __restrict__ weight,
                                                    int64_t n, int H, int W, const int* __restrict__ perm,
                                                    const WindowHeader* __restrict__ h, int normalize_t,
                                                    float* __restrict__ sx, float* __restrict__ sy,
                                                    float* __restrict__ sd, float* __restrict__ sw) {
  TimeRef<float> tr{h->t_ref, h->period};
  const int64_t hw = (int64_t)H * W;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    int i = __ldg(perm + j);
    float4 e = __ldg(reinterpret_cast<const float4*>(ev) + i);
    int64_t k = (int64_t)e.x * W + (int64_t)e.y;
    bool ok = k >= 0 && k < hw && fabsf(e.x) <= FLT_MAX && fabsf(e.y) <= FLT_MAX;
    // invalid events are parked at a coordinate whose k is negative: the kernels skip them.
    sx[j] = ok ? e.x : -2.0f;
    sy[j] = ok ? e.y : -2.0f;
    sd[j] = event_dt<float>(e.z, tr, normalize_t);
    if (sw) sw[j] = __ldg(weight + i);
  }
}

// ---- forward: fused warp + bilinear vote -----------------------------------------------------
__device__ __forceinline__ void flush_cell(float* __restrict__ iwe, int Hp, int Wp, int r, int c, float a0, float a1,
                                           float a2, float a3) {
  const bool r0 = (unsigned)r < (unsigned)Hp, r1 = (unsigned)(r + 1) < (unsigned)Hp;
  const bool c0 = (unsigned)c < (unsigned)Wp, c1 = (unsigned)(c + 1) < (unsigned)Wp;
  float* p = iwe + (int64_t)r * Wp + c;
  if (r0 & c0) red_add(p, a0);
  if (r1 & c0) red_add(p + Wp, a1);
  if (r0 & c1) red_add(p + 1, a2);
  if (r1 & c1) red_add(p + Wp + 1, a3);
}

template <bool HAS_W>
__global__ void __launch_bounds__(256) k_win_splat(const float* __restrict__ sx, const float* __restrict__ sy,
                                                   const float* __restrict__ sd, const float* __restrict__ sw,
                                                   int64_t n, const float* __restrict__ flow, int H, int W, int pad_h,
                                                   int pad_w, float* __restrict__ iwe) {
  constexpr int EPT = kSplatEpt;
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * EPT;
  if (base >= n) return;
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  const int hw = H * W;
  float x[EPT], y[EPT], d[EPT], wt[EPT];
  if (base + EPT <= n) {
#pragma unroll
    for (int j = 0; j < EPT; j += 4) {
      float4 vx = __ldg(reinterpret_cast<const float4*>(sx + base + j));
      float4 vy = __ldg(reinterpret_cast<const float4*>(sy + base + j));
      float4 vd = __ldg(reinterpret_cast<const float4*>(sd + base + j));
      x[j] = vx.x; x[j + 1] = vx.y; x[j + 2] = vx.z; x[j + 3] = vx.w;
      y[j] = vy.x; y[j + 1] = vy.y; y[j + 2] = vy.z; y[j + 3] = vy.w;
      d[j] = vd.x; d[j + 1] = vd.y; d[j + 2] = vd.z; d[j + 3] = vd.w;
      if (HAS_W) {
        float4 vw = __ldg(reinterpret_cast<const float4*>(sw + base + j));
        wt[j] = vw.x; wt[j + 1] = vw.y; wt[j + 2] = vw.z; wt[j + 3] = vw.w;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const bool in = base + j < n;
      x[j] = in ? sx[base + j] : -2.0f;
      y[j] = in ? sy[base + j] : -2.0f;
      d[j] = in ? sd[base + j] : 0.0f;
      if (HAS_W) wt[j] = in ? sw[base + j] : 0.0f;
    }
  }
  int cr = INT_MIN, cc = INT_MIN;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    const int k = (int)x[j] * W + (int)y[j];
    if ((unsigned)k >= (unsigned)hw) continue;  // parked (invalid) event or tail
    const float xw = __fsub_rn(x[j], __fmul_rn(d[j], __ldg(flow + k)));
    const float yw = __fsub_rn(y[j], __fmul_rn(d[j], __ldg(flow + hw + k)));
    Taps<float> t = make_taps<float>(xw, yw, pad_h, pad_w);
    if (HAS_W) {
      t.w0 = __fmul_rn(t.w0, wt[j]); t.w1 = __fmul_rn(t.w1, wt[j]);
      t.w2 = __fmul_rn(t.w2, wt[j]); t.w3 = __fmul_rn(t.w3, wt[j]);
    }
    if (!(fabsf(xw) <= FLT_MAX && fabsf(yw) <= FLT_MAX)) {
      // non-finite warped coordinate (e.g. zero-length window: dt = 0/0): the reference masks all
      // four taps and adds vals*0 = NaN to pixel 0.
      red_add(iwe, __fmul_rn(t.w0, 0.0f));
      continue;
    }
    if (t.r != cr || t.c != cc) {
      if (cr != INT_MIN) flush_cell(iwe, Hp, Wp, cr, cc, a0, a1, a2, a3);
      cr = t.r; cc = t.c;
      a0 = t.w0; a1 = t.w1; a2 = t.w2; a3 = t.w3;
    } else {
      a0 += t.w0; a1 += t.w1; a2 += t.w2; a3 += t.w3;
    }
  }
  if (cr != INT_MIN) flush_cell(iwe, Hp, Wp, cr, cc, a0, a1, a2, a3);
}

// ---- backward ------------------------------------------------------------------------------------
// GSRC 0: dL/dIWE read from a plane.  GSRC 1: variance objective, dL/dIWE = cv * (IWE - mean)
// derived on the fly from the IWE itself (saves writing and re-reading a gradient plane).
struct VarCoef { float mean, cv; int omit; };

template <int GSRC>
__device__ __forceinline__ float fetch_g(const float* __restrict__ g, int Hp, int Wp, int r, int c, const VarCoef& vc) {
  if ((unsigned)r >= (unsigned)Hp || (unsigned)c >= (unsigned)Wp) return 0.f;
  float v = __ldg(g + (int64_t)r * Wp + c);
  if (GSRC == 1) {
    if (vc.omit && (r == 0 || c == 0 || r == Hp - 1 || c == Wp - 1)) return 0.f;
    v = vc.cv * (v - vc.mean);
  }
  return v;
}

template <int GSRC, bool HAS_W>
__global__ void __launch_bounds__(256) k_win_bwd(const float* __restrict__ sx, const float* __restrict__ sy,
                                                 const float* __restrict__ sd, const float* __restrict__ sw, int64_t n,
                                                 const float* __restrict__ flow, int H, int W, int pad_h, int pad_w,
                                                 const float* __restrict__ g, const double* __restrict__ acc, int omit,
                                                 float scale, float* __restrict__ dflow) {
  constexpr int EPT = kBwdEpt;
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * EPT;
  if (base >= n) return;
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  const int hw = H * W;
  VarCoef vc{0.f, 0.f, omit};
  if (GSRC == 1) {
    const double cnt = omit ? (double)(Hp - 2) * (double)(Wp - 2) : (double)Hp * (double)Wp;
    const double mean = acc[0] / cnt;
    vc.mean = (float)mean;
    vc.cv = (float)(-2.0 * (double)scale / (cnt - 1.0));
  }
  float x[EPT], y[EPT], d[EPT], wt[EPT];
  if (base + EPT <= n) {
    float4 vx = __ldg(reinterpret_cast<const float4*>(sx + base));
    float4 vy = __ldg(reinterpret_cast<const float4*>(sy + base));
    float4 vd = __ldg(reinterpret_cast<const float4*>(sd + base));
    x[0] = vx.x; x[1] = vx.y; x[2] = vx.z; x[3] = vx.w;
    y[0] = vy.x; y[1] = vy.y; y[2] = vy.z; y[3] = vy.w;
    d[0] = vd.x; d[1] = vd.y; d[2] = vd.z; d[3] = vd.w;
    if (HAS_W) {
      float4 vw = __ldg(reinterpret_cast<const float4*>(sw + base));
      wt[0] = vw.x; wt[1] = vw.y; wt[2] = vw.z; wt[3] = vw.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const bool in = base + j < n;
      x[j] = in ? sx[base + j] : -2.0f;
      y[j] = in ? sy[base + j] : -2.0f;
      d[j] = in ? sd[base + j] : 0.0f;
      if (HAS_W) wt[j] = in ? sw[base + j] : 0.0f;
    }
  }
  int ck = -1;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    const int k = (int)x[j] * W + (int)y[j];
    if ((unsigned)k >= (unsigned)hw) continue;
    const float xw = __fsub_rn(x[j], __fmul_rn(d[j], __ldg(flow + k)));
    const float yw = __fsub_rn(y[j], __fmul_rn(d[j], __ldg(flow + hw + k)));
    if (!(fabsf(xw) <= FLT_MAX && fabsf(yw) <= FLT_MAX)) continue;  // all taps masked: zero gradient
    const Taps<float> t = make_taps<float>(xw, yw, pad_h, pad_w);
    const float g00 = fetch_g<GSRC>(g, Hp, Wp, t.r, t.c, vc);
    const float g10 = fetch_g<GSRC>(g, Hp, Wp, t.r + 1, t.c, vc);
    const float g01 = fetch_g<GSRC>(g, Hp, Wp, t.r, t.c + 1, vc);
    const float g11 = fetch_g<GSRC>(g, Hp, Wp, t.r + 1, t.c + 1, vc);
    float dx = (1.f - t.b) * (g10 - g00) + t.b * (g11 - g01);
    float dy = (1.f - t.a) * (g01 - g00) + t.a * (g11 - g10);
    if (HAS_W) { dx *= wt[j]; dy *= wt[j]; }
    if (k != ck) {
      if (ck >= 0) { red_add(dflow + ck, s0); red_add(dflow + hw + ck, s1); }
      ck = k; s0 = 0.f; s1 = 0.f;
    }
    s0 -= d[j] * dx;
    s1 -= d[j] * dy;
  }
  if (ck >= 0) { red_add(dflow + ck, s0); red_add(dflow + hw + ck, s1); }
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

int window_splat_launch(const void* window, int64_t n, int has_weight, const float* flow, int H, int W, int pad_h,
                        int pad_w, float* iwe, cudaStream_t st) {
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  cudaError_t e = cudaMemsetAsync(iwe, 0, (size_t)Hp * Wp * sizeof(float), st);
  if (e != cudaSuccess) return cuda_fail(e, "ebos_window_splat memset");
  if (n == 0) return EBOS_OK;
  WindowLayout L = window_layout(n);
  const char* b = reinterpret_cast<const char*>(window);
  const float* sx = reinterpret_cast<const float*>(b + L.off_x);
  const float* sy = reinterpret_cast<const float*>(b + L.off_y);
  const float* sd = reinterpret_cast<const float*>(b + L.off_d);
  const float* sw = reinterpret_cast<const float*>(b + L.off_w);
  int64_t threads = (n + kSplatEpt - 1) / kSplatEpt;
  unsigned grid = (unsigned)((threads + 255) / 256);
  if (has_weight) k_win_splat<true><<<grid, 256, 0, st>>>(sx, sy, sd, sw, n, flow, H, W, pad_h, pad_w, iwe);
  else k_win_splat<false><<<grid, 256, 0, st>>>(sx, sy, sd, sw, n, flow, H, W, pad_h, pad_w, iwe);
  EBOS_LAUNCH_CHECK("ebos_window_splat");
  return EBOS_OK;
}

int window_backward_launch(const void* window, int64_t n, int has_weight, const float* flow, int H, int W, int pad_h,
                           int pad_w, const float* grad_iwe, int kind, const float* iwe, const double* acc,
                           int omit_boundary, float scale, float* dflow, cudaStream_t st) {
  if (n == 0) return EBOS_OK;
  WindowLayout L = window_layout(n);
  const char* b = reinterpret_cast<const char*>(window);
  const float* sx = reinterpret_cast<const float*>(b + L.off_x);
  const float* sy = reinterpret_cast<const float*>(b + L.off_y);
  const float* sd = reinterpret_cast<const float*>(b + L.off_d);
  const float* sw = reinterpret_cast<const float*>(b + L.off_w);
  int64_t threads = (n + kBwdEpt - 1) / kBwdEpt;
  unsigned grid = (unsigned)((threads + 255) / 256);
  const bool affine = grad_iwe == nullptr;
  if (affine) {
    if (kind != EBOS_COST_VARIANCE || !iwe || !acc) {
      set_error("ebos_window_backward: grad_iwe == NULL needs kind == VARIANCE with iwe and acc");
      return EBOS_ERR_BAD_ARG;
    }
    if (has_weight) k_win_bwd<1, true><<<grid, 256, 0, st>>>(sx, sy, sd, sw, n, flow, H, W, pad_h, pad_w, iwe, acc, omit_boundary, scale, dflow);
    else k_win_bwd<1, false><<<grid, 256, 0, st>>>(sx, sy, sd, sw, n, flow, H, W, pad_h, pad_w, iwe, acc, omit_boundary, scale, dflow);
  } else {
    if (has_weight) k_win_bwd<0, true><<<grid, 256, 0, st>>>(sx, sy, sd, sw, n, flow, H, W, pad_h, pad_w, grad_iwe, acc, omit_boundary, scale, dflow);
    else k_win_bwd<0, false><<<grid, 256, 0, st>>>(sx, sy, sd, sw, n, flow, H, W, pad_h, pad_w, grad_iwe, acc, omit_boundary, scale, dflow);
  }
  EBOS_LAUNCH_CHECK("ebos_window_backward");
  return EBOS_OK;
}

}  // namespace ebos

using namespace ebos;

extern "C" {

size_t ebos_window_bytes(int64_t n) { return n < 0 ? 0 : window_layout(n).total; }

size_t ebos_window_workspace_bytes(int64_t n, int H, int W) {
  (void)H; (void)W;
  return n < 0 ? 0 : prep_ws(n).total;
}

int ebos_window_prepare(const float* events, int64_t n, int H, int W, int direction, double direction_frac,
                        int normalize_t, const float* weight, const float* tminmax, void* window, void* workspace,
                        size_t workspace_bytes, int32_t* status, void* stream) {
  EBOS_REQUIRE(n >= 0 && n < (int64_t)INT_MAX && H > 0 && W > 0 && window && status && (n == 0 || events),
               "ebos_window_prepare: bad argument");
  EBOS_REQUIRE((int64_t)H * W < ((int64_t)1 << 31) - 1, "ebos_window_prepare: grid too large");
  EBOS_REQUIRE(direction >= EBOS_DIR_FIRST && direction <= EBOS_DIR_FRAC, "ebos_window_prepare: bad direction");
  EBOS_REQUIRE((reinterpret_cast<size_t>(window) & 255) == 0, "ebos_window_prepare: window buffer must be 256-byte aligned");
  cudaStream_t st = as_stream(stream);
  WindowHeader* hdr = reinterpret_cast<WindowHeader*>(window);
  k_win_hdr_init<<<1, 1, 0, st>>>(hdr);
  if (n == 0) {
    k_win_hdr_final<<<1, 1, 0, st>>>(hdr, direction, direction_frac, tminmax);
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
  WindowLayout L = window_layout(n);
  char* b = reinterpret_cast<char*>(window);
  int* perm = reinterpret_cast<int*>(b + L.off_perm);
  int bx = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
  if (!tminmax) k_win_minmax<<<bx, 256, 0, st>>>(events, n, hdr);
  k_win_hdr_final<<<1, 1, 0, st>>>(hdr, direction, direction_frac, tminmax);
  k_win_keys<<<bx, 256, 0, st>>>(events, n, H, W, k_in, i_in, status);
  size_t cub_bytes = ws.cub_bytes;
  cudaError_t ce = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, k_in, k_out, i_in, perm, (int)n, 0,
                                                   key_bits_for((int64_t)H * W), st);
  if (ce != cudaSuccess) return cuda_fail(ce, "ebos_window_prepare(sort)");
  k_win_gather<<<bx, 256, 0, st>>>(events, weight, n, H, W, perm, hdr, normalize_t,
                                   reinterpret_cast<float*>(b + L.off_x), reinterpret_cast<float*>(b + L.off_y),
                                   reinterpret_cast<float*>(b + L.off_d),
                                   weight ? reinterpret_cast<float*>(b + L.off_w) : nullptr);
  EBOS_LAUNCH_CHECK("ebos_window_prepare");
  return EBOS_OK;
}

int ebos_window_info(const void* window, int64_t n, int32_t* perm_out, float* tinfo_out, void* stream) {
  EBOS_REQUIRE(window && n >= 0, "ebos_window_info: bad argument");
  cudaStream_t st = as_stream(stream);
  const char* b = reinterpret_cast<const char*>(window);
  if (perm_out && n > 0) {
    cudaError_t e = cudaMemcpyAsync(perm_out, b + window_layout(n).off_perm, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_window_info(perm)");
  }
  if (tinfo_out) {
    cudaError_t e = cudaMemcpyAsync(tinfo_out, b, 16, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_window_info(tinfo)");
  }
  return EBOS_OK;
}

int ebos_window_splat(const void* window, int64_t n, int has_weight, const float* flow, int H, int W, int pad_h,
                      int pad_w, float* iwe, void* stream) {
  EBOS_REQUIRE(window && flow && iwe && n >= 0 && H > 0 && W > 0 && pad_h >= 0 && pad_w >= 0, "ebos_window_splat: bad argument");
  return window_splat_launch(window, n, has_weight, flow, H, W, pad_h, pad_w, iwe, as_stream(stream));
}

int ebos_window_backward(const void* window, int64_t n, int has_weight, const float* flow, int H, int W, int pad_h,
                         int pad_w, const float* grad_iwe, int kind, const float* iwe, const double* acc,
                         int omit_boundary, float scale, float* dflow, void* stream) {
  EBOS_REQUIRE(window && flow && dflow && n >= 0 && H > 0 && W > 0 && pad_h >= 0 && pad_w >= 0, "ebos_window_backward: bad argument");
  return window_backward_launch(window, n, has_weight, flow, H, W, pad_h, pad_w, grad_iwe, kind, iwe, acc, omit_boundary,
                                scale, dflow, as_stream(stream));
}

}  // extern "C"
