// Operator-level kernels: drop-in replacements for the individual reference operators
//   Warp.warp_event            (src/warp.py:193-342)
//   EventImageConverter.bilinear_vote_tensor (src/event_image_converter.py:562-620)
// and their analytic backwards.  fp32 and fp64, batched.  Events are [batch, n, 4] AoS rows.
#include <cub/device/device_radix_sort.cuh>

#include "ebos_common.cuh"

namespace ebos {

// thread-local error text -------------------------------------------------------------------
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* where) {
  g_last_error = std::string(where) + ": " + cudaGetErrorString(e);
  return EBOS_ERR_CUDA;
}
bool pdl_enabled() {
  static const bool on = getenv("EBOS_NO_PDL") == nullptr;
  return on;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

// ---- AoS row access --------------------------------------------------------------------------
template <typename T> struct Row { T x, y, t, p; };
__device__ __forceinline__ Row<float> load_row(const float* ev, int64_t i) {
  float4 v = __ldg(reinterpret_cast<const float4*>(ev) + i);
  return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Row<double> load_row(const double* ev, int64_t i) {
  double2 a = __ldg(reinterpret_cast<const double2*>(ev) + 2 * i);
  double2 b = __ldg(reinterpret_cast<const double2*>(ev) + 2 * i + 1);
  return {a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void store_row(float* ev, int64_t i, float x, float y, float t, float p) {
  reinterpret_cast<float4*>(ev)[i] = make_float4(x, y, t, p);
}
__device__ __forceinline__ void store_row(double* ev, int64_t i, double x, double y, double t, double p) {
  reinterpret_cast<double2*>(ev)[2 * i] = make_double2(x, y);
  reinterpret_cast<double2*>(ev)[2 * i + 1] = make_double2(t, p);
}
__device__ __forceinline__ void load_xy(const float* ev, int64_t i, float& x, float& y) {
  float2 v = __ldg(reinterpret_cast<const float2*>(ev) + 2 * i);
  x = v.x; y = v.y;
}
__device__ __forceinline__ void load_xy(const double* ev, int64_t i, double& x, double& y) {
  double2 v = __ldg(reinterpret_cast<const double2*>(ev) + 2 * i);
  x = v.x; y = v.y;
}

template <typename T>
__global__ void k_tstats_init(typename Enc<T>::U* out, int batch) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) { out[2 * i] = ~(typename Enc<T>::U)0; out[2 * i + 1] = 0; }
}
template <typename T>
__global__ void __launch_bounds__(256) k_tstats_reduce(const T* __restrict__ ev, int64_t n, typename Enc<T>::U* out) {
  using U = typename Enc<T>::U;
  const T* e = ev + (int64_t)blockIdx.y * n * 4;
  U lo = ~(U)0, hi = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T t = __ldg(e + 4 * i + 2);
    U u = Enc<T>::enc(t);
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
    atomicMin(out + 2 * blockIdx.y, lo);
    atomicMax(out + 2 * blockIdx.y + 1, hi);
  }
}
template <typename T>
__global__ void k_tstats_decode(void* out, int batch) {
  using U = typename Enc<T>::U;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * batch) {
    U u = reinterpret_cast<U*>(out)[i];
    reinterpret_cast<T*>(out)[i] = Enc<T>::dec(u);
  }
}

template <typename T>
int time_stats_impl(const void* events, int64_t n, int batch, void* out, cudaStream_t st) {
  using U = typename Enc<T>::U;
  k_tstats_init<T><<<(batch + 127) / 128, 128, 0, st>>>(reinterpret_cast<U*>(out), batch);
  if (n > 0) {
    int bx = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
    k_tstats_reduce<T><<<dim3(bx, batch), 256, 0, st>>>(reinterpret_cast<const T*>(events), n, reinterpret_cast<U*>(out));
  }
  k_tstats_decode<T><<<(2 * batch + 127) / 128, 128, 0, st>>>(out, batch);
  EBOS_LAUNCH_CHECK("ebos_time_stats");
  return EBOS_OK;
}

// ---- warp: dense flow ----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_warp_dense(const T* __restrict__ ev, int64_t n, const T* __restrict__ flow,
                                                    int64_t flow_bs, int H, int W, const T* __restrict__ tstats,
                                                    int direction, double frac, int normalize_t, T* __restrict__ out,
                                                    int32_t* __restrict__ status) {
  int b = blockIdx.y;
  const T* e = ev + (int64_t)b * n * 4;
  T* o = out + (int64_t)b * n * 4;
  const T* f0 = flow + (int64_t)b * flow_bs;
  const T* f1 = f0 + (int64_t)H * W;
  TimeRef<T> tr = make_time_ref<T>(tstats[2 * b], tstats[2 * b + 1], direction, frac);
  const int64_t hw = (int64_t)H * W;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    Row<T> r = load_row(e, i);
    T d = event_dt<T>(r.t, tr, normalize_t);
    // trunc toward zero like Tensor.long(); the flat index is what torch.gather bound-checks.
    int64_t k = (int64_t)r.x * W + (int64_t)r.y;
    T xw = r.x, yw = r.y;
    if (k >= 0 && k < hw && Rn<T>::finite(r.x) && Rn<T>::finite(r.y)) {
      xw = Rn<T>::sub(r.x, Rn<T>::mul(d, __ldg(f0 + k)));
      yw = Rn<T>::sub(r.y, Rn<T>::mul(d, __ldg(f1 + k)));
    } else {
      bad = true;
    }
    store_row(o, i, xw, yw, d, r.p);
  }
  if (bad) atomicOr(status, EBOS_STATUS_PIXEL_OOB);
}

template <typename T>
__global__ void __launch_bounds__(256) k_warp_dense_bwd(const T* __restrict__ ev, int64_t n, int H, int W,
                                                        const T* __restrict__ tstats, int direction, double frac,
                                                        int normalize_t, const T* __restrict__ gw, T* __restrict__ dflow,
                                                        int64_t dflow_bs) {
  int b = blockIdx.y;
  const T* e = ev + (int64_t)b * n * 4;
  const T* g = gw + (int64_t)b * n * 4;
  T* d0 = dflow + (int64_t)b * dflow_bs;
  T* d1 = d0 + (int64_t)H * W;
  TimeRef<T> tr = make_time_ref<T>(tstats[2 * b], tstats[2 * b + 1], direction, frac);
  const int64_t hw = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    Row<T> r = load_row(e, i);
    T d = event_dt<T>(r.t, tr, normalize_t);
    int64_t k = (int64_t)r.x * W + (int64_t)r.y;
    if (k < 0 || k >= hw) continue;
    T gx, gy;
    load_xy(g, i, gx, gy);
    red_add(d0 + k, -(d * gx));
    red_add(d1 + k, -(d * gy));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_warp_2dof(const T* __restrict__ ev, int64_t n, const T* __restrict__ theta,
                                                   const T* __restrict__ tstats, int direction, double frac,
                                                   int normalize_t, T* __restrict__ out) {
  int b = blockIdx.y;
  const T* e = ev + (int64_t)b * n * 4;
  T* o = out + (int64_t)b * n * 4;
  TimeRef<T> tr = make_time_ref<T>(tstats[2 * b], tstats[2 * b + 1], direction, frac);
  T th0 = theta[0], th1 = theta[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    Row<T> r = load_row(e, i);
    T d = event_dt<T>(r.t, tr, normalize_t);
    store_row(o, i, Rn<T>::add(r.x, Rn<T>::mul(d, th0)), Rn<T>::add(r.y, Rn<T>::mul(d, th1)), d, r.p);
  }
}

// ---- bilinear vote ----------------------------------------------------------------------------------
// One tap of the reference's (inds*mask, vals*mask) pair: a masked tap targets pixel 0 with w*0.
template <typename T> struct TapOut { int64_t idx; bool m; T val; };
template <typename T>
__device__ __forceinline__ TapOut<T> tap_out(int r, int c, T w, T wt, bool has_w, int Hp, int Wp, bool coord_ok) {
  TapOut<T> o;
  o.m = coord_ok && r >= 0 && r < Hp && c >= 0 && c < Wp;
  T v = has_w ? Rn<T>::mul(w, wt) : w;
  o.idx = o.m ? ((int64_t)c + (int64_t)r * Wp) : 0;
  o.val = o.m ? v : Rn<T>::mul(v, (T)0);
  return o;
}

template <typename T>
__global__ void __launch_bounds__(256) k_splat_atomic(const T* __restrict__ ev, int64_t n, int Hp, int Wp, int pad_h,
                                                      int pad_w, const T* __restrict__ weight, T bias, T* __restrict__ image,
                                                      int64_t* __restrict__ dbg_idx, uint8_t* __restrict__ dbg_mask) {
  int b = blockIdx.y;
  const T* e = ev + (int64_t)b * n * 4;
  T* img = image + (int64_t)b * Hp * Wp;
  const T* wv = weight ? weight + (int64_t)b * n : nullptr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T x, y;
    load_xy(e, i, x, y);
    Taps<T> t = make_taps<T>(x, y, pad_h, pad_w, bias);
    // NaN/Inf coordinates: Tensor.long() of a non-finite floor is INT64_MIN on the CPU reference
    // -> every tap masked; the (NaN) value still lands on pixel 0 through vals*mask.
    bool ok = Rn<T>::finite(x) && Rn<T>::finite(y);
    T wt = wv ? __ldg(wv + i) : (T)1;
    TapOut<T> o[4] = {tap_out<T>(t.r, t.c, t.w0, wt, wv != nullptr, Hp, Wp, ok),
                      tap_out<T>(t.r + 1, t.c, t.w1, wt, wv != nullptr, Hp, Wp, ok && t.r != INT_MAX),
                      tap_out<T>(t.r, t.c + 1, t.w2, wt, wv != nullptr, Hp, Wp, ok && t.c != INT_MAX),
                      tap_out<T>(t.r + 1, t.c + 1, t.w3, wt, wv != nullptr, Hp, Wp, ok && t.r != INT_MAX && t.c != INT_MAX)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // a masked tap adds +-0 to pixel 0: a no-op unless the value is NaN
      if (o[k].m || o[k].val != o[k].val) red_add(img + o[k].idx, o[k].val);
      if (dbg_idx) dbg_idx[((int64_t)b * 4 + k) * n + i] = o[k].idx;
      if (dbg_mask) dbg_mask[((int64_t)b * 4 + k) * n + i] = o[k].m ? 1 : 0;
    }
  }
}

// Deterministic mode, stage 1: emit the 4n (key, val) pairs in the reference's concatenation order.
// key = pixel*4 + tap, so that a STABLE sort by key reproduces, inside every pixel, the reference's
// accumulation order (tap-major, then event order).
template <typename T>
__global__ void __launch_bounds__(256) k_splat_pairs(const T* __restrict__ e, int64_t n, int Hp, int Wp, int pad_h,
                                                     int pad_w, const T* __restrict__ wv, T bias, unsigned int* __restrict__ keys,
                                                     T* __restrict__ vals, int64_t* __restrict__ dbg_idx,
                                                     uint8_t* __restrict__ dbg_mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T x, y;
    load_xy(e, i, x, y);
    Taps<T> t = make_taps<T>(x, y, pad_h, pad_w, bias);
    bool ok = Rn<T>::finite(x) && Rn<T>::finite(y);
    T wt = wv ? __ldg(wv + i) : (T)1;
    TapOut<T> o[4] = {tap_out<T>(t.r, t.c, t.w0, wt, wv != nullptr, Hp, Wp, ok),
                      tap_out<T>(t.r + 1, t.c, t.w1, wt, wv != nullptr, Hp, Wp, ok && t.r != INT_MAX),
                      tap_out<T>(t.r, t.c + 1, t.w2, wt, wv != nullptr, Hp, Wp, ok && t.c != INT_MAX),
                      tap_out<T>(t.r + 1, t.c + 1, t.w3, wt, wv != nullptr, Hp, Wp, ok && t.r != INT_MAX && t.c != INT_MAX)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      keys[(int64_t)k * n + i] = (unsigned int)(o[k].idx * 4 + k);
      vals[(int64_t)k * n + i] = o[k].val;
      if (dbg_idx) dbg_idx[(int64_t)k * n + i] = o[k].idx;
      if (dbg_mask) dbg_mask[(int64_t)k * n + i] = o[k].m ? 1 : 0;
    }
  }
}

// Stage 3: one thread per pixel walks its (sorted) segment and accumulates sequentially in T.
template <typename T>
__global__ void __launch_bounds__(256) k_splat_segsum(const unsigned int* __restrict__ keys, const T* __restrict__ vals,
                                                      int64_t m, int64_t npix, T* __restrict__ image) {
  int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  unsigned int lo_key = (unsigned int)(pix * 4);
  // lower bound of lo_key
  int64_t lo = 0, hi = m;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < lo_key) lo = mid + 1; else hi = mid;
  }
  T acc = (T)0;
  for (int64_t j = lo; j < m && __ldg(keys + j) < lo_key + 4u; ++j) acc = Rn<T>::add(acc, __ldg(vals + j));
  image[pix] = acc;
}

template <typename T>
size_t splat_ws_bytes(int64_t n, int batch, int Hp, int Wp, int mode) {
  if (mode == 0) return 0;
  int64_t m = 4 * n;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned int*)nullptr, (unsigned int*)nullptr,
                                  (const T*)nullptr, (T*)nullptr, (int)m);
  return align256(m * 4) * 2 + align256(m * sizeof(T)) * 2 + align256(cub_bytes) + 256;
}

template <typename T>
int splat_impl(const void* events, int64_t n, int batch, int Hp, int Wp, int pad_h, int pad_w, const void* weight,
               double floor_bias, int mode, void* image, int64_t* dbg_idx, uint8_t* dbg_mask, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t npix = (int64_t)Hp * Wp;
  if (mode == 0) {
    cudaError_t e = cudaMemsetAsync(image, 0, (size_t)batch * npix * sizeof(T), st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_iwe_splat memset");
    if (n > 0) {
      int bx = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
      k_splat_atomic<T><<<dim3(bx, batch), 256, 0, st>>>(reinterpret_cast<const T*>(events), n, Hp, Wp, pad_h, pad_w,
                                                         reinterpret_cast<const T*>(weight), (T)floor_bias,
                                                         reinterpret_cast<T*>(image), dbg_idx, dbg_mask);
    }
    EBOS_LAUNCH_CHECK("ebos_iwe_splat(atomic)");
    return EBOS_OK;
  }
  // deterministic
  EBOS_REQUIRE(npix * 4 + 4 < (int64_t)1 << 32, "deterministic splat: image too large for 32-bit keys");
  EBOS_REQUIRE(4 * n < (int64_t)INT_MAX, "deterministic splat: too many events for one sort");
  size_t need = splat_ws_bytes<T>(n, batch, Hp, Wp, mode);
  if (ws_bytes < need || (need && !ws)) { set_error("ebos_iwe_splat: workspace too small"); return EBOS_ERR_WORKSPACE; }
  const int64_t m = 4 * n;
  char* p = reinterpret_cast<char*>(ws);
  p = reinterpret_cast<char*>(align256(reinterpret_cast<size_t>(p)));
  unsigned int* k_in = reinterpret_cast<unsigned int*>(p); p += align256(m * 4);
  unsigned int* k_out = reinterpret_cast<unsigned int*>(p); p += align256(m * 4);
  T* v_in = reinterpret_cast<T*>(p); p += align256(m * sizeof(T));
  T* v_out = reinterpret_cast<T*>(p); p += align256(m * sizeof(T));
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned int*)nullptr, (unsigned int*)nullptr,
                                  (const T*)nullptr, (T*)nullptr, (int)m);
  int key_bits = 1;
  while (((int64_t)1 << key_bits) < npix * 4 + 4 && key_bits < 32) ++key_bits;
  for (int b = 0; b < batch; ++b) {
    const T* e = reinterpret_cast<const T*>(events) + (int64_t)b * n * 4;
    const T* wv = weight ? reinterpret_cast<const T*>(weight) + (int64_t)b * n : nullptr;
    T* img = reinterpret_cast<T*>(image) + (int64_t)b * npix;
    if (n > 0) {
      int bx = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
      k_splat_pairs<T><<<bx, 256, 0, st>>>(e, n, Hp, Wp, pad_h, pad_w, wv, (T)floor_bias, k_in, v_in,
                                           dbg_idx ? dbg_idx + (int64_t)b * m : nullptr,
                                           dbg_mask ? dbg_mask + (int64_t)b * m : nullptr);
      cudaError_t ce = cub::DeviceRadixSort::SortPairs(p, cub_bytes, k_in, k_out, v_in, v_out, (int)m, 0, key_bits, st);
      if (ce != cudaSuccess) return cuda_fail(ce, "ebos_iwe_splat(sort)");
    }
    k_splat_segsum<T><<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(k_out, v_out, m, npix, img);
  }
  EBOS_LAUNCH_CHECK("ebos_iwe_splat(deterministic)");
  return EBOS_OK;
}

template <typename T>
__global__ void __launch_bounds__(256) k_splat_bwd(const T* __restrict__ ev, int64_t n, int Hp, int Wp, int pad_h,
                                                   int pad_w, const T* __restrict__ weight, T bias, const T* __restrict__ gimg,
                                                   T* __restrict__ gev, T* __restrict__ gweight) {
  int b = blockIdx.y;
  const T* e = ev + (int64_t)b * n * 4;
  const T* g = gimg + (int64_t)b * Hp * Wp;
  const T* wv = weight ? weight + (int64_t)b * n : nullptr;
  T* ge = gev + (int64_t)b * n * 4;
  T* gwt = gweight ? gweight + (int64_t)b * n : nullptr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T x, y;
    load_xy(e, i, x, y);
    Taps<T> t = make_taps<T>(x, y, pad_h, pad_w, bias);
    bool ok = Rn<T>::finite(x) && Rn<T>::finite(y);
    bool r0 = ok && t.r >= 0 && t.r < Hp, r1 = ok && t.r != INT_MAX && t.r + 1 >= 0 && t.r + 1 < Hp;
    bool c0 = t.c >= 0 && t.c < Wp, c1 = t.c != INT_MAX && t.c + 1 >= 0 && t.c + 1 < Wp;
    T g00 = (r0 && c0) ? __ldg(g + (int64_t)t.r * Wp + t.c) : (T)0;
    T g10 = (r1 && c0) ? __ldg(g + (int64_t)(t.r + 1) * Wp + t.c) : (T)0;
    T g01 = (r0 && c1) ? __ldg(g + (int64_t)t.r * Wp + t.c + 1) : (T)0;
    T g11 = (r1 && c1) ? __ldg(g + (int64_t)(t.r + 1) * Wp + t.c + 1) : (T)0;
    T wt = wv ? __ldg(wv + i) : (T)1;
    T dx = (((T)1 - t.b) * (g10 - g00) + t.b * (g11 - g01)) * wt;
    T dy = (((T)1 - t.a) * (g01 - g00) + t.a * (g11 - g10)) * wt;
    store_row(ge, i, dx, dy, (T)0, (T)0);
    if (gwt) gwt[i] = t.w0 * g00 + t.w1 * g10 + t.w2 * g01 + t.w3 * g11;
  }
}

inline dim3 stream_grid(int64_t n, int batch, int per_sm = 16) {
  int bx = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * per_sm));
  return dim3(bx, batch);
}

}  // namespace ebos

using namespace ebos;

extern "C" {

int ebos_version(void) { return EBOS_VERSION; }
const char* ebos_last_error(void) { return g_last_error.c_str(); }
int ebos_device_available(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n > 0 ? 1 : 0;
}

int ebos_time_stats(const void* events, int64_t n, int batch, int dtype, void* out_min_max, void* stream) {
  EBOS_REQUIRE(n >= 0 && batch >= 1 && out_min_max && (events || n == 0), "ebos_time_stats: bad argument");
  if (dtype == EBOS_F32) return time_stats_impl<float>(events, n, batch, out_min_max, as_stream(stream));
  if (dtype == EBOS_F64) return time_stats_impl<double>(events, n, batch, out_min_max, as_stream(stream));
  set_error("ebos_time_stats: unsupported dtype");
  return EBOS_ERR_UNSUPPORTED;
}

int ebos_warp_dense_flow(const void* events, int64_t n, int batch, const void* flow, int64_t flow_batch_stride, int H,
                         int W, const void* tstats, int direction, double direction_frac, int normalize_t, int dtype,
                         void* warped, int32_t* status, void* stream) {
  EBOS_REQUIRE(n >= 0 && batch >= 1 && H > 0 && W > 0 && flow && tstats && status && (n == 0 || (events && warped)),
               "ebos_warp_dense_flow: bad argument");
  EBOS_REQUIRE(direction >= EBOS_DIR_FIRST && direction <= EBOS_DIR_FRAC, "ebos_warp_dense_flow: bad direction");
  if (n == 0) return EBOS_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == EBOS_F32)
    k_warp_dense<float><<<stream_grid(n, batch), 256, 0, st>>>((const float*)events, n, (const float*)flow, flow_batch_stride, H, W,
                                                                (const float*)tstats, direction, direction_frac, normalize_t,
                                                                (float*)warped, status);
  else if (dtype == EBOS_F64)
    k_warp_dense<double><<<stream_grid(n, batch), 256, 0, st>>>((const double*)events, n, (const double*)flow, flow_batch_stride, H,
                                                                 W, (const double*)tstats, direction, direction_frac,
                                                                 normalize_t, (double*)warped, status);
  else { set_error("ebos_warp_dense_flow: unsupported dtype"); return EBOS_ERR_UNSUPPORTED; }
  EBOS_LAUNCH_CHECK("ebos_warp_dense_flow");
  return EBOS_OK;
}

int ebos_warp_dense_flow_bwd(const void* events, int64_t n, int batch, int H, int W, const void* tstats, int direction,
                             double direction_frac, int normalize_t, int dtype, const void* grad_warped, void* dflow,
                             int64_t dflow_batch_stride, void* stream) {
  EBOS_REQUIRE(n >= 0 && batch >= 1 && H > 0 && W > 0 && tstats && dflow && (n == 0 || (events && grad_warped)),
               "ebos_warp_dense_flow_bwd: bad argument");
  if (n == 0) return EBOS_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == EBOS_F32)
    k_warp_dense_bwd<float><<<stream_grid(n, batch), 256, 0, st>>>((const float*)events, n, H, W, (const float*)tstats, direction,
                                                                    direction_frac, normalize_t, (const float*)grad_warped,
                                                                    (float*)dflow, dflow_batch_stride);
  else if (dtype == EBOS_F64)
    k_warp_dense_bwd<double><<<stream_grid(n, batch), 256, 0, st>>>((const double*)events, n, H, W, (const double*)tstats,
                                                                     direction, direction_frac, normalize_t,
                                                                     (const double*)grad_warped, (double*)dflow,
                                                                     dflow_batch_stride);
  else { set_error("ebos_warp_dense_flow_bwd: unsupported dtype"); return EBOS_ERR_UNSUPPORTED; }
  EBOS_LAUNCH_CHECK("ebos_warp_dense_flow_bwd");
  return EBOS_OK;
}

int ebos_warp_2dof(const void* events, int64_t n, int batch, const void* theta, const void* tstats, int direction,
                   double direction_frac, int normalize_t, int dtype, void* warped, void* stream) {
  EBOS_REQUIRE(n >= 0 && batch >= 1 && theta && tstats && (n == 0 || (events && warped)), "ebos_warp_2dof: bad argument");
  if (n == 0) return EBOS_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == EBOS_F32)
    k_warp_2dof<float><<<stream_grid(n, batch), 256, 0, st>>>((const float*)events, n, (const float*)theta, (const float*)tstats,
                                                               direction, direction_frac, normalize_t, (float*)warped);
  else if (dtype == EBOS_F64)
    k_warp_2dof<double><<<stream_grid(n, batch), 256, 0, st>>>((const double*)events, n, (const double*)theta,
                                                                (const double*)tstats, direction, direction_frac, normalize_t,
                                                                (double*)warped);
  else { set_error("ebos_warp_2dof: unsupported dtype"); return EBOS_ERR_UNSUPPORTED; }
  EBOS_LAUNCH_CHECK("ebos_warp_2dof");
  return EBOS_OK;
}

size_t ebos_splat_workspace_bytes(int64_t n, int batch, int Hp, int Wp, int dtype, int mode) {
  if (n < 0 || mode == 0) return 0;
  return dtype == EBOS_F64 ? splat_ws_bytes<double>(n, batch, Hp, Wp, mode) : splat_ws_bytes<float>(n, batch, Hp, Wp, mode);
}

int ebos_iwe_splat(const void* events, int64_t n, int batch, int Hp, int Wp, int pad_h, int pad_w, const void* weight,
                   double floor_bias, int dtype, int mode, void* image, int64_t* dbg_idx, uint8_t* dbg_mask, void* workspace,
                   size_t workspace_bytes, void* stream) {
  EBOS_REQUIRE(n >= 0 && batch >= 1 && Hp > 0 && Wp > 0 && image && (n == 0 || events), "ebos_iwe_splat: bad argument");
  EBOS_REQUIRE(mode == 0 || mode == 1, "ebos_iwe_splat: mode must be 0 (atomic) or 1 (deterministic)");
  if (dtype == EBOS_F32)
    return splat_impl<float>(events, n, batch, Hp, Wp, pad_h, pad_w, weight, floor_bias, mode, image, dbg_idx, dbg_mask, workspace,
                             workspace_bytes, as_stream(stream));
  if (dtype == EBOS_F64)
    return splat_impl<double>(events, n, batch, Hp, Wp, pad_h, pad_w, weight, floor_bias, mode, image, dbg_idx, dbg_mask, workspace,
                              workspace_bytes, as_stream(stream));
  set_error("ebos_iwe_splat: unsupported dtype");
  return EBOS_ERR_UNSUPPORTED;
}

int ebos_iwe_splat_bwd(const void* events, int64_t n, int batch, int Hp, int Wp, int pad_h, int pad_w, const void* weight,
                       double floor_bias, int dtype, const void* grad_image, void* grad_events, void* grad_weight, void* stream) {
  EBOS_REQUIRE(n >= 0 && batch >= 1 && Hp > 0 && Wp > 0 && grad_image && (n == 0 || (events && grad_events)),
               "ebos_iwe_splat_bwd: bad argument");
  if (n == 0) return EBOS_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == EBOS_F32)
    k_splat_bwd<float><<<stream_grid(n, batch), 256, 0, st>>>((const float*)events, n, Hp, Wp, pad_h, pad_w, (const float*)weight,
                                                               (float)floor_bias, (const float*)grad_image, (float*)grad_events, (float*)grad_weight);
  else if (dtype == EBOS_F64)
    k_splat_bwd<double><<<stream_grid(n, batch), 256, 0, st>>>((const double*)events, n, Hp, Wp, pad_h, pad_w,
                                                                (const double*)weight, floor_bias, (const double*)grad_image,
                                                                (double*)grad_events, (double*)grad_weight);
  else { set_error("ebos_iwe_splat_bwd: unsupported dtype"); return EBOS_ERR_UNSUPPORTED; }
  EBOS_LAUNCH_CHECK("ebos_iwe_splat_bwd");
  return EBOS_OK;
}

}  // extern "C"
