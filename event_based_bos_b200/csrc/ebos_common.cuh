// Shared device/host helpers for libebos (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include <float.h>
#include <string>

#include "../../include/ebos.h"

namespace ebos {

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* where);

#define EBOS_REQUIRE(cond, msg)            \
  do {                                     \
    if (!(cond)) {                         \
      ebos::set_error(msg);                \
      return EBOS_ERR_BAD_ARG;             \
    }                                      \
  } while (0)

#define EBOS_LAUNCH_CHECK(where)                                  \
  do {                                                            \
    cudaError_t e__ = cudaGetLastError();                         \
    if (e__ != cudaSuccess) return ebos::cuda_fail(e__, where);   \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs (cached); grids of streaming kernels are sized from it.
int sm_count();

// ---- unfused IEEE arithmetic ----------------------------------------------------------------
// The reference computes every step as a separately rounded torch op (src/warp.py:335,
// src/event_image_converter.py:586-614).  The intrinsics below are never contracted to FMA.
template <typename T> struct Rn;
template <> struct Rn<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float flr(float a) { return floorf(a); }
  static __device__ __forceinline__ float bias() { return 1e-6f; }
  static __device__ __forceinline__ bool finite(float a) { return fabsf(a) <= FLT_MAX; }
};
template <> struct Rn<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double flr(double a) { return floor(a); }
  static __device__ __forceinline__ double bias() { return 1e-6; }
  static __device__ __forceinline__ bool finite(double a) { return fabs(a) <= DBL_MAX; }
};

// Time reference and period of one window from (tmin, tmax).  src/warp.py:230-262, 283-287.
//   t_ref per direction; dt = t - t_ref; period = max(dt) - min(dt) = (tmax - t_ref) - (tmin - t_ref)
//   (fl() is monotone, so max/min of the rounded dt are the rounded dt of tmax/tmin).
template <typename T> struct TimeRef { T t_ref; T period; };
template <typename T>
__device__ __forceinline__ TimeRef<T> make_time_ref(T tmin, T tmax, int direction, double frac) {
  TimeRef<T> r;
  if (direction == EBOS_DIR_FIRST) r.t_ref = tmin;
  else if (direction == EBOS_DIR_LAST) r.t_ref = tmax;
  else r.t_ref = Rn<T>::add(tmin, Rn<T>::mul(Rn<T>::sub(tmax, tmin), (T)frac));
  r.period = Rn<T>::sub(Rn<T>::sub(tmax, r.t_ref), Rn<T>::sub(tmin, r.t_ref));
  return r;
}
template <typename T>
__device__ __forceinline__ T event_dt(T t, const TimeRef<T>& tr, int normalize_t) {
  T d = Rn<T>::sub(t, tr.t_ref);
  return normalize_t ? Rn<T>::div(d, tr.period) : d;
}

// ---- order-preserving float <-> unsigned encodings for atomic min/max ---------------------------
template <typename T> struct Enc;
template <> struct Enc<float> {
  using U = unsigned int;
  static __device__ __forceinline__ U enc(float f) { U b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
  static __device__ __forceinline__ float dec(U u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
};
template <> struct Enc<double> {
  using U = unsigned long long;
  static __device__ __forceinline__ U enc(double f) { U b = (U)__double_as_longlong(f); return (b >> 63) ? ~b : (b | 0x8000000000000000ull); }
  static __device__ __forceinline__ double dec(U u) { return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u)); }
};

// ---- bilinear-vote taps ---------------------------------------------------------------------
// src/event_image_converter.py:586-614.  r/c are the padded integer cell; taps in reference order
// 0:(r,c) 1:(r+1,c) 2:(r,c+1) 3:(r+1,c+1).
template <typename T> struct Taps {
  int r, c;          // saturated int32 (an out-of-range cell is masked either way)
  T a, b;            // fractional parts (from the un-biased coordinate)
  T w0, w1, w2, w3;  // weights, before the optional per-event weight
};

// r = (int)floor + pad.  The float->int conversion saturates; adding a small non-negative pad to a
// saturated value wraps INT_MAX to a NEGATIVE number and leaves INT_MIN very negative, so a wrapped cell
// can never alias a valid one: every tap of such an event stays masked, as in the reference.
template <typename T>
__device__ __forceinline__ int sat_add(int v, int p) { return (int)((unsigned)v + (unsigned)p); }

template <typename T>
__device__ __forceinline__ Taps<T> make_taps(T xw, T yw, int pad_h, int pad_w, T bias = Rn<T>::bias()) {
  // bias: 1e-6 in the tensor branch (src/event_image_converter.py:586), 1e-8 in the numpy branch (:528)
  Taps<T> t;
  T fr = Rn<T>::flr(Rn<T>::add(xw, bias));
  T fc = Rn<T>::flr(Rn<T>::add(yw, bias));
  t.a = Rn<T>::sub(xw, fr);
  t.b = Rn<T>::sub(yw, fc);
  // float->int conversion saturates; NaN converts to 0 and is handled by the caller via `finite`.
  t.r = sat_add<T>((int)fr, pad_h);
  t.c = sat_add<T>((int)fc, pad_w);
  T na = Rn<T>::sub((T)1, t.a), nb = Rn<T>::sub((T)1, t.b);
  t.w0 = Rn<T>::mul(na, nb);
  t.w1 = Rn<T>::mul(t.a, nb);
  t.w2 = Rn<T>::mul(na, t.b);
  t.w3 = Rn<T>::mul(t.a, t.b);
  return t;
}

// floor() and (int)floor() at full FP32 issue rate (FRND/F2I run at quarter rate): for |v| < 2^22 adding
// 1.5*2^23 leaves the round-to-nearest integer in the low mantissa bits; one compare turns it into floor.
// Bit-identical to floorf()/saturating cast on that range (the sign of a zero result excepted, which
// cannot change any sum); anything else (huge, Inf, NaN) takes the exact slow path.
__device__ __forceinline__ float floor_to_int(float v, int& iv) {
  if (fabsf(v) < 4194304.0f) {
    const float M = 12582912.0f;
    float f = __fsub_rn(__fadd_rn(v, M), M);
    if (f > v) f = __fsub_rn(f, 1.0f);
    iv = __float_as_int(__fadd_rn(f, M)) - 0x4B400000;
    return f;
  }
  const float f = floorf(v);
  iv = (int)f;
  return f;
}
__device__ __forceinline__ double floor_to_int(double v, int& iv) {
  const double f = floor(v);
  iv = (int)f;
  return f;
}

// no-return global reduction (REDG.E.ADD.F32 / .F64)
__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// Same without the compiler-level memory clobber: for kernels that never read the accumulation target,
// so that independent loads may be scheduled across the reductions.
__device__ __forceinline__ void red_add_nc(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v));
}
__device__ __forceinline__ void red_add_nc(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v));
}
// 16-byte vector reduction (REDG.E.ADD.F32x4): one LSU lane-op for four consecutive floats.  Measured on
// B200 (profiles/microbench/r01_red_throughput.txt): a RED lane-op costs ~1.3 SM-cycles whatever its width.
// two horizontally adjacent cells in one reduction (8-byte aligned address: even column, even row stride)
__device__ __forceinline__ void red_add_v2(float* p8, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p8), "f"(a), "f"(b));
}
__device__ __forceinline__ void red_add_v4(float* p16, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p16), "f"(a), "f"(b), "f"(c), "f"(d));
}

// ---- packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2) ------------------------------------------
// One issue slot for two IEEE round-to-nearest operations; add/sub/mul are NOT fused, so the per-lane results
// are bit-identical to __fadd_rn/__fsub_rn/__fmul_rn.  Used for the (row, col) pair arithmetic of the
// streaming kernels, which are bound by issue slots.
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {  // fused: only where fusion is allowed
  float2 r;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; "
      "mov.b64 {%0,%1}, rd;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}

// ---- TMA bulk async copy (global -> shared) with mbarrier completion (sm_90+/sm_100a) ---------------
// One elected thread arms the mbarrier with the expected byte count and issues cp.async.bulk (SASS
// UBLKCP); the copy engine lands the bytes in shared memory and completes the transaction on the
// barrier; consumers spin on try_wait.parity.  Addresses and sizes must be multiples of 16 bytes.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- block reductions (warp shuffle) ----------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum `v` over the CTA; result valid in thread 0.  `smem` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.0;
  if (wid == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------
// The kernels of one evaluation form a chain (splat -> cost -> backward -> Adam) in which every link needs the
// COMPLETE output of the previous one, so they cannot be fused; but a plane kernel is ~5 us of work inside ~9 us
// of launch ramp and drain (ncu r01e: SMs active 54 % of the gradient-magnitude kernel's elapsed time).  With PDL
// the next kernel's CTAs are scheduled while the last wave of the previous kernel drains, run their prologue (index
// math, loads that do not depend on the predecessor) and block in pdl_wait() until the predecessor has completed
// and its writes are visible.  Primaries call pdl_launch_dependents() at their start; launched without the
// attribute (or after a non-kernel stream operation) both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();   // EBOS_NO_PDL=1 disables the launch attribute (A/B runs)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- prepared window ----------------------------------------------------------------------------
// Layout of the caller-owned window buffer (ebos_window_bytes): a 256-byte header followed by
// 256-byte aligned SoA arrays of the events sorted by origin pixel.
// The header is written by the device (time statistics are never read back by the host); the
// array offsets are a pure function of n, recomputed on the host by window_layout().
struct WindowHeader {
  double t_ref, period, t_min, t_max;       // stored as double; exact for fp32 windows as well
  unsigned long long enc_min, enc_max;      // order-preserving encodings used by the atomic min/max pass
  int not_packable;                         // some coordinate is not a small non-negative integer
  int packed;                               // 1: x slot holds (row << 16 | col) as uint32, y slot unused
  int n_items;                              // work items (tile, chunk) of the tile kernels
};
static_assert(sizeof(WindowHeader) <= 256, "header must fit its slot");

inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// Spatial tiling of the ORIGIN pixel grid used by the sort key and the shared-memory tile kernels.
constexpr int kTileH = 32, kTileW = 32;
constexpr int kItemEvents = 4080;           // smallest work-item size = 255 x 16 events (sizes the item array); see item_events()
inline int tiles_x(int W) { return (W + kTileW - 1) / kTileW; }
inline int tiles_y(int H) { return (H + kTileH - 1) / kTileH; }
inline int64_t max_items(int64_t n, int H, int W) { return n / kItemEvents + (int64_t)tiles_x(W) * tiles_y(H) + 1; }

// Events per work item (a busy tile is split EVENLY into items of at most this many events).  4080 = 16 events per
// thread of a 256-thread CTA; 8176 = 32 per thread: the per-CTA latency chain (item descriptor -> event stream / flow
// table -> barrier -> ... -> barrier -> flush) is paid once per item, so longer items amortise it (EBOS_ITEM_EVENTS).
int item_events();

// "Blocked-striped" storage of DENSE fp32 windows (round 2).  The tile kernels give every thread 16 CONSECUTIVE sorted
// events (so that flow gathers and cells repeat inside a thread) and read them as four 16-byte groups; with the sorted
// stream stored linearly the 32 lanes of such a load are 64 bytes apart, i.e. 16 L1 wavefronts per LDG.128 instead of 4
// -- ncu r02: a quarter of the L1 data-pipe wavefronts of the splat, the pipe that bounds it.  In a blocked window the
// stream is cut into aligned blocks of 512 events (one warp's work); inside a block, group g (0..3) of lane L sits at
// (g * 32 + L) * 4: a warp's load of group g is one contiguous 512-byte run.  The ragged tail (n % 512) stays linear.
// Only the ORDER IN MEMORY changes: logical (sorted) indices, item ranges and the exported permutation do not.
constexpr int kBlockEvents = 512;
int blocked_limit(int64_t n, int H, int W, size_t elem, bool has_weight);   // events stored blocked (multiple of 512), 0 = linear
__host__ __device__ __forceinline__ int phys_group(int b /*logical index of a group of 4, multiple of 4*/, int n_blocked) {
  if (b >= n_blocked) return b;
  const int r = b & (kBlockEvents - 1);
  return (b & ~(kBlockEvents - 1)) + ((r >> 2) & 3) * 128 + (r >> 4) * 4;
}

struct WindowLayout {
  size_t off_x, off_y, off_d, off_w, off_perm, off_tiles, off_items, total;
};
// events sorted by (tile, pixel-in-tile, time); tile_off[t] = first event of tile t;
// items = int4 (first | last << 16 tile-local pixel, begin, end, tile row << 16 | tile col)
inline WindowLayout window_layout(int64_t n, size_t elem, int H, int W) {
  WindowLayout L;
  size_t a = align256((size_t)n * elem);
  L.off_x = 256;
  L.off_y = L.off_x + a;
  L.off_d = L.off_y + a;
  L.off_w = L.off_d + a;
  L.off_perm = L.off_w + a;
  L.off_tiles = L.off_perm + align256((size_t)n * 4);
  L.off_items = L.off_tiles + align256(((size_t)tiles_x(W) * tiles_y(H) + 1) * 4);
  L.total = L.off_items + align256((size_t)max_items(n, H, W) * 16);
  return L;
}
inline size_t dtype_size(int dtype) { return dtype == EBOS_F64 ? 8 : 4; }

}  // namespace ebos
