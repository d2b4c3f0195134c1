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
// Diagnostics: EBOS_ABLATE bit mask removes one cost component from the one-shot streaming kernels so that its
// share of the run time can be measured directly (results are then WRONG; never set outside profiling):
//   1 no REDs   2 no flow gathers (constant flow)   4 no event loads (synthetic events)   8 no dL/dIWE gathers
// Compiled in only with -DEBOS_ABLATION (EBOS_BUILD_ABLATION=1 python -m event_based_bos_b200._build --force): the
// production kernels must not pay a global load + branch per flush for a diagnostic.
#ifdef EBOS_ABLATION
__device__ int g_ablate_word = 0;
#define g_ablate g_ablate_word
#else
#define g_ablate 0
#endif
// experiment knob (EBOS_SPLAT_EPT / EBOS_BWD_EPT environment variables; 0 = default)
static int env_int(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}
int blocked_limit(int64_t n, int H, int W, size_t elem, bool has_weight) {
  static const int off = env_int("EBOS_LINEAR");    // EBOS_LINEAR=1: keep dense windows linear (A/B runs)
  const bool dense = n >= (int64_t)16 * H * W;
  return (elem == 4 && !has_weight && dense && !off) ? (int)(n & ~(int64_t)(kBlockEvents - 1)) : 0;
}
int item_events();
// blocks of a full work item of a blocked-striped window (0: even cut, see k_win_items); EBOS_ITEM_CUT=0 disables
int item_blocks() {
  static const int off = getenv("EBOS_ITEM_CUT") && atoi(getenv("EBOS_ITEM_CUT")) == 0;
  const int per = (item_events() + 16) / kBlockEvents;
  return (off || per < 8) ? 0 : per;
}
// CTAs of a tile-kernel launch = an upper bound of the number of work items (one CTA per item slot)
static unsigned item_slots(int64_t n, int H, int W, int nb) {
  const int64_t tiles = (int64_t)tiles_x(W) * tiles_y(H);
  if (nb > 0 && item_blocks() > 0) return (unsigned)(n / ((int64_t)item_blocks() * kBlockEvents) + 2 * tiles + 1);
  return (unsigned)(n / (item_events() - (nb > 0 ? kBlockEvents : 0)) + tiles + 1);
}
int item_events() {
  // default 8176 = 32 events per thread: measured on B200 at 16 Mi events the direct tile splat takes 70.7 us with
  // 8176-event items against 72.1 us with 4080 (profiles/README.md, round 2)
  static const int v = env_int("EBOS_ITEM_EVENTS");
  return (v >= kItemEvents && v <= 65520) ? (v & ~15) : 8176;
}

// ---- prepare --------------------------------------------------------------------------------
__global__ void k_win_hdr_init(WindowHeader* h) {
  h->enc_min = ~0ull;
  h->enc_max = 0ull;
  h->not_packable = 0;
  h->packed = 0;
  h->n_items = 0;
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
  const int tx = (W + kTileW - 1) / kTileW, ty = (H + kTileH - 1) / kTileH;
  const unsigned int n_keys = (unsigned int)(tx * ty) * (kTileH * kTileW) + 1;
  bool bad = false, frac = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T x = __ldg(ev + 4 * i), y = __ldg(ev + 4 * i + 1);
    int64_t k = (int64_t)x * W + (int64_t)y;
    bool ok = k >= 0 && k < hw && Rn<T>::finite(x) && Rn<T>::finite(y);
    bad |= !ok;
    frac |= !(ok && x >= (T)0 && y >= (T)0 && x < (T)65536 && y < (T)65536 && x == (T)(int)x && y == (T)(int)y);
    // tile-major key: (tile, pixel inside the tile); the origin pixel is k -> (k / W, k % W)
    unsigned int key = n_keys - 1;  // invalid events: after every valid key
    if (ok) {
      const int r = (int)(k / W), c = (int)(k - (int64_t)r * W);
      const int tile = (r / kTileH) * tx + (c / kTileW);
      key = (unsigned int)tile * (kTileH * kTileW) + (unsigned int)((r % kTileH) * kTileW + (c % kTileW));
    }
    keys[i] = key;
    idx[i] = (int)i;
  }
  if (bad) atomicOr(status, EBOS_STATUS_PIXEL_OOB);
  if (frac) atomicOr(&h->not_packable, 1);
}

__global__ void k_win_layout(WindowHeader* h, int allow_packed, int32_t* __restrict__ status) {
  h->packed = (allow_packed && !h->not_packable) ? 1 : 0;
  if (h->packed) atomicOr(status, EBOS_STATUS_PACKED);
}

// tile_off[t] = first sorted position whose key belongs to tile >= t (binary search on the sorted keys)
__global__ void __launch_bounds__(256) k_win_tile_offsets(const unsigned int* __restrict__ sorted_keys, int64_t n,
                                                          int n_tiles, int* __restrict__ tile_off) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  const unsigned int target = (unsigned int)t * (kTileH * kTileW);
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(sorted_keys + mid) < target) lo = mid + 1; else hi = mid;
  }
  tile_off[t] = (int)lo;   // t == n_tiles: end of the valid events (invalid ones carry the largest key)
}

// Work items (tile, begin, end): every tile's event range cut EVENLY into ceil(len / kItemEvents) pieces (piece
// length rounded up to a multiple of 16), so that the items of a busy tile carry the same load.
// One CTA: per-tile item counts, block-wide exclusive scan (looped), fill.
// items[i] = (first | last << 16 tile-local pixel of the piece, begin, end, tile row << 16 | tile col): the tile kernels
// stage only the rows of the tile (flow table, IWE / dL/dIWE window) that the piece's origin pixels can reach.
//
// Blocked-striped windows with `blocks_per_item` > 0 (default, round 2): a tile's range is cut on ABSOLUTE block
// boundaries into pieces of `blocks_per_item` whole 512-event blocks plus one remainder piece.  A CTA of 8 warps takes 8
// blocks per pass, so a piece of 16 blocks is two passes with every warp busy, whereas the even cut leaves a typical
// benchmark tile (36 blocks -> 3 pieces of 12) with four of the eight warps idle in the second pass, waiting at the
// flush barrier (ncu r02: 16 % of the splat's stall samples): 5 instead of 6 CTA passes per tile.
__global__ void __launch_bounds__(1024) k_win_items(const int* __restrict__ tile_off, int n_tiles, int tiles_per_row,
                                                    const unsigned int* __restrict__ sorted_keys, int item_events,
                                                    int n_blocked, int blocks_per_item, int4* __restrict__ items,
                                                    WindowHeader* __restrict__ h) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += blockDim.x) {
    const int t = base + threadIdx.x;
    int cnt = 0, b = 0, e = 0;
    const bool by_blocks = n_blocked > 0 && blocks_per_item > 0;
    int fbk = 0;
    if (t < n_tiles) {
      b = tile_off[t]; e = tile_off[t + 1];
      if (by_blocks) {
        fbk = b / kBlockEvents;
        cnt = e > b ? ((e - 1) / kBlockEvents - fbk + blocks_per_item) / blocks_per_item : 0;
      } else {
        cnt = (e - b + item_events - 1) / item_events;
      }
    }
    // inclusive scan of cnt over the block
    int v = cnt;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    if (lane == 31) warp_sums[wid] = v;
    __syncthreads();
    if (wid == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int excl = carry + (wid ? warp_sums[wid - 1] : 0) + v - cnt;
    const int piece = cnt ? ((((e - b) + cnt - 1) / cnt + 15) & ~15) : 0;   // <= item_events (a multiple of 16)
    const int tcoord = ((t / tiles_per_row) << 16) | (t % tiles_per_row);   // (tile row, tile col) for the tile kernels
    for (int i = 0; i < cnt; ++i) {
      int pb = min(b + i * piece, e), pe = min(b + (i + 1) * piece, e);
      if (by_blocks) {
        pb = max(b, (fbk + i * blocks_per_item) * kBlockEvents);
        pe = min(e, (fbk + (i + 1) * blocks_per_item) * kBlockEvents);
      } else if (n_blocked > 0) {
        // even cut: the cuts INSIDE a tile sit on block boundaries (a CTA then starts on a whole block; only the two
        // blocks a tile shares with its neighbours are ragged).  item_events leaves room for the rounding (<= 511 events).
        if (i > 0) pb = min(max((pb + kBlockEvents - 1) & ~(kBlockEvents - 1), b), e);
        if (i + 1 < cnt) pe = min(max((pe + kBlockEvents - 1) & ~(kBlockEvents - 1), b), e);
      }
      int range = 0;
      if (pe > pb) {
        const unsigned int lp0 = __ldg(sorted_keys + pb) & (kTileH * kTileW - 1), lp1 = __ldg(sorted_keys + pe - 1) & (kTileH * kTileW - 1);
        range = (int)(lp0 | (lp1 << 16));
      }
      items[excl + i] = make_int4(range, pb, pe, tcoord);
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + cnt;
    __syncthreads();
  }
  if (threadIdx.x == 0) h->n_items = carry;
}

template <typename T>
__global__ void __launch_bounds__(256) k_win_gather(const T* __restrict__ ev, const T* __restrict__ weight, int64_t n,
                                                    int H, int W, const int* __restrict__ perm,
                                                    const WindowHeader* __restrict__ h, int normalize_t, int n_blocked,
                                                    T* __restrict__ sx, T* __restrict__ sy, T* __restrict__ sd,
                                                    T* __restrict__ sw) {
  TimeRef<T> tr{(T)h->t_ref, (T)h->period};
  const int64_t hw = (int64_t)H * W;
  for (int64_t jl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; jl < n; jl += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = __ldg(perm + jl);
    // storage position of sorted event jl (blocked-striped inside aligned 512-event blocks, see ebos_common.cuh)
    const int64_t j = jl < n_blocked ? (int64_t)phys_group((int)(jl & ~(int64_t)3), n_blocked) + (jl & 3) : jl;
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

  // raw fields straight from global memory (vectorised, coalesced)
  __device__ __forceinline__ void load_global(const T* __restrict__ sx, const T* __restrict__ sy,
                                              const T* __restrict__ sd, const T* __restrict__ sw, int64_t base,
                                              int64_t n) {
    load_block<T, EPT>(sd, base, n, (T)0, d);
    if (HAS_W) load_block<T, EPT>(sw, base, n, (T)0, wt);
    if constexpr (PACKED) {
      load_block<float, EPT>(reinterpret_cast<const float*>(sx), base, n, __uint_as_float(0xffffffffu), x);
    } else {
      load_block<T, EPT>(sx, base, n, (T)NAN, x);
      load_block<T, EPT>(sy, base, n, (T)0, y);
    }
  }

  // events [base, base+EPT) of the sorted stream, restricted to the half-open range [lo, hi) of one work item;
  // base is a multiple of EPT (aligned vector loads); events outside the range are marked as skipped
  __device__ __forceinline__ void load_range(const T* __restrict__ sx, const T* __restrict__ sy,
                                             const T* __restrict__ sd, const T* __restrict__ sw, int64_t base,
                                             int64_t lo, int64_t hi) {
    if (base >= lo && base + EPT <= hi) {
#pragma unroll
      for (int j = 0; j < EPT; j += 4) {
        load4(sd + base + j, d + j);
        if (HAS_W) load4(sw + base + j, wt + j);
        load4(sx + base + j, x + j);
        if (!PACKED) load4(sy + base + j, y + j);
      }
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const bool in = base + j >= lo && base + j < hi;
        d[j] = in ? sd[base + j] : (T)0;
        if (HAS_W) wt[j] = in ? sw[base + j] : (T)0;
        if (PACKED) x[j] = in ? sx[base + j] : (T)__uint_as_float(0xffffffffu);
        else { x[j] = in ? sx[base + j] : (T)NAN; y[j] = in ? sy[base + j] : (T)0; }
      }
    }
  }

  // 32-bit index version of load_range (unweighted windows), used by the direct tile splat
  // (`nb`: events stored blocked-striped, see ebos_common.cuh -- `base` is the LOGICAL index, ph the storage position)
  __device__ __forceinline__ void load_range32(const T* __restrict__ sx, const T* __restrict__ sy, const T* __restrict__ sd,
                                               int base, int lo, int hi, int nb = 0) {
    static_assert(EPT == 4, "load_range32 loads one group of four events");
    const int ph = phys_group(base, nb);
    if (base >= lo && base + EPT <= hi) {
      load4(sd + ph, d);
      load4(sx + ph, x);
      if (!PACKED) load4(sy + ph, y);
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const bool in = base + j >= lo && base + j < hi;
        d[j] = in ? sd[ph + j] : (T)0;
        if (PACKED) x[j] = in ? sx[ph + j] : (T)__uint_as_float(0xffffffffu);
        else { x[j] = in ? sx[ph + j] : (T)NAN; y[j] = in ? sy[ph + j] : (T)0; }
      }
    }
  }
  // one group of four events of the whole stream [0, n) (flat kernels), blocked-aware
  __device__ __forceinline__ void load_group(const T* __restrict__ sx, const T* __restrict__ sy, const T* __restrict__ sd,
                                             const T* __restrict__ sw, int64_t base, int64_t n, int nb) {
    static_assert(EPT == 4, "load_group loads one group of four events");
    if (nb == 0) { load_global(sx, sy, sd, sw, base, n); return; }
    const int ph = phys_group((int)base, nb);
    if (base + EPT <= n) {
      load4(sd + ph, d);
      if (HAS_W) load4(sw + ph, wt);
      load4(sx + ph, x);
      if (!PACKED) load4(sy + ph, y);
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const bool in = base + j < n;
        d[j] = in ? sd[ph + j] : (T)0;
        if (HAS_W) wt[j] = in ? sw[ph + j] : (T)0;
        if (PACKED) x[j] = in ? sx[ph + j] : (T)__uint_as_float(0xffffffffu);
        else { x[j] = in ? sx[ph + j] : (T)NAN; y[j] = in ? sy[ph + j] : (T)0; }
      }
    }
  }

  // The events are sorted by origin pixel: a thread's consecutive events mostly share it, so a gather is only issued
  // when the pixel changes (ablation r01c: the redundant gathers cost 15-17 us per pass at 16 Mi events).
  // EBOS_GATHER_SELECT (build flag, A/B): the conditional gathers go to registers of their own, preset with a constant,
  // and the "same pixel -> previous value" choice is made by selects afterwards.  Written as `f[j] = changed ? load :
  // f[j-1]` the compiler emits MOV f[j], f[j-1]; @P LDG f[j] -- and the MOV waits for the FIRST gather to land before the
  // second can even issue: two to four serial L2 round trips per group (ncu r02s: 5-11 % of the stall samples of the two
  // event kernels sit on those MOVs).
  __device__ __forceinline__ void gather_flow(const T* __restrict__ flow, int hw) {
#if defined(EBOS_GATHER_SELECT)
    T a0[EPT], a1[EPT];
    bool ch[EPT];
    a0[0] = __ldg(flow + k[0]);
    a1[0] = __ldg(flow + hw + k[0]);
#pragma unroll
    for (int j = 1; j < EPT; ++j) {
      ch[j] = k[j] != k[j - 1];
      a0[j] = (T)0;
      a1[j] = (T)0;
      if (ch[j]) {
        a0[j] = __ldg(flow + k[j]);
        a1[j] = __ldg(flow + hw + k[j]);
      }
    }
    f0[0] = a0[0];
    f1[0] = a1[0];
#pragma unroll
    for (int j = 1; j < EPT; ++j) {
      f0[j] = ch[j] ? a0[j] : f0[j - 1];
      f1[j] = ch[j] ? a1[j] : f1[j - 1];
    }
#else
    f0[0] = __ldg(flow + k[0]);
    f1[0] = __ldg(flow + hw + k[0]);
#pragma unroll
    for (int j = 1; j < EPT; ++j) {
      if (k[j] != k[j - 1]) {
        f0[j] = __ldg(flow + k[j]);
        f1[j] = __ldg(flow + hw + k[j]);
      } else {
        f0[j] = f0[j - 1];
        f1[j] = f1[j - 1];
      }
    }
#endif
  }

  // finish() for a group that lies entirely inside its item: no end-of-stream markers to test
  __device__ __forceinline__ void finish_full(const T* __restrict__ flow, int W, int hw) {
    if constexpr (PACKED) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const unsigned rc = __float_as_uint(x[j]);
        const int r = rc >> 16, c = rc & 0xffff;
        k[j] = r * W + c;
        x[j] = (T)r;
        y[j] = (T)c;
      }
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int kk = (int)x[j] * W + (int)y[j];   // NaN converts to 0: parked events gather flow[0]
        k[j] = (unsigned)kk < (unsigned)hw ? kk : 0;
      }
    }
    gather_flow(flow, hw);
  }

  // raw fields from a shared-memory stage filled by the TMA bulk copies: arrays of `ch` elements in the
  // order (x|rc, [y], d, [w]); `valid` = number of real events of this thread's block (tail of the stream)
  __device__ __forceinline__ void load_stage(const T* __restrict__ stage, int ch, int off, int valid) {
    const T* px = stage + off;
    const T* py = stage + ch + off;
    const T* pd = stage + (PACKED ? 1 : 2) * ch + off;
    const T* pw = stage + (PACKED ? 2 : 3) * ch + off;
#pragma unroll
    for (int j = 0; j < EPT; j += 4) {
      load4s(px + j, x + j);
      if (!PACKED) load4s(py + j, y + j);
      load4s(pd + j, d + j);
      if (HAS_W) load4s(pw + j, wt + j);
    }
    if (valid < EPT) {
#pragma unroll
      for (int j = 0; j < EPT; ++j)
        if (j >= valid) { x[j] = PACKED ? (T)__uint_as_float(0xffffffffu) : (T)NAN; y[j] = 0; d[j] = 0; }
    }
  }
  static __device__ __forceinline__ void load4s(const float* p, float* o) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void load4s(const double* p, double* o) {
    const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }

  // origin pixel, float coordinates and the flow gathers (all issued before any is consumed)
  __device__ __forceinline__ void finish(const T* __restrict__ flow, int W, int hw) {
    if constexpr (PACKED) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const unsigned rc = __float_as_uint(x[j]);
        const bool skip = rc == 0xffffffffu;  // only past the end of the stream
        const int r = rc >> 16, c = rc & 0xffff;
        k[j] = skip ? 0 : r * W + c;
        x[j] = skip ? (T)NAN : (T)r;
        y[j] = (T)c;
      }
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const int kk = (int)x[j] * W + (int)y[j];   // NaN converts to 0: parked events gather flow[0]
        k[j] = (unsigned)kk < (unsigned)hw ? kk : 0;
      }
    }
    gather_flow(flow, hw);
  }

  __device__ __forceinline__ void load(const T* __restrict__ sx, const T* __restrict__ sy, const T* __restrict__ sd,
                                       const T* __restrict__ sw, int64_t base, int64_t n, const T* __restrict__ flow,
                                       int W, int hw) {
    const int abl = g_ablate;
    if (abl & 4) {  // diagnostics: synthetic events instead of loads
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const unsigned pix = (unsigned)((base + j) / 17) % (unsigned)hw;
        if (PACKED) x[j] = (T)__uint_as_float(((pix / W) << 16) | (pix % W)); else { x[j] = (T)(pix / W); y[j] = (T)(pix % W); }
        d[j] = (T)((base + j) % 17) * (T)0.0588f;
        if (HAS_W) wt[j] = 1;
      }
    } else {
      load_global(sx, sy, sd, sw, base, n);
    }
    if (abl & 2) {  // diagnostics: no flow gathers
      finish_nogather(W, hw);
    } else {
      finish(flow, W, hw);
    }
  }
  __device__ __forceinline__ void finish_nogather(int W, int hw) {
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      if constexpr (PACKED) {
        const unsigned rc = __float_as_uint(x[j]);
        const int r = rc >> 16, c = rc & 0xffff;
        k[j] = r * W + c; x[j] = (T)r; y[j] = (T)c;
      } else {
        const int kk = (int)x[j] * W + (int)y[j];
        k[j] = (unsigned)kk < (unsigned)hw ? kk : 0;
      }
      f0[j] = (T)1.37 + (T)(k[j] & 3) * (T)0.4; f1[j] = (T)-2.11 + (T)(k[j] & 7) * (T)0.3;
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
// floor()/float->int run on the conversion (XU) pipe: ONE issue slot each.  The kernels are bound by issue
// slots at ~0.5 IPC (ncu r01b: 45.4 M warp instructions / (592 SMSP x 0.51) = the measured 80 us), so the
// FP32-pipe emulation of floor (4 slots) that an earlier revision used was a net loss.
template <typename T> struct FastFloor;
template <> struct FastFloor<float> {
  static __device__ __forceinline__ float flr(float v) { return floorf(v); }
  static __device__ __forceinline__ int to_int(float f) { return (int)f; }  // saturating; NaN -> 0
};
template <> struct FastFloor<double> {
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
  if (g_ablate & 1) { if (a0 + a1 + a2 + a3 == (T)-12345) iwe[0] = a0; return; }  // diagnostics: no REDs
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
#if !defined(EBOS_NO_RED_V2)
    // the two column-adjacent taps of a row as ONE 8-byte vector reduction when the cell's column is even (no zero lanes,
    // unlike the red.v4 form above): 2 instead of 4 reductions for half of the cells.  Measured on B200 (r02d2, build flag
    // A/B): sparse splat 12.5 -> 11.9 us at 500 k events, 1 Mi-event evaluation 46.1 -> 44.9 us, solve +1 %
    if constexpr (sizeof(T) == 4) {
      if (!((c | Wp) & 1) && (reinterpret_cast<size_t>(iwe) & 7) == 0) {
        red_add_v2(reinterpret_cast<float*>(p), a0, a2);
        red_add_v2(reinterpret_cast<float*>(p + Wp), a1, a3);
        return;
      }
    }
#endif
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

// fp32 version on packed f32x2 arithmetic: identical per-lane rounding, about half the FP issue slots.
template <bool HAS_W, int EPT, bool VEC, bool PACKED>
__device__ __forceinline__ void splat_block_f32(const EventBlock<float, EPT, HAS_W, PACKED>& e, float* __restrict__ iwe,
                                                int Hp, int Wp, int pad_h, int pad_w) {
  const int Hm1 = Hp - 1, Wm1 = Wp - 1;
  const float2 bias2 = make_float2(1e-6f, 1e-6f), one2 = make_float2(1.f, 1.f), zero2 = make_float2(0.f, 0.f);
  float cfr = NAN, cfc = 0.f;           // current run: floor values of the cell (NaN = none)
  float2 a01 = zero2, a23 = zero2;      // tap sums (w0, w1), (w2, w3)
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    // x' = x - (dt * f): the two roundings must stay separate (bit-exact cells).  ptxas 12.9 contracts
    // mul.rn.f32x2 + sub.rn.f32x2 into FFMA2 (even with --fmad=false), so this step uses the scalar forms.
    const float2 w = make_float2(__fsub_rn(e.x[j], __fmul_rn(e.d[j], e.f0[j])), __fsub_rn(e.y[j], __fmul_rn(e.d[j], e.f1[j])));
    const float2 wb = add2(w, bias2);
    const float fr = floorf(wb.x), fc = floorf(wb.y);
    const float2 ab = sub2(w, make_float2(fr, fc));
    const float2 nab = sub2(one2, ab);
    const float2 lhs = make_float2(nab.x, ab.x);                        // (1-a, a)
    float2 w01 = mul2(lhs, make_float2(nab.y, nab.y));                  // (w0, w1) = ((1-a)(1-b), a(1-b))
    float2 w23 = mul2(lhs, make_float2(ab.y, ab.y));                    // (w2, w3) = ((1-a)b, ab)
    if (w01.x != w01.x) {
      // NaN weight <=> non-finite warped coordinate or an event marked to be skipped (x = NaN)
      splat_event_exact<float>(iwe, Hp, Wp, pad_h, pad_w, e.x[j], w.x, w.y, HAS_W ? e.wt[j] : 1.f);
      continue;
    }
    if (HAS_W) {
      const float2 ww = make_float2(e.wt[j], e.wt[j]);
      w01 = mul2(w01, ww);
      w23 = mul2(w23, ww);
    }
    if (!((fr == cfr) & (fc == cfc))) {
      if (cfr == cfr)
        flush_cell<float, VEC>(iwe, Hp, Wp, Hm1, Wm1, (int)cfr + pad_h, (int)cfc + pad_w, a01.x, a01.y, a23.x, a23.y);
      cfr = fr; cfc = fc;
      a01 = zero2; a23 = zero2;
    }
    a01 = add2(a01, w01);
    a23 = add2(a23, w23);
  }
  if (cfr == cfr) flush_cell<float, VEC>(iwe, Hp, Wp, Hm1, Wm1, (int)cfr + pad_h, (int)cfc + pad_w, a01.x, a01.y, a23.x, a23.y);
}

// EPT consecutive events of one thread: warp, vote, combine runs of equal cells, flush.
template <typename T, bool HAS_W, int EPT, bool VEC, bool PACKED>
__device__ __forceinline__ void splat_block(const EventBlock<T, EPT, HAS_W, PACKED>& e, T* __restrict__ iwe, int Hp,
                                            int Wp, int pad_h, int pad_w) {
  if constexpr (sizeof(T) == 4) {
    splat_block_f32<HAS_W, EPT, VEC, PACKED>(e, iwe, Hp, Wp, pad_h, pad_w);
    return;
  }
  const int Hm1 = Hp - 1, Wm1 = Wp - 1;
  // current run: floor values of the cell (NaN = none) and the four tap sums
  T cfr = (T)NAN, cfc = (T)0;
  T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    const T xw = Rn<T>::sub(e.x[j], Rn<T>::mul(e.d[j], e.f0[j]));
    const T yw = Rn<T>::sub(e.y[j], Rn<T>::mul(e.d[j], e.f1[j]));
    const T fr = FastFloor<T>::flr(Rn<T>::add(xw, Rn<T>::bias())), fc = FastFloor<T>::flr(Rn<T>::add(yw, Rn<T>::bias()));
    const T a = Rn<T>::sub(xw, fr), b = Rn<T>::sub(yw, fc);
    const T na = Rn<T>::sub((T)1, a), nb = Rn<T>::sub((T)1, b);
    T w0 = Rn<T>::mul(na, nb), w1 = Rn<T>::mul(a, nb), w2 = Rn<T>::mul(na, b), w3 = Rn<T>::mul(a, b);
    if (w0 != w0) {
      // NaN weight <=> non-finite warped coordinate (NaN, or Inf whose fraction is Inf - Inf), or an event
      // marked to be skipped (x = NaN): the exact path reproduces the reference (NaN lands on pixel 0).
      splat_event_exact<T>(iwe, Hp, Wp, pad_h, pad_w, e.x[j], xw, yw, HAS_W ? e.wt[j] : (T)1);
      continue;
    }
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

// one-shot variant (fp64 windows, small windows): every thread loads its events straight from global memory
template <typename T, bool HAS_W, int EPT, bool VEC, bool PACKED>
__global__ void __launch_bounds__(256, (sizeof(T) == 4 && EPT <= 8) ? 4 : 1)
k_win_splat(const T* __restrict__ sx, const T* __restrict__ sy, const T* __restrict__ sd, const T* __restrict__ sw,
            int64_t n, const T* __restrict__ flow, int H, int W, int pad_h, int pad_w, T* __restrict__ iwe) {
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * EPT;
  if (base >= n) return;
  EventBlock<T, EPT, HAS_W, PACKED> e;
  e.load(sx, sy, sd, sw, base, n, flow, W, H * W);
  splat_block<T, HAS_W, EPT, VEC, PACKED>(e, iwe, H + 2 * pad_h, W + 2 * pad_w, pad_h, pad_w);
}

// ---- grouped one-shot kernels (fp32 default) ------------------------------------------------------------
// Lean instruction stream AND high occupancy: a thread walks over NG groups of 4 consecutive events; only one
// group lives in registers at a time (<= 40 registers -> 6 CTAs of 256 threads per SM), while the run state
// (current cell + tap sums, or current origin pixel + gradient sums) is carried across the groups, so the
// register-level combining sees 4*NG consecutive events.
struct SplatRun { float cfr, cfc; float2 a01, a23; };   // a01 = taps (r,c),(r+1,c);  a23 = taps (r,c+1),(r+1,c+1)

template <bool HAS_W, bool PACKED>
__device__ __forceinline__ void splat_group4(const EventBlock<float, 4, HAS_W, PACKED>& e, SplatRun& run,
                                             float* __restrict__ iwe, int Hp, int Wp, int pad_h, int pad_w) {
  const float2 bias2 = make_float2(1e-6f, 1e-6f), one2 = make_float2(1.f, 1.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 w = make_float2(__fsub_rn(e.x[j], __fmul_rn(e.d[j], e.f0[j])), __fsub_rn(e.y[j], __fmul_rn(e.d[j], e.f1[j])));
    const float2 wb = add2(w, bias2);
    const float fr = floorf(wb.x), fc = floorf(wb.y);
    const float2 ab = sub2(w, make_float2(fr, fc));
    const float2 nab = sub2(one2, ab);
    const float2 lhs = make_float2(nab.x, ab.x);
    float2 w01 = mul2(lhs, make_float2(nab.y, nab.y));
    float2 w23 = mul2(lhs, make_float2(ab.y, ab.y));
    if (w01.x != w01.x) {
      splat_event_exact<float>(iwe, Hp, Wp, pad_h, pad_w, e.x[j], w.x, w.y, HAS_W ? e.wt[j] : 1.f);
      continue;
    }
    if (HAS_W) {
      const float2 ww = make_float2(e.wt[j], e.wt[j]);
      w01 = mul2(w01, ww);
      w23 = mul2(w23, ww);
    }
    if (!((fr == run.cfr) & (fc == run.cfc))) {
      // (tried: carrying the two taps shared with an edge-adjacent next cell instead of flushing all four --
      //  38 % fewer RED lane-ops, no measurable gain: 87.9 vs 84.2 us, so the simple flush stays)
      if (run.cfr == run.cfr)
        flush_cell<float, false>(iwe, Hp, Wp, Hp - 1, Wp - 1, (int)run.cfr + pad_h, (int)run.cfc + pad_w, run.a01.x, run.a01.y,
                                 run.a23.x, run.a23.y);
      run.cfr = fr; run.cfc = fc;
      run.a01 = make_float2(0.f, 0.f); run.a23 = make_float2(0.f, 0.f);
    }
    run.a01 = add2(run.a01, w01);
    run.a23 = add2(run.a23, w23);
  }
}

template <bool HAS_W, bool PACKED, int NG>
__global__ void __launch_bounds__(256, 6)
k_win_splat_g(const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ sd,
              const float* __restrict__ sw, int64_t n, const float* __restrict__ flow, int H, int W, int pad_h,
              int pad_w, float* __restrict__ iwe) {
  pdl_launch_dependents();
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * (4 * NG);
  if (base >= n) return;
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w, hw = H * W;
  SplatRun run{NAN, 0.f, make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll 1
  for (int g = 0; g < NG; ++g) {
    const int64_t b = base + 4 * g;
    if (b >= n) break;
    EventBlock<float, 4, HAS_W, PACKED> e;
    e.load(sx, sy, sd, sw, b, n, flow, W, hw);
    splat_group4<HAS_W, PACKED>(e, run, iwe, Hp, Wp, pad_h, pad_w);
  }
  if (run.cfr == run.cfr)
    flush_cell<float, false>(iwe, Hp, Wp, Hp - 1, Wp - 1, (int)run.cfr + pad_h, (int)run.cfc + pad_w, run.a01.x, run.a01.y,
                             run.a23.x, run.a23.y);
}

// ---- persistent, TMA-staged streaming (fp32) ---------------------------------------------------------
// The one-shot kernels were latency-bound: a thread's only loads from HBM sit at its very beginning and
// nothing else in the thread can run until they land (ncu r01b: 37-46 % of stall samples on the long
// scoreboard at the first use of the event fields, 44 % occupancy).  Here a persistent CTA walks over
// chunks of 256 x EPT consecutive events; the chunk's SoA slices are brought into shared memory by the
// copy engine (cp.async.bulk + mbarrier transaction count), double buffered, so the HBM latency of chunk
// i+1 and i+2 overlaps the arithmetic of chunk i and the threads only ever wait on shared memory.
constexpr int kPipeEpt = 8;
constexpr int kPipeChunk = 256 * kPipeEpt;
constexpr int kPipeStages = 2;

template <bool HAS_W, bool PACKED> struct PipeCfg {
  static constexpr int narr = (PACKED ? 2 : 3) + (HAS_W ? 1 : 0);
  static constexpr int stage_floats = narr * kPipeChunk;
  static constexpr size_t smem_bytes = (size_t)kPipeStages * stage_floats * sizeof(float);
};

template <bool HAS_W, bool PACKED>
__device__ __forceinline__ void pipe_issue(float* __restrict__ stage, unsigned long long* bar, const float* __restrict__ sx,
                                           const float* __restrict__ sy, const float* __restrict__ sd,
                                           const float* __restrict__ sw, int64_t chunk, int64_t n) {
  const int64_t base = chunk * kPipeChunk;
  const int64_t cnt = min((int64_t)kPipeChunk, n - base);
  const unsigned bytes = (unsigned)((cnt * 4 + 15) & ~(int64_t)15);  // the window arrays are padded to 256 B
  mbar_expect_tx(bar, bytes * PipeCfg<HAS_W, PACKED>::narr);
  int a = 0;
  bulk_g2s(stage + (a++) * kPipeChunk, sx + base, bytes, bar);
  if (!PACKED) bulk_g2s(stage + (a++) * kPipeChunk, sy + base, bytes, bar);
  bulk_g2s(stage + (a++) * kPipeChunk, sd + base, bytes, bar);
  if (HAS_W) bulk_g2s(stage + (a++) * kPipeChunk, sw + base, bytes, bar);
}

template <bool HAS_W, bool PACKED, typename BODY>
__device__ __forceinline__ void pipe_loop(const float* __restrict__ sx, const float* __restrict__ sy,
                                          const float* __restrict__ sd, const float* __restrict__ sw, int64_t n,
                                          const float* __restrict__ flow, int W, int hw, BODY body) {
  extern __shared__ __align__(128) unsigned char pipe_smem[];
  __shared__ __align__(8) unsigned long long full[kPipeStages];
  float* stages = reinterpret_cast<float*>(pipe_smem);
  constexpr int SF = PipeCfg<HAS_W, PACKED>::stage_floats;
  const int64_t n_chunks = (n + kPipeChunk - 1) / kPipeChunk;
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kPipeStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kPipeStages; ++s) {
      const int64_t c = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (c < n_chunks) pipe_issue<HAS_W, PACKED>(stages + s * SF, &full[s], sx, sy, sd, sw, c, n);
    }
  }
  unsigned it = 0;
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
    const int s = it % kPipeStages;
    mbar_wait(&full[s], (it / kPipeStages) & 1);
    EventBlock<float, kPipeEpt, HAS_W, PACKED> e;
    const int64_t first = chunk * kPipeChunk + (int64_t)tid * kPipeEpt;
    const int valid = (int)max((int64_t)0, min((int64_t)kPipeEpt, n - first));
    e.load_stage(stages + s * SF, kPipeChunk, tid * kPipeEpt, valid);
    __syncthreads();  // every thread has copied its events out: the stage may be refilled
    if (tid == 0) {
      const int64_t nxt = chunk + (int64_t)kPipeStages * gridDim.x;
      if (nxt < n_chunks) pipe_issue<HAS_W, PACKED>(stages + s * SF, &full[s], sx, sy, sd, sw, nxt, n);
    }
    if (valid > 0) {
      e.finish(flow, W, hw);
      body(e);
    }
  }
}

template <bool HAS_W, bool VEC, bool PACKED>
__global__ void __launch_bounds__(256, 3)
k_win_splat_pipe(const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ sd,
                 const float* __restrict__ sw, int64_t n, const float* __restrict__ flow, int H, int W, int pad_h,
                 int pad_w, float* __restrict__ iwe) {
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  pipe_loop<HAS_W, PACKED>(sx, sy, sd, sw, n, flow, W, H * W, [&](const EventBlock<float, kPipeEpt, HAS_W, PACKED>& e) {
    splat_block<float, HAS_W, kPipeEpt, VEC, PACKED>(e, iwe, Hp, Wp, pad_h, pad_w);
  });
}

// ---- shared-memory tile kernels (fp32) ----------------------------------------------------------------
// The one-shot kernels sit on two LSU walls measured on B200 (profiles/microbench/r01_red_throughput.txt and the
// r01c ablation): global REDs cost ~1.15 SM-cycles per lane (17 M lane-REDs = 70 us of the 80 us splat) and the
// four dL/dIWE gathers of the backward 0.24 cycles per lane (67 M lane-loads = 57 us of 90 us).  Shared memory
// does the same operations 3-4x faster (float atomicAdd 0.42, LDS 0.14 cycles per lane).  Events are therefore
// sorted by 32x32 TILE of their origin pixel; a CTA takes one work item (tile, <= 8192 events), accumulates /
// gathers in a shared-memory window of the IWE that covers the tile plus a halo of kHalo pixels, and only taps
// outside the window (|flow * dt| > kHalo) fall back to global memory, so correctness never depends on the halo.
constexpr int kHalo = 4;
constexpr int kSH = kTileH + 2 * kHalo + 1;   // window rows  (taps reach one past the last cell)
constexpr int kSW = kTileW + 2 * kHalo + 1;   // window cols; 41 is odd: consecutive rows start in different banks

// ---- shared-memory tile splat with FIXED-POINT accumulation (fp32 default for dense windows) -----------
// Shared-memory float atomics are a CAS loop on sm_100a, integer ATOMS.ADD is native and 8x cheaper per lane than
// a global RED (0.14 vs 1.15 SM-cycles, profiles/microbench/r01_red_throughput.txt).  One CTA takes one work item
// (<= kItemEvents events of one 32x32 tile), every thread walks over 16 consecutive events, combines runs of
// equal cells in registers exactly like k_win_splat_g and adds each run's four tap sums to an int32 window as
// round(sum * 2^S).  S is chosen per item so that the window can never overflow:
//   |tap weight| <= 1 + 2e-6, so every partial sum of an item of cnt events is < 2^(ilog2(cnt)+1), and with
//   S = 30 - ilog2(cnt) (<= 24) every int32 cell stays below 2^31 in magnitude.
// Quantisation: one rounding of <= 2^-(S+1) per flushed tap sum (S >= 19 for 4080 events, i.e. <= 9.5e-7: the
// size of ONE fp32 rounding of a cell value in [8, 16)); integer addition is exact and order independent inside an
// item.  The window is then added to the global IWE with coalesced fp32 REDs (windows of neighbouring tiles / items
// overlap, so the last few additions per cell are fp32 and unordered like in the other kernels).  Taps outside the window (|flow * dt| > kHalo) and
// non-finite events take the global fp32 path, so correctness never depends on the halo.  Items with fewer than
// kWinMinEvents events skip the window (zero + flush of 1681 cells would cost more than their REDs).
constexpr int kWinMinEvents = 1024;

// Direct variant: NO run tracking.  The divergent "cell changed -> flush" branch of the run-combining kernels is
// executed by a warp at almost every event step (one lane in four flushes), so its ~25 instructions are paid
// per event anyway; with native integer shared-memory atomics at 4.5 SM-cycles per warp instruction it is cheaper
// to add every event's four taps straight to the window: ~45 instead of ~95 issue slots per event.
// float -> fixed point on the FP32 pipe: fma(w, 2^S, 1.5 * 2^23) leaves round(w * 2^S) in the low mantissa bits
// (|w * 2^S| < 2^22, i.e. S <= 21); no quarter-rate F2I.
__device__ __noinline__ void splat_taps_global(float* __restrict__ iwe, int Hp, int Wp, int r, int c, float w0, float w1,
                                               float w2, float w3) {
  flush_cell<float, false>(iwe, Hp, Wp, Hp - 1, Wp - 1, r, c, w0, w1, w2, w3);
}

template <bool PACKED, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_tile_splat_d(const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ sd,
               const int4* __restrict__ items, const WindowHeader* __restrict__ hdr, const float* __restrict__ flow,
               int H, int W, int pad_h, int pad_w, float* __restrict__ iwe, int nb) {
  // One CTA per item slot (unused slots exit at once): the hardware block scheduler balances the ragged items.  A
  // persistent variant (static round-robin over items, window re-zeroed by the flush) was measured slower on B200
  // (78-114 vs 77 us at 16 Mi events: more live state -> spills, and the per-item barriers idle whole CTAs).
  __shared__ int win[kSH * kSW];
  pdl_launch_dependents();   // the cost kernel may be scheduled while this grid drains (it waits before reading the IWE)
  const int n_items = hdr->n_items;
  if ((int)blockIdx.x >= n_items) return;
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w, hw = H * W;
  const float M = 12582912.0f;   // 1.5 * 2^23
  const float2 bias2 = make_float2(1e-6f, 1e-6f), one2 = make_float2(1.f, 1.f);
  for (int i = threadIdx.x; i < kSH * kSW; i += blockDim.x) win[i] = 0;
  __syncthreads();
  const int4 it = __ldg(items + blockIdx.x);
  {
    const int cnt = it.z - it.y;
    if (cnt > 0) {
      const int r_org = (it.w >> 16) * kTileH - kHalo + pad_h, c_org = (it.w & 0xffff) * kTileW - kHalo + pad_w;
      const bool use_win = cnt >= kWinMinEvents;
      const int S = min(21, 30 - (31 - __clz(cnt)));
      const float qs = __int_as_float((127 + S) << 23);
      // 32-bit indices (n < 2^31 is checked by ebos_window_prepare).  A group of 4 events takes ONE branch: when all
      // four are regular (inside the item, finite, all taps inside the window) the 16 atomics are issued back to
      // back; anything else goes through the per-event path.  The per-event NaN / window branches of the first
      // version cost ~8 issue slots per event (ncu r01f: 79 instructions per event, issue-active 75 %).
      const int lo = it.y, hi = it.z;
      // blocked-striped windows (nb > 0): a thread's 16 events must be lane L's slot of an aligned 512-event block, so
      // the CTA starts on the block that holds `lo` (groups in front of `lo` belong to the previous item: skipped)
      const int start = nb > 0 ? (lo & ~(kBlockEvents - 1)) : (lo & ~3);
      for (int base = start + (int)threadIdx.x * 16; base < hi; base += (int)blockDim.x * 16) {
        if (base + 16 <= lo) continue;
        // software pipelining: the raw fields of group g+1 are requested before group g is processed (the first use
        // of a freshly loaded group was the top stall site, ncu r01d)
        EventBlock<float, 4, false, PACKED> e, nxt;
        e.load_range32(sx, sy, sd, base, lo, hi, nb);
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          const int b = base + 4 * g;
          if (b >= hi) break;
          if (g < 3 && b + 4 < hi) nxt.load_range32(sx, sy, sd, b + 4, lo, hi, nb);
          if (b + 4 <= lo) {      // whole group in front of the item
#pragma unroll
            for (int j = 0; j < 4; ++j) { e.x[j] = nxt.x[j]; e.y[j] = nxt.y[j]; e.d[j] = nxt.d[j]; }
            continue;
          }
          const bool full = b >= lo && b + 4 <= hi;
          if (full) e.finish_full(flow, W, hw); else e.finish(flow, W, hw);
          float2 w[4], w01[4], w23[4];
          int off[4], rr[4], cc[4];
          bool ok = full && use_win;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // x' = x - (dt * f) with two roundings (see splat_block_f32)
            w[j] = make_float2(__fsub_rn(e.x[j], __fmul_rn(e.d[j], e.f0[j])), __fsub_rn(e.y[j], __fmul_rn(e.d[j], e.f1[j])));
            const float2 wb = add2(w[j], bias2);
            const float fr = floorf(wb.x), fc = floorf(wb.y);
            const float2 ab = sub2(w[j], make_float2(fr, fc));
            const float2 nab = sub2(one2, ab);
            const float2 lhs = make_float2(nab.x, ab.x);
            w01[j] = mul2(lhs, make_float2(nab.y, nab.y));
            w23[j] = mul2(lhs, make_float2(ab.y, ab.y));
            rr[j] = (int)fr + pad_h; cc[j] = (int)fc + pad_w;
            const int lr = rr[j] - r_org, lc = cc[j] - c_org;
            off[j] = lr * kSW + lc;
            // (a NaN weight fails the first test; a huge coordinate saturates the conversion and fails the range tests)
            ok = ok && (w01[j].x == w01[j].x) && (unsigned)lr < (unsigned)(kSH - 1) && (unsigned)lc < (unsigned)(kSW - 1);
          }
          if (ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              int* p = win + off[j];
              atomicAdd(p, __float_as_int(fmaf(w01[j].x, qs, M)) - 0x4B400000);            // (r  , c  )
              atomicAdd(p + kSW, __float_as_int(fmaf(w01[j].y, qs, M)) - 0x4B400000);      // (r+1, c  )
              atomicAdd(p + 1, __float_as_int(fmaf(w23[j].x, qs, M)) - 0x4B400000);        // (r  , c+1)
              atomicAdd(p + kSW + 1, __float_as_int(fmaf(w23[j].y, qs, M)) - 0x4B400000);  // (r+1, c+1)
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // (unrolled: a dynamic index would put the per-event arrays into local memory)
              if (w01[j].x != w01[j].x) {
                splat_event_exact<float>(iwe, Hp, Wp, pad_h, pad_w, e.x[j], w[j].x, w[j].y, 1.f);
                continue;
              }
              const int lr = rr[j] - r_org, lc = cc[j] - c_org;
              if (use_win && (unsigned)lr < (unsigned)(kSH - 1) && (unsigned)lc < (unsigned)(kSW - 1)) {
                int* p = win + off[j];
                atomicAdd(p, __float_as_int(fmaf(w01[j].x, qs, M)) - 0x4B400000);
                atomicAdd(p + kSW, __float_as_int(fmaf(w01[j].y, qs, M)) - 0x4B400000);
                atomicAdd(p + 1, __float_as_int(fmaf(w23[j].x, qs, M)) - 0x4B400000);
                atomicAdd(p + kSW + 1, __float_as_int(fmaf(w23[j].y, qs, M)) - 0x4B400000);
              } else {
                splat_taps_global(iwe, Hp, Wp, rr[j], cc[j], w01[j].x, w01[j].y, w23[j].x, w23[j].y);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) { e.x[j] = nxt.x[j]; e.y[j] = nxt.y[j]; e.d[j] = nxt.d[j]; }
        }
      }
      if (use_win) {
        __syncthreads();
        // window -> global: coalesced rows of fp32 REDs (an int32 -> fp32 conversion is one rounding)
        const float inv = __int_as_float((127 - S) << 23);
        for (int i = threadIdx.x; i < kSH * kSW; i += blockDim.x) {
          const int v = win[i];
          if (v != 0) {
            const int lr = i / kSW, lc = i - lr * kSW;
            const int r = r_org + lr, c = c_org + lc;
            if ((unsigned)r < (unsigned)Hp && (unsigned)c < (unsigned)Wp) red_add_nc(iwe + r * Wp + c, (float)v * inv);
          }
        }
      }
    }
  }
}

// ---- fixed-point tile splat, second version (round 2) --------------------------------------------------------
// SASS of k_tile_splat_d: of its ~69 instructions per event, ~25 are the origin-pixel unpack + the "gather the flow only
// when the pixel changes" branch ladder (two dependent global loads behind it), ~7 the zero-fill and flush of the whole
// 41 x 41 window per 4080-event item, 2 the register copies of the software pipeline.  This version
//  * stages the flow at the item's origin pixels, together with the pixel coordinates as floats, in shared memory
//    once per CTA: the per-event gather is ONE conflict-free LDS.128 (the lanes of a warp read the same few pixels:
//    broadcast) and the two int -> float conversions disappear;
//  * knows the item's pixel range (prepare stores first | last << 16 tile-local pixel in the item descriptor): an
//    item of ~224 pixels spans ~7 of the 32 tile rows, so only rows [row0, row1 + 2 * halo + 2) of the window are
//    zeroed, addressable and flushed, and only the item's pixels are put into the table;
//  * requests its first event group BEFORE that set-up, so that the DRAM latency of the stream and the L2 latency of
//    the table overlap; two register buffers used alternately replace the copies;
//  * converts to fixed point on packed FFMA2.
// MERGE = true additionally keeps the four tap sums of a run of events that stay in one IWE cell in registers and adds
// them to the window when the cell changes.  ncu r01g blamed shared-memory atomic wavefronts (8.83 M for 2.36 M ATOMS:
// the +-3 px random flow sends the 32 lanes to random banks); merging halves them, but ptxas never predicates ATOMS on
// sm_100a (a predicated `red.shared` becomes one BSSY / BRA / BSYNC region per instruction), a warp takes the flush
// branch at nearly every event step, and the kernel turned out to be bound by issue slots, not by wavefronts: measured
// on B200 at 16 Mi events 84 us against 72 us (profiles/README.md, round 2).  Kept as an A/B knob (EBOS_TILE=6).
constexpr int kMergeS = 17;   // a merged run is at most the 16 events of a thread: 16 * (1 + 2e-6) * 2^17 < 2^22

// four tap values of one cell -> fixed point (scale qs = 2^S) -> window
__device__ __forceinline__ void taps_to_window(int* __restrict__ win, int off, float2 a01, float2 a23, float qs) {
  const float M = 12582912.0f;   // 1.5 * 2^23: fma(w, 2^S, M) leaves round(w * 2^S) in the low mantissa bits
  const float2 q01 = fma2(a01, make_float2(qs, qs), make_float2(M, M));
  const float2 q23 = fma2(a23, make_float2(qs, qs), make_float2(M, M));
  int* p = win + off;
  atomicAdd(p, __float_as_int(q01.x) - 0x4B400000);            // (r  , c  )
  atomicAdd(p + kSW, __float_as_int(q01.y) - 0x4B400000);      // (r+1, c  )
  atomicAdd(p + 1, __float_as_int(q23.x) - 0x4B400000);        // (r  , c+1)
  atomicAdd(p + kSW + 1, __float_as_int(q23.y) - 0x4B400000);  // (r+1, c+1)
}

// tile-local index of an origin pixel from the packed (row << 16 | col) word: (row % 32) * 32 + col % 32
__device__ __forceinline__ int tile_local_pixel(unsigned rc) { return (int)(((rc >> 11) & 0x3e0u) | (rc & 31u)); }

// Geometry of one work item: event range, tile origin, window origin, and the rows of the window its pixels can reach.
struct ItemGeom {
  int lo, hi;          // events [lo, hi) of the sorted stream
  int tr0, tc0;        // image coordinates of the tile's first pixel
  int r_org, c_org;    // padded-image coordinates of window element (0, 0)
  int lp0, lp1;        // first / last tile-local origin pixel of the item
  int row0, nrows;     // window rows [row0, row0 + nrows) are in use (cells: one row fewer)
};
__device__ __forceinline__ ItemGeom item_geom(const int4 it, int pad_h, int pad_w) {
  ItemGeom g;
  g.lo = it.y; g.hi = it.z;
  g.tr0 = (it.w >> 16) * kTileH; g.tc0 = (it.w & 0xffff) * kTileW;
  g.r_org = g.tr0 - kHalo + pad_h; g.c_org = g.tc0 - kHalo + pad_w;
  g.lp0 = it.x & 0xffff; g.lp1 = (int)((unsigned)it.x >> 16);
  g.row0 = g.lp0 >> 5;
  g.nrows = (g.lp1 >> 5) - g.row0 + 2 * kHalo + 2;
  return g;
}

// (f0, f1, row, col) of the item's origin pixels -> shared memory
__device__ __forceinline__ void fill_pixel_table(float4* __restrict__ pix, const float* __restrict__ flow, const ItemGeom& g,
                                                 int H, int W, int hw) {
  for (int i = g.lp0 + (int)threadIdx.x; i <= g.lp1; i += blockDim.x) {
    const int r = g.tr0 + (i >> 5), c = g.tc0 + (i & 31);
    const bool in = r < H && c < W;
    const int k = in ? r * W + c : 0;
    pix[i] = make_float4(in ? __ldg(flow + k) : 0.f, in ? __ldg(flow + hw + k) : 0.f, (float)r, (float)c);
  }
}

// One event outside the regular case of the tile splat (item boundary, cell outside the window rows in use, non-finite,
// item too small for the window): same arithmetic as the fast path, recomputed from the warped coordinate so that the
// fast path keeps nothing alive for it.  x0 = NaN marks an event that belongs to the neighbouring item.
__device__ __noinline__ void splat_event_general(float* __restrict__ iwe, int Hp, int Wp, int pad_h, int pad_w, float x0,
                                                 float xw, float yw, int* __restrict__ win, int r_org, int c_org, int row0,
                                                 int nrows, float qs) {
  if (x0 != x0) return;
  const float2 w = make_float2(xw, yw);
  const float2 wb = add2(w, make_float2(1e-6f, 1e-6f));
  const float fr = floorf(wb.x), fc = floorf(wb.y);
  const float2 ab = sub2(w, make_float2(fr, fc));
  const float2 nab = sub2(make_float2(1.f, 1.f), ab);
  const float2 lhs = make_float2(nab.x, ab.x);
  const float2 w01 = mul2(lhs, make_float2(nab.y, nab.y));
  const float2 w23 = mul2(lhs, make_float2(ab.y, ab.y));
  if (w01.x != w01.x) {   // non-finite warped coordinate: the reference's NaN lands on pixel 0
    splat_event_exact<float>(iwe, Hp, Wp, pad_h, pad_w, x0, xw, yw, 1.f);
    return;
  }
  const int r = (int)fr + pad_h, c = (int)fc + pad_w;
  const int lr = r - r_org, lc = c - c_org;
  if (nrows > 0 && (unsigned)(lr - row0) < (unsigned)(nrows - 1) && (unsigned)lc < (unsigned)(kSW - 1))
    taps_to_window(win, lr * kSW + lc, w01, w23, qs);
  else
    flush_cell<float, false>(iwe, Hp, Wp, Hp - 1, Wp - 1, r, c, w01.x, w01.y, w23.x, w23.y);
}

template <bool PACKED, int MINB, bool MERGE>
__global__ void __launch_bounds__(256, MINB)
k_tile_splat_m(const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ sd,
               const int4* __restrict__ items, const WindowHeader* __restrict__ hdr, const float* __restrict__ flow,
               int H, int W, int pad_h, int pad_w, float* __restrict__ iwe, int nb) {
  __shared__ int win[kSH * kSW];
  __shared__ float4 pix[kTileH * kTileW];
  pdl_launch_dependents();   // the cost kernel may be scheduled while this grid drains (it waits before reading the IWE)
  ItemGeom g = item_geom(__ldg(items + blockIdx.x), pad_h, pad_w);   // (unused slots are empty: no header read before it)
  const int lo = g.lo, hi = g.hi;
  if (hi <= lo) return;
  const int base0 = (nb > 0 ? (lo & ~(kBlockEvents - 1)) : (lo & ~3)) + (int)threadIdx.x * 16;
  // two event buffers used alternately: the raw fields of group g+1 are requested before group g is processed; the very
  // first group is requested before the set-up below
  EventBlock<float, 4, false, PACKED> ea, eb;
  if (base0 < hi) ea.load_range32(sx, sy, sd, base0, lo, hi, nb);
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w, hw = H * W;
  if (hi - lo < kWinMinEvents) g.nrows = 0;                  // small item: every event takes the global path
  const int cnt_log = 31 - __clz(hi - lo);
  const int S = MERGE ? kMergeS : min(21, 30 - cnt_log);     // the int32 window cannot overflow: cnt * 2^S < 2^31
  const float qs = __int_as_float((127 + S) << 23);
  for (int i = threadIdx.x; i < g.nrows * kSW; i += blockDim.x) win[g.row0 * kSW + i] = 0;
  fill_pixel_table(pix, flow, g, H, W, hw);
  __syncthreads();
  const int dr = pad_h - g.r_org - g.row0, dc = pad_w - g.c_org;   // (cell row relative to row0, cell column) = floor + (dr, dc)
  const unsigned span = (unsigned)max(g.nrows - 1, 0);
  const int off0 = g.row0 * kSW;
  const float2 bias2 = make_float2(1e-6f, 1e-6f);
  if (base0 < hi) {
    // MERGE: pending run = window offset of its cell (-1: none) and the four tap sums
    int poff = -1;
    float2 pa01 = make_float2(0.f, 0.f), pa23 = make_float2(0.f, 0.f);
    auto process = [&](EventBlock<float, 4, false, PACKED>& e, const int b) {
      const bool full = b >= lo && b + 4 <= hi;
      float2 w[4], w01[4], w23[4];
      int off[4];
      bool ok = full;
      if (full) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int lp;
          if constexpr (PACKED) lp = tile_local_pixel(__float_as_uint(e.x[j]));
          else lp = (((int)e.x[j] & 31) << 5) | ((int)e.y[j] & 31);
          const float4 q = pix[lp];
          const float xo = PACKED ? q.z : e.x[j], yo = PACKED ? q.w : e.y[j];
          // x' = x - (dt * f) with two roundings (see splat_block_f32)
          w[j] = make_float2(__fsub_rn(xo, __fmul_rn(e.d[j], q.x)), __fsub_rn(yo, __fmul_rn(e.d[j], q.y)));
        }
      } else {
        e.finish(flow, W, hw);   // item boundary: events of the neighbouring item are marked (x = NaN)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          w[j] = make_float2(__fsub_rn(e.x[j], __fmul_rn(e.d[j], e.f0[j])), __fsub_rn(e.y[j], __fmul_rn(e.d[j], e.f1[j])));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 wb = add2(w[j], bias2);
        const float fr = floorf(wb.x), fc = floorf(wb.y);
        const float2 ab = sub2(w[j], make_float2(fr, fc));                                            // (a, b)
        // (1 - a, a) in one packed FMA: a * (-1) + 1 is ONE rounding of 1 - a, a * 1 + 0 is a
        const float2 lhs = fma2(make_float2(ab.x, ab.x), make_float2(-1.f, 1.f), make_float2(1.f, 0.f));
        const float nb = __fsub_rn(1.f, ab.y);
        w01[j] = mul2(lhs, make_float2(nb, nb));
        w23[j] = mul2(lhs, make_float2(ab.y, ab.y));
        const int lr = (int)fr + dr, lc = (int)fc + dc;
        off[j] = lr * kSW + lc + off0;
        // (a NaN weight fails the first test; a huge coordinate saturates the conversion and fails the range tests)
        ok = ok && (w01[j].x == w01[j].x) && (unsigned)lr < span && (unsigned)lc < (unsigned)(kSW - 1);
      }
      if (ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (MERGE) {
            const bool same = off[j] == poff;
            if (!same & (poff >= 0)) taps_to_window(win, poff, pa01, pa23, qs);
            const float sel = same ? 1.f : 0.f;                 // 1 * sum + w is one rounding of sum + w
            pa01 = fma2(make_float2(sel, sel), pa01, w01[j]);
            pa23 = fma2(make_float2(sel, sel), pa23, w23[j]);
            poff = off[j];
          } else {
            taps_to_window(win, off[j], w01[j], w23[j], qs);
          }
        }
      } else {
        if constexpr (MERGE) {
          if (poff >= 0) taps_to_window(win, poff, pa01, pa23, qs);
          poff = -1;
          pa01 = make_float2(0.f, 0.f); pa23 = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          splat_event_general(iwe, Hp, Wp, pad_h, pad_w, full ? 0.f : e.x[j], w[j].x, w[j].y, win, g.r_org, g.c_org, g.row0,
                              g.nrows, qs);
      }
    };
    // a thread's groups: 4 consecutive groups of 4 events per run of 16, runs 256 * 16 events apart; the raw fields of
    // the next group (also across runs) are requested before the current one is processed
    auto group_base = [&](int gi) { return base0 + (gi >> 2) * (int)(blockDim.x * 16) + (gi & 3) * 4; };
#pragma unroll 1
    for (int gi = 0;; gi += 2) {
      const int b0 = group_base(gi);
      if (b0 >= hi) break;
      const int b1 = group_base(gi + 1);
      if (b1 < hi) eb.load_range32(sx, sy, sd, b1, lo, hi, nb);
      if (b0 + 4 > lo) process(ea, b0);   // (blocked windows: groups in front of the item belong to its predecessor)
      if (b1 >= hi) break;
      const int b2 = group_base(gi + 2);
      if (b2 < hi) ea.load_range32(sx, sy, sd, b2, lo, hi, nb);
      if (b1 + 4 > lo) process(eb, b1);
    }
    if constexpr (MERGE) { if (poff >= 0) taps_to_window(win, poff, pa01, pa23, qs); }
  }
  if (g.nrows > 0) {
    __syncthreads();
    // window rows in use -> global: coalesced fp32 REDs (an int32 -> fp32 conversion is one rounding)
    const float inv = __int_as_float((127 - S) << 23);
    for (int i = threadIdx.x; i < g.nrows * kSW; i += blockDim.x) {
      const int v = win[off0 + i];
      if (v != 0) {
        const int lr = i / kSW, lc = i - lr * kSW;
        const int r = g.r_org + g.row0 + lr, c = g.c_org + lc;
        if ((unsigned)r < (unsigned)Hp && (unsigned)c < (unsigned)Wp) red_add_nc(iwe + r * Wp + c, (float)v * inv);
      }
    }
  }
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

// EPT consecutive events of one thread, processed in groups of 4: (phase 1) cells and fractions, then all
// gathers of dL/dIWE of the group in flight together; (phase 2) gradients, combined over runs of the same
// origin pixel (the run state is carried across groups), one REDG pair per pixel run.
template <typename T> struct BwdParams {
  int Hp, Wp, pad_h, pad_w, hw, lo;
  unsigned r_span, c_span;
  VarCoef<T> vc;
};
template <typename T, int GSRC>
__device__ __forceinline__ BwdParams<T> make_bwd_params(int H, int W, int pad_h, int pad_w, const double* acc, int omit,
                                                        double scale) {
  BwdParams<T> P;
  P.Hp = H + 2 * pad_h; P.Wp = W + 2 * pad_w; P.pad_h = pad_h; P.pad_w = pad_w; P.hw = H * W;
  P.vc = VarCoef<T>{(T)0, (T)0, omit};
  if (GSRC == 1) {
    const double cnt = omit ? (double)(P.Hp - 2) * (double)(P.Wp - 2) : (double)P.Hp * (double)P.Wp;
    P.vc.mean = (T)(acc[0] / cnt);
    P.vc.cv = (T)(-2.0 * scale / (cnt - 1.0));
  }
  // fast range of cells whose four taps are all inside (and, for the cropped variance, all counted)
  P.lo = (GSRC == 1 && omit) ? 1 : 0;
  P.r_span = (unsigned)max(P.Hp - 1 - 2 * P.lo, 0);
  P.c_span = (unsigned)max(P.Wp - 1 - 2 * P.lo, 0);
  return P;
}

template <typename T, int GSRC, bool HAS_W, int EPT, bool VEC, bool PACKED>
__device__ __forceinline__ void bwd_block(const EventBlock<T, EPT, HAS_W, PACKED>& e, const BwdParams<T>& P,
                                          const T* __restrict__ g, T* __restrict__ dflow) {
  constexpr int G = 4;
  static_assert(EPT % G == 0, "EPT must be a multiple of the gather group");
  int ck = -1;
  T s0 = 0, s1 = 0;
#pragma unroll
  for (int h = 0; h < EPT; h += G) {
    T a[G], b[G], g00[G], g01[G], g10[G], g11[G];
    bool fast[G];
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int j = h + i;
      T xw, yw, fr, fc;
      if constexpr (sizeof(T) == 4) {
        const float2 w = make_float2(__fsub_rn(e.x[j], __fmul_rn(e.d[j], e.f0[j])), __fsub_rn(e.y[j], __fmul_rn(e.d[j], e.f1[j])));
        const float2 wb = add2(w, make_float2(1e-6f, 1e-6f));
        fr = floorf(wb.x); fc = floorf(wb.y);
        const float2 ab = sub2(w, make_float2(fr, fc));
        xw = w.x; yw = w.y; a[i] = ab.x; b[i] = ab.y;
      } else {
        xw = Rn<T>::sub(e.x[j], Rn<T>::mul(e.d[j], e.f0[j]));
        yw = Rn<T>::sub(e.y[j], Rn<T>::mul(e.d[j], e.f1[j]));
        fr = FastFloor<T>::flr(Rn<T>::add(xw, Rn<T>::bias())); fc = FastFloor<T>::flr(Rn<T>::add(yw, Rn<T>::bias()));
        a[i] = Rn<T>::sub(xw, fr);
        b[i] = Rn<T>::sub(yw, fc);
      }
      // saturating conversions: a huge / Inf coordinate lands outside the fast range; NaN converts to cell 0
      // and is caught by the fraction test (NaN a or b)
      const int r = FastFloor<T>::to_int(fr) + P.pad_h, c = FastFloor<T>::to_int(fc) + P.pad_w;
      fast[i] = (unsigned)(r - P.lo) < P.r_span && (unsigned)(c - P.lo) < P.c_span && (a[i] + b[i] == a[i] + b[i]);
      if (fast[i] && (g_ablate & 8)) {  // diagnostics: no dL/dIWE gathers
        g00[i] = a[i]; g01[i] = b[i]; g10[i] = (T)r; g11[i] = (T)c;
      } else if (fast[i]) {
        const T* p = g + (r * P.Wp + c);
        load_pair<VEC>(p, c & 3, g00[i], g01[i]);
        load_pair<VEC>(p + P.Wp, c & 3, g10[i], g11[i]);
      } else {
        // rare: border cell, out-of-range or skipped event -- exact masked gathers, result kept in g00/g01
        T dx, dy;
        bwd_event_exact<T, GSRC>(g, P.Hp, P.Wp, P.pad_h, P.pad_w, e.x[j] == e.x[j] ? xw : (T)NAN, yw, P.vc, dx, dy);
        g00[i] = dx; g01[i] = dy; g10[i] = 0; g11[i] = 0; a[i] = 0; b[i] = 0;
      }
    }
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int j = h + i;
      T dx, dy;
      if (fast[i]) {
        if constexpr (sizeof(T) == 4) {
          // (dx, dy) = (1-b, 1-a) * (g10-g00, g01-g00) + (b, a) * (g11-g01, g11-g10)   on packed f32x2
          const float2 d1 = sub2(make_float2(g10[i], g01[i]), make_float2(g00[i], g00[i]));
          const float2 d2 = sub2(make_float2(g11[i], g11[i]), make_float2(g01[i], g10[i]));
          const float2 ba = make_float2(b[i], a[i]);
          const float2 r = fma2(ba, d2, mul2(sub2(make_float2(1.f, 1.f), ba), d1));
          dx = r.x; dy = r.y;
        } else {
          dx = ((T)1 - b[i]) * (g10[i] - g00[i]) + b[i] * (g11[i] - g01[i]);
          dy = ((T)1 - a[i]) * (g01[i] - g00[i]) + a[i] * (g11[i] - g10[i]);
        }
        if (GSRC == 1) { dx *= P.vc.cv; dy *= P.vc.cv; }  // differences: the mean cancels
      } else {
        dx = g00[i]; dy = g01[i];
      }
      if (HAS_W) { dx *= e.wt[j]; dy *= e.wt[j]; }
      if (e.x[j] != e.x[j]) continue;  // skipped event
      if (e.k[j] != ck) {
        if (ck >= 0 && !(g_ablate & 1)) { red_add_nc(dflow + ck, s0); red_add_nc(dflow + P.hw + ck, s1); }
        else if (ck >= 0 && s0 + s1 == (T)-12345) dflow[0] = s0;
        ck = e.k[j]; s0 = 0; s1 = 0;
      }
      s0 -= e.d[j] * dx;
      s1 -= e.d[j] * dy;
    }
  }
  if (ck >= 0 && !(g_ablate & 1)) { red_add_nc(dflow + ck, s0); red_add_nc(dflow + P.hw + ck, s1); }
  else if (ck >= 0 && s0 + s1 == (T)-12345) dflow[0] = s0;
}

template <typename T, int GSRC, bool HAS_W, int EPT, bool VEC, bool PACKED>
__global__ void __launch_bounds__(256, (sizeof(T) == 4 && EPT <= 4) ? 4 : 1)
k_win_bwd(const T* __restrict__ sx, const T* __restrict__ sy, const T* __restrict__ sd, const T* __restrict__ sw,
          int64_t n, const T* __restrict__ flow, int H, int W, int pad_h, int pad_w, const T* __restrict__ g,
          const double* __restrict__ acc, int omit, double scale, T* __restrict__ dflow) {
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * EPT;
  if (base >= n) return;
  const BwdParams<T> P = make_bwd_params<T, GSRC>(H, W, pad_h, pad_w, acc, omit, scale);
  EventBlock<T, EPT, HAS_W, PACKED> e;
  e.load(sx, sy, sd, sw, base, n, flow, W, P.hw);
  bwd_block<T, GSRC, HAS_W, EPT, VEC, PACKED>(e, P, g, dflow);
}

template <int GSRC, bool HAS_W, bool VEC, bool PACKED>
__global__ void __launch_bounds__(256, 3)
k_win_bwd_pipe(const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ sd,
               const float* __restrict__ sw, int64_t n, const float* __restrict__ flow, int H, int W, int pad_h,
               int pad_w, const float* __restrict__ g, const double* __restrict__ acc, int omit, double scale,
               float* __restrict__ dflow) {
  const BwdParams<float> P = make_bwd_params<float, GSRC>(H, W, pad_h, pad_w, acc, omit, scale);
  pipe_loop<HAS_W, PACKED>(sx, sy, sd, sw, n, flow, W, P.hw, [&](const EventBlock<float, kPipeEpt, HAS_W, PACKED>& e) {
    bwd_block<float, GSRC, HAS_W, kPipeEpt, VEC, PACKED>(e, P, g, dflow);
  });
}

// Run state of the backward carried across groups: current origin pixel + gradient sums, and the last
// gathered cell with its four dL/dIWE values.  Consecutive events of a pixel mostly fall into the same
// IWE cell, so the four gathers are only issued when the cell changes (the kernel is L1-bound on them:
// ncu r01c l1tex 78 %).
struct BwdRun { int ck; float2 s01; int pr, pc; float g00, g01, g10, g11; };

template <int GSRC, bool HAS_W, bool PACKED>
__device__ __forceinline__ void bwd_group4(const EventBlock<float, 4, HAS_W, PACKED>& e, BwdRun& run,
                                           const BwdParams<float>& P, const float* __restrict__ g, float* __restrict__ dflow) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 w = make_float2(__fsub_rn(e.x[i], __fmul_rn(e.d[i], e.f0[i])), __fsub_rn(e.y[i], __fmul_rn(e.d[i], e.f1[i])));
    const float2 wb = add2(w, make_float2(1e-6f, 1e-6f));
    const float fr = floorf(wb.x), fc = floorf(wb.y);
    const float2 ab = sub2(w, make_float2(fr, fc));
    const int r = (int)fr + P.pad_h, c = (int)fc + P.pad_w;
    const bool fast = (unsigned)(r - P.lo) < P.r_span && (unsigned)(c - P.lo) < P.c_span && (ab.x + ab.y == ab.x + ab.y);
    float2 dxy;
    if (fast) {
      if (r != run.pr || c != run.pc) {
        const float* p = g + (r * P.Wp + c);
        run.g00 = __ldg(p); run.g01 = __ldg(p + 1); run.g10 = __ldg(p + P.Wp); run.g11 = __ldg(p + P.Wp + 1);
        run.pr = r; run.pc = c;
      }
      // (dx, dy) = (1-b, 1-a) * (g10-g00, g01-g00) + (b, a) * (g11-g01, g11-g10)
      const float2 d1 = sub2(make_float2(run.g10, run.g01), make_float2(run.g00, run.g00));
      const float2 d2 = sub2(make_float2(run.g11, run.g11), make_float2(run.g01, run.g10));
      const float2 ba = make_float2(ab.y, ab.x);
      dxy = fma2(ba, d2, mul2(sub2(make_float2(1.f, 1.f), ba), d1));
      if (GSRC == 1) dxy = mul2(dxy, make_float2(P.vc.cv, P.vc.cv));  // differences: the mean cancels
    } else {
      // rare: border cell, out-of-range or skipped event -- exact masked gathers
      float dx, dy;
      bwd_event_exact<float, GSRC>(g, P.Hp, P.Wp, P.pad_h, P.pad_w, e.x[i] == e.x[i] ? w.x : NAN, w.y, P.vc, dx, dy);
      dxy = make_float2(dx, dy);
    }
    if (HAS_W) dxy = mul2(dxy, make_float2(e.wt[i], e.wt[i]));
    if (e.x[i] != e.x[i]) continue;  // skipped event
    if (e.k[i] != run.ck) {
      if (run.ck >= 0) { red_add_nc(dflow + run.ck, run.s01.x); red_add_nc(dflow + P.hw + run.ck, run.s01.y); }
      run.ck = e.k[i];
      run.s01 = make_float2(0.f, 0.f);
    }
    run.s01 = fma2(make_float2(-e.d[i], -e.d[i]), dxy, run.s01);
  }
}

// Group of four events that lies entirely inside the stream (no end markers): the cells of all four are computed
// first; when every one is regular (all taps inside, finite) the per-event exact-path branches disappear from the
// instruction stream.  Returns false (nothing done) when the group needs the general code.
template <int GSRC, bool HAS_W, bool PACKED>
__device__ __forceinline__ bool bwd_group4_regular(const EventBlock<float, 4, HAS_W, PACKED>& e, BwdRun& run,
                                                   const BwdParams<float>& P, const float* __restrict__ g,
                                                   float* __restrict__ dflow) {
  float2 ab[4];
  int r[4], c[4];
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 w = make_float2(__fsub_rn(e.x[i], __fmul_rn(e.d[i], e.f0[i])), __fsub_rn(e.y[i], __fmul_rn(e.d[i], e.f1[i])));
    const float2 wb = add2(w, make_float2(1e-6f, 1e-6f));
    const float fr = floorf(wb.x), fc = floorf(wb.y);
    ab[i] = sub2(w, make_float2(fr, fc));
    r[i] = (int)fr + P.pad_h; c[i] = (int)fc + P.pad_w;
    ok = ok && (unsigned)(r[i] - P.lo) < P.r_span && (unsigned)(c[i] - P.lo) < P.c_span && (ab[i].x + ab[i].y == ab[i].x + ab[i].y);
  }
  if (!ok) return false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (r[i] != run.pr || c[i] != run.pc) {
      const float* p = g + (r[i] * P.Wp + c[i]);
      run.g00 = __ldg(p); run.g01 = __ldg(p + 1); run.g10 = __ldg(p + P.Wp); run.g11 = __ldg(p + P.Wp + 1);
      run.pr = r[i]; run.pc = c[i];
    }
    // (dx, dy) = (1-b, 1-a) * (g10-g00, g01-g00) + (b, a) * (g11-g01, g11-g10)
    const float2 d1 = sub2(make_float2(run.g10, run.g01), make_float2(run.g00, run.g00));
    const float2 d2 = sub2(make_float2(run.g11, run.g11), make_float2(run.g01, run.g10));
    const float2 ba = make_float2(ab[i].y, ab[i].x);
    float2 dxy = fma2(ba, d2, mul2(sub2(make_float2(1.f, 1.f), ba), d1));
    if (GSRC == 1) dxy = mul2(dxy, make_float2(P.vc.cv, P.vc.cv));  // differences: the mean cancels
    if (HAS_W) dxy = mul2(dxy, make_float2(e.wt[i], e.wt[i]));
    if (e.k[i] != run.ck) {
      if (run.ck >= 0) { red_add_nc(dflow + run.ck, run.s01.x); red_add_nc(dflow + P.hw + run.ck, run.s01.y); }
      run.ck = e.k[i];
      run.s01 = make_float2(0.f, 0.f);
    }
    run.s01 = fma2(make_float2(-e.d[i], -e.d[i]), dxy, run.s01);
  }
  return true;
}

// (4 CTAs/SM = 64 registers: with the prefetched group and the regular-group fast path the 48-register build spills)
template <int GSRC, bool HAS_W, bool PACKED, int NG, int MINB = 4>
__global__ void __launch_bounds__(256, MINB)
k_win_bwd_g(const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ sd,
            const float* __restrict__ sw, int64_t n, const float* __restrict__ flow, int H, int W, int pad_h, int pad_w,
            const float* __restrict__ g, const double* __restrict__ acc, int omit, double scale,
            float* __restrict__ dflow, int nb) {
  pdl_launch_dependents();   // (fused solver iteration: Adam follows and waits before reading dflow)
  // Blocked-striped windows (nb > 0, NG in {1, 2, 4}): lane L of block B owns the 16 logical events B * 512 + L * 16 ...;
  // a warp takes the SAME run of 4 * NG events of all 32 lanes of one block, so that its group loads are contiguous.
  const int64_t tid_g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t base = tid_g * (4 * NG);
  if (NG <= 4 && base < nb) {
    constexpr int kRuns = 4 / (NG <= 4 ? NG : 4);                  // runs of 4 * NG events per lane slot
    const int64_t warp = tid_g >> 5;
    base = (warp / kRuns) * kBlockEvents + (tid_g & 31) * 16 + (warp % kRuns) * (4 * NG);
  }
  if (base >= n) return;
  if (GSRC == 1) pdl_wait();   // the variance coefficients come from the accumulators of the preceding cost kernel
  const BwdParams<float> P = make_bwd_params<float, GSRC>(H, W, pad_h, pad_w, acc, omit, scale);
  BwdRun run{-1, make_float2(0.f, 0.f), INT_MIN, INT_MIN, 0.f, 0.f, 0.f, 0.f};
  // Software pipelining of the event stream: the raw fields of group g+1 are requested before group g is processed.
  // ncu r01e showed each group start exposing three dependent latencies (event load -> flow gather -> dL/dIWE
  // gather); this removes the first (and longest, DRAM) one from every group but a thread's first.
  EventBlock<float, 4, HAS_W, PACKED> cur, nxt;
  cur.load_group(sx, sy, sd, sw, base, n, nb);
  // PDL: everything above (and the event loads in flight) overlaps the drain of the preceding cost kernel; dL/dIWE
  // and the dflow accumulation target are only touched below
  if (GSRC == 0) pdl_wait();
#pragma unroll
  for (int gi = 0; gi < NG; ++gi) {
    const int64_t b = base + 4 * gi;
    if (b >= n) break;
    if (gi + 1 < NG && b + 4 < n) nxt.load_group(sx, sy, sd, sw, b + 4, n, nb);
    if (b + 4 <= n) {
      cur.finish_full(flow, W, P.hw);
      if (!bwd_group4_regular<GSRC, HAS_W, PACKED>(cur, run, P, g, dflow)) bwd_group4<GSRC, HAS_W, PACKED>(cur, run, P, g, dflow);
    } else {
      if (g_ablate & 2) cur.finish_nogather(W, P.hw); else cur.finish(flow, W, P.hw);
      bwd_group4<GSRC, HAS_W, PACKED>(cur, run, P, g, dflow);
    }
    if (gi + 1 < NG) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { cur.x[j] = nxt.x[j]; cur.y[j] = nxt.y[j]; cur.d[j] = nxt.d[j]; if (HAS_W) cur.wt[j] = nxt.wt[j]; }
    }
  }
  if (run.ck >= 0) { red_add_nc(dflow + run.ck, run.s01.x); red_add_nc(dflow + P.hw + run.ck, run.s01.y); }
}

// ---- shared-memory tile backward, second version (round 2) ----------------------------------------------------
// ncu r01g of k_win_bwd_g: L1TEX 52 % busy on 20 M gather sectors, `long_scoreboard` the top stall -- three dependent
// global latencies per group (event load -> flow gather -> dL/dIWE gather).  Here one CTA takes one work item of the
// tile sort (like the splat): the rows of the tile's dL/dIWE window that the item's pixels can reach (masked and --
// for the variance objective -- already transformed) and the per-pixel table (flow + float coordinates) are staged in
// shared memory once, so that both gathers of an event are LDS (the four taps only when the cell changes), and the
// only global traffic per event is the prefetched stream itself.  All events of an origin pixel are combined in
// registers; the per-pixel flush is one pair of global REDs per pixel run.  The first tile backward (round 1) lost to
// the one-shot kernel because of its coarse persistent items, the whole 41 x 41 window filled per item, 8 events of
// state in registers (80 registers) and no prefetch.
// One event outside the regular case (item boundary handled by the caller; here: cell outside the window rows in use,
// non-finite coordinate): exact masked gathers from global memory.
template <int GSRC>
__device__ __noinline__ float2 bwd_event_general(const float* __restrict__ g, const BwdParams<float>& P, float xw, float yw) {
  float dx, dy;
  bwd_event_exact<float, GSRC>(g, P.Hp, P.Wp, P.pad_h, P.pad_w, xw, yw, P.vc, dx, dy);
  return make_float2(dx, dy);
}

template <int GSRC, bool PACKED, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_tile_bwd_m(const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ sd,
             const int4* __restrict__ items, const WindowHeader* __restrict__ hdr, const float* __restrict__ flow,
             int H, int W, int pad_h, int pad_w, const float* __restrict__ g, const double* __restrict__ acc, int omit,
             double scale, float* __restrict__ dflow, int nb) {
  __shared__ float gwin[kSH * kSW];
  __shared__ float4 pix[kTileH * kTileW];
  pdl_launch_dependents();   // (fused solver iteration: Adam follows and waits before reading dflow)
  const ItemGeom ig = item_geom(__ldg(items + blockIdx.x), pad_h, pad_w);   // (unused slots are empty)
  const int lo = ig.lo, hi = ig.hi;
  if (hi <= lo) return;
  const int hw = H * W;
  const int base0 = (nb > 0 ? (lo & ~(kBlockEvents - 1)) : (lo & ~3)) + (int)threadIdx.x * 16;
  EventBlock<float, 4, false, PACKED> ea, eb;
  if (base0 < hi) ea.load_range32(sx, sy, sd, base0, lo, hi, nb);
  fill_pixel_table(pix, flow, ig, H, W, hw);
  // PDL: everything above overlaps the drain of the preceding cost kernel; dL/dIWE (and, for the variance objective,
  // the accumulators) are only read below
  pdl_wait();
  const BwdParams<float> P = make_bwd_params<float, GSRC>(H, W, pad_h, pad_w, acc, omit, scale);
  const int off0 = ig.row0 * kSW;
  for (int i = threadIdx.x; i < ig.nrows * kSW; i += blockDim.x) {
    const int lr = i / kSW, lc = i - lr * kSW;
    gwin[off0 + i] = fetch_g<float, GSRC>(g, P.Hp, P.Wp, ig.r_org + ig.row0 + lr, ig.c_org + lc, P.vc);
  }
  __syncthreads();
  const int dr = pad_h - ig.r_org - ig.row0, dc = pad_w - ig.c_org;
  const unsigned span = (unsigned)(ig.nrows - 1);
  const float2 bias2 = make_float2(1e-6f, 1e-6f), one2 = make_float2(1.f, 1.f);
  if (base0 < hi) {
    int poff = -1;                                   // cell whose four dL/dIWE values are in registers
    float g00 = 0.f, g01 = 0.f, g10 = 0.f, g11 = 0.f;
    int pk = -1;                                     // origin pixel of the pending gradient sums
    float2 s01 = make_float2(0.f, 0.f);
    auto process = [&](EventBlock<float, 4, false, PACKED>& e, const int b) {
      const bool full = b >= lo && b + 4 <= hi;
      float2 w[4], ab[4];
      int off[4], kk[4];
      bool inw[4];
      bool ok = full;
      if (full) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int lp;
          if constexpr (PACKED) {
            const unsigned rc = __float_as_uint(e.x[j]);
            lp = tile_local_pixel(rc);
            kk[j] = (int)(rc >> 16) * W + (int)(rc & 0xffffu);
          } else {
            const int r = (int)e.x[j], c = (int)e.y[j];
            lp = ((r & 31) << 5) | (c & 31);
            kk[j] = r * W + c;
          }
          const float4 q = pix[lp];
          const float xo = PACKED ? q.z : e.x[j], yo = PACKED ? q.w : e.y[j];
          w[j] = make_float2(__fsub_rn(xo, __fmul_rn(e.d[j], q.x)), __fsub_rn(yo, __fmul_rn(e.d[j], q.y)));
        }
      } else {
        e.finish(flow, W, hw);   // item boundary: events of the neighbouring item are marked (x = NaN)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          kk[j] = e.k[j];
          w[j] = make_float2(__fsub_rn(e.x[j], __fmul_rn(e.d[j], e.f0[j])), __fsub_rn(e.y[j], __fmul_rn(e.d[j], e.f1[j])));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 wb = add2(w[j], bias2);
        const float fr = floorf(wb.x), fc = floorf(wb.y);
        ab[j] = sub2(w[j], make_float2(fr, fc));
        const int lr = (int)fr + dr, lc = (int)fc + dc;
        off[j] = lr * kSW + lc + off0;
        // inside the window rows in use every tap is either in the image or reads the masked zero the window holds
        // for it; NaN fractions (non-finite coordinates, marked events) and other cells take the exact path
        inw[j] = (unsigned)lr < span && (unsigned)lc < (unsigned)(kSW - 1) && (ab[j].x + ab[j].y == ab[j].x + ab[j].y);
        ok = ok && inw[j];
      }
      if (ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (off[j] != poff) {
            const float* p = gwin + off[j];
            g00 = p[0]; g01 = p[1]; g10 = p[kSW]; g11 = p[kSW + 1];
            poff = off[j];
          }
          // (dx, dy) = (1-b, 1-a) * (g10-g00, g01-g00) + (b, a) * (g11-g01, g11-g10)
          const float2 d1 = sub2(make_float2(g10, g01), make_float2(g00, g00));
          const float2 d2 = sub2(make_float2(g11, g11), make_float2(g01, g10));
          const float2 ba = make_float2(ab[j].y, ab[j].x);
          const float2 dxy = fma2(ba, d2, mul2(sub2(one2, ba), d1));
          const bool same = kk[j] == pk;
          if (!same & (pk >= 0)) { red_add_nc(dflow + pk, s01.x); red_add_nc(dflow + hw + pk, s01.y); }
          const float sel = same ? 1.f : 0.f;
          s01 = fma2(make_float2(-e.d[j], -e.d[j]), dxy, mul2(make_float2(sel, sel), s01));
          pk = kk[j];
        }
      } else {
        poff = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!full && e.x[j] != e.x[j]) continue;   // marked event (belongs to the neighbouring item)
          float2 dxy;
          if (inw[j]) {
            const float* p = gwin + off[j];
            const float h00 = p[0], h01 = p[1], h10 = p[kSW], h11 = p[kSW + 1];
            const float2 d1 = sub2(make_float2(h10, h01), make_float2(h00, h00));
            const float2 d2 = sub2(make_float2(h11, h11), make_float2(h01, h10));
            const float2 ba = make_float2(ab[j].y, ab[j].x);
            dxy = fma2(ba, d2, mul2(sub2(one2, ba), d1));
          } else {
            dxy = bwd_event_general<GSRC>(g, P, w[j].x, w[j].y);
          }
          if (kk[j] != pk) {
            if (pk >= 0) { red_add_nc(dflow + pk, s01.x); red_add_nc(dflow + hw + pk, s01.y); }
            pk = kk[j];
            s01 = make_float2(0.f, 0.f);
          }
          s01 = fma2(make_float2(-e.d[j], -e.d[j]), dxy, s01);
        }
      }
    };
    auto group_base = [&](int gi) { return base0 + (gi >> 2) * (int)(blockDim.x * 16) + (gi & 3) * 4; };
#pragma unroll 1
    for (int gi = 0;; gi += 2) {
      const int b0 = group_base(gi);
      if (b0 >= hi) break;
      const int b1 = group_base(gi + 1);
      if (b1 < hi) eb.load_range32(sx, sy, sd, b1, lo, hi, nb);
      if (b0 + 4 > lo) process(ea, b0);   // (blocked windows: groups in front of the item belong to its predecessor)
      if (b1 >= hi) break;
      const int b2 = group_base(gi + 2);
      if (b2 < hi) ea.load_range32(sx, sy, sd, b2, lo, hi, nb);
      if (b1 + 4 > lo) process(eb, b1);
    }
    if (pk >= 0) { red_add_nc(dflow + pk, s01.x); red_add_nc(dflow + hw + pk, s01.y); }
  }
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
  WindowLayout L = window_layout(n, sizeof(T), H, W);
  char* b = reinterpret_cast<char*>(window);
  int* perm = reinterpret_cast<int*>(b + L.off_perm);
  int bx = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
  if (!tminmax) k_win_minmax<T><<<bx, 256, 0, st>>>(events, n, hdr);
  k_win_hdr_final<T><<<1, 1, 0, st>>>(hdr, direction, direction_frac, tminmax);
  k_win_keys<T><<<bx, 256, 0, st>>>(events, n, H, W, k_in, i_in, status, hdr);
  k_win_layout<<<1, 1, 0, st>>>(hdr, allow_packed && sizeof(T) == 4 && H <= 65536 && W <= 65536, status);
  size_t cub_bytes = ws.cub_bytes;
  cudaError_t ce = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, k_in, k_out, i_in, perm, (int)n, 0,
                                                   key_bits_for((int64_t)tiles_x(W) * tiles_y(H) * kTileH * kTileW + 1), st);
  if (ce != cudaSuccess) return cuda_fail(ce, "ebos_window_prepare(sort)");
  const int n_tiles = tiles_x(W) * tiles_y(H);
  int* tile_off = reinterpret_cast<int*>(b + L.off_tiles);
  k_win_tile_offsets<<<(n_tiles + 1 + 255) / 256, 256, 0, st>>>(k_out, n, n_tiles, tile_off);
  // unused item slots stay empty (begin == end == 0): the tile kernels launch one CTA per slot and read only their slot
  cudaError_t me = cudaMemsetAsync(b + L.off_items, 0, (size_t)max_items(n, H, W) * 16, st);
  if (me != cudaSuccess) return cuda_fail(me, "ebos_window_prepare(items)");
  const int n_blocked = blocked_limit(n, H, W, sizeof(T), weight != nullptr);
  k_win_items<<<1, 1024, 0, st>>>(tile_off, n_tiles, tiles_x(W), k_out, n_blocked ? item_events() - kBlockEvents : item_events(),
                                  n_blocked, item_blocks(), reinterpret_cast<int4*>(b + L.off_items), hdr);
  k_win_gather<T><<<bx, 256, 0, st>>>(events, weight, n, H, W, perm, hdr, normalize_t, n_blocked,
                                      reinterpret_cast<T*>(b + L.off_x), reinterpret_cast<T*>(b + L.off_y),
                                      reinterpret_cast<T*>(b + L.off_d),
                                      weight ? reinterpret_cast<T*>(b + L.off_w) : nullptr);
  EBOS_LAUNCH_CHECK("ebos_window_prepare");
  return EBOS_OK;
}

static void apply_ablate_env() {
  static bool done = false;
  if (done) return;
  done = true;
#ifdef EBOS_ABLATION
  const int v = env_int("EBOS_ABLATE");
  if (v) cudaMemcpyToSymbol(g_ablate_word, &v, sizeof(int));
#endif
}

template <typename T>
int window_splat_t(const void* window, int64_t n, int flags, const T* flow, int H, int W, int pad_h, int pad_w, T* iwe,
                   cudaStream_t st, bool zero_iwe = true) {
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  apply_ablate_env();
  if (zero_iwe) {   // (the fused entry zeroes the accumulators and an adjacent IWE plane with ONE memset node)
    cudaError_t e = cudaMemsetAsync(iwe, 0, (size_t)Hp * Wp * sizeof(T), st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_window_splat memset");
  }
  if (n == 0) return EBOS_OK;
  const bool has_weight = flags & EBOS_WIN_HAS_WEIGHT, packed = flags & EBOS_WIN_PACKED;
  if (packed && sizeof(T) != 4) { set_error("ebos_window_splat: the packed layout exists for fp32 windows only"); return EBOS_ERR_BAD_ARG; }
  const int nb = blocked_limit(n, H, W, sizeof(T), has_weight);   // same rule as ebos_window_prepare: blocked-striped storage
  WindowLayout L = window_layout(n, sizeof(T), H, W);
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
  // (experiment knob, default off: red.v4 was measured slower than four scalar REDs in this kernel)
  static const int vec_env = env_int("EBOS_VEC_RED");
  const bool vec = sizeof(T) == 4 && vec_env && (Wp % 4 == 0) && ((reinterpret_cast<size_t>(iwe) & 15) == 0);
  // shared-memory tile kernel (fp32): opt-in with EBOS_TILE=1.  Measured on B200 at 16 Mi events it is SLOWER than
  // the one-shot kernel (130 vs 84 us): adjacent lanes flush the same cell, and shared-memory float atomics are a
  // CAS loop that serialises on same-address conflicts, whereas the L2 RED unit merges them.
  static const int tile_env = env_int("EBOS_TILE");
  if constexpr (sizeof(T) == 4) {
    // fixed-point shared-memory tile kernel (direct variant): the default for DENSE unweighted windows, mean >= 16
    // events per pixel.  Its quantisation (one rounding of <= 2^-(S+1), S = 19 for full items) is below the rounding
    // noise of sequential fp32 accumulation once a cell holds ~10 events or more, and above it for sparse windows
    // (measured: a 40-iteration Adam solve at 5 events/pixel ends 1.1e-3 px from the fp64 reference instead of
    // 0.8e-3), so sparse windows keep the fp32 REDs.  EBOS_TILE=4 forces it, 2 the run-combining variant, 3 the
    // grouped kernel.
    const bool dense_all = n >= (int64_t)16 * H * W;
    static const int v2_default = env_int("EBOS_SPLAT_V2");   // round-2 kernel as the dense default (A/B knob)
    const bool dense_v2 = dense_all && v2_default == 1;
    const bool dense = dense_all && !dense_v2;
    if (!has_weight && (tile_env == 4 || (tile_env == 0 && dense))) {
      const int4* items = reinterpret_cast<const int4*>(b + L.off_items);
      const WindowHeader* hdr = reinterpret_cast<const WindowHeader*>(b);
      static const int occ_env = env_int("EBOS_QOCC");
      // 4 CTAs/SM = 64 registers: with the prefetched next group the 48-register build spills (97 vs 75 us)
      const int occ = (occ_env == 5 || occ_env == 6) ? occ_env : 4;
      const unsigned qgrid = item_slots(n, H, W, nb);   // one CTA per item slot
      const float* fx = reinterpret_cast<const float*>(sx); const float* fy = reinterpret_cast<const float*>(sy);
      const float* fd = reinterpret_cast<const float*>(sd);
      const float* ff = reinterpret_cast<const float*>(flow); float* fi = reinterpret_cast<float*>(iwe);
#define EBOS_SD(P, B) k_tile_splat_d<P, B><<<qgrid, 256, 0, st>>>(fx, fy, fd, items, hdr, ff, H, W, pad_h, pad_w, fi, nb)
      if (packed) { if (occ == 4) EBOS_SD(true, 4); else if (occ == 5) EBOS_SD(true, 5); else EBOS_SD(true, 6); }
      else { if (occ == 4) EBOS_SD(false, 4); else if (occ == 5) EBOS_SD(false, 5); else EBOS_SD(false, 6); }
#undef EBOS_SD
      EBOS_LAUNCH_CHECK("ebos_window_splat(tile, direct)");
      return EBOS_OK;
    }
    if (!has_weight && (tile_env == 5 || tile_env == 6 || (tile_env == 0 && dense_v2))) {
      // round-2 tile kernel (per-pixel table in shared memory, window rows in use only); EBOS_TILE=6: with run merging
      const int4* items = reinterpret_cast<const int4*>(b + L.off_items);
      const WindowHeader* hdr = reinterpret_cast<const WindowHeader*>(b);
      static const int occ_env = env_int("EBOS_QOCC");
      const int occ = (occ_env == 3 || occ_env == 5 || occ_env == 6) ? occ_env : 4;
      const bool merge = tile_env == 6;
      const unsigned qgrid = item_slots(n, H, W, nb);   // one CTA per item slot
      const float* fx = reinterpret_cast<const float*>(sx); const float* fy = reinterpret_cast<const float*>(sy);
      const float* fd = reinterpret_cast<const float*>(sd);
      const float* ff = reinterpret_cast<const float*>(flow); float* fi = reinterpret_cast<float*>(iwe);
#define EBOS_SM(P, B, M) k_tile_splat_m<P, B, M><<<qgrid, 256, 0, st>>>(fx, fy, fd, items, hdr, ff, H, W, pad_h, pad_w, fi, nb)
#define EBOS_SM_O(P, M) do { if (occ == 3) EBOS_SM(P, 3, M); else if (occ == 4) EBOS_SM(P, 4, M); else if (occ == 5) EBOS_SM(P, 5, M); else EBOS_SM(P, 6, M); } while (0)
      if (packed) { if (merge) EBOS_SM_O(true, true); else EBOS_SM_O(true, false); }
      else { if (merge) EBOS_SM_O(false, true); else EBOS_SM_O(false, false); }
#undef EBOS_SM_O
#undef EBOS_SM
      EBOS_LAUNCH_CHECK("ebos_window_splat(tile, round 2)");
      return EBOS_OK;
    }
  }
  // persistent TMA-staged kernel (fp32): measured slower than the one-shot kernel, kept as an option (EBOS_PIPE=1)
  static const int pipe_env = env_int("EBOS_PIPE");
  if constexpr (sizeof(T) == 4) {
    const int64_t n_chunks = (n + kPipeChunk - 1) / kPipeChunk;
    const int slots = sm_count() * 3;
    if (pipe_env == 1) {
      const unsigned pgrid = (unsigned)std::min<int64_t>(n_chunks, slots);
      const float* fx = reinterpret_cast<const float*>(sx); const float* fy = reinterpret_cast<const float*>(sy);
      const float* fd = reinterpret_cast<const float*>(sd); const float* fw = reinterpret_cast<const float*>(sw);
      const float* ff = reinterpret_cast<const float*>(flow); float* fi = reinterpret_cast<float*>(iwe);
#define EBOS_SPLAT_P(WGT, V, P)                                                                                          \
  do {                                                                                                                   \
    auto kern = k_win_splat_pipe<WGT, V, P>;                                                                             \
    const size_t smem = PipeCfg<WGT, P>::smem_bytes;                                                                     \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                  \
    kern<<<pgrid, 256, smem, st>>>(fx, fy, fd, fw, n, ff, H, W, pad_h, pad_w, fi);                                        \
  } while (0)
      if (has_weight) { if (packed) { if (vec) EBOS_SPLAT_P(true, true, true); else EBOS_SPLAT_P(true, false, true); }
                        else { if (vec) EBOS_SPLAT_P(true, true, false); else EBOS_SPLAT_P(true, false, false); } }
      else { if (packed) { if (vec) EBOS_SPLAT_P(false, true, true); else EBOS_SPLAT_P(false, false, true); }
             else { if (vec) EBOS_SPLAT_P(false, true, false); else EBOS_SPLAT_P(false, false, false); } }
#undef EBOS_SPLAT_P
      EBOS_LAUNCH_CHECK("ebos_window_splat(pipe)");
      return EBOS_OK;
    }
  }
  static const int ng_env = env_int("EBOS_GROUPS");   // 0 default (4 groups of 4 events), 2/4/8 groups, -1 legacy kernels
  if constexpr (sizeof(T) == 4) {
    if (ng_env >= 0) {
      // events per thread = 4 * ng: 16 by default; small windows take fewer so that the grid still holds >= 2 CTAs per
      // SM (a 500 k-event window at 16 events per thread is 122 CTAs on 148 SMs: latency-bound, 16.6 us)
      int ng = (ng_env == 1 || ng_env == 2 || ng_env == 8) ? ng_env : 4;
      if (ng_env == 0) {
        const int64_t want = (int64_t)2 * sm_count() * 256;   // threads
        while (ng > 1 && n / (4 * ng) < want) ng >>= 1;
      }
      const unsigned ggrid = (unsigned)((((n + 4 * ng - 1) / (4 * ng)) + 255) / 256);
      const float* fx = reinterpret_cast<const float*>(sx); const float* fy = reinterpret_cast<const float*>(sy);
      const float* fd = reinterpret_cast<const float*>(sd); const float* fw = reinterpret_cast<const float*>(sw);
      const float* ff = reinterpret_cast<const float*>(flow); float* fi = reinterpret_cast<float*>(iwe);
#define EBOS_SG(WGT, P, NGV) k_win_splat_g<WGT, P, NGV><<<ggrid, 256, 0, st>>>(fx, fy, fd, fw, n, ff, H, W, pad_h, pad_w, fi)
#define EBOS_SG_N(WGT, P) do { if (ng == 1) EBOS_SG(WGT, P, 1); else if (ng == 2) EBOS_SG(WGT, P, 2); else if (ng == 8) EBOS_SG(WGT, P, 8); else EBOS_SG(WGT, P, 4); } while (0)
      if (has_weight) { if (packed) EBOS_SG_N(true, true); else EBOS_SG_N(true, false); }
      else { if (packed) EBOS_SG_N(false, true); else EBOS_SG_N(false, false); }
#undef EBOS_SG_N
#undef EBOS_SG
      EBOS_LAUNCH_CHECK("ebos_window_splat(grouped)");
      return EBOS_OK;
    }
  }
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
  const int nb = blocked_limit(n, H, W, sizeof(T), has_weight);
  WindowLayout L = window_layout(n, sizeof(T), H, W);
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
  // (experiment knob, default off: float4 pair loads were measured slower -- L1 moves 4x the bytes)
  static const int vec_env = env_int("EBOS_VEC_LOAD");
  const int Wp = W + 2 * pad_w;
  const bool vec = sizeof(T) == 4 && vec_env && (Wp % 4 == 0) && ((reinterpret_cast<size_t>(gsrc) & 15) == 0);
  static const int tile_env = env_int("EBOS_TILE");
  if constexpr (sizeof(T) == 4) {
    // second tile backward (round 2): dL/dIWE window + per-pixel table in shared memory, one CTA per work item
    static const int tbwd_env = env_int("EBOS_TILE_BWD");   // 1 force, 2 default for dense windows (A/B knob)
    const bool dense = n >= (int64_t)16 * H * W;
    if (!has_weight && (tbwd_env == 1 || (tbwd_env == 2 && dense && tile_env == 0))) {
      const int4* items = reinterpret_cast<const int4*>(b + L.off_items);
      const WindowHeader* hdr = reinterpret_cast<const WindowHeader*>(b);
      static const int occ_env = env_int("EBOS_BOCC");
      const int occ = (occ_env == 3 || occ_env == 5 || occ_env == 6) ? occ_env : 4;
      const unsigned qgrid = item_slots(n, H, W, nb);
      const float* fx = reinterpret_cast<const float*>(sx); const float* fy = reinterpret_cast<const float*>(sy);
      const float* fd = reinterpret_cast<const float*>(sd);
      const float* ff = reinterpret_cast<const float*>(flow); const float* fg = reinterpret_cast<const float*>(gsrc);
      float* fo = reinterpret_cast<float*>(dflow);
      cudaError_t le;
#define EBOS_TBM(G, P, B) le = launch_pdl(k_tile_bwd_m<G, P, B>, dim3(qgrid), dim3(256), st, fx, fy, fd, items, hdr, ff, H, W, pad_h, pad_w, fg, acc, omit_boundary, scale, fo, nb)
#define EBOS_TBM_O(G, P) do { if (occ == 3) EBOS_TBM(G, P, 3); else if (occ == 4) EBOS_TBM(G, P, 4); else if (occ == 5) EBOS_TBM(G, P, 5); else EBOS_TBM(G, P, 6); } while (0)
      if (affine) { if (packed) EBOS_TBM_O(1, true); else EBOS_TBM_O(1, false); }
      else { if (packed) EBOS_TBM_O(0, true); else EBOS_TBM_O(0, false); }
#undef EBOS_TBM_O
#undef EBOS_TBM
      if (le != cudaSuccess) return cuda_fail(le, "ebos_window_backward(tile, round 2)");
      return EBOS_OK;
    }
  }
  static const int pipe_env = env_int("EBOS_PIPE");
  if constexpr (sizeof(T) == 4) {
    const int64_t n_chunks = (n + kPipeChunk - 1) / kPipeChunk;
    const int slots = sm_count() * 3;
    if (pipe_env == 1) {
      const unsigned pgrid = (unsigned)std::min<int64_t>(n_chunks, slots);
      const float* fx = reinterpret_cast<const float*>(sx); const float* fy = reinterpret_cast<const float*>(sy);
      const float* fd = reinterpret_cast<const float*>(sd); const float* fw = reinterpret_cast<const float*>(sw);
      const float* ff = reinterpret_cast<const float*>(flow); const float* fg = reinterpret_cast<const float*>(gsrc);
      float* fo = reinterpret_cast<float*>(dflow);
#define EBOS_BWD_P(G, WGT, V, P)                                                                                         \
  do {                                                                                                                   \
    auto kern = k_win_bwd_pipe<G, WGT, V, P>;                                                                            \
    const size_t smem = PipeCfg<WGT, P>::smem_bytes;                                                                     \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                  \
    kern<<<pgrid, 256, smem, st>>>(fx, fy, fd, fw, n, ff, H, W, pad_h, pad_w, fg, acc, omit_boundary, scale, fo);         \
  } while (0)
#define EBOS_BWD_PG(G)                                                                                                   \
  do {                                                                                                                   \
    if (has_weight) { if (packed) { if (vec) EBOS_BWD_P(G, true, true, true); else EBOS_BWD_P(G, true, false, true); }    \
                      else { if (vec) EBOS_BWD_P(G, true, true, false); else EBOS_BWD_P(G, true, false, false); } }       \
    else { if (packed) { if (vec) EBOS_BWD_P(G, false, true, true); else EBOS_BWD_P(G, false, false, true); }             \
           else { if (vec) EBOS_BWD_P(G, false, true, false); else EBOS_BWD_P(G, false, false, false); } }                \
  } while (0)
      if (affine) EBOS_BWD_PG(1); else EBOS_BWD_PG(0);
#undef EBOS_BWD_PG
#undef EBOS_BWD_P
      EBOS_LAUNCH_CHECK("ebos_window_backward(pipe)");
      return EBOS_OK;
    }
  }
  static const int ng_env = env_int("EBOS_GROUPS");   // 0 default (2 groups of 4 events: measured best), 4/8, -1 legacy
  if constexpr (sizeof(T) == 4) {
    if (ng_env >= 0) {
      int ng = (ng_env == 1 || ng_env == 4 || ng_env == 8) ? ng_env : 2;
      if (ng_env == 0 && n / (4 * ng) < (int64_t)2 * sm_count() * 256) ng = 1;   // small windows: see window_splat_t
      if (nb > 0 && ng == 8) ng = 4;                                              // blocked-striped windows: at most one lane slot per thread
      const unsigned ggrid = (unsigned)((((n + 4 * ng - 1) / (4 * ng)) + 255) / 256);
      const float* fx = reinterpret_cast<const float*>(sx); const float* fy = reinterpret_cast<const float*>(sy);
      const float* fd = reinterpret_cast<const float*>(sd); const float* fw = reinterpret_cast<const float*>(sw);
      const float* ff = reinterpret_cast<const float*>(flow); const float* fg = reinterpret_cast<const float*>(gsrc);
      float* fo = reinterpret_cast<float*>(dflow);
#define EBOS_BG(G, WGT, P, NGV) launch_pdl(k_win_bwd_g<G, WGT, P, NGV>, dim3(ggrid), dim3(256), st, fx, fy, fd, fw, n, ff, H, W, pad_h, pad_w, fg, acc, omit_boundary, scale, fo, nb)
#define EBOS_BGO(G, WGT, P, NGV, O) k_win_bwd_g<G, WGT, P, NGV, O><<<ggrid, 256, 0, st>>>(fx, fy, fd, fw, n, ff, H, W, pad_h, pad_w, fg, acc, omit_boundary, scale, fo, nb)
      static const int bocc = env_int("EBOS_BOCC");   // experiment knob (unweighted packed gradient-plane kernel only)
      if (bocc && !affine && !has_weight && packed && (ng == 2 || ng == 4)) {
        if (bocc == 5 && ng == 2) EBOS_BGO(0, false, true, 2, 5); else if (bocc == 5) EBOS_BGO(0, false, true, 4, 5);
        else if (bocc == 3 && ng == 2) EBOS_BGO(0, false, true, 2, 3); else if (bocc == 3) EBOS_BGO(0, false, true, 4, 3);
        else if (bocc == 6 && ng == 2) EBOS_BGO(0, false, true, 2, 6); else EBOS_BGO(0, false, true, 4, 6);
        EBOS_LAUNCH_CHECK("ebos_window_backward(grouped, occ)");
        return EBOS_OK;
      }
#define EBOS_BG_N(G, WGT, P) do { if (ng == 1) EBOS_BG(G, WGT, P, 1); else if (ng == 2) EBOS_BG(G, WGT, P, 2); else if (ng == 8) EBOS_BG(G, WGT, P, 8); else EBOS_BG(G, WGT, P, 4); } while (0)
#define EBOS_BG_W(G) do { if (has_weight) { if (packed) EBOS_BG_N(G, true, true); else EBOS_BG_N(G, true, false); } \
                          else { if (packed) EBOS_BG_N(G, false, true); else EBOS_BG_N(G, false, false); } } while (0)
      if (affine) EBOS_BG_W(1); else EBOS_BG_W(0);
#undef EBOS_BG_W
#undef EBOS_BG_N
#undef EBOS_BG
      EBOS_LAUNCH_CHECK("ebos_window_backward(grouped)");
      return EBOS_OK;
    }
  }
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
                        int pad_w, int dtype, void* iwe, cudaStream_t st, bool zero_iwe) {
  if (dtype == EBOS_F64)
    return window_splat_t<double>(window, n, flags, (const double*)flow, H, W, pad_h, pad_w, (double*)iwe, st, zero_iwe);
  return window_splat_t<float>(window, n, flags, (const float*)flow, H, W, pad_h, pad_w, (float*)iwe, st, zero_iwe);
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

size_t ebos_window_bytes(int64_t n, int H, int W, int dtype) {
  return (n < 0 || H <= 0 || W <= 0) ? 0 : window_layout(n, dtype_size(dtype), H, W).total;
}

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

int ebos_window_info(const void* window, int64_t n, int H, int W, int dtype, int32_t* perm_out, double* tinfo_out,
                     void* stream) {
  EBOS_REQUIRE(window && n >= 0, "ebos_window_info: bad argument");
  EBOS_CHECK_DTYPE(dtype, "ebos_window_info");
  cudaStream_t st = as_stream(stream);
  const char* b = reinterpret_cast<const char*>(window);
  if (perm_out && n > 0) {
    cudaError_t e = cudaMemcpyAsync(perm_out, b + window_layout(n, dtype_size(dtype), H, W).off_perm, (size_t)n * 4,
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
  return window_splat_launch(window, n, flags, flow, H, W, pad_h, pad_w, dtype, iwe, as_stream(stream), true);
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
