// Objectives on the IWE / flow planes, Adam, and the fused per-iteration entry point (fp32 / fp64).
//
//   variance            L = -var(IWE) (unbiased)                    SURVEY.md A.4 (not upstream)
//   gradient magnitude  L = -mean((Sx/8)^2 + (Sy/8)^2), Sobel 3x3 with replicate padding, kernels of
//                       src/utils/stat_utils.py:62-139                SURVEY.md A.4 (not upstream)
//   total variation     ImageGradient.calculate_torch, src/costs/image_gradient.py:60-75
//   Adam                torch.optim.Adam as used at src/solver/patch_eklt_pyramid2.py:262-264,284
//
// All reductions: per-thread double accumulators -> warp shuffle -> one atomicAdd(double) per CTA.
// acc layout (double[EBOS_ACC_DOUBLES = 40]): [0] sum(IWE) [1] sum(IWE^2) [2] + [8..23] sum(gx^2+gy^2)
// [3] + [24..39] sum(|TV terms|); [4] ticket counter of the fused Adam + TV kernel; [5], [6] the Adam bias-correction
// factors of the current iteration (written by the TV kernel, NOT reset).  Same-address atomics from different CTAs serialise in the L2 atomic unit
// (~6 ns each on B200: 3.5 us for the ~600 CTAs of the TV kernel, as long as its real work), so the two plane
// kernels of the hot path spread their per-CTA partial sums over 16 slots; the consumers add the slots up.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "ebos_common.cuh"

namespace ebos {

constexpr int kAccSpread = 16, kAccGradSlots = 8, kAccTvSlots = 24;
constexpr int64_t kTvAfterSplatEvents = (int64_t)1 << 22;   // windows from 4 Mi events: TV kernel next to the cost kernel, not the splat
static_assert(kAccTvSlots + kAccSpread == EBOS_ACC_DOUBLES, "accumulator layout");

// The TV term depends only on the flow, the splat only on events + flow: inside the fused evaluation
// they run concurrently (fork/join on a cached auxiliary stream; the pattern is CUDA-graph capturable).
// One non-blocking stream and two events per device are created on first use and kept for the process.
struct AuxLane {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, cost_done = nullptr, fin_done = nullptr;
  bool ok = false;
};
// One lane per (device, caller stream), created on first use and kept for the process: solves that run concurrently on
// different streams (ContrastMaximizationDense.estimate_many in eager mode, several host threads) must not share
// the lane's stream and events -- a shared lane adds false cross-window dependencies, and a second thread re-recording
// `fork` between the first one's record and wait would let its TV kernel read a flow that is still being updated.
static AuxLane* aux_lane(cudaStream_t caller) {
  static std::map<std::pair<int, cudaStream_t>, AuxLane> lanes;
  static std::mutex mu;
  static const bool disabled = getenv("EBOS_NO_OVERLAP") != nullptr;
  if (disabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (lanes.size() > 4096) return nullptr;   // (streams come and go in a long-lived process: fall back to no overlap)
  AuxLane& L = lanes[std::make_pair(dev, caller)];
  if (!L.ok) {
    if (cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&L.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&L.cost_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&L.fin_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    L.ok = true;
  }
  return &L;
}

int window_splat_launch(const void* window, int64_t n, int flags, const void* flow, int H, int W, int pad_h,
                        int pad_w, int dtype, void* iwe, cudaStream_t st, bool zero_iwe);
int blur3_launch(const void* in, int H, int W, double sigma, int adjoint, int dtype, void* out, cudaStream_t st);
int window_backward_launch(const void* window, int64_t n, int flags, const void* flow, int H, int W, int pad_h,
                           int pad_w, int dtype, const void* grad_iwe, int kind, const void* iwe, const double* acc,
                           int omit_boundary, double scale, void* dflow, cudaStream_t st);

// ---- variance: one pass, sum and sum of squares in double --------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_var_reduce(const T* __restrict__ iwe, int Hp, int Wp, int omit,
                                                    double* __restrict__ acc) {
  __shared__ double sm[32];
  const int r_lo = omit ? 1 : 0, r_hi = omit ? Hp - 1 : Hp;
  const int c_lo = omit ? 1 : 0, c_hi = omit ? Wp - 1 : Wp;
  double s = 0.0, q = 0.0;
  for (int r = r_lo + blockIdx.y; r < r_hi; r += gridDim.y) {
    const T* row = iwe + (int64_t)r * Wp;
    for (int c = c_lo + blockIdx.x * blockDim.x + threadIdx.x; c < c_hi; c += gridDim.x * blockDim.x) {
      const double v = (double)__ldg(row + c);
      s += v;
      q += v * v;
    }
  }
  s = block_sum(s, sm);
  q = block_sum(q, sm);
  if (threadIdx.x == 0) {
    atomicAdd(acc + 0, s);
    atomicAdd(acc + 1, q);
  }
}

// dL/dIWE for the variance objective as an explicit plane (only needed by the operator-level
// cost classes; the fused path derives it on the fly inside the backward kernel).
template <typename T>
__global__ void __launch_bounds__(256) k_var_grad(const T* __restrict__ iwe, int Hp, int Wp, int omit, double scale,
                                                  const double* __restrict__ acc, T* __restrict__ g) {
  const double cnt = omit ? (double)(Hp - 2) * (double)(Wp - 2) : (double)Hp * (double)Wp;
  const T mean = (T)(acc[0] / cnt);
  const T cv = (T)(-2.0 * scale / (cnt - 1.0));
  for (int r = blockIdx.y; r < Hp; r += gridDim.y) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < Wp; c += gridDim.x * blockDim.x) {
      const bool border = r == 0 || c == 0 || r == Hp - 1 || c == Wp - 1;
      const int64_t i = (int64_t)r * Wp + c;
      g[i] = (omit && border) ? (T)0 : cv * (__ldg(iwe + i) - mean);
    }
  }
}

// ---- gradient magnitude: Sobel forward + adjoint in one tiled pass ---------------------------------
constexpr int GTH = 16, GTW = 64;  // output tile (rows x cols); 256 threads, 4 pixels each

// Adjoint at a pixel on the image border: the replicate padding folds the padded positions that clamp
// to p back onto it.  Rare (perimeter only), kept out of line so that the interior path stays lean.
template <typename T>
__device__ __noinline__ T gradmag_border_adjoint(const T* __restrict__ sDx, const T* __restrict__ sDy, int pitch, int r,
                                                 int c, int r0, int c0, int Hp, int Wp) {
  T out = 0;
  for (int pr = (r == 0 ? -1 : r); pr <= (r == Hp - 1 ? Hp : r); ++pr) {
    for (int pc = (c == 0 ? -1 : c); pc <= (c == Wp - 1 ? Wp : c); ++pc) {
      for (int u = -1; u <= 1; ++u) {
        for (int v = -1; v <= 1; ++v) {
          const int qr = pr - u, qc = pc - v;
          if (qr < 0 || qr >= Hp || qc < 0 || qc >= Wp) continue;
          const int sr = qr - (r0 - 1), sc = qc - (c0 - 1);
          if (sr < 0 || sr >= GTH + 2 || sc < 0 || sc >= GTW + 2) continue;
          const T kx = (T)(u * (2 - (v < 0 ? -v : v)));
          const T ky = (T)(v * (2 - (u < 0 ? -u : u)));
          out += kx * sDx[sr * pitch + sc] + ky * sDy[sr * pitch + sc];
        }
      }
    }
  }
  return out;
}

template <typename T>
__global__ void __launch_bounds__(256, 4) k_gradmag(const T* __restrict__ iwe, int Hp, int Wp, int omit, T coef,
                                                 double* __restrict__ acc, T* __restrict__ g) {
  // coef = -2 * scale / (8 * n_el): D = coef * gx is dL/d(Sx) including the forward's 1/8.
  __shared__ T sI[GTH + 4][GTW + 4 + 1];
  __shared__ T sDx[GTH + 2][GTW + 2 + 1];
  __shared__ T sDy[GTH + 2][GTW + 2 + 1];
  __shared__ double sm[32];
  const int tiles_x = (Wp + GTW - 1) / GTW, tiles_y = (Hp + GTH - 1) / GTH;
  double part = 0.0;
  for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
  const int r0 = (tile / tiles_x) * GTH, c0 = (tile % tiles_x) * GTW;
  __syncthreads();  // the previous tile's stage 3 is done with the shared planes
  // stage 1: image tile with a 2-pixel halo, replicate padding materialised by clamping
  for (int i = threadIdx.x; i < (GTH + 4) * (GTW + 4); i += blockDim.x) {
    const int lr = i / (GTW + 4), lc = i - lr * (GTW + 4);
    const int r = min(max(r0 - 2 + lr, 0), Hp - 1), c = min(max(c0 - 2 + lc, 0), Wp - 1);
    sI[lr][lc] = __ldg(iwe + (int64_t)r * Wp + c);
  }
  __syncthreads();
  // stage 2: Sobel/8 on the tile + 1-pixel halo; positions outside the image (or outside the
  // omit_boundary crop) carry no gradient.
  for (int i = threadIdx.x; i < (GTH + 2) * (GTW + 2); i += blockDim.x) {
    const int lr = i / (GTW + 2), lc = i - lr * (GTW + 2);
    const int r = r0 - 1 + lr, c = c0 - 1 + lc;
    T dx = 0, dy = 0;
    const bool inside = r >= 0 && r < Hp && c >= 0 && c < Wp;
    const bool counted = inside && !(omit && (r == 0 || c == 0 || r == Hp - 1 || c == Wp - 1));
    if (counted) {
      // sI index of pixel (r,c) is [lr+1][lc+1]
      const T i00 = sI[lr][lc], i01 = sI[lr][lc + 1], i02 = sI[lr][lc + 2];
      const T i10 = sI[lr + 1][lc], i12 = sI[lr + 1][lc + 2];
      const T i20 = sI[lr + 2][lc], i21 = sI[lr + 2][lc + 1], i22 = sI[lr + 2][lc + 2];
      const T gx = ((i20 - i00) + (T)2 * (i21 - i01) + (i22 - i02)) * (T)0.125;  // d/drow
      const T gy = ((i02 - i00) + (T)2 * (i12 - i10) + (i22 - i20)) * (T)0.125;  // d/dcol
      dx = coef * gx;
      dy = coef * gy;
      if (lr >= 1 && lr <= GTH && lc >= 1 && lc <= GTW) part += (double)gx * gx + (double)gy * gy;
    }
    sDx[lr][lc] = dx;
    sDy[lr][lc] = dy;
  }
  __syncthreads();
  // stage 3: adjoint.  dI[p] = sum over padded positions pp that clamp to p, over the 3x3 window:
  //   Kx[u][v]*Dx[pp-(u,v)] + Ky[u][v]*Dy[pp-(u,v)],   with D = 0 outside the image (already zero in
  //   smem), Kx[u][v] = u*(2-|v|), Ky[u][v] = v*(2-|u|).
  for (int i = threadIdx.x; i < GTH * GTW; i += blockDim.x) {
    const int lr = i / GTW, lc = i - lr * GTW;
    const int r = r0 + lr, c = c0 + lc;
    if (r >= Hp || c >= Wp) continue;
    T out = 0;
    if (r > 0 && r < Hp - 1 && c > 0 && c < Wp - 1) {
      // interior pixel: only pp = p; q = p - (u,v) is inside the image; s index of p is [lr+1][lc+1]
      const int sr = lr + 1, sc = lc + 1;
      // u = -1 -> q row r+1 ; u = +1 -> q row r-1
      out = -(sDx[sr + 1][sc + 1] + (T)2 * sDx[sr + 1][sc] + sDx[sr + 1][sc - 1])
            + (sDx[sr - 1][sc + 1] + (T)2 * sDx[sr - 1][sc] + sDx[sr - 1][sc - 1])
            - (sDy[sr + 1][sc + 1] + (T)2 * sDy[sr][sc + 1] + sDy[sr - 1][sc + 1])
            + (sDy[sr + 1][sc - 1] + (T)2 * sDy[sr][sc - 1] + sDy[sr - 1][sc - 1]);
    } else {
      out = gradmag_border_adjoint<T>(&sDx[0][0], &sDy[0][0], GTW + 2 + 1, r, c, r0, c0, Hp, Wp);
    }
    g[(int64_t)r * Wp + c] = out;
  }
  }  // tile loop
  part = block_sum(part, sm);
  if (threadIdx.x == 0) atomicAdd(acc + 2, part);
}

// ---- gradient magnitude, separable version (default) -------------------------------------------------
// The tiled kernel above spends ~220 instructions per pixel (ncu r01c: issue-bound at 15 us for 7.4 MB).  Away
// from the border the whole forward + adjoint collapses into two separable stencils.  With s = [1,2,1] (smooth),
// d = [-1,0,1] (difference), Sx = d_row (x) s_col, Sy = s_row (x) d_col:
//   gx = (S[r+1] - S[r-1]) / 8,  gy = (D[r-1] + 2 D[r] + D[r+1]) / 8      S = s_col * I,  D = d_col * I
//   dL/dI = coef/8 * (Sx^T Sx + Sy^T Sy) I = coef/8 * (a_row (x) b_col + b_row (x) a_col) I
//   a = d (*) d = [-1,0,2,0,-1],  b = s (*) s = [1,4,6,4,1];   P = b_col * I,  Q = a_col * I
// A thread owns one column of a 32-row x 64-column tile band and marches down the rows with the last five rows of
// (P, Q, S, D) in registers: 5 conflict-free LDS + ~35 flops per pixel.  This "fast region" is r in [3, Hp-4],
// c in [3, Wp-4]: there no tap is clamped by the replicate padding and no g on the (optionally omitted) border
// ring is involved.  The 3-pixel frame around it is evaluated exactly, straight from the definition, by extra
// CTAs of the same launch (one thread per frame pixel).
constexpr int GM_TW = 64, GM_ROWS = 8;               // tile width, rows per thread
#ifndef EBOS_GM_GROUPS
#define EBOS_GM_GROUPS 4                              // row groups per CTA (CTA = 64 columns x groups threads)
#endif
constexpr int GM_GROUPS = EBOS_GM_GROUPS, GM_TH = GM_GROUPS * GM_ROWS, GM_THREADS = GM_TW * GM_GROUPS;
constexpr int GM_LW = GM_TW + 8;                      // shared tile row: columns c0 - 4 .. c0 + 67 (18 aligned float4)
constexpr int GM_STRIP = 64;                          // pixels along a frame strip (3 deep): one CTA each
static_assert((GM_TH + 4) * GM_LW >= 7 * (GM_STRIP + 4) + 2 * 5 * (GM_STRIP + 2), "frame strips reuse the tile's shared memory");

// exact value at one frame pixel p = (r, c): sum over the counted positions q in the 3x3 neighbourhood of p and
// the Sobel taps (u, v) whose clamped target clamp(q + (u, v)) is p.  The 5x5 clamped neighbourhood of p is
// loaded up front (25 independent loads, one round trip); everything else is register arithmetic with
// compile-time indices (a first version with nested loops around dependent loads made these few threads the
// critical path of the whole launch).
// The IWE as the sum of up to kMaxPeers partial planes (event-sharded windows: every rank's partial IWE lives in
// symmetric memory, mapped into this process over NVLink).  The partial planes are summed IN RANK ORDER while the
// tiles are loaded, so the all-reduce of the partial IWEs is fused into the cost kernel and every rank computes
// bit-identical results.  n == 1 is the ordinary single-plane case.
constexpr int kMaxPeers = 8;
template <typename T> struct PeerPlanes {
  const T* p[kMaxPeers];
  int n;
  __device__ __forceinline__ T load(int64_t idx) const {
    T v = p[0][idx];
    // (not unrolled: the compiler otherwise expands the seven optional peers at every call site -- 5000 of the 7160 SASS
    //  instructions of k_gradmag_sep were this loop, 25 times over in the frame-pixel code, and the kernel stalled on
    //  instruction fetches; with one plane the loop body never runs)
#pragma unroll 1
    for (int r = 1; r < n; ++r) v += p[r][idx];
    return v;
  }
};

template <typename T>
__device__ void gradmag_frame_pixel(const PeerPlanes<T>& iwe, int Hp, int Wp, int omit, T coef, int r, int c,
                                    T* __restrict__ g, double& part) {
  T v[5][5];
#pragma unroll
  for (int a = -2; a <= 2; ++a)
#pragma unroll
    for (int b = -2; b <= 2; ++b)
      v[a + 2][b + 2] = iwe.load((int64_t)min(max(r + a, 0), Hp - 1) * Wp + min(max(c + b, 0), Wp - 1));
  T out = 0;
  // (fully unrolled on purpose: with the outer loop rolled the 5 x 5 neighbourhood moves to local memory and these few
  //  threads become the critical path of the launch -- measured, r02t: 12.1 instead of 10.5 us)
#pragma unroll
  for (int dr = -1; dr <= 1; ++dr) {
#pragma unroll
    for (int dc = -1; dc <= 1; ++dc) {
      const int qr = r + dr, qc = c + dc;
      if (qr < 0 || qr >= Hp || qc < 0 || qc >= Wp) continue;
      if (omit && (qr == 0 || qc == 0 || qr == Hp - 1 || qc == Wp - 1)) continue;
      // Sobel/8 at q from the clamped neighbourhood: I(clamp(q + (u,v))) = v[dr + u + 2][dc + v + 2]
      const T gx = ((v[dr + 3][dc + 1] - v[dr + 1][dc + 1]) + (T)2 * (v[dr + 3][dc + 2] - v[dr + 1][dc + 2]) +
                    (v[dr + 3][dc + 3] - v[dr + 1][dc + 3])) * (T)0.125;  // d/drow
      const T gy = ((v[dr + 1][dc + 3] - v[dr + 1][dc + 1]) + (T)2 * (v[dr + 2][dc + 3] - v[dr + 2][dc + 1]) +
                    (v[dr + 3][dc + 3] - v[dr + 3][dc + 1])) * (T)0.125;  // d/dcol
      if (dr == 0 && dc == 0) part += (double)gx * gx + (double)gy * gy;
#pragma unroll
      for (int u = -1; u <= 1; ++u) {
#pragma unroll
        for (int w = -1; w <= 1; ++w) {
          if (min(max(qr + u, 0), Hp - 1) != r || min(max(qc + w, 0), Wp - 1) != c) continue;
          out += (T)(u * (2 - (w < 0 ? -w : w))) * (coef * gx) + (T)(w * (2 - (u < 0 ? -u : u))) * (coef * gy);
        }
      }
    }
  }
  g[(int64_t)r * Wp + c] = out;
}

// The frame, one STRIP per CTA (3 x <= 64 pixels of the top / bottom band, <= 64 x 3 of the left / right band).  With one
// thread per frame pixel running the function above, the 374 frame warps of a 1280 x 720 launch lived eight times as long
// as the 3680 tile warps (ncu r02s: 9 % of the warps, 36 % of the stall samples) and set the duration of the whole kernel:
// every thread re-derived the Sobel pair of its nine neighbours, ~2000 dependent instructions.  Here the strip's clamped
// image tile goes to shared memory, phase 1 evaluates coef * (gx, gy) ONCE per position q of the strip dilated by one
// (zero where q is outside the image or on the omitted ring, so phase 2 needs no validity tests) and sums the loss terms of
// the positions the strip owns, phase 2 gathers the adjoint for the strip's pixels.  Same expressions and the same
// summation order per pixel as gradmag_frame_pixel.
template <typename T>
__device__ void gradmag_frame_strip(const PeerPlanes<T>& iwe, int Hp, int Wp, int omit, T coef, int r0, int nr, int c0,
                                    int nc, T* __restrict__ g, double& part, T* __restrict__ sm) {
  const int tw = nc + 4, th = nr + 4;   // clamped image tile, origin (r0 - 2, c0 - 2)
  const int gw = nc + 2, gh = nr + 2;   // Sobel pairs, origin (r0 - 1, c0 - 1)
  T* sImg = sm;
  T* sGx = sm + th * tw;
  T* sGy = sGx + gh * gw;
  for (int i = threadIdx.x; i < th * tw; i += blockDim.x) {
    const int lr = i / tw, lc = i - lr * tw;
    sImg[i] = iwe.load((int64_t)min(max(r0 - 2 + lr, 0), Hp - 1) * Wp + min(max(c0 - 2 + lc, 0), Wp - 1));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < gh * gw; i += blockDim.x) {
    const int lr = i / gw, lc = i - lr * gw;
    const int qr = r0 - 1 + lr, qc = c0 - 1 + lc;
    T cgx = 0, cgy = 0;
    const bool counted = qr >= 0 && qr < Hp && qc >= 0 && qc < Wp &&
                         !(omit && (qr == 0 || qc == 0 || qr == Hp - 1 || qc == Wp - 1));
    if (counted) {
      const T* a = sImg + lr * tw + lc;   // a[(u + 1) * tw + (w + 1)] = I(clamp(q + (u, w)))
      const T gx = ((a[2 * tw] - a[0]) + (T)2 * (a[2 * tw + 1] - a[1]) + (a[2 * tw + 2] - a[2])) * (T)0.125;          // d/drow
      const T gy = ((a[2] - a[0]) + (T)2 * (a[tw + 2] - a[tw]) + (a[2 * tw + 2] - a[2 * tw])) * (T)0.125;              // d/dcol
      if (lr >= 1 && lr <= nr && lc >= 1 && lc <= nc) part += (double)gx * gx + (double)gy * gy;
      cgx = coef * gx;
      cgy = coef * gy;
    }
    sGx[i] = cgx;
    sGy[i] = cgy;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nr * nc; i += blockDim.x) {
    const int pl = i / nc, pc = i - pl * nc;
    const int r = r0 + pl, c = c0 + pc;
    T out = 0;
#pragma unroll
    for (int dr = -1; dr <= 1; ++dr) {
#pragma unroll
      for (int dc = -1; dc <= 1; ++dc) {
        const int qr = r + dr, qc = c + dc;
        const T cgx = sGx[(pl + 1 + dr) * gw + pc + 1 + dc], cgy = sGy[(pl + 1 + dr) * gw + pc + 1 + dc];
#pragma unroll
        for (int u = -1; u <= 1; ++u) {
#pragma unroll
          for (int w = -1; w <= 1; ++w) {
            if (min(max(qr + u, 0), Hp - 1) != r || min(max(qc + w, 0), Wp - 1) != c) continue;
            out += (T)(u * (2 - (w < 0 ? -w : w))) * cgx + (T)(w * (2 - (u < 0 ? -u : u))) * cgy;
          }
        }
      }
    }
    g[(int64_t)r * Wp + c] = out;
  }
}

template <typename T>
__global__ void __launch_bounds__(GM_THREADS)
k_gradmag_sep(const PeerPlanes<T> iwe, int Hp, int Wp, int omit, T coef, double* __restrict__ acc, T* __restrict__ g,
              int n_frame_ctas, int n_tiles) {
  // image tile with a 2-pixel halo, stored from column c0 - 4 (the 16-byte aligned start of the vector loads below):
  // image column c sits at sI[.][c - c0 + 4]
  __shared__ __align__(16) T sI[GM_TH + 4][GM_LW];
  __shared__ double sm[32];
  pdl_launch_dependents();   // the backward may be scheduled while this grid drains (it waits before reading g)
  double part = 0.0;
  const bool has_fast = Hp >= 7 && Wp >= 7;
  if ((int)blockIdx.x < n_frame_ctas) {
    pdl_wait();   // the IWE (and the zeroed accumulators) of the preceding splat
    if (!has_fast) {
      // images without a fast region are all frame: one pixel per thread
      const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (idx < (int64_t)Hp * Wp) {
        const int r = (int)(idx / Wp), c = (int)(idx - (int64_t)r * Wp);
        gradmag_frame_pixel<T>(iwe, Hp, Wp, omit, coef, r, c, g, part);
      }
    } else {
      // frame strips: rows 0..2 and Hp-3..Hp-1 in 64-column segments, then columns 0..2 and Wp-3..Wp-1 of the other
      // rows in 64-row segments (same order as the host's count: 2 * nsx + 2 * nsy)
      const int nsx = (Wp + GM_STRIP - 1) / GM_STRIP, nsy = (Hp - 6 + GM_STRIP - 1) / GM_STRIP;
      int sidx = blockIdx.x, r0, nr, c0, nc;
      if (sidx < 2 * nsx) {
        r0 = sidx < nsx ? 0 : Hp - 3; nr = 3;
        c0 = (sidx % nsx) * GM_STRIP; nc = min(GM_STRIP, Wp - c0);
      } else {
        sidx -= 2 * nsx;
        c0 = sidx < nsy ? 0 : Wp - 3; nc = 3;
        r0 = 3 + (sidx % nsy) * GM_STRIP; nr = min(GM_STRIP, Hp - 3 - r0);
      }
      gradmag_frame_strip<T>(iwe, Hp, Wp, omit, coef, r0, nr, c0, nc, g, part, &sI[0][0]);
    }
  } else if ((int)blockIdx.x - n_frame_ctas < n_tiles) {
    const int tile = blockIdx.x - n_frame_ctas;
    const int tiles_x = (Wp + GM_TW - 1) / GM_TW;
    const int r0 = (tile / tiles_x) * GM_TH, c0 = (tile % tiles_x) * GM_TW;
    pdl_wait();   // the IWE (and the zeroed accumulators) of the preceding splat
    // tile load (clamped reads: values outside the image are never used by the fast region).  The first version
    // walked over the 36 x 68 elements one by one -- 33 instructions per element with the store waiting on its own
    // load, ten times per thread: a third of the kernel's instructions and 43 % of its stall samples (ncu r02).  One
    // plane, fp32, rows a multiple of 4 wide: whole rows as aligned float4, all of a thread's (<= 3) loads in flight
    // before the first store.
    bool vec_load = false;
    if constexpr (sizeof(T) == 4) vec_load = iwe.n == 1 && (Wp & 3) == 0 && (reinterpret_cast<size_t>(iwe.p[0]) & 15) == 0;
    if (vec_load) {
      if constexpr (sizeof(T) == 4) {
        constexpr int Q = GM_LW / 4, NV = (GM_TH + 4) * Q, PER = (NV + GM_THREADS - 1) / GM_THREADS;
        const float4* src = reinterpret_cast<const float4*>(iwe.p[0]);
        const int wq = Wp >> 2, q0 = (c0 >> 2) - 1;
        float4 v[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int i = threadIdx.x + j * GM_THREADS;
          const int lr = min(i / Q, GM_TH + 3), q = i - (i / Q) * Q;
          const int r = min(max(r0 - 2 + lr, 0), Hp - 1), cq = min(max(q0 + q, 0), wq - 1);
          v[j] = __ldg(src + (int64_t)r * wq + cq);
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int i = threadIdx.x + j * GM_THREADS;
          if (i < NV) *reinterpret_cast<float4*>(&sI[i / Q][(i - (i / Q) * Q) * 4]) = v[j];
        }
      }
    } else {
      for (int i = threadIdx.x; i < (GM_TH + 4) * (GM_TW + 4); i += blockDim.x) {
        const int lr = i / (GM_TW + 4), lc = i - lr * (GM_TW + 4);
        const int r = min(max(r0 - 2 + lr, 0), Hp - 1), c = min(max(c0 - 2 + lc, 0), Wp - 1);
        sI[lr][lc + 2] = iwe.load((int64_t)r * Wp + c);
      }
    }
    __syncthreads();
    const int tx = threadIdx.x & (GM_TW - 1), ty = threadIdx.x / GM_TW;
    const int c = c0 + tx;
    const bool col_ok = c >= 3 && c <= Wp - 4;
    const T c8 = coef * (T)0.125;
    T P[5], Q[5], S[5], D[5];
    T tpart = 0;   // at most GM_ROWS terms per thread before the double accumulation
#pragma unroll
    for (int k = 0; k < GM_ROWS + 4; ++k) {
      // horizontal pass on image row r0 + ty*GM_ROWS - 2 + k
      const T* row = &sI[ty * GM_ROWS + k][tx + 2];
      const T i0 = row[0], i1 = row[1], i2 = row[2], i3 = row[3], i4 = row[4];
      const T h0 = i0 + i4, h1 = i1 + i3;
#pragma unroll
      for (int j = 0; j < 4; ++j) { P[j] = P[j + 1]; Q[j] = Q[j + 1]; S[j] = S[j + 1]; D[j] = D[j + 1]; }
      P[4] = h0 + (T)4 * h1 + (T)6 * i2;
      Q[4] = (T)2 * i2 - h0;
      S[4] = h1 + (T)2 * i2;
      D[4] = i3 - i1;
      if (k >= 4) {
        const int r = r0 + ty * GM_ROWS + k - 4;   // centre row of the five-row window
        if (col_ok && r >= 3 && r <= Hp - 4) {
          const T gx = (S[3] - S[1]) * (T)0.125, gy = (D[1] + (T)2 * D[2] + D[3]) * (T)0.125;
          tpart += gx * gx + gy * gy;
          const T v = ((T)2 * P[2] - P[0] - P[4]) + ((Q[0] + Q[4]) + (T)4 * (Q[1] + Q[3]) + (T)6 * Q[2]);
          g[(int64_t)r * Wp + c] = c8 * v;
        }
      }
    }
    part = (double)tpart;
  } else {
    pdl_wait();
  }
  part = block_sum(part, sm);
  if (threadIdx.x == 0) atomicAdd(acc + kAccGradSlots + (blockIdx.x & (kAccSpread - 1)), part);
}

// ---- total variation of the flow: value + gradient -------------------------------------------------
// torch.gradient (spacing 1, edge_order 1): interior G(q) = (f[q+1]-f[q-1])/2, edges one-sided.
// d/df[i] of sum_q |G(q) w(q)|  =  sum_q s(q) dG(q)/df[i],  s(q) = sign(G(q) w(q)) w(q).  Only q = i-1, i, i+1
// touch f[i]:   q=i-1: +1/2 (or +1 when q is the first sample);  q=i+1: -1/2 (or -1 when q is the last
// sample);  q=i: -1 at the first sample, +1 at the last, 0 inside.  The five samples f[i-2..i+2] suffice.
template <typename T>
__device__ __forceinline__ T sgn(T v) { return (v > (T)0) ? (T)1 : ((v < (T)0) ? (T)-1 : (T)0); }

// v[0..4] = f[i-2..i+2] (entries outside [0,n) are never used), w[0..2] = weights at i-1, i, i+1.
// Returns the adjoint at i and the forward |G(i) w(i)| term through `absterm`.
template <typename T>
__device__ __forceinline__ T tv_axis(const T v[5], const T w[3], int i, int n, T& absterm) {
  const bool first = i == 0, last = i == n - 1;
  // G(i)
  const T gi = first ? (v[3] - v[2]) : (last ? (v[2] - v[1]) : (v[3] - v[1]) * (T)0.5);
  const T giw = gi * w[1];
  absterm = giw < (T)0 ? -giw : giw;
  T out = 0;
  if (first) out -= sgn(giw) * w[1];
  if (last) out += sgn(giw) * w[1];
  if (i >= 1) {  // q = i-1
    const bool qfirst = i - 1 == 0;
    const T gq = qfirst ? (v[2] - v[1]) : (v[2] - v[0]) * (T)0.5;
    out += (qfirst ? (T)1 : (T)0.5) * sgn(gq * w[0]) * w[0];
  }
  if (i <= n - 2) {  // q = i+1
    const bool qlast = i + 1 == n - 1;
    const T gq = qlast ? (v[3] - v[2]) : (v[4] - v[2]) * (T)0.5;
    out -= (qlast ? (T)1 : (T)0.5) * sgn(gq * w[2]) * w[2];
  }
  return out;
}

// One thread = 4 consecutive columns of one row of one channel.  Deep-interior quads (two samples away from
// every edge, unit weights) take a lean path: five float4 row loads + two float4 column neighbours, the
// adjoint reduces to  0.5*(sgn(f[i]-f[i-2]) - sgn(f[i+2]-f[i]))  per axis.  Everything else (edges, per-pixel
// weights, unaligned widths) goes through the general per-element code.  The kernel was instruction-bound
// (~200 instructions per element in the first version).
template <typename T, bool HAS_WTS>
__device__ __noinline__ T tv_element_value(const T* __restrict__ f, const T* __restrict__ weights, int H, int W,
                                           int r, int c, T coef, double& part) {
  T vr[5], vc[5], wr[3] = {(T)1, (T)1, (T)1}, wc[3] = {(T)1, (T)1, (T)1};
#pragma unroll
  for (int o = -2; o <= 2; ++o) {
    const int rr = min(max(r + o, 0), H - 1), cc = min(max(c + o, 0), W - 1);
    vr[o + 2] = __ldg(f + (int64_t)rr * W + c);
    vc[o + 2] = __ldg(f + (int64_t)r * W + cc);
  }
  if (HAS_WTS) {
#pragma unroll
    for (int o = -1; o <= 1; ++o) {
      const int rr = min(max(r + o, 0), H - 1), cc = min(max(c + o, 0), W - 1);
      wr[o + 1] = __ldg(weights + (int64_t)rr * W + c);
      wc[o + 1] = __ldg(weights + (int64_t)r * W + cc);
    }
  }
  T ar, ac;
  const T adj = tv_axis<T>(vr, wr, r, H, ar) + tv_axis<T>(vc, wc, c, W, ac);
  part += (double)ar + (double)ac;
  return coef * adj;
}
template <typename T, bool HAS_WTS>
__device__ __forceinline__ void tv_element_general(const T* __restrict__ f, const T* __restrict__ weights, int H, int W,
                                                   int r, int c, T coef, T* __restrict__ out, double& part) {
  out[(int64_t)r * W + c] = tv_element_value<T, HAS_WTS>(f, weights, H, W, r, c, coef, part);
}

__device__ __forceinline__ float sgn_diff(float a, float b) {  // sign(a - b) without forming NaN surprises
  const float d = a - b;
  return (d > 0.f ? 1.f : 0.f) - (d < 0.f ? 1.f : 0.f);
}

// Fused solver iteration: the TV kernel runs once per iteration, strictly before that iteration's Adam kernel and off
// the critical path (beside the splat).  Its first thread therefore also advances the device-side Adam step counter and
// evaluates the two bias-correction factors (two double pow() calls, ~1 us) into dst[0..1] = acc[kAccAdamCoefs..]:
// the Adam kernel used to do that itself, one thread per CTA with the other 255 waiting at a barrier (ncu r02s: 25 % of
// its stall samples).
constexpr int kAccAdamCoefs = 5;   // acc[5], acc[6]: lr / (1 - b1^t), 1 / sqrt(1 - b2^t) of the current iteration
struct StepCoefs {
  int32_t* step_dev;
  double lr, b1, b2;
  double* dst;
};
__device__ __forceinline__ void advance_step(const StepCoefs& sc) {
  const int step = *sc.step_dev + 1;
  *sc.step_dev = step;
  if (sc.dst) {
    sc.dst[0] = sc.lr / (1.0 - pow(sc.b1, (double)step));
    sc.dst[1] = 1.0 / sqrt(1.0 - pow(sc.b2, (double)step));
  }
}

template <typename T, bool HAS_WTS>
__global__ void __launch_bounds__(256) k_flow_tv(const T* __restrict__ flow, const T* __restrict__ weights, int H, int W,
                                                 T coef, double* __restrict__ acc, T* __restrict__ dflow, StepCoefs sc) {
  if (sc.step_dev && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) advance_step(sc);
  // coef = tv_scale / (2*H*W).  Grid (ceil(quads/64), ceil(H/4), 2), block = 64 quads x 4 rows: no div/mod in the
  // index math (the first lean version spent most of its 111 instructions per element on 64-bit div/mod).
  __shared__ double sm[32];
  const int quads = (W + 3) / 4;
  const bool lean_ok = sizeof(T) == 4 && !HAS_WTS && (W % 4 == 0) && ((reinterpret_cast<size_t>(flow) & 15) == 0) &&
                       ((reinterpret_cast<size_t>(dflow) & 15) == 0);
  double part = 0.0;
  float fpart = 0.f;
  const int q = blockIdx.x * 64 + (threadIdx.x & 63), r = blockIdx.y * 4 + (threadIdx.x >> 6), ch = blockIdx.z;
  for (int once = 0; once < 1 && q < quads && r < H; ++once) {
    const int c0 = q * 4;
    const T* f = flow + (int64_t)ch * H * W;
    T* out = dflow + (int64_t)ch * H * W;
    if constexpr (sizeof(T) == 4) {
      if (lean_ok && r >= 2 && r <= H - 3 && c0 >= 4 && c0 + 8 <= W) {
        const float* p = reinterpret_cast<const float*>(f) + (int64_t)r * W + c0;
        const float4 m2 = __ldg(reinterpret_cast<const float4*>(p - 2 * W)), m1 = __ldg(reinterpret_cast<const float4*>(p - W));
        const float4 ce = __ldg(reinterpret_cast<const float4*>(p));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(p + W)), p2 = __ldg(reinterpret_cast<const float4*>(p + 2 * W));
        const float4 lf = __ldg(reinterpret_cast<const float4*>(p - 4)), rt = __ldg(reinterpret_cast<const float4*>(p + 4));
        const float row[12] = {lf.x, lf.y, lf.z, lf.w, ce.x, ce.y, ce.z, ce.w, rt.x, rt.y, rt.z, rt.w};
        const float um2[4] = {m2.x, m2.y, m2.z, m2.w}, um1[4] = {m1.x, m1.y, m1.z, m1.w};
        const float up1[4] = {p1.x, p1.y, p1.z, p1.w}, up2[4] = {p2.x, p2.y, p2.z, p2.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float v = row[4 + k];
          // rows: q = r-1 and r+1 are interior samples here (2 <= r <= H-3)
          float adj = 0.5f * (sgn_diff(v, um2[k]) - sgn_diff(up2[k], v));
          // columns: c = c0+k with 4 <= c0 and c0+8 <= W  =>  c-2 >= 2 and c+2 <= W-3: interior
          adj += 0.5f * (sgn_diff(v, row[2 + k]) - sgn_diff(row[6 + k], v));
          fpart += 0.5f * (fabsf(up1[k] - um1[k]) + fabsf(row[5 + k] - row[3 + k]));
          o[k] = (float)coef * adj;
        }
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (int64_t)r * W + c0) = make_float4(o[0], o[1], o[2], o[3]);
        continue;
      }
    }
    for (int k = 0; k < 4 && c0 + k < W; ++k) tv_element_general<T, HAS_WTS>(f, weights, H, W, r, c0 + k, coef, out, part);
  }
  part += (double)fpart;
  part = block_sum(part, sm);
  if (threadIdx.x == 0 && acc) atomicAdd(acc + 3, part);
}

// Row-marching version for the lean case (fp32, unit weights, W % 4 == 0, 16-byte aligned planes): a thread owns
// one quad of columns and TV_ROWS consecutive rows and keeps the five rows r-2..r+2 of its quad in registers, so
// a flow row is loaded ~1.5 times instead of 5, and the index set-up and the block reduction are paid once per
// 32 elements instead of once per 4 (the one-quad-per-thread kernel above executes ~100 instructions per element).
constexpr int TV_ROWS = 8;

// frame element idx -> (r, c): rows {0, 1, H-2, H-1} over all columns, then columns {0..3, W-4..W-1} of the rows between
__device__ __forceinline__ bool tv_frame_coord(int64_t idx, int H, int W, int& r, int& c) {
  if (idx < (int64_t)4 * W) {
    const int k = (int)(idx / W);
    r = k < 2 ? k : H - 4 + k; c = (int)(idx - (int64_t)k * W);
    return true;
  }
  const int64_t j = idx - (int64_t)4 * W;
  if (j >= (int64_t)8 * (H - 4)) return false;
  const int k = (int)(j & 7);
  r = 2 + (int)(j >> 3); c = k < 4 ? k : W - 8 + k;
  return true;
}

// grid.x = n_frame_ctas + fast CTAs (64 quads x 2 row strips each), grid.y = channel.  Needs H >= 5, W >= 12.
__global__ void __launch_bounds__(128) k_flow_tv_march(const float* __restrict__ flow, int H, int W, float coef,
                                                       double* __restrict__ acc, float* __restrict__ dflow,
                                                       StepCoefs sc, int n_frame_ctas, int fast_gx) {
  if (sc.step_dev && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) advance_step(sc);
  __shared__ double sm[32];
  const int ch = blockIdx.y;
  const float* f = flow + (int64_t)ch * H * W;
  float* out = dflow + (int64_t)ch * H * W;
  double part = 0.0;
  if ((int)blockIdx.x < n_frame_ctas) {
    // edges: one element per thread through the general code (its ten loads are independent: one round trip)
    int r, c;
    if (tv_frame_coord((int64_t)blockIdx.x * blockDim.x + threadIdx.x, H, W, r, c))
      tv_element_general<float, false>(f, nullptr, H, W, r, c, coef, out, part);
  } else {
    const int t = blockIdx.x - n_frame_ctas;
    const int q = (t % fast_gx) * 64 + (threadIdx.x & 63), rs = 2 + ((t / fast_gx) * 2 + (threadIdx.x >> 6)) * TV_ROWS;
    const int c0 = q * 4;
    if (c0 >= 4 && c0 + 8 <= W && rs <= H - 3) {
      // every load is unconditional (row index clamped), only the stores are guarded: the unrolled loop has no
      // control flow around its loads, so they are all in flight together
      auto rowp = [&](int r) { return f + (int64_t)min(r, H - 1) * W + c0; };
      float4 m2 = __ldg(reinterpret_cast<const float4*>(rowp(rs - 2))), m1 = __ldg(reinterpret_cast<const float4*>(rowp(rs - 1)));
      float4 ce = __ldg(reinterpret_cast<const float4*>(rowp(rs))), p1 = __ldg(reinterpret_cast<const float4*>(rowp(rs + 1)));
      float fpart = 0.f;
#pragma unroll
      for (int i = 0; i < TV_ROWS; ++i) {
        const int r = rs + i;
        const float4 p2 = __ldg(reinterpret_cast<const float4*>(rowp(r + 2)));
        const float4 lf = __ldg(reinterpret_cast<const float4*>(rowp(r) - 4)), rt = __ldg(reinterpret_cast<const float4*>(rowp(r) + 4));
        const float row[8] = {lf.z, lf.w, ce.x, ce.y, ce.z, ce.w, rt.x, rt.y};   // columns c0-2 .. c0+5
        const float um2[4] = {m2.x, m2.y, m2.z, m2.w}, um1[4] = {m1.x, m1.y, m1.z, m1.w};
        const float up1[4] = {p1.x, p1.y, p1.z, p1.w}, up2[4] = {p2.x, p2.y, p2.z, p2.w};
        float o[4], rpart = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float v = row[2 + k];
          // rows and columns: every q = i-1, i, i+1 involved is an interior sample here (see k_flow_tv)
          float adj = 0.5f * (sgn_diff(v, um2[k]) - sgn_diff(up2[k], v));
          adj += 0.5f * (sgn_diff(v, row[k]) - sgn_diff(row[4 + k], v));
          rpart += 0.5f * (fabsf(up1[k] - um1[k]) + fabsf(row[3 + k] - row[1 + k]));
          o[k] = coef * adj;
        }
        if (r <= H - 3) {
          fpart += rpart;
          *reinterpret_cast<float4*>(out + (int64_t)r * W + c0) = make_float4(o[0], o[1], o[2], o[3]);
        }
        m2 = m1; m1 = ce; ce = p1; p1 = p2;
      }
      part = (double)fpart;   // 64 non-negative terms
    }
  }
  part = block_sum(part, sm);
  if (threadIdx.x == 0 && acc) atomicAdd(acc + kAccTvSlots + ((blockIdx.x + 5 * blockIdx.y) & (kAccSpread - 1)), part);
}

// ---- loss scalar -------------------------------------------------------------------------------------
__device__ __forceinline__ double acc_total(const double* __restrict__ acc, int single, int spread) {
  double t = acc[single];
#pragma unroll
  for (int i = 0; i < kAccSpread; ++i) t += acc[spread + i];
  return t;
}
template <typename T>
__global__ void k_loss_finalize(int kind, const double* __restrict__ acc, int Hp, int Wp, int H, int W, int omit,
                                double data_scale, double tv_scale, T* __restrict__ loss) {
  const double cnt = omit ? (double)(Hp - 2) * (double)(Wp - 2) : (double)Hp * (double)Wp;
  double data = 0.0;
  if (kind == EBOS_COST_VARIANCE) data = -((acc[1] - acc[0] * acc[0] / cnt) / (cnt - 1.0));
  else if (kind == EBOS_COST_GRADMAG) data = -(acc_total(acc, 2, kAccGradSlots) / cnt);
  const double tv = acc_total(acc, 3, kAccTvSlots) / (2.0 * (double)H * (double)W);
  loss[0] = (T)(data_scale * data + tv_scale * tv);
}

// ---- Adam ----------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void adam_one(T& p, T g, T& m, T& v, T b1, T b2, T eps, T step_size, T inv_bc2_sqrt) {
  m = m * b1 + ((T)1 - b1) * g;
  v = v * b2 + ((T)1 - b2) * g * g;
  const T denom = sqrt(v) * inv_bc2_sqrt + eps;
  p -= step_size * (m / denom);
}

// Optional tail work of the fused solver iteration, done by one thread of the Adam kernel (it runs after every
// producer/consumer of `acc`): scalar loss from the accumulators, then reset them for the next iteration.
struct FinalizeArgs {
  int enabled, kind, Hp, Wp, H, W, omit;
  double data_scale, tv_scale;
  double* acc;
  void* loss;
};

template <typename T>
__device__ __forceinline__ void finalize_loss(const FinalizeArgs& f) {
  const double cnt = f.omit ? (double)(f.Hp - 2) * (double)(f.Wp - 2) : (double)f.Hp * (double)f.Wp;
  double data = 0.0;
  if (f.kind == EBOS_COST_VARIANCE) data = -((f.acc[1] - f.acc[0] * f.acc[0] / cnt) / (cnt - 1.0));
  else if (f.kind == EBOS_COST_GRADMAG) data = -(acc_total(f.acc, 2, kAccGradSlots) / cnt);
  const double tv = acc_total(f.acc, 3, kAccTvSlots) / (2.0 * (double)f.H * (double)f.W);
  reinterpret_cast<T*>(f.loss)[0] = (T)(f.data_scale * data + f.tv_scale * tv);
  // (the Adam coefficients of this iteration stay: other CTAs of the Adam kernel may not have read them yet)
#pragma unroll
  for (int i = 0; i < EBOS_ACC_DOUBLES; ++i)
    if (i != kAccAdamCoefs && i != kAccAdamCoefs + 1) f.acc[i] = 0.0;
}

// step_mode 0: `step_host`;  1: *step_dev + 1 (bumped afterwards by k_adam_bump);  2: *step_dev (already advanced);
// 3: the two factors are read from coefs_dev (written by the TV kernel of the same iteration, see advance_step)
// (4 CTAs/SM = 62 registers: 10.7 us on the [2,720,1280] flow; 13.4 us unconstrained at 78 registers / 3 CTAs, 11.2 us at 40
//  registers with spills)
template <typename T, int MINB = 4>
__global__ void __launch_bounds__(256, MINB) k_adam(T* __restrict__ p, const T* __restrict__ g, T* __restrict__ m,
                                              T* __restrict__ v, int64_t n, double lr, double b1, double b2, double eps,
                                              int step_host, const int32_t* __restrict__ step_dev, int step_mode,
                                              FinalizeArgs fin, const double* __restrict__ coefs_dev = nullptr) {
  // bias corrections: two double pow() calls cost ~300 instructions; done by one thread per CTA and shared through
  // shared memory (the first version had every thread do them: 163 instructions per float4, ncu r01e).  The first
  // loads of every thread are issued before the barrier so that their latency overlaps the pow().
  __shared__ T s_coef[2];
  pdl_wait();   // dflow of the preceding backward (fused iteration: launched with the PDL attribute)
  auto coefs = [&]() {
    if (step_mode == 3) {        // block-uniform: no barrier, every thread reads the two factors (L2 hits)
      s_coef[0] = (T)coefs_dev[0];
      s_coef[1] = (T)coefs_dev[1];
      __syncwarp();
      return;
    }
    if (threadIdx.x == 0) {
      int step = step_host;
      if (step_mode == 1) step = *step_dev + 1;
      else if (step_mode == 2) step = *step_dev;
      const double bc1 = 1.0 - pow(b1, (double)step);
      const double bc2 = 1.0 - pow(b2, (double)step);
      s_coef[0] = (T)(lr / bc1);
      s_coef[1] = (T)(1.0 / sqrt(bc2));
    }
    __syncthreads();
  };
  const T tb1 = (T)b1, tb2 = (T)b2, teps = (T)eps;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if constexpr (sizeof(T) == 4) {
    // 16-byte accesses: 7 plane streams, the kernel is pure bandwidth
    if ((((size_t)p | (size_t)g | (size_t)m | (size_t)v) & 15) == 0) {
      const int64_t n4 = n >> 2;
      int64_t i = tid;
      float4 pp, mm, vv, gg;
      if (i < n4) {
        pp = reinterpret_cast<float4*>(p)[i]; mm = reinterpret_cast<float4*>(m)[i]; vv = reinterpret_cast<float4*>(v)[i];
        gg = __ldg(reinterpret_cast<const float4*>(g) + i);
      }
      coefs();
      const float step_size = s_coef[0], inv_bc2_sqrt = s_coef[1];
      while (i < n4) {
        adam_one<float>(pp.x, gg.x, mm.x, vv.x, tb1, tb2, teps, step_size, inv_bc2_sqrt);
        adam_one<float>(pp.y, gg.y, mm.y, vv.y, tb1, tb2, teps, step_size, inv_bc2_sqrt);
        adam_one<float>(pp.z, gg.z, mm.z, vv.z, tb1, tb2, teps, step_size, inv_bc2_sqrt);
        adam_one<float>(pp.w, gg.w, mm.w, vv.w, tb1, tb2, teps, step_size, inv_bc2_sqrt);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        i += nth;
        if (i < n4) {
          pp = reinterpret_cast<float4*>(p)[i]; mm = reinterpret_cast<float4*>(m)[i]; vv = reinterpret_cast<float4*>(v)[i];
          gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        }
      }
      for (int64_t k = n4 * 4 + tid; k < n; k += nth) {
        T p1 = p[k], m1 = m[k], v1 = v[k];
        adam_one<T>(p1, __ldg(g + k), m1, v1, tb1, tb2, teps, step_size, inv_bc2_sqrt);
        p[k] = p1; m[k] = m1; v[k] = v1;
      }
      if (fin.enabled && tid == 0) finalize_loss<T>(fin);
      return;
    }
  }
  coefs();
  const T step_size = s_coef[0], inv_bc2_sqrt = s_coef[1];
  for (int64_t i = tid; i < n; i += nth) {
    T pp = p[i], mm = m[i], vv = v[i];
    adam_one<T>(pp, __ldg(g + i), mm, vv, tb1, tb2, teps, step_size, inv_bc2_sqrt);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
  if (fin.enabled && tid == 0) finalize_loss<T>(fin);
}
__global__ void k_adam_bump(int32_t* step_dev) { *step_dev += 1; }

// ---- Adam with the TV term folded in (fused solver iteration, fp32 lean case) -----------------------------------
// The unfused iteration spends a whole launch and 4P of traffic on the TV kernel (read the flow, write lambda * dTV as
// the start value of the gradient plane) before the Adam kernel reads that plane back.  Here the Adam kernel evaluates
// the TV stencil itself: it marches over the OLD flow exactly like k_flow_tv_march (rows r-2..r+2 of a quad of columns
// in registers), adds lambda * dTV to the data gradient the backward accumulated on a ZEROED plane, applies the Adam
// update, writes the NEW flow to a second buffer (neighbouring threads still need the old values: ping-pong) and
// zero-fills the gradient plane for the next iteration.  16P instead of 14P + 4P per iteration, one launch less.
// The TV value of this iteration's loss is summed here as well, so the scalar loss, the accumulator reset and the step
// counter are done by whichever CTA finishes last (ticket in acc[kAccTicket]).
constexpr int kAccTicket = 4;   // acc[4] (unused by the sums) reinterpreted as the unsigned ticket counter
constexpr int ADAM_ROWS = 2;    // rows per thread of the fused Adam + TV kernel

__global__ void __launch_bounds__(128) k_adam_tv_march(const float* __restrict__ flow_in, float* __restrict__ flow_out,
                                                       float* __restrict__ dflow, float* __restrict__ em,
                                                       float* __restrict__ ev, int H, int W, float coef, double lr,
                                                       double b1, double b2, double eps, int32_t* __restrict__ step_dev,
                                                       FinalizeArgs fin, int n_frame_ctas, int fast_gx) {
  __shared__ double sm[32];
  __shared__ float s_coef[2];
  const int ch = blockIdx.y;
  const int64_t plane = (int64_t)ch * H * W;
  const float* f = flow_in + plane;
  float* fo = flow_out + plane;
  float* g = dflow + plane;
  float* pm = em + plane;
  float* pv = ev + plane;
  const float tb1 = (float)b1, tb2 = (float)b2, teps = (float)eps;
  double part = 0.0;
  const bool frame = (int)blockIdx.x < n_frame_ctas;
  // geometry of the fast region: like k_flow_tv_march, but only ADAM_ROWS rows per thread -- this kernel streams seven
  // planes from HBM and wants every load of a thread in flight at once (a first version marched over 8 rows with the
  // loads of row i+1 behind the stores of row i: 12 warps per SM, eight serial memory round trips each, slower than the
  // two kernels it replaced)
  const int t = blockIdx.x - n_frame_ctas;
  const int q = (t % fast_gx) * 64 + (threadIdx.x & 63), rs = 2 + ((t / fast_gx) * 2 + (threadIdx.x >> 6)) * ADAM_ROWS;
  const int c0 = q * 4;
  const bool fast = !frame && c0 >= 4 && c0 + 8 <= W && rs <= H - 3;
  auto rowp = [&](int r) { return f + (int64_t)min(r, H - 1) * W + c0; };
  // everything that does not depend on the backward is requested before the grid dependency is awaited: the old flow
  // is read-only during the iteration, the moments are touched by this kernel only
  float4 rows[ADAM_ROWS + 4], lf[ADAM_ROWS], rt[ADAM_ROWS], m4[ADAM_ROWS], v4[ADAM_ROWS], g4[ADAM_ROWS];
  int64_t idx[ADAM_ROWS];
  if (fast) {
#pragma unroll
    for (int j = 0; j < ADAM_ROWS + 4; ++j) rows[j] = __ldg(reinterpret_cast<const float4*>(rowp(rs - 2 + j)));
#pragma unroll
    for (int i = 0; i < ADAM_ROWS; ++i) {
      lf[i] = __ldg(reinterpret_cast<const float4*>(rowp(rs + i) - 4));
      rt[i] = __ldg(reinterpret_cast<const float4*>(rowp(rs + i) + 4));
      idx[i] = (int64_t)min(rs + i, H - 3) * W + c0;
      m4[i] = *reinterpret_cast<const float4*>(pm + idx[i]);
      v4[i] = *reinterpret_cast<const float4*>(pv + idx[i]);
    }
  }
  int fr = -1, fc = -1;
  float ftv = 0.f;
  if (frame && tv_frame_coord((int64_t)blockIdx.x * blockDim.x + threadIdx.x, H, W, fr, fc))
    ftv = tv_element_value<float, false>(f, nullptr, H, W, fr, fc, coef, part);
  else
    fr = -1;
  if (threadIdx.x == 0) {
    const int step = *step_dev + 1;
    const double bc1 = 1.0 - pow(b1, (double)step);
    const double bc2 = 1.0 - pow(b2, (double)step);
    s_coef[0] = (float)(lr / bc1);
    s_coef[1] = (float)(1.0 / sqrt(bc2));
  }
  pdl_wait();          // dflow of the preceding backward
  if (fast) {
#pragma unroll
    for (int i = 0; i < ADAM_ROWS; ++i) g4[i] = *reinterpret_cast<const float4*>(g + idx[i]);
  }
  __syncthreads();
  const float step_size = s_coef[0], inv_bc2_sqrt = s_coef[1];
  if (fr >= 0) {
    const int64_t i = (int64_t)fr * W + fc;
    float x = f[i], mm = pm[i], vv = pv[i];
    adam_one<float>(x, g[i] + ftv, mm, vv, tb1, tb2, teps, step_size, inv_bc2_sqrt);
    fo[i] = x; pm[i] = mm; pv[i] = vv; g[i] = 0.f;
  }
  if (fast) {
    float fpart = 0.f;
#pragma unroll
    for (int i = 0; i < ADAM_ROWS; ++i) {
      const int r = rs + i;
      const float4 m2 = rows[i], m1 = rows[i + 1], ce = rows[i + 2], p1 = rows[i + 3], p2 = rows[i + 4];
      const float row[8] = {lf[i].z, lf[i].w, ce.x, ce.y, ce.z, ce.w, rt[i].x, rt[i].y};   // columns c0-2 .. c0+5
      const float um2[4] = {m2.x, m2.y, m2.z, m2.w}, um1[4] = {m1.x, m1.y, m1.z, m1.w};
      const float up1[4] = {p1.x, p1.y, p1.z, p1.w}, up2[4] = {p2.x, p2.y, p2.z, p2.w};
      const float gd[4] = {g4[i].x, g4[i].y, g4[i].z, g4[i].w};
      float xm[4] = {m4[i].x, m4[i].y, m4[i].z, m4[i].w}, xv[4] = {v4[i].x, v4[i].y, v4[i].z, v4[i].w}, xo[4];
      float rpart = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float val = row[2 + k];
        float adj = 0.5f * (sgn_diff(val, um2[k]) - sgn_diff(up2[k], val));
        adj += 0.5f * (sgn_diff(val, row[k]) - sgn_diff(row[4 + k], val));
        rpart += 0.5f * (fabsf(up1[k] - um1[k]) + fabsf(row[3 + k] - row[1 + k]));
        xo[k] = val;
        adam_one<float>(xo[k], gd[k] + coef * adj, xm[k], xv[k], tb1, tb2, teps, step_size, inv_bc2_sqrt);
      }
      if (r <= H - 3) {
        fpart += rpart;
        *reinterpret_cast<float4*>(fo + idx[i]) = make_float4(xo[0], xo[1], xo[2], xo[3]);
        *reinterpret_cast<float4*>(pm + idx[i]) = make_float4(xm[0], xm[1], xm[2], xm[3]);
        *reinterpret_cast<float4*>(pv + idx[i]) = make_float4(xv[0], xv[1], xv[2], xv[3]);
        *reinterpret_cast<float4*>(g + idx[i]) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    part = (double)fpart;
  }
  part = block_sum(part, sm);
  if (threadIdx.x == 0) {
    atomicAdd(fin.acc + kAccTvSlots + ((blockIdx.x + 5 * blockIdx.y) & (kAccSpread - 1)), part);
    __threadfence();
    const unsigned total = gridDim.x * gridDim.y;
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(fin.acc + kAccTicket), 1u);
    if (ticket == total - 1) {
      __threadfence();
      // every CTA's sums are in: scalar loss (reads through L2), reset of ALL accumulators incl. the ticket, step++
      const volatile double* a = fin.acc;
      const double cnt = fin.omit ? (double)(fin.Hp - 2) * (double)(fin.Wp - 2) : (double)fin.Hp * (double)fin.Wp;
      double data = 0.0, tv = a[3];
      if (fin.kind == EBOS_COST_VARIANCE) data = -((a[1] - a[0] * a[0] / cnt) / (cnt - 1.0));
      else if (fin.kind == EBOS_COST_GRADMAG) {
        double sgm = a[2];
        for (int i = 0; i < kAccSpread; ++i) sgm += a[kAccGradSlots + i];
        data = -(sgm / cnt);
      }
      for (int i = 0; i < kAccSpread; ++i) tv += a[kAccTvSlots + i];
      tv /= 2.0 * (double)fin.H * (double)fin.W;
      reinterpret_cast<float*>(fin.loss)[0] = (float)(fin.data_scale * data + fin.tv_scale * tv);
      for (int i = 0; i < EBOS_ACC_DOUBLES; ++i) fin.acc[i] = 0.0;
      *step_dev += 1;
    }
  }
}

// out[i] = sum over the peer buffers, in rank order (the one-shot all-reduce of the partial flow gradients)
template <typename T>
__global__ void __launch_bounds__(256) k_sum_peers(const PeerPlanes<T> pl, int64_t n, T* __restrict__ out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  bool vec = sizeof(T) == 4 && (reinterpret_cast<size_t>(out) & 15) == 0;
  for (int r = 0; r < pl.n; ++r) vec = vec && (reinterpret_cast<size_t>(pl.p[r]) & 15) == 0;
  int64_t done = 0;
  if constexpr (sizeof(T) == 4) {
    if (vec) {
      const int64_t n4 = n >> 2;
      for (int64_t i = tid; i < n4; i += nth) {
        float4 a = reinterpret_cast<const float4*>(pl.p[0])[i];
        for (int r = 1; r < pl.n; ++r) {
          const float4 b = reinterpret_cast<const float4*>(pl.p[r])[i];
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        reinterpret_cast<float4*>(out)[i] = a;
      }
      done = n4 * 4;
    }
  }
  for (int64_t i = done + tid; i < n; i += nth) out[i] = pl.load(i);
}

// Two-shot form of the same reduction for larger rank counts (the one-shot pass reads R whole planes per rank: 7 remote
// planes = 52 MB per rank for the flow gradient at R = 8): (1) reduce-scatter -- rank r sums slice r of every peer's
// buffer, IN PLACE into its own buffer (only rank r ever reads slice r of rank r's buffer); (2) after a barrier,
// all-gather -- every rank copies slice q from rank q's buffer.  2 (R-1)/R planes per rank instead of R-1.
template <typename T>
__global__ void __launch_bounds__(256) k_reduce_slice(const PeerPlanes<T> pl, int64_t begin, int64_t end, T* dst) {   // (dst may alias a peer plane)
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  bool vec = sizeof(T) == 4 && ((begin | end) & 3) == 0 && (reinterpret_cast<size_t>(dst) & 15) == 0;
  for (int r = 0; r < pl.n; ++r) vec = vec && (reinterpret_cast<size_t>(pl.p[r]) & 15) == 0;
  if constexpr (sizeof(T) == 4) {
    if (vec) {
      for (int64_t i = (begin >> 2) + tid; i < (end >> 2); i += nth) {
        float4 a = reinterpret_cast<const float4*>(pl.p[0])[i];
        for (int r = 1; r < pl.n; ++r) {
          const float4 b = reinterpret_cast<const float4*>(pl.p[r])[i];
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        reinterpret_cast<float4*>(dst)[i] = a;
      }
      return;
    }
  }
  for (int64_t i = begin + tid; i < end; i += nth) dst[i] = pl.load(i);
}

template <typename T>
__global__ void __launch_bounds__(256) k_gather_slices(const PeerPlanes<T> pl, int64_t n, int64_t slice, T* __restrict__ out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  bool vec = sizeof(T) == 4 && (slice & 3) == 0 && (n & 3) == 0 && (reinterpret_cast<size_t>(out) & 15) == 0;
  for (int r = 0; r < pl.n; ++r) vec = vec && (reinterpret_cast<size_t>(pl.p[r]) & 15) == 0;
  if constexpr (sizeof(T) == 4) {
    if (vec) {
      for (int64_t i = tid; i < (n >> 2); i += nth) {
        const int owner = (int)min((i << 2) / slice, (int64_t)pl.n - 1);
        reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(pl.p[owner])[i];
      }
      return;
    }
  }
  for (int64_t i = tid; i < n; i += nth) out[i] = pl.p[(int)min(i / slice, (int64_t)pl.n - 1)][i];
}

// In-switch form (NVLS): `mc` is the MULTICAST mapping of a symmetric buffer -- one address that stands for the same
// offset in every rank's copy.  multimem.ld_reduce has the NVSwitch fetch and add the R copies of an element and return the
// sum; multimem.st writes a value into all R copies.  Rank r does both for slice r only (one reducer per element: every
// rank ends with bit-identical values): the reduce-scatter, the barrier between the two shots and the all-gather of the
// two-shot form collapse into this one pass, and a rank moves 2/R of a plane over its links instead of 2 (R-1)/R.
__global__ void __launch_bounds__(256) k_multimem_allreduce_slice(float* __restrict__ mc, int64_t begin4, int64_t end4) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = begin4 + tid; i < end4; i += nth) {
    float* a = mc + 4 * i;
    float x, y, z, w;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "l"(a) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
  }
  __threadfence_system();   // the broadcast stores are performed before this rank enters the barrier that follows
}

// ---- launch helpers ----------------------------------------------------------------------------------
// 2-D grid-stride launch shape for plane kernels: about `per_sm` CTAs per SM in total (the reductions end in
// one same-address atomic per CTA, so fewer, longer-lived CTAs are better).
static dim3 plane_grid2d(int rows, int cols, int per_sm = 8) {
  int bx = std::max(1, std::min((cols + 255) / 256, 8));
  int by = std::max(1, std::min(rows, (sm_count() * per_sm + bx - 1) / bx));
  return dim3(bx, by);
}

template <typename T>
int iwe_cost_t(int kind, const T* iwe, int Hp, int Wp, int omit, double scale, double* acc, T* grad_iwe, cudaStream_t st,
               const void* const* peers = nullptr, int n_peers = 0) {
  const double cnt = omit ? (double)(Hp - 2) * (double)(Wp - 2) : (double)Hp * (double)Wp;
  if (kind == EBOS_COST_VARIANCE) {
    k_var_reduce<T><<<plane_grid2d(Hp, Wp, 2), 256, 0, st>>>(iwe, Hp, Wp, omit, acc);
    if (grad_iwe) k_var_grad<T><<<plane_grid2d(Hp, Wp), 256, 0, st>>>(iwe, Hp, Wp, omit, scale, acc, grad_iwe);
  } else if (kind == EBOS_COST_GRADMAG) {
    if (!grad_iwe) { set_error("ebos_iwe_cost: GRADMAG needs grad_iwe"); return EBOS_ERR_BAD_ARG; }
    const T coef = (T)(-2.0 * scale / (8.0 * cnt));
    static const bool legacy = getenv("EBOS_GRADMAG_LEGACY") != nullptr;   // the tiled kernel, kept for A/B runs
    if (legacy) {
      const int n_tiles = ((Wp + GTW - 1) / GTW) * ((Hp + GTH - 1) / GTH);
      const int grid = std::max(1, std::min(n_tiles, sm_count() * 4));
      k_gradmag<T><<<grid, 256, 0, st>>>(iwe, Hp, Wp, omit, coef, acc, grad_iwe);
    } else {
      const bool has_fast = Hp >= 7 && Wp >= 7;
      const int n_frame_ctas = has_fast ? 2 * ((Wp + GM_STRIP - 1) / GM_STRIP) + 2 * ((Hp - 6 + GM_STRIP - 1) / GM_STRIP)
                                        : (int)(((int64_t)Hp * Wp + GM_THREADS - 1) / GM_THREADS);
      const int n_tiles = has_fast ? ((Wp + GM_TW - 1) / GM_TW) * ((Hp + GM_TH - 1) / GM_TH) : 0;
      PeerPlanes<T> planes{};
      planes.n = peers ? n_peers : 1;
      for (int r = 0; r < planes.n; ++r) planes.p[r] = peers ? reinterpret_cast<const T*>(peers[r]) : iwe;
      cudaError_t le = launch_pdl(k_gradmag_sep<T>, dim3(n_frame_ctas + n_tiles), dim3(GM_THREADS), st, planes, Hp, Wp, omit, coef, acc,
                                  grad_iwe, n_frame_ctas, n_tiles);
      if (le != cudaSuccess) return cuda_fail(le, "ebos_iwe_cost(gradmag)");
    }
  } else if (kind != EBOS_COST_NONE) {
    set_error("ebos_iwe_cost: unknown cost kind");
    return EBOS_ERR_BAD_ARG;
  }
  EBOS_LAUNCH_CHECK("ebos_iwe_cost");
  return EBOS_OK;
}

template <typename T>
int flow_tv_t(const T* flow, const T* weights, int H, int W, double tv_scale, double* acc, T* dflow, cudaStream_t st,
              StepCoefs sc = StepCoefs{}) {
  int32_t* step_dev = sc.step_dev;
  if ((tv_scale == 0.0 && !step_dev) || H < 2 || W < 2) {
    cudaError_t e = cudaMemsetAsync(dflow, 0, (size_t)2 * H * W * sizeof(T), st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_flow_tv memset");
    if (tv_scale == 0.0) return EBOS_OK;
    set_error("ebos_flow_tv: torch.gradient needs at least 2 samples per axis");
    return EBOS_ERR_BAD_ARG;
  }
  const T coef = (T)(tv_scale / (2.0 * (double)H * (double)W));
  if constexpr (sizeof(T) == 4) {
    static const bool legacy = getenv("EBOS_TV_LEGACY") != nullptr;   // the one-quad-per-thread kernel, for A/B runs
    if (!legacy && !weights && W % 4 == 0 && W >= 12 && H >= 5 &&
        ((reinterpret_cast<size_t>(flow) | reinterpret_cast<size_t>(dflow)) & 15) == 0) {
      const int64_t n_frame = (int64_t)4 * W + (int64_t)8 * (H - 4);
      const int n_frame_ctas = (int)((n_frame + 127) / 128);
      const int fast_gx = ((W >> 2) + 63) / 64, fast_gy = (H - 4 + 2 * TV_ROWS - 1) / (2 * TV_ROWS);
      const dim3 mgrid(n_frame_ctas + fast_gx * fast_gy, 2);
      k_flow_tv_march<<<mgrid, 128, 0, st>>>(flow, H, W, coef, acc, dflow, sc, n_frame_ctas, fast_gx);
      EBOS_LAUNCH_CHECK("ebos_flow_tv");
      return EBOS_OK;
    }
  }
  const dim3 grid(((W + 3) / 4 + 63) / 64, (H + 3) / 4, 2);
  if (weights) k_flow_tv<T, true><<<grid, 256, 0, st>>>(flow, weights, H, W, coef, acc, dflow, sc);
  else k_flow_tv<T, false><<<grid, 256, 0, st>>>(flow, weights, H, W, coef, acc, dflow, sc);
  EBOS_LAUNCH_CHECK("ebos_flow_tv");
  return EBOS_OK;
}

static int adam_grid(int64_t n) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n / 4 + 255) / 256, (int64_t)sm_count() * 8));
}

// Data objective on the (optionally blurred) IWE.  blur_sigma > 0: B = blur3(IWE) (the sigma > 0 branch of
// EventImageConverter.create_image_from_events_tensor, src/event_image_converter.py:399-404), cost and dL/dB on B, then
// dL/dIWE = blur3^T(dL/dB) back into the plane that held B -- the backward reads an explicit gradient plane in that
// case also for the variance objective.  Returns in *gplane the plane ebos_window_backward must be given (NULL: derive
// the variance gradient from the IWE on the fly).
static int cost_with_blur(int kind, const void* iwe, int Hp, int Wp, int omit_boundary, double data_scale, int dtype,
                          double* acc, void* grad_iwe, double blur_sigma, void* blur_plane, cudaStream_t st, void** gplane) {
  int rc;
  if (blur_sigma > 0.0) {
    if (!blur_plane || !grad_iwe) { set_error("blur_sigma > 0 needs the blur_plane and grad_iwe scratch planes"); return EBOS_ERR_BAD_ARG; }
    if (Hp < 2 || Wp < 2) { set_error("blur_sigma > 0 needs an image of at least 2 x 2 (reflect padding)"); return EBOS_ERR_BAD_ARG; }
    rc = blur3_launch(iwe, Hp, Wp, blur_sigma, 0, dtype, blur_plane, st);
    if (rc) return rc;
    if (dtype == EBOS_F64) rc = iwe_cost_t<double>(kind, (const double*)blur_plane, Hp, Wp, omit_boundary, data_scale, acc, (double*)grad_iwe, st);
    else rc = iwe_cost_t<float>(kind, (const float*)blur_plane, Hp, Wp, omit_boundary, data_scale, acc, (float*)grad_iwe, st);
    if (rc) return rc;
    rc = blur3_launch(grad_iwe, Hp, Wp, blur_sigma, 1, dtype, blur_plane, st);
    *gplane = blur_plane;
    return rc;
  }
  // variance: no gradient plane, the backward derives it from (iwe, acc)
  *gplane = kind == EBOS_COST_GRADMAG ? grad_iwe : nullptr;
  if (dtype == EBOS_F64) return iwe_cost_t<double>(kind, (const double*)iwe, Hp, Wp, omit_boundary, data_scale, acc, (double*)*gplane, st);
  return iwe_cost_t<float>(kind, (const float*)iwe, Hp, Wp, omit_boundary, data_scale, acc, (float*)*gplane, st);
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

int ebos_iwe_cost(int kind, const void* iwe, int Hp, int Wp, int omit_boundary, double scale, int dtype, double* acc,
                  void* grad_iwe, void* stream) {
  EBOS_REQUIRE(iwe && acc && Hp > 0 && Wp > 0, "ebos_iwe_cost: bad argument");
  EBOS_REQUIRE(!omit_boundary || (Hp > 2 && Wp > 2), "ebos_iwe_cost: omit_boundary needs an image larger than 2x2");
  EBOS_CHECK_DTYPE(dtype, "ebos_iwe_cost");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(acc, 0, 3 * sizeof(double), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(acc + kAccGradSlots, 0, kAccSpread * sizeof(double), st);
  if (e != cudaSuccess) return cuda_fail(e, "ebos_iwe_cost memset");
  if (dtype == EBOS_F64) return iwe_cost_t<double>(kind, (const double*)iwe, Hp, Wp, omit_boundary, scale, acc, (double*)grad_iwe, st);
  return iwe_cost_t<float>(kind, (const float*)iwe, Hp, Wp, omit_boundary, scale, acc, (float*)grad_iwe, st);
}

int ebos_iwe_cost_peers(int kind, const void* const* iwe_peers, int n_peers, int Hp, int Wp, int omit_boundary, double scale,
                        int dtype, double* acc, void* grad_iwe, void* stream) {
  EBOS_REQUIRE(iwe_peers && acc && grad_iwe && Hp > 0 && Wp > 0 && n_peers >= 1 && n_peers <= kMaxPeers,
               "ebos_iwe_cost_peers: bad argument");
  EBOS_REQUIRE(kind == EBOS_COST_GRADMAG, "ebos_iwe_cost_peers: only the gradient-magnitude objective reads peer planes");
  EBOS_REQUIRE(!omit_boundary || (Hp > 2 && Wp > 2), "ebos_iwe_cost_peers: omit_boundary needs an image larger than 2x2");
  EBOS_REQUIRE(getenv("EBOS_GRADMAG_LEGACY") == nullptr, "ebos_iwe_cost_peers: not available with EBOS_GRADMAG_LEGACY");
  EBOS_CHECK_DTYPE(dtype, "ebos_iwe_cost_peers");
  for (int r = 0; r < n_peers; ++r) EBOS_REQUIRE(iwe_peers[r] != nullptr, "ebos_iwe_cost_peers: null peer plane");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(acc, 0, 3 * sizeof(double), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(acc + kAccGradSlots, 0, kAccSpread * sizeof(double), st);
  if (e != cudaSuccess) return cuda_fail(e, "ebos_iwe_cost_peers memset");
  if (dtype == EBOS_F64)
    return iwe_cost_t<double>(kind, (const double*)iwe_peers[0], Hp, Wp, omit_boundary, scale, acc, (double*)grad_iwe, st, iwe_peers, n_peers);
  return iwe_cost_t<float>(kind, (const float*)iwe_peers[0], Hp, Wp, omit_boundary, scale, acc, (float*)grad_iwe, st, iwe_peers, n_peers);
}

int ebos_sum_peers(const void* const* peers, int n_peers, int64_t n, int dtype, void* out, void* stream) {
  EBOS_REQUIRE(peers && out && n >= 0 && n_peers >= 1 && n_peers <= kMaxPeers, "ebos_sum_peers: bad argument");
  EBOS_CHECK_DTYPE(dtype, "ebos_sum_peers");
  if (n == 0) return EBOS_OK;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n / 4 + 255) / 256, (int64_t)sm_count() * 8));
  if (dtype == EBOS_F64) {
    PeerPlanes<double> pl{};
    pl.n = n_peers;
    for (int r = 0; r < n_peers; ++r) pl.p[r] = (const double*)peers[r];
    k_sum_peers<double><<<grid, 256, 0, as_stream(stream)>>>(pl, n, (double*)out);
  } else {
    PeerPlanes<float> pl{};
    pl.n = n_peers;
    for (int r = 0; r < n_peers; ++r) pl.p[r] = (const float*)peers[r];
    k_sum_peers<float><<<grid, 256, 0, as_stream(stream)>>>(pl, n, (float*)out);
  }
  EBOS_LAUNCH_CHECK("ebos_sum_peers");
  return EBOS_OK;
}

int ebos_reduce_peers_slice(const void* const* peers, int n_peers, int64_t begin, int64_t end, int dtype, void* dst,
                            void* stream) {
  EBOS_REQUIRE(peers && dst && n_peers >= 1 && n_peers <= kMaxPeers && begin >= 0 && end >= begin, "ebos_reduce_peers_slice: bad argument");
  EBOS_CHECK_DTYPE(dtype, "ebos_reduce_peers_slice");
  if (end == begin) return EBOS_OK;
  const int64_t cnt = end - begin;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((cnt / 4 + 255) / 256, (int64_t)sm_count() * 8));
  if (dtype == EBOS_F64) {
    PeerPlanes<double> pl{};
    pl.n = n_peers;
    for (int r = 0; r < n_peers; ++r) pl.p[r] = (const double*)peers[r];
    k_reduce_slice<double><<<grid, 256, 0, as_stream(stream)>>>(pl, begin, end, (double*)dst);
  } else {
    PeerPlanes<float> pl{};
    pl.n = n_peers;
    for (int r = 0; r < n_peers; ++r) pl.p[r] = (const float*)peers[r];
    k_reduce_slice<float><<<grid, 256, 0, as_stream(stream)>>>(pl, begin, end, (float*)dst);
  }
  EBOS_LAUNCH_CHECK("ebos_reduce_peers_slice");
  return EBOS_OK;
}

int ebos_multimem_allreduce_slice(void* multicast_base, int64_t begin, int64_t end, int dtype, void* stream) {
  EBOS_REQUIRE(multicast_base && begin >= 0 && end >= begin, "ebos_multimem_allreduce_slice: bad argument");
  if (dtype != EBOS_F32 || ((begin | end) & 3) || (reinterpret_cast<size_t>(multicast_base) & 15)) {
    set_error("ebos_multimem_allreduce_slice: needs fp32, a 16-byte aligned multicast address and a slice on multiples of 4 elements");
    return EBOS_ERR_UNSUPPORTED;
  }
  if (end == begin) return EBOS_OK;
  const int64_t n4 = (end - begin) >> 2;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n4 + 255) / 256, (int64_t)sm_count() * 4));
  k_multimem_allreduce_slice<<<grid, 256, 0, as_stream(stream)>>>((float*)multicast_base, begin >> 2, end >> 2);
  EBOS_LAUNCH_CHECK("ebos_multimem_allreduce_slice");
  return EBOS_OK;
}

int ebos_gather_peers_slices(const void* const* peers, int n_peers, int64_t n, int64_t slice, int dtype, void* out, void* stream) {
  EBOS_REQUIRE(peers && out && n_peers >= 1 && n_peers <= kMaxPeers && n >= 0 && slice >= 1, "ebos_gather_peers_slices: bad argument");
  EBOS_CHECK_DTYPE(dtype, "ebos_gather_peers_slices");
  if (n == 0) return EBOS_OK;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n / 4 + 255) / 256, (int64_t)sm_count() * 8));
  if (dtype == EBOS_F64) {
    PeerPlanes<double> pl{};
    pl.n = n_peers;
    for (int r = 0; r < n_peers; ++r) pl.p[r] = (const double*)peers[r];
    k_gather_slices<double><<<grid, 256, 0, as_stream(stream)>>>(pl, n, slice, (double*)out);
  } else {
    PeerPlanes<float> pl{};
    pl.n = n_peers;
    for (int r = 0; r < n_peers; ++r) pl.p[r] = (const float*)peers[r];
    k_gather_slices<float><<<grid, 256, 0, as_stream(stream)>>>(pl, n, slice, (float*)out);
  }
  EBOS_LAUNCH_CHECK("ebos_gather_peers_slices");
  return EBOS_OK;
}

int ebos_flow_tv(const void* flow, const void* weights, int H, int W, double tv_scale, int dtype, double* acc,
                 void* dflow, void* stream) {
  EBOS_REQUIRE(flow && dflow && H > 0 && W > 0, "ebos_flow_tv: bad argument");
  EBOS_CHECK_DTYPE(dtype, "ebos_flow_tv");
  cudaStream_t st = as_stream(stream);
  if (acc) {
    cudaError_t e = cudaMemsetAsync(acc + 3, 0, sizeof(double), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(acc + kAccTvSlots, 0, kAccSpread * sizeof(double), st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_flow_tv memset");
  }
  if (dtype == EBOS_F64) return flow_tv_t<double>((const double*)flow, (const double*)weights, H, W, tv_scale, acc, (double*)dflow, st);
  return flow_tv_t<float>((const float*)flow, (const float*)weights, H, W, tv_scale, acc, (float*)dflow, st);
}

int ebos_loss_finalize(int kind, const double* acc, int Hp, int Wp, int H, int W, int omit_boundary, double data_scale,
                       double tv_scale, int dtype, void* loss, void* stream) {
  EBOS_REQUIRE(acc && loss, "ebos_loss_finalize: bad argument");
  EBOS_CHECK_DTYPE(dtype, "ebos_loss_finalize");
  if (dtype == EBOS_F64)
    k_loss_finalize<double><<<1, 1, 0, as_stream(stream)>>>(kind, acc, Hp, Wp, H, W, omit_boundary, data_scale, tv_scale, (double*)loss);
  else
    k_loss_finalize<float><<<1, 1, 0, as_stream(stream)>>>(kind, acc, Hp, Wp, H, W, omit_boundary, data_scale, tv_scale, (float*)loss);
  EBOS_LAUNCH_CHECK("ebos_loss_finalize");
  return EBOS_OK;
}

int ebos_cmax_value_and_grad(const void* window, int64_t n, int flags, const void* flow, int H, int W, int pad_h,
                             int pad_w, int kind, int omit_boundary, double data_scale, double tv_scale,
                             const void* tv_weights, int dtype, void* iwe, void* grad_iwe, void* dflow, void* loss,
                             double* acc, int clean_workspace, double blur_sigma, void* blur_plane, void* stream) {
  EBOS_REQUIRE(window && flow && iwe && dflow && loss && acc && n >= 0 && H > 0 && W > 0 && pad_h >= 0 && pad_w >= 0,
               "ebos_cmax_value_and_grad: bad argument");
  EBOS_REQUIRE(blur_sigma >= 0.0, "ebos_cmax_value_and_grad: blur_sigma must not be negative");
  EBOS_REQUIRE(kind == EBOS_COST_VARIANCE || kind == EBOS_COST_GRADMAG, "ebos_cmax_value_and_grad: unknown cost kind");
  EBOS_REQUIRE(kind != EBOS_COST_GRADMAG || grad_iwe, "ebos_cmax_value_and_grad: GRADMAG needs the grad_iwe scratch plane");
  EBOS_CHECK_DTYPE(dtype, "ebos_cmax_value_and_grad");
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  EBOS_REQUIRE(!omit_boundary || (Hp > 2 && Wp > 2), "ebos_cmax_value_and_grad: omit_boundary needs an image larger than 2x2");
  cudaStream_t st = as_stream(stream);
  // one memset node for the accumulators AND the IWE when the caller laid them out back to back (ops.CmaxWorkspace does)
  const size_t iwe_bytes = (size_t)Hp * Wp * dtype_size(dtype);
  const bool adjacent = reinterpret_cast<char*>(acc) + EBOS_ACC_DOUBLES * sizeof(double) == reinterpret_cast<char*>(iwe);
  // clean_workspace: the caller guarantees acc and iwe are all zero on entry and gets them back all zero -- the zero-fill
  // for the NEXT evaluation then runs on the auxiliary lane concurrently with the backward instead of in front of the splat
  auto zero_workspace = [&](cudaStream_t s) -> cudaError_t {
    cudaError_t z = cudaMemsetAsync(acc, 0, EBOS_ACC_DOUBLES * sizeof(double) + (adjacent ? iwe_bytes : 0), s);
    if (z == cudaSuccess && !adjacent) z = cudaMemsetAsync(iwe, 0, iwe_bytes, s);
    return z;
  };
  if (!clean_workspace) {
    cudaError_t e = cudaMemsetAsync(acc, 0, EBOS_ACC_DOUBLES * sizeof(double) + (adjacent ? iwe_bytes : 0), st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_cmax_value_and_grad memset");
  }
  // fork: TV(flow) -> dflow on the auxiliary lane.  Small windows: concurrently with the splat (a 500 k-event splat is
  // ~120 CTAs and leaves most SMs idle).  Large windows: the splat fills the machine on its own and a concurrent TV kernel
  // only slows it down, so the fork comes AFTER the splat and the TV kernel shares the GPU with the (short, ramp-bound)
  // cost kernel instead.
  AuxLane* lane = aux_lane(st);
  const bool tv_after_splat = n >= kTvAfterSplatEvents;
  int rc;
  if (tv_after_splat) {
    rc = window_splat_launch(window, n, flags, flow, H, W, pad_h, pad_w, dtype, iwe, st, !adjacent && !clean_workspace);
    if (rc) return rc;
  }
  cudaStream_t tv_st = st;
  if (lane) {
    if (cudaEventRecord(lane->fork, st) == cudaSuccess && cudaStreamWaitEvent(lane->stream, lane->fork, 0) == cudaSuccess)
      tv_st = lane->stream;
    else
      lane = nullptr;
  }
  if (dtype == EBOS_F64)
    rc = flow_tv_t<double>((const double*)flow, (const double*)tv_weights, H, W, tv_scale, acc, (double*)dflow, tv_st);
  else
    rc = flow_tv_t<float>((const float*)flow, (const float*)tv_weights, H, W, tv_scale, acc, (float*)dflow, tv_st);
  if (lane && cudaEventRecord(lane->join, tv_st) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_value_and_grad(join)");
  if (rc) return rc;
  if (!tv_after_splat) {
    rc = window_splat_launch(window, n, flags, flow, H, W, pad_h, pad_w, dtype, iwe, st, !adjacent && !clean_workspace);
    if (rc) return rc;
  }
  void* gplane = nullptr;
  rc = cost_with_blur(kind, iwe, Hp, Wp, omit_boundary, data_scale, dtype, acc, grad_iwe, blur_sigma, blur_plane, st, &gplane);
  if (rc) return rc;
  // The scalar loss only needs the accumulators (complete once the cost kernel here and the TV kernel on the lane are
  // done), not the backward: it is computed on the lane, concurrently with the backward.
  auto finalize = [&](cudaStream_t s) {
    if (dtype == EBOS_F64)
      k_loss_finalize<double><<<1, 1, 0, s>>>(kind, acc, Hp, Wp, H, W, omit_boundary, data_scale, tv_scale, (double*)loss);
    else
      k_loss_finalize<float><<<1, 1, 0, s>>>(kind, acc, Hp, Wp, H, W, omit_boundary, data_scale, tv_scale, (float*)loss);
  };
  // (the gradient-magnitude backward reads neither the IWE nor the accumulators: their zero-fill follows the loss on the lane)
  const bool zero_on_lane = clean_workspace && (kind == EBOS_COST_GRADMAG || blur_sigma > 0.0);
  bool fin_on_lane = false;
  if (lane && cudaEventRecord(lane->cost_done, st) == cudaSuccess &&
      cudaStreamWaitEvent(lane->stream, lane->cost_done, 0) == cudaSuccess) {
    finalize(lane->stream);
    if (zero_on_lane && zero_workspace(lane->stream) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_value_and_grad(zero)");
    fin_on_lane = cudaEventRecord(lane->fin_done, lane->stream) == cudaSuccess;
    if (!fin_on_lane) return cuda_fail(cudaGetLastError(), "ebos_cmax_value_and_grad(fin)");
  }
  // join: the backward accumulates into the dflow the TV kernel wrote
  if (lane && cudaStreamWaitEvent(st, lane->join, 0) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_value_and_grad(wait)");
  rc = window_backward_launch(window, n, flags, flow, H, W, pad_h, pad_w, dtype, gplane, kind, iwe, acc,
                              omit_boundary, data_scale, dflow, st);
  if (rc) return rc;
  if (fin_on_lane) {
    if (cudaStreamWaitEvent(st, lane->fin_done, 0) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_value_and_grad(wait fin)");
  } else {
    finalize(st);
  }
  if (clean_workspace && !(fin_on_lane && zero_on_lane) && zero_workspace(st) != cudaSuccess)
    return cuda_fail(cudaGetLastError(), "ebos_cmax_value_and_grad(zero)");
  EBOS_LAUNCH_CHECK("ebos_cmax_value_and_grad");
  return EBOS_OK;
}

int ebos_cmax_adam_iteration(const void* window, int64_t n, int flags, void* flow, int H, int W, int pad_h, int pad_w,
                             int kind, int omit_boundary, double data_scale, double tv_scale, const void* tv_weights,
                             int dtype, void* iwe, void* grad_iwe, void* dflow, void* loss, double* acc, void* exp_avg,
                             void* exp_avg_sq, double lr, double beta1, double beta2, double eps, int32_t* step_dev,
                             double blur_sigma, void* blur_plane, void* stream) {
  EBOS_REQUIRE(blur_sigma >= 0.0, "ebos_cmax_adam_iteration: blur_sigma must not be negative");
  EBOS_REQUIRE(window && flow && iwe && dflow && loss && acc && exp_avg && exp_avg_sq && step_dev && n >= 0 && H > 1 &&
                   W > 1 && pad_h >= 0 && pad_w >= 0,
               "ebos_cmax_adam_iteration: bad argument");
  EBOS_REQUIRE(kind == EBOS_COST_VARIANCE || kind == EBOS_COST_GRADMAG, "ebos_cmax_adam_iteration: unknown cost kind");
  EBOS_REQUIRE(kind != EBOS_COST_GRADMAG || grad_iwe, "ebos_cmax_adam_iteration: GRADMAG needs the grad_iwe scratch plane");
  EBOS_CHECK_DTYPE(dtype, "ebos_cmax_adam_iteration");
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  EBOS_REQUIRE(!omit_boundary || (Hp > 2 && Wp > 2), "ebos_cmax_adam_iteration: omit_boundary needs an image larger than 2x2");
  cudaStream_t st = as_stream(stream);
  // graph nodes of one iteration: [TV + step++ + Adam factors | splat] [cost] [backward | IWE memset] [Adam + loss + acc reset]
  // (large windows: [splat] [TV + step++ | cost] ..., see ebos_cmax_value_and_grad)
  static const bool coefs_in_adam = getenv("EBOS_ADAM_OWN_COEFS") != nullptr;   // A/B: pow() + barrier inside k_adam as before
  const StepCoefs sc{step_dev, lr, beta1, beta2, coefs_in_adam ? nullptr : acc + kAccAdamCoefs};
  AuxLane* lane = aux_lane(st);
  const bool tv_after_splat = n >= kTvAfterSplatEvents;
  int rc;
  if (tv_after_splat) {
    rc = window_splat_launch(window, n, flags, flow, H, W, pad_h, pad_w, dtype, iwe, st, false);   // iwe is zero on entry
    if (rc) return rc;
  }
  cudaStream_t tv_st = st;
  if (lane) {
    if (cudaEventRecord(lane->fork, st) == cudaSuccess && cudaStreamWaitEvent(lane->stream, lane->fork, 0) == cudaSuccess)
      tv_st = lane->stream;
    else
      lane = nullptr;
  }
  if (dtype == EBOS_F64)
    rc = flow_tv_t<double>((const double*)flow, (const double*)tv_weights, H, W, tv_scale, acc, (double*)dflow, tv_st, sc);
  else
    rc = flow_tv_t<float>((const float*)flow, (const float*)tv_weights, H, W, tv_scale, acc, (float*)dflow, tv_st, sc);
  if (lane && cudaEventRecord(lane->join, tv_st) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_adam_iteration(join)");
  if (rc) return rc;
  if (!tv_after_splat) {
    rc = window_splat_launch(window, n, flags, flow, H, W, pad_h, pad_w, dtype, iwe, st, false);   // iwe is zero on entry
    if (rc) return rc;
  }
  void* gplane = nullptr;
  rc = cost_with_blur(kind, iwe, Hp, Wp, omit_boundary, data_scale, dtype, acc, grad_iwe, blur_sigma, blur_plane, st, &gplane);
  if (rc) return rc;
  // zero-fill of the IWE for the next iteration: the gradient-magnitude backward does not read the IWE, so it runs on
  // the lane concurrently with the backward (2 us off the critical path of a ~30 us iteration); the variance backward
  // derives dL/dIWE from the IWE, so there it follows the backward
  const size_t iwe_bytes = (size_t)Hp * Wp * dtype_size(dtype);
  bool zero_on_lane = false;
  if (lane && (kind == EBOS_COST_GRADMAG || blur_sigma > 0.0) && cudaEventRecord(lane->cost_done, st) == cudaSuccess &&
      cudaStreamWaitEvent(lane->stream, lane->cost_done, 0) == cudaSuccess) {
    if (cudaMemsetAsync(iwe, 0, iwe_bytes, lane->stream) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_adam_iteration(zero)");
    zero_on_lane = cudaEventRecord(lane->fin_done, lane->stream) == cudaSuccess;
    if (!zero_on_lane) return cuda_fail(cudaGetLastError(), "ebos_cmax_adam_iteration(zero event)");
  }
  if (lane && cudaStreamWaitEvent(st, lane->join, 0) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_adam_iteration(wait)");
  rc = window_backward_launch(window, n, flags, flow, H, W, pad_h, pad_w, dtype, gplane, kind, iwe, acc, omit_boundary,
                              data_scale, dflow, st);
  if (rc) return rc;
  if (!zero_on_lane && cudaMemsetAsync(iwe, 0, iwe_bytes, st) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_adam_iteration(zero)");
  const FinalizeArgs fin{1, kind, Hp, Wp, H, W, omit_boundary, data_scale, tv_scale, acc, loss};
  const int64_t np = (int64_t)2 * H * W;
  cudaError_t le;
  if (dtype == EBOS_F64)
    le = launch_pdl(k_adam<double>, dim3(adam_grid(np)), dim3(256), st, (double*)flow, (const double*)dflow, (double*)exp_avg,
                    (double*)exp_avg_sq, np, lr, beta1, beta2, eps, 0, (const int32_t*)step_dev, coefs_in_adam ? 2 : 3, fin,
                    (const double*)(acc + kAccAdamCoefs));
  else
    le = launch_pdl(k_adam<float>, dim3(adam_grid(np)), dim3(256), st, (float*)flow, (const float*)dflow, (float*)exp_avg,
                    (float*)exp_avg_sq, np, lr, beta1, beta2, eps, 0, (const int32_t*)step_dev, coefs_in_adam ? 2 : 3, fin,
                    (const double*)(acc + kAccAdamCoefs));
  if (le != cudaSuccess) return cuda_fail(le, "ebos_cmax_adam_iteration(adam)");
  if (zero_on_lane && cudaStreamWaitEvent(st, lane->fin_done, 0) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_adam_iteration(wait zero)");
  EBOS_LAUNCH_CHECK("ebos_cmax_adam_iteration");
  return EBOS_OK;
}

int ebos_cmax_adam_iteration_fused_tv(const void* window, int64_t n, int flags, const void* flow_in, void* flow_out, int H,
                                      int W, int pad_h, int pad_w, int kind, int omit_boundary, double data_scale,
                                      double tv_scale, int dtype, void* iwe, void* grad_iwe, void* dflow, void* loss,
                                      double* acc, void* exp_avg, void* exp_avg_sq, double lr, double beta1, double beta2,
                                      double eps, int32_t* step_dev, double blur_sigma, void* blur_plane, void* stream) {
  EBOS_REQUIRE(blur_sigma >= 0.0, "ebos_cmax_adam_iteration_fused_tv: blur_sigma must not be negative");
  EBOS_REQUIRE(window && flow_in && flow_out && iwe && dflow && loss && acc && exp_avg && exp_avg_sq && step_dev && n >= 0 &&
                   H > 1 && W > 1 && pad_h >= 0 && pad_w >= 0 && flow_in != flow_out,
               "ebos_cmax_adam_iteration_fused_tv: bad argument");
  EBOS_REQUIRE(kind == EBOS_COST_VARIANCE || kind == EBOS_COST_GRADMAG, "ebos_cmax_adam_iteration_fused_tv: unknown cost kind");
  EBOS_REQUIRE(kind != EBOS_COST_GRADMAG || grad_iwe, "ebos_cmax_adam_iteration_fused_tv: GRADMAG needs the grad_iwe scratch plane");
  const size_t align = reinterpret_cast<size_t>(flow_in) | reinterpret_cast<size_t>(flow_out) | reinterpret_cast<size_t>(dflow) |
                       reinterpret_cast<size_t>(exp_avg) | reinterpret_cast<size_t>(exp_avg_sq);
  if (dtype != EBOS_F32 || (W & 3) || W < 12 || H < 5 || (align & 15)) {
    set_error("ebos_cmax_adam_iteration_fused_tv: needs fp32, W % 4 == 0, W >= 12, H >= 5 and 16-byte aligned planes "
              "(use ebos_cmax_adam_iteration)");
    return EBOS_ERR_UNSUPPORTED;
  }
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  EBOS_REQUIRE(!omit_boundary || (Hp > 2 && Wp > 2), "ebos_cmax_adam_iteration_fused_tv: omit_boundary needs an image larger than 2x2");
  cudaStream_t st = as_stream(stream);
  // graph nodes of one iteration: [splat] [cost] [backward | IWE memset] [Adam + TV + loss + accumulator reset + step++]
  AuxLane* lane = aux_lane(st);
  int rc = window_splat_launch(window, n, flags, flow_in, H, W, pad_h, pad_w, dtype, iwe, st, false);   // iwe is zero on entry
  if (rc) return rc;
  void* gplane = nullptr;
  rc = cost_with_blur(kind, iwe, Hp, Wp, omit_boundary, data_scale, dtype, acc, grad_iwe, blur_sigma, blur_plane, st, &gplane);
  if (rc) return rc;
  const size_t iwe_bytes = (size_t)Hp * Wp * sizeof(float);
  bool zero_on_lane = false;
  if (lane && (kind == EBOS_COST_GRADMAG || blur_sigma > 0.0) && cudaEventRecord(lane->cost_done, st) == cudaSuccess &&
      cudaStreamWaitEvent(lane->stream, lane->cost_done, 0) == cudaSuccess) {
    if (cudaMemsetAsync(iwe, 0, iwe_bytes, lane->stream) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_adam_iteration_fused_tv(zero)");
    zero_on_lane = cudaEventRecord(lane->fin_done, lane->stream) == cudaSuccess;
    if (!zero_on_lane) return cuda_fail(cudaGetLastError(), "ebos_cmax_adam_iteration_fused_tv(zero event)");
  }
  rc = window_backward_launch(window, n, flags, flow_in, H, W, pad_h, pad_w, dtype, gplane, kind, iwe, acc, omit_boundary,
                              data_scale, dflow, st);
  if (rc) return rc;
  if (!zero_on_lane && cudaMemsetAsync(iwe, 0, iwe_bytes, st) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_adam_iteration_fused_tv(zero)");
  const FinalizeArgs fin{1, kind, Hp, Wp, H, W, omit_boundary, data_scale, tv_scale, acc, loss};
  const float coef = (float)(tv_scale / (2.0 * (double)H * (double)W));
  const int64_t n_frame = (int64_t)4 * W + (int64_t)8 * (H - 4);
  const int n_frame_ctas = (int)((n_frame + 127) / 128);
  const int fast_gx = ((W >> 2) + 63) / 64, fast_gy = (H - 4 + 2 * ADAM_ROWS - 1) / (2 * ADAM_ROWS);
  cudaError_t le = launch_pdl(k_adam_tv_march, dim3(n_frame_ctas + fast_gx * fast_gy, 2), dim3(128), st, (const float*)flow_in,
                              (float*)flow_out, (float*)dflow, (float*)exp_avg, (float*)exp_avg_sq, H, W, coef, lr, beta1, beta2,
                              eps, step_dev, fin, n_frame_ctas, fast_gx);
  if (le != cudaSuccess) return cuda_fail(le, "ebos_cmax_adam_iteration_fused_tv(adam)");
  if (zero_on_lane && cudaStreamWaitEvent(st, lane->fin_done, 0) != cudaSuccess) return cuda_fail(cudaGetLastError(), "ebos_cmax_adam_iteration_fused_tv(wait zero)");
  EBOS_LAUNCH_CHECK("ebos_cmax_adam_iteration_fused_tv");
  return EBOS_OK;
}

int ebos_adam_step(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, int64_t n, double lr, double beta1,
                   double beta2, double eps, int step, int dtype, void* stream) {
  EBOS_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "ebos_adam_step: bad argument");
  EBOS_CHECK_DTYPE(dtype, "ebos_adam_step");
  if (n == 0) return EBOS_OK;
  if (dtype == EBOS_F64)
    k_adam<double><<<adam_grid(n), 256, 0, as_stream(stream)>>>((double*)param, (const double*)grad, (double*)exp_avg,
                                                                 (double*)exp_avg_sq, n, lr, beta1, beta2, eps, step, nullptr, 0, FinalizeArgs{});
  else
    k_adam<float><<<adam_grid(n), 256, 0, as_stream(stream)>>>((float*)param, (const float*)grad, (float*)exp_avg,
                                                                (float*)exp_avg_sq, n, lr, beta1, beta2, eps, step, nullptr, 0, FinalizeArgs{});
  EBOS_LAUNCH_CHECK("ebos_adam_step");
  return EBOS_OK;
}

int ebos_adam_step_graph(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, int64_t n, double lr,
                         double beta1, double beta2, double eps, int32_t* step_dev, int dtype, void* stream) {
  EBOS_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_dev && n >= 0, "ebos_adam_step_graph: bad argument");
  EBOS_CHECK_DTYPE(dtype, "ebos_adam_step_graph");
  cudaStream_t st = as_stream(stream);
  if (n > 0) {
    if (dtype == EBOS_F64)
      k_adam<double><<<adam_grid(n), 256, 0, st>>>((double*)param, (const double*)grad, (double*)exp_avg,
                                                    (double*)exp_avg_sq, n, lr, beta1, beta2, eps, 0, step_dev, 1, FinalizeArgs{});
    else
      k_adam<float><<<adam_grid(n), 256, 0, st>>>((float*)param, (const float*)grad, (float*)exp_avg,
                                                   (float*)exp_avg_sq, n, lr, beta1, beta2, eps, 0, step_dev, 1, FinalizeArgs{});
  }
  k_adam_bump<<<1, 1, 0, st>>>(step_dev);
  EBOS_LAUNCH_CHECK("ebos_adam_step_graph");
  return EBOS_OK;
}

}  // extern "C"
