// Objectives on the IWE / flow planes, Adam, and the fused per-iteration entry point.
//
//   variance            L = -var(IWE) (unbiased)                    SURVEY.md A.4 (not upstream)
//   gradient magnitude  L = -mean((Sx/8)^2 + (Sy/8)^2), Sobel 3x3 with replicate padding, kernels of
//                       src/utils/stat_utils.py:62-139                SURVEY.md A.4 (not upstream)
//   total variation     ImageGradient.calculate_torch, src/costs/image_gradient.py:60-75
//   Adam                torch.optim.Adam as used at src/solver/patch_eklt_pyramid2.py:262-264,284
//
// All reductions: per-thread double accumulators -> warp shuffle -> one atomicAdd(double) per CTA.
// acc layout (double[8]): [0] sum(IWE) [1] sum(IWE^2) [2] sum(gx^2+gy^2) [3] sum(|TV terms|).
#include <algorithm>

#include "ebos_common.cuh"

namespace ebos {

int window_splat_launch(const void* window, int64_t n, int has_weight, const float* flow, int H, int W, int pad_h,
                        int pad_w, float* iwe, cudaStream_t st);
int window_backward_launch(const void* window, int64_t n, int has_weight, const float* flow, int H, int W, int pad_h,
                           int pad_w, const float* grad_iwe, int kind, const float* iwe, const double* acc,
                           int omit_boundary, float scale, float* dflow, cudaStream_t st);

// ---- variance: one pass, sum and sum of squares in double --------------------------------------
__global__ void __launch_bounds__(256) k_var_reduce(const float* __restrict__ iwe, int Hp, int Wp, int omit,
                                                    double* __restrict__ acc) {
  __shared__ double sm[32];
  const int r_lo = omit ? 1 : 0, r_hi = omit ? Hp - 1 : Hp;
  const int c_lo = omit ? 1 : 0, c_hi = omit ? Wp - 1 : Wp;
  const int64_t rows = r_hi - r_lo, cols = c_hi - c_lo;
  const int64_t total = rows > 0 && cols > 0 ? rows * cols : 0;
  double s = 0.0, q = 0.0;
  if (!omit && (Wp & 3) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(iwe);
    const int64_t n4 = total >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      float4 v = __ldg(p4 + i);
      s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
      q += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = r_lo + (int)(i / cols), c = c_lo + (int)(i % cols);
      const double v = (double)__ldg(iwe + (int64_t)r * Wp + c);
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
__global__ void __launch_bounds__(256) k_var_grad(const float* __restrict__ iwe, int Hp, int Wp, int omit, float scale,
                                                  const double* __restrict__ acc, float* __restrict__ g) {
  const double cnt = omit ? (double)(Hp - 2) * (double)(Wp - 2) : (double)Hp * (double)Wp;
  const float mean = (float)(acc[0] / cnt);
  const float cv = (float)(-2.0 * (double)scale / (cnt - 1.0));
  const int64_t total = (int64_t)Hp * Wp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / Wp), c = (int)(i % Wp);
    const bool border = r == 0 || c == 0 || r == Hp - 1 || c == Wp - 1;
    g[i] = (omit && border) ? 0.f : cv * (__ldg(iwe + i) - mean);
  }
}

// ---- gradient magnitude: Sobel forward + adjoint in one tiled pass ---------------------------------
constexpr int GT = 32;  // output tile edge
__global__ void __launch_bounds__(256) k_gradmag(const float* __restrict__ iwe, int Hp, int Wp, int omit, float coef,
                                                 double* __restrict__ acc, float* __restrict__ g) {
  // coef = -2 * scale / (8 * n_el): D = coef * gx is dL/d(Sx) including the forward's 1/8.
  __shared__ float sI[GT + 4][GT + 4 + 1];
  __shared__ float sDx[GT + 2][GT + 2 + 1];
  __shared__ float sDy[GT + 2][GT + 2 + 1];
  __shared__ double sm[32];
  const int r0 = blockIdx.y * GT, c0 = blockIdx.x * GT;
  // stage 1: image tile with a 2-pixel halo, replicate padding materialised by clamping
  for (int i = threadIdx.x; i < (GT + 4) * (GT + 4); i += blockDim.x) {
    const int lr = i / (GT + 4), lc = i % (GT + 4);
    const int r = min(max(r0 - 2 + lr, 0), Hp - 1), c = min(max(c0 - 2 + lc, 0), Wp - 1);
    sI[lr][lc] = __ldg(iwe + (int64_t)r * Wp + c);
  }
  __syncthreads();
  // stage 2: Sobel/8 on the tile + 1-pixel halo; positions outside the image (or outside the
  // omit_boundary crop) carry no gradient.
  double part = 0.0;
  for (int i = threadIdx.x; i < (GT + 2) * (GT + 2); i += blockDim.x) {
    const int lr = i / (GT + 2), lc = i % (GT + 2);
    const int r = r0 - 1 + lr, c = c0 - 1 + lc;
    float dx = 0.f, dy = 0.f;
    const bool inside = r >= 0 && r < Hp && c >= 0 && c < Wp;
    const bool counted = inside && !(omit && (r == 0 || c == 0 || r == Hp - 1 || c == Wp - 1));
    if (counted) {
      // sI index of pixel (r,c) is [lr+1][lc+1]
      const float i00 = sI[lr][lc], i01 = sI[lr][lc + 1], i02 = sI[lr][lc + 2];
      const float i10 = sI[lr + 1][lc], i12 = sI[lr + 1][lc + 2];
      const float i20 = sI[lr + 2][lc], i21 = sI[lr + 2][lc + 1], i22 = sI[lr + 2][lc + 2];
      const float gx = ((i20 - i00) + 2.f * (i21 - i01) + (i22 - i02)) * 0.125f;  // d/drow
      const float gy = ((i02 - i00) + 2.f * (i12 - i10) + (i22 - i20)) * 0.125f;  // d/dcol
      dx = coef * gx;
      dy = coef * gy;
      if (lr >= 1 && lr <= GT && lc >= 1 && lc <= GT) part += (double)gx * gx + (double)gy * gy;
    }
    sDx[lr][lc] = dx;
    sDy[lr][lc] = dy;
  }
  __syncthreads();
  // stage 3: adjoint.  dI[p] = sum over padded positions pp that clamp to p, over the 3x3 window:
  //   Kx[u][v]*Dx[pp-(u,v)] + Ky[u][v]*Dy[pp-(u,v)],   with D = 0 outside the image.
  for (int i = threadIdx.x; i < GT * GT; i += blockDim.x) {
    const int lr = i / GT, lc = i % GT;
    const int r = r0 + lr, c = c0 + lc;
    if (r >= Hp || c >= Wp) continue;
    float out = 0.f;
    for (int pr = (r == 0 ? -1 : r); pr <= (r == Hp - 1 ? Hp : r); ++pr) {
      for (int pc = (c == 0 ? -1 : c); pc <= (c == Wp - 1 ? Wp : c); ++pc) {
#pragma unroll
        for (int u = -1; u <= 1; ++u) {
#pragma unroll
          for (int v = -1; v <= 1; ++v) {
            const int qr = pr - u, qc = pc - v;
            if (qr < 0 || qr >= Hp || qc < 0 || qc >= Wp) continue;
            // Kx[u][v] = u * (2 - |v|),  Ky[u][v] = v * (2 - |u|)
            const float kx = (float)(u * (2 - (v < 0 ? -v : v)));
            const float ky = (float)(v * (2 - (u < 0 ? -u : u)));
            const int sr = qr - (r0 - 1), sc = qc - (c0 - 1);
            out += kx * sDx[sr][sc] + ky * sDy[sr][sc];
          }
        }
      }
    }
    g[(int64_t)r * Wp + c] = out;
  }
  part = block_sum(part, sm);
  if (threadIdx.x == 0) atomicAdd(acc + 2, part);
}

// ---- total variation of the flow: value + gradient -------------------------------------------------
// torch.gradient (spacing 1, edge_order 1): interior (f[i+1]-f[i-1])/2, edges one-sided.
__device__ __forceinline__ float grad1d(const float* __restrict__ f, int i, int n, int64_t stride) {
  if (i == 0) return f[stride] - f[0];
  if (i == n - 1) return f[(int64_t)(n - 1) * stride] - f[(int64_t)(n - 2) * stride];
  return (f[(int64_t)(i + 1) * stride] - f[(int64_t)(i - 1) * stride]) * 0.5f;
}
__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

// adjoint of grad1d along one axis at index i: sum_q s(q) * d grad(q) / d f[i]
template <typename S>
__device__ __forceinline__ float grad1d_adjoint(int i, int n, S s) {
  float out = 0.f;
  if (i + 1 <= n - 2) out -= 0.5f * s(i + 1);          // interior q = i+1
  if (i - 1 >= 1) out += 0.5f * s(i - 1);              // interior q = i-1
  if (i == 0) out -= s(0);
  if (i == 1) out += s(0);
  if (i == n - 1) out += s(n - 1);
  if (i == n - 2) out -= s(n - 1);
  return out;
}

__global__ void __launch_bounds__(256) k_flow_tv(const float* __restrict__ flow, const float* __restrict__ weights, int H,
                                                 int W, float coef, double* __restrict__ acc, float* __restrict__ dflow) {
  // coef = tv_scale / (2*H*W)
  __shared__ double sm[32];
  const int64_t hw = (int64_t)H * W;
  double part = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * hw; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i / hw);
    const int64_t p = i - (int64_t)ch * hw;
    const int r = (int)(p / W), c = (int)(p % W);
    const float* f = flow + (int64_t)ch * hw;
    auto wgt = [&](int rr, int cc) { return weights ? __ldg(weights + (int64_t)rr * W + cc) : 1.f; };
    auto s_row = [&](int q) { const float w = wgt(q, c); return sgn(grad1d(f + c, q, H, W) * w) * w; };
    auto s_col = [&](int q) { const float w = wgt(r, q); return sgn(grad1d(f + (int64_t)r * W, q, W, 1) * w) * w; };
    const float w_here = wgt(r, c);
    part += (double)fabsf(grad1d(f + c, r, H, W) * w_here) + (double)fabsf(grad1d(f + (int64_t)r * W, c, W, 1) * w_here);
    dflow[i] = coef * (grad1d_adjoint(r, H, s_row) + grad1d_adjoint(c, W, s_col));
  }
  part = block_sum(part, sm);
  if (threadIdx.x == 0 && acc) atomicAdd(acc + 3, part);
}

// ---- loss scalar -------------------------------------------------------------------------------------
__global__ void k_loss_finalize(int kind, const double* __restrict__ acc, int Hp, int Wp, int H, int W, int omit,
                                float data_scale, float tv_scale, float* __restrict__ loss) {
  const double cnt = omit ? (double)(Hp - 2) * (double)(Wp - 2) : (double)Hp * (double)Wp;
  double data = 0.0;
  if (kind == EBOS_COST_VARIANCE) data = -((acc[1] - acc[0] * acc[0] / cnt) / (cnt - 1.0));
  else if (kind == EBOS_COST_GRADMAG) data = -(acc[2] / cnt);
  const double tv = acc[3] / (2.0 * (double)H * (double)W);
  loss[0] = (float)((double)data_scale * data + (double)tv_scale * tv);
}

// ---- Adam ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float b1, float b2, float eps,
                                         float step_size, float inv_bc2_sqrt) {
  m = m * b1 + (1.f - b1) * g;
  v = v * b2 + (1.f - b2) * g * g;
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                              int step_host, int32_t* __restrict__ step_dev) {
  int step = step_host;
  if (step_dev) step = *step_dev + 1;  // every thread reads the pre-increment value (bumped by k_adam_bump)
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const int64_t n4 = n >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if ((((size_t)p | (size_t)g | (size_t)m | (size_t)v) & 15) == 0) {
    for (int64_t i = tid; i < n4; i += nth) {
      float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
      adam_one(pp.x, gg.x, mm.x, vv.x, b1, b2, eps, step_size, inv_bc2_sqrt);
      adam_one(pp.y, gg.y, mm.y, vv.y, b1, b2, eps, step_size, inv_bc2_sqrt);
      adam_one(pp.z, gg.z, mm.z, vv.z, b1, b2, eps, step_size, inv_bc2_sqrt);
      adam_one(pp.w, gg.w, mm.w, vv.w, b1, b2, eps, step_size, inv_bc2_sqrt);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += nth) adam_one(p[i], g[i], m[i], v[i], b1, b2, eps, step_size, inv_bc2_sqrt);
  } else {
    for (int64_t i = tid; i < n; i += nth) adam_one(p[i], g[i], m[i], v[i], b1, b2, eps, step_size, inv_bc2_sqrt);
  }
}
__global__ void k_adam_bump(int32_t* step_dev) { *step_dev += 1; }

// ---- launch helpers ----------------------------------------------------------------------------------
static int plane_grid(int64_t elems, int per_thread = 4) {
  int64_t blocks = (elems / per_thread + 255) / 256;
  return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)sm_count() * 8));
}

int iwe_cost_launch(int kind, const float* iwe, int Hp, int Wp, int omit, float scale, double* acc, float* grad_iwe,
                    cudaStream_t st) {
  const double cnt = omit ? (double)(Hp - 2) * (double)(Wp - 2) : (double)Hp * (double)Wp;
  if (kind == EBOS_COST_VARIANCE) {
    k_var_reduce<<<plane_grid((int64_t)Hp * Wp), 256, 0, st>>>(iwe, Hp, Wp, omit, acc);
    if (grad_iwe) k_var_grad<<<plane_grid((int64_t)Hp * Wp), 256, 0, st>>>(iwe, Hp, Wp, omit, scale, acc, grad_iwe);
  } else if (kind == EBOS_COST_GRADMAG) {
    if (!grad_iwe) { set_error("ebos_iwe_cost: GRADMAG needs grad_iwe"); return EBOS_ERR_BAD_ARG; }
    const float coef = (float)(-2.0 * (double)scale / (8.0 * cnt));
    dim3 grid((Wp + GT - 1) / GT, (Hp + GT - 1) / GT);
    k_gradmag<<<grid, 256, 0, st>>>(iwe, Hp, Wp, omit, coef, acc, grad_iwe);
  } else if (kind != EBOS_COST_NONE) {
    set_error("ebos_iwe_cost: unknown cost kind");
    return EBOS_ERR_BAD_ARG;
  }
  EBOS_LAUNCH_CHECK("ebos_iwe_cost");
  return EBOS_OK;
}

int flow_tv_launch(const float* flow, const float* weights, int H, int W, float tv_scale, double* acc, float* dflow,
                   cudaStream_t st) {
  if (tv_scale == 0.f || H < 2 || W < 2) {
    cudaError_t e = cudaMemsetAsync(dflow, 0, (size_t)2 * H * W * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_flow_tv memset");
    if (tv_scale == 0.f) return EBOS_OK;
    set_error("ebos_flow_tv: torch.gradient needs at least 2 samples per axis");
    return EBOS_ERR_BAD_ARG;
  }
  const float coef = (float)((double)tv_scale / (2.0 * (double)H * (double)W));
  k_flow_tv<<<plane_grid((int64_t)2 * H * W, 1), 256, 0, st>>>(flow, weights, H, W, coef, acc, dflow);
  EBOS_LAUNCH_CHECK("ebos_flow_tv");
  return EBOS_OK;
}

}  // namespace ebos

using namespace ebos;

extern "C" {

int ebos_iwe_cost(int kind, const float* iwe, int Hp, int Wp, int omit_boundary, float scale, double* acc,
                  float* grad_iwe, void* stream) {
  EBOS_REQUIRE(iwe && acc && Hp > 0 && Wp > 0, "ebos_iwe_cost: bad argument");
  EBOS_REQUIRE(!omit_boundary || (Hp > 2 && Wp > 2), "ebos_iwe_cost: omit_boundary needs an image larger than 2x2");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(acc, 0, 3 * sizeof(double), st);
  if (e != cudaSuccess) return cuda_fail(e, "ebos_iwe_cost memset");
  return iwe_cost_launch(kind, iwe, Hp, Wp, omit_boundary, scale, acc, grad_iwe, st);
}

int ebos_flow_tv(const float* flow, const float* weights, int H, int W, float tv_scale, double* acc, float* dflow,
                 void* stream) {
  EBOS_REQUIRE(flow && dflow && H > 0 && W > 0, "ebos_flow_tv: bad argument");
  cudaStream_t st = as_stream(stream);
  if (acc) {
    cudaError_t e = cudaMemsetAsync(acc + 3, 0, sizeof(double), st);
    if (e != cudaSuccess) return cuda_fail(e, "ebos_flow_tv memset");
  }
  return flow_tv_launch(flow, weights, H, W, tv_scale, acc, dflow, st);
}

int ebos_loss_finalize(int kind, const double* acc, int Hp, int Wp, int H, int W, int omit_boundary, float data_scale,
                       float tv_scale, float* loss, void* stream) {
  EBOS_REQUIRE(acc && loss, "ebos_loss_finalize: bad argument");
  k_loss_finalize<<<1, 1, 0, as_stream(stream)>>>(kind, acc, Hp, Wp, H, W, omit_boundary, data_scale, tv_scale, loss);
  EBOS_LAUNCH_CHECK("ebos_loss_finalize");
  return EBOS_OK;
}

int ebos_cmax_value_and_grad(const void* window, int64_t n, int has_weight, const float* flow, int H, int W, int pad_h,
                             int pad_w, int kind, int omit_boundary, float data_scale, float tv_scale,
                             const float* tv_weights, float* iwe, float* grad_iwe, float* dflow, float* loss, double* acc,
                             void* stream) {
  EBOS_REQUIRE(window && flow && iwe && dflow && loss && acc && n >= 0 && H > 0 && W > 0 && pad_h >= 0 && pad_w >= 0,
               "ebos_cmax_value_and_grad: bad argument");
  EBOS_REQUIRE(kind == EBOS_COST_VARIANCE || kind == EBOS_COST_GRADMAG, "ebos_cmax_value_and_grad: unknown cost kind");
  EBOS_REQUIRE(kind != EBOS_COST_GRADMAG || grad_iwe, "ebos_cmax_value_and_grad: GRADMAG needs the grad_iwe scratch plane");
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  EBOS_REQUIRE(!omit_boundary || (Hp > 2 && Wp > 2), "ebos_cmax_value_and_grad: omit_boundary needs an image larger than 2x2");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(acc, 0, 8 * sizeof(double), st);
  if (e != cudaSuccess) return cuda_fail(e, "ebos_cmax_value_and_grad memset");
  int rc = window_splat_launch(window, n, has_weight, flow, H, W, pad_h, pad_w, iwe, st);
  if (rc) return rc;
  // variance: no gradient plane, the backward derives it from (iwe, acc)
  rc = iwe_cost_launch(kind, iwe, Hp, Wp, omit_boundary, data_scale, acc, kind == EBOS_COST_GRADMAG ? grad_iwe : nullptr, st);
  if (rc) return rc;
  rc = flow_tv_launch(flow, tv_weights, H, W, tv_scale, acc, dflow, st);
  if (rc) return rc;
  rc = window_backward_launch(window, n, has_weight, flow, H, W, pad_h, pad_w,
                              kind == EBOS_COST_GRADMAG ? grad_iwe : nullptr, kind, iwe, acc, omit_boundary, data_scale,
                              dflow, st);
  if (rc) return rc;
  k_loss_finalize<<<1, 1, 0, st>>>(kind, acc, Hp, Wp, H, W, omit_boundary, data_scale, tv_scale, loss);
  EBOS_LAUNCH_CHECK("ebos_cmax_value_and_grad");
  return EBOS_OK;
}

int ebos_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, int step, void* stream) {
  EBOS_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "ebos_adam_step: bad argument");
  if (n == 0) return EBOS_OK;
  k_adam<<<plane_grid(n), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, nullptr);
  EBOS_LAUNCH_CHECK("ebos_adam_step");
  return EBOS_OK;
}

int ebos_adam_step_graph(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                         float beta1, float beta2, float eps, int32_t* step_dev, void* stream) {
  EBOS_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_dev && n >= 0, "ebos_adam_step_graph: bad argument");
  cudaStream_t st = as_stream(stream);
  if (n > 0) k_adam<<<plane_grid(n), 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, 0, step_dev);
  k_adam_bump<<<1, 1, 0, st>>>(step_dev);
  EBOS_LAUNCH_CHECK("ebos_adam_step_graph");
  return EBOS_OK;
}

}  // extern "C"
