// Plane kernels either side of the hot path, fp32 / fp64:
//
//   flow-error metrics   calculate_flow_error_numpy, src/utils/flow_utils.py:769-821 (EPE, N-pixel outlier ratios, AE
//                        with the event mask of src/solver/base.py:289-317): one reduction pass, sums in double
//   3x3 Gaussian blur    torchvision gaussian_blur(kernel_size=3, sigma) on the IWE, src/event_image_converter.py:399-404
//                        (reflect padding, kernel2d = outer(k, k) with k = exp(-x^2 / 2 sigma^2) / sum, x = -1, 0, 1),
//                        forward and exact adjoint (the reflect border makes the operator non-symmetric)
#include <cmath>

#include "ebos_common.cuh"

namespace ebos {

// ---- flow error -------------------------------------------------------------------------------------------
// per batch row: [0] n_points  [1] sum EPE  [2..7] counts EPE > 1, 2, 3, 5, 10, 20  [8] sum AE
constexpr int kErrSlots = 9;

template <typename T>
__global__ void __launch_bounds__(256) k_flow_error(const T* __restrict__ gt, const T* __restrict__ pred,
                                                    const uint8_t* __restrict__ mask, int64_t mask_batch_stride,
                                                    const T* __restrict__ time_scale, int64_t hw, double* __restrict__ sums) {
  __shared__ double red[32];
  const int b = blockIdx.y;
  const T* g0 = gt + (int64_t)b * 2 * hw;
  const T* g1 = g0 + hw;
  const T* p0 = pred + (int64_t)b * 2 * hw;
  const T* p1 = p0 + hw;
  const uint8_t* m = mask ? mask + (int64_t)b * mask_batch_stride : nullptr;
  const bool scaled = time_scale != nullptr;
  const T ts = scaled ? time_scale[b] : (T)1;
  double n = 0.0, epe = 0.0, ae = 0.0;
  int c1 = 0, c2 = 0, c3 = 0, c5 = 0, c10 = 0, c20 = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    const T a0 = __ldg(g0 + i), a1 = __ldg(g1 + i);
    // valid in the ground truth: not +-inf and non-zero in BOTH channels (a NaN fails |x| > 0), and inside the event mask
    bool ok = !isinf(a0) && !isinf(a1) && fabs(a0) > (T)0 && fabs(a1) > (T)0;
    if (m) ok = ok && m[i] != 0;
    const T w = ok ? (T)1 : (T)0;
    // the reference MULTIPLIES by the mask: an infinite value at a masked-out pixel becomes NaN and poisons the sums
    T ug = Rn<T>::mul(a0, w), vg = Rn<T>::mul(a1, w);
    T u = Rn<T>::mul(__ldg(p0 + i), w), v = Rn<T>::mul(__ldg(p1 + i), w);
    if (scaled) { ug = Rn<T>::mul(ug, ts); vg = Rn<T>::mul(vg, ts); u = Rn<T>::mul(u, ts); v = Rn<T>::mul(v, ts); }   // tensor variant, :742-745
    const T du = Rn<T>::sub(ug, u), dv = Rn<T>::sub(vg, v);
    const T e = sqrt(Rn<T>::add(Rn<T>::mul(du, du), Rn<T>::mul(dv, dv)));
    n += ok ? 1.0 : 0.0;
    epe += (double)e;
    c1 += e > (T)1; c2 += e > (T)2; c3 += e > (T)3; c5 += e > (T)5; c10 += e > (T)10; c20 += e > (T)20;
    const T num = Rn<T>::add(Rn<T>::add((T)1, Rn<T>::mul(u, ug)), Rn<T>::mul(v, vg));
    const T den = Rn<T>::mul(sqrt(Rn<T>::add(Rn<T>::add((T)1, Rn<T>::mul(u, u)), Rn<T>::mul(v, v))),
                             sqrt(Rn<T>::add(Rn<T>::add((T)1, Rn<T>::mul(ug, ug)), Rn<T>::mul(vg, vg))));
    ae += (double)acos(Rn<T>::div(num, den));   // a cosine above 1 by rounding gives NaN, as in numpy
  }
  double* out = sums + (int64_t)b * kErrSlots;
  const double vals[kErrSlots] = {n, epe, (double)c1, (double)c2, (double)c3, (double)c5, (double)c10, (double)c20, ae};
#pragma unroll
  for (int k = 0; k < kErrSlots; ++k) {
    const double t = block_sum(vals[k], red);
    if (threadIdx.x == 0) atomicAdd(out + k, t);
  }
}

// errors[k] = mean over the batch of sum_k / (n_points + 1e-5); order EPE, 1PE, 2PE, 3PE, 5PE, 10PE, 20PE, AE
// n_fp32 (the tensor variant): n_points is an int64 tensor + 1e-5, which torch types as FLOAT32 (:741); the outlier
// counts are int64 tensors, so their ratios and the batch mean are float32 as well, while EPE and AE stay in the flow
// dtype divided by that float32 count.
__global__ void k_flow_error_final(const double* __restrict__ sums, int batch, int n_fp32, double* __restrict__ errors) {
  const int k = threadIdx.x;
  if (k >= 8) return;
  const bool ratio32 = n_fp32 && k >= 1 && k <= 6;
  double acc = 0.0;
  float acc32 = 0.f;
  for (int b = 0; b < batch; ++b) {
    const double cnt = sums[b * kErrSlots], v = sums[b * kErrSlots + 1 + k];
    if (ratio32) acc32 = __fadd_rn(acc32, __fdiv_rn((float)v, __fadd_rn((float)cnt, 1e-5f)));
    else acc += v / (n_fp32 ? (double)__fadd_rn((float)cnt, 1e-5f) : cnt + 1e-5);
  }
  errors[k] = ratio32 ? (double)__fdiv_rn(acc32, (float)batch) : acc / (double)batch;
}

// ---- 3x3 Gaussian blur with reflect padding ------------------------------------------------------------------
template <typename T> struct Blur3 { T k2[3][3]; };   // kernel2d[a][b] = k[a] * k[b], rounded in T like torch.mm does

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// forward: out[r,c] = sum_ab k2[a][b] * in[refl(r+a-1), refl(c+b-1)]
// adjoint: out[r,c] = sum_ab mult_r(a) * mult_c(b) * k2[a][b] * in[r+a-1, c+b-1] over in-range neighbours, where the
//          multiplicity counts how often the forward reads pixel r from row r+a-1: once directly, and once more through
//          the reflection when r is next to the border (row 1 is also what row 0 sees at "-1"; row n-2 what row n-1 sees at "n")
template <typename T, bool ADJOINT>
__global__ void __launch_bounds__(256) k_blur3(const T* __restrict__ in, int H, int W, Blur3<T> kw, T* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= W) return;
  const T* src = in + (int64_t)blockIdx.z * H * W;
  T acc = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int rr = r + a - 1, cc = c + b - 1;
      if constexpr (!ADJOINT) {
        acc += kw.k2[a][b] * __ldg(src + (int64_t)reflect1(rr, H) * W + reflect1(cc, W));
      } else {
        if (rr < 0 || rr >= H || cc < 0 || cc >= W) continue;
        // output row rr reads input row r with tap index (2 - a) directly, and with tap index a through the mirror
        int mr = 1, mc = 1;
        if (a == 0 && r == 1) mr = 2;            // rr = 0 also reads row 1 as its "-1" neighbour
        if (a == 2 && r == H - 2) mr = 2;        // rr = H-1 also reads row H-2 as its "H" neighbour
        if (b == 0 && c == 1) mc = 2;
        if (b == 2 && c == W - 2) mc = 2;
        acc += (T)(mr * mc) * kw.k2[2 - a][2 - b] * __ldg(src + (int64_t)rr * W + cc);
      }
    }
  }
  out[(int64_t)blockIdx.z * H * W + (int64_t)r * W + c] = acc;
}

template <typename T>
static Blur3<T> make_blur3(double sigma) {
  // torchvision _get_gaussian_kernel1d in the image dtype: x = linspace(-1, 1, 3); pdf = exp(-0.5 * (x / sigma)^2);
  // k = pdf / pdf.sum(); kernel2d = k[:, None] * k[None, :]
  T pdf[3], k[3];
  const T s = (T)sigma;
  for (int i = 0; i < 3; ++i) {
    const T x = (T)(i - 1) / s;
    pdf[i] = (T)std::exp((T)(-0.5) * (x * x));
  }
  const T sum = (pdf[0] + pdf[1]) + pdf[2];
  for (int i = 0; i < 3; ++i) k[i] = pdf[i] / sum;
  Blur3<T> w;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) w.k2[a][b] = k[a] * k[b];
  return w;
}

template <typename T>
int blur3_t(const T* in, int batch, int H, int W, double sigma, int adjoint, T* out, cudaStream_t st) {
  const Blur3<T> kw = make_blur3<T>(sigma);
  const dim3 grid((W + 255) / 256, H, batch);
  if (adjoint) k_blur3<T, true><<<grid, 256, 0, st>>>(in, H, W, kw, out);
  else k_blur3<T, false><<<grid, 256, 0, st>>>(in, H, W, kw, out);
  EBOS_LAUNCH_CHECK("ebos_blur3");
  return EBOS_OK;
}

// type-erased launcher for the fused entries (ebos_costs.cu)
int blur3_launch(const void* in, int H, int W, double sigma, int adjoint, int dtype, void* out, cudaStream_t st) {
  if (dtype == EBOS_F64) return blur3_t<double>((const double*)in, 1, H, W, sigma, adjoint, (double*)out, st);
  return blur3_t<float>((const float*)in, 1, H, W, sigma, adjoint, (float*)out, st);
}

}  // namespace ebos

using namespace ebos;

extern "C" {

int ebos_flow_error(const void* flow_gt, const void* flow_pred, const uint8_t* event_mask, int mask_batched,
                    const void* time_scale, int n_points_fp32, int batch, int H, int W, int dtype, double* workspace,
                    double* errors, void* stream) {
  EBOS_REQUIRE(flow_gt && flow_pred && workspace && errors && batch > 0 && H > 0 && W > 0, "ebos_flow_error: bad argument");
  if (dtype != EBOS_F32 && dtype != EBOS_F64) { set_error("ebos_flow_error: unsupported dtype"); return EBOS_ERR_UNSUPPORTED; }
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)batch * kErrSlots * sizeof(double), st);
  if (e != cudaSuccess) return cuda_fail(e, "ebos_flow_error memset");
  const int64_t hw = (int64_t)H * W;
  const int bx = (int)std::max<int64_t>(1, std::min<int64_t>((hw + 255) / 256, (int64_t)sm_count() * 4));
  const dim3 grid(bx, batch);
  const int64_t mstride = mask_batched ? hw : 0;
  if (dtype == EBOS_F64)
    k_flow_error<double><<<grid, 256, 0, st>>>((const double*)flow_gt, (const double*)flow_pred, event_mask, mstride,
                                               (const double*)time_scale, hw, workspace);
  else
    k_flow_error<float><<<grid, 256, 0, st>>>((const float*)flow_gt, (const float*)flow_pred, event_mask, mstride,
                                              (const float*)time_scale, hw, workspace);
  k_flow_error_final<<<1, 32, 0, st>>>(workspace, batch, n_points_fp32, errors);
  EBOS_LAUNCH_CHECK("ebos_flow_error");
  return EBOS_OK;
}

size_t ebos_flow_error_workspace_doubles(int batch) { return batch > 0 ? (size_t)batch * kErrSlots : 0; }

int ebos_capture_begin(void* stream) {
  cudaError_t e = cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) return cuda_fail(e, "ebos_capture_begin");
  return EBOS_OK;
}

int ebos_capture_end_count(void* stream, int32_t* n_kernel_nodes, int32_t* n_other_nodes) {
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(as_stream(stream), &graph);
  if (e != cudaSuccess || !graph) return cuda_fail(e, "ebos_capture_end_count");
  size_t n = 0;
  int32_t kernels = 0, others = 0;
  e = cudaGraphGetNodes(graph, nullptr, &n);
  if (e == cudaSuccess && n > 0) {
    cudaGraphNode_t* nodes = new cudaGraphNode_t[n];
    e = cudaGraphGetNodes(graph, nodes, &n);
    for (size_t i = 0; e == cudaSuccess && i < n; ++i) {
      cudaGraphNodeType t;
      e = cudaGraphNodeGetType(nodes[i], &t);
      if (e == cudaSuccess) { if (t == cudaGraphNodeTypeKernel) ++kernels; else ++others; }
    }
    delete[] nodes;
  }
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return cuda_fail(e, "ebos_capture_end_count(nodes)");
  if (n_kernel_nodes) *n_kernel_nodes = kernels;
  if (n_other_nodes) *n_other_nodes = others;
  return EBOS_OK;
}

// ---- replayable launch sequences (solver loops) ------------------------------------------------------------------
// A solver captures `unroll` iterations once per window and replays them n_iter / unroll times.  Instantiating and
// destroying an executable graph per window was measured to cost more than it looks (B200, r02l): destroying an
// executable graph (and freeing the per-graph memory pool torch.cuda.CUDAGraph attaches) waits for ALL work in flight on
// the device, which serialised the "concurrent" windows of estimate_many completely.  Here the executable graph
// belongs to a SLOT that lives as long as the solver: a new window is captured into a throw-away cudaGraph_t (host
// object) and the slot's executable is UPDATED in place (cudaGraphExecUpdate: same topology, new pointers / sizes /
// grids) -- no instantiation, no destruction, no device synchronisation in steady state.  When the topology differs
// (another kernel variant, another unroll) the executable is rebuilt.
int ebos_capture_end_exec(void* stream, void** exec_inout, int32_t* updated_in_place) {
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(as_stream(stream), &graph);
  if (e != cudaSuccess || !graph) return cuda_fail(e == cudaSuccess ? cudaErrorUnknown : e, "ebos_capture_end_exec");
  if (!exec_inout) { cudaGraphDestroy(graph); set_error("ebos_capture_end_exec: exec_inout is NULL"); return EBOS_ERR_BAD_ARG; }
  cudaGraphExec_t exec = reinterpret_cast<cudaGraphExec_t>(*exec_inout);
  int32_t updated = 0;
  if (exec) {
    cudaGraphExecUpdateResultInfo info;
    if (cudaGraphExecUpdate(exec, graph, &info) == cudaSuccess) {
      updated = 1;
    } else {
      (void)cudaGetLastError();          // not an error of ours: the executable no longer matches, rebuild it
      cudaGraphExecDestroy(exec);
      exec = nullptr;
      *exec_inout = nullptr;
    }
  }
  if (!exec) {
    e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return cuda_fail(e, "ebos_capture_end_exec(instantiate)"); }
    *exec_inout = exec;
  }
  cudaGraphDestroy(graph);
  if (updated_in_place) *updated_in_place = updated;
  return EBOS_OK;
}

int ebos_exec_launch(void* exec, void* stream) {
  EBOS_REQUIRE(exec, "ebos_exec_launch: no executable (capture first)");
  cudaError_t e = cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(exec), as_stream(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ebos_exec_launch");
  return EBOS_OK;
}

int ebos_exec_destroy(void* exec) {
  if (!exec) return EBOS_OK;
  cudaError_t e = cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(exec));
  if (e != cudaSuccess) return cuda_fail(e, "ebos_exec_destroy");
  return EBOS_OK;
}

int ebos_blur3(const void* image, int batch, int H, int W, double sigma, int adjoint, int dtype, void* out, void* stream) {
  EBOS_REQUIRE(image && out && batch > 0 && H >= 2 && W >= 2 && sigma > 0.0 && image != out, "ebos_blur3: bad argument");
  if (dtype != EBOS_F32 && dtype != EBOS_F64) { set_error("ebos_blur3: unsupported dtype"); return EBOS_ERR_UNSUPPORTED; }
  if (dtype == EBOS_F64) return blur3_t<double>((const double*)image, batch, H, W, sigma, adjoint, (double*)out, as_stream(stream));
  return blur3_t<float>((const float*)image, batch, H, W, sigma, adjoint, (float*)out, as_stream(stream));
}

}  // extern "C"
