// Event ingestion on the device: raw sensor arrays -> the reference's event rows.
//
//   raw stream   x:int16 (sensor column), y:int16 (sensor row), t:int32 [us], p:bool      9 B/event, time ordered
//                (the datasets of src/data_loader/ccs.py:50-69, `raw_events/{x,y,t,p}`)
//   event rows   [n,4] = (row = y, col = x, t = t_us / 1e6 [s], p)                        src/data_loader/ccs.py:288-296
//   CROP filter  keep row0 <= row < row1 and col0 <= col < col1, order preserved, coordinates
//                NOT shifted                                                              src/utils/event_utils.py:109-129
//   window       time_to_index(time) = searchsorted(t_us / 1e6, time) - 1                 src/data_loader/ccs.py:345-357
//
// The raw stream stays resident in HBM in its compact form (9 B/event: 180 GB hold 2e10 events); a window is cut
// out, filtered and widened to fp32/fp64 rows by one pass here and then handed to ebos_window_prepare.  In float64
// with rebase == 0 the rows are bit-identical to the loader's (IEEE double division); `rebase` subtracts the window
// origin in INTEGER microseconds first, which is what makes fp32 rows usable (the fp32 ulp at t = 10 s is 1 us).
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "ebos_common.cuh"

namespace ebos {

struct CropPred {
  const int16_t* x;   // sensor column -> event col
  const int16_t* y;   // sensor row    -> event row
  int row0, row1, col0, col1;
  __device__ __forceinline__ bool operator()(int i) const {
    const int r = y[i], c = x[i];
    return row0 <= r && r < row1 && col0 <= c && c < col1;
  }
};

template <typename T>
__global__ void __launch_bounds__(256) k_ingest_rows(const int16_t* __restrict__ x, const int16_t* __restrict__ y,
                                                     const int32_t* __restrict__ t_us, const uint8_t* __restrict__ p,
                                                     const int* __restrict__ sel, const int64_t* __restrict__ n_sel,
                                                     int64_t n, long long t_origin_us, int rebase, T* __restrict__ out) {
  const int64_t n_out = sel ? min(*n_sel, n) : n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = sel ? sel[i] : i;
    const long long tu = (long long)t_us[j] - (rebase ? t_origin_us : 0ll);
    const double ts = (double)tu / 1e6;   // "/ 1e6  # from micro sec to sec", in float64 like the loader
    // one 16-byte (fp32) or two 16-byte (fp64) stores per row
    if constexpr (sizeof(T) == 4) {
      reinterpret_cast<float4*>(out)[i] = make_float4((float)y[j], (float)x[j], (float)ts, p[j] ? 1.f : 0.f);
    } else {
      reinterpret_cast<double2*>(out)[2 * i] = make_double2((double)y[j], (double)x[j]);
      reinterpret_cast<double2*>(out)[2 * i + 1] = make_double2(ts, p[j] ? 1.0 : 0.0);
    }
  }
}

__global__ void k_copy_count(const int64_t* __restrict__ src, int64_t n, int has_sel, int64_t* __restrict__ dst) {
  *dst = has_sel ? *src : n;
}

// lower_bound over the (implicit) float64 time axis t_us / 1e6, then - 1: one thread, log2(n) dependent loads
__global__ void k_time_to_index(const int32_t* __restrict__ t_us, int64_t n, double time, int64_t* __restrict__ out) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((double)t_us[mid] / 1e6 < time) lo = mid + 1; else hi = mid;
  }
  *out = lo - 1;
}

static size_t select_tmp_bytes(int64_t n) {
  size_t b = 0;
  CropPred pred{nullptr, nullptr, 0, 0, 0, 0};
  cub::DeviceSelect::If(nullptr, b, thrust::counting_iterator<int>(0), (int*)nullptr, (int64_t*)nullptr,
                        (int)std::max<int64_t>(n, 1), pred);
  return b;
}

}  // namespace ebos

using namespace ebos;

extern "C" {

size_t ebos_ingest_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return align256((size_t)std::max<int64_t>(n, 1) * 4) + align256(select_tmp_bytes(n)) + 512;
}

int ebos_ingest_raw(const int16_t* x, const int16_t* y, const int32_t* t_us, const uint8_t* p, int64_t n, int crop,
                    int row0, int row1, int col0, int col1, int64_t t_origin_us, int rebase, int dtype, void* events_out,
                    int64_t* n_kept, void* workspace, size_t workspace_bytes, void* stream) {
  EBOS_REQUIRE(n >= 0 && n < (int64_t)INT_MAX && n_kept && (n == 0 || (x && y && t_us && p && events_out)),
               "ebos_ingest_raw: bad argument");
  if (dtype != EBOS_F32 && dtype != EBOS_F64) { set_error("ebos_ingest_raw: unsupported dtype"); return EBOS_ERR_UNSUPPORTED; }
  EBOS_REQUIRE((reinterpret_cast<size_t>(events_out) & 15) == 0, "ebos_ingest_raw: events_out must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  int* sel = nullptr;
  int64_t* n_sel = nullptr;
  if (crop && n > 0) {
    if (!workspace || workspace_bytes < ebos_ingest_workspace_bytes(n)) {
      set_error("ebos_ingest_raw: workspace too small");
      return EBOS_ERR_WORKSPACE;
    }
    char* wp = reinterpret_cast<char*>(align256(reinterpret_cast<size_t>(workspace)));
    sel = reinterpret_cast<int*>(wp);
    n_sel = reinterpret_cast<int64_t*>(wp + align256((size_t)n * 4));
    void* tmp = wp + align256((size_t)n * 4) + 256;
    size_t tmp_bytes = select_tmp_bytes(n);
    CropPred pred{x, y, row0, row1, col0, col1};
    cudaError_t ce = cub::DeviceSelect::If(tmp, tmp_bytes, thrust::counting_iterator<int>(0), sel, n_sel, (int)n, pred, st);
    if (ce != cudaSuccess) return cuda_fail(ce, "ebos_ingest_raw(select)");
  }
  k_copy_count<<<1, 1, 0, st>>>(n_sel, n, sel != nullptr, n_kept);
  if (n > 0) {
    const int bx = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8));
    if (dtype == EBOS_F64)
      k_ingest_rows<double><<<bx, 256, 0, st>>>(x, y, t_us, p, sel, n_sel, n, (long long)t_origin_us, rebase, (double*)events_out);
    else
      k_ingest_rows<float><<<bx, 256, 0, st>>>(x, y, t_us, p, sel, n_sel, n, (long long)t_origin_us, rebase, (float*)events_out);
  }
  EBOS_LAUNCH_CHECK("ebos_ingest_raw");
  return EBOS_OK;
}

int ebos_time_to_index(const int32_t* t_us, int64_t n, double time, int64_t* index_out, void* stream) {
  EBOS_REQUIRE(n >= 0 && index_out && (n == 0 || t_us), "ebos_time_to_index: bad argument");
  k_time_to_index<<<1, 1, 0, as_stream(stream)>>>(t_us, n, time, index_out);
  EBOS_LAUNCH_CHECK("ebos_time_to_index");
  return EBOS_OK;
}

}  // extern "C"
