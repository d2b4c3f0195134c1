// Micro-benchmark: candidate scatter strategies for the warp+splat forward on B200.
// Standalone (nvcc -gencode arch=compute_100a,code=sm_100a -O3 splat_variants.cu -o splat_variants).
// Not part of the product; it exists to pick the splat design from measured rates.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>
#include <cmath>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %s:%d\n",cudaGetErrorString(e),__FILE__,__LINE__); exit(1);} }while(0)

constexpr int H = 720, W = 1280;
constexpr int TH = 16, TW = 64;         // tile of original pixels
constexpr int NTR = H / TH, NTC = W / TW;
constexpr int R = 4;                    // halo
constexpr int SH = TH + 2 * R + 1, SW = TW + 2 * R + 1 + 3;  // 25 x 76

struct Taps { int fr, fc; float w0, w1, w2, w3; };

__device__ __forceinline__ Taps warp_one(float x, float y, float d, const float* __restrict__ flow, int k) {
  float f0 = __ldg(flow + k), f1 = __ldg(flow + H * W + k);
  float xw = __fsub_rn(x, __fmul_rn(d, f0));
  float yw = __fsub_rn(y, __fmul_rn(d, f1));
  float flr = floorf(__fadd_rn(xw, 1e-6f)), flc = floorf(__fadd_rn(yw, 1e-6f));
  float a = __fsub_rn(xw, flr), b = __fsub_rn(yw, flc);
  float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
  Taps t; t.fr = (int)flr; t.fc = (int)flc;
  t.w0 = __fmul_rn(na, nb); t.w1 = __fmul_rn(a, nb); t.w2 = __fmul_rn(na, b); t.w3 = __fmul_rn(a, b);
  return t;
}

__device__ __forceinline__ void red1(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red2(float* p, float a, float b) { asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory"); }
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) { asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory"); }

// global flush of one cell (fr,fc) with weights (row r: wA at c, wB at c+1 ; row r+1: wC, wD)
template <int VEC>
__device__ __forceinline__ void flush_global(float* __restrict__ iwe, int fr, int fc, float w0, float w1, float w2, float w3) {
  // taps: 0:(r,c) 1:(r+1,c) 2:(r,c+1) 3:(r+1,c+1)
  bool r0 = (fr >= 0) & (fr < H), r1 = (fr + 1 >= 0) & (fr + 1 < H);
  bool c0 = (fc >= 0) & (fc < W), c1 = (fc + 1 >= 0) & (fc + 1 < W);
  if (VEC == 1) {
    if (r0 & c0) red1(iwe + fr * W + fc, w0);
    if (r1 & c0) red1(iwe + (fr + 1) * W + fc, w1);
    if (r0 & c1) red1(iwe + fr * W + fc + 1, w2);
    if (r1 & c1) red1(iwe + (fr + 1) * W + fc + 1, w3);
  } else if (VEC == 2) {
    if (c0 & c1 & ((fc & 1) == 0)) {
      if (r0) red2(iwe + fr * W + fc, w0, w2);
      if (r1) red2(iwe + (fr + 1) * W + fc, w1, w3);
    } else {
      if (r0 & c0) red1(iwe + fr * W + fc, w0);
      if (r1 & c0) red1(iwe + (fr + 1) * W + fc, w1);
      if (r0 & c1) red1(iwe + fr * W + fc + 1, w2);
      if (r1 & c1) red1(iwe + (fr + 1) * W + fc + 1, w3);
    }
  } else {
    int j = fc & 3;
    if (c0 & c1 & (j != 3)) {
      int cb = fc - j;
      float z = 0.f;
      if (r0) red4(iwe + fr * W + cb, j == 0 ? w0 : z, j == 0 ? w2 : (j == 1 ? w0 : z), j == 1 ? w2 : (j == 2 ? w0 : z), j == 2 ? w2 : z);
      if (r1) red4(iwe + (fr + 1) * W + cb, j == 0 ? w1 : z, j == 0 ? w3 : (j == 1 ? w1 : z), j == 1 ? w3 : (j == 2 ? w1 : z), j == 2 ? w3 : z);
    } else {
      if (r0 & c0) red1(iwe + fr * W + fc, w0);
      if (r1 & c0) red1(iwe + (fr + 1) * W + fc, w1);
      if (r0 & c1) red1(iwe + fr * W + fc + 1, w2);
      if (r1 & c1) red1(iwe + (fr + 1) * W + fc + 1, w3);
    }
  }
}

// ---------------- global-RED variants ----------------
// EPT events per thread, consecutive; AGG: combine runs of equal cell in registers.
template <int VEC, int EPT, bool AGG>
__global__ void __launch_bounds__(256) k_global(const float* __restrict__ ex, const float* __restrict__ ey, const float* __restrict__ ed,
                                                const float* __restrict__ flow, float* __restrict__ iwe, int n) {
  int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * EPT;
  if (base >= n) return;
  float x[EPT], y[EPT], d[EPT];
  if (EPT % 4 == 0 && base + EPT <= n) {
#pragma unroll
    for (int j = 0; j < EPT; j += 4) {
      float4 vx = *reinterpret_cast<const float4*>(ex + base + j);
      float4 vy = *reinterpret_cast<const float4*>(ey + base + j);
      float4 vd = *reinterpret_cast<const float4*>(ed + base + j);
      x[j] = vx.x; x[j + 1] = vx.y; x[j + 2] = vx.z; x[j + 3] = vx.w;
      y[j] = vy.x; y[j + 1] = vy.y; y[j + 2] = vy.z; y[j + 3] = vy.w;
      d[j] = vd.x; d[j + 1] = vd.y; d[j + 2] = vd.z; d[j + 3] = vd.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      bool ok = base + j < n;
      x[j] = ok ? ex[base + j] : -100.f; y[j] = ok ? ey[base + j] : -100.f; d[j] = ok ? ed[base + j] : 0.f;
    }
  }
  int cfr = INT_MIN, cfc = INT_MIN; float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    if (x[j] < 0.f) continue;
    int k = (int)x[j] * W + (int)y[j];
    Taps t = warp_one(x[j], y[j], d[j], flow, k);
    if (AGG) {
      if (t.fr != cfr || t.fc != cfc) {
        if (cfr != INT_MIN) flush_global<VEC>(iwe, cfr, cfc, a0, a1, a2, a3);
        cfr = t.fr; cfc = t.fc; a0 = t.w0; a1 = t.w1; a2 = t.w2; a3 = t.w3;
      } else { a0 += t.w0; a1 += t.w1; a2 += t.w2; a3 += t.w3; }
    } else {
      flush_global<VEC>(iwe, t.fr, t.fc, t.w0, t.w1, t.w2, t.w3);
    }
  }
  if (AGG && cfr != INT_MIN) flush_global<VEC>(iwe, cfr, cfc, a0, a1, a2, a3);
}

// ---------------- smem tile variants ----------------
// one CTA per tile; events of the tile are [toff[t], toff[t+1]).
// SMODE 0: f32 atomicAdd (CAS loop)   1: s32 fixed point ATOMS.ADD   2: racy plain RMW (rate reference only)
constexpr float FIX = 262144.f;  // 2^18
template <int SMODE, int EPT, bool AGG>
__global__ void __launch_bounds__(512) k_tile(const float* __restrict__ ex, const float* __restrict__ ey, const float* __restrict__ ed,
                                              const int* __restrict__ toff, const float* __restrict__ flow, float* __restrict__ iwe) {
  __shared__ float sf[SH * SW];
  int* si = reinterpret_cast<int*>(sf);
  for (int i = threadIdx.x; i < SH * SW; i += blockDim.x) sf[i] = 0.f;  // 0 bits for both
  __syncthreads();
  int t = blockIdx.x, tr = t / NTC, tc = t % NTC;
  int r_org = tr * TH - R, c_org = tc * TW - R;
  int s = toff[t], e = toff[t + 1];
  auto flush = [&](int fr, int fc, float w0, float w1, float w2, float w3) {
    int lr = fr - r_org, lc = fc - c_org;
    if (lr >= 0 && lr + 1 < SH && lc >= 0 && lc + 1 < SW - 3) {
      // bounds against image: smem tile may extend outside the image; the flush masks that.
      int o = lr * SW + lc;
      if (SMODE == 0) { atomicAdd(sf + o, w0); atomicAdd(sf + o + SW, w1); atomicAdd(sf + o + 1, w2); atomicAdd(sf + o + SW + 1, w3); }
      else if (SMODE == 1) {
        atomicAdd(si + o, __float2int_rn(w0 * FIX)); atomicAdd(si + o + SW, __float2int_rn(w1 * FIX));
        atomicAdd(si + o + 1, __float2int_rn(w2 * FIX)); atomicAdd(si + o + SW + 1, __float2int_rn(w3 * FIX));
      } else { sf[o] += w0; sf[o + SW] += w1; sf[o + 1] += w2; sf[o + SW + 1] += w3; }
    } else {
      flush_global<1>(iwe, fr, fc, w0, w1, w2, w3);
    }
  };
  for (int base = s + threadIdx.x * EPT; base < e; base += blockDim.x * EPT) {
    int cfr = INT_MIN, cfc = INT_MIN; float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      int i = base + j;
      if (i >= e) break;
      float x = ex[i], y = ey[i], d = ed[i];
      int k = (int)x * W + (int)y;
      Taps tp = warp_one(x, y, d, flow, k);
      if (AGG) {
        if (tp.fr != cfr || tp.fc != cfc) {
          if (cfr != INT_MIN) flush(cfr, cfc, a0, a1, a2, a3);
          cfr = tp.fr; cfc = tp.fc; a0 = tp.w0; a1 = tp.w1; a2 = tp.w2; a3 = tp.w3;
        } else { a0 += tp.w0; a1 += tp.w1; a2 += tp.w2; a3 += tp.w3; }
      } else flush(tp.fr, tp.fc, tp.w0, tp.w1, tp.w2, tp.w3);
    }
    if (AGG && cfr != INT_MIN) flush(cfr, cfc, a0, a1, a2, a3);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SH * SW; i += blockDim.x) {
    int lr = i / SW, lc = i % SW;
    int r = r_org + lr, c = c_org + lc;
    float v = (SMODE == 1) ? (float)si[i] * (1.f / FIX) : sf[i];
    if (lc < SW - 3 && r >= 0 && r < H && c >= 0 && c < W && v != 0.f) red1(iwe + r * W + c, v);
  }
}

// gather-rate reference for the backward: 4 LDG taps from a plane + segmented-free sum
template <int EPT>
__global__ void __launch_bounds__(256) k_gather(const float* __restrict__ ex, const float* __restrict__ ey, const float* __restrict__ ed,
                                                const float* __restrict__ flow, const float* __restrict__ g, float* __restrict__ dflow, int n) {
  int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * EPT;
  int ck = -1; float s0 = 0, s1 = 0;
#pragma unroll
  for (int j = 0; j < EPT; ++j) {
    int64_t i = base + j; if (i >= n) break;
    float x = ex[i], y = ey[i], d = ed[i];
    int k = (int)x * W + (int)y;
    Taps t = warp_one(x, y, d, flow, k);
    auto G = [&](int r, int c) { return (r >= 0 && r < H && c >= 0 && c < W) ? __ldg(g + r * W + c) : 0.f; };
    float g00 = G(t.fr, t.fc), g10 = G(t.fr + 1, t.fc), g01 = G(t.fr, t.fc + 1), g11 = G(t.fr + 1, t.fc + 1);
    float a = t.w1 + t.w3, b = t.w2 + t.w3;  // ~frac parts
    float dx = (1 - b) * (g10 - g00) + b * (g11 - g01), dy = (1 - a) * (g01 - g00) + a * (g11 - g10);
    if (k != ck) { if (ck >= 0) { red1(dflow + ck, s0); red1(dflow + H * W + ck, s1); } ck = k; s0 = 0; s1 = 0; }
    s0 -= d * dx; s1 -= d * dy;
  }
  if (ck >= 0) { red1(dflow + ck, s0); red1(dflow + H * W + ck, s1); }
}

template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms); }
  return best * 1000.f;
}

int main(int argc, char** argv) {
  int logn = argc > 1 ? atoi(argv[1]) : 24;
  float fmax = argc > 2 ? atof(argv[2]) : 3.f;
  int n = 1 << logn;
  printf("# N=%d (2^%d) flow U(-%g,%g) H=%d W=%d\n", n, logn, fmax, fmax, H, W);
  std::mt19937_64 rng(0);
  std::vector<float> x(n), y(n), t(n), flow(2 * H * W);
  for (int i = 0; i < n; ++i) { x[i] = (float)(rng() % H); y[i] = (float)(rng() % W); t[i] = (rng() >> 11) * (1.0 / 9007199254740992.0); }
  std::sort(t.begin(), t.end());
  for (auto& f : flow) f = ((rng() >> 11) * (1.0 / 9007199254740992.0) * 2 - 1) * fmax;
  // tile-major pixel key
  std::vector<int> key(n), perm(n);
  for (int i = 0; i < n; ++i) { int r = (int)x[i], c = (int)y[i]; int tile = (r / TH) * NTC + c / TW; key[i] = tile * (TH * TW) + (r % TH) * TW + (c % TW); }
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
  std::vector<float> sx(n), sy(n), sd(n);
  std::vector<int> toff(NTR * NTC + 1, 0);
  for (int i = 0; i < n; ++i) { sx[i] = x[perm[i]]; sy[i] = y[perm[i]]; sd[i] = t[perm[i]]; toff[key[perm[i]] / (TH * TW) + 1]++; }
  for (int i = 0; i < NTR * NTC; ++i) toff[i + 1] += toff[i];
  // CPU reference (double)
  std::vector<double> ref(H * W, 0.0);
  for (int i = 0; i < n; ++i) {
    int k = (int)x[i] * W + (int)y[i];
    float xw = x[i] - t[i] * flow[k], yw = y[i] - t[i] * flow[H * W + k];
    float fr = floorf(xw + 1e-6f), fc = floorf(yw + 1e-6f); float a = xw - fr, b = yw - fc; int r = (int)fr, c = (int)fc;
    auto add = [&](int rr, int cc, float w) { if (rr >= 0 && rr < H && cc >= 0 && cc < W) ref[rr * W + cc] += w; };
    add(r, c, (1 - a) * (1 - b)); add(r + 1, c, a * (1 - b)); add(r, c + 1, (1 - a) * b); add(r + 1, c + 1, a * b);
  }
  double refmax = *std::max_element(ref.begin(), ref.end());
  float *dx, *dy, *dd, *dsx, *dsy, *dsd, *dflow, *diwe, *dg; int* dtoff;
  CK(cudaMalloc(&dx, n * 4)); CK(cudaMalloc(&dy, n * 4)); CK(cudaMalloc(&dd, n * 4));
  CK(cudaMalloc(&dsx, n * 4)); CK(cudaMalloc(&dsy, n * 4)); CK(cudaMalloc(&dsd, n * 4));
  CK(cudaMalloc(&dflow, 2 * H * W * 4)); CK(cudaMalloc(&diwe, H * W * 4)); CK(cudaMalloc(&dg, 2 * H * W * 4)); CK(cudaMalloc(&dtoff, toff.size() * 4));
  CK(cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dy, y.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dd, t.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dsx, sx.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dsy, sy.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dsd, sd.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dflow, flow.data(), 2 * H * W * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dtoff, toff.data(), toff.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> out(H * W);
  auto check = [&](const char* name, float us) {
    CK(cudaMemcpy(out.data(), diwe, H * W * 4, cudaMemcpyDeviceToHost));
    double md = 0, sum = 0; for (int i = 0; i < H * W; ++i) { md = std::max(md, std::fabs(out[i] - ref[i])); sum += out[i]; }
    printf("%-34s %9.1f us  %7.2f Gev/s   maxabs/max=%.2e sum=%.6e\n", name, us, n / us * 1e-3, md / refmax, sum);
    fflush(stdout);
  };
#define RUN_G(name, VEC, EPT, AGG, X, Y, D) { auto f = [&]() { cudaMemsetAsync(diwe, 0, H * W * 4); k_global<VEC, EPT, AGG><<<(n / EPT + 255) / 256, 256>>>(X, Y, D, dflow, diwe, n); }; float us = timeit(f); f(); check(name, us); }
  { auto f = [&]() { cudaMemsetAsync(diwe, 0, H * W * 4); }; printf("%-34s %9.1f us\n", "memset plane", timeit(f)); }
  RUN_G("unsorted red1 ept1", 1, 1, false, dx, dy, dd);
  RUN_G("unsorted red4 ept1", 4, 1, false, dx, dy, dd);
  RUN_G("unsorted red1 ept4", 1, 4, false, dx, dy, dd);
  RUN_G("sorted red1 ept1", 1, 1, false, dsx, dsy, dsd);
  RUN_G("sorted red2 ept1", 2, 1, false, dsx, dsy, dsd);
  RUN_G("sorted red4 ept1", 4, 1, false, dsx, dsy, dsd);
  RUN_G("sorted red1 ept4 agg", 1, 4, true, dsx, dsy, dsd);
  RUN_G("sorted red1 ept8 agg", 1, 8, true, dsx, dsy, dsd);
  RUN_G("sorted red4 ept4 agg", 4, 4, true, dsx, dsy, dsd);
  RUN_G("sorted red4 ept8 agg", 4, 8, true, dsx, dsy, dsd);
  RUN_G("sorted red4 ept16 agg", 4, 16, true, dsx, dsy, dsd);
  RUN_G("sorted red2 ept8 agg", 2, 8, true, dsx, dsy, dsd);
  RUN_G("sorted red4 ept8 noagg", 4, 8, false, dsx, dsy, dsd);
#define RUN_T(name, SMODE, EPT, AGG, TPB) { auto f = [&]() { cudaMemsetAsync(diwe, 0, H * W * 4); k_tile<SMODE, EPT, AGG><<<NTR * NTC, TPB>>>(dsx, dsy, dsd, dtoff, dflow, diwe); }; float us = timeit(f); f(); check(name, us); }
  RUN_T("tile smem f32cas ept1 512", 0, 1, false, 512);
  RUN_T("tile smem s32fix ept1 512", 1, 1, false, 512);
  RUN_T("tile smem racy   ept1 512", 2, 1, false, 512);
  RUN_T("tile smem f32cas ept8 agg 512", 0, 8, true, 512);
  RUN_T("tile smem s32fix ept8 agg 512", 1, 8, true, 512);
  RUN_T("tile smem s32fix ept8 agg 256", 1, 8, true, 256);
  RUN_T("tile smem s32fix ept4 agg 512", 1, 4, true, 512);
  RUN_T("tile smem racy   ept8 agg 512", 2, 8, true, 512);
  // backward-like gather
  CK(cudaMemcpy(dg, dflow, H * W * 4, cudaMemcpyDeviceToDevice));
  float* ddf; CK(cudaMalloc(&ddf, 2 * H * W * 4));
#define RUN_B(name, EPT, X, Y, D) { auto f = [&]() { cudaMemsetAsync(ddf, 0, 2 * H * W * 4); k_gather<EPT><<<(n / EPT + 255) / 256, 256>>>(X, Y, D, dflow, dg, ddf, n); }; float us = timeit(f); printf("%-34s %9.1f us  %7.2f Gev/s\n", name, us, n / us * 1e-3); }
  RUN_B("bwd gather unsorted ept1", 1, dx, dy, dd);
  RUN_B("bwd gather sorted ept1", 1, dsx, dsy, dsd);
  RUN_B("bwd gather sorted ept4", 4, dsx, dsy, dsd);
  RUN_B("bwd gather sorted ept8", 8, dsx, dsy, dsd);
  RUN_B("bwd gather sorted ept16", 16, dsx, dsy, dsd);
  return 0;
}
