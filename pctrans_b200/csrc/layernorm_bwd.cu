// layernorm_bwd.cu -- backward of the LayerNorm that the tensor-core Linear fuses into its epilogue
// (include/pct_linear.h: pct_linear_layernorm_backward).  Replaces, for the encoder layer's two
// `norm(residual + linear(x))` blocks (PD/msdeformattn.py:107-126), aten's two-kernel LayerNorm
// backward with ONE pass over the rows: a warp owns a row (lane l holds columns 4l..4l+3 of every
// 128-column group, so each load instruction of the warp covers 512 contiguous bytes), two warp
// reductions give the row's means, grad_pre goes out, and the warp keeps running column sums of
// grad_gamma / grad_beta in registers that are combined per CTA in shared memory and per grid with
// one fp32 reduction per column.  HBM traffic: read grad_y and pre_norm, write grad_pre.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/pct_linear.h"
#include "../../include/pct_msda.h"

namespace pct {
void count_launch();
}

namespace {

constexpr int kLnThreads = 256;        // 8 warps per CTA
constexpr int ln_ctas_per_sm(int nv) { return nv <= 2 ? 4 : 2; }   // 64 / 128 registers per thread

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int w = 16; w > 0; w >>= 1) v += __shfl_xor_sync(0xffffffffu, v, w);
  return v;
}

// NV = features / 128: float4 pieces per lane and row
template <int NV>
__global__ void __launch_bounds__(kLnThreads, ln_ctas_per_sm(NV))
layernorm_bwd_kernel(const float* __restrict__ grad_y, const float* __restrict__ pre,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, float* __restrict__ grad_pre,
                     float* __restrict__ grad_gamma, float* __restrict__ grad_beta, int rows) {
  constexpr int N = NV * 128;
  static_assert(N <= 512, "layernorm_bwd_kernel: up to 512 features (32 KB of column sums per CTA)");
  __shared__ float s_acc[2][kLnThreads / 32][N];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (gridDim.x * kLnThreads) >> 5;
  const float inv_n = 1.f / (float)N;

  float4 gam[NV], dg[NV], db[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    gam[j] = __ldg(reinterpret_cast<const float4*>(gamma) + j * 32 + lane);
    dg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r = blockIdx.x * (kLnThreads / 32) + warp; r < rows; r += nwarps) {
    const float4* gy4 = reinterpret_cast<const float4*>(grad_y + (size_t)r * N);
    const float4* x4 = reinterpret_cast<const float4*>(pre + (size_t)r * N);
    float4 gy[NV], xh[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      gy[j] = __ldcs(gy4 + j * 32 + lane);
      xh[j] = __ldcs(x4 + j * 32 + lane);
    }
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      xh[j].x = (xh[j].x - mu) * rs; xh[j].y = (xh[j].y - mu) * rs;
      xh[j].z = (xh[j].z - mu) * rs; xh[j].w = (xh[j].w - mu) * rs;
      dg[j].x = fmaf(gy[j].x, xh[j].x, dg[j].x); dg[j].y = fmaf(gy[j].y, xh[j].y, dg[j].y);
      dg[j].z = fmaf(gy[j].z, xh[j].z, dg[j].z); dg[j].w = fmaf(gy[j].w, xh[j].w, dg[j].w);
      db[j].x += gy[j].x; db[j].y += gy[j].y; db[j].z += gy[j].z; db[j].w += gy[j].w;
      // g = grad_y * gamma from here on
      gy[j].x *= gam[j].x; gy[j].y *= gam[j].y; gy[j].z *= gam[j].z; gy[j].w *= gam[j].w;
      s1 += (gy[j].x + gy[j].y) + (gy[j].z + gy[j].w);
      s2 += fmaf(gy[j].x, xh[j].x, gy[j].y * xh[j].y) + fmaf(gy[j].z, xh[j].z, gy[j].w * xh[j].w);
    }
    s1 = warp_sum(s1) * inv_n;
    s2 = warp_sum(s2) * inv_n;
    float4* out4 = reinterpret_cast<float4*>(grad_pre + (size_t)r * N);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 o;
      o.x = rs * (gy[j].x - s1 - xh[j].x * s2);
      o.y = rs * (gy[j].y - s1 - xh[j].y * s2);
      o.z = rs * (gy[j].z - s1 - xh[j].z * s2);
      o.w = rs * (gy[j].w - s1 - xh[j].w * s2);
      out4[j * 32 + lane] = o;
    }
  }
  // column sums: warps of the CTA through shared memory, then one reduction per column and CTA
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    *reinterpret_cast<float4*>(&s_acc[0][warp][j * 128 + lane * 4]) = dg[j];
    *reinterpret_cast<float4*>(&s_acc[1][warp][j * 128 + lane * 4]) = db[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * N; c += kLnThreads) {
    const int which = c >= N, col = which ? c - N : c;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kLnThreads / 32; ++w) s += s_acc[which][w][col];
    atomicAdd((which ? grad_beta : grad_gamma) + col, s);
  }
}

template <int NV>
int launch(const float* gy, const float* pre, const float* mean, const float* rstd, const float* gamma,
           float* gpre, float* ggam, float* gbeta, int rows, cudaStream_t st) {
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  const long long want = ((long long)rows + kLnThreads / 32 - 1) / (kLnThreads / 32);
  const long long cap = (long long)sms * ln_ctas_per_sm(NV);
  const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
  layernorm_bwd_kernel<NV><<<grid, kLnThreads, 0, st>>>(gy, pre, mean, rstd, gamma, gpre, ggam, gbeta, rows);
  pct::count_launch();
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int pct_linear_layernorm_backward(const float* grad_y, const float* pre_norm, const float* mean,
                                      const float* rstd, const float* gamma, float* grad_pre,
                                      float* grad_gamma, float* grad_beta, int rows, int features,
                                      void* cuda_stream) {
  if (!grad_y || !pre_norm || !mean || !rstd || !gamma || !grad_pre || !grad_gamma || !grad_beta)
    return PCT_MSDA_E_NULL_POINTER;
  if (rows < 1 || features < 128 || features > 512 || features % 128 != 0) return PCT_MSDA_E_UNSUPPORTED;
  if ((((uintptr_t)grad_y | (uintptr_t)pre_norm | (uintptr_t)grad_pre | (uintptr_t)gamma) & 15u) ||
      (((uintptr_t)mean | (uintptr_t)rstd | (uintptr_t)grad_gamma | (uintptr_t)grad_beta) & 3u))
    return PCT_MSDA_E_MISALIGNED;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cudaError_t e = cudaMemsetAsync(grad_gamma, 0, sizeof(float) * features, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(grad_beta, 0, sizeof(float) * features, st);
  if (e != cudaSuccess) return (int)e;
  switch (features / 128) {
    case 1: return launch<1>(grad_y, pre_norm, mean, rstd, gamma, grad_pre, grad_gamma, grad_beta, rows, st);
    case 2: return launch<2>(grad_y, pre_norm, mean, rstd, gamma, grad_pre, grad_gamma, grad_beta, rows, st);
    case 3: return launch<3>(grad_y, pre_norm, mean, rstd, gamma, grad_pre, grad_gamma, grad_beta, rows, st);
    case 4: return launch<4>(grad_y, pre_norm, mean, rstd, gamma, grad_pre, grad_gamma, grad_beta, rows, st);
  }
  return PCT_MSDA_E_UNSUPPORTED;
}
