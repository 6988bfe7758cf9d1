// msda_kernels.cuh -- sm_100a kernels for multi-scale deformable attention (MSDeformAttn).
//
// Written from scratch for B200; the arithmetic follows the reference's definition of the op
// (OPS/src/cuda/ms_deform_im2col_cuda.cuh:38-89 forward taps, :93-164 gradients, :290-293 pixel
// coordinates and in-range predicate) but none of its structure:
//
//  * Work is laid out as per-(image, head) STREAMS of queries.  A persistent CTA owns a
//    contiguous slice of the stream-ordered task list, so all of its warps walk neighbouring
//    queries of ONE head of ONE image: the bilinear footprints of neighbouring queries overlap,
//    and the SM's L1 acts as a sliding window over each level's value map (the gather moves 15x
//    the HBM bytes through L1, so L1 locality -- not HBM -- decides the speed of this op).
//  * A (query, head) pair is handled by LPP = D*sizeof(T)/16 lanes, each owning one 128-bit
//    channel vector, so every tap is one LDG.128 per lane and a warp covers 32/LPP pairs.
//  * grad_value leaves the SM as 128-bit vector reductions (red.global.add.v4.f32, one per tap per
//    lane) instead of per-channel scalar atomics; grad_sampling_loc / grad_attn_weight are reduced
//    over the LPP channel lanes with xor-shuffles and written once (no shared memory, no
//    __syncthreads in either kernel's main loop).
//  * The generic kernels (any D, fp64, L > kMaxFastLevels) keep the same stream ordering with one
//    warp per pair.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pct {

constexpr int kMaxFastLevels = 16;
constexpr int kVecThreads = 1024;  // forward: one persistent 32-warp CTA per SM
constexpr int kBwdVecThreads = 768; // backward needs ~80 registers per thread
constexpr int kGenericThreads = 256;

// ---------------------------------------------------------------------------------------------
// 128-bit channel vectors
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Vec16;

template <>
struct Vec16<float> {
  static constexpr int kElems = 4;
  static __device__ __forceinline__ void load(const float* p, bool pred, float (&v)[4]) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pred) r = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void load_stream(const float* p, float (&v)[4]) {
    float4 r = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};

template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int kElems = 8;
  static __device__ __forceinline__ void unpack(const uint4& r, float (&v)[8]) {
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, bool pred, float (&v)[8]) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (pred) r = __ldg(reinterpret_cast<const uint4*>(p));
    unpack(r, v);
  }
  static __device__ __forceinline__ void load_stream(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = __ldcs(reinterpret_cast<const uint4*>(p));
    unpack(r, v);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    unsigned* w = reinterpret_cast<unsigned*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<unsigned*>(&h);
    }
    __stcs(reinterpret_cast<uint4*>(p), r);
  }
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// Bilinear taps (cuh:45-58 geometry, cuh:61-83 per-corner bounds, cuh:293 in-range predicate)
// ---------------------------------------------------------------------------------------------
struct Taps {
  int off00;           // element offset of pixel (y0, x0) inside the level map, head 0 channel 0
  bool m00, m01, m10, m11;  // corner is inside the map (and the point passed the predicate)
  float ly, lx, hy, hx;
};

__device__ __forceinline__ Taps make_taps(float loc_x, float loc_y, int H, int W, int MD) {
  Taps t;
  // The reference evaluates loc*size - 0.5 in double and rounds once (the literal 0.5 promotes,
  // cuh:290-291): that is exactly one fused multiply-add in fp32.
  const float y = fmaf(loc_y, (float)H, -0.5f);
  const float x = fmaf(loc_x, (float)W, -0.5f);
  const bool inside = (y > -1.f) && (x > -1.f) && (y < (float)H) && (x < (float)W);
  const float fy = floorf(y), fx = floorf(x);
  const int y0 = (int)fy, x0 = (int)fx;
  t.ly = y - fy;
  t.lx = x - fx;
  t.hy = 1.f - t.ly;
  t.hx = 1.f - t.lx;
  const bool top = inside && (y0 >= 0), bot = inside && (y0 + 1 <= H - 1);
  const bool lef = (x0 >= 0), rig = (x0 + 1 <= W - 1);
  t.m00 = top && lef;
  t.m01 = top && rig;
  t.m10 = bot && lef;
  t.m11 = bot && rig;
  t.off00 = (y0 * W + x0) * MD;
  return t;
}

struct StreamTask {
  int b, m, q;      // q = first query of the warp's group
  long long pair;   // ((b*Lq + q) * M + m) for the lane's own pair (clamped to a valid query)
  bool valid;
};

// ---------------------------------------------------------------------------------------------
// Forward, vector path.  T = float | bf16; LPP lanes per (query, head) pair; D = LPP * 16/sizeof(T)
// ---------------------------------------------------------------------------------------------
template <typename T, int LPP>
__global__ void __launch_bounds__(kVecThreads, 1)
msda_fwd_vec_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                    const float* __restrict__ attn, T* __restrict__ out, int N, int S, int M,
                    int L, int Lq, int P) {
  constexpr int VE = Vec16<T>::kElems;
  constexpr int D = LPP * VE;
  constexpr int PPW = 32 / LPP;  // pairs per warp
  __shared__ int sH[kMaxFastLevels], sW[kMaxFastLevels], sStart[kMaxFastLevels];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = (int)shapes[2 * threadIdx.x];
    sW[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
    sStart[threadIdx.x] = (int)lsi[threadIdx.x];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int grp = lane / LPP, sub = lane % LPP;
  const int MD = M * D;
  const int LP = L * P;
  const int tasks_per_stream = (Lq + PPW - 1) / PPW;
  const long long n_tasks = (long long)N * M * tasks_per_stream;
  const long long t_begin = n_tasks * blockIdx.x / gridDim.x;
  const long long t_end = n_tasks * (blockIdx.x + 1) / gridDim.x;

  for (long long t = t_begin + warp; t < t_end; t += nwarps) {
    const int stream = (int)(t / tasks_per_stream);
    const int tq = (int)(t - (long long)stream * tasks_per_stream);
    const int b = stream / M, m = stream - b * M;
    const int q = tq * PPW + grp;
    const bool qvalid = q < Lq;
    const long long pair = ((long long)b * Lq + (qvalid ? q : Lq - 1)) * M + m;
    const float2* lp = reinterpret_cast<const float2*>(loc) + pair * LP;
    const float* ap = attn + pair * LP;
    const T* vimg = value + ((long long)b * S * M + m) * D + sub * VE;

    float acc[VE];
#pragma unroll
    for (int i = 0; i < VE; ++i) acc[i] = 0.f;

    for (int l = 0; l < L; ++l) {
      const int H = sH[l], W = sW[l];
      const T* vl = vimg + (long long)sStart[l] * MD;
      const int rowstep = W * MD;
#pragma unroll 2
      for (int p = 0; p < P; ++p) {
        const float2 xy = __ldg(lp + l * P + p);
        const float a = __ldg(ap + l * P + p);
        const Taps tp = make_taps(xy.x, xy.y, H, W, MD);
        const T* p00 = vl + tp.off00;
        float v00[VE], v01[VE], v10[VE], v11[VE];
        Vec16<T>::load(p00, tp.m00, v00);
        Vec16<T>::load(p00 + MD, tp.m01, v01);
        Vec16<T>::load(p00 + rowstep, tp.m10, v10);
        Vec16<T>::load(p00 + rowstep + MD, tp.m11, v11);
        const float w00 = tp.hy * tp.hx * a, w01 = tp.hy * tp.lx * a;
        const float w10 = tp.ly * tp.hx * a, w11 = tp.ly * tp.lx * a;
#pragma unroll
        for (int i = 0; i < VE; ++i)
          acc[i] = fmaf(w00, v00[i], fmaf(w01, v01[i], fmaf(w10, v10[i], fmaf(w11, v11[i], acc[i]))));
      }
    }
    if (qvalid) Vec16<T>::store(out + pair * D + sub * VE, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// Backward, vector path.  grad_value is accumulated in fp32 (`gacc`: grad_value itself for fp32
// values, the caller's workspace for bf16 values) with 128-bit vector reductions.
// ---------------------------------------------------------------------------------------------
template <typename T, int LPP>
__global__ void __launch_bounds__(kBwdVecThreads, 1)
msda_bwd_vec_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                    const float* __restrict__ attn, const T* __restrict__ grad_out,
                    float* __restrict__ gacc, float* __restrict__ grad_loc,
                    float* __restrict__ grad_attn, int N, int S, int M, int L, int Lq, int P) {
  constexpr int VE = Vec16<T>::kElems;
  constexpr int D = LPP * VE;
  constexpr int PPW = 32 / LPP;
  __shared__ int sH[kMaxFastLevels], sW[kMaxFastLevels], sStart[kMaxFastLevels];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = (int)shapes[2 * threadIdx.x];
    sW[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
    sStart[threadIdx.x] = (int)lsi[threadIdx.x];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int grp = lane / LPP, sub = lane % LPP;
  const int MD = M * D;
  const int LP = L * P;
  const int tasks_per_stream = (Lq + PPW - 1) / PPW;
  const long long n_tasks = (long long)N * M * tasks_per_stream;
  const long long t_begin = n_tasks * blockIdx.x / gridDim.x;
  const long long t_end = n_tasks * (blockIdx.x + 1) / gridDim.x;

  for (long long t = t_begin + warp; t < t_end; t += nwarps) {
    const int stream = (int)(t / tasks_per_stream);
    const int tq = (int)(t - (long long)stream * tasks_per_stream);
    const int b = stream / M, m = stream - b * M;
    const int q = tq * PPW + grp;
    const bool qvalid = q < Lq;
    const long long pair = ((long long)b * Lq + (qvalid ? q : Lq - 1)) * M + m;
    const float2* lp = reinterpret_cast<const float2*>(loc) + pair * LP;
    const float* ap = attn + pair * LP;
    const long long img_off = ((long long)b * S * M + m) * D + sub * VE;
    const T* vimg = value + img_off;
    float* gimg = gacc + img_off;

    float go[VE];
    Vec16<T>::load_stream(grad_out + pair * D + sub * VE, go);

    for (int l = 0; l < L; ++l) {
      const int H = sH[l], W = sW[l];
      const long long lvl_off = (long long)sStart[l] * MD;
      const T* vl = vimg + lvl_off;
      float* gl = gimg + lvl_off;
      const int rowstep = W * MD;
#pragma unroll 2
      for (int p = 0; p < P; ++p) {
        const float2 xy = __ldg(lp + l * P + p);
        const float a = __ldg(ap + l * P + p);
        Taps tp = make_taps(xy.x, xy.y, H, W, MD);
        tp.m00 = tp.m00 && qvalid;
        tp.m01 = tp.m01 && qvalid;
        tp.m10 = tp.m10 && qvalid;
        tp.m11 = tp.m11 && qvalid;
        const T* p00 = vl + tp.off00;
        float v00[VE], v01[VE], v10[VE], v11[VE];
        Vec16<T>::load(p00, tp.m00, v00);
        Vec16<T>::load(p00 + MD, tp.m01, v01);
        Vec16<T>::load(p00 + rowstep, tp.m10, v10);
        Vec16<T>::load(p00 + rowstep + MD, tp.m11, v11);

        const float w00 = tp.hy * tp.hx, w01 = tp.hy * tp.lx;
        const float w10 = tp.ly * tp.hx, w11 = tp.ly * tp.lx;
        float s_val = 0.f, s_gh = 0.f, s_gw = 0.f;
#pragma unroll
        for (int i = 0; i < VE; ++i) {
          const float g = go[i];
          const float val = fmaf(w00, v00[i], fmaf(w01, v01[i], fmaf(w10, v10[i], w11 * v11[i])));
          // grad_h_weight = -hw*v1 - lw*v2 + hw*v3 + lw*v4 ; grad_w_weight = -hh*v1 + hh*v2 - lh*v3 + lh*v4
          const float gh = fmaf(tp.hx, v10[i] - v00[i], tp.lx * (v11[i] - v01[i]));
          const float gw = fmaf(tp.hy, v01[i] - v00[i], tp.ly * (v11[i] - v10[i]));
          s_val = fmaf(g, val, s_val);
          s_gh = fmaf(g, gh, s_gh);
          s_gw = fmaf(g, gw, s_gw);
        }
        // grad_value taps: w_k * attn * grad_out  (cuh:118,130,139,148,157)
        float* g00 = gl + tp.off00;
        const float a00 = w00 * a, a01 = w01 * a, a10 = w10 * a, a11 = w11 * a;
#pragma unroll
        for (int i = 0; i < VE; i += 4) {
          if (tp.m00) red_add_v4(g00 + i, a00 * go[i], a00 * go[i + 1], a00 * go[i + 2], a00 * go[i + 3]);
          if (tp.m01) red_add_v4(g00 + MD + i, a01 * go[i], a01 * go[i + 1], a01 * go[i + 2], a01 * go[i + 3]);
          if (tp.m10) red_add_v4(g00 + rowstep + i, a10 * go[i], a10 * go[i + 1], a10 * go[i + 2], a10 * go[i + 3]);
          if (tp.m11) red_add_v4(g00 + rowstep + MD + i, a11 * go[i], a11 * go[i + 1], a11 * go[i + 2], a11 * go[i + 3]);
        }
        // reduce the three partial sums over the pair's LPP channel lanes
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) {
          s_val += __shfl_xor_sync(0xffffffffu, s_val, o);
          s_gh += __shfl_xor_sync(0xffffffffu, s_gh, o);
          s_gw += __shfl_xor_sync(0xffffffffu, s_gw, o);
        }
        if (sub == 0 && qvalid) {
          const long long s = pair * LP + l * P + p;
          grad_attn[s] = s_val;                                                     // cuh:161
          reinterpret_cast<float2*>(grad_loc)[s] =
              make_float2((float)W * s_gw * a, (float)H * s_gh * a);               // cuh:162-163
        }
      }
    }
  }
}

// fp32 scratch image -> bf16 grad_value (bf16 path only)
__global__ void msda_round_to_bf16_kernel(const float* __restrict__ src,
                                          __nv_bfloat16* __restrict__ dst, long long n8) {
  // n8 = number of 8-element groups
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    Vec16<__nv_bfloat16>::store(dst + 8 * i, v);
  }
}

__global__ void msda_round_to_bf16_tail_kernel(const float* __restrict__ src,
                                               __nv_bfloat16* __restrict__ dst, long long begin,
                                               long long n) {
  const long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

// ---------------------------------------------------------------------------------------------
// Generic path: any D, any L, fp32 / fp64 / bf16.  One warp per (query, head) pair, lanes stride
// over the channels; same stream ordering.  CT = compute / location type, AT = grad_value
// accumulator type in global memory.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Scalar {
  template <typename CT>
  static __device__ __forceinline__ CT load(const T* p) { return (CT)(*p); }
  template <typename CT>
  static __device__ __forceinline__ void store(T* p, CT v) { *p = (T)v; }
};
template <>
struct Scalar<__nv_bfloat16> {
  template <typename CT>
  static __device__ __forceinline__ CT load(const __nv_bfloat16* p) { return (CT)__bfloat162float(*p); }
  template <typename CT>
  static __device__ __forceinline__ void store(__nv_bfloat16* p, CT v) { *p = __float2bfloat16_rn((float)v); }
};

template <typename CT>
struct TapsG {
  long long off[4];
  bool ok[4];
  CT w[4];
  CT ly, lx, hy, hx;
};

template <typename CT>
__device__ __forceinline__ CT fma_coord(CT loc, int size);
template <>
__device__ __forceinline__ float fma_coord<float>(float loc, int size) { return fmaf(loc, (float)size, -0.5f); }
template <>
__device__ __forceinline__ double fma_coord<double>(double loc, int size) { return loc * (double)size - 0.5; }

template <typename CT>
__device__ __forceinline__ TapsG<CT> make_taps_g(CT loc_x, CT loc_y, int H, int W, long long MD) {
  TapsG<CT> t;
  const CT y = fma_coord<CT>(loc_y, H), x = fma_coord<CT>(loc_x, W);
  const bool inside = (y > (CT)-1) && (x > (CT)-1) && (y < (CT)H) && (x < (CT)W);
  const CT fy = floor(y), fx = floor(x);
  const int y0 = inside ? (int)fy : 0, x0 = inside ? (int)fx : 0;
  t.ly = y - fy; t.lx = x - fx; t.hy = (CT)1 - t.ly; t.hx = (CT)1 - t.lx;
  const bool top = inside && y0 >= 0, bot = inside && y0 + 1 <= H - 1;
  const bool lef = x0 >= 0, rig = x0 + 1 <= W - 1;
  t.ok[0] = top && lef; t.ok[1] = top && rig; t.ok[2] = bot && lef; t.ok[3] = bot && rig;
  const long long base = ((long long)y0 * W + x0) * MD;
  t.off[0] = base; t.off[1] = base + MD; t.off[2] = base + (long long)W * MD; t.off[3] = t.off[2] + MD;
  t.w[0] = t.hy * t.hx; t.w[1] = t.hy * t.lx; t.w[2] = t.ly * t.hx; t.w[3] = t.ly * t.lx;
  return t;
}

template <typename T, typename CT>
__global__ void __launch_bounds__(kGenericThreads)
msda_fwd_generic_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ lsi, const CT* __restrict__ loc,
                        const CT* __restrict__ attn, T* __restrict__ out, int N, int S, int M,
                        int D, int L, int Lq, int P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long long MD = (long long)M * D;
  const long long LP = (long long)L * P;
  const long long n_tasks = (long long)N * M * Lq;
  const long long t_begin = n_tasks / gridDim.x * blockIdx.x +
                            (n_tasks % gridDim.x < blockIdx.x ? n_tasks % gridDim.x : blockIdx.x);
  const long long t_end = t_begin + n_tasks / gridDim.x + (blockIdx.x < n_tasks % gridDim.x ? 1 : 0);
  for (long long t = t_begin + warp; t < t_end; t += nwarps) {
    const long long stream = t / Lq;
    const int q = (int)(t - stream * Lq);
    const int b = (int)(stream / M), m = (int)(stream - (long long)b * M);
    const long long pair = ((long long)b * Lq + q) * M + m;
    const T* vimg = value + ((long long)b * S * M + m) * D;
    for (int c0 = 0; c0 < D; c0 += 32) {
      const int c = c0 + lane;
      const bool cok = c < D;
      CT acc = (CT)0;
      for (int l = 0; l < L; ++l) {
        const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
        const T* vl = vimg + (long long)lsi[l] * MD;
        for (int p = 0; p < P; ++p) {
          const long long s = pair * LP + (long long)l * P + p;
          const TapsG<CT> tp = make_taps_g<CT>(loc[2 * s], loc[2 * s + 1], H, W, MD);
          const CT a = attn[s];
          CT val = (CT)0;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (tp.ok[k] && cok) val += tp.w[k] * Scalar<T>::template load<CT>(vl + tp.off[k] + c);
          acc += val * a;
        }
      }
      if (cok) Scalar<T>::template store<CT>(out + pair * D + c, acc);
    }
  }
}

template <typename T, typename CT, typename AT>
__global__ void __launch_bounds__(kGenericThreads)
msda_bwd_generic_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ lsi, const CT* __restrict__ loc,
                        const CT* __restrict__ attn, const T* __restrict__ grad_out,
                        AT* __restrict__ gacc, CT* __restrict__ grad_loc,
                        CT* __restrict__ grad_attn, int N, int S, int M, int D, int L, int Lq,
                        int P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long long MD = (long long)M * D;
  const long long LP = (long long)L * P;
  const long long n_tasks = (long long)N * M * Lq;
  const long long t_begin = n_tasks / gridDim.x * blockIdx.x +
                            (n_tasks % gridDim.x < blockIdx.x ? n_tasks % gridDim.x : blockIdx.x);
  const long long t_end = t_begin + n_tasks / gridDim.x + (blockIdx.x < n_tasks % gridDim.x ? 1 : 0);
  for (long long t = t_begin + warp; t < t_end; t += nwarps) {
    const long long stream = t / Lq;
    const int q = (int)(t - stream * Lq);
    const int b = (int)(stream / M), m = (int)(stream - (long long)b * M);
    const long long pair = ((long long)b * Lq + q) * M + m;
    const long long img_off = ((long long)b * S * M + m) * D;
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const long long lvl_off = img_off + (long long)lsi[l] * MD;
      for (int p = 0; p < P; ++p) {
        const long long s = pair * LP + (long long)l * P + p;
        const TapsG<CT> tp = make_taps_g<CT>(loc[2 * s], loc[2 * s + 1], H, W, MD);
        const CT a = attn[s];
        CT s_val = (CT)0, s_gh = (CT)0, s_gw = (CT)0;
        for (int c = lane; c < D; c += 32) {
          const CT g = Scalar<T>::template load<CT>(grad_out + pair * D + c);
          CT v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            v[k] = (CT)0;
            if (tp.ok[k]) {
              v[k] = Scalar<T>::template load<CT>(value + lvl_off + tp.off[k] + c);
              atomicAdd(gacc + lvl_off + tp.off[k] + c, (AT)(tp.w[k] * a * g));
            }
          }
          const CT val = tp.w[0] * v[0] + tp.w[1] * v[1] + tp.w[2] * v[2] + tp.w[3] * v[3];
          const CT gh = -tp.hx * v[0] - tp.lx * v[1] + tp.hx * v[2] + tp.lx * v[3];
          const CT gw = -tp.hy * v[0] + tp.hy * v[1] - tp.ly * v[2] + tp.ly * v[3];
          s_val += g * val;
          s_gh += g * gh;
          s_gw += g * gw;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s_val += __shfl_xor_sync(0xffffffffu, s_val, o);
          s_gh += __shfl_xor_sync(0xffffffffu, s_gh, o);
          s_gw += __shfl_xor_sync(0xffffffffu, s_gw, o);
        }
        if (lane == 0) {
          grad_attn[s] = s_val;
          grad_loc[2 * s] = (CT)W * s_gw * a;
          grad_loc[2 * s + 1] = (CT)H * s_gh * a;
        }
      }
    }
  }
}

}  // namespace pct
