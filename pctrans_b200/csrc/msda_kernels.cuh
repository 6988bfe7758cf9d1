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
//  * A (query, head) pair is handled by LPP lanes, each owning 16 or 32 contiguous bytes of the
//    head's channels, so every bilinear tap is one 128-bit or 256-bit load per lane (LDG.E.128 /
//    LDG.E.ENL2.256) and a warp covers 32/LPP pairs.
//  * The per-sample scalar work (pixel coordinates, in-range predicate, bilinear weights, corner
//    addresses) is done by ONE lane per sample and broadcast with warp shuffles, not recomputed
//    by every channel lane.
//  * grad_value leaves the SM as 128-bit vector reductions (red.global.add.v4.f32) instead of
//    per-channel scalar atomics; grad_sampling_loc / grad_attn_weight come from four per-corner
//    dot products reduced over the LPP channel lanes with xor-shuffles and are written coalesced
//    by the lane that owns the sample (no __syncthreads in either kernel's main loop).
//  * The generic kernels (any D, fp64, L*P > kMaxFastSamples) keep the same stream ordering with
//    one warp per pair.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

// Debug build (-DPCT_MSDA_BOUNDS_CHECK, libpctrans_msda_dbg.so): every tap address of the vector
// kernels is checked against the image extent and the kernel traps on a violation.  This stands
// in for compute-sanitizer, which is closed on the B200 pool; tests/test_gpu_bounds.py runs it.
#ifdef PCT_MSDA_BOUNDS_CHECK
#define PCT_CHECK_TAP(o, ox, oy, lane_off, ve, img_elems)                                        \
  do {                                                                                           \
    const long long lo_ = (long long)(o) + (lane_off);                                           \
    const long long hi_ = lo_ + (ox) + (oy) + (ve);                                              \
    if ((o) < 0 || (ox) < 0 || (oy) < 0 || hi_ > (long long)(img_elems)) __trap();               \
  } while (0)
#else
#define PCT_CHECK_TAP(o, ox, oy, lane_off, ve, img_elems) do { } while (0)
#endif

namespace pct {

// Tuning overrides (scripts/build_variants.sh): -DPCT_FWD_THREADS=..., -DPCT_BWD_THREADS=...
#ifndef PCT_FWD_THREADS
#define PCT_FWD_THREADS 1024
#endif
#ifndef PCT_BWD_THREADS
#define PCT_BWD_THREADS 768
#endif
#ifndef PCT_BWD_FUSED_THREADS
#define PCT_BWD_FUSED_THREADS 640
#endif
// samples per branch-free block of the tap loops (tuning: 1 = test every sample)
#ifndef PCT_TAP_BLOCK
#define PCT_TAP_BLOCK 4
#endif
#ifndef PCT_PACK_TAPS
#define PCT_PACK_TAPS 1
#endif
// forward: taps travel from the lane that computed them to the lanes of its pair through a small
// per-warp shared-memory patch (2 broadcast loads per sample) instead of 5 warp shuffles -- used
// when a pair has 8 or more lanes (fp32 D >= 32: 223 -> 213 us; with 4-lane pairs the shuffles
// are as fast or faster, profiles/r02_fwd_smem_taps.log)
#ifndef PCT_FWD_SMEM_TAPS
#define PCT_FWD_SMEM_TAPS 1
#endif
#ifndef PCT_TAP_BLOCK_BWD
#define PCT_TAP_BLOCK_BWD 4
#endif
constexpr int kBwdFusedThreads = PCT_BWD_FUSED_THREADS;  // fused backward carries more state
constexpr int kVecThreads = PCT_FWD_THREADS;     // forward: one persistent 32-warp CTA per SM
constexpr int kBwdVecThreads = PCT_BWD_THREADS;  // backward needs ~80 registers per thread
constexpr int kGenericThreads = 256;

// ---------------------------------------------------------------------------------------------
// Channel vectors: one lane owns VB contiguous bytes of a (pixel, head) row, moved with one
// 64-bit, 128-bit or 256-bit (LDG.E.ENL2.256 on sm_100a) instruction.
// ---------------------------------------------------------------------------------------------
template <typename T, int VB>
struct ChanVec;

template <>
struct ChanVec<float, 16> {
  static constexpr int kElems = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void load_stream(const float* p, float (&v)[4]) {
    const float4 r = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void store_stream(float* p, const float (&v)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};

template <>
struct ChanVec<float, 32> {
  static constexpr int kElems = 8;
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]),
                   "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
  }
  static __device__ __forceinline__ void load_stream(const float* p, float (&v)[8]) {
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]),
                   "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
  }
  static __device__ __forceinline__ void store_stream(float* p, const float (&v)[8]) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]),
                 "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
  }
};

template <>
struct ChanVec<__nv_bfloat16, 16> {
  static constexpr int kElems = 8;
  static __device__ __forceinline__ void unpack(const uint4& r, float (&v)[8]) {
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    unpack(__ldg(reinterpret_cast<const uint4*>(p)), v);
  }
  static __device__ __forceinline__ void load_stream(const __nv_bfloat16* p, float (&v)[8]) {
    unpack(__ldcs(reinterpret_cast<const uint4*>(p)), v);
  }
  static __device__ __forceinline__ void store_stream(__nv_bfloat16* p, const float (&v)[8]) {
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

template <>
struct ChanVec<__nv_bfloat16, 8> {
  static constexpr int kElems = 4;
  static __device__ __forceinline__ void unpack(const uint2& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x << 16);
    v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16);
    v[3] = __uint_as_float(r.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    unpack(__ldg(reinterpret_cast<const uint2*>(p)), v);
  }
  static __device__ __forceinline__ void load_stream(const __nv_bfloat16* p, float (&v)[4]) {
    unpack(__ldcs(reinterpret_cast<const uint2*>(p)), v);
  }
  static __device__ __forceinline__ void store_stream(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 r;
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
    r.x = *reinterpret_cast<unsigned*>(&lo);
    r.y = *reinterpret_cast<unsigned*>(&hi);
    __stcs(reinterpret_cast<uint2*>(p), r);
  }
};

// scalar element access with conversion (generic kernels, deterministic helpers)
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

struct DetScratch {       // tail of the deterministic / scaled-fp16 workspace
  unsigned int max_bits;  // bit pattern of max |grad_out| (non-negative floats order like ints)
  float scale;            // power of two the scattered products are multiplied with
  unsigned int rowsum_bits;   // scaled-fp16 path: bit pattern of max over (b,q,m) of sum_{l,p} |attn|
  unsigned int pad;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// bf16 accumulator (opt-in, PCT_MSDA_FLAG_BF16_ACCUM): four channels as one 8-byte packed reduction.
// Halves the L2 sector operations of the scatter, but every reduction rounds the running sum to
// bf16 (~48 roundings per pixel), so it is not the default.
__device__ __forceinline__ void red_add_v4(__nv_bfloat16* p, float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1, %2};" ::"l"(p),
               "r"(*reinterpret_cast<const unsigned*>(&lo)),
               "r"(*reinterpret_cast<const unsigned*>(&hi))
               : "memory");
}

// Scaled fp16 accumulator (bf16 values, opt-in PCT_MSDA_FLAG_BF16_F16_ACCUM): four channels as one 8-byte packed reduction --
// half the L2 sector operations of the fp32 scratch image, which is what bounds the per-tap
// backward.  The products arrive pre-multiplied by a power-of-two scale chosen from max|grad_out|,
// max sum|attn| and Lq so that no sum can overflow fp16 (msda_half_scale_kernel); fp16 keeps 11
// significand bits in the running sum where the bf16 accumulator keeps 8.
__device__ __forceinline__ void red_add_v4(__half* p, float a, float b, float c, float d) {
  const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
  asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(p),
               "r"(*reinterpret_cast<const unsigned*>(&lo)),
               "r"(*reinterpret_cast<const unsigned*>(&hi))
               : "memory");
}

__device__ __forceinline__ void red_add_v4(long long*, float, float, float, float) {}  // unused (kDet path scatters itself)

template <typename ACC> struct AccScaled { static constexpr bool value = false; };
template <> struct AccScaled<__half> { static constexpr bool value = true; };

// ---------------------------------------------------------------------------------------------
// Bilinear taps (cuh:45-58 geometry, cuh:61-83 per-corner bounds, cuh:293 in-range predicate).
//
// One lane computes the taps of ONE sample and broadcasts them to the lanes that share its
// (query, head) pair, instead of every channel lane recomputing them.  Taps are made branch-free
// for the consumers: the four corner addresses are clamped into the map (always loadable) and a
// corner that the reference would skip gets weight 0, so it contributes exactly nothing (for
// finite values) and its grad_value update is skipped.
// ---------------------------------------------------------------------------------------------
struct SampleTaps {
  int off;   // element offset (inside one image, head 0, channel 0) of the clamped (y0, x0) corner
  int dx;    // element step to the x+1 corner (0 when clamped onto the same pixel)
  int dy;    // element step to the y+1 corner
  float ly, lx;          // fractional parts
  bool v00, v01, v10, v11;  // corner is sampled by the reference
};

// level table entry = {H, W, level_start, level index}
__device__ __forceinline__ SampleTaps make_sample_taps(float loc_x, float loc_y, int H, int W,
                                                       int start, int MD) {
  SampleTaps t;
  // cuh:290-291 `loc * size - 0.5`: the product is rounded to float (float * int), the literal
  // 0.5 promotes the subtraction to double and the assignment rounds once more -- bit for bit a
  // float multiply followed by a float subtract (NOT a fused multiply-add).
  const float y = __fsub_rn(__fmul_rn(loc_y, (float)H), 0.5f);
  const float x = __fsub_rn(__fmul_rn(loc_x, (float)W), 0.5f);
  const bool inside = (y > -1.f) && (x > -1.f) && (y < (float)H) && (x < (float)W);
  const float fy = floorf(y), fx = floorf(x);
  t.ly = y - fy;
  t.lx = x - fx;
  const int y0 = inside ? (int)fy : 0, x0 = inside ? (int)fx : 0;
  const bool top = inside && (y0 >= 0), bot = inside && (y0 + 1 <= H - 1);
  const bool lef = (x0 >= 0), rig = (x0 + 1 <= W - 1);
  t.v00 = top && lef;
  t.v01 = top && rig;
  t.v10 = bot && lef;
  t.v11 = bot && rig;
  // inside => H >= 1 and W >= 1, so the clamps below stay inside the map
  const int y0c = max(y0, 0), x0c = max(x0, 0);
  const int y1c = min(y0 + 1, H - 1), x1c = min(x0 + 1, W - 1);
  t.off = inside ? (start + y0c * W + x0c) * MD : 0;
  t.dx = inside ? (x1c - x0c) * MD : 0;
  t.dy = inside ? (y1c - y0c) * W * MD : 0;
  return t;
}

constexpr int kMaxFastSamples = 128;  // L*P limit of the vector path (level table in smem)

__device__ __forceinline__ void fill_level_table(int4* tab, const int64_t* shapes,
                                                 const int64_t* lsi, int L, int P) {
  for (int s = threadIdx.x; s < L * P; s += blockDim.x) {
    const int l = s / P;
    tab[s] = make_int4((int)shapes[2 * l], (int)shapes[2 * l + 1], (int)lsi[l], l);
  }
}

// ---------------------------------------------------------------------------------------------
// Task scheduling of the persistent vector kernels: "static head, dynamic tail".
// Every CTA first walks a contiguous static slice of the stream-ordered task list (this is what
// keeps the L1 sliding window hot), covering `static_per_cta` tasks; the remaining tasks are
// handed out in mini-chunks of kMiniChunk consecutive tasks per warp through a global counter, so
// SMs that run faster (near-die L2, no co-running NCCL kernel) absorb the tail instead of idling
// -- static partitioning alone left 8 % (forward) / 11 % (backward) of the SM time idle.
// The counters live in a ring of module-scope device words; the last CTA to leave resets its slot.
// ---------------------------------------------------------------------------------------------
constexpr int kSchedSlots = 512;
#ifndef PCT_MINI_CHUNK
#define PCT_MINI_CHUNK 2
#endif
constexpr int kMiniChunk = PCT_MINI_CHUNK;
// one ring per translation unit that launches persistent kernels (internal linkage)
static __device__ unsigned int g_sched[kSchedSlots * 2];

struct SchedArgs {
  int slot;            // ring slot of this launch
  int static_per_cta;  // tasks of the static slice of every CTA
};

struct TaskIter {
  int t, t_static_end, n_tasks, dyn_begin, nwarps, lane, chunk_left;
  unsigned int* ctr;
  __device__ __forceinline__ TaskIter(const SchedArgs& sa, int n_tasks_, int warp, int nwarps_,
                                      int lane_)
      : n_tasks(n_tasks_), nwarps(nwarps_), lane(lane_), chunk_left(0) {
    const int base = blockIdx.x * sa.static_per_cta;
    t = base + warp - nwarps_;
    t_static_end = base + sa.static_per_cta;
    dyn_begin = gridDim.x * sa.static_per_cta;
    ctr = g_sched + 2 * sa.slot;
  }
  // next task index for this warp, or -1 when the work is exhausted (warp-uniform)
  __device__ __forceinline__ int next() {
    if (t_static_end >= 0) {                 // static phase
      t += nwarps;
      if (t < t_static_end) return t;
      t_static_end = -1;
    }
    if (chunk_left == 0) {                   // dynamic phase: grab kMiniChunk consecutive tasks
      unsigned int c = 0;
      if (lane == 0) c = atomicAdd(ctr, 1u);
      c = __shfl_sync(0xffffffffu, c, 0);
      const long long first = (long long)dyn_begin + (long long)c * kMiniChunk;
      if (first >= n_tasks) return -1;
      t = (int)first - 1;
      chunk_left = (int)((n_tasks - first) < kMiniChunk ? (n_tasks - first) : kMiniChunk);
    }
    --chunk_left;
    return ++t;
  }
  // call once per CTA after all warps are done (after a __syncthreads)
  __device__ __forceinline__ void finish() {
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(ctr + 1, 1u) == gridDim.x - 1) {
        ctr[0] = 0u;
        ctr[1] = 0u;
        __threadfence();
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Fused prologue.  With FUSED = true the kernels consume what the module's
// Linears produce instead of materialised locations / weights:
//   `loc`  = raw sampling offsets (N,Lq,M,L,P,2),  `attn` = raw attention logits (N,Lq,M,L*P),
//   `ref`  = reference points (N,Lq,L,ref_dim) with ref_dim 2 (points) or 4 (boxes cx,cy,w,h),
// and evaluate  loc = ref + off/(W_l,H_l)            (mod.py:106-109)
//          or   loc = ref.xy + off/P * ref.wh * 0.5  (mod.py:110-112)
// and the softmax over the L*P logits of a (query, head) pair (mod.py:104) in registers.
// ---------------------------------------------------------------------------------------------
struct FusedRef {
  const float* ref;  // nullptr when not fused
  int ref_dim;       // 2 or 4
};

template <int LPP>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int w = LPP / 2; w > 0; w >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, w));
  return v;
}
template <int LPP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int w = LPP / 2; w > 0; w >>= 1) v += __shfl_xor_sync(0xffffffffu, v, w);
  return v;
}

// sampling location of sample s from the raw offset (fused) -- same operation order as the module
__device__ __forceinline__ float2 fused_location(const float* __restrict__ rp, int ref_dim, int l,
                                                 float2 off, int H, int W, int P) {
  if (ref_dim == 2) {
    const float2 r = __ldg(reinterpret_cast<const float2*>(rp) + l);
    return make_float2(r.x + __fdiv_rn(off.x, (float)W), r.y + __fdiv_rn(off.y, (float)H));
  }
  const float4 r = __ldg(reinterpret_cast<const float4*>(rp) + l);
  return make_float2(r.x + __fdiv_rn(off.x, (float)P) * r.z * 0.5f,
                     r.y + __fdiv_rn(off.y, (float)P) * r.w * 0.5f);
}

// four fp32 channels of a staged value row (STAGED forward, fp32 values with 16-byte lanes only)
__device__ __forceinline__ void load_shared_f4(unsigned addr, float (&v)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
               : "r"(addr));
}

// ---------------------------------------------------------------------------------------------
// Forward, vector path.  T = float | bf16; LPP lanes per (query, head) pair, each owning VB
// bytes of the head's channels: D = LPP * VB/sizeof(T); a warp covers 32/LPP pairs.
// ---------------------------------------------------------------------------------------------
// STAGED (experiment of round 2, PCT_MSDA_FLAG_FWD_STAGED; BASELINE.json's north_star: "per-level value
// tiles staged in shared memory"): before a CTA walks the queries of one (image, head) stream it
// copies the value rows of that head for the smallest levels -- as many whole levels as fit the
// dynamic shared memory, a contiguous pixel range -- into shared memory (cp.async) and takes the
// taps of those levels with LDS.128 instead of LDG.128.  The static slice of a CTA is walked in
// per-stream segments (CTA-wide barriers around the copy); the dynamic tail reads global memory.
template <typename T, int LPP, int VB, bool FUSED, bool STAGED = false>
__global__ void __launch_bounds__(kVecThreads, 1)
msda_fwd_vec_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                    const float* __restrict__ attn, T* __restrict__ out, int N, int S, int M,
                    int L, int Lq, int P, FusedRef fr, SchedArgs sched, int stage_cap_pix = 0) {
  using CV = ChanVec<T, VB>;
  constexpr int VE = CV::kElems;
  constexpr int D = LPP * VE;
  constexpr int PPW = 32 / LPP;  // pairs per warp
  __shared__ int4 s_tab[kMaxFastSamples];
  constexpr bool kSmemTaps = PCT_FWD_SMEM_TAPS && PCT_PACK_TAPS && LPP >= 8;
#if PCT_FWD_SMEM_TAPS
  // Tap exchange patch of this warp: slot (pair, sample-of-round) holds {w00,w01,w10,w11} and the
  // packed corner word.  A pair's slots are padded to LPP+1 so that the PPW pairs of a warp read
  // different banks (the LPP lanes of a pair read the same slot: a broadcast).
  constexpr int kSlots = kSmemTaps ? PPW * (LPP + 1) : 1;
  __shared__ float4 s_tapw[kSmemTaps ? kVecThreads / 32 : 1][kSlots];
  __shared__ int s_tapo[kSmemTaps ? kVecThreads / 32 : 1][kSlots];
#endif
  fill_level_table(s_tab, shapes, lsi, L, P);
  extern __shared__ __align__(16) unsigned char s_tile[];   // STAGED: value rows of one head
  __shared__ int s_stage[3];                                 // first pixel, pixels, level mask
  if (STAGED && threadIdx.x == 0) {
    // whole levels from the front or from the back of the pixel order, whichever covers more levels
    int npre = 0, lpre = 0, nsuf = 0, lsuf = 0;
    for (int l = 0; l < L; ++l) {
      const int n = (int)(shapes[2 * l] * shapes[2 * l + 1]);
      if (npre + n > stage_cap_pix) break;
      npre += n; ++lpre;
    }
    for (int l = L - 1; l >= 0; --l) {
      const int n = (int)(shapes[2 * l] * shapes[2 * l + 1]);
      if (nsuf + n > stage_cap_pix) break;
      nsuf += n; ++lsuf;
    }
    const bool pre = lpre > lsuf || (lpre == lsuf && npre >= nsuf);
    s_stage[0] = pre ? 0 : S - nsuf;
    s_stage[1] = pre ? npre : nsuf;
    s_stage[2] = pre ? ((1 << lpre) - 1) : (((1 << lsuf) - 1) << (L - lsuf));
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int grp = lane / LPP, sub = lane % LPP;
  const int MD = M * D;
  const int LP = L * P;
  const int tasks_per_stream = (Lq + PPW - 1) / PPW;
  // the launcher guarantees n_tasks < 2^31, so the per-task index math stays 32-bit
  const int n_tasks = N * M * tasks_per_stream;
  TaskIter iter(sched, n_tasks, warp, nwarps, lane);
  // STAGED state: the stream whose rows sit in s_tile, the shift that turns an element offset
  // (multiple of M*D; the launcher guarantees M is a power of two) into pixel*D, this lane's base
  const int stage_pix0 = STAGED ? s_stage[0] : 0, stage_npix = STAGED ? s_stage[1] : 0;
  const unsigned stage_lmask = STAGED ? (unsigned)s_stage[2] : 0u;
  const int sh_m = STAGED ? __ffs(M) - 1 : 0;
  const unsigned tile_lane = (unsigned)__cvta_generic_to_shared(s_tile) + (unsigned)(sub * VB) -
                             (unsigned)(stage_pix0 * D * (int)sizeof(T));
  int staged_stream = -1;
#if PCT_FWD_SMEM_TAPS
  float4* const my_w = &s_tapw[kSmemTaps ? warp : 0][kSmemTaps ? grp * (LPP + 1) : 0];
  int* const my_o = &s_tapo[kSmemTaps ? warp : 0][kSmemTaps ? grp * (LPP + 1) : 0];
#endif

  auto process = [&](const int t) {
    const int stream = t / tasks_per_stream;
    const int tq = t - stream * tasks_per_stream;
    const int b = stream / M, m = stream - b * M;
    const int q = tq * PPW + grp;
    const bool qvalid = q < Lq;
    const bool in_tile = STAGED && stream == staged_stream;     // warp-uniform
    const long long pair = ((long long)b * Lq + (qvalid ? q : Lq - 1)) * M + m;
    const float2* lp = reinterpret_cast<const float2*>(loc) + pair * LP;
    const float* ap = attn + pair * LP;
    // The lane's image base as an opaque 64-bit register pair: a corner address is then one
    // IMAD.WIDE (base + 32-bit element offset) instead of a sign-extended 64-bit add + LEA pair.
    unsigned long long vbase = reinterpret_cast<unsigned long long>(
        value + ((long long)b * S * M + m) * D + sub * VE);
    asm volatile("" : "+l"(vbase));
    const T* vimg = reinterpret_cast<const T*>(vbase);

    float acc[VE];
#pragma unroll
    for (int i = 0; i < VE; ++i) acc[i] = 0.f;

    // fused softmax: stabilise with the pair's max logit, accumulate with un-normalised weights
    // exp(logit - max) and divide the output once by their sum
    float lmax = 0.f, esum = 0.f;
    const float* rp = nullptr;
    // up to three rounds of this lane's logits stay in registers (L*P <= 3*LPP: every PCTrans
    // shape); longer sample lists re-read them through L1
    float lg0 = -INFINITY, lg1 = -INFINITY, lg2 = -INFINITY;
    const bool lg_regs = FUSED && (LP <= 3 * LPP);
    if (FUSED) {
      rp = fr.ref + ((long long)b * Lq + (qvalid ? q : Lq - 1)) * L * fr.ref_dim;
      float mx = -INFINITY;
      if (lg_regs) {
        if (sub < LP) lg0 = __ldg(ap + sub);
        if (sub + LPP < LP) lg1 = __ldg(ap + sub + LPP);
        if (sub + 2 * LPP < LP) lg2 = __ldg(ap + sub + 2 * LPP);
        mx = fmaxf(lg0, fmaxf(lg1, lg2));
      } else {
        for (int s = sub; s < LP; s += LPP) mx = fmaxf(mx, __ldg(ap + s));
      }
      lmax = group_max<LPP>(mx);
    }

    // software pipeline: the location / weight of this lane's sample of the NEXT round
    float2 xy_n = make_float2(0.f, 0.f);
    float a_n = 0.f;
    if (sub < LP) {
      xy_n = __ldcs(lp + sub);
      a_n = lg_regs ? lg0 : __ldcs(ap + sub);
    }
    for (int s0 = 0; s0 < LP; s0 += LPP) {
      const int s = s0 + sub;
      float2 xy = xy_n;
      float a = a_n;
      if (s + LPP < LP) {
        xy_n = __ldcs(lp + s + LPP);
        a_n = lg_regs ? (s0 == 0 ? lg1 : lg2) : __ldcs(ap + s + LPP);
      }
      int off = 0, dx = 0, dy = 0;
      float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
      if (s < LP) {
        const int4 tb = s_tab[s];
        if (FUSED) {
          xy = fused_location(rp, fr.ref_dim, tb.w, xy, tb.x, tb.y, P);
          a = expf(a - lmax);
          esum += a;
        }
        const SampleTaps tp = make_sample_taps(xy.x, xy.y, tb.x, tb.y, tb.z, MD);
        const float hy = 1.f - tp.ly, hx = 1.f - tp.lx;
        off = tp.off; dx = tp.dx; dy = tp.dy;
        w00 = tp.v00 ? hy * hx * a : 0.f;
        w01 = tp.v01 ? hy * tp.lx * a : 0.f;
        w10 = tp.v10 ? tp.ly * hx * a : 0.f;
        w11 = tp.v11 ? tp.ly * tp.lx * a : 0.f;
      }
      // one sample of the round: lane j of every pair broadcasts its taps, all lanes gather.
      // Shuffles cost one L1 data-pipe wavefront each (ncu: 31 % of this kernel's wavefronts are
      // shuffles, and the pipe is what bounds it), so the three corner offsets travel as ONE word:
      // off is a multiple of M*D (D >= 4), its two low bits say whether x+1 / y+1 is another pixel;
      // the x step is M*D and the y step is the level's row pitch, read once per block.
#if PCT_PACK_TAPS
      const int word = off | (dx != 0 ? 1 : 0) | (dy != 0 ? 2 : 0);
#endif
#if PCT_FWD_SMEM_TAPS && PCT_PACK_TAPS
      if (kSmemTaps) {
        __syncwarp();                                 // the previous round's reads are done
        my_w[sub] = make_float4(w00, w01, w10, w11);
        my_o[sub] = word;
        __syncwarp();
      }
#endif
      auto tap = [&](const int j, const int pitch, const bool from_tile = false) {
#if PCT_FWD_SMEM_TAPS && PCT_PACK_TAPS
        int wd;
        float b00, b01, b10, b11;
        if (kSmemTaps) {
          wd = my_o[j];
          const float4 wv = my_w[j];
          b00 = wv.x; b01 = wv.y; b10 = wv.z; b11 = wv.w;
        } else {
          wd = __shfl_sync(0xffffffffu, word, j, LPP);
          b00 = __shfl_sync(0xffffffffu, w00, j, LPP);
          b01 = __shfl_sync(0xffffffffu, w01, j, LPP);
          b10 = __shfl_sync(0xffffffffu, w10, j, LPP);
          b11 = __shfl_sync(0xffffffffu, w11, j, LPP);
        }
        const int o = wd & ~3;
        const int ox = (wd & 1) ? MD : 0;
        const int oy = (wd & 2) ? pitch : 0;
#elif PCT_PACK_TAPS
        const int wd = __shfl_sync(0xffffffffu, word, j, LPP);
        const int o = wd & ~3;
        const int ox = (wd & 1) ? MD : 0;
        const int oy = (wd & 2) ? pitch : 0;
#else
        const int o = __shfl_sync(0xffffffffu, off, j, LPP);
        const int ox = __shfl_sync(0xffffffffu, dx, j, LPP);
        const int oy = __shfl_sync(0xffffffffu, dy, j, LPP);
#endif
#if !(PCT_FWD_SMEM_TAPS && PCT_PACK_TAPS)
        const float b00 = __shfl_sync(0xffffffffu, w00, j, LPP);
        const float b01 = __shfl_sync(0xffffffffu, w01, j, LPP);
        const float b10 = __shfl_sync(0xffffffffu, w10, j, LPP);
        const float b11 = __shfl_sync(0xffffffffu, w11, j, LPP);
#endif
        PCT_CHECK_TAP(o, ox, oy, (m * D + sub * VE), VE, (long long)S * MD);
        // 32-bit corner offsets first: each address is then ONE wide multiply-add off the base
        const int o01 = o + ox, o10 = o + oy, o11 = o10 + ox;
        float v00[VE], v01[VE], v10[VE], v11[VE];
        bool done = false;
        if constexpr (STAGED) {
          if (from_tile) {
            // the same four rows out of the staged tile: element offset / M = pixel * D
            const unsigned a00 = tile_lane + ((unsigned)o >> sh_m) * (unsigned)sizeof(T);
            const unsigned sx = ((unsigned)ox >> sh_m) * (unsigned)sizeof(T);
            const unsigned sy = ((unsigned)oy >> sh_m) * (unsigned)sizeof(T);
            load_shared_f4(a00, v00);
            load_shared_f4(a00 + sx, v01);
            load_shared_f4(a00 + sy, v10);
            load_shared_f4(a00 + sy + sx, v11);
            done = true;
          }
        }
        if (!done) {
          CV::load(vimg + o, v00);
          CV::load(vimg + o01, v01);
          CV::load(vimg + o10, v10);
          CV::load(vimg + o11, v11);
        }
#pragma unroll
        for (int i = 0; i < VE; ++i)
          acc[i] = fmaf(b00, v00[i], fmaf(b01, v01[i], fmaf(b10, v10[i], fmaf(b11, v11[i], acc[i]))));
      };
      // Samples are taken in blocks of kBlk with ONE warp-uniform test per block: inside a block
      // there are no branches, so the row loads of neighbouring samples are scheduled together
      // (the kernel is partly latency-bound); only a ragged tail is tested per sample.
      constexpr int kBlk = LPP < PCT_TAP_BLOCK ? LPP : PCT_TAP_BLOCK;
#pragma unroll
      for (int j0 = 0; j0 < LPP; j0 += kBlk) {
        if (s0 + j0 + kBlk <= LP) {
          if (P % kBlk == 0) {   // the block lies inside one level (s0 + j0 is a multiple of kBlk)
            const int pitch = PCT_PACK_TAPS ? s_tab[s0 + j0].y * MD : 0;
            if (STAGED && in_tile && ((stage_lmask >> s_tab[s0 + j0].w) & 1u)) {
#pragma unroll
              for (int j = 0; j < kBlk; ++j) tap(j0 + j, pitch, true);
            } else {
#pragma unroll
              for (int j = 0; j < kBlk; ++j) tap(j0 + j, pitch);
            }
          } else {
#pragma unroll
            for (int j = 0; j < kBlk; ++j) tap(j0 + j, PCT_PACK_TAPS ? s_tab[s0 + j0 + j].y * MD : 0);
          }
        } else {
#pragma unroll
          for (int j = 0; j < kBlk; ++j)
            if (s0 + j0 + j < LP) tap(j0 + j, PCT_PACK_TAPS ? s_tab[s0 + j0 + j].y * MD : 0);
        }
      }
    }
    if (FUSED) {
      const float inv = 1.f / group_sum<LPP>(esum);
#pragma unroll
      for (int i = 0; i < VE; ++i) acc[i] *= inv;
    }
    if (qvalid) CV::store_stream(out + pair * D + sub * VE, acc);
  };
  if (STAGED) {
    // the static slice in per-stream segments, the head's small levels staged before each
    const int seg_last = min(n_tasks, blockIdx.x * sched.static_per_cta + sched.static_per_cta);
    for (int seg = blockIdx.x * sched.static_per_cta; seg < seg_last;) {
      const int stream = seg / tasks_per_stream;
      const int seg_end = min(seg_last, (stream + 1) * tasks_per_stream);
      __syncthreads();                                      // the previous tile is no longer read
      {
        const int b = stream / M, m = stream - b * M;
        constexpr int PPR = D * (int)sizeof(T) / 16;        // 16-byte pieces per row
        const T* src = value + (((long long)b * S + stage_pix0) * M + m) * D;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s_tile);
        for (int i = threadIdx.x; i < stage_npix * PPR; i += blockDim.x) {
          const int pix = i / PPR, part = i - pix * PPR;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)i * 16u),
                       "l"(src + (long long)pix * MD + part * (16 / (int)sizeof(T)))
                       : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      staged_stream = stream;
      for (int t = seg + warp; t < seg_end; t += nwarps) process(t);
      seg = seg_end;
    }
    staged_stream = -1;
    iter.t_static_end = -1;                                 // only the dynamic tail is left
  }
  for (int t = iter.next(); t >= 0; t = iter.next()) process(t);
  __syncthreads();
  iter.finish();
}

// ---------------------------------------------------------------------------------------------
// Backward, vector path.  grad_value is accumulated in fp32 (`gacc`: grad_value itself for fp32
// values, the caller's workspace for bf16 values) with 128-bit vector reductions.  The launcher
// picks VB so that a lane owns 4 channels whenever it can: the LPP lanes of a pair then cover
// whole 32-byte sectors with ONE red.v4 each, and the L2 -- which retires reductions per sector,
// ~199 sectors/ns measured on B200 -- sees 4 sector operations per 128-byte tap instead of 8.  Per sample the
// channel lanes only form the four dot products <grad_out, corner value>; the lane that owns the
// sample combines them into grad_attn and grad_loc (cuh:128-163) and writes them coalesced.
// ---------------------------------------------------------------------------------------------
// With FUSED the kernel writes grad w.r.t. the raw offsets into `grad_loc` and grad w.r.t. the raw
// logits into `grad_attn` (softmax backward: a_s * (g_s - sum_t a_t g_t)).
template <typename T, int LPP, int VB, bool FUSED, typename ACC = float>
__global__ void __launch_bounds__(FUSED ? kBwdFusedThreads : kBwdVecThreads, 1)
msda_bwd_vec_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                    const float* __restrict__ attn, const T* __restrict__ grad_out,
                    ACC* __restrict__ gacc, float* __restrict__ grad_loc,
                    float* __restrict__ grad_attn, int N, int S, int M, int L, int Lq, int P,
                    FusedRef fr, SchedArgs sched, const DetScratch* det) {
  using CV = ChanVec<T, VB>;
  constexpr int VE = CV::kElems;
  constexpr int D = LPP * VE;
  constexpr int PPW = 32 / LPP;
  __shared__ int4 s_tab[kMaxFastSamples];
  fill_level_table(s_tab, shapes, lsi, L, P);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int grp = lane / LPP, sub = lane % LPP;
  const int MD = M * D;
  const int LP = L * P;
  const int tasks_per_stream = (Lq + PPW - 1) / PPW;
  // the launcher guarantees n_tasks < 2^31, so the per-task index math stays 32-bit
  const int n_tasks = N * M * tasks_per_stream;
  TaskIter iter(sched, n_tasks, warp, nwarps, lane);

  for (int t = iter.next(); t >= 0; t = iter.next()) {
    const int stream = t / tasks_per_stream;
    const int tq = t - stream * tasks_per_stream;
    const int b = stream / M, m = stream - b * M;
    const int q = tq * PPW + grp;
    const bool qvalid = q < Lq;
    const long long pair = ((long long)b * Lq + (qvalid ? q : Lq - 1)) * M + m;
    const float2* lp = reinterpret_cast<const float2*>(loc) + pair * LP;
    const float* ap = attn + pair * LP;
    const long long img_off = ((long long)b * S * M + m) * D + sub * VE;
    // opaque 64-bit bases: a corner address is one IMAD.WIDE off the base (see the forward)
    unsigned long long vbase = reinterpret_cast<unsigned long long>(value + img_off);
    unsigned long long gbase = reinterpret_cast<unsigned long long>(gacc + img_off);
    asm volatile("" : "+l"(vbase), "+l"(gbase));
    const T* vimg = reinterpret_cast<const T*>(vbase);
    ACC* gimg = reinterpret_cast<ACC*>(gbase);

    float go[VE];
    CV::load_stream(grad_out + pair * D + sub * VE, go);
    // deterministic mode: this lane scatters channels {i*LPP + sub} so that the LPP lanes of a
    // pair cover contiguous 8-byte accumulators (whole sectors) with each integer reduction
    constexpr bool kDet = sizeof(ACC) == 8;
    constexpr bool kScaled = AccScaled<ACC>::value;   // fp16 accumulator: scatter scale * grad_out
    float gs[VE];
    ACC* grow = gacc;
    float gsc[VE];
#pragma unroll
    for (int i = 0; i < VE; ++i) gsc[i] = go[i];
    if (kScaled) {
      const float scale = det->scale;
#pragma unroll
      for (int i = 0; i < VE; ++i) gsc[i] = go[i] * scale;
    }
    if (kDet) {
      const float scale = det->scale;
#pragma unroll
      for (int i = 0; i < VE; ++i)
        gs[i] = Scalar<T>::template load<float>(grad_out + pair * D + i * LPP + sub) * scale;
      grow = gacc + ((long long)b * S * M + m) * D + sub;
    }

    // fused softmax: max and normaliser of the pair's logits before the first tap
    float lmax = 0.f, inv_sum = 1.f, ag_dot = 0.f;
    const float* rp = nullptr;
    // this lane's logits of up to three rounds, then its softmax weights (pa) and the raw
    // d/d(attn) values (pg) of the same samples, kept for the softmax backward at the end
    float lg0 = -INFINITY, lg1 = -INFINITY, lg2 = -INFINITY;
    float pa0 = 0.f, pa1 = 0.f, pa2 = 0.f, pg0 = 0.f, pg1 = 0.f, pg2 = 0.f;
    const bool lg_regs = FUSED && (LP <= 3 * LPP);
    if (FUSED) {
      rp = fr.ref + ((long long)b * Lq + (qvalid ? q : Lq - 1)) * L * fr.ref_dim;
      float mx = -INFINITY, es = 0.f;
      if (lg_regs) {
        if (sub < LP) lg0 = __ldg(ap + sub);
        if (sub + LPP < LP) lg1 = __ldg(ap + sub + LPP);
        if (sub + 2 * LPP < LP) lg2 = __ldg(ap + sub + 2 * LPP);
        lmax = group_max<LPP>(fmaxf(lg0, fmaxf(lg1, lg2)));
        es = expf(lg0 - lmax) + expf(lg1 - lmax) + expf(lg2 - lmax);   // exp(-inf) = 0 for unused slots
      } else {
        for (int s = sub; s < LP; s += LPP) mx = fmaxf(mx, __ldg(ap + s));
        lmax = group_max<LPP>(mx);
        for (int s = sub; s < LP; s += LPP) es += expf(__ldg(ap + s) - lmax);
      }
      inv_sum = 1.f / group_sum<LPP>(es);
    }

    float2 xy_n = make_float2(0.f, 0.f);
    float a_n = 0.f;
    if (sub < LP) {
      xy_n = __ldcs(lp + sub);
      a_n = lg_regs ? lg0 : __ldcs(ap + sub);
    }
    for (int s0 = 0; s0 < LP; s0 += LPP) {
      const int s = s0 + sub;
      float2 xy = xy_n;
      float a = a_n;
      const float2 raw_off = xy_n;
      if (s + LPP < LP) {
        xy_n = __ldcs(lp + s + LPP);
        a_n = lg_regs ? (s0 == 0 ? lg1 : lg2) : __ldcs(ap + s + LPP);
      }
      // this lane's sample
      int off = 0, dx = 0, dy = 0;
      float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;   // bilinear weights, 0 for skipped corners
      float ch00 = 0.f, ch01 = 0.f, ch10 = 0.f, ch11 = 0.f;  // d(val)/d(y) coefficients
      float cw00 = 0.f, cw01 = 0.f, cw10 = 0.f, cw11 = 0.f;  // d(val)/d(x) coefficients
      float Hf = 0.f, Wf = 0.f;
      float2 dloc = make_float2(1.f, 1.f);   // d(location)/d(raw offset), fused only
      if (s < LP) {
        const int4 tb = s_tab[s];
        if (FUSED) {
          xy = fused_location(rp, fr.ref_dim, tb.w, raw_off, tb.x, tb.y, P);
          a = expf(a - lmax) * inv_sum;
          if (fr.ref_dim == 2) {
            dloc = make_float2(__fdiv_rn(1.f, (float)tb.y), __fdiv_rn(1.f, (float)tb.x));
          } else {
            const float4 r = __ldg(reinterpret_cast<const float4*>(rp) + tb.w);
            dloc = make_float2(__fdiv_rn(r.z * 0.5f, (float)P), __fdiv_rn(r.w * 0.5f, (float)P));
          }
        }
        const SampleTaps tp = make_sample_taps(xy.x, xy.y, tb.x, tb.y, tb.z, MD);
        const float hy = 1.f - tp.ly, hx = 1.f - tp.lx;
        off = tp.off; dx = tp.dx; dy = tp.dy;
        Hf = (float)tb.x; Wf = (float)tb.y;
        // grad_h_weight = -hw*v1 - lw*v2 + hw*v3 + lw*v4 ; grad_w_weight = -hh*v1 + hh*v2 - lh*v3 + lh*v4
        if (tp.v00) { w00 = hy * hx;       ch00 = -hx;    cw00 = -hy; }
        if (tp.v01) { w01 = hy * tp.lx;    ch01 = -tp.lx; cw01 = hy; }
        if (tp.v10) { w10 = tp.ly * hx;    ch10 = hx;     cw10 = -tp.ly; }
        if (tp.v11) { w11 = tp.ly * tp.lx; ch11 = tp.lx;  cw11 = tp.ly; }
      }
      const float ae = qvalid ? a : 0.f;   // padded tail pairs scatter nothing
      const float aw00 = w00 * ae, aw01 = w01 * ae, aw10 = w10 * ae, aw11 = w11 * ae;
      // Scatter weights of MY sample's four rows.  All samples of a pair scatter the SAME
      // grad_out vector, so two samples that touch the same value row can share one reduction:
      // the weights add.  Neighbouring points of a level (p, p+1) overlap by one or two rows
      // when the head walks in pixel steps (mod.py:66-74), which is 14-28 % of all reductions --
      // and the backward is bound by L2 reduction throughput.  Each lane hands the rows it
      // shares with the previous lane's sample to that lane and absorbs the rows the next
      // lane's sample shares with it (both sides evaluate the same match on the original
      // weights, so nothing is counted twice or lost).
      float rw0 = aw00, rw1 = aw01, rw2 = aw10, rw3 = aw11;
      if (LPP > 1) {
        const int r0 = off, r1 = off + dx, r2 = off + dy, r3 = off + dx + dy;
        const int n_off = __shfl_down_sync(0xffffffffu, off, 1, LPP);
        const int n_dx = __shfl_down_sync(0xffffffffu, dx, 1, LPP);
        const int n_dy = __shfl_down_sync(0xffffffffu, dy, 1, LPP);
        const float nw0 = __shfl_down_sync(0xffffffffu, aw00, 1, LPP);
        const float nw1 = __shfl_down_sync(0xffffffffu, aw01, 1, LPP);
        const float nw2 = __shfl_down_sync(0xffffffffu, aw10, 1, LPP);
        const float nw3 = __shfl_down_sync(0xffffffffu, aw11, 1, LPP);
        const int p_off = __shfl_up_sync(0xffffffffu, off, 1, LPP);
        const int p_dx = __shfl_up_sync(0xffffffffu, dx, 1, LPP);
        const int p_dy = __shfl_up_sync(0xffffffffu, dy, 1, LPP);
        if (sub > 0 && s < LP) {   // rows also touched by sample s-1: that lane scatters them
          const int p0 = p_off, p1 = p_off + p_dx, p2 = p_off + p_dy, p3 = p_off + p_dx + p_dy;
          if (r0 == p0 || r0 == p1 || r0 == p2 || r0 == p3) rw0 = 0.f;
          if (r1 == p0 || r1 == p1 || r1 == p2 || r1 == p3) rw1 = 0.f;
          if (r2 == p0 || r2 == p1 || r2 == p2 || r2 == p3) rw2 = 0.f;
          if (r3 == p0 || r3 == p1 || r3 == p2 || r3 == p3) rw3 = 0.f;
        }
        if (sub + 1 < LPP && s + 1 < LP) {   // rows of sample s+1 that coincide with mine
          const int n0 = n_off, n1 = n_off + n_dx, n2 = n_off + n_dy, n3 = n_off + n_dx + n_dy;
#define PCT_ABSORB(nr, nw)              \
  if ((nr) == r0) rw0 += (nw);          \
  else if ((nr) == r1) rw1 += (nw);     \
  else if ((nr) == r2) rw2 += (nw);     \
  else if ((nr) == r3) rw3 += (nw);
          PCT_ABSORB(n0, nw0)
          PCT_ABSORB(n1, nw1)
          PCT_ABSORB(n2, nw2)
          PCT_ABSORB(n3, nw3)
#undef PCT_ABSORB
        }
      }
      float D00 = 0.f, D01 = 0.f, D10 = 0.f, D11 = 0.f;   // full dot products of MY sample
      // one sample of the round: lane j of every pair broadcasts its taps; all lanes gather the
      // four rows, scatter their grad_value share and reduce the four dot products
      auto tap = [&](const int j) {
        const int o = __shfl_sync(0xffffffffu, off, j, LPP);
        const int ox = __shfl_sync(0xffffffffu, dx, j, LPP);
        const int oy = __shfl_sync(0xffffffffu, dy, j, LPP);
        const float b00 = __shfl_sync(0xffffffffu, rw0, j, LPP);
        const float b01 = __shfl_sync(0xffffffffu, rw1, j, LPP);
        const float b10 = __shfl_sync(0xffffffffu, rw2, j, LPP);
        const float b11 = __shfl_sync(0xffffffffu, rw3, j, LPP);
        PCT_CHECK_TAP(o, ox, oy, (m * D + sub * VE), VE, (long long)S * MD);
        // 32-bit corner offsets first: each address is then ONE wide multiply-add off the base
        const int o01 = o + ox, o10 = o + oy, o11 = o10 + ox;
        float v00[VE], v01[VE], v10[VE], v11[VE];
        CV::load(vimg + o, v00);
        CV::load(vimg + o01, v01);
        CV::load(vimg + o10, v10);
        CV::load(vimg + o11, v11);
        // grad_value taps: w_k * attn * grad_out  (cuh:118,130,139,148,157)
        if (kDet) {
          unsigned long long* r = reinterpret_cast<unsigned long long*>(grow);
#pragma unroll
          for (int i = 0; i < VE; ++i) {
            if (b00 != 0.f) atomicAdd(r + o + i * LPP, (unsigned long long)__float2ll_rn(b00 * gs[i]));
            if (b01 != 0.f) atomicAdd(r + o01 + i * LPP, (unsigned long long)__float2ll_rn(b01 * gs[i]));
            if (b10 != 0.f) atomicAdd(r + o10 + i * LPP, (unsigned long long)__float2ll_rn(b10 * gs[i]));
            if (b11 != 0.f) atomicAdd(r + o11 + i * LPP, (unsigned long long)__float2ll_rn(b11 * gs[i]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < VE; i += 4) {
            if (b00 != 0.f) red_add_v4(gimg + o + i, b00 * gsc[i], b00 * gsc[i + 1], b00 * gsc[i + 2], b00 * gsc[i + 3]);
            if (b01 != 0.f) red_add_v4(gimg + o01 + i, b01 * gsc[i], b01 * gsc[i + 1], b01 * gsc[i + 2], b01 * gsc[i + 3]);
            if (b10 != 0.f) red_add_v4(gimg + o10 + i, b10 * gsc[i], b10 * gsc[i + 1], b10 * gsc[i + 2], b10 * gsc[i + 3]);
            if (b11 != 0.f) red_add_v4(gimg + o11 + i, b11 * gsc[i], b11 * gsc[i + 1], b11 * gsc[i + 2], b11 * gsc[i + 3]);
          }
        }
        float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
#pragma unroll
        for (int i = 0; i < VE; ++i) {
          d00 = fmaf(go[i], v00[i], d00);
          d01 = fmaf(go[i], v01[i], d01);
          d10 = fmaf(go[i], v10[i], d10);
          d11 = fmaf(go[i], v11[i], d11);
        }
#pragma unroll
        for (int w = LPP / 2; w > 0; w >>= 1) {
          d00 += __shfl_xor_sync(0xffffffffu, d00, w);
          d01 += __shfl_xor_sync(0xffffffffu, d01, w);
          d10 += __shfl_xor_sync(0xffffffffu, d10, w);
          d11 += __shfl_xor_sync(0xffffffffu, d11, w);
        }
        if (sub == j) { D00 = d00; D01 = d01; D10 = d10; D11 = d11; }
      };
      // blocks of kBlk samples behind one warp-uniform test (see the forward kernel)
      constexpr int kBlk = LPP < PCT_TAP_BLOCK_BWD ? LPP : PCT_TAP_BLOCK_BWD;
#pragma unroll
      for (int j0 = 0; j0 < LPP; j0 += kBlk) {
        if (s0 + j0 + kBlk <= LP) {
#pragma unroll
          for (int j = 0; j < kBlk; ++j) tap(j0 + j);
        } else {
#pragma unroll
          for (int j = 0; j < kBlk; ++j)
            if (s0 + j0 + j < LP) tap(j0 + j);
        }
      }
      if (s < LP && qvalid) {
        const float s_val = w00 * D00 + w01 * D01 + w10 * D10 + w11 * D11;
        const float s_gh = ch00 * D00 + ch01 * D01 + ch10 * D10 + ch11 * D11;
        const float s_gw = cw00 * D00 + cw01 * D01 + cw10 * D10 + cw11 * D11;
        const long long si = pair * LP + s;
        if (!FUSED) __stcs(grad_attn + si, s_val);                                   // cuh:161
        else if (!lg_regs) grad_attn[si] = s_val;
        float2 gl = make_float2(Wf * s_gw * a, Hf * s_gh * a);                      // cuh:162-163
        if (FUSED) {
          gl.x *= dloc.x;
          gl.y *= dloc.y;
          ag_dot = fmaf(a, s_val, ag_dot);
          if (s0 == 0) { pa0 = a; pg0 = s_val; }
          else if (s0 == LPP) { pa1 = a; pg1 = s_val; }
          else { pa2 = a; pg2 = s_val; }
        }
        __stcs(reinterpret_cast<float2*>(grad_loc) + si, gl);
      }
    }
    if (FUSED) {
      // softmax backward over the pair's L*P logits: a_s * (g_s - sum_t a_t g_t)
      const float dot = group_sum<LPP>(ag_dot);
      if (qvalid) {
        if (lg_regs) {
          const long long si = pair * LP + sub;
          if (sub < LP) grad_attn[si] = pa0 * (pg0 - dot);
          if (sub + LPP < LP) grad_attn[si + LPP] = pa1 * (pg1 - dot);
          if (sub + 2 * LPP < LP) grad_attn[si + 2 * LPP] = pa2 * (pg2 - dot);
        } else {   // each lane re-reads the g_s it wrote above
          for (int s = sub; s < LP; s += LPP) {
            const long long si = pair * LP + s;
            const float a = expf(__ldg(ap + s) - lmax) * inv_sum;
            grad_attn[si] = a * (grad_attn[si] - dot);
          }
        }
      }
    }
  }
  __syncthreads();
  iter.finish();
}

// Zero-fill of the grad_value accumulator.  A kernel, not cudaMemsetAsync: a memset may be placed on
// a copy engine, where it queues behind the host<->device copies of the end-to-end pipeline
// (pctrans_b200/host_pipeline.py) and stalls the backward of every chunk.
static __global__ void msda_zero_kernel(uint4* __restrict__ dst, long long n16, unsigned char* tail,
                                 int n_tail) {
  // programmatic dependent launch: the gradient kernel that follows on the stream may start its
  // CTAs now -- the cell-binned backward sorts its first chunk while this kernel still writes, and
  // waits (griddepcontrol.wait) before its first reduction into the cleared image
  asm volatile("griddepcontrol.launch_dependents;");
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {   // four independent 16-byte stores in flight
    dst[i] = z;
    dst[i + stride] = z;
    dst[i + 2 * stride] = z;
    dst[i + 3 * stride] = z;
  }
  for (; i < n16; i += stride) dst[i] = z;
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) tail[threadIdx.x] = 0;
}

// ---------------------------------------------------------------------------------------------
// Deterministic grad_value (PCT_MSDA_FLAG_DETERMINISTIC).  Floating-point
// atomics make the summation order -- and therefore the low bits of grad_value -- vary from run to
// run (the reference has the same property, cuh:130-157).  Here every contribution is converted to
// 64-bit fixed point with a power-of-two scale derived from max|grad_out| and accumulated with
// integer reductions, which are associative: the result is bit-identical across runs and launch
// shapes.  Costs 8-byte accumulators (twice the L2 reduction traffic) and one conversion pass.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void msda_absmax_kernel(const T* __restrict__ x, long long n, DetScratch* sc) {
  unsigned int m = 0u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = fabsf(Scalar<T>::template load<float>(x + i));
    if (v == v && v <= 3.4e38f) m = max(m, __float_as_uint(v));   // ignore NaN / Inf
  }
#pragma unroll
  for (int w = 16; w > 0; w >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, w));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(&sc->max_bits, m);
}

static __global__ void msda_det_scale_kernel(DetScratch* sc) {
  const float mx = __uint_as_float(sc->max_bits);
  // max*scale lands in [2^38, 2^39): bilinear weight * attention <= 1, so 2^24 contributions per
  // element still fit in 63 bits
  sc->scale = mx > 0.f ? ldexpf(1.f, 38 - ilogbf(mx)) : 1.f;
}

template <typename T>
__global__ void msda_fixed_to_value_kernel(const long long* __restrict__ acc, T* __restrict__ dst,
                                           long long n, const DetScratch* sc) {
  const float inv = 1.f / sc->scale;   // exact: power of two
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    Scalar<T>::template store<float>(dst + i, __ll2float_rn(acc[i]) * inv);
}

// ---------------------------------------------------------------------------------------------
// Scaled fp16 accumulator (bf16 values): bound, scale, and the conversion back.
// Any element of grad_value is  sum over (q, l, p) of  w * attn * grad_out  with bilinear weights
// w <= 1, so  |element| <= max|grad_out| * Lq * max_{b,q,m} sum_{l,p} |attn|.  The scale is the
// largest power of two that keeps that bound below the fp16 maximum.
// ---------------------------------------------------------------------------------------------
static __global__ void msda_attn_rowsum_max_kernel(const float* __restrict__ attn, long long n_pairs,
                                                   int LP, DetScratch* sc) {
  unsigned int m = 0u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < LP; ++k) s += fabsf(__ldg(attn + i * LP + k));
    if (s == s && s <= 3.4e38f) m = max(m, __float_as_uint(s));
  }
#pragma unroll
  for (int w = 16; w > 0; w >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, w));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(&sc->rowsum_bits, m);
}

static __global__ void msda_half_scale_kernel(DetScratch* sc, int Lq) {
  const float mx = __uint_as_float(sc->max_bits);
  const float rows = fmaxf(__uint_as_float(sc->rowsum_bits), 1e-30f);
  const float bound = mx * (float)Lq * rows;          // no accumulated element can exceed this
  // bound * scale < 2^15 (fp16 max is 65504): scale = 2^(14 - exponent(bound))
  sc->scale = (bound > 0.f && bound < 3.0e38f) ? ldexpf(1.f, 14 - ilogbf(bound)) : 1.f;
}

static __global__ void msda_half_to_bf16_kernel(const __half* __restrict__ acc,
                                                __nv_bfloat16* __restrict__ dst, long long n,
                                                const DetScratch* sc) {
  const float inv = 1.f / sc->scale;   // exact: power of two
  const long long n8 = n / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 r = __ldcs(reinterpret_cast<const uint4*>(acc) + i);
    const __half2* h = reinterpret_cast<const __half2*>(&r);
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(h[k]);
      v[2 * k] = f.x * inv;
      v[2 * k + 1] = f.y * inv;
    }
    ChanVec<__nv_bfloat16, 16>::store_stream(dst + 8 * i, v);
  }
  if (blockIdx.x == 0)
    for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x)
      dst[i] = __float2bfloat16_rn(__half2float(acc[i]) * inv);
}

// fp32 scratch image -> bf16 grad_value (bf16 path only)
static __global__ void msda_round_to_bf16_kernel(const float* __restrict__ src,
                                          __nv_bfloat16* __restrict__ dst, long long n8) {
  // n8 = number of 8-element groups
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    ChanVec<__nv_bfloat16, 16>::store_stream(dst + 8 * i, v);
  }
}

static __global__ void msda_round_to_bf16_tail_kernel(const float* __restrict__ src,
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
__device__ __forceinline__ float fma_coord<float>(float loc, int size) { return __fsub_rn(__fmul_rn(loc, (float)size), 0.5f); }
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
