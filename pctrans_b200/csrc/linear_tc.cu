// linear_tc.cu -- fp32 Linear on the tcgen05 tensor cores (3xTF32 split, or bf16 operands with
// PCT_LINEAR_BF16: struct Prec below), with the encoder layer's epilogues fused: bias, ReLU,
// residual add, ReLU gate, row mask, LayerNorm.  C ABI in include/pct_linear.h.  The description
// that follows is the 3xTF32 instantiation; the bf16 one differs in the K step (32 elements, one
// rounded piece per operand, two kind::f16 MMAs) and may read / write bf16 tensors.
//
// One CTA computes a 128-row x (<=256)-column tile of y = x @ W^T:
//   * the accumulator is a 128-lane x 256-column fp32 block of tensor memory (TMEM);
//   * the K dimension is streamed in 16-float steps through a two-stage shared-memory ring.  The
//     weight was split and re-tiled once (pct_linear_pack_weight) into exactly the bytes a stage
//     holds, so one thread fetches a step's W pieces with a single 32 KB bulk copy (TMA, no
//     tensor map) that bypasses the load/store pipe.  Four loader warps read the activation step
//     with 16-byte loads, split it into its tf32 head and the fp32 remainder on the fly, and store
//     the pieces as 8-row x 16-byte core matrices (the K-major, no-swizzle canonical layout of the
//     tensor-core shared-memory descriptor);
//   * one thread of a fifth warp issues six tcgen05.mma (2 k-steps x {hi*hi, hi*lo, lo*hi}) per
//     stage and commits them to the stage's mbarrier, which hands the stage back to the loaders;
//     the roles talk through mbarriers only, so loads, copies and MMAs of different steps overlap;
//   * the epilogue reads the accumulator with tcgen05.ld -- thread t of loader warp w owns row
//     32*w+t and walks all its columns -- so bias / ReLU / LayerNorm are per-thread arithmetic
//     (two-pass mean / variance, the pre-norm row parked in TMEM in between, no exchange).
//     Global traffic of the epilogue (residual in, y out) is transposed through a per-warp
//     shared-memory patch so that every warp access covers whole 128-byte lines: the kernel is
//     bound by load/store-pipe wavefronts, not by the tensor core (profiles/r01_linear_*).
// Two CTAs fit per SM (96 KB of shared memory, 256 TMEM columns each), so one CTA's loads and
// epilogue overlap the other's MMAs; in the first wave the second CTA of an SM starts half a phase
// late (PCT_LIN_STAGGER_*) so that the pair stays out of phase.
#include "../../include/pct_linear.h"
#include "../../include/pct_msda.h"

#include <cuda_runtime.h>
#ifndef PCT_PAIR_STAGES
#define PCT_PAIR_STAGES 3
#endif
#ifndef PCT_LIN_XPF
#define PCT_LIN_XPF 0   // L2 prefetch of the tile's x rows: 0 none, 1 every CTA, 2 first column tile only
#endif
#ifndef PCT_LIN_RPF
#define PCT_LIN_RPF 1   // L2 prefetch of the tile's residual / gate rows ...
#endif
#ifndef PCT_LIN_RPF_LEAD
#define PCT_LIN_RPF_LEAD 4   // ... this many K steps before the end of the main loop
#endif
#ifndef PCT_LIN_STAGGER_PER_STEP
#define PCT_LIN_STAGGER_PER_STEP 200 // ns per K step: start delay of the second CTA of every SM (first wave)
#endif
#ifndef PCT_LIN_STAGGER_FIXED
#define PCT_LIN_STAGGER_FIXED 2000   // ns
#endif
#ifndef PCT_LIN_L2PF
#define PCT_LIN_L2PF 256 // L2 prefetch hint of the loaders' x loads: 0 none, 128 / 256 bytes
#endif
#ifndef PCT_LIN_AHEAD_BF16
#define PCT_LIN_AHEAD_BF16 4
#endif
#ifndef PCT_LIN_AHEAD_TF32
#define PCT_LIN_AHEAD_TF32 4
#endif
#include <stdlib.h>

namespace pct {
void count_launch();   // msda_capi.cu
}

namespace {

constexpr int kBlockM = 128;   // rows per tile = TMEM lanes
constexpr int kBlockN = 256;   // max columns per tile = TMEM columns allocated
constexpr int kBlockK = 16;    // floats of K per stage: four 16-byte core-matrix columns
constexpr int kStages = 2;
constexpr int kLoaderWarps = 4;                 // load/split x, then run the epilogue
constexpr int kThreads = (kLoaderWarps + 2) * 32;   // + MMA issuer warp (TMEM owner) + W-copy warp
constexpr int kMmaK = 8;       // K of one tcgen05.mma.kind::tf32

constexpr int kCoreBytes = 128;                       // one 8 x 16 B core matrix
constexpr int kLbo = kCoreBytes;                      // K-adjacent core matrices
constexpr int kSbo = (kBlockK / 4) * kCoreBytes;      // next 8-row group
constexpr int kABytes = kBlockM * kBlockK * 4;        // one piece (hi or lo) of the x tile
constexpr int kBBytes = kBlockN * kBlockK * 4;        // one piece of the W tile
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
constexpr int kWStepBytes = 2 * kBBytes;              // packed W: hi piece then lo piece of one K step
constexpr int kBarrierBytes = 128;
constexpr int kVecBytes = 3 * kBlockN * 4;            // bias, gamma, beta of the tile's columns
constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + kVecBytes;
constexpr int kAhead = 3;                             // K steps of x in flight per loader thread
constexpr int kPatchStride = 36;                      // floats per patch row: conflict-free both ways
static_assert(kLoaderWarps * 32 * kPatchStride * 4 <= kStages * kStageBytes, "epilogue patches reuse the ring");

// The forward / input-gradient kernel exists in two operand precisions with the same byte geometry
// (a K step is 64 bytes of every row: four 16-byte core-matrix columns, two MMAs' worth of K):
//   3xTF32 (fp32 accuracy):  16 floats of K per step, every operand as a tf32 head and an fp32
//                            remainder piece, 6 MMAs per step, 48 KB per stage, 2 stages;
//   bf16 (PCT_LINEAR_BF16):  32 elements of K per step, operands rounded once to bf16, 2 MMAs per
//                            step (kind::f16, fp32 accumulate), 24 KB per stage, 4 stages.
template <bool BF16>
struct Prec {
  static constexpr int kPieces = BF16 ? 1 : 2;
  static constexpr int kStepK = BF16 ? 32 : 16;             // elements of K per stage
  static constexpr int kNumStages = BF16 ? 4 : 2;
  static constexpr int kStage = kPieces * (kABytes + kBBytes);
  static constexpr int kWStep = kPieces * kBBytes;
  static constexpr int kXAhead = BF16 ? PCT_LIN_AHEAD_BF16 : PCT_LIN_AHEAD_TF32;   // K steps of x in flight per loader thread
  static constexpr int kSmem = kNumStages * kStage + kBarrierBytes + kVecBytes;
  static_assert(kLoaderWarps * 32 * kPatchStride * 4 <= kNumStages * kStage, "epilogue patches reuse the ring");
  static_assert((3 * kNumStages + 1) * 8 + 4 <= kBarrierBytes, "barriers + TMEM slot");
};
static_assert(Prec<false>::kSmem == kSmemBytes && Prec<false>::kStage == kStageBytes, "3xTF32 geometry");

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// Shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kLbo >> 4) << 16) |
         ((uint64_t)(kSbo >> 4) << 32) | ((uint64_t)1 << 46);
}

// Instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t make_instr_desc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kBlockM >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n (kind::f16).
__device__ __forceinline__ uint32_t make_instr_desc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kBlockM >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// eight fp32 -> eight bf16 (round to nearest even), element 0 at the lowest address
__device__ __forceinline__ uint4 pack_bf16x8(const float4& a, const float4& b) {
  uint4 r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.x) : "f"(a.y), "f"(a.x));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.y) : "f"(a.w), "f"(a.z));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.z) : "f"(b.y), "f"(b.x));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.w) : "f"(b.w), "f"(b.z));
  return r;
}

// four bf16 (8 bytes, element 0 lowest) -> fp32; exact
__device__ __forceinline__ float4 bf16x4_to_float4(uint2 u) {
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}

// four fp32 -> four bf16 (round to nearest even), element 0 lowest
__device__ __forceinline__ uint2 pack_bf16x4(const float4& a) {
  uint2 r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.x) : "f"(a.y), "f"(a.x));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.y) : "f"(a.w), "f"(a.z));
  return r;
}

// 16-byte read-only load of an activation chunk (optionally with an L2 prefetch-size hint)
__device__ __forceinline__ float4 ldg_x(const void* p) {
#if PCT_LIN_L2PF == 256
  float4 v;
  asm volatile("ld.global.nc.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#elif PCT_LIN_L2PF == 128
  float4 v;
  asm volatile("ld.global.nc.L2::128B.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  return __ldg(reinterpret_cast<const float4*>(p));
#endif
}

__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// Bounded wait: a pipeline bug traps (launch error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}

// 1-D bulk copy global -> shared (TMA without a tensor map); completion counted on `bar`.
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

__device__ __forceinline__ float tf32_head(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// TMEM -> registers, 16 columns of this thread's lane.  Asynchronous: the registers are valid only
// after tmem_ld_wait() and tmem_tie16() (which pins later uses of r[] behind the wait).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_tie16(uint32_t* r) {
  asm volatile(""
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                 "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]),
                 "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
      "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
}

struct LinearArgs {
  const float* x;
  const unsigned char* w_packed;
  const float* bias;
  const float* residual;
  const float* gamma;
  const float* beta;
  float* y;
  float* pre_norm;   // LayerNorm only, may be null: the normalised tensor's input (rows, N)
  float* mean;       // LayerNorm only, may be null: row statistics for the backward pass
  float* rstd;
  float eps;
  int rows, K, N;
  unsigned flags;
  // SPLIT instantiations (pct_linear_forward_split / _cat): x is the K-concatenation of two dense
  // tensors [x (rows, k_split) | x2 (rows, K - k_split)], y the N-concatenation [y (rows, n_split) |
  // y2 (rows, N - n_split)]; 0 = that side is a single tensor.  k_split is a multiple of the K step,
  // n_split of 32 (the epilogue's column block).
  const float* x2;
  float* y2;
  int k_split, n_split;
};

// Epilogue of one 128 x n_tile accumulator, run by the four loader warps (warp w owns rows
// 32w .. 32w+31 of the tile).  Row domain: a thread owns one row and walks its columns in TMEM
// order.  Line domain: a warp instruction covers 4 rows x 128 contiguous bytes of global memory;
// the two are bridged by a per-warp 32 x 36-float patch at the start of shared memory.
template <bool IO16 = false, bool SPLIT = false>   // bf16-in-memory options / split tensors: own instantiations
__device__ __forceinline__ void linear_epilogue(const LinearArgs& a, unsigned char* smem, uint32_t tmem,
                                                const float* vec_bias, const float* vec_gamma,
                                                const float* vec_beta, int warp, int lane, int tid,
                                                int m0, int n0, int n_tile) {
  const int patch_id = warp, c_lo = 0, c_hi = n_tile;
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  const bool relu = a.flags & PCT_LINEAR_RELU, has_res = a.flags & PCT_LINEAR_RESIDUAL;
  const bool has_gate = a.flags & PCT_LINEAR_GATE;   // `residual` holds the gate tensor instead
  const bool res_bf16 = IO16 && (a.flags & PCT_LINEAR_RES_BF16), y_bf16 = IO16 && (a.flags & PCT_LINEAR_Y_BF16);
  // PCT_LINEAR_ROWMASK: `residual` holds one byte per row; rows with a non-zero byte are written
  // as zeros (the masked_fill after value_proj, mod.py:100-101) -- row domain: a thread owns a row
  const int my_row = m0 + warp * 32 + lane;
  const bool row_masked = (a.flags & PCT_LINEAR_ROWMASK) && my_row < a.rows &&
                          reinterpret_cast<const unsigned char*>(a.residual)[my_row] != 0;
  float* patch = reinterpret_cast<float*>(smem) + patch_id * (32 * kPatchStride);
  float* patch_row = patch + lane * kPatchStride;
  const int l_row = lane >> 3, l_c4 = (lane & 7) * 4;
  const size_t warp_off = (size_t)(m0 + warp * 32) * a.N + n0;   // first row of this warp
  const int rows_left = a.rows - (m0 + warp * 32);
  float v[32];

  // residual block (32 rows x cw columns at column c): line-domain loads into registers (issued
  // one block ahead of their use), then through the patch into the row domain
  float4 res[8];
  auto fetch_residual = [&](int c, int cw) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = 4 * j + l_row;
      if (l_c4 < cw && r < rows_left) {
        const size_t at = warp_off + (size_t)r * a.N + c + l_c4;
        if (res_bf16) {   // raw bits now, widened where they are used: a conversion here would
                          // wait for every load in turn
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(
              reinterpret_cast<const unsigned short*>(a.residual) + at));
          res[j].x = __uint_as_float(u.x), res[j].y = __uint_as_float(u.y);
        } else {
          res[j] = __ldg(reinterpret_cast<const float4*>(a.residual + at));
        }
      }
    }
  };
  auto res_value = [&](int j) {
    return res_bf16 ? bf16x4_to_float4(make_uint2(__float_as_uint(res[j].x), __float_as_uint(res[j].y)))
                    : res[j];
  };
  auto add_residual = [&](int cw) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(patch + (4 * j + l_row) * kPatchStride + l_c4) = res_value(j);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (4 * q < cw) {
        const float4 t = *reinterpret_cast<const float4*>(patch_row + 4 * q);
        v[4 * q] += t.x, v[4 * q + 1] += t.y, v[4 * q + 2] += t.z, v[4 * q + 3] += t.w;
      }
    __syncwarp();
  };
  // v (row domain) -> y block, through the patch
  auto store_block = [&](float* dst_base, int c, int cw, bool gated = false, bool as_bf16 = false,
                         const uint2* gate_raw = nullptr) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (4 * q < cw)
        *reinterpret_cast<float4*>(patch_row + 4 * q) =
            make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    __syncwarp();
    float4 t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      t[j] = *reinterpret_cast<const float4*>(patch + (4 * j + l_row) * kPatchStride + l_c4);
    __syncwarp();
    if (gated) {   // res[] holds the gate block in exactly this (line-domain) arrangement
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 gt = gate_raw ? bf16x4_to_float4(gate_raw[j]) : res_value(j);
        t[j].x = gt.x > 0.f ? t[j].x : 0.f, t[j].y = gt.y > 0.f ? t[j].y : 0.f;
        t[j].z = gt.z > 0.f ? t[j].z : 0.f, t[j].w = gt.w > 0.f ? t[j].w : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = 4 * j + l_row;
      if (l_c4 < cw && r < rows_left) {
        size_t at = warp_off + (size_t)r * a.N + c + l_c4;
        float* dst = dst_base;
        if (SPLIT && a.n_split && dst_base == a.y) {   // block-uniform: n_split is a multiple of 32
          const int gc = n0 + c;                       // first column of the block in [y | y2]
          const bool second = gc >= a.n_split;
          dst = second ? a.y2 : a.y;
          at = (size_t)(m0 + warp * 32 + r) * (second ? a.N - a.n_split : a.n_split) +
               (second ? gc - a.n_split : gc) + l_c4;
        }
        if (as_bf16)
          *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(dst) + at) = pack_bf16x4(t[j]);
        else
          *reinterpret_cast<float4*>(dst + at) = t[j];
      }
    }
  };
  auto load_acc = [&](int c, int cw) {
    uint32_t u[32];
    tmem_ld16(t_row + c, u);
    if (cw > 16) tmem_ld16(t_row + c + 16, u + 16);
    tmem_ld_wait();
    tmem_tie16(u);
    tmem_tie16(u + 16);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(u[i]);
  };
  auto add_bias = [&](int c, int cw) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (4 * q < cw) {
        const float4 b = *reinterpret_cast<const float4*>(vec_bias + c + 4 * q);
        v[4 * q] += b.x, v[4 * q + 1] += b.y, v[4 * q + 2] += b.z, v[4 * q + 3] += b.w;
      }
  };

  if (has_gate && res_bf16) {
    // bf16 gate (the FFN's hidden activation, dgrad of linear2): raw gate blocks are fetched TWO
    // blocks ahead into two small register buffers (8 x 8 bytes each) -- one block ahead leaves
    // most of the load latency exposed, and this epilogue reads as many bytes as it writes
    uint2 g0[8], g1[8];
    auto fetch_gate = [&](uint2* dst, int c) {
      const int cw = min(32, c_hi - c);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = 4 * j + l_row;
        if (l_c4 < cw && r < rows_left)
          dst[j] = __ldg(reinterpret_cast<const uint2*>(
              reinterpret_cast<const unsigned short*>(a.residual) + warp_off + (size_t)r * a.N + c + l_c4));
      }
    };
    auto block = [&](int c, const uint2* gate) {
      const int cw = min(32, c_hi - c);
      load_acc(c, cw);
      add_bias(c, cw);
      if (relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      store_block(a.y, c, cw, true, y_bf16, gate);
    };
    if (c_lo < c_hi) fetch_gate(g0, c_lo);
    if (c_lo + 32 < c_hi) fetch_gate(g1, c_lo + 32);
    for (int c = c_lo; c < c_hi; c += 64) {
      block(c, g0);
      if (c + 64 < c_hi) fetch_gate(g0, c + 64);
      if (c + 32 < c_hi) {
        block(c + 32, g1);
        if (c + 96 < c_hi) fetch_gate(g1, c + 96);
      }
    }
  } else if (!(a.flags & PCT_LINEAR_LAYERNORM)) {
    if ((has_res || has_gate) && c_lo < c_hi) fetch_residual(c_lo, min(32, c_hi - c_lo));
    for (int c = c_lo; c < c_hi; c += 32) {
      const int cw = min(32, c_hi - c);
      load_acc(c, cw);
      add_bias(c, cw);
      if (has_res) {
        add_residual(cw);
        if (c + 32 < c_hi) fetch_residual(c + 32, min(32, c_hi - c - 32));
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      if (row_masked) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
      store_block(a.y, c, cw, has_gate, y_bf16);
      if (has_gate && c + 32 < c_hi) fetch_residual(c + 32, min(32, c_hi - c - 32));
    }
  } else {
    // pass 1: pre-norm row = acc + bias (+ residual), parked back in TMEM; row sum
    float sum = 0.f;
    if (has_res && c_lo < c_hi) fetch_residual(c_lo, min(32, c_hi - c_lo));
    for (int c = c_lo; c < c_hi; c += 32) {
      const int cw = min(32, c_hi - c);
      load_acc(c, cw);
      add_bias(c, cw);
      if (has_res) {
        add_residual(cw);
        if (c + 32 < c_hi) fetch_residual(c + 32, min(32, c_hi - c - 32));
      }
      float part[4] = {0.f, 0.f, 0.f, 0.f};   // four independent chains instead of one of 32 adds
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < cw) part[i & 3] += v[i];
      sum += (part[0] + part[1]) + (part[2] + part[3]);
      tmem_st16(t_row + c, v);
      if (cw > 16) tmem_st16(t_row + c + 16, v + 16);
      if (a.pre_norm) store_block(a.pre_norm, c, cw);   // training: LayerNorm's input is saved
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const float mean = sum / (float)a.N;
    // pass 2: centred sum of squares
    float ssq = 0.f;
    for (int c = c_lo; c < c_hi; c += 32) {
      const int cw = min(32, c_hi - c);
      load_acc(c, cw);
      float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < cw) part[i & 3] += (v[i] - mean) * (v[i] - mean);
      ssq += (part[0] + part[1]) + (part[2] + part[3]);
    }
    const float rstd = rsqrtf(ssq / (float)a.N + a.eps);
    if (a.mean && m0 + tid < a.rows) a.mean[m0 + tid] = mean, a.rstd[m0 + tid] = rstd;
    // pass 3: normalise, scale, shift, store
    for (int c = c_lo; c < c_hi; c += 32) {
      const int cw = min(32, c_hi - c);
      load_acc(c, cw);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (4 * q < cw) {
          const float4 g = *reinterpret_cast<const float4*>(vec_gamma + c + 4 * q);
          const float4 b = *reinterpret_cast<const float4*>(vec_beta + c + 4 * q);
          v[4 * q] = (v[4 * q] - mean) * rstd * g.x + b.x;
          v[4 * q + 1] = (v[4 * q + 1] - mean) * rstd * g.y + b.y;
          v[4 * q + 2] = (v[4 * q + 2] - mean) * rstd * g.z + b.z;
          v[4 * q + 3] = (v[4 * q + 3] - mean) * rstd * g.w + b.w;
        }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      store_block(a.y, c, cw, false, y_bf16);
    }
  }
}

// Phase timestamps of one CTA per 64 (variant builds with -DPCT_LINEAR_TRACE only; read back with
// pct_linear_debug_trace): 0 start, 1 set-up done, 2 first K step stored, 3 last K step stored,
// 4 accumulator complete, 5 epilogue done -- taken by thread 0 (a loader / epilogue thread).
#ifdef PCT_LINEAR_TRACE
__device__ long long g_linear_trace[8 * 4096];
#define PCT_TRACE(slot)                                                                         \
  do {                                                                                          \
    const unsigned cta_ = blockIdx.y * gridDim.x + blockIdx.x;                                  \
    if (threadIdx.x == 0 && (cta_ & 63u) == 0 && (cta_ >> 6) < 4096u)                           \
      g_linear_trace[(cta_ >> 6) * 8 + (slot)] = clock64();                                     \
  } while (0)
#else
#define PCT_TRACE(slot) do { } while (0)
#endif

template <bool BF16, bool IO16 = false, bool SPLIT = false>
__device__ __forceinline__ void linear_body(const LinearArgs& a, unsigned char* smem) {
  static_assert(BF16 || !IO16, "bf16 tensors only with the bf16 kernel");
  static_assert(!(IO16 && SPLIT), "split tensors are fp32");
  using P = Prec<BF16>;
  constexpr int kStages = P::kNumStages, kStageBytes = P::kStage, kWStepBytes = P::kWStep;
  constexpr int kAhead = P::kXAhead;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n0 = blockIdx.x * kBlockN;
  const int m0 = blockIdx.y * kBlockM;
  if (PCT_LIN_STAGGER_PER_STEP || PCT_LIN_STAGGER_FIXED) {
    // Two CTAs share an SM and a tile is a load phase followed by an epilogue phase: started
    // together they stay in phase (both loading, then both storing).  The CTAs that fill the second
    // slot of every SM in the first wave start about half a tile late, so that one CTA's main loop
    // runs against the other's epilogue from then on (grids of at least two full waves only;
    // measured: 1024->256 272 -> 250 us 3xTF32, 171 -> 160 us bf16, LayerNorm 120 -> 113 us, others
    // within 2 %, profiles/r02_linear_stagger.log).
    unsigned nsm;
    asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (cta >= nsm && cta < 2 * nsm && gridDim.x * gridDim.y >= 4 * nsm) {
      long long ns = PCT_LIN_STAGGER_FIXED + (long long)(a.K / Prec<BF16>::kStepK) * PCT_LIN_STAGGER_PER_STEP;
      for (; ns > 0; ns -= 100000) __nanosleep((unsigned)(ns < 100000 ? ns : 100000));
    }
  }
  PCT_TRACE(0);
  const int n_tile = min(kBlockN, a.N - n0);

  // barriers, kStages each: stage free (its MMAs are done) | x pieces stored | W pieces landed;
  // then one for "accumulator complete"
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);
  float* vec_bias = reinterpret_cast<float*>(smem + kStages * kStageBytes + kBarrierBytes);
  float* vec_gamma = vec_bias + kBlockN;
  float* vec_beta = vec_gamma + kBlockN;
  const uint32_t bar_empty = smem_u32(bars);
  const uint32_t bar_xfull = bar_empty + 8 * kStages;
  const uint32_t bar_wfull = bar_empty + 16 * kStages;
  const uint32_t bar_done = bar_empty + 24 * kStages;

  if (warp == kLoaderWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)kBlockN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_xfull + 8 * s, kLoaderWarps);
      mbar_init(bar_wfull + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = tid; c < n_tile; c += kThreads) {   // L1 is streamed through by x: keep these close
    vec_bias[c] = a.bias ? a.bias[n0 + c] : 0.f;
    if (a.flags & PCT_LINEAR_LAYERNORM) vec_gamma[c] = a.gamma[n0 + c], vec_beta[c] = a.beta[n0 + c];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const int n_steps = a.K / P::kStepK;
  const uint32_t smem0 = smem_u32(smem);
  PCT_TRACE(1);

  if (warp == kLoaderWarps) {
    // ---- issuer: one thread feeds the tensor core ----------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = BF16 ? make_instr_desc_bf16(n_tile) : make_instr_desc(n_tile);
      for (int step = 0; step < n_steps; ++step) {
        const int s = step % kStages;
        const uint32_t parity = (step / kStages) & 1;
        mbar_wait(bar_wfull + 8 * s, parity);
        mbar_wait(bar_xfull + 8 * s, parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem0 + s * kStageBytes;
#pragma unroll
        for (int ks = 0; ks < kBlockK / kMmaK; ++ks) {
          const uint32_t koff = ks * (kMmaK / 4) * kLbo;   // 32 bytes of K per MMA in both precisions
          if constexpr (BF16) {
            mma_bf16(tmem, make_smem_desc(sa + koff), make_smem_desc(sa + kABytes + koff), idesc,
                     (step | ks) != 0);
          } else {
            const uint64_t d_ahi = make_smem_desc(sa + koff);
            const uint64_t d_alo = make_smem_desc(sa + kABytes + koff);
            const uint64_t d_bhi = make_smem_desc(sa + 2 * kABytes + koff);
            const uint64_t d_blo = make_smem_desc(sa + 2 * kABytes + kBBytes + koff);
            // consecutive MMAs share one operand (W_hi, then x_hi)
            mma_tf32(tmem, d_alo, d_bhi, idesc, (step | ks) != 0);
            mma_tf32(tmem, d_ahi, d_bhi, idesc, 1u);
            mma_tf32(tmem, d_ahi, d_blo, idesc, 1u);
          }
        }
        mma_commit(bar_empty + 8 * s);             // stage free once these MMAs have read it
        if (step == n_steps - 1) mma_commit(bar_done);
      }
    }
  } else if (warp == kLoaderWarps + 1) {
    // ---- W producer: one thread streams the packed weight steps with bulk copies --------------
    const unsigned char* w_tile = a.w_packed + (size_t)blockIdx.x * n_steps * kWStepBytes;
    auto fetch_w = [&](int step) {
      const int s = step % kStages;
      mbar_expect_tx(bar_wfull + 8 * s, kWStepBytes);
      bulk_load(smem0 + s * kStageBytes + P::kPieces * kABytes, w_tile + (size_t)step * kWStepBytes,
                kWStepBytes, bar_wfull + 8 * s);
    };
    if (lane == 0)
      for (int step = 0; step < kStages && step < n_steps; ++step) fetch_w(step);
    // Pull the tile's residual / gate rows towards L2 shortly before the epilogue reads them (a few
    // K steps before the end of the main loop: issued at the start they compete with the loaders'
    // first loads, measured).  The rows of x are NOT prefetched: with the loaders' own look-ahead a
    // whole-tile prefetch only queues in front of the loads that are needed first
    // (profiles/r02_linear_trace_prefetch.log).
    const int tile_rows = min(kBlockM, a.rows - m0);
    auto prefetch_residual = [&]() {
      const int esz = (a.flags & PCT_LINEAR_RES_BF16) ? 2 : 4;
      if (PCT_LIN_RPF && (a.flags & (PCT_LINEAR_RESIDUAL | PCT_LINEAR_GATE)))
        for (int r = lane; r < tile_rows; r += 32)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(
                           reinterpret_cast<const char*>(a.residual) +
                           ((size_t)(m0 + r) * a.N + n0) * esz),
                       "r"((uint32_t)(n_tile * esz))
                       : "memory");
    };
    if (PCT_LIN_XPF == 1 || (PCT_LIN_XPF == 2 && blockIdx.x == 0)) {
      const size_t x_bytes = (size_t)max(tile_rows, 0) * a.K * 4;
      const char* xp = reinterpret_cast<const char*>(a.x + (size_t)m0 * a.K);
      for (size_t off = (size_t)lane * 32768; off < x_bytes; off += 32 * 32768) {
        const size_t nb = x_bytes - off < 32768 ? x_bytes - off : 32768;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xp + off), "r"((uint32_t)nb)
                     : "memory");
      }
    }
    const int pf_step = max(kStages, n_steps - PCT_LIN_RPF_LEAD);
    if (pf_step >= n_steps) prefetch_residual();
    for (int step = kStages; step < n_steps; ++step) {
      if (step == pf_step) prefetch_residual();
      if (lane == 0) {
        mbar_wait(bar_empty + 8 * (step % kStages), ((step / kStages) - 1) & 1);
        fetch_w(step);
      }
      __syncwarp();
    }
  } else {
    // ---- loaders: read a K step of x, split it, store the pieces in core-matrix order ---------
    // float4 i of a piece lives at byte 16*i with i = 32*(row/8) + 8*(k/4) + row%8: eight
    // consecutive threads fill one core matrix and a warp reads 8 rows x 64 contiguous bytes.
    const int r8 = tid & 7, kc = (tid >> 3) & 3, rg = tid >> 5;
    constexpr int kPerThread = kBlockM * kBlockK / 4 / (kLoaderWarps * 32);
    // PCT_LINEAR_X_BF16 (bf16 kernel only): x is already bf16 in global memory -- the thread's
    // 16-byte chunk of a stage is copied as it is (half the bytes, no conversion)
    const bool xb = IO16 && (a.flags & PCT_LINEAR_X_BF16);
    const int x_esz = xb ? 2 : 4;
    const char* xrow[kPerThread];
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int row = m0 + (rg + kLoaderWarps * j) * 8 + r8;
      xrow[j] = row < a.rows ? reinterpret_cast<const char*>(a.x) +
                                   ((size_t)row * a.K + kc * (P::kStepK / 4)) * x_esz
                             : nullptr;
    }
    // kAhead K steps in flight: the loads are DRAM-latency bound.  A thread's share of a row is one
    // 16-byte chunk of the stage: 4 floats (3xTF32) or 8 floats rounded to bf16
    constexpr int kQ = BF16 ? 2 : 1;
    float4 xq[kAhead][kPerThread * kQ];
    // SPLIT: the K steps below k_split come from x (row pitch k_split), the rest from x2
    const int ks_steps = (SPLIT && a.k_split) ? a.k_split / P::kStepK : 0;
    const char* xrow2[SPLIT ? kPerThread : 1];
    if constexpr (SPLIT) {
      if (ks_steps) {
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
          const int row = m0 + (rg + kLoaderWarps * j) * 8 + r8;
          if (row < a.rows) {
            xrow[j] = reinterpret_cast<const char*>(a.x) + ((size_t)row * a.k_split + kc * (P::kStepK / 4)) * 4;
            xrow2[j] = reinterpret_cast<const char*>(a.x2) +
                       ((size_t)row * (a.K - a.k_split) + kc * (P::kStepK / 4)) * 4;
          }
        }
      }
    }
    auto fetch_x = [&](float4* dst, int step) {
      size_t at = (size_t)step * P::kStepK * x_esz;
      if constexpr (SPLIT) {
        if (ks_steps && step >= ks_steps) {
          at = (size_t)(step - ks_steps) * P::kStepK * 4;
#pragma unroll
          for (int j = 0; j < kPerThread; ++j)
#pragma unroll
            for (int h = 0; h < kQ; ++h)
              dst[j * kQ + h] = xrow[j] ? __ldg(reinterpret_cast<const float4*>(xrow2[j] + at) + h)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
          return;
        }
      }
#pragma unroll
      for (int j = 0; j < kPerThread; ++j) {
        if (xb) {
          dst[j * kQ] = xrow[j] ? ldg_x(xrow[j] + at)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
          for (int h = 0; h < kQ; ++h)
            dst[j * kQ + h] = xrow[j] ? ldg_x(xrow[j] + at + 16 * h)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
#pragma unroll
    for (int d = 0; d < kAhead; ++d)
      if (d < n_steps) fetch_x(xq[d], d);
    for (int step0 = 0; step0 < n_steps; step0 += kAhead) {
#pragma unroll
      for (int d = 0; d < kAhead; ++d) {
        const int step = step0 + d;
        if (step < n_steps) {
          const int s = step % kStages;
          if (step >= kStages) mbar_wait(bar_empty + 8 * s, ((step / kStages) - 1) & 1);
          float4* a_hi = reinterpret_cast<float4*>(smem + s * kStageBytes);
          [[maybe_unused]] float4* a_lo = reinterpret_cast<float4*>(smem + s * kStageBytes + kABytes);
#pragma unroll
          for (int j = 0; j < kPerThread; ++j) {
            if constexpr (BF16) {
              if (xb)
                a_hi[tid + kLoaderWarps * 32 * j] = xq[d][2 * j];
              else
                reinterpret_cast<uint4*>(a_hi)[tid + kLoaderWarps * 32 * j] =
                    pack_bf16x8(xq[d][2 * j], xq[d][2 * j + 1]);
            } else {
              const float4 x4 = xq[d][j];
              float4 h, l;
              h.x = tf32_head(x4.x), h.y = tf32_head(x4.y), h.z = tf32_head(x4.z), h.w = tf32_head(x4.w);
              l.x = x4.x - h.x, l.y = x4.y - h.y, l.z = x4.z - h.z, l.w = x4.w - h.w;
              a_hi[tid + kLoaderWarps * 32 * j] = h;
              a_lo[tid + kLoaderWarps * 32 * j] = l;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic -> async proxy
          __syncwarp();                                                    // one arrival per warp
          if (lane == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_xfull + 8 * s) : "memory");
          if (step == 0) PCT_TRACE(2);
          if (step == n_steps - 1) PCT_TRACE(3);
          if (step + kAhead < n_steps) fetch_x(xq[d], step + kAhead);
        }
      }
    }

    // ---- epilogue (the loader warps; warp w owns rows 32w .. 32w+31 of the tile) --------------
    // Row domain: this thread owns one row and walks its columns in TMEM order.
    // Line domain: a warp instruction covers 4 rows x 128 contiguous bytes of global memory.
    mbar_wait(bar_done, 0);
    PCT_TRACE(4);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    linear_epilogue<IO16, SPLIT>(a, smem, tmem, vec_bias, vec_gamma, vec_beta, warp, lane, tid, m0, n0, n_tile);
    PCT_TRACE(5);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kLoaderWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"((uint32_t)kBlockN)
                 : "memory");
  }
}

__global__ void __launch_bounds__(kThreads, 2) linear_tf32x3_kernel(const LinearArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  linear_body<false>(a, smem);
}

// bf16 operands (PCT_LINEAR_BF16): same roles, barriers and epilogue; fp32 tensors in global memory
__global__ void __launch_bounds__(kThreads, 2) linear_bf16_kernel(const LinearArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  linear_body<true>(a, smem);
}

// split tensors (pct_linear_forward_split / _cat): x = [x | x2] along K and / or y = [y | y2] along N
__global__ void __launch_bounds__(kThreads, 2) linear_tf32x3_split_kernel(const LinearArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  linear_body<false, false, true>(a, smem);
}

__global__ void __launch_bounds__(kThreads, 2) linear_bf16_split_kernel(const LinearArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  linear_body<true, false, true>(a, smem);
}

// ... and with some of x / y / residual bf16 in global memory (PCT_LINEAR_X_BF16 / Y_BF16 / RES_BF16)
__global__ void __launch_bounds__(kThreads, 2) linear_bf16_io16_kernel(const LinearArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  linear_body<true, true>(a, smem);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster, on the two SMs of a TPC, compute 256 rows
// x n_tile columns with one M=256 MMA stream issued by the leader.  Each CTA keeps its own 128
// rows of x and only HALF of the W step (n_tile/2 columns) in shared memory; the tensor cores of
// both SMs read the other half from the partner.  Per K step a CTA therefore moves 80 KB through
// its shared memory instead of 120 KB (the single-CTA kernel loads the shared-memory data pipe to
// ~100 %) and the ring holds three stages in the same 96 KB.
// Status (round 1): bit-for-bit the same results as the single-CTA kernel and 33 % fewer
// tensor-core shared-memory wavefronts (ncu), but not yet faster -- the cross-CTA hand-offs (remote
// arrivals, multicast commits) leave the pair latency-bound -- so it is opt-in:
// PCT_LINEAR_PAIR=1 in the environment selects it (tests/test_gpu_linear.py runs it that way).
constexpr int kPairStages = PCT_PAIR_STAGES;
constexpr int kPairStageBytes = 2 * kABytes + kBBytes;          // x hi/lo + half of W hi/lo
constexpr int kPairBarrierBytes = 128;
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + kPairBarrierBytes + kVecBytes;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}

__device__ __forceinline__ void mma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// completion of the pair's MMAs, signalled on the same barrier offset in both CTAs
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 2) linear_tf32x3_pair_kernel(const LinearArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues the MMAs)
  // the pair sits along x (cluster 2x1x1: cta_group::2 pairs the CTAs of ranks 2i, 2i+1)
  const int n0 = blockIdx.y * kBlockN;
  const int m0 = blockIdx.x * kBlockM;                     // may lie past the last row (padding CTA)
  const int n_tile = min(kBlockN, a.N - n0);
  const int half_bytes = n_tile * 32;                      // this CTA's n_tile/2 columns of a W piece

  // barriers, kPairStages each: stage free | stage ready (leader's copy counts: both CTAs' x
  // pieces and the partner's W) | this CTA's W landed; then "accumulator complete"
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPairStages * kPairStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kPairStages + 1);
  float* vec_bias = reinterpret_cast<float*>(smem + kPairStages * kPairStageBytes + kPairBarrierBytes);
  float* vec_gamma = vec_bias + kBlockN;
  float* vec_beta = vec_gamma + kBlockN;
  const uint32_t bar_empty = smem_u32(bars);
  const uint32_t bar_ready = bar_empty + 8 * kPairStages;
  const uint32_t bar_wloc = bar_empty + 16 * kPairStages;
  const uint32_t bar_done = bar_empty + 24 * kPairStages;
  const uint32_t leader_ready = map_to_cta(bar_ready, 0);

  if (warp == kLoaderWarps) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)kBlockN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_ready + 8 * s, 2 * kLoaderWarps + 1);
      mbar_init(bar_wloc + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = tid; c < n_tile; c += kThreads) {
    vec_bias[c] = a.bias ? a.bias[n0 + c] : 0.f;
    if (a.flags & PCT_LINEAR_LAYERNORM) vec_gamma[c] = a.gamma[n0 + c], vec_beta[c] = a.beta[n0 + c];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                                      // both CTAs' barriers exist from here on
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const int n_steps = a.K / kBlockK;
  const uint32_t smem0 = smem_u32(smem);

  if (warp == kLoaderWarps) {
    if (lane == 0 && rank == 0) {
      // ---- leader: issue the M=256 MMAs of the pair ------------------------------------------
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_tile >> 3) << 17) |
                             ((uint32_t)(2 * kBlockM >> 4) << 24);
      for (int step = 0; step < n_steps; ++step) {
        const int s = step % kPairStages;
        const uint32_t parity = (step / kPairStages) & 1;
        mbar_wait(bar_wloc + 8 * s, parity);               // own half of W
        mbar_wait_cluster(bar_ready + 8 * s, parity);      // both CTAs' x pieces, partner's W
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem0 + s * kPairStageBytes;
#pragma unroll
        for (int ks = 0; ks < kBlockK / kMmaK; ++ks) {
          const uint32_t koff = ks * (kMmaK / 4) * kLbo;
          const uint64_t d_ahi = make_smem_desc(sa + koff);
          const uint64_t d_alo = make_smem_desc(sa + kABytes + koff);
          const uint64_t d_bhi = make_smem_desc(sa + 2 * kABytes + koff);
          const uint64_t d_blo = make_smem_desc(sa + 2 * kABytes + kBBytes / 2 + koff);
          mma_tf32_pair(tmem, d_alo, d_bhi, idesc, (step | ks) != 0);
          mma_tf32_pair(tmem, d_ahi, d_bhi, idesc, 1u);
          mma_tf32_pair(tmem, d_ahi, d_blo, idesc, 1u);
        }
        mma_commit_pair(bar_empty + 8 * s);
        if (step == n_steps - 1) mma_commit_pair(bar_done);
      }
    } else if (lane == 0) {
      // ---- partner: tell the leader when this CTA's half of W has landed ----------------------
      for (int step = 0; step < n_steps; ++step) {
        const int s = step % kPairStages;
        mbar_wait(bar_wloc + 8 * s, (step / kPairStages) & 1);
        mbar_arrive_remote(leader_ready + 8 * s);
      }
    }
  } else if (warp == kLoaderWarps + 1) {
    // ---- W producer (both CTAs): this CTA's columns of the hi piece, then of the lo piece ------
    const unsigned char* w_tile =
        a.w_packed + (size_t)blockIdx.y * n_steps * kWStepBytes + (size_t)rank * half_bytes;
    auto fetch_w = [&](int step) {
      const int s = step % kPairStages;
      const uint32_t dst = smem0 + s * kPairStageBytes + 2 * kABytes;
      mbar_expect_tx(bar_wloc + 8 * s, 2 * half_bytes);
      bulk_load(dst, w_tile + (size_t)step * kWStepBytes, half_bytes, bar_wloc + 8 * s);
      bulk_load(dst + kBBytes / 2, w_tile + (size_t)step * kWStepBytes + kBBytes, half_bytes,
                bar_wloc + 8 * s);
    };
    if (lane == 0)
      for (int step = 0; step < kPairStages && step < n_steps; ++step) fetch_w(step);
    const int tile_rows = min(kBlockM, a.rows - m0);
    if (tile_rows > 0) {
      const size_t x_bytes = (size_t)tile_rows * a.K * 4;
      const char* xp = reinterpret_cast<const char*>(a.x + (size_t)m0 * a.K);
      for (size_t off = (size_t)lane * 32768; off < x_bytes; off += 32 * 32768) {
        const size_t nb = x_bytes - off < 32768 ? x_bytes - off : 32768;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xp + off), "r"((uint32_t)nb)
                     : "memory");
      }
      if (a.flags & (PCT_LINEAR_RESIDUAL | PCT_LINEAR_GATE)) {
        for (int r = lane; r < tile_rows; r += 32)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(
                           a.residual + (size_t)(m0 + r) * a.N + n0),
                       "r"((uint32_t)(n_tile * 4))
                       : "memory");
      }
    }
    if (lane == 0)
      for (int step = kPairStages; step < n_steps; ++step) {
        mbar_wait(bar_empty + 8 * (step % kPairStages), ((step / kPairStages) - 1) & 1);
        fetch_w(step);
      }
  } else {
    // ---- loaders (both CTAs): as in the single-CTA kernel, arriving on the LEADER's barrier ----
    const int r8 = tid & 7, kc = (tid >> 3) & 3, rg = tid >> 5;
    constexpr int kPerThread = kBlockM * kBlockK / 4 / (kLoaderWarps * 32);
    const float* xrow[kPerThread];
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int row = m0 + (rg + kLoaderWarps * j) * 8 + r8;
      xrow[j] = row < a.rows ? a.x + (size_t)row * a.K + kc * 4 : nullptr;
    }
    float4 xq[kAhead][kPerThread];
    auto fetch_x = [&](float4* dst, int step) {
#pragma unroll
      for (int j = 0; j < kPerThread; ++j)
        dst[j] = xrow[j] ? __ldg(reinterpret_cast<const float4*>(xrow[j] + step * kBlockK))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
    };
#pragma unroll
    for (int d = 0; d < kAhead; ++d)
      if (d < n_steps) fetch_x(xq[d], d);
    for (int step0 = 0; step0 < n_steps; step0 += kAhead) {
#pragma unroll
      for (int d = 0; d < kAhead; ++d) {
        const int step = step0 + d;
        if (step < n_steps) {
          const int s = step % kPairStages;
          if (step >= kPairStages)
            mbar_wait(bar_empty + 8 * s, ((step / kPairStages) - 1) & 1);
          float4* a_hi = reinterpret_cast<float4*>(smem + s * kPairStageBytes);
          float4* a_lo = reinterpret_cast<float4*>(smem + s * kPairStageBytes + kABytes);
#pragma unroll
          for (int j = 0; j < kPerThread; ++j) {
            const float4 x4 = xq[d][j];
            float4 h, l;
            h.x = tf32_head(x4.x), h.y = tf32_head(x4.y), h.z = tf32_head(x4.z), h.w = tf32_head(x4.w);
            l.x = x4.x - h.x, l.y = x4.y - h.y, l.z = x4.z - h.z, l.w = x4.w - h.w;
            a_hi[tid + kLoaderWarps * 32 * j] = h;
            a_lo[tid + kLoaderWarps * 32 * j] = l;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();                                  // one (remote) arrival per warp
          if (lane == 0) mbar_arrive_remote(leader_ready + 8 * s);
          if (step + kAhead < n_steps) fetch_x(xq[d], step + kAhead);
        }
      }
    }

    mbar_wait_cluster(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    linear_epilogue(a, smem, tmem, vec_bias, vec_gamma, vec_beta, warp, lane, tid, m0, n0, n_tile);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                  // nobody leaves while the partner may still signal or read
  if (warp == kLoaderWarps) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"((uint32_t)kBlockN)
                 : "memory");
  }
}

// W (N, K) fp32 -> per (256-column tile, 16-float K step): [hi piece 16 KB][lo piece 16 KB], each in
// the stage's core-matrix order (float4 i <-> column 8*(i/32) + i%8, k-quad (i/8)%4); columns
// past N are zero.
__global__ void pack_weight_kernel(const float* __restrict__ w, float4* __restrict__ packed, int N,
                                   int K) {
  const int n_steps = K / kBlockK;
  const size_t per_piece = kBBytes / 16;                       // float4 per piece
  const size_t total = (size_t)((N + kBlockN - 1) / kBlockN) * n_steps * per_piece;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % per_piece);
    const size_t ts = idx / per_piece;                         // tile * n_steps + step
    const int step = (int)(ts % n_steps), tile = (int)(ts / n_steps);
    const int col = tile * kBlockN + (i >> 5) * 8 + (i & 7);
    const int k = step * kBlockK + ((i >> 3) & 3) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < N) x = *reinterpret_cast<const float4*>(w + (size_t)col * K + k);
    float4 h, l;
    h.x = tf32_head(x.x), h.y = tf32_head(x.y), h.z = tf32_head(x.z), h.w = tf32_head(x.w);
    l.x = x.x - h.x, l.y = x.y - h.y, l.z = x.z - h.z, l.w = x.w - h.w;
    packed[ts * 2 * per_piece + i] = h;
    packed[ts * 2 * per_piece + per_piece + i] = l;
  }
}

// bf16 operand precision: W (N, K) fp32 -> per (256-column tile, 32-element K step) one 16 KB piece of
// bf16 in the same core-matrix order (16-byte chunk i <-> column 8*(i/32) + i%8, elements
// 8*((i/8)%4) .. +7 of the step); columns past N are zero.
__global__ void pack_weight_bf16_kernel(const float* __restrict__ w, uint4* __restrict__ packed, int N,
                                        int K) {
  const int n_steps = K / Prec<true>::kStepK;
  const size_t per_piece = kBBytes / 16;
  const size_t total = (size_t)((N + kBlockN - 1) / kBlockN) * n_steps * per_piece;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % per_piece);
    const size_t ts = idx / per_piece;                         // tile * n_steps + step
    const int step = (int)(ts % n_steps), tile = (int)(ts / n_steps);
    const int col = tile * kBlockN + (i >> 5) * 8 + (i & 7);
    const int k = step * Prec<true>::kStepK + ((i >> 3) & 3) * 8;
    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
    if (col < N) {
      x0 = *reinterpret_cast<const float4*>(w + (size_t)col * K + k);
      x1 = *reinterpret_cast<const float4*>(w + (size_t)col * K + k + 4);
    }
    packed[idx] = pack_bf16x8(x0, x1);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient: dW (N, K) = dY^T (N, rows) @ X (rows, K), the reduction running over the rows.
// Both operands are activations, so both are split on the fly, and both arrive transposed: global
// memory is contiguous along the features, the tensor core wants 16-byte chunks along the
// reduction (row) dimension.  A loader thread therefore reads a 4-row x 4-feature block (four
// 16-byte loads, a warp covering 512 contiguous bytes of each row), transposes it in registers and
// stores four 16-byte chunks in the same K-major core-matrix order the forward kernel uses; the
// 8-row groups are padded to 528 bytes so those stores are bank-conflict free.  A CTA owns a
// 128 x (<=256) tile of dW and a slice of the rows (split-K); slices are combined with vector fp32
// reductions into the zero-filled dW.
constexpr int kWgLoaderWarps = 8;
constexpr int kWgThreads = (kWgLoaderWarps + 1) * 32;
constexpr int kWgAhead = 2;
constexpr int kWgSbo = kSbo + 16;                              // padded 8-feature group
constexpr int kWgABytes = (kBlockM / 8) * kWgSbo;              // one piece of the dY^T tile
constexpr int kWgBBytes = (kBlockN / 8) * kWgSbo;              // one piece of the X^T tile
constexpr int kWgStageBytes = 2 * kWgABytes + 2 * kWgBBytes;
constexpr int kWgSmemBytes = kStages * kWgStageBytes + kBarrierBytes;
constexpr int kWgUnitsA = (kBlockM / 4) * (kBlockK / 4);       // 4x4 blocks of the dY tile per step
constexpr int kWgUnits = kWgUnitsA + (kBlockN / 4) * (kBlockK / 4);
constexpr int kWgPer = (kWgUnits + kWgLoaderWarps * 32 - 1) / (kWgLoaderWarps * 32);

__device__ __forceinline__ uint64_t make_smem_desc_wg(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kLbo >> 4) << 16) |
         ((uint64_t)(kWgSbo >> 4) << 32) | ((uint64_t)1 << 46);
}

struct WgradArgs {
  unsigned flags;    // PCT_LINEAR_X_BF16 / PCT_LINEAR_GY_BF16: that operand is bf16 in global memory
  const float* gy;   // (rows, N)
  const float* x;    // (rows, K)
  float* gw;         // (N, K), zero on entry
  float* gb;         // (N,) column sums of gy, zero on entry; may be null
  int rows, K, N;
  int steps_per_split;
  const float* gy2;  // second part of a column-concatenated grad_y, (rows, N - n_split); n_split = 0: none
  int n_split;
};

// BF16 (PCT_LINEAR_BF16): both operands rounded to bf16; a K step is then 32 rows (a 16-byte chunk
// holds 8 consecutive rows of one feature), the shared-memory geometry is the same, one piece
// per operand and two kind::f16 MMAs per step instead of six kind::tf32.
// GY16 / X16 (bf16 kernel only): that operand is bf16 in global memory (PCT_LINEAR_GY_BF16 /
// PCT_LINEAR_X_BF16): its units load four bf16 (8 bytes) per row, kept raw until the step is
// stored (a conversion next to the load would wait for every load in turn) and widened exactly.
template <bool BF16, bool GY16 = false, bool X16 = false>
__device__ __forceinline__ void wgrad_body(const WgradArgs& a, unsigned char* smem) {
  static_assert(BF16 || !(GY16 || X16), "bf16 tensors only with the bf16 kernel");
  constexpr int kRowsPerChunk = BF16 ? 8 : 4;               // rows (K) in one 16-byte chunk
  constexpr int kRowsPerStep = 4 * kRowsPerChunk;           // four chunk columns per stage
  constexpr int kPieces = BF16 ? 1 : 2;
  constexpr int kWgStage = kPieces * (kWgABytes + kWgBBytes);
  constexpr int kAheadW = BF16 ? 1 : kWgAhead;              // the same bytes in flight either way
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n0 = blockIdx.x * kBlockM;                   // dW rows = output features
  const int k0 = blockIdx.y * kBlockN;                   // dW columns = input features
  const int n_valid = min(kBlockM, a.N - n0);
  const int k_tile = min(kBlockN, a.K - k0);
  const int total_steps = (a.rows + kRowsPerStep - 1) / kRowsPerStep;
  const int step_lo = blockIdx.z * a.steps_per_split;
  const int n_steps = min(a.steps_per_split, total_steps - step_lo);
  if (n_steps <= 0) return;                              // uniform per CTA, before any barrier

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kWgStage);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);
  const uint32_t bar_empty = smem_u32(bars);
  const uint32_t bar_full = bar_empty + 8 * kStages;
  const uint32_t bar_done = bar_empty + 16 * kStages;

  if (warp == kWgLoaderWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)kBlockN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_full + 8 * s, kWgLoaderWarps);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t smem0 = smem_u32(smem);

  if (warp == kWgLoaderWarps) {
    if (lane == 0) {
      const uint32_t idesc = BF16 ? make_instr_desc_bf16(k_tile) : make_instr_desc(k_tile);
      for (int step = 0; step < n_steps; ++step) {
        const int s = step % kStages;
        mbar_wait(bar_full + 8 * s, (step / kStages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem0 + s * kWgStage;
#pragma unroll
        for (int ks = 0; ks < kBlockK / kMmaK; ++ks) {
          const uint32_t koff = ks * (kMmaK / 4) * kLbo;
          if constexpr (BF16) {
            mma_bf16(tmem, make_smem_desc_wg(sa + koff), make_smem_desc_wg(sa + kWgABytes + koff), idesc,
                     (step | ks) != 0);
          } else {
            const uint64_t d_ahi = make_smem_desc_wg(sa + koff);
            const uint64_t d_alo = make_smem_desc_wg(sa + kWgABytes + koff);
            const uint64_t d_bhi = make_smem_desc_wg(sa + 2 * kWgABytes + koff);
            const uint64_t d_blo = make_smem_desc_wg(sa + 2 * kWgABytes + kWgBBytes + koff);
            mma_tf32(tmem, d_alo, d_bhi, idesc, (step | ks) != 0);
            mma_tf32(tmem, d_ahi, d_bhi, idesc, 1u);
            mma_tf32(tmem, d_ahi, d_blo, idesc, 1u);
          }
        }
        mma_commit(bar_empty + 8 * s);
        if (step == n_steps - 1) mma_commit(bar_done);
      }
    }
  } else {
    // ---- loaders.  Unit = 4 rows x 4 features (bf16: 8 rows x 4 features); unit ids [0, kWgUnitsA) belong to the dY tile,
    // the rest to the X tile; within a tile, id = featureGroup + groups * rowQuad, so a warp reads
    // 512 contiguous bytes of one row per load.
    const char* src[kWgPer];      // first of the unit's 4 features, row 0 (nullptr: past the edge)
    bool s16[kWgPer];             // the unit's operand is bf16 in global memory
    int rq[kWgPer];               // row quad within the step
    uint32_t dst[kWgPer];         // byte offset of the unit's first chunk within the hi piece
    uint32_t lo_off[kWgPer];      // distance from the hi piece to the lo piece
    int ld[kWgPer];               // row pitch of the source
#pragma unroll
    for (int j = 0; j < kWgPer; ++j) {
      const int u = tid + kWgLoaderWarps * 32 * j;
      const bool is_a = u < kWgUnitsA;
      const int v = is_a ? u : u - kWgUnitsA;
      const int groups = is_a ? kBlockM / 4 : kBlockN / 4;
      const int g = v % groups;
      rq[j] = v / groups;
      const bool ok = u < kWgUnits && 4 * g < (is_a ? n_valid : k_tile);
      s16[j] = is_a ? GY16 : X16;
      const int esz = s16[j] ? 2 : 4;
      // grad_y may be the column-concatenation [gy (rows, n_split) | gy2 (rows, N - n_split)]
      // (pct_linear_wgrad_split; n_split is a multiple of 4, so a unit lies in one of the two)
      const int feat = n0 + 4 * g;
      const bool second = is_a && a.n_split > 0 && feat >= a.n_split;
      src[j] = !ok ? nullptr
                   : is_a ? (second ? reinterpret_cast<const char*>(a.gy2) + (size_t)(feat - a.n_split) * esz
                                    : reinterpret_cast<const char*>(a.gy) + (size_t)feat * esz)
                          : reinterpret_cast<const char*>(a.x) + (size_t)(k0 + 4 * g) * esz;
      ld[j] = (is_a ? (a.n_split > 0 ? (second ? a.N - a.n_split : a.n_split) : a.N) : a.K) * esz;   // row pitch in bytes
      // features 4g..4g+3 are rows (4g+i)%8 of 8-row group (4g)/8; the row quad is the k-chunk
      dst[j] = (is_a ? 0 : kPieces * kWgABytes) + (g >> 1) * kWgSbo + rq[j] * kCoreBytes + (g & 1) * 64;
      lo_off[j] = is_a ? kWgABytes : kWgBBytes;
      if (u >= kWgUnits) dst[j] = 0xffffffffu;
    }
    float4 q[kAheadW][kWgPer][kRowsPerChunk];
    // bias gradient = column sums of dY: the threads that load dY blocks keep running sums of
    // their four features (first k tile only, so that every (row, feature) is counted once)
    const bool sum_bias = a.gb != nullptr && blockIdx.y == 0 && tid < kWgUnitsA && src[0] != nullptr;
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](float4 (*d)[kRowsPerChunk], int step) {
      const int row0 = (step_lo + step) * kRowsPerStep;
#pragma unroll
      for (int j = 0; j < kWgPer; ++j)
#pragma unroll
        for (int i = 0; i < kRowsPerChunk; ++i) {
          const int row = row0 + kRowsPerChunk * rq[j] + i;
          const bool in = src[j] && row < a.rows;
          if ((GY16 || X16) && s16[j]) {
            const uint2 u = in ? __ldg(reinterpret_cast<const uint2*>(src[j] + (size_t)row * ld[j]))
                               : make_uint2(0u, 0u);
            d[j][i] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), 0.f, 0.f);
          } else {
            d[j][i] = in ? __ldg(reinterpret_cast<const float4*>(src[j] + (size_t)row * ld[j]))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
    };
    auto split_store = [&](unsigned char* hi, unsigned char* lo, float x0, float x1, float x2, float x3) {
      float4 h, l;
      h.x = tf32_head(x0), h.y = tf32_head(x1), h.z = tf32_head(x2), h.w = tf32_head(x3);
      l.x = x0 - h.x, l.y = x1 - h.y, l.z = x2 - h.z, l.w = x3 - h.w;
      *reinterpret_cast<float4*>(hi) = h;
      *reinterpret_cast<float4*>(lo) = l;
    };
#pragma unroll
    for (int d = 0; d < kAheadW; ++d)
      if (d < n_steps) fetch(q[d], d);
    for (int step0 = 0; step0 < n_steps; step0 += kAheadW) {
#pragma unroll
      for (int d = 0; d < kAheadW; ++d) {
        const int step = step0 + d;
        if (step < n_steps) {
          const int s = step % kStages;
          if (step >= kStages) mbar_wait(bar_empty + 8 * s, ((step / kStages) - 1) & 1);
          unsigned char* stage = smem + s * kWgStage;
          if constexpr (GY16 || X16) {
#pragma unroll
            for (int j = 0; j < kWgPer; ++j)
              if (s16[j]) {
#pragma unroll
                for (int i = 0; i < kRowsPerChunk; ++i)
                  q[d][j][i] = bf16x4_to_float4(
                      make_uint2(__float_as_uint(q[d][j][i].x), __float_as_uint(q[d][j][i].y)));
              }
          }
          if (sum_bias) {
#pragma unroll
            for (int i = 0; i < kRowsPerChunk; ++i)
              bsum.x += q[d][0][i].x, bsum.y += q[d][0][i].y, bsum.z += q[d][0][i].z,
                  bsum.w += q[d][0][i].w;
          }
#pragma unroll
          for (int j = 0; j < kWgPer; ++j)
            if (dst[j] != 0xffffffffu) {
              unsigned char* hi = stage + dst[j];
              unsigned char* lo = hi + lo_off[j];
              const float4* r = q[d][j];      // r[i] = features 4g..4g+3 of row i: transpose
              if constexpr (BF16) {
                uint4* c = reinterpret_cast<uint4*>(hi);
                c[0] = pack_bf16x8(make_float4(r[0].x, r[1].x, r[2].x, r[3].x),
                                   make_float4(r[4].x, r[5].x, r[6].x, r[7].x));
                c[1] = pack_bf16x8(make_float4(r[0].y, r[1].y, r[2].y, r[3].y),
                                   make_float4(r[4].y, r[5].y, r[6].y, r[7].y));
                c[2] = pack_bf16x8(make_float4(r[0].z, r[1].z, r[2].z, r[3].z),
                                   make_float4(r[4].z, r[5].z, r[6].z, r[7].z));
                c[3] = pack_bf16x8(make_float4(r[0].w, r[1].w, r[2].w, r[3].w),
                                   make_float4(r[4].w, r[5].w, r[6].w, r[7].w));
              } else {
                split_store(hi, lo, r[0].x, r[1].x, r[2].x, r[3].x);
                split_store(hi + 16, lo + 16, r[0].y, r[1].y, r[2].y, r[3].y);
                split_store(hi + 32, lo + 32, r[0].z, r[1].z, r[2].z, r[3].z);
                split_store(hi + 48, lo + 48, r[0].w, r[1].w, r[2].w, r[3].w);
              }
            }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_full + 8 * s) : "memory");
          if (step + kAheadW < n_steps) fetch(q[d], step + kAheadW);
        }
      }
    }

    if (sum_bias)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.gb + n0 + 4 * (tid & (kBlockM / 4 - 1))),
                   "f"(bsum.x), "f"(bsum.y), "f"(bsum.z), "f"(bsum.w)
                   : "memory");

    // ---- epilogue: this thread owns dW row n0 + 32*(warp%4) + lane, one half of the columns ----
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int c_lo = (warp >> 2) * (kBlockN / 2);
    const int c_hi = min(k_tile, c_lo + kBlockN / 2);
    float* out = a.gw + (size_t)(n0 + row) * a.K + k0;
    for (int c = c_lo; c < c_hi; c += 16) {
      uint32_t u[16];
      tmem_ld16(t_row + c, u);
      tmem_ld_wait();
      tmem_tie16(u);
      if (row < n_valid) {
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c + 4 * qd),
                       "f"(__uint_as_float(u[4 * qd])), "f"(__uint_as_float(u[4 * qd + 1])),
                       "f"(__uint_as_float(u[4 * qd + 2])), "f"(__uint_as_float(u[4 * qd + 3]))
                       : "memory");
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kWgLoaderWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"((uint32_t)kBlockN)
                 : "memory");
  }
}

__global__ void __launch_bounds__(kWgThreads, 2) wgrad_tf32x3_kernel(const WgradArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  wgrad_body<false>(a, smem);
}

template <bool GY16, bool X16>
__global__ void __launch_bounds__(kWgThreads, 2) wgrad_bf16_kernel(const WgradArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  wgrad_body<true, GY16, X16>(a, smem);
}

bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

bool bad_dims(int in_features, int out_features) {
  return in_features < kBlockK || out_features < 16 || in_features % kBlockK || out_features % 16;
}

}  // namespace

extern "C" {

size_t pct_linear_packed_bytes(int in_features, int out_features) {
  if (bad_dims(in_features, out_features)) return 0;
  return (size_t)((out_features + kBlockN - 1) / kBlockN) * (in_features / kBlockK) * kWStepBytes;
}

int pct_linear_pack_weight(const float* w, void* packed, int in_features, int out_features,
                           void* cuda_stream) {
  if (!w || !packed) return PCT_MSDA_E_NULL_POINTER;
  if (bad_dims(in_features, out_features)) return PCT_MSDA_E_BAD_SHAPE;
  if (misaligned(w) || misaligned(packed)) return PCT_MSDA_E_MISALIGNED;
  const size_t total = pct_linear_packed_bytes(in_features, out_features) / 32;
  const int block = 256;
  const size_t want = (total + block - 1) / block;
  pack_weight_kernel<<<(int)(want < 1184 ? want : 1184), block, 0, (cudaStream_t)cuda_stream>>>(
      w, reinterpret_cast<float4*>(packed), out_features, in_features);
  pct::count_launch();
  return (int)cudaGetLastError();
}

size_t pct_linear_packed_bytes_bf16(int in_features, int out_features) {
  if (bad_dims(in_features, out_features) || in_features % Prec<true>::kStepK) return 0;
  return (size_t)((out_features + kBlockN - 1) / kBlockN) * (in_features / Prec<true>::kStepK) *
         Prec<true>::kWStep;
}

int pct_linear_pack_weight_bf16(const float* w, void* packed, int in_features, int out_features,
                                void* cuda_stream) {
  if (!w || !packed) return PCT_MSDA_E_NULL_POINTER;
  if (bad_dims(in_features, out_features) || in_features % Prec<true>::kStepK)
    return PCT_MSDA_E_BAD_SHAPE;
  if (misaligned(w) || misaligned(packed)) return PCT_MSDA_E_MISALIGNED;
  const size_t total = pct_linear_packed_bytes_bf16(in_features, out_features) / 16;
  const int block = 256;
  const size_t want = (total + block - 1) / block;
  pack_weight_bf16_kernel<<<(int)(want < 1184 ? want : 1184), block, 0, (cudaStream_t)cuda_stream>>>(
      w, reinterpret_cast<uint4*>(packed), out_features, in_features);
  pct::count_launch();
  return (int)cudaGetLastError();
}

#ifdef PCT_LINEAR_TRACE
// variant builds only (scripts/linear_trace.py): the phase timestamps of the last launch
int pct_linear_debug_trace(long long* host_out, int n_values) {
  return (int)cudaMemcpyFromSymbol(host_out, g_linear_trace, (size_t)n_values * sizeof(long long));
}
#endif

int pct_linear_forward(const float* x, const void* w_packed, const float* bias,
                       const float* residual, const float* ln_gamma, const float* ln_beta,
                       float ln_eps, float* y, int rows, int in_features, int out_features,
                       unsigned flags, void* cuda_stream) {
  return pct_linear_forward_train(x, w_packed, bias, residual, ln_gamma, ln_beta, ln_eps, y, nullptr,
                                  nullptr, nullptr, rows, in_features, out_features, flags,
                                  cuda_stream);
}

int pct_linear_forward_train(const float* x, const void* w_packed, const float* bias,
                             const float* residual, const float* ln_gamma, const float* ln_beta,
                             float ln_eps, float* y, float* pre_norm, float* ln_mean,
                             float* ln_rstd, int rows, int in_features, int out_features,
                             unsigned flags, void* cuda_stream) {
  if (!x || !w_packed || !y) return PCT_MSDA_E_NULL_POINTER;
  if ((ln_mean == nullptr) != (ln_rstd == nullptr)) return PCT_MSDA_E_NULL_POINTER;
  if ((pre_norm || ln_mean) && !(flags & PCT_LINEAR_LAYERNORM)) return PCT_MSDA_E_UNSUPPORTED;
  if (misaligned(pre_norm)) return PCT_MSDA_E_MISALIGNED;
  if ((flags & (PCT_LINEAR_RESIDUAL | PCT_LINEAR_GATE | PCT_LINEAR_ROWMASK)) && !residual)
    return PCT_MSDA_E_NULL_POINTER;
  if ((flags & PCT_LINEAR_LAYERNORM) && (!ln_gamma || !ln_beta)) return PCT_MSDA_E_NULL_POINTER;
  if (rows < 1 || bad_dims(in_features, out_features)) return PCT_MSDA_E_BAD_SHAPE;
  if ((flags & PCT_LINEAR_LAYERNORM) && out_features > kBlockN) return PCT_MSDA_E_UNSUPPORTED;
  if (flags & ~(PCT_LINEAR_RELU | PCT_LINEAR_RESIDUAL | PCT_LINEAR_LAYERNORM | PCT_LINEAR_GATE |
                PCT_LINEAR_ROWMASK | PCT_LINEAR_BF16 | PCT_LINEAR_X_BF16 | PCT_LINEAR_Y_BF16 |
                PCT_LINEAR_RES_BF16))
    return PCT_MSDA_E_UNSUPPORTED;
  // bf16 tensors in global memory only with the bf16 kernel; a bf16 residual only where one is read
  if ((flags & (PCT_LINEAR_X_BF16 | PCT_LINEAR_Y_BF16 | PCT_LINEAR_RES_BF16)) && !(flags & PCT_LINEAR_BF16))
    return PCT_MSDA_E_UNSUPPORTED;
  if ((flags & PCT_LINEAR_RES_BF16) && !(flags & (PCT_LINEAR_RESIDUAL | PCT_LINEAR_GATE)))
    return PCT_MSDA_E_UNSUPPORTED;
  if ((flags & PCT_LINEAR_BF16) && in_features % Prec<true>::kStepK) return PCT_MSDA_E_BAD_SHAPE;
  if ((flags & PCT_LINEAR_GATE) && (flags & (PCT_LINEAR_RESIDUAL | PCT_LINEAR_LAYERNORM)))
    return PCT_MSDA_E_UNSUPPORTED;
  if ((flags & PCT_LINEAR_ROWMASK) &&
      (flags & (PCT_LINEAR_RESIDUAL | PCT_LINEAR_LAYERNORM | PCT_LINEAR_GATE)))
    return PCT_MSDA_E_UNSUPPORTED;
  // 16-byte accesses: x, packed W, residual, y, pre_norm; bias / gamma / beta are read as scalars
  // (and the row mask as bytes)
  if (misaligned(x) || misaligned(w_packed) || misaligned(y) ||
      (!(flags & PCT_LINEAR_ROWMASK) && misaligned(residual)))
    return PCT_MSDA_E_MISALIGNED;
  if (((uintptr_t)bias | (uintptr_t)ln_gamma | (uintptr_t)ln_beta | (uintptr_t)ln_mean |
       (uintptr_t)ln_rstd) & 3u)
    return PCT_MSDA_E_MISALIGNED;
  const long long m_tiles = ((long long)rows + kBlockM - 1) / kBlockM;
  if (m_tiles > 65535) return PCT_MSDA_E_TOO_LARGE;

  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    e = cudaFuncSetAttribute(linear_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  LinearArgs a{x, static_cast<const unsigned char*>(w_packed), bias, residual, ln_gamma, ln_beta,
               y, pre_norm, ln_mean, ln_rstd, ln_eps, rows, in_features, out_features, flags,
               nullptr, nullptr, 0, 0};
  dim3 grid((out_features + kBlockN - 1) / kBlockN, (unsigned)m_tiles);
  if (flags & PCT_LINEAR_BF16) {
    static bool bf16_attr[64] = {};
    if (dev < 0 || dev >= 64 || !bf16_attr[dev]) {
      e = cudaFuncSetAttribute(linear_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Prec<true>::kSmem);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(linear_bf16_io16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Prec<true>::kSmem);
      if (e != cudaSuccess) return (int)e;
      if (dev >= 0 && dev < 64) bf16_attr[dev] = true;
    }
    if (flags & (PCT_LINEAR_X_BF16 | PCT_LINEAR_Y_BF16 | PCT_LINEAR_RES_BF16))
      linear_bf16_io16_kernel<<<grid, kThreads, Prec<true>::kSmem, (cudaStream_t)cuda_stream>>>(a);
    else
      linear_bf16_kernel<<<grid, kThreads, Prec<true>::kSmem, (cudaStream_t)cuda_stream>>>(a);
    pct::count_launch();
    return (int)cudaGetLastError();
  }
  static const int use_pair = [] {
    const char* e = getenv("PCT_LINEAR_PAIR");
    return e ? atoi(e) : 0;
  }();
  if (use_pair) {
    static bool pair_attr[64] = {};
    if (dev < 0 || dev >= 64 || !pair_attr[dev]) {
      e = cudaFuncSetAttribute(linear_tf32x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               kPairSmemBytes);
      if (e != cudaSuccess) return (int)e;
      if (dev >= 0 && dev < 64) pair_attr[dev] = true;
    }
    grid = dim3((unsigned)((m_tiles + 1) / 2 * 2), grid.x);   // whole pairs along x; a padding CTA has no rows
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kPairSmemBytes;
    cfg.stream = (cudaStream_t)cuda_stream;
    cudaLaunchAttribute attr = {};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, linear_tf32x3_pair_kernel, a);
    pct::count_launch();
    return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
  }
  linear_tf32x3_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)cuda_stream>>>(a);
  pct::count_launch();
  return (int)cudaGetLastError();
}

int pct_linear_forward_split(const float* x_a, const float* x_b, int in_a, int in_b,
                             const void* w_packed, const float* bias, const float* residual,
                             float* y_a, float* y_b, int out_a, int out_b, int rows, unsigned flags,
                             void* cuda_stream) {
  if (!x_a || !w_packed || !y_a) return PCT_MSDA_E_NULL_POINTER;
  if ((in_b > 0 && !x_b) || (out_b > 0 && !y_b)) return PCT_MSDA_E_NULL_POINTER;
  if (in_b < 0 || out_b < 0) return PCT_MSDA_E_BAD_SHAPE;
  if (flags & ~(PCT_LINEAR_RELU | PCT_LINEAR_RESIDUAL | PCT_LINEAR_BF16)) return PCT_MSDA_E_UNSUPPORTED;
  if ((flags & PCT_LINEAR_RESIDUAL) && (!residual || out_b > 0)) return PCT_MSDA_E_UNSUPPORTED;
  const bool bf16 = flags & PCT_LINEAR_BF16;
  const int in_features = in_a + in_b, out_features = out_a + out_b;
  const int kstep = bf16 ? Prec<true>::kStepK : Prec<false>::kStepK;
  if (rows < 1 || bad_dims(in_features, out_features) || in_features % kstep) return PCT_MSDA_E_BAD_SHAPE;
  if ((in_b > 0 && in_a % kstep) || (out_b > 0 && out_a % 32)) return PCT_MSDA_E_BAD_SHAPE;
  if (misaligned(x_a) || misaligned(x_b) || misaligned(w_packed) || misaligned(y_a) || misaligned(y_b) ||
      misaligned(residual) || (in_b > 0 && in_b % 4) || (out_b > 0 && out_b % 4))
    return PCT_MSDA_E_MISALIGNED;
  if (((uintptr_t)bias) & 3u) return PCT_MSDA_E_MISALIGNED;
  const long long m_tiles = ((long long)rows + kBlockM - 1) / kBlockM;
  if (m_tiles > 65535) return PCT_MSDA_E_TOO_LARGE;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    e = cudaFuncSetAttribute(linear_tf32x3_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(linear_bf16_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Prec<true>::kSmem);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  LinearArgs a{x_a, static_cast<const unsigned char*>(w_packed), bias, residual, nullptr, nullptr,
               y_a, nullptr, nullptr, nullptr, 0.f, rows, in_features, out_features, flags,
               x_b, y_b, in_b > 0 ? in_a : 0, out_b > 0 ? out_a : 0};
  dim3 grid((out_features + kBlockN - 1) / kBlockN, (unsigned)m_tiles);
  if (bf16)
    linear_bf16_split_kernel<<<grid, kThreads, Prec<true>::kSmem, (cudaStream_t)cuda_stream>>>(a);
  else
    linear_tf32x3_split_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)cuda_stream>>>(a);
  pct::count_launch();
  return (int)cudaGetLastError();
}

int pct_linear_wgrad(const float* grad_y, const float* x, float* grad_w, float* grad_bias, int rows,
                     int in_features, int out_features, void* cuda_stream) {
  return pct_linear_wgrad_ex(grad_y, x, grad_w, grad_bias, rows, in_features, out_features, 0u,
                             cuda_stream);
}

int pct_linear_wgrad_ex(const float* grad_y, const float* x, float* grad_w, float* grad_bias, int rows,
                        int in_features, int out_features, unsigned flags, void* cuda_stream) {
  return pct_linear_wgrad_split(grad_y, nullptr, out_features, 0, x, grad_w, grad_bias, rows, in_features,
                                flags, cuda_stream);
}

int pct_linear_wgrad_split(const float* grad_y, const float* grad_y_b, int out_a, int out_b, const float* x,
                           float* grad_w, float* grad_bias, int rows, int in_features, unsigned flags,
                           void* cuda_stream) {
  if (out_b < 0 || (out_b > 0 && (!grad_y_b || out_a % 4 || out_b % 4 || misaligned(grad_y_b))))
    return out_b > 0 && !grad_y_b ? PCT_MSDA_E_NULL_POINTER : PCT_MSDA_E_BAD_SHAPE;
  if (out_b > 0 && (flags & PCT_LINEAR_GY_BF16)) return PCT_MSDA_E_UNSUPPORTED;
  const int out_features = out_a + out_b;
  if (!grad_y || !x || !grad_w) return PCT_MSDA_E_NULL_POINTER;
  if (flags & ~(PCT_LINEAR_BF16 | PCT_LINEAR_X_BF16 | PCT_LINEAR_GY_BF16)) return PCT_MSDA_E_UNSUPPORTED;
  const bool bf16 = flags & PCT_LINEAR_BF16;
  if (!bf16 && (flags & (PCT_LINEAR_X_BF16 | PCT_LINEAR_GY_BF16))) return PCT_MSDA_E_UNSUPPORTED;
  if (rows < 1 || bad_dims(in_features, out_features)) return PCT_MSDA_E_BAD_SHAPE;
  if (misaligned(grad_y) || misaligned(x) || misaligned(grad_w) || misaligned(grad_bias))
    return PCT_MSDA_E_MISALIGNED;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cudaError_t e = cudaMemsetAsync(grad_w, 0, (size_t)in_features * out_features * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  if (grad_bias) {
    e = cudaMemsetAsync(grad_bias, 0, (size_t)out_features * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
  }

  static bool attr_set[64] = {};
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    e = cudaFuncSetAttribute(wgrad_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kWgSmemBytes);
    if (e != cudaSuccess) return (int)e;
    for (auto* k : {wgrad_bf16_kernel<false, false>, wgrad_bf16_kernel<true, false>,
                    wgrad_bf16_kernel<false, true>, wgrad_bf16_kernel<true, true>}) {
      e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes);
      if (e != cudaSuccess) return (int)e;
    }
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  // split the rows so that about two CTAs per SM are busy; at least 8 K steps per slice
  const int tiles = ((out_features + kBlockM - 1) / kBlockM) * ((in_features + kBlockN - 1) / kBlockN);
  const int rows_per_step = bf16 ? 32 : kBlockK;
  const int total_steps = (rows + rows_per_step - 1) / rows_per_step;
  int splits = (2 * 148 + tiles - 1) / tiles;
  if (splits > (total_steps + 7) / 8) splits = (total_steps + 7) / 8;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  const int steps_per_split = (total_steps + splits - 1) / splits;
  splits = (total_steps + steps_per_split - 1) / steps_per_split;
  WgradArgs a{flags, grad_y, x, grad_w, grad_bias, rows, in_features, out_features, steps_per_split,
              grad_y_b, out_b > 0 ? out_features - out_b : 0};
  dim3 grid((out_features + kBlockM - 1) / kBlockM, (in_features + kBlockN - 1) / kBlockN, splits);
  if (bf16) {
    auto* k = (flags & PCT_LINEAR_GY_BF16)
                  ? ((flags & PCT_LINEAR_X_BF16) ? wgrad_bf16_kernel<true, true> : wgrad_bf16_kernel<true, false>)
                  : ((flags & PCT_LINEAR_X_BF16) ? wgrad_bf16_kernel<false, true> : wgrad_bf16_kernel<false, false>);
    k<<<grid, kWgThreads, kWgSmemBytes, st>>>(a);
  } else
    wgrad_tf32x3_kernel<<<grid, kWgThreads, kWgSmemBytes, st>>>(a);
  pct::count_launch();
  return (int)cudaGetLastError();
}

}  // extern "C"
