/*
 * pct_msda.h -- C ABI of the B200-native multi-scale deformable attention (MSDeformAttn) path.
 *
 * This is the drop-in boundary for PCTrans's pixel-decoder op.  It replaces the pybind11 module
 * `MultiScaleDeformableAttention` of the reference
 * (OPS = connectomics/model/maskformer_block/pixel_decoder/ops):
 *
 *   pct_msda_forward   <-  ms_deform_attn_forward   OPS/src/vision.cpp:19,
 *                          OPS/src/ms_deform_attn.h:25-44, OPS/src/cuda/ms_deform_attn_cuda.cu:25-85
 *   pct_msda_backward  <-  ms_deform_attn_backward  OPS/src/vision.cpp:20,
 *                          OPS/src/ms_deform_attn.h:46-66, OPS/src/cuda/ms_deform_attn_cuda.cu:88-158
 *
 * Plain C: pointers and sizes only, no torch / ATen types, no C++ in the signatures.  The library
 * links against the CUDA runtime only.  See INTEGRATION.md for the ctypes binding the reference's
 * OPS/functions/ms_deform_attn_func.py would use.
 *
 * Tensor layouts (row-major, contiguous, identical to the reference, cu:45-53):
 *   value              (N, S, M, D)         value_dtype
 *   spatial_shapes     (L, 2) int64 rows (H_l, W_l)            DEVICE memory (cu:39)
 *   level_start_index  (L,)   int64                             DEVICE memory (cu:40)
 *   sampling_loc       (N, Lq, M, L, P, 2)  (x, y) in [0,1]     loc_dtype
 *   attn_weight        (N, Lq, M, L, P)                         loc_dtype
 *   output/grad_output (N, Lq, M*D)         value_dtype
 * with loc_dtype = float32 for PCT_MSDA_F32 and PCT_MSDA_BF16, float64 for PCT_MSDA_F64
 * (the reference dispatches fp32/fp64 only, cu:69; bf16 values with fp32 locations/weights and
 * fp32 accumulation is this library's extension).
 *
 * Ownership: the caller allocates every buffer; the library never allocates, frees or retains
 * device memory.  Outputs are write-only: they need not be initialised (the reference zero-fills
 * with at::zeros, cu:59,126-128; here grad_value is cleared by the callee on `stream`).
 *
 * Threading / streams: every call is asynchronous on `cuda_stream` (a cudaStream_t cast to
 * void*; NULL = legacy default stream) of the CURRENT device, never synchronises the host, is
 * re-entrant and may be called from any host thread (cu:70,140 use the current stream too).
 *
 * The persistent kernels balance their tail through a ring of 512 module-scope counter slots, one
 * per launch (the slot index is chosen on the host at launch time and is baked into a captured
 * CUDA graph).  Eager launches may overlap freely.  A captured graph must not run concurrently
 * with itself, nor with more than ~500 other pct_msda launches issued after its capture (the ring
 * would wrap onto its slot): serialise graph replays of this library on one stream, or re-capture.
 *
 * Numerics: fp32 arithmetic with fp32 accumulation for F32 and BF16 values, fp64 for F64.  grad_value
 * is accumulated with fp32 global reductions (red.global.add.v4.f32; the reference uses atomicAdd,
 * cuh:130-157): the summation order varies from run to run unless PCT_MSDA_FLAG_DETERMINISTIC is
 * set, and -- a property of the hardware instruction, REDG.E.ADD.F32x4.FTZ -- subnormal addends and
 * sums are flushed to zero (|x| < 1.18e-38; the reference's atomicAdd flushes them too).
 *
 * Errors: 0 on success; 1..999 = the cudaError_t of the failing runtime call or launch
 * (the reference only printf()s launch errors, cuh:953-957); >= 1000 = PCT_MSDA_E_* below.
 * There is no CPU fallback of any kind.
 */
#ifndef PCT_MSDA_H_
#define PCT_MSDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCT_MSDA_ABI_VERSION 4

/* value_dtype */
#define PCT_MSDA_F32 0
#define PCT_MSDA_BF16 1
#define PCT_MSDA_F64 2

/* own error codes */
#define PCT_MSDA_E_NULL_POINTER 1001
#define PCT_MSDA_E_BAD_SHAPE 1002
#define PCT_MSDA_E_BAD_DTYPE 1003
#define PCT_MSDA_E_MISALIGNED 1004
#define PCT_MSDA_E_WORKSPACE 1005
#define PCT_MSDA_E_TOO_LARGE 1006
#define PCT_MSDA_E_UNSUPPORTED 1007

/* flags for the *_ex entry points (0 = library default; used by the tuning harness) */
#define PCT_MSDA_FLAG_FORCE_GENERIC 0x1u /* use the any-D scalar kernels even when a vector path fits */
#define PCT_MSDA_FLAG_FWD_STAGED 0x4u    /* forward, fp32 values, 32-channel heads, power-of-two head count: stage the
                                            smallest levels of a head in shared memory (an experiment kept for
                                            measurements; same results) */
#define PCT_MSDA_FLAG_BF16_ACCUM 0x8u    /* backward, bf16 values: accumulate grad_value directly in bf16
                                            (packed bf16x2 reductions, no workspace; every tap rounds) */
#define PCT_MSDA_FLAG_BF16_F16_ACCUM 0x80u /* backward, bf16 values: accumulate grad_value in a scaled fp16
                                            image inside the workspace instead of the fp32 scratch image:
                                            half the L2 reduction traffic (6-20 % faster), about twice
                                            the error (4-7e-3 against 2-3e-3 of max|grad_value|).  The
                                            scale (from max|grad_out|, max sum|attn| and Lq) rules out
                                            overflow, but an element that receives hundreds of EQUAL
                                            same-signed contributions is off by percents (biased
                                            rounding) and beyond ~2000 it stops growing (11-bit
                                            running sum) -- hence opt-in */
#define PCT_MSDA_FLAG_DETERMINISTIC 0x10u /* backward: bit-reproducible grad_value (64-bit fixed-point
                                            integer reductions; needs pct_msda_backward_workspace_bytes_ex
                                            bytes of workspace; vector-path shapes only, else
                                            PCT_MSDA_E_UNSUPPORTED) */
#define PCT_MSDA_FLAG_BWD_PER_TAP 0x20u   /* backward: use the per-tap scatter kernel (one L2 reduction per
                                            bilinear corner) even where the cell-binned kernel is the default */
#define PCT_MSDA_FLAG_BWD_BINNED 0x40u    /* backward: use the cell-binned kernel whenever the shape fits it
                                            (since round 2 it is the default there; the flag is kept
                                            for callers that set it) */
/* bits 8..15: threads/32 of the persistent CTA; bits 16..19: CTAs per SM; bits 20..24: 1 + static
 * share of the task list in sixteenths; bits 25..31: cap of the binned backward's chunk size in
 * units of 8 queries (tuning only) */

int pct_msda_abi_version(void);

/* Human-readable text for a code returned by any entry point (static storage). */
const char* pct_msda_error_string(int code);

/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
uint64_t pct_msda_launch_count(void);

/* Which kernel family a call with these dimensions runs: 0 = generic scalar, 1 = 128-bit vector
 * path.  `backward` selects the gradient kernels. */
int pct_msda_kernel_family(int channels, int num_levels, int num_point, int value_dtype,
                           int backward);

/* Which gradient kernel pct_msda_backward_ex runs for these dimensions and flags on the current
 * device (buffers assumed 16-byte aligned): 0 = generic scalar, 1 = per-tap vector scatter,
 * 2 = cell-binned (grad_value aggregated per bilinear cell on chip before any global reduction;
 * available when the head dim is 16/32/64/128 channels, L <= 8, L*P <= 64 and the level maps are
 * small enough for shared memory -- 12 bytes per pixel of one image beside a 64-query chunk, about
 * 18 000 pixels on B200, or 6 bytes per bilinear cell when the level shapes are passed on the host,
 * see pct_msda_backward_plan_ex; the default wherever it fits, PCT_MSDA_FLAG_BWD_PER_TAP overrides).
 * Negative = error code negated. */
int pct_msda_backward_plan(int batch, int spatial_size, int num_heads, int channels,
                           int num_levels, int num_query, int num_point, int value_dtype,
                           unsigned flags);

/* The level shapes live in DEVICE memory, so for the cell-binned kernel the library sizes shared
 * memory from a bound that holds for any shapes (2S + 2L cells, about twice the count of square
 * maps).  A caller that also has spatial_shapes on the HOST can pass that copy ((L, 2) int64 rows
 * (H_l, W_l), must describe the same S pixels, else PCT_MSDA_E_BAD_SHAPE) to the _ex2 / _plan_ex
 * entry points below: the exact cell count then decides, which extends the binned kernel from
 * ~18 000 to ~30 000 pixels per image (e.g. a 1024 x 1024 input at strides 8/16/32).  NULL = the
 * plain entry points' behaviour.  The host copy is read during the call only. */
int pct_msda_backward_plan_ex(int batch, int spatial_size, int num_heads, int channels,
                              int num_levels, int num_query, int num_point, int value_dtype,
                              unsigned flags, const int64_t* host_spatial_shapes);

/* out[b,q,m,:] = sum_{l,p} attn[b,q,m,l,p] * bilinear(value_l[b,:,m,:], loc[b,q,m,l,p]).
 * Replaces ms_deform_attn_cuda_forward (cu:25-85).  The reference's `im2col_step` has no
 * counterpart: the whole batch is processed by one launch with 64-bit offsets. */
int pct_msda_forward(const void* value, const int64_t* spatial_shapes,
                     const int64_t* level_start_index, const void* sampling_loc,
                     const void* attn_weight, void* output, int batch, int spatial_size,
                     int num_heads, int channels, int num_levels, int num_query, int num_point,
                     int value_dtype, void* cuda_stream);

int pct_msda_forward_ex(const void* value, const int64_t* spatial_shapes,
                        const int64_t* level_start_index, const void* sampling_loc,
                        const void* attn_weight, void* output, int batch, int spatial_size,
                        int num_heads, int channels, int num_levels, int num_query,
                        int num_point, int value_dtype, void* cuda_stream, unsigned flags);

/* Bytes of device scratch pct_msda_backward needs for these dimensions (0 for fp32/fp64: only the
 * bf16 path accumulates grad_value in an fp32 scratch image before rounding it once). */
size_t pct_msda_backward_workspace_bytes(int batch, int spatial_size, int num_heads, int channels,
                                         int value_dtype);

/* Workspace for pct_msda_backward_ex with `flags` (PCT_MSDA_FLAG_DETERMINISTIC changes it). */
size_t pct_msda_backward_workspace_bytes_ex(int batch, int spatial_size, int num_heads,
                                            int channels, int value_dtype, unsigned flags);

/* Gradients of pct_msda_forward w.r.t. value, sampling_loc and attn_weight.
 * Replaces ms_deform_attn_cuda_backward (cu:88-158).  grad_value (N,S,M,D) value_dtype,
 * grad_sampling_loc / grad_attn_weight shaped like their inputs, loc_dtype; all three are fully
 * overwritten.  `workspace` may be NULL when workspace_bytes is 0. */
int pct_msda_backward(const void* value, const int64_t* spatial_shapes,
                      const int64_t* level_start_index, const void* sampling_loc,
                      const void* attn_weight, const void* grad_output, void* grad_value,
                      void* grad_sampling_loc, void* grad_attn_weight, int batch,
                      int spatial_size, int num_heads, int channels, int num_levels,
                      int num_query, int num_point, int value_dtype, void* workspace,
                      size_t workspace_bytes, void* cuda_stream);

int pct_msda_backward_ex(const void* value, const int64_t* spatial_shapes,
                         const int64_t* level_start_index, const void* sampling_loc,
                         const void* attn_weight, const void* grad_output, void* grad_value,
                         void* grad_sampling_loc, void* grad_attn_weight, int batch,
                         int spatial_size, int num_heads, int channels, int num_levels,
                         int num_query, int num_point, int value_dtype, void* workspace,
                         size_t workspace_bytes, void* cuda_stream, unsigned flags);

/* The same with an optional host copy of spatial_shapes (see pct_msda_backward_plan_ex). */
int pct_msda_backward_ex2(const void* value, const int64_t* spatial_shapes,
                          const int64_t* level_start_index, const void* sampling_loc,
                          const void* attn_weight, const void* grad_output, void* grad_value,
                          void* grad_sampling_loc, void* grad_attn_weight, int batch,
                          int spatial_size, int num_heads, int channels, int num_levels,
                          int num_query, int num_point, int value_dtype, void* workspace,
                          size_t workspace_bytes, void* cuda_stream, unsigned flags,
                          const int64_t* host_spatial_shapes);

/*
 * Fused prologue: the same op, fed with what the module's Linears
 * produce instead of materialised locations / weights.  Replaces, together with the kernels,
 * OPS/modules/ms_deform_attn.py:104 (softmax over L*P) and :106-112 (location arithmetic):
 *   reference_points (N, Lq, L, ref_dim) fp32, ref_dim = 2 (points) or 4 (boxes cx,cy,w,h)
 *   sampling_offsets (N, Lq, M, L, P, 2) fp32 raw     attn_logits (N, Lq, M, L*P) fp32 raw
 *   loc = ref + off / (W_l, H_l)              (ref_dim 2)
 *   loc = ref.xy + off / P * ref.wh * 0.5     (ref_dim 4)        attn = softmax(logits)
 * value_dtype F32 or BF16.  Only shapes served by the vector kernels are accepted
 * (pct_msda_kernel_family == 1, S > 0); otherwise PCT_MSDA_E_UNSUPPORTED is returned and the
 * caller composes softmax / locations itself and calls pct_msda_forward.
 * The backward writes d/d(sampling_offsets) and d/d(attn_logits); reference points get no
 * gradient from these entry points.
 */
int pct_msda_fused_forward(const void* value, const int64_t* spatial_shapes,
                           const int64_t* level_start_index, const void* reference_points,
                           int ref_dim, const void* sampling_offsets, const void* attn_logits,
                           void* output, int batch, int spatial_size, int num_heads, int channels,
                           int num_levels, int num_query, int num_point, int value_dtype,
                           void* cuda_stream);

int pct_msda_fused_backward(const void* value, const int64_t* spatial_shapes,
                            const int64_t* level_start_index, const void* reference_points,
                            int ref_dim, const void* sampling_offsets, const void* attn_logits,
                            const void* grad_output, void* grad_value, void* grad_offsets,
                            void* grad_logits, int batch, int spatial_size, int num_heads,
                            int channels, int num_levels, int num_query, int num_point,
                            int value_dtype, void* workspace, size_t workspace_bytes,
                            void* cuda_stream);

/* The same with option bits: PCT_MSDA_FLAG_BWD_PER_TAP / PCT_MSDA_FLAG_BWD_BINNED choose between the
 * fused per-tap and the fused cell-binned gradient kernel (default: cell-binned wherever it fits,
 * as for pct_msda_backward_ex), PCT_MSDA_FLAG_BF16_F16_ACCUM as above; any other option bit returns
 * PCT_MSDA_E_UNSUPPORTED. */
int pct_msda_fused_backward_ex(const void* value, const int64_t* spatial_shapes,
                               const int64_t* level_start_index, const void* reference_points,
                               int ref_dim, const void* sampling_offsets, const void* attn_logits,
                               const void* grad_output, void* grad_value, void* grad_offsets,
                               void* grad_logits, int batch, int spatial_size, int num_heads,
                               int channels, int num_levels, int num_query, int num_point,
                               int value_dtype, void* workspace, size_t workspace_bytes,
                               void* cuda_stream, unsigned flags);

/* The same with an optional host copy of spatial_shapes (see pct_msda_backward_plan_ex). */
int pct_msda_fused_backward_ex2(const void* value, const int64_t* spatial_shapes,
                                const int64_t* level_start_index, const void* reference_points,
                                int ref_dim, const void* sampling_offsets, const void* attn_logits,
                                const void* grad_output, void* grad_value, void* grad_offsets,
                                void* grad_logits, int batch, int spatial_size, int num_heads,
                                int channels, int num_levels, int num_query, int num_point,
                                int value_dtype, void* workspace, size_t workspace_bytes,
                                void* cuda_stream, unsigned flags,
                                const int64_t* host_spatial_shapes);

/*
 * Host pipeline: the same op for callers whose tensors live in HOST memory (the end-to-end path
 * bench.py reports as `e2e`).  A context owns three CUDA streams and device staging buffers that
 * grow on demand; a call cuts the batch into chunks of images (the op is per-image independent),
 * overlaps H2D copies, kernels and D2H copies, and returns when the results are in host memory.
 * Pinned (page-locked) host buffers give full-duplex copies; pageable buffers work but serialise.
 * spatial_shapes / level_start_index are HOST arrays here.  One call at a time per context;
 * contexts are independent (one per device / thread).  Every entry point below works on the
 * context's device and restores the calling thread's current device before it returns; when a
 * call fails, its streams are drained first, so the caller's host buffers are no longer in use,
 * and a failed growth of the staging buffers leaves the context empty but reusable.
 */
typedef struct pct_msda_host_ctx pct_msda_host_ctx;

int pct_msda_host_ctx_create(pct_msda_host_ctx** ctx, int device, int chunk_images /* 0 = batch/16 */);
void pct_msda_host_ctx_destroy(pct_msda_host_ctx* ctx);

int pct_msda_host_forward(pct_msda_host_ctx* ctx, const void* value, const int64_t* spatial_shapes,
                          const int64_t* level_start_index, const void* sampling_loc,
                          const void* attn_weight, void* output, int batch, int spatial_size,
                          int num_heads, int channels, int num_levels, int num_query,
                          int num_point, int value_dtype);

int pct_msda_host_forward_backward(pct_msda_host_ctx* ctx, const void* value,
                                   const int64_t* spatial_shapes, const int64_t* level_start_index,
                                   const void* sampling_loc, const void* attn_weight,
                                   const void* grad_output, void* output, void* grad_value,
                                   void* grad_sampling_loc, void* grad_attn_weight, int batch,
                                   int spatial_size, int num_heads, int channels, int num_levels,
                                   int num_query, int num_point, int value_dtype);

#ifdef __cplusplus
}
#endif
#endif /* PCT_MSDA_H_ */
