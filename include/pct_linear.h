/*
 * pct_linear.h -- C ABI of the tensor-core Linear used around the MSDeformAttn op
 * (the encoder layer's projections and their epilogues).
 *
 * Replaces, for fp32 activations on B200, the cuBLAS SIMT-fp32 GEMMs behind the reference's
 *   value_proj / sampling_offsets / attention_weights / output_proj   OPS/modules/ms_deform_attn.py:55-58
 *   linear1 / linear2 (FFN) and the residual + LayerNorm after them   PD/msdeformattn.py:107-120,125-126
 * which the reference runs in fp32 with autocast disabled (PD/msdeformattn.py:314-320).
 *
 *   y = epilogue(x @ W^T + bias)        x (rows, in) fp32, W (out, in) fp32, y (rows, out) fp32
 *
 * Arithmetic: tcgen05 tensor cores, kind::tf32, with the 3xTF32 split
 *   x = x_hi + x_lo,  W = W_hi + W_lo,  x@W^T ~= x_hi@W_hi^T + x_hi@W_lo^T + x_lo@W_hi^T
 * (the dropped x_lo@W_lo^T term is ~2^-22 relative), fp32 accumulation in tensor memory.  The
 * result matches an fp32 GEMM to a few 1e-7 relative, so the encoder keeps the reference's fp32
 * parity bar while the GEMMs run on the tensor cores.  The activation is split inside the kernel;
 * the weight is split and re-tiled once by pct_linear_pack_weight into the order the kernel's
 * shared-memory stages use, so a stage is one bulk copy (cache the packed copy per weight version).
 *
 * Same conventions as pct_msda.h: device pointers, caller-owned buffers, asynchronous on
 * `cuda_stream`, 0 or an error code (pct_msda_error_string understands them), no CPU fallback.
 */
#ifndef PCT_LINEAR_H_
#define PCT_LINEAR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* epilogue flags */
#define PCT_LINEAR_RELU 0x1u      /* y = max(., 0)                                        */
#define PCT_LINEAR_RESIDUAL 0x2u  /* y = residual + (x W^T + b); residual (rows, out)      */
#define PCT_LINEAR_GATE 0x8u      /* y = (x W^T + b) where gate > 0, else 0; the (rows, out) gate is passed
                                     in `residual` (ReLU backward fused into the input-gradient GEMM);
                                     not combinable with RESIDUAL / LAYERNORM */
#define PCT_LINEAR_LAYERNORM 0x4u /* y = LayerNorm(. ) * gamma + beta over the out dim;
                                     needs out_features <= 256 (one tile holds a whole row) */
#define PCT_LINEAR_ROWMASK 0x10u  /* rows whose byte in the (rows,) uint8 / bool mask passed in `residual` is
                                     non-zero are written as zeros: the masked_fill of padded positions
                                     after value_proj (OPS/modules/ms_deform_attn.py:100-101) in the
                                     epilogue; not combinable with RESIDUAL / GATE / LAYERNORM */

#define PCT_LINEAR_BF16 0x20u     /* operand precision: x and W are rounded to bf16 (nearest even) on their
                                     way into the tensor cores, products accumulate in fp32; tensors in
                                     global memory stay fp32 and every epilogue above applies.  One
                                     tcgen05.mma.kind::f16 instead of three kind::tf32 per K step: what
                                     torch.autocast(bfloat16) computes for a Linear (without rounding
                                     the result), for callers that trade the reference's fp32 accuracy
                                     (PD/msdeformattn.py:314-320 disables autocast) for speed.
                                     `w_packed` must come from pct_linear_pack_weight_bf16;
                                     in_features % 32 == 0 (forward); also accepted by pct_linear_wgrad_ex. */

/* bf16 tensors in global memory (only together with PCT_LINEAR_BF16; the pointer keeps its
 * `float*` type in the signature and is read / written as bf16).  Rounding an fp32 tensor to bf16 is
 * what the bf16 operand path does on load anyway, so storing an intermediate that is only consumed
 * as a GEMM operand or as a gate in bf16 changes no result -- it halves that tensor's traffic.  Used
 * for the FFN's hidden activation and its gradient (PD/msdeformattn.py:107-120). */
#define PCT_LINEAR_X_BF16 0x40u    /* x (rows, in) is bf16                  forward, wgrad_ex */
#define PCT_LINEAR_Y_BF16 0x80u    /* y (rows, out) is written as bf16      forward (pre_norm stays fp32) */
#define PCT_LINEAR_RES_BF16 0x100u /* the residual / gate tensor is bf16    forward */
#define PCT_LINEAR_GY_BF16 0x200u  /* grad_y (rows, out) is bf16            wgrad_ex */

/* Constraints everywhere: in_features % 16 == 0, out_features % 16 == 0; x, packed weight, residual,
 * y, pre_norm, grad_y, grad_w, grad_bias 16-byte aligned (bias / LayerNorm vectors: 4-byte). */

/* Bytes of the packed weight (0 if the dimensions are not supported). */
size_t pct_linear_packed_bytes(int in_features, int out_features);

/* w (out, in) fp32 -> packed: per 256-column tile and 16-float K step, the tf32 head of the tile
 * followed by the fp32 remainder (w - head), each in tensor-core core-matrix order. */
int pct_linear_pack_weight(const float* w, void* packed, int in_features, int out_features,
                           void* cuda_stream);

/* The same for PCT_LINEAR_BF16: w (out, in) fp32 -> packed bf16, per 256-column tile and 32-element K
 * step one piece in tensor-core core-matrix order.  in_features % 32 == 0. */
size_t pct_linear_packed_bytes_bf16(int in_features, int out_features);
int pct_linear_pack_weight_bf16(const float* w, void* packed, int in_features, int out_features,
                                void* cuda_stream);

/* y (rows, out) = epilogue(x (rows, in) @ W^T + bias).  bias may be NULL.  residual / ln_gamma /
 * ln_beta are read only when their flag is set.  rows >= 1. */
int pct_linear_forward(const float* x, const void* w_packed, const float* bias,
                       const float* residual, const float* ln_gamma, const float* ln_beta,
                       float ln_eps, float* y, int rows, int in_features, int out_features,
                       unsigned flags, void* cuda_stream);

/* Same, for training with PCT_LINEAR_LAYERNORM: also writes what LayerNorm's backward needs --
 * pre_norm (rows, out) = x W^T + b (+ residual), ln_mean / ln_rstd (rows,) -- each may be NULL
 * (mean and rstd together). */
int pct_linear_forward_train(const float* x, const void* w_packed, const float* bias,
                             const float* residual, const float* ln_gamma, const float* ln_beta,
                             float ln_eps, float* y, float* pre_norm, float* ln_mean,
                             float* ln_rstd, int rows, int in_features, int out_features,
                             unsigned flags, void* cuda_stream);

/* Two Linears that read the same input as ONE GEMM, without a concatenated copy of anything
 * (sampling_offsets and attention_weights both project the query, OPS/modules/ms_deform_attn.py:102-104):
 *   forward:         x_b = NULL, in_b = 0;   [y_a (rows, out_a) | y_b (rows, out_b)] = x_a @ W^T + bias
 *                    with W (out_a + out_b, in_a) the row-concatenation of the two weights (packed as one)
 *                    and bias the concatenation of the two biases; out_a % 32 == 0;
 *   input gradient:  y_b = NULL, out_b = 0;  y_a (rows, out_a) = [x_a (rows, in_a) | x_b (rows, in_b)] @ W^T
 *                    with W the packed transpose of that concatenated weight, x_a / x_b the two output
 *                    gradients; in_a a multiple of the K step (16, or 32 with PCT_LINEAR_BF16).
 * flags: PCT_LINEAR_RELU, PCT_LINEAR_RESIDUAL (single y only), PCT_LINEAR_BF16.  fp32 tensors. */
int pct_linear_forward_split(const float* x_a, const float* x_b, int in_a, int in_b,
                             const void* w_packed, const float* bias, const float* residual,
                             float* y_a, float* y_b, int out_a, int out_b, int rows, unsigned flags,
                             void* cuda_stream);

/* grad_w (out, in) = grad_y (rows, out)^T @ x (rows, in): the weight gradient of the Linear, same
 * 3xTF32 arithmetic, the row dimension split over CTAs and combined with fp32 reductions (so, like
 * the reference's atomicAdd backward, not bit-reproducible run to run).  grad_w is fully
 * overwritten (cleared by the callee on `cuda_stream`).  If grad_bias (out,) is not NULL it receives
 * the bias gradient, the column sums of grad_y, from the same pass over grad_y. */
int pct_linear_wgrad(const float* grad_y, const float* x, float* grad_w, float* grad_bias, int rows,
                     int in_features, int out_features, void* cuda_stream);

/* The same with option flags: PCT_LINEAR_BF16 rounds both operands (grad_y, x) to bf16 on their way
 * into the tensor cores (fp32 accumulation, fp32 grad_w / grad_bias; the bias gradient is summed
 * from the unrounded grad_y) -- what torch.autocast(bfloat16) computes for a Linear's weight
 * gradient.  flags = 0 is pct_linear_wgrad. */
int pct_linear_wgrad_ex(const float* grad_y, const float* x, float* grad_w, float* grad_bias, int rows,
                        int in_features, int out_features, unsigned flags, void* cuda_stream);

/* Weight gradient of two Linears that read the same x (see pct_linear_forward_split): grad_y is the
 * pair [grad_y_a (rows, out_a) | grad_y_b (rows, out_b)], grad_w (out_a + out_b, in) and grad_bias
 * (out_a + out_b,) are the row-concatenations of the two gradients.  out_a % 4 == 0, out_b % 4 == 0;
 * flags: PCT_LINEAR_BF16, PCT_LINEAR_X_BF16.  out_b = 0 is pct_linear_wgrad_ex. */
int pct_linear_wgrad_split(const float* grad_y_a, const float* grad_y_b, int out_a, int out_b,
                           const float* x, float* grad_w, float* grad_bias, int rows, int in_features,
                           unsigned flags, void* cuda_stream);

/* Backward of the LayerNorm fused into pct_linear_forward_train (PD/msdeformattn.py:107-126: the two
 * `norm(residual + linear(x))` blocks of an encoder layer), one pass over the rows:
 *   x_hat = (pre_norm - mean) * rstd,   g = grad_y * gamma
 *   grad_pre   = rstd * (g - mean_n(g) - x_hat * mean_n(g * x_hat))         (rows, features)
 *   grad_gamma = sum_rows grad_y * x_hat,   grad_beta = sum_rows grad_y     (features,)
 * features in {128, 256, 384, 512}; grad_y / pre_norm / grad_pre / gamma 16-byte aligned.  grad_gamma
 * and grad_beta are fully overwritten (cleared by the callee on `cuda_stream`, combined with fp32
 * reductions: like pct_linear_wgrad not bit-reproducible run to run).  grad_pre is also the gradient
 * of the residual input and the operand of the Linear's dgrad / wgrad. */
int pct_linear_layernorm_backward(const float* grad_y, const float* pre_norm, const float* mean,
                           const float* rstd, const float* gamma, float* grad_pre,
                           float* grad_gamma, float* grad_beta, int rows, int features,
                           void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* PCT_LINEAR_H_ */
