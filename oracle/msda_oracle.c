/*
 * msda_oracle.c -- CPU restatement of the reference's multi-scale deformable attention.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pctrans_b200/ may import, link or execute this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and there only as the checker / reported baseline.
 *
 * What it restates (paths relative to /root/reference, OPS = connectomics/model/
 * maskformer_block/pixel_decoder/ops):
 *   - forward closed form : OPS/src/cuda/ms_deform_im2col_cuda.cuh:38-89 (bilinear with
 *     per-corner zero padding) and :258-303 (loop over levels and points, the in-range
 *     predicate h_im > -1 && w_im > -1 && h_im < H && w_im < W at :293, pixel coordinates
 *     h_im = loc_y*H - 0.5, w_im = loc_x*W - 0.5 at :290-291).
 *   - backward closed form: OPS/src/cuda/ms_deform_im2col_cuda.cuh:93-164 (grad_value
 *     scatter, grad_h_weight / grad_w_weight, grad_attn = top_grad*val,
 *     grad_loc = (W*grad_w_weight*top_grad*attn, H*grad_h_weight*top_grad*attn)).
 *   - results are the same function as OPS/functions/ms_deform_attn_func.py:52-72
 *     (grid_sample bilinear / zeros / align_corners=False), which is the designated oracle;
 *     tests/test_oracle.py pins this file against golden vectors generated from that
 *     function by tests/golden/make_golden.py.
 *
 * Parity status: PINNED (golden vectors from the reference's own Python oracle + the
 * OPS/test.py recipe, see tests/golden/).
 *
 * Layouts (all row-major, identical to the reference):
 *   value (N,S,M,D)  shapes (L,2) int64 rows (H,W)  lsi (L) int64
 *   loc (N,Lq,M,L,P,2) last dim (x,y) in [0,1]   attn (N,Lq,M,L,P)
 *   out / grad_out (N,Lq,M*D)
 *
 * The file is compiled twice through REAL: double (msda_oracle_*_f64) and float
 * (msda_oracle_*_f32, literal fp32 arithmetic in the reference's operation order;
 * build with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define MSDA_ORACLE_OK 0
#define MSDA_ORACLE_BAD_ARG 1

int msda_oracle_abi_version(void) { return 1; }

#define CONCAT_(a, b) a##b
#define CONCAT(a, b) CONCAT_(a, b)

#define REAL double
#define SUFFIX _f64
#define FLOOR floor
#include "msda_oracle_impl.inc"
#undef REAL
#undef SUFFIX
#undef FLOOR

#define REAL float
#define SUFFIX _f32
#define FLOOR floorf
#include "msda_oracle_impl.inc"
#undef REAL
#undef SUFFIX
#undef FLOOR
