// msda_capi.cu -- C ABI (include/pct_msda.h) and launch logic for the MSDeformAttn kernels.
//
// Host side of the drop-in boundary: argument validation (the reference's AT_ASSERTM block,
// OPS/src/cuda/ms_deform_attn_cuda.cu:33-57,98-124, becomes error codes), kernel selection and
// asynchronous launches on the caller's stream.  Links against the CUDA runtime only.
#include "../../include/pct_msda.h"

#include <atomic>

#include "msda_bwd_binned.cuh"
#include "msda_kernels.cuh"
#include "msda_launch.h"

namespace {

std::atomic<uint64_t> g_launches{0};
std::atomic<unsigned> g_sched_seq{0};   // ring position of the next persistent launch

// Static share of the task list (in 1/16) before the dynamic tail; tuning flag bits 20..24 carry
// (value + 1) to override.
constexpr int kStaticSixteenths = 14;

}  // namespace

namespace pct {
// launch accounting shared with linear_tc.cu
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace pct

namespace pctl {

void count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
unsigned next_sched_slot() {
  return g_sched_seq.fetch_add(1, std::memory_order_relaxed) % pct::kSchedSlots;
}

pct::SchedArgs make_sched(long long n_tasks, int grid, int warps, unsigned flags) {
  int sixteenths = kStaticSixteenths;
  const int f = (int)((flags >> 20) & 0x1fu);
  if (f > 0) sixteenths = f - 1;
  long long per_cta = n_tasks * sixteenths / 16 / grid;
  per_cta -= per_cta % warps;             // whole rounds of the CTA's warps
  pct::SchedArgs sa;
  sa.slot = (int)next_sched_slot();
  sa.static_per_cta = (int)(per_cta < 0 ? 0 : per_cta);
  return sa;
}

// Persistent launch: one CTA per SM (or fewer when there is little work).
LaunchShape vec_launch_shape(long long n_tasks, int sm_count, unsigned flags, int default_block) {
  int block = default_block;
  int ctas_per_sm = 1;
  const int fb = (int)((flags >> 8) & 0xffu);   // threads / 32
  const int fc = (int)((flags >> 16) & 0xfu);
  if (fb > 0 && fb * 32 <= default_block) block = fb * 32;
  if (fc > 0) ctas_per_sm = fc;
  const int warps = block / 32;
  long long want = (n_tasks + warps - 1) / warps;
  long long cap = (long long)sm_count * ctas_per_sm;
  LaunchShape s;
  s.grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
  s.block = block;
  return s;
}

}  // namespace pctl

namespace {

using namespace pctl;

int device_sm_count(int* out) {
  // Cached per device ordinal; racing initialisers write the same value.
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev >= 0 && dev < 64) {
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v > 0) { *out = v; return 0; }
  }
  int n = 0;
  e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
  *out = n;
  return 0;
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

inline size_t elem_size(int dtype) {
  return dtype == PCT_MSDA_F32 ? 4 : dtype == PCT_MSDA_BF16 ? 2 : 8;
}

inline bool pow2_le32(size_t x) { return x >= 1 && x <= 32 && (x & (x - 1)) == 0; }

// Forward prefers 16-byte lanes (LDG.128); backward wants 4 channels per lane so that every
// red.v4.f32 covers whole sectors together with its neighbours (16 B fp32, 8 B bf16).  Wider lanes
// are the fall-back for head dims that would need more than 32 lanes.
inline VecPlan vec_plan(int channels, int num_levels, int num_point, int dtype, bool backward) {
  VecPlan p;
  if (dtype != PCT_MSDA_F32 && dtype != PCT_MSDA_BF16) return p;
  if ((long long)num_levels * num_point > pct::kMaxFastSamples) return p;
  const size_t bytes = (size_t)channels * elem_size(dtype);
  const int first = (dtype == PCT_MSDA_BF16 && backward) ? 8 : 16;
  const int second = first * 2;
  if (bytes % first == 0 && pow2_le32(bytes / first)) {
    p.lanes = (int)(bytes / first);
    p.vb = first;
  } else if (bytes % second == 0 && bytes / second == 32) {
    p.lanes = 32;
    p.vb = second;
  } else if (dtype == PCT_MSDA_BF16 && !backward && bytes == 8) {
    p.lanes = 1;
    p.vb = 8;
  }
  return p;
}

int check_dims(const Dims& d, int dtype) {
  if (dtype != PCT_MSDA_F32 && dtype != PCT_MSDA_BF16 && dtype != PCT_MSDA_F64) return PCT_MSDA_E_BAD_DTYPE;
  if (d.N < 0 || d.S < 0 || d.Lq < 0 || d.M <= 0 || d.D <= 0 || d.L <= 0 || d.P <= 0) return PCT_MSDA_E_BAD_SHAPE;
  // per-image element counts are indexed with 64-bit offsets, but keep every count representable
  const long double lim = 9.0e18L;
  if ((long double)d.N * d.S * d.M * d.D > lim || (long double)d.N * d.Lq * d.M * d.L * d.P * 2 > lim)
    return PCT_MSDA_E_TOO_LARGE;
  return 0;
}

#ifndef PCT_ZERO_CTAS_PER_SM
#define PCT_ZERO_CTAS_PER_SM 4
#endif
// zero `bytes` at a 16-byte aligned address with a kernel (see msda_zero_kernel)
int zero_fill(void* p, size_t bytes, int sm_count, cudaStream_t st) {
  if (bytes == 0) return 0;
  if (!aligned(p, 16)) return (int)cudaMemsetAsync(p, 0, bytes, st);
  const long long n16 = (long long)(bytes / 16);
  const int n_tail = (int)(bytes % 16);
  long long want = (n16 + 1023) / 1024;   // 4 stores per thread per pass
  // 4 CTAs = 1024 threads per SM: enough stores in flight for the HBM rate, and room for the
  // 1024-thread CTA of the cell-binned backward to become resident beside them (see msda_zero_kernel)
  long long cap = (long long)sm_count * PCT_ZERO_CTAS_PER_SM;
  const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
  pct::msda_zero_kernel<<<grid, 256, 0, st>>>((uint4*)p, n16, (unsigned char*)p + n16 * 16, n_tail);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int round_workspace_to_bf16(const float* ws, void* grad_value, long long n, int sm_count,
                            cudaStream_t st) {
  const long long n8 = n / 8;
  if (n8 > 0) {
    long long want = (n8 + 255) / 256;
    long long cap = (long long)sm_count * 16;
    pct::msda_round_to_bf16_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(
        ws, (__nv_bfloat16*)grad_value, n8);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  if (n8 * 8 < n) {
    pct::msda_round_to_bf16_tail_kernel<<<1, 32, 0, st>>>(ws, (__nv_bfloat16*)grad_value, n8 * 8, n);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Cell-binned backward (msda_bwd_binned.cuh): the default gradient kernel whenever it fits.
// The level shapes live in device memory, so the host sizes shared memory from an upper bound of
// the number of bilinear cells,  sum_l (H_l+1)(W_l+1) <= 2*S + 2*L  (H + W <= H*W + 1), and the
// kernel derives the chunk size from the real shapes.
// ---------------------------------------------------------------------------------------------
int device_smem_optin(int* out) {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev >= 0 && dev < 64) {
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v > 0) { *out = v; return 0; }
  }
  int n = 0;
  e = cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return (int)e;
  if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
  *out = n;
  return 0;
}

// 0 = not eligible, else lanes per (query, head) row
// `cells` = sum_l (H_l+1)(W_l+1) when the caller supplied the level shapes on the host, else 0: the
// bound 2S + 2L (H + W <= H*W + 1) stands in, which is about twice the count of square maps.
long long cells_bound(const Dims& d, long long cells) { return cells > 0 ? cells : 2ll * d.S + 2ll * d.L; }

int binned_plan(const Dims& d, int dtype, unsigned flags, int smem_optin, const void* value,
                const void* grad_out, const void* acc, const void* loc, const void* attn,
                const void* grad_loc, bool fused = false, long long cells = 0) {
  if (flags & (PCT_MSDA_FLAG_FORCE_GENERIC | PCT_MSDA_FLAG_BWD_PER_TAP)) return 0;
  // Measured on B200 (profiles/r02_binned_v3_*.log): in its round-2 form (bilinear-basis sweep in
  // lock-step rounds) the binned kernel is faster than the per-tap kernel on every shape it fits
  // -- 64-, 128- and 256-byte head rows, fp32 and bf16, PCTrans's local sampling pattern and
  // uniform-random locations, batch 1 to 32 -- so it is the default; PCT_MSDA_FLAG_BWD_PER_TAP
  // selects the other one.
  if (dtype != PCT_MSDA_F32 && dtype != PCT_MSDA_BF16) return 0;
  if (d.D % 4 != 0) return 0;
  const int lanes = d.D / 4;
  if (lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32) return 0;
  const long long LP = (long long)d.L * d.P;
  if (d.L > pct::kBinMaxLevels || LP > pct::kBinMaxSamples) return 0;
  if ((long long)d.N * d.S * d.M * d.D >= (1ll << 31) || d.S == 0) return 0;   // 32-bit element offsets
  if ((long long)d.N * d.M * ((long long)d.Lq / 8 + 1) >= (1ll << 31)) return 0;   // task index
  const long long nc_bound = cells_bound(d, cells);
  if (nc_bound > pct::kBinMaxCells || d.S > 32767) return 0;   // 16-bit cell ids, 15-bit pixel indices
  const int row_bytes = d.D * (int)elem_size(dtype);
  const long long fixed = pct::bin_bytes_fixed(nc_bound) + 1536;   // + the kernel's static arrays
  if (fixed + (long long)kBinMinChunk * pct::bin_bytes_per_query((int)LP, row_bytes, fused) > smem_optin) return 0;
  const size_t vb = 4 * elem_size(dtype);
  if (!aligned(value, vb) || !aligned(grad_out, 16) || !aligned(acc, 16) || !aligned(loc, 8) ||
      !aligned(attn, 4) || !aligned(grad_loc, 8))
    return 0;
  return lanes;
}

// `io` is output (forward) or grad_output (backward); `acc` the fp32 grad_value accumulator
// (backward only).
VecPlan vec_eligible(const Dims& d, int dtype, unsigned flags, const void* value, const void* io,
                     const void* acc, const void* loc, const void* attn) {
  VecPlan none;
  if (flags & PCT_MSDA_FLAG_FORCE_GENERIC) return none;
  if ((long long)d.S * d.M * d.D >= (1ll << 31) || d.S == 0) return none;  // 32-bit in-image offsets
  if ((long long)d.N * d.M * d.Lq >= (1ll << 31)) return none;             // 32-bit task index
  if (!aligned(loc, 8) || !aligned(attn, 4)) return none;
  VecPlan p = vec_plan(d.D, d.L, d.P, dtype, acc != nullptr);
  if (p.lanes == 0) return none;
  if (!aligned(value, p.vb) || !aligned(io, p.vb) || (acc && !aligned(acc, 16))) return none;
  return p;
}

}  // namespace

extern "C" {

int pct_msda_abi_version(void) { return PCT_MSDA_ABI_VERSION; }

uint64_t pct_msda_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* pct_msda_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case PCT_MSDA_E_NULL_POINTER: return "pct_msda: null pointer argument";
    case PCT_MSDA_E_BAD_SHAPE: return "pct_msda: invalid dimension";
    case PCT_MSDA_E_BAD_DTYPE: return "pct_msda: unsupported value_dtype";
    case PCT_MSDA_E_MISALIGNED: return "pct_msda: buffer is not aligned to its element size";
    case PCT_MSDA_E_WORKSPACE: return "pct_msda: workspace missing or too small";
    case PCT_MSDA_E_TOO_LARGE: return "pct_msda: tensor too large";
    case PCT_MSDA_E_UNSUPPORTED: return "pct_msda: shape / dtype not supported by this entry point or flag (fused and deterministic paths need the vector kernels)";
    default:
      if (code > 0 && code < 1000) return cudaGetErrorString((cudaError_t)code);
      return "pct_msda: unknown error code";
  }
}

int pct_msda_kernel_family(int channels, int num_levels, int num_point, int value_dtype,
                           int backward) {
  return vec_plan(channels, num_levels, num_point, value_dtype, backward != 0).lanes > 0 ? 1 : 0;
}

}  // extern "C"

namespace {

// Shared by the plain and the fused entry points.  fused: `sampling_loc` holds raw offsets,
// `attn_weight` raw logits, fr the reference points.
int forward_impl(const void* value, const int64_t* spatial_shapes,
                 const int64_t* level_start_index, const void* sampling_loc,
                 const void* attn_weight, void* output, int batch, int spatial_size,
                 int num_heads, int channels, int num_levels, int num_query, int num_point,
                 int value_dtype, void* cuda_stream, unsigned flags, pct::FusedRef fr) {
  const bool fused = fr.ref != nullptr;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  int rc = check_dims(d, value_dtype);
  if (rc) return rc;
  if ((long long)d.N * d.Lq == 0) return 0;  // empty output, nothing to write
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output)
    return PCT_MSDA_E_NULL_POINTER;
  const size_t es = elem_size(value_dtype);
  const size_t ls = value_dtype == PCT_MSDA_F64 ? 8 : 4;
  if (!aligned(value, es) || !aligned(output, es) || !aligned(sampling_loc, ls) ||
      !aligned(attn_weight, ls) || !aligned(spatial_shapes, 8) || !aligned(level_start_index, 8))
    return PCT_MSDA_E_MISALIGNED;
  int sm_count = 0;
  rc = device_sm_count(&sm_count);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;

  if ((long long)d.S * d.M * d.D == 0) {
    // no value to sample: every point is out of range, the output is all zeros
    return (int)cudaMemsetAsync(output, 0, (size_t)d.N * d.Lq * d.M * d.D * es, st);
  }
  const VecPlan vp = vec_eligible(d, value_dtype, flags, value, output, nullptr, sampling_loc, attn_weight);
  if (FwdLauncher fn = vp.lanes == 0 ? nullptr : pick_fwd(fused, vp, value_dtype))
    return fn(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, d, sm_count,
              st, flags, fr);
  if (fused) return PCT_MSDA_E_UNSUPPORTED;  // the fused prologue exists on the vector path only
  return launch_fwd_generic(value_dtype, value, spatial_shapes, level_start_index, sampling_loc,
                            attn_weight, output, d, sm_count, st);
}

}  // namespace

extern "C" {

int pct_msda_forward_ex(const void* value, const int64_t* spatial_shapes,
                        const int64_t* level_start_index, const void* sampling_loc,
                        const void* attn_weight, void* output, int batch, int spatial_size,
                        int num_heads, int channels, int num_levels, int num_query,
                        int num_point, int value_dtype, void* cuda_stream, unsigned flags) {
  return forward_impl(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output,
                      batch, spatial_size, num_heads, channels, num_levels, num_query, num_point,
                      value_dtype, cuda_stream, flags, pct::FusedRef{nullptr, 0});
}

int pct_msda_fused_forward(const void* value, const int64_t* spatial_shapes,
                           const int64_t* level_start_index, const void* reference_points,
                           int ref_dim, const void* sampling_offsets, const void* attn_logits,
                           void* output, int batch, int spatial_size, int num_heads, int channels,
                           int num_levels, int num_query, int num_point, int value_dtype,
                           void* cuda_stream) {
  if (ref_dim != 2 && ref_dim != 4) return PCT_MSDA_E_BAD_SHAPE;
  if ((long long)batch * num_query > 0 && !reference_points) return PCT_MSDA_E_NULL_POINTER;
  if (!aligned(reference_points, ref_dim == 2 ? 8 : 16)) return PCT_MSDA_E_MISALIGNED;
  if (value_dtype == PCT_MSDA_F64) return PCT_MSDA_E_UNSUPPORTED;
  return forward_impl(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                      output, batch, spatial_size, num_heads, channels, num_levels, num_query,
                      num_point, value_dtype, cuda_stream, 0u,
                      pct::FusedRef{(const float*)reference_points, ref_dim});
}

int pct_msda_forward(const void* value, const int64_t* spatial_shapes,
                     const int64_t* level_start_index, const void* sampling_loc,
                     const void* attn_weight, void* output, int batch, int spatial_size,
                     int num_heads, int channels, int num_levels, int num_query, int num_point,
                     int value_dtype, void* cuda_stream) {
  return pct_msda_forward_ex(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                             output, batch, spatial_size, num_heads, channels, num_levels,
                             num_query, num_point, value_dtype, cuda_stream, 0u);
}

size_t pct_msda_backward_workspace_bytes_ex(int batch, int spatial_size, int num_heads,
                                            int channels, int value_dtype, unsigned flags) {
  if (batch <= 0 || spatial_size <= 0 || num_heads <= 0 || channels <= 0) return 0;
  const size_t n = (size_t)batch * spatial_size * num_heads * channels;
  if (flags & PCT_MSDA_FLAG_DETERMINISTIC) return n * sizeof(long long) + 16;  // + DetScratch
  return pct_msda_backward_workspace_bytes(batch, spatial_size, num_heads, channels, value_dtype);
}

size_t pct_msda_backward_workspace_bytes(int batch, int spatial_size, int num_heads, int channels,
                                         int value_dtype) {
  if (value_dtype != PCT_MSDA_BF16) return 0;
  if (batch <= 0 || spatial_size <= 0 || num_heads <= 0 || channels <= 0) return 0;
  return (size_t)batch * spatial_size * num_heads * channels * sizeof(float);
}

}  // extern "C"

namespace {

// Exact cell count from a HOST copy of spatial_shapes (optional): must describe the same S pixels.
int host_cells(const int64_t* host_shapes, const Dims& d, long long* cells) {
  *cells = 0;
  if (!host_shapes) return 0;
  long long pixels = 0, n = 0;
  for (int l = 0; l < d.L; ++l) {
    const long long H = host_shapes[2 * l], W = host_shapes[2 * l + 1];
    if (H < 0 || W < 0 || H > (1ll << 30) || W > (1ll << 30)) return PCT_MSDA_E_BAD_SHAPE;
    pixels += H * W;
    if (H > 0 && W > 0) n += (H + 1) * (W + 1);
  }
  if (pixels != d.S) return PCT_MSDA_E_BAD_SHAPE;
  *cells = n > 0 ? n : 1;
  return 0;
}

int backward_impl(const void* value, const int64_t* spatial_shapes,
                  const int64_t* level_start_index, const void* sampling_loc,
                  const void* attn_weight, const void* grad_output, void* grad_value,
                  void* grad_sampling_loc, void* grad_attn_weight, int batch, int spatial_size,
                  int num_heads, int channels, int num_levels, int num_query, int num_point,
                  int value_dtype, void* workspace, size_t workspace_bytes, void* cuda_stream,
                  unsigned flags, pct::FusedRef fr, const int64_t* host_shapes = nullptr) {
  const bool fused = fr.ref != nullptr;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  int rc = check_dims(d, value_dtype);
  if (rc) return rc;
  long long cells = 0;
  rc = host_cells(host_shapes, d, &cells);
  if (rc) return rc;
  const long long n_value = (long long)d.N * d.S * d.M * d.D;
  const long long n_pairs = (long long)d.N * d.Lq * d.M;
  if (n_value == 0 && n_pairs == 0) return 0;
  if (!spatial_shapes || !level_start_index) return PCT_MSDA_E_NULL_POINTER;
  if (n_value > 0 && (!value || !grad_value)) return PCT_MSDA_E_NULL_POINTER;
  if (n_pairs > 0 && (!sampling_loc || !attn_weight || !grad_output || !grad_sampling_loc || !grad_attn_weight))
    return PCT_MSDA_E_NULL_POINTER;
  const size_t es = elem_size(value_dtype);
  const size_t ls = value_dtype == PCT_MSDA_F64 ? 8 : 4;
  if (!aligned(value, es) || !aligned(grad_value, es) || !aligned(grad_output, es) ||
      !aligned(sampling_loc, ls) || !aligned(attn_weight, ls) || !aligned(grad_sampling_loc, ls) ||
      !aligned(grad_attn_weight, ls) || !aligned(spatial_shapes, 8) || !aligned(level_start_index, 8))
    return PCT_MSDA_E_MISALIGNED;
  // opt-in bf16 accumulator: grad_value itself is the accumulator (packed bf16x2 reductions), no
  // fp32 workspace; only on the unfused vector path with 8-byte lanes
  BwdLauncher bf16_accum = nullptr;
  if ((flags & PCT_MSDA_FLAG_BF16_ACCUM) && value_dtype == PCT_MSDA_BF16 && !fused && n_value > 0 &&
      n_pairs > 0 && aligned(grad_sampling_loc, 8) && aligned(grad_value, 8)) {
    const VecPlan p = vec_eligible(d, value_dtype, flags, value, grad_output, grad_value, sampling_loc,
                                   attn_weight);
    if (p.lanes > 0) bf16_accum = pick_bwd_bf16_accum(p);
  }
  // opt-in deterministic accumulator: 64-bit fixed point in the workspace (+ DetScratch tail)
  BwdLauncher det_fn = nullptr;
  if (flags & PCT_MSDA_FLAG_DETERMINISTIC) {
    if (fused || value_dtype == PCT_MSDA_F64 || n_value == 0 || n_pairs == 0 || !aligned(grad_sampling_loc, 8))
      return PCT_MSDA_E_UNSUPPORTED;
    const VecPlan p = vec_eligible(d, value_dtype, flags, value, grad_output, workspace, sampling_loc,
                                   attn_weight);
    det_fn = p.lanes > 0 ? pick_bwd_deterministic(p, value_dtype) : nullptr;
    if (!det_fn) return PCT_MSDA_E_UNSUPPORTED;
    bf16_accum = nullptr;
  }
  int smem_optin = 0;
  rc = device_smem_optin(&smem_optin);
  if (rc) return rc;
  // opt-in for bf16 values on the per-tap vector path: accumulate in a SCALED FP16 image inside the
  // workspace (half the L2 reduction sectors of the fp32 scratch image, msda_kernels.cuh)
  BwdLauncher half_fn = nullptr;
  if ((flags & PCT_MSDA_FLAG_BF16_F16_ACCUM) && value_dtype == PCT_MSDA_BF16 && !det_fn && !bf16_accum &&
      n_value >= 64 && n_pairs > 0 && !(flags & PCT_MSDA_FLAG_FORCE_GENERIC) &&
      aligned(grad_sampling_loc, 8) && workspace && aligned(workspace, 16) &&
      // (an explicit request for the binned kernel wins over this per-tap option, the default does not)
      !binned_plan(d, value_dtype, !(flags & PCT_MSDA_FLAG_BWD_BINNED) ? (flags | PCT_MSDA_FLAG_BWD_PER_TAP) : flags, smem_optin, value,
                   grad_output, workspace, sampling_loc, attn_weight, grad_sampling_loc, fused, cells)) {
    const VecPlan p = vec_eligible(d, value_dtype, flags, value, grad_output, workspace, sampling_loc,
                                   attn_weight);
    if (p.lanes > 0) half_fn = pick_bwd_half_accum(fused, p);
  }
  const size_t half_image = (((size_t)n_value * 2 + 15) / 16) * 16;   // DetScratch follows, 16-aligned
  const size_t need = det_fn ? pct_msda_backward_workspace_bytes_ex(batch, spatial_size, num_heads, channels,
                                                                    value_dtype, flags)
                      : bf16_accum ? 0
                                   : pct_msda_backward_workspace_bytes(batch, spatial_size, num_heads,
                                                                       channels, value_dtype);
  if (need > 0 && (!workspace || workspace_bytes < need || !aligned(workspace, 16)))
    return PCT_MSDA_E_WORKSPACE;
  int sm_count = 0;
  rc = device_sm_count(&sm_count);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;

  // grad_value accumulates through reductions: clear the accumulator image first.
  const bool ws_accum = det_fn || (value_dtype == PCT_MSDA_BF16 && !bf16_accum);
  void* acc_image = ws_accum ? workspace : grad_value;
  const size_t acc_bytes = half_fn    ? half_image + sizeof(pct::DetScratch)
                           : ws_accum ? need
                                      : (size_t)n_value * es;
  pct::DetScratch* det = det_fn    ? reinterpret_cast<pct::DetScratch*>((char*)workspace + (size_t)n_value * 8)
                         : half_fn ? reinterpret_cast<pct::DetScratch*>((char*)workspace + half_image)
                                   : nullptr;
  if (acc_bytes > 0) {
    rc = zero_fill(acc_image, acc_bytes, sm_count, st);
    if (rc) return rc;
  }
  if (n_pairs == 0 || n_value == 0) {
    // no samples can touch a value: grads w.r.t. loc/attn are all zero
    if (n_pairs > 0) {
      cudaError_t e = cudaMemsetAsync(grad_sampling_loc, 0, (size_t)n_pairs * d.L * d.P * 2 * ls, st);
      if (e != cudaSuccess) return (int)e;
      e = cudaMemsetAsync(grad_attn_weight, 0, (size_t)n_pairs * d.L * d.P * ls, st);
      if (e != cudaSuccess) return (int)e;
    }
    if (value_dtype == PCT_MSDA_BF16 && n_value > 0) {
      cudaError_t e = cudaMemsetAsync(grad_value, 0, (size_t)n_value * es, st);
      if (e != cudaSuccess) return (int)e;
    }
    return 0;
  }

  const VecPlan vp = aligned(grad_sampling_loc, 8)
                         ? vec_eligible(d, value_dtype, flags, value, grad_output, acc_image,
                                        sampling_loc, attn_weight)
                         : VecPlan{};
  if (det_fn) {
    // scale from max|grad_out| (accumulator and DetScratch were just zeroed)
    const long long n_go = n_pairs * d.D;
    long long want = (n_go + 1023) / 1024;
    const int grid = (int)(want < 1 ? 1 : (want < (long long)sm_count * 8 ? want : (long long)sm_count * 8));
    if (value_dtype == PCT_MSDA_F32)
      pct::msda_absmax_kernel<float><<<grid, 256, 0, st>>>((const float*)grad_output, n_go, det);
    else
      pct::msda_absmax_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)grad_output, n_go, det);
    pct::msda_det_scale_kernel<<<1, 1, 0, st>>>(det);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  if (half_fn) {
    // scale from max|grad_out|, max sum|attn| and Lq (accumulator and DetScratch were just zeroed)
    const long long n_go = n_pairs * d.D;
    const long long cap = (long long)sm_count * 8;
    long long want = (n_go + 1023) / 1024;
    pct::msda_absmax_kernel<__nv_bfloat16><<<(int)(want < 1 ? 1 : (want < cap ? want : cap)), 256, 0, st>>>(
        (const __nv_bfloat16*)grad_output, n_go, det);
    if (fused) {
      // the kernel's own softmax produces the weights: every row sums to exactly one
      const unsigned one = 0x3f800000u;
      cudaError_t e = cudaMemcpyAsync(&det->rowsum_bits, &one, sizeof(one), cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return (int)e;
    } else {
      want = (n_pairs + 255) / 256;
      pct::msda_attn_rowsum_max_kernel<<<(int)(want < 1 ? 1 : (want < cap ? want : cap)), 256, 0, st>>>(
          (const float*)attn_weight, n_pairs, d.L * d.P, det);
    }
    pct::msda_half_scale_kernel<<<1, 1, 0, st>>>(det, d.Lq);
    g_launches.fetch_add(fused ? 2 : 3, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  const int bin_lanes = (det_fn || bf16_accum || half_fn)
                            ? 0
                            : binned_plan(d, value_dtype, flags, smem_optin, value, grad_output, acc_image,
                                          sampling_loc, attn_weight, grad_sampling_loc, fused, cells);
  if (bin_lanes) {
    rc = dispatch_bwd_binned(bin_lanes, value_dtype, value, spatial_shapes, level_start_index,
                             sampling_loc, attn_weight, grad_output, acc_image, grad_sampling_loc,
                             grad_attn_weight, d, sm_count, smem_optin, st, flags,
                             fused ? fr : pct::FusedRef{nullptr, 0}, cells_bound(d, cells));
  } else if (BwdLauncher fn = det_fn          ? det_fn
                              : half_fn       ? half_fn
                              : bf16_accum    ? bf16_accum
                              : vp.lanes == 0 ? nullptr
                                              : pick_bwd(fused, vp, value_dtype)) {
    rc = fn(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
            acc_image, grad_sampling_loc, grad_attn_weight, d, sm_count, st, flags, fr, det);
  } else if (fused) {
    return PCT_MSDA_E_UNSUPPORTED;
  } else {
    rc = launch_bwd_generic(value_dtype, value, spatial_shapes, level_start_index, sampling_loc,
                            attn_weight, grad_output, acc_image, grad_sampling_loc, grad_attn_weight,
                            d, sm_count, st);
  }
  if (rc) return rc;
  if (det_fn) {
    long long want = (n_value + 255) / 256;
    const int grid = (int)(want < (long long)sm_count * 16 ? want : (long long)sm_count * 16);
    if (value_dtype == PCT_MSDA_F32)
      pct::msda_fixed_to_value_kernel<float><<<grid, 256, 0, st>>>((const long long*)workspace, (float*)grad_value, n_value, det);
    else
      pct::msda_fixed_to_value_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const long long*)workspace, (__nv_bfloat16*)grad_value, n_value, det);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
  }
  if (half_fn) {
    long long want = (n_value / 8 + 255) / 256;
    const long long cap = (long long)sm_count * 16;
    pct::msda_half_to_bf16_kernel<<<(int)(want < 1 ? 1 : (want < cap ? want : cap)), 256, 0, st>>>(
        (const __half*)workspace, (__nv_bfloat16*)grad_value, n_value, det);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
  }
  if (ws_accum)
    return round_workspace_to_bf16((const float*)workspace, grad_value, n_value, sm_count, st);
  return 0;
}

}  // namespace

extern "C" {

int pct_msda_backward_plan(int batch, int spatial_size, int num_heads, int channels,
                           int num_levels, int num_query, int num_point, int value_dtype,
                           unsigned flags) {
  return pct_msda_backward_plan_ex(batch, spatial_size, num_heads, channels, num_levels, num_query,
                                   num_point, value_dtype, flags, nullptr);
}

int pct_msda_backward_plan_ex(int batch, int spatial_size, int num_heads, int channels,
                              int num_levels, int num_query, int num_point, int value_dtype,
                              unsigned flags, const int64_t* host_spatial_shapes) {
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  int rc = check_dims(d, value_dtype);
  if (rc) return -rc;
  long long cells = 0;
  rc = host_cells(host_spatial_shapes, d, &cells);
  if (rc) return -rc;
  const bool special = flags & (PCT_MSDA_FLAG_DETERMINISTIC | PCT_MSDA_FLAG_BF16_ACCUM);
  if (!special) {
    int smem_optin = 0;
    rc = device_smem_optin(&smem_optin);
    if (rc) return -rc;
    alignas(16) static const char dummy[16] = {0};
    if (binned_plan(d, value_dtype, flags, smem_optin, dummy, dummy, dummy, dummy, dummy, dummy, false, cells))
      return 2;
  }
  if (flags & PCT_MSDA_FLAG_FORCE_GENERIC) return 0;
  if ((long long)d.S * d.M * d.D >= (1ll << 31) || d.S == 0) return 0;
  if ((long long)d.N * d.M * d.Lq >= (1ll << 31)) return 0;
  return vec_plan(d.D, d.L, d.P, value_dtype, true).lanes > 0 ? 1 : 0;
}

int pct_msda_backward_ex(const void* value, const int64_t* spatial_shapes,
                         const int64_t* level_start_index, const void* sampling_loc,
                         const void* attn_weight, const void* grad_output, void* grad_value,
                         void* grad_sampling_loc, void* grad_attn_weight, int batch,
                         int spatial_size, int num_heads, int channels, int num_levels,
                         int num_query, int num_point, int value_dtype, void* workspace,
                         size_t workspace_bytes, void* cuda_stream, unsigned flags) {
  return pct_msda_backward_ex2(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                               grad_output, grad_value, grad_sampling_loc, grad_attn_weight, batch,
                               spatial_size, num_heads, channels, num_levels, num_query, num_point,
                               value_dtype, workspace, workspace_bytes, cuda_stream, flags, nullptr);
}

int pct_msda_backward_ex2(const void* value, const int64_t* spatial_shapes,
                          const int64_t* level_start_index, const void* sampling_loc,
                          const void* attn_weight, const void* grad_output, void* grad_value,
                          void* grad_sampling_loc, void* grad_attn_weight, int batch,
                          int spatial_size, int num_heads, int channels, int num_levels,
                          int num_query, int num_point, int value_dtype, void* workspace,
                          size_t workspace_bytes, void* cuda_stream, unsigned flags,
                          const int64_t* host_spatial_shapes) {
  return backward_impl(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                       grad_output, grad_value, grad_sampling_loc, grad_attn_weight, batch,
                       spatial_size, num_heads, channels, num_levels, num_query, num_point,
                       value_dtype, workspace, workspace_bytes, cuda_stream, flags,
                       pct::FusedRef{nullptr, 0}, host_spatial_shapes);
}

int pct_msda_fused_backward(const void* value, const int64_t* spatial_shapes,
                            const int64_t* level_start_index, const void* reference_points,
                            int ref_dim, const void* sampling_offsets, const void* attn_logits,
                            const void* grad_output, void* grad_value, void* grad_offsets,
                            void* grad_logits, int batch, int spatial_size, int num_heads,
                            int channels, int num_levels, int num_query, int num_point,
                            int value_dtype, void* workspace, size_t workspace_bytes,
                            void* cuda_stream) {
  return pct_msda_fused_backward_ex(value, spatial_shapes, level_start_index, reference_points, ref_dim,
                                    sampling_offsets, attn_logits, grad_output, grad_value,
                                    grad_offsets, grad_logits, batch, spatial_size, num_heads,
                                    channels, num_levels, num_query, num_point, value_dtype,
                                    workspace, workspace_bytes, cuda_stream, 0u);
}

int pct_msda_fused_backward_ex(const void* value, const int64_t* spatial_shapes,
                               const int64_t* level_start_index, const void* reference_points,
                               int ref_dim, const void* sampling_offsets, const void* attn_logits,
                               const void* grad_output, void* grad_value, void* grad_offsets,
                               void* grad_logits, int batch, int spatial_size, int num_heads,
                               int channels, int num_levels, int num_query, int num_point,
                               int value_dtype, void* workspace, size_t workspace_bytes,
                               void* cuda_stream, unsigned flags) {
  return pct_msda_fused_backward_ex2(value, spatial_shapes, level_start_index, reference_points, ref_dim,
                                     sampling_offsets, attn_logits, grad_output, grad_value,
                                     grad_offsets, grad_logits, batch, spatial_size, num_heads,
                                     channels, num_levels, num_query, num_point, value_dtype,
                                     workspace, workspace_bytes, cuda_stream, flags, nullptr);
}

int pct_msda_fused_backward_ex2(const void* value, const int64_t* spatial_shapes,
                                const int64_t* level_start_index, const void* reference_points,
                                int ref_dim, const void* sampling_offsets, const void* attn_logits,
                                const void* grad_output, void* grad_value, void* grad_offsets,
                                void* grad_logits, int batch, int spatial_size, int num_heads,
                                int channels, int num_levels, int num_query, int num_point,
                                int value_dtype, void* workspace, size_t workspace_bytes,
                                void* cuda_stream, unsigned flags,
                                const int64_t* host_spatial_shapes) {
  if (flags & ~(PCT_MSDA_FLAG_BWD_PER_TAP | PCT_MSDA_FLAG_BWD_BINNED | PCT_MSDA_FLAG_BF16_F16_ACCUM |
                0xffffff00u))
    return PCT_MSDA_E_UNSUPPORTED;   // the other options exist for the plain entry point only
  if (ref_dim != 2 && ref_dim != 4) return PCT_MSDA_E_BAD_SHAPE;
  if ((long long)batch * num_query > 0 && !reference_points) return PCT_MSDA_E_NULL_POINTER;
  if (!aligned(reference_points, ref_dim == 2 ? 8 : 16)) return PCT_MSDA_E_MISALIGNED;
  if (value_dtype == PCT_MSDA_F64) return PCT_MSDA_E_UNSUPPORTED;
  if ((long long)batch * spatial_size == 0) return PCT_MSDA_E_UNSUPPORTED;
  return backward_impl(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                       grad_output, grad_value, grad_offsets, grad_logits, batch, spatial_size,
                       num_heads, channels, num_levels, num_query, num_point, value_dtype,
                       workspace, workspace_bytes, cuda_stream, flags,
                       pct::FusedRef{(const float*)reference_points, ref_dim}, host_spatial_shapes);
}

int pct_msda_backward(const void* value, const int64_t* spatial_shapes,
                      const int64_t* level_start_index, const void* sampling_loc,
                      const void* attn_weight, const void* grad_output, void* grad_value,
                      void* grad_sampling_loc, void* grad_attn_weight, int batch,
                      int spatial_size, int num_heads, int channels, int num_levels,
                      int num_query, int num_point, int value_dtype, void* workspace,
                      size_t workspace_bytes, void* cuda_stream) {
  return pct_msda_backward_ex(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                              grad_output, grad_value, grad_sampling_loc, grad_attn_weight, batch,
                              spatial_size, num_heads, channels, num_levels, num_query, num_point,
                              value_dtype, workspace, workspace_bytes, cuda_stream, 0u);
}

}  // extern "C"
