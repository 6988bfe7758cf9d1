// msda_inst_bwd.cu -- instantiations and launchers of the per-tap backward kernels (msda_kernels.cuh).
#include "../../include/pct_msda.h"
#include "msda_launch.h"

namespace pctl {
namespace {
template <typename T, int LPP, int VB, bool FUSED, typename ACC = float>
int launch_bwd_vec(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                   const void* attn, const void* grad_out, void* gacc, void* grad_loc,
                   void* grad_attn, const Dims& d, int sm_count, cudaStream_t st, unsigned flags,
                   pct::FusedRef fr, const pct::DetScratch* det) {
  constexpr int PPW = 32 / LPP;
  const long long n_tasks = (long long)d.N * d.M * ((d.Lq + PPW - 1) / PPW);
  const LaunchShape ls = vec_launch_shape(n_tasks, sm_count, flags,
                                          FUSED ? pct::kBwdFusedThreads : pct::kBwdVecThreads);
  pct::msda_bwd_vec_kernel<T, LPP, VB, FUSED, ACC><<<ls.grid, ls.block, 0, st>>>(
      (const T*)value, shapes, lsi, (const float*)loc, (const float*)attn, (const T*)grad_out,
      (ACC*)gacc, (float*)grad_loc, (float*)grad_attn, d.N, d.S, d.M, d.L, d.Lq, d.P, fr,
      make_sched(n_tasks, ls.grid, ls.block / 32, flags), det);
  count_launches(1);
  return (int)cudaGetLastError();
}

#define PCT_LANE_CASES(LAUNCH, T, VB, FUSED)                                             \
  switch (plan.lanes) {                                                                 \
    case 1: return LAUNCH<T, 1, VB, FUSED>;                                             \
    case 2: return LAUNCH<T, 2, VB, FUSED>;                                             \
    case 4: return LAUNCH<T, 4, VB, FUSED>;                                             \
    case 8: return LAUNCH<T, 8, VB, FUSED>;                                             \
    case 16: return LAUNCH<T, 16, VB, FUSED>;                                           \
    case 32: return LAUNCH<T, 32, VB, FUSED>;                                           \
  }                                                                                     \
  return nullptr;

template <bool FUSED>
BwdLauncher pick_bwd_t(const VecPlan& plan, int dtype) {
  if (dtype == PCT_MSDA_F32) {
    if (plan.vb == 16) { PCT_LANE_CASES(launch_bwd_vec, float, 16, FUSED) }
    if (plan.vb == 32 && plan.lanes == 32) return launch_bwd_vec<float, 32, 32, FUSED>;
  } else {
    if (plan.vb == 8) { PCT_LANE_CASES(launch_bwd_vec, __nv_bfloat16, 8, FUSED) }
    if (plan.vb == 16 && plan.lanes == 32) return launch_bwd_vec<__nv_bfloat16, 32, 16, FUSED>;
  }
  return nullptr;
}
#undef PCT_LANE_CASES

int generic_grid(long long n_tasks, int sm_count) {
  const int warps = pct::kGenericThreads / 32;
  long long want = (n_tasks + warps - 1) / warps;
  long long cap = (long long)sm_count * 8;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <typename T, typename CT, typename AT>
int launch_bwd_generic_t(const void* value, const int64_t* shapes, const int64_t* lsi,
                       const void* loc, const void* attn, const void* grad_out, AT* gacc,
                       void* grad_loc, void* grad_attn, const Dims& d, int sm_count,
                       cudaStream_t st) {
  const long long n_tasks = (long long)d.N * d.M * d.Lq;
  pct::msda_bwd_generic_kernel<T, CT, AT><<<generic_grid(n_tasks, sm_count), pct::kGenericThreads, 0, st>>>(
      (const T*)value, shapes, lsi, (const CT*)loc, (const CT*)attn, (const T*)grad_out, gacc,
      (CT*)grad_loc, (CT*)grad_attn, d.N, d.S, d.M, d.D, d.L, d.Lq, d.P);
  count_launches(1);
  return (int)cudaGetLastError();
}

}  // namespace

// deterministic fixed-point accumulator (opt-in): lanes of 4 channels only
BwdLauncher pick_bwd_deterministic(const VecPlan& plan, int dtype) {
#define PCT_DET_CASES(T, VB)                                                   \
  switch (plan.lanes) {                                                        \
    case 1: return launch_bwd_vec<T, 1, VB, false, long long>;                 \
    case 2: return launch_bwd_vec<T, 2, VB, false, long long>;                 \
    case 4: return launch_bwd_vec<T, 4, VB, false, long long>;                 \
    case 8: return launch_bwd_vec<T, 8, VB, false, long long>;                 \
    case 16: return launch_bwd_vec<T, 16, VB, false, long long>;               \
    case 32: return launch_bwd_vec<T, 32, VB, false, long long>;               \
  }
  if (dtype == PCT_MSDA_F32 && plan.vb == 16) { PCT_DET_CASES(float, 16) }
  if (dtype == PCT_MSDA_BF16 && plan.vb == 8) { PCT_DET_CASES(__nv_bfloat16, 8) }
#undef PCT_DET_CASES
  return nullptr;
}

// bf16 values, bf16 accumulator (opt-in): 8-byte lanes only
BwdLauncher pick_bwd_bf16_accum(const VecPlan& plan) {
  if (plan.vb != 8) return nullptr;
  switch (plan.lanes) {
    case 1: return launch_bwd_vec<__nv_bfloat16, 1, 8, false, __nv_bfloat16>;
    case 2: return launch_bwd_vec<__nv_bfloat16, 2, 8, false, __nv_bfloat16>;
    case 4: return launch_bwd_vec<__nv_bfloat16, 4, 8, false, __nv_bfloat16>;
    case 8: return launch_bwd_vec<__nv_bfloat16, 8, 8, false, __nv_bfloat16>;
    case 16: return launch_bwd_vec<__nv_bfloat16, 16, 8, false, __nv_bfloat16>;
    case 32: return launch_bwd_vec<__nv_bfloat16, 32, 8, false, __nv_bfloat16>;
  }
  return nullptr;
}


// bf16 values, scaled fp16 accumulator in the workspace (opt-in): 8-byte lanes only
BwdLauncher pick_bwd_half_accum(bool fused, const VecPlan& plan) {
  if (plan.vb != 8) return nullptr;
#define PCT_HALF_CASES(FUSED)                                                               \
  switch (plan.lanes) {                                                                     \
    case 2: return launch_bwd_vec<__nv_bfloat16, 2, 8, FUSED, __half>;                      \
    case 4: return launch_bwd_vec<__nv_bfloat16, 4, 8, FUSED, __half>;                      \
    case 8: return launch_bwd_vec<__nv_bfloat16, 8, 8, FUSED, __half>;                      \
    case 16: return launch_bwd_vec<__nv_bfloat16, 16, 8, FUSED, __half>;                    \
    case 32: return launch_bwd_vec<__nv_bfloat16, 32, 8, FUSED, __half>;                    \
  }
  if (fused) { PCT_HALF_CASES(true) } else { PCT_HALF_CASES(false) }
#undef PCT_HALF_CASES
  return nullptr;
}

BwdLauncher pick_bwd(bool fused, const VecPlan& plan, int dtype) {
  return fused ? pick_bwd_t<true>(plan, dtype) : pick_bwd_t<false>(plan, dtype);
}

int launch_bwd_generic(int dtype, const void* value, const int64_t* shapes, const int64_t* lsi,
                       const void* loc, const void* attn, const void* grad_out, void* gacc,
                       void* grad_loc, void* grad_attn, const Dims& d, int sm_count, cudaStream_t st) {
  switch (dtype) {
    case PCT_MSDA_F32:
      return launch_bwd_generic_t<float, float, float>(value, shapes, lsi, loc, attn, grad_out, (float*)gacc,
                                                       grad_loc, grad_attn, d, sm_count, st);
    case PCT_MSDA_BF16:
      return launch_bwd_generic_t<__nv_bfloat16, float, float>(value, shapes, lsi, loc, attn, grad_out,
                                                               (float*)gacc, grad_loc, grad_attn, d, sm_count, st);
    default:
      return launch_bwd_generic_t<double, double, double>(value, shapes, lsi, loc, attn, grad_out,
                                                          (double*)gacc, grad_loc, grad_attn, d, sm_count, st);
  }
}

}  // namespace pctl
