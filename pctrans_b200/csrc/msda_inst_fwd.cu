// msda_inst_fwd.cu -- instantiations and launchers of the forward kernels (msda_kernels.cuh).
#include <type_traits>

#include "../../include/pct_msda.h"
#include "msda_launch.h"

namespace pctl {
namespace {
template <typename T, int LPP, int VB, bool FUSED>
int launch_fwd_vec(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                   const void* attn, void* out, const Dims& d, int sm_count, cudaStream_t st,
                   unsigned flags, pct::FusedRef fr) {
  constexpr int PPW = 32 / LPP;
  const long long n_tasks = (long long)d.N * d.M * ((d.Lq + PPW - 1) / PPW);
  const LaunchShape ls = vec_launch_shape(n_tasks, sm_count, flags, pct::kVecThreads);
  if constexpr (std::is_same<T, float>::value && LPP == 8 && VB == 16 && !FUSED) {
    // experiment (PCT_MSDA_FLAG_FWD_STAGED): the smallest levels of a head staged in shared memory
    if ((flags & PCT_MSDA_FLAG_FWD_STAGED) && d.M > 0 && (d.M & (d.M - 1)) == 0) {
      auto kern = pct::msda_fwd_vec_kernel<float, 8, 16, false, true>;
      static int dyn_bytes[64] = {};
      int dev = 0;
      cudaError_t e = cudaGetDevice(&dev);
      if (e != cudaSuccess) return (int)e;
      if (dev < 0 || dev >= 64) return PCT_MSDA_E_UNSUPPORTED;
      if (dyn_bytes[dev] == 0) {
        int optin = 0;
        cudaFuncAttributes fa;
        e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return (int)e;
        const int dyn = (optin - (int)fa.sharedSizeBytes) & ~127;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        if (e != cudaSuccess) return (int)e;
        dyn_bytes[dev] = dyn;
      }
      const int dyn = dyn_bytes[dev];
      kern<<<ls.grid, ls.block, dyn, st>>>(
          (const float*)value, shapes, lsi, (const float*)loc, (const float*)attn, (float*)out, d.N,
          d.S, d.M, d.L, d.Lq, d.P, fr, make_sched(n_tasks, ls.grid, ls.block / 32, flags),
          dyn / (32 * (int)sizeof(float)));
      count_launches(1);
      return (int)cudaGetLastError();
    }
  }
  pct::msda_fwd_vec_kernel<T, LPP, VB, FUSED><<<ls.grid, ls.block, 0, st>>>(
      (const T*)value, shapes, lsi, (const float*)loc, (const float*)attn, (T*)out, d.N, d.S, d.M,
      d.L, d.Lq, d.P, fr, make_sched(n_tasks, ls.grid, ls.block / 32, flags));
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
FwdLauncher pick_fwd_t(const VecPlan& plan, int dtype) {
  if (dtype == PCT_MSDA_F32) {
    if (plan.vb == 16) { PCT_LANE_CASES(launch_fwd_vec, float, 16, FUSED) }
    if (plan.vb == 32 && plan.lanes == 32) return launch_fwd_vec<float, 32, 32, FUSED>;
  } else {
    if (plan.vb == 16) { PCT_LANE_CASES(launch_fwd_vec, __nv_bfloat16, 16, FUSED) }
    if (plan.vb == 8 && plan.lanes == 1) return launch_fwd_vec<__nv_bfloat16, 1, 8, FUSED>;
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

template <typename T, typename CT>
int launch_fwd_generic_t(const void* value, const int64_t* shapes, const int64_t* lsi,
                       const void* loc, const void* attn, void* out, const Dims& d, int sm_count,
                       cudaStream_t st) {
  const long long n_tasks = (long long)d.N * d.M * d.Lq;
  pct::msda_fwd_generic_kernel<T, CT><<<generic_grid(n_tasks, sm_count), pct::kGenericThreads, 0, st>>>(
      (const T*)value, shapes, lsi, (const CT*)loc, (const CT*)attn, (T*)out, d.N, d.S, d.M, d.D,
      d.L, d.Lq, d.P);
  count_launches(1);
  return (int)cudaGetLastError();
}


}  // namespace

FwdLauncher pick_fwd(bool fused, const VecPlan& plan, int dtype) {
  return fused ? pick_fwd_t<true>(plan, dtype) : pick_fwd_t<false>(plan, dtype);
}

int launch_fwd_generic(int dtype, const void* value, const int64_t* shapes, const int64_t* lsi,
                       const void* loc, const void* attn, void* out, const Dims& d, int sm_count,
                       cudaStream_t st) {
  switch (dtype) {
    case PCT_MSDA_F32:
      return launch_fwd_generic_t<float, float>(value, shapes, lsi, loc, attn, out, d, sm_count, st);
    case PCT_MSDA_BF16:
      return launch_fwd_generic_t<__nv_bfloat16, float>(value, shapes, lsi, loc, attn, out, d, sm_count, st);
    default:
      return launch_fwd_generic_t<double, double>(value, shapes, lsi, loc, attn, out, d, sm_count, st);
  }
}

}  // namespace pctl
