// msda_launch.h -- internal interface between the C-ABI translation unit (msda_capi.cu) and the
// translation units that instantiate the kernels (msda_inst_*.cu).  The kernels are heavy
// templates; one translation unit per kernel family keeps rebuilds short and lets them compile in
// parallel.  Not part of the public ABI (include/pct_msda.h is).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_kernels.cuh"

namespace pctl {

struct Dims {
  int N, S, M, D, L, Lq, P;
};

// How a (query, head) pair is spread over lanes on the vector path.
struct VecPlan {
  int lanes = 0;  // lanes per pair (0 = shape does not fit the vector path)
  int vb = 0;     // bytes of channels per lane
};

using FwdLauncher = int (*)(const void*, const int64_t*, const int64_t*, const void*, const void*,
                            void*, const Dims&, int, cudaStream_t, unsigned, pct::FusedRef);
using BwdLauncher = int (*)(const void*, const int64_t*, const int64_t*, const void*, const void*,
                            const void*, void*, void*, void*, const Dims&, int, cudaStream_t,
                            unsigned, pct::FusedRef, const pct::DetScratch*);

// msda_capi.cu
void count_launches(int n);
unsigned next_sched_slot();
pct::SchedArgs make_sched(long long n_tasks, int grid, int warps, unsigned flags);

struct LaunchShape {
  int grid, block;
};
LaunchShape vec_launch_shape(long long n_tasks, int sm_count, unsigned flags, int default_block);

// msda_inst_fwd.cu
FwdLauncher pick_fwd(bool fused, const VecPlan& plan, int dtype);
int launch_fwd_generic(int dtype, const void* value, const int64_t* shapes, const int64_t* lsi,
                       const void* loc, const void* attn, void* out, const Dims& d, int sm_count,
                       cudaStream_t st);

// msda_inst_bwd.cu
BwdLauncher pick_bwd(bool fused, const VecPlan& plan, int dtype);
BwdLauncher pick_bwd_deterministic(const VecPlan& plan, int dtype);
BwdLauncher pick_bwd_bf16_accum(const VecPlan& plan);
BwdLauncher pick_bwd_half_accum(bool fused, const VecPlan& plan);   // bf16 values, scaled fp16 accumulator
int launch_bwd_generic(int dtype, const void* value, const int64_t* shapes, const int64_t* lsi,
                       const void* loc, const void* attn, const void* grad_out, void* gacc,
                       void* grad_loc, void* grad_attn, const Dims& d, int sm_count, cudaStream_t st);

// msda_inst_binned.cu
constexpr int kBinMinChunk = 64;   // below this many queries per chunk the per-tap kernel is used
int dispatch_bwd_binned(int lanes, int dtype, const void* value, const int64_t* shapes,
                        const int64_t* lsi, const void* loc, const void* attn, const void* grad_out,
                        void* gacc, void* grad_loc, void* grad_attn, const Dims& d, int sm_count,
                        int smem_optin, cudaStream_t st, unsigned flags, pct::FusedRef fr,
                        long long cells_bound);

}  // namespace pctl
