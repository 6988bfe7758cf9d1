// msda_inst_binned.cu -- instantiations and launcher of the cell-binned backward (msda_bwd_binned.cuh).
#include <atomic>

#include "../../include/pct_msda.h"
#include "msda_bwd_binned.cuh"
#include "msda_launch.h"

#ifndef PCT_BIN_PDL
#define PCT_BIN_PDL 1
#endif

namespace pctl {
namespace {

template <typename T, int LPP, bool FUSED>
int launch_bwd_binned(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                      const void* attn, const void* grad_out, void* gacc, void* grad_loc,
                      void* grad_attn, const Dims& d, int sm_count, int smem_optin, cudaStream_t st,
                      unsigned flags, pct::FusedRef fr, long long cells_bound) {
  static std::atomic<int> configured[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  auto kern = pct::msda_bwd_binned_kernel<T, LPP, FUSED>;
  // dynamic shared memory = the opt-in maximum minus the kernel's static arrays
  int dyn = (dev >= 0 && dev < 64) ? configured[dev].load(std::memory_order_acquire) : 0;
  if (dyn == 0) {
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return (int)e;
    // (per resident CTA: the SM's shared memory holds kBinCtasPerSm of them, each with 1 KB of
    // system reserve)
    dyn = ((smem_optin + 1024) / pct::kBinCtasPerSm - 1024 - (int)fa.sharedSizeBytes) & ~15;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) configured[dev].store(dyn, std::memory_order_release);
  }
  pct::BinLaunch bl;
  bl.smem_bytes = dyn;
  bl.qc_cap = (int)((flags >> 25) & 0x7fu) * 8;
  if (bl.qc_cap == 0) {
    // Small problems: when the chunks that fit shared memory would give the SMs fewer than one and
    // a half tasks each, shorter chunks do (the kernel takes the smaller of this cap and what fits).
    const long long streams = (long long)d.N * d.M;
    const long long per_q = pct::bin_bytes_per_query(d.L * d.P, LPP * 4 * (int)sizeof(T), FUSED);
    long long qc_fit = (dyn - pct::bin_bytes_fixed(cells_bound)) / per_q;
    if (qc_fit > pct::kBinMaxQc) qc_fit = pct::kBinMaxQc;
    if (qc_fit < kBinMinChunk) qc_fit = kBinMinChunk;
    const long long chunks_fit = (d.Lq + qc_fit - 1) / qc_fit;
    const long long sms = (long long)sm_count * pct::kBinCtasPerSm;
    if (streams * chunks_fit * 2 < sms * 3) {
      // Measured (profiles/r02_binned_v3_small_batches.log): with fewer tasks than SMs, one task per
      // SM; with about one per SM, one and a half -- tasks differ in length (the chunk that holds
      // the coarse levels' queries touches cells all over the maps) and are handed out dynamically.
      const long long want_chunks = streams * chunks_fit * 4 < sms * 3
                                        ? (sms + streams - 1) / streams
                                        : (3 * sms + 2 * streams - 1) / (2 * streams);
      long long cap = ((d.Lq + want_chunks - 1) / want_chunks + 7) & ~7ll;
      if (cap < kBinMinChunk) cap = kBinMinChunk;
      if (cap < qc_fit) bl.qc_cap = (int)cap;
    }
  }
  bl.slot = (int)(next_sched_slot());
  const long long want = (long long)d.N * d.M * ((d.Lq + kBinMinChunk - 1) / kBinMinChunk);
  const long long cap = (long long)sm_count * pct::kBinCtasPerSm;
  const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
  // Programmatic dependent launch behind the zero-fill of the accumulator image: the CTAs start
  // while it still runs, sort their first chunk, and wait before the first reduction.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)pct::kBinThreads);
  cfg.dynamicSmemBytes = (size_t)dyn;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = PCT_BIN_PDL;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, (const T*)value, shapes, lsi, (const float*)loc,
                         (const float*)attn, (const T*)grad_out, (float*)gacc, (float*)grad_loc,
                         (float*)grad_attn, d.N, d.S, d.M, d.L, d.Lq, d.P, bl, fr);
  count_launches(1);
  return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

}  // namespace

int dispatch_bwd_binned(int lanes, int dtype, const void* value, const int64_t* shapes,
                        const int64_t* lsi, const void* loc, const void* attn, const void* grad_out,
                        void* gacc, void* grad_loc, void* grad_attn, const Dims& d, int sm_count,
                        int smem_optin, cudaStream_t st, unsigned flags, pct::FusedRef fr,
                        long long cells_bound) {
#define PCT_BIN_CASE(T, LPP)                                                                       \
  return fr.ref ? launch_bwd_binned<T, LPP, true>(value, shapes, lsi, loc, attn, grad_out, gacc,   \
                                                  grad_loc, grad_attn, d, sm_count, smem_optin,    \
                                                  st, flags, fr, cells_bound)                      \
                : launch_bwd_binned<T, LPP, false>(value, shapes, lsi, loc, attn, grad_out, gacc,  \
                                                   grad_loc, grad_attn, d, sm_count, smem_optin,   \
                                                   st, flags, fr, cells_bound)
  if (dtype == PCT_MSDA_F32) {
    switch (lanes) {
      case 4: PCT_BIN_CASE(float, 4);
      case 8: PCT_BIN_CASE(float, 8);
      case 16: PCT_BIN_CASE(float, 16);
      case 32: PCT_BIN_CASE(float, 32);
    }
  } else {
    switch (lanes) {
      case 4: PCT_BIN_CASE(__nv_bfloat16, 4);
      case 8: PCT_BIN_CASE(__nv_bfloat16, 8);
      case 16: PCT_BIN_CASE(__nv_bfloat16, 16);
      case 32: PCT_BIN_CASE(__nv_bfloat16, 32);
    }
  }
#undef PCT_BIN_CASE
  return PCT_MSDA_E_UNSUPPORTED;
}

}  // namespace pctl
