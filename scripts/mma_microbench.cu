// mma_microbench.cu -- what does one tcgen05.mma of the Linear kernel's shape cost on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_microbench mma_microbench.cu
// One thread per CTA issues a long chain of M=128 x N x K MMAs out of shared memory (contents are
// irrelevant), commits them to an mbarrier and times the whole chain with clock64().  Variants:
// kind::tf32 (K=8) vs kind::f16 (bf16, K=16); N = 256 / 128; operands from the same or from
// rotating shared-memory addresses; 1 or 2 CTAs per SM; K-major SWIZZLE_NONE descriptors as in
// pctrans_b200/csrc/linear_tc.cu.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// layout: 0 = SWIZZLE_NONE (8 x 16 B core matrices, lbo = next k chunk, sbo = next 8 rows),
// 6 / 4 / 2 = SWIZZLE_32B / 64B / 128B (rows of 32 / 64 / 128 B, sbo = 8 rows)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}

template <int KIND>  // 0 = tf32, 1 = bf16
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(64) mma_chain(long long* cycles, int n_mma, int N, int rotate, int a_stride,
                                               int b_stride, int layout, int n_acc) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x;
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(n_acc == 2 ? 512u : 256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 24576; i += 64) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 127);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t fmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t s0 = smem_u32(smem);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t ra = (rotate && n_acc != 4) ? (uint32_t)(i % rotate) * a_stride : 0u;
      const uint32_t rb = (rotate && n_acc != 3) ? (uint32_t)(i % rotate) * b_stride : 0u;
      const uint32_t lbo = layout == 0 ? 128u : 16u;
      const uint32_t sbo = layout == 0 ? 512u : layout == 6 ? 256u : layout == 4 ? 512u : 1024u;
      // n_acc = 2: consecutive MMAs go to two independent accumulators (two row tiles sharing B)
      mma<KIND>(tmem + (n_acc == 2 ? (i & 1) * 256u : 0u), make_desc(s0 + ra, lbo, sbo, layout),
                make_desc(s0 + 32768 + rb, lbo, sbo, layout), idesc, i >= (n_acc == 2 ? 2 : 1));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
    const long long t1 = clock64();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
      cycles[0] = t1 - t0;
      cycles[1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                          "r"(n_acc == 2 ? 512u : 256u) : "memory");
}

template <int KIND>
void run(const char* name, int ctas_per_sm, int N, int rotate, int layout = 0, int n_acc = 1) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 98304, n_mma = 3000;
  cudaFuncSetAttribute(mma_chain<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_chain<KIND><<<148 * ctas_per_sm, 64, smem>>>(d, n_mma, N, rotate, 4096, 8192, layout, n_acc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double per = (double)h[1] / n_mma;
  const double flop = 2.0 * 128 * N * (KIND == 0 ? 8 : 16);
  printf("%-5s acc=%d layout=%d N=%3d ctas/SM=%d rotate=%d : issue %.1f cyc/MMA, complete %.1f cyc/MMA -> %.0f FLOP/clk/SM per CTA (%s)\n",
         name, n_acc, layout, N, ctas_per_sm, rotate, (double)h[0] / n_mma, per, flop / per, cudaGetErrorString(e));
  cudaFree(d);
}


// ---- cta_group::2: the leader of a 2-CTA cluster issues M=256 MMAs; each CTA supplies its own 128
// rows of A and N/2 columns of B from its shared memory ------------------------------------------
__global__ void __launch_bounds__(64) mma_chain_pair(long long* cycles, int n_mma, int N, int rotate) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 24576; i += 64) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 127);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (16u << 24);
    const uint32_t s0 = smem_u32(smem);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t ra = rotate ? (uint32_t)(i % rotate) * 4096u : 0u;
      const uint32_t rb = rotate ? (uint32_t)(i % rotate) * 4096u : 0u;
      const uint32_t acc = i != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                   "l"(make_desc(s0 + ra, 128, 512)), "l"(make_desc(s0 + 32768 + rb, 128, 512)), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    const long long t1 = clock64();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
      cycles[0] = t1 - t0;
      cycles[1] = t2 - t0;
    }
  } else if (tid == 0) {
    uint32_t done = 0;   // the partner waits for the same (multicast) completion before tearing down
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

void run_pair(int ctas_per_sm, int N, int rotate) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 98304, n_mma = 3000;
  cudaFuncSetAttribute(mma_chain_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148 * ctas_per_sm);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr = {};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  cudaError_t e0 = cudaLaunchKernelEx(&cfg, mma_chain_pair, d, n_mma, N, rotate);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double per = (double)h[1] / n_mma;
  printf("pair  cta_group::2 M=256 N=%3d ctas/SM=%d rotate=%d : issue %.1f cyc/MMA, complete %.1f cyc/MMA -> %.0f FLOP/clk/SM per stream (%s, %s)\n",
         N, ctas_per_sm, rotate, (double)h[0] / n_mma, per, 2.0 * 128 * N * 8 / per, cudaGetErrorString(e0),
         cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int N : {256, 128})
      for (int rotate : {0, 4}) {
        run<0>("tf32", ctas, N, rotate);
        run<1>("bf16", ctas, N, rotate);
      }
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int layout : {6, 4, 2}) run<0>("tf32", ctas, 256, 4, layout);
  // one CTA per SM, two accumulators fed alternately (512 TMEM columns)
  for (int rotate : {0, 4}) run<0>("tf32", 1, 256, rotate, 0, 2);
  // only one operand changes between consecutive MMAs: is the unchanged one fetched again?
  printf("rotate A only / B only (n_acc field: 3 = A only, 4 = B only):\n");
  for (int ctas = 1; ctas <= 2; ++ctas) {
    run<0>("tf32", ctas, 256, 4, 0, 3);
    run<0>("tf32", ctas, 256, 4, 0, 4);
  }
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int rotate : {0, 4}) run_pair(ctas, 256, rotate);
  return 0;
}
