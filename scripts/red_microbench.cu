// red_microbench.cu -- how fast can B200's L2 absorb vector reductions?  (tuning aid, GPU box)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_microbench red_microbench.cu
// Every warp issues RED.128 (red.global.add.v4.f32) where each group of 8 lanes covers one full
// 128-byte line, lines chosen pseudo-randomly inside a `region` of the buffer.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// mode 0: v4.f32 full lines (8 lanes / line); mode 1: scalar f32 atomics, 32 lanes / line;
// mode 2: bf16x8 (16 B per lane, 4 lanes cover 64 B = "bf16 line of 32 channels");
// mode 3: v4.f32 but each lane's 16 B lands in a different half-sector pattern (lane*32 B stride)
// mode 4: v2.f32 (8 B per lane, 16 lanes / line)
template <int MODE>
__global__ void red_kernel(float* buf, unsigned n_lines, int iters, unsigned seed) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      const unsigned grp = lane >> 3, sub = lane & 7;
      const unsigned line = hash32(seed + (gwarp * iters + it) * 4u + grp) % n_lines;
      float* p = buf + (size_t)line * 32 + sub * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
    } else if (MODE == 1) {
      const unsigned line = hash32(seed + (gwarp * iters + it)) % n_lines;
      atomicAdd(buf + (size_t)line * 32 + lane, 1.0f);
    } else if (MODE == 2) {
      const unsigned grp = lane >> 2, sub = lane & 3;
      const unsigned line = hash32(seed + (gwarp * iters + it) * 8u + grp) % (n_lines * 2);
      __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(buf) + (size_t)line * 32 + sub * 8;
      unsigned one2 = 0x3f803f80u;
      asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(one2) : "memory");
    } else if (MODE == 3) {
      const unsigned grp = lane >> 2, sub = lane & 3;
      const unsigned line = hash32(seed + (gwarp * iters + it) * 8u + grp) % n_lines;
      float* p = buf + (size_t)line * 32 + sub * 8;   // 16 B at 32 B stride: half sectors
      asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
    } else if (MODE == 5) {   // plain vector stores, same random-line pattern (L2 write path)
      const unsigned grp = lane >> 3, sub = lane & 7;
      const unsigned line = hash32(seed + (gwarp * iters + it) * 4u + grp) % n_lines;
      float* p = buf + (size_t)line * 32 + sub * 4;
      asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
    } else if (MODE == 6) {   // plain vector loads that bypass L1 (L2 read path)
      const unsigned grp = lane >> 3, sub = lane & 7;
      const unsigned line = hash32(seed + (gwarp * iters + it) * 4u + grp) % n_lines;
      const float* p = buf + (size_t)line * 32 + sub * 4;
      float a, b, c, d;
      asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(p));
      if (a == 123456.f) buf[0] = b + c + d;
    } else if (MODE == 7) {   // TMA bulk reduction: one elected lane adds a 512-B smem row block
      __shared__ __align__(128) float srow[32][128];   // per warp 512 B
      const unsigned warp_in_cta = threadIdx.x >> 5;
      for (int i = lane; i < 128; i += 32) srow[warp_in_cta][i] = 1.0f;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        for (int g = 0; g < 4; ++g) {
          const unsigned line = hash32(seed + (gwarp * iters + it) * 4u + g) % n_lines;
          float* p = buf + (size_t)line * 32;
          unsigned sa = (unsigned)__cvta_generic_to_shared(&srow[warp_in_cta][g * 32]);
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;"
                       ::"l"(p), "r"(sa) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if ((it & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
    } else {
      const unsigned grp = lane >> 4, sub = lane & 15;
      const unsigned line = hash32(seed + (gwarp * iters + it) * 2u + grp) % n_lines;
      float* p = buf + (size_t)line * 32 + sub * 2;
      asm volatile("red.global.add.v2.f32 [%0], {%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
    }
  }
}

template <int MODE>
void run(const char* name, float* buf, size_t region_bytes, int blocks, int threads, int iters,
         double bytes_per_warp_iter) {
  unsigned n_lines = (unsigned)(region_bytes / 128);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  red_kernel<MODE><<<blocks, threads>>>(buf, n_lines, iters, 1u);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    red_kernel<MODE><<<blocks, threads>>>(buf, n_lines, iters, 77u + r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double warps = (double)blocks * threads / 32;
  double bytes = warps * iters * bytes_per_warp_iter;
  printf("%-34s region %6.1f MB: %8.1f us  %7.1f GB/s payload  %7.1f sectors/us\n", name,
         region_bytes / 1e6, best * 1e3, bytes / best / 1e6, bytes / 32 / (best * 1e3));
}

int main() {
  size_t cap = 512ull << 20;
  float* buf;
  cudaMalloc(&buf, cap);
  cudaMemset(buf, 0, cap);
  int sms = 148;
  const size_t regions[] = {22ull << 20, 88ull << 20};
  for (size_t region : regions) {
    run<0>("v4.f32 full line (8 lanes/line)", buf, region, sms * 2, 512, 2000, 512);
    run<3>("v4.f32 half sectors", buf, region, sms * 2, 512, 2000, 512);
    run<4>("v2.f32 (16 lanes/line)", buf, region, sms * 2, 512, 2000, 256);
    run<1>("scalar f32 atomicAdd (32 lanes/line)", buf, region, sms * 2, 512, 2000, 128);
    run<2>("bf16x8 (4 lanes / 64 B)", buf, region, sms * 2, 512, 2000, 512);
    run<5>("st.global.v4 full line", buf, region, sms * 2, 512, 2000, 512);
    run<6>("ld.global.cg.v4 full line", buf, region, sms * 2, 512, 2000, 512);
    run<7>("TMA cp.reduce.async.bulk 128 B rows", buf, region, sms * 2, 512, 500, 512);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
