// msda_bwd_binned.cuh -- cell-binned backward of MSDeformAttn for sm_100a.
//
// Replaces the reference's seven col2im kernels (OPS/src/cuda/ms_deform_im2col_cuda.cuh:306-925)
// and their scatter of 4 scalar atomics per tap and channel (cuh:124-158).  The arithmetic is the
// reference's (cuh:93-164); the organisation is not:
//
// The per-tap backward (msda_bwd_vec_kernel) sends every bilinear corner of every sample to L2 as
// its own 128-byte reduction and re-reads the four value rows of every sample: 8 row transfers
// per sample through the SM's L1 data pipe and 16 L2 reduction sectors per sample, and the L2
// reduction path (measured ceiling 199 sectors/ns) bounds the kernel.  But the samples of
// neighbouring queries fall into the SAME bilinear cells -- a 16x16 level receives ~80 samples
// per cell and (image, head) -- so here grad_value is aggregated per CELL on chip before any
// global atomic:
//
//   task  = (image b, head m, chunk of QC consecutive queries); one persistent CTA per SM pulls
//           tasks from a global counter.
//   B1/B2/B3  counting sort of the chunk's QC*L*P samples by bilinear cell (level, y0, x0) in
//           shared memory: one shared-memory integer atomic per sample (ATOMS.ADD is native for
//           32-bit integers), a CTA-wide scan, and a 16-byte record {cell geometry | query, lw,
//           lh, attn} per in-range sample.  The scan also lists the non-empty cells in the order
//           of their record counts (a 64-bin histogram).  The chunk's grad_out rows are staged in
//           shared memory with cp.async meanwhile.
//   sweep  in ROUNDS: a round gives each group of LPP lanes (a lane owns 4 channels) one cell, the
//           groups of a warp taking neighbours of the count-ordered list, i.e. cells with the same
//           number of records.  All groups load the four corner value rows of their cells
//           together, walk their records together -- per record one grad_out row from shared
//           memory, four dot products (-> grad_attn, grad_loc, cuh:160-163) and four
//           multiply-adds into fp32 running sums in registers, no branch -- and then send the
//           sums to L2 (red.global.add.v4.f32) together.  Value rows and running sums live in the
//           BILINEAR BASIS (1, lw, lh, lw*lh) instead of per corner, so a record needs no corner
//           weights; the change of basis is paid once per cell.
//           Per-sample cost drops from 8 row transfers + 16 L2 sectors to 1 shared-memory row
//           read, with 8 row transfers + 16 sectors per CELL.
//   final  grad_sampling_loc / grad_attn_weight are written coalesced in sample order (the sweep
//           left them in the records; out-of-range samples get exact zeros as in the reference).
//
// Everything is fp32 arithmetic with fp32 accumulation (bf16 values accumulate into the caller's
// fp32 workspace exactly like the per-tap kernel).  Summation order differs from the reference
// (as any atomic scatter does), and the corner sums are formed as differences of basis sums
// (error of the order of 1e-7 of the cell's largest sum); tests hold the result to 1e-4.
#pragma once

#include "msda_kernels.cuh"

namespace pct {

#ifndef PCT_BIN_THREADS
#define PCT_BIN_THREADS 1024
#endif
#ifndef PCT_BIN_CTAS
#define PCT_BIN_CTAS 1       // resident CTAs per SM (each gets 1/PCT_BIN_CTAS of the shared memory)
#endif
constexpr int kBinThreads = PCT_BIN_THREADS;
constexpr int kBinCtasPerSm = PCT_BIN_CTAS;
constexpr int kBinMaxLevels = 8;
constexpr int kBinMaxSamples = 64;    // L*P
constexpr int kBinMaxQc = 1024;       // queries per chunk: 10 bits of the record word
constexpr int kBinMaxCells = 65535;   // 16 bits of the record word
constexpr unsigned kBinNoPos = 0xffffu;
#ifndef PCT_BIN_UB
#define PCT_BIN_UB 4        // samples per thread and batch of the binning passes (loads in flight)
#endif
#ifndef PCT_BIN_DYN_TAIL
#define PCT_BIN_DYN_TAIL 4
#endif
constexpr int kBinDynTail = PCT_BIN_DYN_TAIL;   // rounds per warp handed out dynamically at the end of a sweep

struct __align__(16) BinLevel {   // read as two 16-byte words
  int cell_base;                // cells of level l: cell_base + (y0+1)*(W+1) + (x0+1), y0 in [-1,H-1]
  unsigned magic;               // ceil(2^32 / wp1): __umulhi(c, magic) == c / wp1 for c < 2^16
  int wp1;                      // W + 1
  int W;
  int H, start;
  float Hf, Wf;
};

struct BinLaunch {
  int smem_bytes;   // dynamic shared memory of the launch
  int qc_cap;       // upper limit for the chunk size (0 = whatever fits)
  int slot;         // g_sched ring slot: [0] task counter, [1] CTA-exit counter
};

// Shared-memory bytes a chunk of one query needs: records + grad_out row + position of each sample.
// (fused: + the pair's softmax maximum, 1/normaliser and sum_t a_t g_t)
__host__ __device__ inline int bin_bytes_per_query(int LP, int row_bytes, bool fused = false) {
  return LP * 16 + row_bytes + LP * 2 + (fused ? 12 : 0);
}
// Shared-memory bytes that depend on the number of cells NC only: the counters / bin offsets and
// the list of non-empty cells in the order of the sweep.
__host__ __device__ inline long long bin_bytes_fixed(long long NC) {
  return ((NC + 1 + 3) & ~3ll) * 4 + ((NC * 2 + 15) & ~15ll) + 16;
}

__device__ __forceinline__ float bin_coord(float loc, float size) {
  // cuh:290-291: `loc * size - 0.5` -- float product (rounded), then the subtraction in double
  // rounded once to float, which is exactly one fp32 subtraction of the rounded product.
  return __fsub_rn(__fmul_rn(loc, size), 0.5f);
}

// transposed xor-shuffle reduction of three values over the LPP lanes of a group (LPP >= 4):
// returns, in lane sub, the group sum of  a (sub&3 == 0),  y (== 1),  x (== 2),  0 (== 3).
template <int LPP>
__device__ __forceinline__ float bin_reduce3(float a, float x, float y, int sub) {
  const bool b0 = sub & 1, b1 = sub & 2;
  const float send0 = b0 ? a : y, send1 = b0 ? x : 0.f;
  float keep0 = b0 ? y : a, keep1 = b0 ? 0.f : x;
  keep0 += __shfl_xor_sync(0xffffffffu, send0, 1);
  keep1 += __shfl_xor_sync(0xffffffffu, send1, 1);
  const float send = b1 ? keep0 : keep1;
  float k = b1 ? keep1 : keep0;
  k += __shfl_xor_sync(0xffffffffu, send, 2);
#pragma unroll
  for (int w = 4; w < LPP; w <<= 1) k += __shfl_xor_sync(0xffffffffu, k, w);
  return k;
}

// ---- packed fp32 pairs (fma.rn.f32x2 = FFMA2 on sm_100a) and raw shared-window accesses ----
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float sum2(u64 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return lo + hi;
}
#ifndef PCT_BIN_FMA2
#define PCT_BIN_FMA2 1
#endif
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
#if PCT_BIN_FMA2
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
#else
  float a0, a1, b0, b1, c0, c1;
  unpack2(a, a0, a1); unpack2(b, b0, b1); unpack2(c, c0, c1);
  return pack2(fmaf(a0, b0, c0), fmaf(a1, b1, c1));
#endif
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
#if PCT_BIN_FMA2
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
#else
  float a0, a1, b0, b1;
  unpack2(a, a0, a1); unpack2(b, b0, b1);
  return pack2(a0 * b0, a1 * b1);
#endif
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 pack2u(unsigned lo, unsigned hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ unsigned lds_u32(unsigned addr) {
  unsigned r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint4 lds_v4(unsigned addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint2 lds_v2(unsigned addr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
#ifndef PCT_BIN_EXPERIMENT
#define PCT_BIN_EXPERIMENT 0   // timing experiments only (wrong results): 1 = never change cell, 2 = no sweep
#endif

// one lane's 4 channels of a value row as two channel pairs; zeros when `keep` is false (the load
// is predicated off, so nothing waits on it)
template <typename T>
struct BinVal;
template <>
struct BinVal<float> {
  static __device__ __forceinline__ void load(const float* p, unsigned keep, u64& v01, u64& v23) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.b64 %0, 0;\n\tmov.b64 %1, 0;\n\t"
        "@p ld.global.nc.v2.u64 {%0, %1}, [%3];\n\t}"
        : "=&l"(v01), "=&l"(v23)
        : "r"(keep), "l"(p));
  }
  static __device__ __forceinline__ void load_all(const float* p, u64& v01, u64& v23) {
    asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(v01), "=l"(v23) : "l"(p));
  }
};
template <>
struct BinVal<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, unsigned keep, u64& v01, u64& v23) {
    unsigned x, y;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\t"
        "@p ld.global.nc.v2.u32 {%0, %1}, [%3];\n\t}"
        : "=&r"(x), "=&r"(y)
        : "r"(keep), "l"(p));
    v01 = pack2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
    v23 = pack2(__uint_as_float(y << 16), __uint_as_float(y & 0xffff0000u));
  }
  static __device__ __forceinline__ void load_all(const __nv_bfloat16* p, u64& v01, u64& v23) {
    unsigned x, y;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "l"(p));
    v01 = pack2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
    v23 = pack2(__uint_as_float(y << 16), __uint_as_float(y & 0xffff0000u));
  }
};

// one lane's 4 channels of a staged grad_out row as two channel pairs
template <typename T>
struct BinRowS;
template <>
struct BinRowS<float> {
  static __device__ __forceinline__ void load(unsigned addr, u64& g01, u64& g23) {
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(g01), "=l"(g23) : "r"(addr));
  }
};
template <>
struct BinRowS<__nv_bfloat16> {
  static __device__ __forceinline__ void load(unsigned addr, u64& g01, u64& g23) {
    unsigned x, y;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(addr));
    g01 = pack2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
    g23 = pack2(__uint_as_float(y << 16), __uint_as_float(y & 0xffff0000u));
  }
};


#ifdef PCT_MSDA_BOUNDS_CHECK
#define PCT_BIN_ASSERT(c) do { if (!(c)) __trap(); } while (0)
#else
#define PCT_BIN_ASSERT(c) do { } while (0)
#endif

// FUSED (same contract as the FUSED per-tap kernels in msda_kernels.cuh):
// `loc` holds the module's raw sampling offsets and `attn` its raw attention logits; the binning
// passes evaluate loc = ref + off/(W,H) (or the box form) and the softmax over the pair's L*P
// logits, and the final pass writes d/d(offsets) and the softmax backward a_s (g_s - sum_t a_t g_t)
// instead of grad_sampling_loc / grad_attn_weight.
template <typename T, int LPP, bool FUSED>
__global__ void __launch_bounds__(kBinThreads, kBinCtasPerSm)
msda_bwd_binned_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                       const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                       const float* __restrict__ attn, const T* __restrict__ grad_out,
                       float* __restrict__ gacc, float* __restrict__ grad_loc,
                       float* __restrict__ grad_attn, int N, int S, int M, int L, int Lq, int P,
                       BinLaunch bl, FusedRef fr) {
  constexpr int VE = 4;
  constexpr int VB = VE * (int)sizeof(T);
  constexpr int D = LPP * VE;
  constexpr int ROWB = D * (int)sizeof(T);     // bytes of one (query, head) grad_out row
  constexpr int PPR = ROWB / 16;               // 16-byte pieces per row
  static_assert(LPP >= 4 && LPP <= 32 && ROWB % 16 == 0, "binned backward: 4..32 lanes of 4 channels");

  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ BinLevel s_lvl[kBinMaxLevels];
  __shared__ unsigned char s_slvl[kBinMaxSamples];   // level of sample s
  __shared__ unsigned int s_wsum[32];
  __shared__ int s_nc, s_task[2];                    // this task and the next one (fetched one ahead)
  __shared__ unsigned int s_total;                   // (non-empty cells << 16) | records
  __shared__ unsigned int s_hist[64];                // non-empty cells per record count (descending)
  __shared__ unsigned int s_round;                   // next round of the dynamic tail of the sweep
  __shared__ __align__(16) unsigned int s_zero[16];   // zeros the compiler cannot see through

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARPS = kBinThreads / 32;
  const int LP = L * P;
  const int MD = M * D;

  if (tid == 0) {
    int cb = 0;
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      BinLevel lv;
      lv.H = H; lv.W = W; lv.start = (int)lsi[l]; lv.cell_base = cb;
      lv.wp1 = W + 1;
      lv.magic = (unsigned)((0x100000000ull + (unsigned long long)W) / (unsigned long long)(W + 1));
      lv.Hf = (float)H; lv.Wf = (float)W;
      s_lvl[l] = lv;
      cb += (H > 0 && W > 0) ? (H + 1) * (W + 1) : 0;
    }
    s_nc = cb;
  }
  for (int s = tid; s < LP; s += kBinThreads) s_slvl[s] = (unsigned char)(s / P);
  if (tid < 16) s_zero[tid] = 0u;
  __syncthreads();

  // ---- chunk geometry (identical in every CTA: a function of the shapes only) ----
  const int NC = s_nc;
  const int cnt_words = (NC + 1 + 3) & ~3;
  const int per_q = bin_bytes_per_query(LP, ROWB, FUSED);
  int qc_max = (bl.smem_bytes - (int)bin_bytes_fixed(NC)) / per_q;
  if (qc_max > kBinMaxQc) qc_max = kBinMaxQc;
  if (qc_max > 65528 / LP) qc_max = 65528 / LP;   // 16-bit record positions, 0xffff = none
  if (bl.qc_cap > 0 && qc_max > bl.qc_cap) qc_max = bl.qc_cap;
  qc_max &= ~7;
  if (qc_max < 8 || NC > kBinMaxCells) __trap();   // the launcher's bound guarantees the fit
  const int nchunks = (Lq + qc_max - 1) / qc_max;
  const int QC = (((Lq + nchunks - 1) / nchunks) + 7) & ~7;   // balanced chunks, <= qc_max
  const int n_tasks = N * nchunks * M;

  // staged grad_out rows first (a compile-time shared address in the sweep), then the records
  unsigned char* const go_s = smem;
  uint4* const rec = reinterpret_cast<uint4*>(smem + (size_t)QC * ROWB);
  unsigned int* const cnt = reinterpret_cast<unsigned int*>(smem + (size_t)QC * ROWB + (size_t)QC * LP * 16);
  unsigned short* const posarr = reinterpret_cast<unsigned short*>(cnt + cnt_words);
  unsigned short* const klist = posarr + (size_t)QC * LP;   // (QC*LP*2 is a multiple of 16)
  // fused softmax state per query of the chunk: maximum logit, 1 / normaliser, sum_t a_t g_t
  float* const sm_max = reinterpret_cast<float*>(klist + ((NC + 7) & ~7));
  float* const sm_inv = sm_max + QC;
  float* const sm_dot = sm_inv + QC;

  const int sub = tid % LPP;
  unsigned int* const ctr = g_sched + 2 * bl.slot;

  // one sample of the binning passes: cell index (or -1 when the reference skips it), fractions
  auto locate = [&](float2 xy, int s, float& lw, float& lh, int& lvl) -> int {
    lvl = s_slvl[s];
    const int4 A = *reinterpret_cast<const int4*>(&s_lvl[lvl]);          // cell_base, magic, wp1, W
    const float2 F = *reinterpret_cast<const float2*>(&s_lvl[lvl].Hf);   // Hf, Wf
    const float y = bin_coord(xy.y, F.x), x = bin_coord(xy.x, F.y);
    if (!(y > -1.f && x > -1.f && y < F.x && x < F.y)) return -1;       // cuh:293 (NaN: outside)
    const float fy = floorf(y), fx = floorf(x);
    lh = y - fy; lw = x - fx;
    return A.x + ((int)fy + 1) * A.z + (int)fx + 1;
  };

  // the same plus the record's first word: the geometry of the sample's cell, worked out by ONE
  // thread per sample -- first pixel of the cell clamped into the map (15 bits), which of its four
  // corners exist (cuh:61-83; 4 bits), level (3 bits); the caller adds the query (10 bits)
  auto locate_full = [&](float2 xy, int s, float& lw, float& lh, unsigned& hdr) -> int {
    const int lvl = s_slvl[s];
    const int4 A = *reinterpret_cast<const int4*>(&s_lvl[lvl]);          // cell_base, magic, wp1, W
    const int4 Bq = *reinterpret_cast<const int4*>(&s_lvl[lvl].H);       // H, start, Hf, Wf
    const float Hf = __int_as_float(Bq.z), Wf = __int_as_float(Bq.w);
    const float y = bin_coord(xy.y, Hf), x = bin_coord(xy.x, Wf);
    if (!(y > -1.f && x > -1.f && y < Hf && x < Wf)) return -1;         // cuh:293 (NaN: outside)
    const float fy = floorf(y), fx = floorf(x);
    lh = y - fy; lw = x - fx;
    const int yy = (int)fy + 1, xx = (int)fx + 1;                        // y0 + 1, x0 + 1
    const bool top = yy >= 1, bot = yy <= Bq.x - 1, lef = xx >= 1, rig = xx <= A.w - 1;
    const unsigned vm = (top && lef ? 1u : 0u) | (top && rig ? 2u : 0u) |
                        (bot && lef ? 4u : 0u) | (bot && rig ? 8u : 0u);
    const unsigned pix = (unsigned)(Bq.y + max(yy - 1, 0) * A.w + max(xx - 1, 0));
    PCT_BIN_ASSERT(pix < 32768u && vm != 0u);
    hdr = (pix << 17) | (vm << 13) | ((unsigned)lvl << 10);
    return A.x + yy * A.z + xx;
  };
  // tasks come from a global counter; the next index is fetched while the final pass of this task
  // runs (its round trip and a barrier per task used to be exposed) -- not earlier: a task reserved
  // long before it starts cannot go to an SM that finishes sooner
  if (tid == 0) s_task[0] = (int)atomicAdd(ctr, 1u);
  __syncthreads();
  for (int it = 0;; ++it) {
    const int t = s_task[it & 1];
    if (t >= n_tasks) break;
    const int m = t % M;
    const int tc = t / M;
    const int c = tc % nchunks, b = tc / nchunks;
    const int q0 = c * QC;
    const int nq = min(QC, Lq - q0);
    const int ntaps = nq * LP;
    // global index of sample (q0, m, 0); a query step is M*LP samples
    const long long samp0 = ((long long)((long long)b * Lq + q0) * M + m) * LP;
    const long long qstep = (long long)M * LP;
    const float2* const loc2 = reinterpret_cast<const float2*>(loc) + samp0;
    const float* const attn1 = attn + samp0;

    // ---- P0: clear the cell counters, start staging the chunk's grad_out rows ----
    for (int i = tid; i < cnt_words; i += kBinThreads) cnt[i] = 0u;
    if (tid < 64) s_hist[tid] = 0u;
    if (tid == 64) s_round = 0u;
    // reference points of the chunk's first query (fused): (L, ref_dim) floats per query
    const float* const ref0 = FUSED ? fr.ref + ((long long)b * Lq + q0) * L * fr.ref_dim : nullptr;
    if (FUSED) {
      // softmax over the L*P logits of every (query, head) pair of the chunk (mod.py:104): four
      // lanes per pair (a thread per pair left most of the CTA waiting at the next barrier behind
      // 2*L*P dependent strided loads, ncu: 13 % of the kernel's stall samples)
      const int n4 = nq * 4;
      for (int i = tid; i < ((n4 + 31) & ~31); i += kBinThreads) {
        const bool ok = i < n4;                      // whole groups of four: n4 is a multiple of 4
        const int ql = i >> 2, part = i & 3;
        const float* ap = attn1 + (ok ? ql : 0) * qstep;
        float mx = -INFINITY, es = 0.f;
        if (ok)
          for (int s = part; s < LP; s += 4) mx = fmaxf(mx, __ldg(ap + s));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        if (ok)
          for (int s = part; s < LP; s += 4) es += expf(__ldg(ap + s) - mx);
        es += __shfl_xor_sync(0xffffffffu, es, 1);
        es += __shfl_xor_sync(0xffffffffu, es, 2);
        if (ok && part == 0) {
          sm_max[ql] = mx;
          sm_inv[ql] = 1.f / es;
        }
      }
    }
    // sampling location of sample (ql, s) from what `loc` holds
    auto sample_xy = [&](float2 raw, int ql, int s) -> float2 {
      if (!FUSED) return raw;
      if (ql >= nq) return make_float2(-4.f, -4.f);   // padding of the last batch: outside every map
      const int lvl = s_slvl[s];
      const int2 hw = make_int2(s_lvl[lvl].H, s_lvl[lvl].W);
      return fused_location(ref0 + (long long)ql * L * fr.ref_dim, fr.ref_dim, lvl, raw, hw.x, hw.y, P);
    };
    {
      const T* gsrc = grad_out + ((long long)((long long)b * Lq + q0) * M + m) * D;
      const unsigned dst0 = (unsigned)__cvta_generic_to_shared(go_s);
      for (int j = tid; j < nq * PPR; j += kBinThreads) {
        const int row = j / PPR, part = j - row * PPR;
        const T* src = gsrc + (long long)row * MD + part * (16 / (int)sizeof(T));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (unsigned)j * 16u), "l"(src)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __syncthreads();

    // ---- B1: count the samples of every cell (the loads of a batch first, then the atomics) ----
    constexpr int UB = PCT_BIN_UB;
    const int dq = kBinThreads / LP, ds = kBinThreads - dq * LP;
    {
      int ql = tid / LP, s = tid - ql * LP;
      for (int k0 = tid; k0 < ntaps; k0 += UB * kBinThreads) {
        float2 xy[UB];
        int ss[UB], qq[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          ss[u] = s; qq[u] = ql;
          xy[u] = (k0 + u * kBinThreads < ntaps) ? __ldg(loc2 + ql * qstep + s)
                                                : make_float2(-4.f, -4.f);   // outside every map
          ql += dq; s += ds;
          if (s >= LP) { s -= LP; ++ql; }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          float lw, lh;
          int lvl;
          const int cell = locate(sample_xy(xy[u], qq[u], ss[u]), ss[u], lw, lh, lvl);
          if (cell >= 0) {
            PCT_BIN_ASSERT(cell < NC);
            atomicAdd(cnt + cell, 1u);
          }
        }
      }
    }
    __syncthreads();

    // ---- B2: exclusive scan of (non-empty << 16 | count) over the cells, in place ----
    {
      const int per = (NC + kBinThreads - 1) / kBinThreads;
      const int i0 = min(tid * per, NC), i1 = min(i0 + per, NC);
      unsigned sum = 0;
      for (int i = i0; i < i1; ++i) {
        const unsigned cv = cnt[i];
        sum += cv + (cv ? 0x10000u : 0u);
        if (cv) atomicAdd(&s_hist[63u - min(cv, 63u)], 1u);     // bin 0 = the fullest cells
      }
      unsigned incl = sum;
#pragma unroll
      for (int w = 1; w < 32; w <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, incl, w);
        if (lane >= w) incl += o;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        const unsigned v = lane < NWARPS ? s_wsum[lane] : 0u;
        unsigned inc2 = v;
#pragma unroll
        for (int w = 1; w < 32; w <<= 1) {
          const unsigned o = __shfl_up_sync(0xffffffffu, inc2, w);
          if (lane >= w) inc2 += o;
        }
        if (lane < NWARPS) s_wsum[lane] = inc2 - v;   // exclusive prefix of the warp totals
        if (lane == 31) { s_total = inc2; cnt[NC] = inc2; }
      }
      if (warp == 1) {   // exclusive scan of the 64 histogram bins (two per lane)
        const unsigned h0 = s_hist[2 * lane], h1 = s_hist[2 * lane + 1];
        unsigned inc2 = h0 + h1;
#pragma unroll
        for (int w = 1; w < 32; w <<= 1) {
          const unsigned o = __shfl_up_sync(0xffffffffu, inc2, w);
          if (lane >= w) inc2 += o;
        }
        s_hist[2 * lane] = inc2 - h0 - h1;
        s_hist[2 * lane + 1] = inc2 - h1;
      }
      __syncthreads();
      unsigned base = s_wsum[warp] + incl - sum;
      for (int i = i0; i < i1; ++i) {
        const unsigned cv = cnt[i];
        cnt[i] = base;
        base += cv + (cv ? 0x10000u : 0u);
        // the sweep visits the non-empty cells in the order of their record counts
        if (cv) klist[atomicAdd(&s_hist[63u - min(cv, 63u)], 1u)] = (unsigned short)i;
      }
    }
    __syncthreads();


    // ---- B3: write the records into their bins ----
    {
      int ql = tid / LP, s = tid - ql * LP;
      for (int k0 = tid; k0 < ntaps; k0 += UB * kBinThreads) {
        float2 xy[UB];
        float aw[UB];
        int ss[UB], qq[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          ss[u] = s; qq[u] = ql;
          const bool ok = k0 + u * kBinThreads < ntaps;
          xy[u] = ok ? __ldg(loc2 + ql * qstep + s) : make_float2(-4.f, -4.f);
          aw[u] = ok ? __ldg(attn1 + ql * qstep + s) : 0.f;
          ql += dq; s += ds;
          if (s >= LP) { s -= LP; ++ql; }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int k = k0 + u * kBinThreads;
          if (k < ntaps) {
            float lw = 0.f, lh = 0.f;
            unsigned hdr = 0u;
            const int cell = locate_full(sample_xy(xy[u], qq[u], ss[u]), ss[u], lw, lh, hdr);
            unsigned pos = kBinNoPos;
            if (cell >= 0) {
              pos = atomicAdd(cnt + cell, 1u) & 0xffffu;
              PCT_BIN_ASSERT(pos < (unsigned)ntaps);
              const float a = FUSED ? expf(aw[u] - sm_max[qq[u]]) * sm_inv[qq[u]] : aw[u];
              rec[pos] = make_uint4(hdr | (unsigned)qq[u], __float_as_uint(lw), __float_as_uint(lh),
                                    __float_as_uint(a));
            }
            posarr[k] = (unsigned short)pos;
          }
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // launched as a programmatic dependent of the zero-fill of gacc: nothing above touches gacc,
    // the reductions below do (a no-op after the first task and for a plain launch)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();

    // ---- sweep: lock-step rounds over the non-empty cells, fullest cells first ----
    // Bilinear-basis form.  With value(lw, lh) = u0 + lw u1 + lh u2 + lw lh u3  (u0 = v00,
    // u1 = v01 - v00, u2 = v10 - v00, u3 = v11 - v10 - v01 + v00) the three per-sample results are
    //   val . grad_out = e0 + lw e1 + lh e2 + lw lh e3,   grad_w_weight = e1 + lh e3,
    //   grad_h_weight  = e2 + lw e3                        (e_k = <grad_out, u_k>; cuh:128-161)
    // and the four corner sums of grad_value are differences of the four running sums
    //   A0 = sum a g, A1 = sum a lw g, A2 = sum a lh g, A3 = sum a lw lh g   (cuh:118-158),
    // so the per-record work needs no corner weights at all; the change of basis is paid once per
    // CELL, and the cell's geometry comes ready-made in the record's first word.
    // A round gives each of the warp's 32/LPP lane groups ONE cell: all groups load their four
    // value rows together, walk their records together (cells of a round hold the same number of
    // records up to the 63-record histogram limit, so hardly a lane idles) and flush together --
    // no branch inside the record loop.
    {
      constexpr int GPW = 32 / LPP;                  // lane groups (cells) per warp and round
      // (the launcher guarantees N*S*M*D < 2^31: one 32-bit element offset serves value and gacc)
      const int img_off = (b * S * M + m) * D + sub * VE;
      unsigned go_sa = (unsigned)__cvta_generic_to_shared(go_s) + (unsigned)(sub * VB);
      unsigned sub4 = (unsigned)sub * 4u;
      asm volatile("" : "+r"(go_sa), "+r"(sub4));   // kept in registers, not re-derived per record
      const unsigned rec_sa = (unsigned)__cvta_generic_to_shared(rec);
      const unsigned lvl_sa = (unsigned)__cvta_generic_to_shared(s_lvl);
      const unsigned zero_sa = (unsigned)__cvta_generic_to_shared(s_zero);
      const int ncell = (int)(s_total >> 16);
      const int nrounds = (ncell + GPW - 1) / GPW;
#if PCT_BIN_EXPERIMENT == 2
      if (false)
#endif
      // rounds go to the warps in snake order (w, 63-w, 64+w, 127-w, ...): the fullest cells come
      // first, and a plain interleave gave warp 0 the longest round of every 32 (11 % of the
      // kernel's stall samples sat at the barrier behind the sweep)
      // The last kBinDynTail * NWARPS rounds (cells with one or two records) are handed out through
      // a shared counter instead: warps that fell behind take fewer of them.
      const int n_static = max(0, (nrounds / NWARPS - kBinDynTail) * NWARPS);
      for (int j = 0;; ++j) {
        int rd;
        if (j * NWARPS < n_static) {
          rd = j * NWARPS + ((j & 1) ? NWARPS - 1 - warp : warp);
        } else {
          rd = 0;
          if (lane == 0) rd = n_static + (int)atomicAdd(&s_round, 1u);
          rd = __shfl_sync(0xffffffffu, rd, 0);
          if (rd >= nrounds) break;
        }
        const int ci = rd * GPW + (lane / LPP);
        unsigned rp = 0u, rend = 0u;                 // shared addresses of this group's records
        if (ci < ncell) {
          const unsigned c = klist[ci];
          rp = rec_sa + (c ? (cnt[c - 1] & 0xffffu) : 0u) * 16u;   // B3 left the END of every bin in cnt
          rend = rec_sa + (cnt[c] & 0xffffu) * 16u;
          PCT_BIN_ASSERT(rp < rend);
        }
        const bool have = rp < rend;
        // ---- open ----
        // element offsets of the four corner rows of cell `key`, clamped into the map (a corner the
        // map does not have shares the row of its neighbour and is never touched); key 0 = no cell
        auto geometry = [&](unsigned key, int& o00, int& o01, int& o10, int& o11) {
          const unsigned vm = key >> 3;
          const int Wl = (int)lds_u32(lvl_sa + (key & 7u) * 32u + 12u);
          const int xs = ((vm & (vm >> 1)) & 5u) ? MD : 0;          // left and right column both in the map
          const int ys = ((vm & (vm >> 2)) & 3u) ? Wl * MD : 0;     // top and bottom row both in the map
          o00 = img_off + (int)(key >> 7) * MD;
          o01 = o00 + xs;
          o10 = o00 + ys;
          o11 = o10 + xs;
        };
        const unsigned key = have ? lds_u32(rp) >> 10 : 0u;    // (pixel, corner mask, level)
        u64 u0[2], u1[2], u2[2], u3[2];    // value rows of the cell in the bilinear basis
        u64 A0[2], A1[2], A2[2], A3[2];    // running sums of grad_value in the same basis
        {
          u64 v01[2], v10[2], v11[2];
          const unsigned vm = (key >> 3) & 15u;                // bit k: corner 00, 01, 10, 11 exists
          int o00, o01, o10, o11;
          geometry(key, o00, o01, o10, o11);
          PCT_CHECK_TAP(o00 - img_off, 0, 0, (m * D + sub * VE), VE, (long long)S * MD);
          PCT_CHECK_TAP(o11 - img_off, 0, 0, (m * D + sub * VE), VE, (long long)S * MD);
          // corners the reference skips read as 0 (cuh:61-83): zero-initialised, load predicated
          BinVal<T>::load(value + o00, vm & 1u, u0[0], u0[1]);
          BinVal<T>::load(value + o01, vm & 2u, v01[0], v01[1]);
          BinVal<T>::load(value + o10, vm & 4u, v10[0], v10[1]);
          BinVal<T>::load(value + o11, vm & 8u, v11[0], v11[1]);
          // the running sums start from zeros READ from shared memory: as constants ptxas routes
          // them through 32 uniform-register moves
          const uint4 z0 = lds_v4(zero_sa), z1 = lds_v4(zero_sa + 16u), z2 = lds_v4(zero_sa + 32u), z3 = lds_v4(zero_sa + 48u);
          A0[0] = pack2u(z0.x, z0.y); A0[1] = pack2u(z0.z, z0.w);
          A1[0] = pack2u(z1.x, z1.y); A1[1] = pack2u(z1.z, z1.w);
          A2[0] = pack2u(z2.x, z2.y); A2[1] = pack2u(z2.z, z2.w);
          A3[0] = pack2u(z3.x, z3.y); A3[1] = pack2u(z3.z, z3.w);
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) {
            u1[k2] = sub2(v01[k2], u0[k2]);
            u2[k2] = sub2(v10[k2], u0[k2]);
            u3[k2] = sub2(sub2(v11[k2], v10[k2]), u1[k2]);
          }
        }
        // ---- the records of the round (warp-uniform trip count: shuffles inside) ----
#if PCT_BIN_EXPERIMENT == 1
        for (int i = 0; i > 0; --i, rp += 16u) {
#else
        for (int i = (int)__reduce_max_sync(0xffffffffu, (rend - rp) >> 4); i > 0; --i, rp += 16u) {
#endif
          float pa = 0.f, px = 0.f, py = 0.f;
          const bool active = rp < rend;
          if (active) {
            const uint4 rc = lds_v4(rp);
            u64 g01, g23;
            BinRowS<T>::load(go_sa + (rc.x & 0x3ffu) * ROWB, g01, g23);
            const float lw = __uint_as_float(rc.y), lh = __uint_as_float(rc.z);
            const float a = __uint_as_float(rc.w);
            const u64 LW = pack2(lw, lw), LH = pack2(lh, lh), AA = pack2(a, a);
            const u64 T1 = mul2(AA, LW), T2 = mul2(AA, LH), T3 = mul2(T1, LH);
            A0[0] = fma2(AA, g01, A0[0]); A0[1] = fma2(AA, g23, A0[1]);
            A1[0] = fma2(T1, g01, A1[0]); A1[1] = fma2(T1, g23, A1[1]);
            A2[0] = fma2(T2, g01, A2[0]); A2[1] = fma2(T2, g23, A2[1]);
            A3[0] = fma2(T3, g01, A3[0]); A3[1] = fma2(T3, g23, A3[1]);
            const u64 e3 = fma2(g23, u3[1], mul2(g01, u3[0]));
            const u64 X = fma2(LH, e3, fma2(g23, u1[1], mul2(g01, u1[0])));   // e1 + lh e3
            const u64 e2 = fma2(g23, u2[1], mul2(g01, u2[0]));
            const u64 Y = fma2(LW, e3, e2);                                   // e2 + lw e3
            const u64 e0 = fma2(g23, u0[1], mul2(g01, u0[0]));
            px = sum2(X);                                     // grad_w_weight (cuh:128-156)
            py = sum2(Y);                                     // grad_h_weight
            pa = sum2(fma2(LW, X, fma2(LH, e2, e0)));         // cuh:161: e0 + lw e1 + lh e2 + lw lh e3
          }
          // group sums: lane sub 0 holds pa, 1 px, 2 py; they replace words 0..2 of the record (its
          // header / lw / lh are used up; word 3 keeps attn for the scaling in the final pass)
          const float red = bin_reduce3<LPP>(pa, py, px, sub);
          if (active && sub4 < 12u) sts_f32(rp + sub4, red);
        }
        // ---- flush: corner sums out of the basis sums, one L2 reduction per existing corner ----
        {
          auto red_row = [&](int off, u64 lo, u64 hi) {
            float x0, x1, x2, x3;
            unpack2(lo, x0, x1);
            unpack2(hi, x2, x3);
            red_add_v4(gacc + off, x0, x1, x2, x3);
          };
          const u64 c01a = sub2(A1[0], A3[0]), c01b = sub2(A1[1], A3[1]);
          const u64 c10a = sub2(A2[0], A3[0]), c10b = sub2(A2[1], A3[1]);
          const u64 c00a = sub2(sub2(A0[0], A1[0]), c10a), c00b = sub2(sub2(A0[1], A1[1]), c10b);
          const unsigned vm = (key >> 3) & 15u;
          int o00, o01, o10, o11;
          geometry(key, o00, o01, o10, o11);
          if (vm & 1u) red_row(o00, c00a, c00b);
          if (vm & 2u) red_row(o01, c01a, c01b);
          if (vm & 4u) red_row(o10, c10a, c10b);
          if (vm & 8u) red_row(o11, A3[0], A3[1]);
        }
      }
    }
    __syncthreads();

    // ---- final: coalesced grad_sampling_loc / grad_attn_weight in sample order ----
    if (tid == 0) s_task[(it + 1) & 1] = (int)atomicAdd(ctr, 1u);   // read after the barrier below
    if (FUSED) {
      // sum_t a_t g_t of every pair for the softmax backward (samples the reference skips have g = 0)
      // (four lanes per pair, as in the softmax pass above)
      const int n4 = nq * 4;
      for (int i = tid; i < ((n4 + 31) & ~31); i += kBinThreads) {
        const bool ok = i < n4;
        const int ql = i >> 2, part = i & 3;
        float dot = 0.f;
        if (ok)
          for (int s = part; s < LP; s += 4) {
            const unsigned pos = posarr[ql * LP + s];
            if (pos != kBinNoPos) {
              const uint4 rc = rec[pos];
              dot = fmaf(__uint_as_float(rc.w), __uint_as_float(rc.x), dot);
            }
          }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        if (ok && part == 0) sm_dot[ql] = dot;
      }
      __syncthreads();
    }
    {
      int ql = tid / LP, s = tid - ql * LP;
      for (int k2 = tid; k2 < ntaps; k2 += kBinThreads) {
        const unsigned pos = posarr[k2];
        float ga = 0.f, gx = 0.f, gy = 0.f;    // out-of-range samples: exact zeros (cu:126-128)
        float a = 0.f;
        const int lvl = s_slvl[s];
        if (pos != kBinNoPos) {
          const uint4 rc = rec[pos];           // {val . grad_out, grad_w_weight, grad_h_weight, attn}
          const float2 F = *reinterpret_cast<const float2*>(&s_lvl[lvl].Hf);
          a = __uint_as_float(rc.w);
          ga = __uint_as_float(rc.x);                       // cuh:161
          gx = F.y * __uint_as_float(rc.y) * a;             // cuh:162  W * grad_w_weight * attn
          gy = F.x * __uint_as_float(rc.z) * a;             // cuh:163  H * grad_h_weight * attn
        }
        const long long si = samp0 + ql * qstep + s;
        if (FUSED) {
          // d(location)/d(raw offset) (mod.py:106-112) and the softmax backward
          if (pos == kBinNoPos) a = expf(__ldg(attn + si) - sm_max[ql]) * sm_inv[ql];
          else if (fr.ref_dim == 2) {
            gx *= __fdiv_rn(1.f, (float)s_lvl[lvl].W);
            gy *= __fdiv_rn(1.f, (float)s_lvl[lvl].H);
          } else {
            const float4 r = __ldg(reinterpret_cast<const float4*>(ref0 + (long long)ql * L * 4) + lvl);
            gx *= __fdiv_rn(r.z * 0.5f, (float)P);
            gy *= __fdiv_rn(r.w * 0.5f, (float)P);
          }
          ga = a * (ga - sm_dot[ql]);
        }
        __stcs(grad_attn + si, ga);
        __stcs(reinterpret_cast<float2*>(grad_loc) + si, make_float2(gx, gy));
        ql += dq; s += ds;
        if (s >= LP) { s -= LP; ++ql; }
      }
    }
    __syncthreads();   // the next task reuses the shared arrays
  }

  // last CTA out resets the counters of this ring slot
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ctr + 1, 1u) == gridDim.x - 1) {
      ctr[0] = 0u;
      ctr[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace pct
