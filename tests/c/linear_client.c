/* Plain-C client of include/pct_linear.h (no Python, no torch): device buffers through the CUDA
 * runtime API, y = LayerNorm(x W^T + b + residual) and the weight / bias gradients checked against
 * double-precision loops on the host.  Built and run by tests/test_gpu_c_host.py:
 *   gcc linear_client.c -I include -I /usr/local/cuda/include -lpctrans_msda -lcudart -lm */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "pct_linear.h"
#include "pct_msda.h"

static uint64_t rng = 0x2545f4914f6cdd1dull;
static float frand(void) {
  rng = rng * 6364136223846793005ull + 1442695040888963407ull;
  return (float)((double)(rng >> 11) / 9007199254740992.0 * 2.0 - 1.0);
}

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (int)(call);                                                           \
    if (rc_ != 0) {                                                                  \
      printf("FAIL %s -> %d (%s)\n", #call, rc_, pct_msda_error_string(rc_));      \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

static float* to_device(const float* h, size_t n) {
  float* d = NULL;
  if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess) return NULL;
  if (h) cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice);
  return d;
}

int main(void) {
  enum { ROWS = 333, K = 96, N = 80 };
  float *x = malloc(4 * ROWS * K), *w = malloc(4 * N * K), *b = malloc(4 * N), *res = malloc(4 * ROWS * N),
        *gamma = malloc(4 * N), *beta = malloc(4 * N), *gy = malloc(4 * ROWS * N);
  for (int i = 0; i < ROWS * K; ++i) x[i] = frand();
  for (int i = 0; i < N * K; ++i) w[i] = frand() * 0.1f;
  for (int i = 0; i < N; ++i) b[i] = frand(), gamma[i] = 1.0f + 0.5f * frand(), beta[i] = frand();
  for (int i = 0; i < ROWS * N; ++i) res[i] = frand(), gy[i] = frand();

  float *dx = to_device(x, ROWS * K), *dw = to_device(w, N * K), *db = to_device(b, N),
        *dres = to_device(res, ROWS * N), *dgam = to_device(gamma, N), *dbet = to_device(beta, N),
        *dgy = to_device(gy, ROWS * N), *dy = to_device(NULL, ROWS * N), *dgw = to_device(NULL, N * K),
        *dgb = to_device(NULL, N);
  size_t packed_bytes = pct_linear_packed_bytes(K, N);
  void* dpacked = NULL;
  if (!packed_bytes || cudaMalloc(&dpacked, packed_bytes) != cudaSuccess || !dx || !dgb) {
    printf("FAIL allocation\n");
    return 1;
  }
  CHECK(pct_linear_pack_weight(dw, dpacked, K, N, NULL));
  CHECK(pct_linear_forward(dx, dpacked, db, dres, dgam, dbet, 1e-5f, dy, ROWS, K, N,
                           PCT_LINEAR_RESIDUAL | PCT_LINEAR_LAYERNORM, NULL));
  CHECK(pct_linear_wgrad(dgy, dx, dgw, dgb, ROWS, K, N, NULL));
  /* bad arguments are error codes, not crashes */
  if (pct_linear_forward(dx, dpacked, db, NULL, NULL, NULL, 0.f, dy, ROWS, K + 1, N, 0, NULL) !=
          PCT_MSDA_E_BAD_SHAPE ||
      pct_linear_forward(NULL, dpacked, db, NULL, NULL, NULL, 0.f, dy, ROWS, K, N, 0, NULL) !=
          PCT_MSDA_E_NULL_POINTER) {
    printf("FAIL error codes\n");
    return 1;
  }
  CHECK(cudaDeviceSynchronize());
  float *y = malloc(4 * ROWS * N), *gw = malloc(4 * N * K), *gb = malloc(4 * N);
  cudaMemcpy(y, dy, 4 * ROWS * N, cudaMemcpyDeviceToHost);
  cudaMemcpy(gw, dgw, 4 * N * K, cudaMemcpyDeviceToHost);
  cudaMemcpy(gb, dgb, 4 * N, cudaMemcpyDeviceToHost);

  double err_y = 0, ref_y = 0, err_w = 0, ref_w = 0, err_b = 0, ref_b = 0;
  for (int r = 0; r < ROWS; ++r) {
    double t[N], mean = 0, var = 0;
    for (int n = 0; n < N; ++n) {
      double acc = (double)b[n] + res[r * N + n];
      for (int k = 0; k < K; ++k) acc += (double)x[r * K + k] * w[n * K + k];
      t[n] = acc;
      mean += acc / N;
    }
    for (int n = 0; n < N; ++n) var += (t[n] - mean) * (t[n] - mean) / N;
    for (int n = 0; n < N; ++n) {
      double ref = (t[n] - mean) / sqrt(var + 1e-5) * gamma[n] + beta[n];
      err_y = fmax(err_y, fabs(ref - y[r * N + n]));
      ref_y = fmax(ref_y, fabs(ref));
    }
  }
  for (int n = 0; n < N; ++n) {
    double sb = 0;
    for (int r = 0; r < ROWS; ++r) sb += gy[r * N + n];
    err_b = fmax(err_b, fabs(sb - gb[n]));
    ref_b = fmax(ref_b, fabs(sb));
    for (int k = 0; k < K; ++k) {
      double acc = 0;
      for (int r = 0; r < ROWS; ++r) acc += (double)gy[r * N + n] * x[r * K + k];
      err_w = fmax(err_w, fabs(acc - gw[n * K + k]));
      ref_w = fmax(ref_w, fabs(acc));
    }
  }
  printf("linear+residual+layernorm rel err %.2e, wgrad %.2e, bias grad %.2e\n", err_y / ref_y,
         err_w / ref_w, err_b / ref_b);
  if (err_y / ref_y > 1e-5 || err_w / ref_w > 1e-5 || err_b / ref_b > 1e-5) {
    printf("FAIL tolerance\n");
    return 1;
  }
  /* bf16 operands (PCT_LINEAR_BF16) from C: K = 96 is a multiple of 32; against a double loop over
   * the operands rounded to bf16 on the host (round to nearest even) the fp32-accumulated result
   * has to be as tight as above; the bias gradient sums the unrounded grad_y */
  {
    size_t pb = pct_linear_packed_bytes_bf16(K, N);
    void* dp16 = NULL;
    if (!pb || cudaMalloc(&dp16, pb) != cudaSuccess) {
      printf("FAIL bf16 allocation\n");
      return 1;
    }
    CHECK(pct_linear_pack_weight_bf16(dw, dp16, K, N, NULL));
    CHECK(pct_linear_forward(dx, dp16, db, dres, NULL, NULL, 0.f, dy, ROWS, K, N,
                             PCT_LINEAR_RESIDUAL | PCT_LINEAR_BF16, NULL));
    CHECK(pct_linear_wgrad_ex(dgy, dx, dgw, dgb, ROWS, K, N, PCT_LINEAR_BF16, NULL));
    if (pct_linear_forward(dx, dp16, db, NULL, NULL, NULL, 0.f, dy, ROWS, K, N, PCT_LINEAR_X_BF16, NULL) !=
        PCT_MSDA_E_UNSUPPORTED) {
      printf("FAIL bf16 flag error code\n");
      return 1;
    }
    CHECK(cudaDeviceSynchronize());
    cudaMemcpy(y, dy, 4 * ROWS * N, cudaMemcpyDeviceToHost);
    cudaMemcpy(gw, dgw, 4 * N * K, cudaMemcpyDeviceToHost);
    float *x16 = malloc(4 * ROWS * K), *w16 = malloc(4 * N * K), *g16 = malloc(4 * ROWS * N);
#define RN_BF16(dst, src, count)                                                     \
  for (int i_ = 0; i_ < (count); ++i_) {                                             \
    union { float f; uint32_t u; } v_;                                               \
    v_.f = (src)[i_];                                                                \
    v_.u = (v_.u + 0x7fffu + ((v_.u >> 16) & 1u)) & 0xffff0000u;                     \
    (dst)[i_] = v_.f;                                                                \
  }
    RN_BF16(x16, x, ROWS * K)
    RN_BF16(w16, w, N * K)
    RN_BF16(g16, gy, ROWS * N)
    double e_y = 0, r_y = 0, e_w = 0, r_w = 0;
    for (int r = 0; r < ROWS; ++r)
      for (int n = 0; n < N; ++n) {
        double acc = (double)b[n] + res[r * N + n];
        for (int k = 0; k < K; ++k) acc += (double)x16[r * K + k] * w16[n * K + k];
        e_y = fmax(e_y, fabs(acc - y[r * N + n]));
        r_y = fmax(r_y, fabs(acc));
      }
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        double acc = 0;
        for (int r = 0; r < ROWS; ++r) acc += (double)g16[r * N + n] * x16[r * K + k];
        e_w = fmax(e_w, fabs(acc - gw[n * K + k]));
        r_w = fmax(r_w, fabs(acc));
      }
    printf("bf16 operands: linear+residual rel err %.2e, wgrad %.2e (vs rounded operands)\n", e_y / r_y,
           e_w / r_w);
    if (e_y / r_y > 1e-5 || e_w / r_w > 1e-5) {
      printf("FAIL bf16 tolerance\n");
      return 1;
    }
  }
  printf("C_LINEAR_CLIENT PASS\n");
  return 0;
}
