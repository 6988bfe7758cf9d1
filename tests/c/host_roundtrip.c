/* Plain-C client of the C ABI (no Python, no torch): fills host buffers, runs the host pipeline
 * (pct_msda_host_forward_backward) and checks the results against the CPU oracle
 * (oracle/msda_oracle.c, test infrastructure).  Built and run by tests/test_gpu_c_host.py:
 *   gcc host_roundtrip.c -I include -L <libs> -lpctrans_msda -lmsda_oracle -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "pct_msda.h"

int msda_oracle_forward_f64(const double*, const int64_t*, const int64_t*, const double*, const double*,
                            double*, int, int, int, int, int, int, int);
int msda_oracle_backward_f64(const double*, const int64_t*, const int64_t*, const double*, const double*,
                             const double*, double*, double*, double*, int, int, int, int, int, int, int);

static uint64_t rng = 0x9e3779b97f4a7c15ull;
static double urand(void) {
  rng = rng * 6364136223846793005ull + 1442695040888963407ull;
  return (double)(rng >> 11) / 9007199254740992.0;
}

static double max_rel(const float* a, const double* b, size_t n) {
  double err = 0, ref = 0;
  for (size_t i = 0; i < n; ++i) {
    double e = fabs((double)a[i] - b[i]);
    if (e > err) err = e;
    if (fabs(b[i]) > ref) ref = fabs(b[i]);
  }
  return err / (ref > 0 ? ref : 1e-30);
}

int main(void) {
  enum { N = 5, M = 8, D = 32, L = 3, P = 4 };
  const int64_t shapes[L * 2] = {6, 9, 12, 17, 23, 31};
  int64_t lsi[L];
  int S = 0;
  for (int l = 0; l < L; ++l) { lsi[l] = S; S += (int)(shapes[2 * l] * shapes[2 * l + 1]); }
  const int Lq = 211;
  const size_t n_val = (size_t)N * S * M * D, n_loc = (size_t)N * Lq * M * L * P * 2;
  const size_t n_att = n_loc / 2, n_out = (size_t)N * Lq * M * D;
  float *value = malloc(4 * n_val), *loc = malloc(4 * n_loc), *attn = malloc(4 * n_att),
        *gout = malloc(4 * n_out), *out = malloc(4 * n_out), *gv = malloc(4 * n_val),
        *gl = malloc(4 * n_loc), *ga = malloc(4 * n_att);
  double *dv = malloc(8 * n_val), *dl = malloc(8 * n_loc), *da = malloc(8 * n_att),
         *dg = malloc(8 * n_out), *ro = malloc(8 * n_out), *rgv = malloc(8 * n_val),
         *rgl = malloc(8 * n_loc), *rga = malloc(8 * n_att);
  for (size_t i = 0; i < n_val; ++i) dv[i] = value[i] = (float)(urand() * 2 - 1);
  for (size_t i = 0; i < n_loc; ++i) {
    /* in [-0.2, 1.2], kept 2e-3 px off the pixel grid of every level (d/d(loc) is one-sided there) */
    float x = (float)(urand() * 1.4 - 0.2);
    int l = (int)((i / 2 / P) % L);
    double size = (double)shapes[2 * l + ((i & 1) ? 0 : 1)];
    double c = (double)x * size - 0.5;
    if (fabs(c - floor(c + 0.5)) < 2e-3) x += (float)(6e-3 / size);
    dl[i] = loc[i] = x;
  }
  for (size_t i = 0; i < n_att; ++i) da[i] = attn[i] = (float)(urand() / (L * P));
  for (size_t i = 0; i < n_out; ++i) dg[i] = gout[i] = (float)(urand() * 2 - 1);

  pct_msda_host_ctx* ctx = NULL;
  int rc = pct_msda_host_ctx_create(&ctx, 0, 2);
  if (rc) { printf("ctx_create failed: %s\n", pct_msda_error_string(rc)); return 1; }
  for (int rep = 0; rep < 2; ++rep) {   /* second call reuses the staging buffers */
    rc = pct_msda_host_forward_backward(ctx, value, shapes, lsi, loc, attn, gout, out, gv, gl, ga,
                                        N, S, M, D, L, Lq, P, PCT_MSDA_F32);
    if (rc) { printf("host_forward_backward failed: %s\n", pct_msda_error_string(rc)); return 1; }
  }
  float* out2 = malloc(4 * n_out);
  rc = pct_msda_host_forward(ctx, value, shapes, lsi, loc, attn, out2, N, S, M, D, L, Lq, P, PCT_MSDA_F32);
  if (rc) { printf("host_forward failed: %s\n", pct_msda_error_string(rc)); return 1; }
  pct_msda_host_ctx_destroy(ctx);

  msda_oracle_forward_f64(dv, shapes, lsi, dl, da, ro, N, S, M, D, L, Lq, P);
  msda_oracle_backward_f64(dv, shapes, lsi, dl, da, dg, rgv, rgl, rga, N, S, M, D, L, Lq, P);
  const double e_out = max_rel(out, ro, n_out), e_out2 = max_rel(out2, ro, n_out);
  const double e_gv = max_rel(gv, rgv, n_val), e_gl = max_rel(gl, rgl, n_loc), e_ga = max_rel(ga, rga, n_att);
  printf("out %.3g out(fwd-only) %.3g grad_value %.3g grad_loc %.3g grad_attn %.3g launches %llu\n",
         e_out, e_out2, e_gv, e_gl, e_ga, (unsigned long long)pct_msda_launch_count());
  const int ok = e_out <= 1e-5 && e_out2 <= 1e-5 && e_gv <= 1e-4 && e_gl <= 1e-4 && e_ga <= 1e-4;
  printf(ok ? "C_HOST_ROUNDTRIP PASS\n" : "C_HOST_ROUNDTRIP FAIL\n");
  return ok ? 0 : 2;
}
