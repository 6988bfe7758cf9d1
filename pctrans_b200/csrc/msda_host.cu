// msda_host.cu -- host-buffer entry points of the C ABI (include/pct_msda.h, "host pipeline").
//
// pct_msda_forward / pct_msda_backward take DEVICE pointers, like the reference's launchers
// (OPS/src/cuda/ms_deform_attn_cuda.cu:25-158, called with CUDA tensors).  A host language that
// holds its tensors in host memory would otherwise have to re-implement the staging itself, so
// this file provides the end-to-end call: images are independent in this op, the
// batch is cut into chunks of images and three CUDA streams overlap the H2D copy of chunk i+1,
// the kernels of chunk i and the D2H copy of chunk i-1.  Staging buffers live in a context object
// the caller creates once (the device entry points stay allocation-free).
#include "../../include/pct_msda.h"

#include <cuda_runtime.h>

#include <new>

namespace {

constexpr int kSlots = 4;

struct Slot {
  void* value = nullptr;
  void* loc = nullptr;
  void* attn = nullptr;
  void* gout = nullptr;
  void* out = nullptr;
  void* gvalue = nullptr;
  void* gloc = nullptr;
  void* gattn = nullptr;
  void* ws = nullptr;
  cudaEvent_t loaded = nullptr, done = nullptr, free_ev = nullptr;
};

inline size_t esize(int dtype) { return dtype == PCT_MSDA_F32 ? 4 : dtype == PCT_MSDA_BF16 ? 2 : 8; }
inline size_t lsize(int dtype) { return dtype == PCT_MSDA_F64 ? 8 : 4; }

}  // namespace

struct pct_msda_host_ctx {
  int device = 0;
  int chunk_images = 0;  // 0 = batch/16
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  Slot slots[kSlots];
  int64_t* d_shapes = nullptr;
  int64_t* d_lsi = nullptr;
  int meta_levels = 0;
  // capacity of the staging buffers, in bytes per slot
  size_t cap_value = 0, cap_loc = 0, cap_attn = 0, cap_io = 0, cap_ws = 0;
  bool has_bwd = false;
};

namespace {

#define PCT_CUDA(expr)                         \
  do {                                         \
    cudaError_t e_ = (expr);                   \
    if (e_ != cudaSuccess) return (int)e_;     \
  } while (0)

int realloc_dev(void** p, size_t bytes) {
  if (*p) PCT_CUDA(cudaFree(*p));
  *p = nullptr;
  if (bytes) PCT_CUDA(cudaMalloc(p, bytes));
  return 0;
}

void free_slots(pct_msda_host_ctx* c) {
  for (int i = 0; i < kSlots; ++i) {
    Slot& s = c->slots[i];
    void** bufs[] = {&s.value, &s.loc, &s.attn, &s.gout, &s.out, &s.gvalue, &s.gloc, &s.gattn, &s.ws};
    for (void** b : bufs) {
      if (*b) cudaFree(*b);
      *b = nullptr;
    }
  }
  c->cap_value = c->cap_loc = c->cap_attn = c->cap_io = c->cap_ws = 0;
  c->has_bwd = false;
}

// Staging buffers only ever grow; when any of them is too small every slot is reallocated at the
// new sizes (after draining the streams -- nothing is in flight between calls anyway).  The new
// capacities are committed only when EVERY allocation succeeded; a failure half way frees all
// slots and resets the capacities, so the next call starts from a clean context.
int ensure_capacity(pct_msda_host_ctx* c, size_t b_value, size_t b_loc, size_t b_attn, size_t b_io,
                    size_t b_ws, bool backward) {
  const bool fits = b_value <= c->cap_value && b_loc <= c->cap_loc && b_attn <= c->cap_attn &&
                    b_io <= c->cap_io && (!backward || (c->has_bwd && b_ws <= c->cap_ws));
  if (fits) return 0;
  PCT_CUDA(cudaStreamSynchronize(c->s_in));
  PCT_CUDA(cudaStreamSynchronize(c->s_run));
  PCT_CUDA(cudaStreamSynchronize(c->s_out));
  auto mx = [](size_t a, size_t b) { return a > b ? a : b; };
  const size_t n_value = mx(c->cap_value, b_value), n_loc = mx(c->cap_loc, b_loc);
  const size_t n_attn = mx(c->cap_attn, b_attn), n_io = mx(c->cap_io, b_io);
  const size_t n_ws = mx(c->cap_ws, backward ? b_ws : 0);
  const bool n_bwd = c->has_bwd || backward;
  int rc = 0;
  for (int i = 0; i < kSlots && !rc; ++i) {
    Slot& s = c->slots[i];
    if (!rc) rc = realloc_dev(&s.value, n_value);
    if (!rc) rc = realloc_dev(&s.loc, n_loc);
    if (!rc) rc = realloc_dev(&s.attn, n_attn);
    if (!rc) rc = realloc_dev(&s.out, n_io);
    if (n_bwd) {
      if (!rc) rc = realloc_dev(&s.gout, n_io);
      if (!rc) rc = realloc_dev(&s.gvalue, n_value);
      if (!rc) rc = realloc_dev(&s.gloc, n_loc);
      if (!rc) rc = realloc_dev(&s.gattn, n_attn);
      if (!rc) rc = realloc_dev(&s.ws, n_ws);
    }
  }
  if (rc) {
    free_slots(c);
    cudaGetLastError();   // clear the sticky allocation error for the caller's next call
    return rc;
  }
  c->cap_value = n_value; c->cap_loc = n_loc; c->cap_attn = n_attn; c->cap_io = n_io; c->cap_ws = n_ws;
  c->has_bwd = n_bwd;
  return 0;
}

// Every entry point works on the context's device and puts the caller's current device back on
// every exit path (the Python wrapper may be destroyed by the garbage collector at any time).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  int enter(int device) {
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) return (int)e;
    if (prev != device) {
      e = cudaSetDevice(device);
      if (e != cudaSuccess) return (int)e;
      switched = true;
    }
    return 0;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

int pipeline_body(pct_msda_host_ctx* c, const void* value, const int64_t* shapes, const int64_t* lsi,
                  const void* loc, const void* attn, const void* grad_out, void* out,
                  void* grad_value, void* grad_loc, void* grad_attn, int N, int S, int M, int D,
                  int L, int Lq, int P, int dtype);

int run_pipeline(pct_msda_host_ctx* c, const void* value, const int64_t* shapes, const int64_t* lsi,
                 const void* loc, const void* attn, const void* grad_out, void* out,
                 void* grad_value, void* grad_loc, void* grad_attn, int N, int S, int M, int D,
                 int L, int Lq, int P, int dtype) {
  if (!c) return PCT_MSDA_E_NULL_POINTER;
  if (dtype != PCT_MSDA_F32 && dtype != PCT_MSDA_BF16 && dtype != PCT_MSDA_F64) return PCT_MSDA_E_BAD_DTYPE;
  if (N < 0 || S < 0 || Lq < 0 || M <= 0 || D <= 0 || L <= 0 || P <= 0) return PCT_MSDA_E_BAD_SHAPE;
  const bool backward = grad_out != nullptr;
  if (N == 0) return 0;
  if (!shapes || !lsi || !out || (S && !value) || (Lq && (!loc || !attn))) return PCT_MSDA_E_NULL_POINTER;
  if (backward && (!grad_value || !grad_loc || !grad_attn)) return PCT_MSDA_E_NULL_POINTER;
  DeviceGuard guard;
  int rc = guard.enter(c->device);
  if (rc) return rc;
  rc = pipeline_body(c, value, shapes, lsi, loc, attn, grad_out, out, grad_value, grad_loc, grad_attn,
                     N, S, M, D, L, Lq, P, dtype);
  if (rc) {
    // an error in the middle of the loop: copies to / from the caller's host buffers may still be
    // in flight -- drain the three streams before the caller gets its buffers back
    cudaStreamSynchronize(c->s_in);
    cudaStreamSynchronize(c->s_run);
    cudaStreamSynchronize(c->s_out);
  }
  return rc;
}

int pipeline_body(pct_msda_host_ctx* c, const void* value, const int64_t* shapes, const int64_t* lsi,
                  const void* loc, const void* attn, const void* grad_out, void* out,
                  void* grad_value, void* grad_loc, void* grad_attn, int N, int S, int M, int D,
                  int L, int Lq, int P, int dtype) {
  const bool backward = grad_out != nullptr;

  const size_t es = esize(dtype), ls = lsize(dtype);
  const size_t img_value = (size_t)S * M * D * es;
  const size_t img_loc = (size_t)Lq * M * L * P * 2 * ls;
  const size_t img_attn = (size_t)Lq * M * L * P * ls;
  const size_t img_io = (size_t)Lq * M * D * es;
  int chunk = c->chunk_images > 0 ? c->chunk_images : (N / 16 > 0 ? N / 16 : 1);
  if (chunk > N) chunk = N;
  const size_t img_ws = pct_msda_backward_workspace_bytes(1, S, M, D, dtype);
  int rc = ensure_capacity(c, img_value * chunk, img_loc * chunk, img_attn * chunk, img_io * chunk,
                           img_ws * chunk, backward);
  if (rc) return rc;

  // level metadata lives on the device for the kernels (cu:39-40)
  if (c->meta_levels < L) {
    if (c->d_shapes) PCT_CUDA(cudaFree(c->d_shapes));
    if (c->d_lsi) PCT_CUDA(cudaFree(c->d_lsi));
    PCT_CUDA(cudaMalloc((void**)&c->d_shapes, sizeof(int64_t) * 2 * L));
    PCT_CUDA(cudaMalloc((void**)&c->d_lsi, sizeof(int64_t) * L));
    c->meta_levels = L;
  }
  PCT_CUDA(cudaMemcpyAsync(c->d_shapes, shapes, sizeof(int64_t) * 2 * L, cudaMemcpyHostToDevice, c->s_in));
  PCT_CUDA(cudaMemcpyAsync(c->d_lsi, lsi, sizeof(int64_t) * L, cudaMemcpyHostToDevice, c->s_in));

  const char* h_value = (const char*)value;
  const char* h_loc = (const char*)loc;
  const char* h_attn = (const char*)attn;
  const char* h_gout = (const char*)grad_out;
  char* h_out = (char*)out;
  char* h_gvalue = (char*)grad_value;
  char* h_gloc = (char*)grad_loc;
  char* h_gattn = (char*)grad_attn;

  int i = 0;
  for (int lo = 0; lo < N; lo += chunk, ++i) {
    const int n = (N - lo < chunk) ? (N - lo) : chunk;
    Slot& s = c->slots[i % kSlots];
    if (i >= kSlots) PCT_CUDA(cudaStreamWaitEvent(c->s_in, s.free_ev, 0));  // slot drained to the host
    if (img_value) PCT_CUDA(cudaMemcpyAsync(s.value, h_value + lo * img_value, n * img_value, cudaMemcpyHostToDevice, c->s_in));
    if (img_loc) {
      PCT_CUDA(cudaMemcpyAsync(s.loc, h_loc + lo * img_loc, n * img_loc, cudaMemcpyHostToDevice, c->s_in));
      PCT_CUDA(cudaMemcpyAsync(s.attn, h_attn + lo * img_attn, n * img_attn, cudaMemcpyHostToDevice, c->s_in));
    }
    if (backward && img_io)
      PCT_CUDA(cudaMemcpyAsync(s.gout, h_gout + lo * img_io, n * img_io, cudaMemcpyHostToDevice, c->s_in));
    PCT_CUDA(cudaEventRecord(s.loaded, c->s_in));

    PCT_CUDA(cudaStreamWaitEvent(c->s_run, s.loaded, 0));
    rc = pct_msda_forward(s.value, c->d_shapes, c->d_lsi, s.loc, s.attn, s.out, n, S, M, D, L, Lq, P,
                          dtype, c->s_run);
    if (rc) return rc;
    if (backward) {
      rc = pct_msda_backward(s.value, c->d_shapes, c->d_lsi, s.loc, s.attn, s.gout, s.gvalue, s.gloc,
                             s.gattn, n, S, M, D, L, Lq, P, dtype, s.ws, img_ws * n, c->s_run);
      if (rc) return rc;
    }
    PCT_CUDA(cudaEventRecord(s.done, c->s_run));

    PCT_CUDA(cudaStreamWaitEvent(c->s_out, s.done, 0));
    if (img_io) PCT_CUDA(cudaMemcpyAsync(h_out + lo * img_io, s.out, n * img_io, cudaMemcpyDeviceToHost, c->s_out));
    if (backward) {
      if (img_value) PCT_CUDA(cudaMemcpyAsync(h_gvalue + lo * img_value, s.gvalue, n * img_value, cudaMemcpyDeviceToHost, c->s_out));
      if (img_loc) {
        PCT_CUDA(cudaMemcpyAsync(h_gloc + lo * img_loc, s.gloc, n * img_loc, cudaMemcpyDeviceToHost, c->s_out));
        PCT_CUDA(cudaMemcpyAsync(h_gattn + lo * img_attn, s.gattn, n * img_attn, cudaMemcpyDeviceToHost, c->s_out));
      }
    }
    PCT_CUDA(cudaEventRecord(s.free_ev, c->s_out));
  }
  PCT_CUDA(cudaStreamSynchronize(c->s_out));   // results are in host memory on return
  return 0;
}

}  // namespace

extern "C" {

int pct_msda_host_ctx_create(pct_msda_host_ctx** ctx, int device, int chunk_images) {
  if (!ctx) return PCT_MSDA_E_NULL_POINTER;
  *ctx = nullptr;
  DeviceGuard guard;
  if (int rc = guard.enter(device)) return rc;
  pct_msda_host_ctx* c = new (std::nothrow) pct_msda_host_ctx();
  if (!c) return (int)cudaErrorMemoryAllocation;
  c->device = device;
  c->chunk_images = chunk_images;
  cudaError_t e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_run, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
  for (int i = 0; i < kSlots && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&c->slots[i].loaded, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->slots[i].done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->slots[i].free_ev, cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    pct_msda_host_ctx_destroy(c);
    return (int)e;
  }
  *ctx = c;
  return 0;
}

void pct_msda_host_ctx_destroy(pct_msda_host_ctx* c) {
  if (!c) return;
  DeviceGuard guard;
  guard.enter(c->device);
  if (c->s_out) cudaStreamSynchronize(c->s_out);
  free_slots(c);
  for (int i = 0; i < kSlots; ++i) {
    Slot& s = c->slots[i];
    if (s.loaded) cudaEventDestroy(s.loaded);
    if (s.done) cudaEventDestroy(s.done);
    if (s.free_ev) cudaEventDestroy(s.free_ev);
  }
  if (c->d_shapes) cudaFree(c->d_shapes);
  if (c->d_lsi) cudaFree(c->d_lsi);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_run) cudaStreamDestroy(c->s_run);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  delete c;
}

int pct_msda_host_forward(pct_msda_host_ctx* ctx, const void* value, const int64_t* spatial_shapes,
                          const int64_t* level_start_index, const void* sampling_loc,
                          const void* attn_weight, void* output, int batch, int spatial_size,
                          int num_heads, int channels, int num_levels, int num_query,
                          int num_point, int value_dtype) {
  return run_pipeline(ctx, value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                      nullptr, output, nullptr, nullptr, nullptr, batch, spatial_size, num_heads,
                      channels, num_levels, num_query, num_point, value_dtype);
}

int pct_msda_host_forward_backward(pct_msda_host_ctx* ctx, const void* value,
                                   const int64_t* spatial_shapes, const int64_t* level_start_index,
                                   const void* sampling_loc, const void* attn_weight,
                                   const void* grad_output, void* output, void* grad_value,
                                   void* grad_sampling_loc, void* grad_attn_weight, int batch,
                                   int spatial_size, int num_heads, int channels, int num_levels,
                                   int num_query, int num_point, int value_dtype) {
  if (!grad_output) return PCT_MSDA_E_NULL_POINTER;
  return run_pipeline(ctx, value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                      grad_output, output, grad_value, grad_sampling_loc, grad_attn_weight, batch,
                      spatial_size, num_heads, channels, num_levels, num_query, num_point,
                      value_dtype);
}

}  // extern "C"
