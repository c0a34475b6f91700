// api.cu — C-ABI glue: error state, device query, residual begin/accumulate/finalize dispatch.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "mlp_thread.cuh"
#include "residual_common.cuh"

namespace pdeip {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
    cached = n;
  } else {
    cudaGetLastError();  // clear the sticky "no device" error on CPU-only hosts
    cached = 148;        // B200
  }
  return cached;
}

// implemented in residual_mlp.cu / residual_tensor.cu / parametric.cu
int mlp_residual_accumulate_fp32(int set_kind, const ResidualArgs& a, int hidden, cudaStream_t st);
int mlp_residual_accumulate_tensor(int set_kind, const ResidualArgs& a, int hidden, cudaStream_t st);
int mlp_eval_fp32(const float* params, int d, int hidden, int layers, const float* x, const float* v,
                  float* out_value, float* out_grad, float* out_vHv, float* out_lap, int64_t n, cudaStream_t st);
int param_residual_accumulate(int set_kind, int model_kind, const ResidualArgs& a, int n_gaussian, cudaStream_t st);
int param_eval(int model_kind, const float* params, int d, int n_gaussian, const float* x, const float* v,
               float* out_value, float* out_grad, float* out_vHv, float* out_lap, int64_t n, cudaStream_t st);
int kmv_mean_grad(int model_kind, const float* params, int d, int hidden, int layers, const float* xv, int64_t n,
                  int nt, const float* ref, int64_t m, float* out_G, float* out_Gtrue, const float* true_A,
                  void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t kmv_ws_bytes(int64_t n, int nt, int d, int64_t m);
int kmv_closure_correction(const float* W, int d, int64_t n, int nt, const float* c, const float* cov, float weight,
                           float* part, cudaStream_t st);
int kmv_density_terms(const float* xv, int64_t n, int nt, int d, const float* coef, float gamma, float* out_c,
                      float* out_ps, float* out_ps2, cudaStream_t st);
int kmv_g_sums(const float* G, const float* Gtrue, int64_t n_rows, int d, float w, float* part_sums, cudaStream_t st);
// residual_tensor.cu: per-device status words of the tcgen05 kernels: [0] residual, [1] integrator (0 = every bounded
// mbarrier wait completed).  begin clears them, finalize turns a non-zero word into NaN sums / gradient, so that a
// timed-out phase reaches the caller's existing NaN check (core/trainer.py:112) without an extra host sync.
int* tensor_status_word();

static int64_t num_params(int model_kind, int d, int hidden, int layers, int n_gaussian) {
  switch (model_kind) {
    case PDEIP_MODEL_MLP:
      return (int64_t)d * hidden + hidden + (int64_t)(layers - 1) * (hidden * hidden + hidden) + hidden * kOut + kOut;
    case PDEIP_MODEL_GMM:
      return (int64_t)n_gaussian * d;
    case PDEIP_MODEL_QUADRATIC:
      return (int64_t)d * d + d;
    default:
      return -1;
  }
}

__global__ void finalize_reduce_kernel(const float* __restrict__ ws, int grid_ctas, int64_t pstride, int64_t P,
                                       float* __restrict__ sums, float* __restrict__ grad,
                                       const int* __restrict__ status) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P + PDEIP_NUM_SUMS) return;
  float s = 0.f;
  for (int b = 0; b < grid_ctas; ++b) s += ws[(int64_t)b * pstride + idx];
  if (status != nullptr && (status[0] | status[1]) != 0) s = __int_as_float(0x7fc00000);  // a tcgen05 phase timed out
  if (idx < P) grad[idx] = s;
  else if (idx - P != PDEIP_SUM_GRADNORM) sums[idx - P] = s;
}

__global__ void grad_norm_kernel(const float* __restrict__ grad, int64_t P, float* __restrict__ sums) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < P; i += blockDim.x) s = fmaf(grad[i], grad[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) sums[PDEIP_SUM_GRADNORM] = sqrtf(t);
  }
}

}  // namespace pdeip

using namespace pdeip;

extern "C" int pdeip_abi_version(void) { return PDEIP_ABI_VERSION; }
extern "C" const char* pdeip_last_error(void) { return g_err; }
extern "C" int pdeip_sm_count(void) { return sm_count(); }

extern "C" int64_t pdeip_model_num_params(int model_kind, int d, int hidden, int layers, int n_gaussian) {
  return num_params(model_kind, d, hidden, layers, n_gaussian);
}

extern "C" size_t pdeip_residual_workspace_bytes(int model_kind, int d, int hidden, int layers, int n_gaussian) {
  const int64_t P = num_params(model_kind, d, hidden, layers, n_gaussian);
  if (P < 0) return 0;
  return residual_ws_bytes(P);
}

extern "C" int pdeip_residual_begin(void* workspace, size_t workspace_bytes, int model_kind, int d, int hidden,
                                    int layers, int n_gaussian, void* stream) {
  const int64_t P = num_params(model_kind, d, hidden, layers, n_gaussian);
  PDEIP_REQUIRE(P > 0, PDEIP_ERR_INVALID_ARG, "unknown model kind %d", model_kind);
  PDEIP_REQUIRE(workspace != nullptr && workspace_bytes >= residual_ws_bytes(P), PDEIP_ERR_WORKSPACE,
                "workspace too small: need %zu bytes, got %zu", residual_ws_bytes(P), workspace_bytes);
  PDEIP_CUDA_OK(cudaMemsetAsync(workspace, 0, residual_ws_bytes(P), (cudaStream_t)stream));
  int* status = tensor_status_word();
  PDEIP_REQUIRE(status != nullptr, PDEIP_ERR_CUDA, "cannot allocate the tensor-path status words");
  PDEIP_CUDA_OK(cudaMemsetAsync(status, 0, 2 * sizeof(int), (cudaStream_t)stream));
  return PDEIP_OK;
}

extern "C" int pdeip_residual_accumulate(void* workspace, size_t workspace_bytes, int set_kind, int model_kind,
                                         const float* params, int d, int hidden, int layers, int n_gaussian,
                                         const float* points, int64_t n_points, int layout, float weight,
                                         float coef, int true_kind, const float* true_params,
                                         int true_n_gaussian, float true_sigma, int path, void* stream) {
  const int64_t P = num_params(model_kind, d, hidden, layers, n_gaussian);
  PDEIP_REQUIRE(P > 0, PDEIP_ERR_INVALID_ARG, "unknown model kind %d", model_kind);
  PDEIP_REQUIRE(workspace != nullptr && workspace_bytes >= residual_ws_bytes(P), PDEIP_ERR_WORKSPACE,
                "workspace too small: need %zu bytes, got %zu", residual_ws_bytes(P), workspace_bytes);
  PDEIP_REQUIRE(params != nullptr, PDEIP_ERR_INVALID_ARG, "params is NULL");
  PDEIP_REQUIRE(n_points >= 0, PDEIP_ERR_INVALID_ARG, "n_points < 0");
  PDEIP_REQUIRE(layout == PDEIP_LAYOUT_AOS || layout == PDEIP_LAYOUT_SOA || layout == PDEIP_LAYOUT_BLOCK128,
                PDEIP_ERR_INVALID_ARG, "bad layout %d", layout);
  PDEIP_REQUIRE(layout != PDEIP_LAYOUT_BLOCK128 || n_points % 128 == 0, PDEIP_ERR_INVALID_ARG,
                "PDEIP_LAYOUT_BLOCK128 needs n_points %% 128 == 0 (got %lld)", (long long)n_points);
  PDEIP_REQUIRE(true_kind == PDEIP_DRIFT_NONE || true_kind == PDEIP_DRIFT_LINEAR || true_kind == PDEIP_DRIFT_GMM ||
                    true_kind == PDEIP_DRIFT_IN_POINTS,
                PDEIP_ERR_INVALID_ARG, "true_kind must be NONE, LINEAR, GMM or IN_POINTS");
  PDEIP_REQUIRE(true_kind == PDEIP_DRIFT_NONE || true_kind == PDEIP_DRIFT_IN_POINTS || true_params != nullptr,
                PDEIP_ERR_INVALID_ARG, "true_params is NULL");
  if (n_points == 0) return PDEIP_OK;
  PDEIP_REQUIRE(points != nullptr, PDEIP_ERR_INVALID_ARG, "points is NULL");
  ResidualArgs a;
  memset(&a, 0, sizeof(a));
  a.params = params; a.points = points; a.n_points = n_points; a.layout = layout; a.d = d; a.layers = layers;
  a.weight = weight; a.coef = coef;
  a.tg.kind = true_kind; a.tg.params = true_params; a.tg.n_gaussian = true_n_gaussian;
  a.tg.inv_sigma2 = (true_kind == PDEIP_DRIFT_GMM && true_sigma > 0.f) ? 1.f / (true_sigma * true_sigma) : 1.f;
  a.ws = (float*)workspace; a.pstride = residual_pstride(P);
  cudaStream_t st = (cudaStream_t)stream;
  if (model_kind == PDEIP_MODEL_MLP) {
    if (path == PDEIP_PATH_TENSOR) return mlp_residual_accumulate_tensor(set_kind, a, hidden, st);
    PDEIP_REQUIRE(path == PDEIP_PATH_FP32, PDEIP_ERR_INVALID_ARG, "unknown path %d", path);
    return mlp_residual_accumulate_fp32(set_kind, a, hidden, st);
  }
  return param_residual_accumulate(set_kind, model_kind, a, n_gaussian, st);
}

extern "C" int pdeip_residual_finalize(void* workspace, size_t workspace_bytes, int model_kind, int d, int hidden,
                                       int layers, int n_gaussian, float* sums, float* grad, void* stream) {
  const int64_t P = num_params(model_kind, d, hidden, layers, n_gaussian);
  PDEIP_REQUIRE(P > 0, PDEIP_ERR_INVALID_ARG, "unknown model kind %d", model_kind);
  PDEIP_REQUIRE(workspace != nullptr && workspace_bytes >= residual_ws_bytes(P), PDEIP_ERR_WORKSPACE,
                "workspace too small: need %zu bytes, got %zu", residual_ws_bytes(P), workspace_bytes);
  PDEIP_REQUIRE(sums && grad, PDEIP_ERR_INVALID_ARG, "sums / grad is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = P + PDEIP_NUM_SUMS;
  finalize_reduce_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>((const float*)workspace, residual_grid(),
                                                                          residual_pstride(P), P, sums, grad,
                                                                          tensor_status_word());
  PDEIP_LAUNCH_OK();
  grad_norm_kernel<<<1, 512, 0, st>>>(grad, P, sums);
  PDEIP_LAUNCH_OK();
  return PDEIP_OK;
}

extern "C" int pdeip_model_eval(int model_kind, const float* params, int d, int hidden, int layers, int n_gaussian,
                                const float* x, const float* v, float* out_value, float* out_grad, float* out_vHv,
                                float* out_lap, int64_t n, void* stream) {
  PDEIP_REQUIRE(params && x, PDEIP_ERR_INVALID_ARG, "params / x is NULL");
  PDEIP_REQUIRE(n >= 0, PDEIP_ERR_INVALID_ARG, "n < 0");
  PDEIP_REQUIRE(!(out_vHv && !v), PDEIP_ERR_INVALID_ARG, "out_vHv requested without v");
  if (n == 0) return PDEIP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (model_kind == PDEIP_MODEL_MLP)
    return mlp_eval_fp32(params, d, hidden, layers, x, v, out_value, out_grad, out_vHv, out_lap, n, st);
  return param_eval(model_kind, params, d, n_gaussian, x, v, out_value, out_grad, out_vHv, out_lap, n, st);
}

extern "C" size_t pdeip_kmv_workspace_bytes_ref(int64_t n, int nt, int d, int64_t m) {
  if (n < 1 || nt < 1 || d < 1 || m < 1) return 0;
  return kmv_ws_bytes(n, nt, d, m);
}
extern "C" size_t pdeip_kmv_workspace_bytes(int64_t n, int nt, int d) { return pdeip_kmv_workspace_bytes_ref(n, nt, d, n); }

extern "C" int pdeip_kmv_mean_grad_ref(int model_kind, const float* params, int d, int hidden, int layers,
                                       const float* xv, int64_t n, int nt, const float* ref, int64_t m, float* out_G,
                                       float* out_Gtrue, const float* true_A, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  PDEIP_REQUIRE(params && xv && out_G, PDEIP_ERR_INVALID_ARG, "NULL argument");
  PDEIP_REQUIRE(n >= 1 && nt >= 1, PDEIP_ERR_INVALID_ARG, "n / nt must be >= 1");
  if (ref == nullptr) { ref = xv; m = n; }
  PDEIP_REQUIRE(m >= 1, PDEIP_ERR_INVALID_ARG, "reference-set size m must be >= 1");
  return kmv_mean_grad(model_kind, params, d, hidden, layers, xv, n, nt, ref, m, out_G, out_Gtrue, true_A, workspace,
                       workspace_bytes, (cudaStream_t)stream);
}
extern "C" int pdeip_kmv_mean_grad(int model_kind, const float* params, int d, int hidden, int layers,
                                   const float* xv, int64_t n, int nt, float* out_G, float* out_Gtrue,
                                   const float* true_A, void* workspace, size_t workspace_bytes, void* stream) {
  return pdeip_kmv_mean_grad_ref(model_kind, params, d, hidden, layers, xv, n, nt, nullptr, n, out_G, out_Gtrue, true_A,
                                 workspace, workspace_bytes, stream);
}

extern "C" int pdeip_residual_accumulate_kmv_ref(void* workspace, size_t workspace_bytes, int model_kind,
                                                 const float* params, int d, int hidden, int layers, const float* xv,
                                                 int64_t n, int nt, const float* ref, int64_t m, const float* G,
                                                 const float* G_true, const float* c, float weight, void* stream) {
  const int64_t P = num_params(model_kind, d, hidden, layers, 0);
  PDEIP_REQUIRE(P > 0, PDEIP_ERR_INVALID_ARG, "unknown model kind %d", model_kind);
  PDEIP_REQUIRE(workspace != nullptr && workspace_bytes >= residual_ws_bytes(P), PDEIP_ERR_WORKSPACE,
                "workspace too small: need %zu bytes, got %zu", residual_ws_bytes(P), workspace_bytes);
  PDEIP_REQUIRE(params && xv && c, PDEIP_ERR_INVALID_ARG, "NULL argument");
  PDEIP_REQUIRE(n >= 1 && nt >= 1, PDEIP_ERR_INVALID_ARG, "n / nt must be >= 1");
  if (ref == nullptr) { ref = xv; m = n; }
  PDEIP_REQUIRE(m >= 1, PDEIP_ERR_INVALID_ARG, "reference-set size m must be >= 1");
  ResidualArgs a;
  memset(&a, 0, sizeof(a));
  a.params = params; a.points = xv; a.n_points = m * n * nt; a.layout = PDEIP_LAYOUT_AOS; a.d = d;
  a.layers = layers; a.weight = weight; a.G = G; a.c = c; a.ref = ref; a.kmv_n = n; a.kmv_nt = nt;
  a.ws = (float*)workspace; a.pstride = residual_pstride(P);
  cudaStream_t st = (cudaStream_t)stream;
  if (G) {  // |G|^2, |G_true|^2, |G_true - G|^2 with weight 1/(n*nt) = weight * m
    int rc = kmv_g_sums(G, G_true, n * nt, d, weight * (float)m, a.ws + P, st);
    if (rc != PDEIP_OK) return rc;
  }
  if (model_kind == PDEIP_MODEL_MLP) return mlp_residual_accumulate_fp32(PDEIP_SET_KMV_PAIRS, a, hidden, st);
  return param_residual_accumulate(PDEIP_SET_KMV_PAIRS, model_kind, a, 0, st);
}
extern "C" int pdeip_residual_accumulate_kmv(void* workspace, size_t workspace_bytes, int model_kind,
                                             const float* params, int d, int hidden, int layers, const float* xv,
                                             int64_t n, int nt, const float* G, const float* G_true,
                                             const float* c, float weight, void* stream) {
  return pdeip_residual_accumulate_kmv_ref(workspace, workspace_bytes, model_kind, params, d, hidden, layers, xv, n, nt,
                                           nullptr, n, G, G_true, c, weight, stream);
}

extern "C" int pdeip_kmv_closure_correction(void* workspace, size_t workspace_bytes, const float* params, int d,
                                            int64_t n, int nt, const float* c, const float* cov, float weight,
                                            void* stream) {
  const int64_t P = num_params(PDEIP_MODEL_QUADRATIC, d, 0, 0, 0);
  PDEIP_REQUIRE(workspace != nullptr && workspace_bytes >= residual_ws_bytes(P), PDEIP_ERR_WORKSPACE,
                "workspace too small: need %zu bytes, got %zu", residual_ws_bytes(P), workspace_bytes);
  PDEIP_REQUIRE(params && c && cov && n >= 1, PDEIP_ERR_INVALID_ARG, "NULL argument / n < 1");
  return kmv_closure_correction(params, d, n, nt, c, cov, weight, (float*)workspace, (cudaStream_t)stream);
}

extern "C" int pdeip_kmv_density_terms(const float* xv, int64_t n, int nt, int d, const float* coef, float gamma,
                                       float* out_c, float* out_ps, float* out_ps2, void* stream) {
  PDEIP_REQUIRE(xv && coef && (out_c || out_ps || out_ps2), PDEIP_ERR_INVALID_ARG, "NULL argument");
  PDEIP_REQUIRE(n >= 1 && nt >= 1, PDEIP_ERR_INVALID_ARG, "n / nt must be >= 1");
  return kmv_density_terms(xv, n, nt, d, coef, gamma, out_c, out_ps, out_ps2, (cudaStream_t)stream);
}
