// extern "C" boundary: argument validation, dtype / architecture dispatch, error text.
// See include/maskunet_b200.h for the contract of every entry point.
#include <cstring>

#include "common.cuh"
#include "det_reduce.cuh"
#include "tma_host.cuh"

namespace mu {

// ---- deterministic-mode scratch, registered per device by the caller (mu_set_deterministic_scratch)
namespace {
struct DetScratch { void* ptr; size_t bytes; };
std::mutex g_det_mu;
DetScratch g_det_scratch[64] = {};
constexpr size_t kDetCounterBytes = 1024;     // 256 uint32 counters in front of the partial area
}  // namespace

bool det_context(int slot, size_t floats_needed, DetCtx* out, const char* who) {
  out->partial = nullptr;
  out->counter = nullptr;
  out->floats = 0;
  if (!get_deterministic()) return true;
  int dev = 0;
  cudaGetDevice(&dev);
  DetScratch sc{nullptr, 0};
  if (dev >= 0 && dev < 64) {
    std::lock_guard<std::mutex> lock(g_det_mu);
    sc = g_det_scratch[dev];
  }
  const size_t need = kDetCounterBytes + floats_needed * sizeof(float);
  if (sc.ptr == nullptr || sc.bytes < need) {
    set_error("%s: deterministic mode needs a scratch buffer of at least %zu bytes on device %d "
              "(mu_set_deterministic_scratch; %zu registered)", who, need, dev, sc.bytes);
    return false;
  }
  out->counter = reinterpret_cast<unsigned*>(sc.ptr) + slot;
  out->partial = reinterpret_cast<float*>(reinterpret_cast<char*>(sc.ptr) + kDetCounterBytes);
  out->floats = (sc.bytes - kDetCounterBytes) / sizeof(float);
  return true;
}

static thread_local char g_err[512] = "";

TmapCache& tmap_cache() {
  static TmapCache cache;
  return cache;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

static int device_cc_major() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  return major;
}

static int check_common(const char* fn, int B, int C, int N, int dtype) {
  MU_REQUIRE(B > 0 && N > 0, MU_ERR_BAD_SHAPE, "%s: B and N must be positive (B=%d N=%d)", fn, B, N);
  MU_REQUIRE(C == 64 || C == 128 || C == 256, MU_ERR_BAD_SHAPE, "%s: channels must be 64, 128 or 256 (got %d)", fn, C);
  MU_REQUIRE(dtype == MU_F32 || dtype == MU_BF16, MU_ERR_BAD_DTYPE, "%s: unknown dtype code %d", fn, dtype);
  return 0;
}
static int check_nkp(const char* fn, int N, int NKP) {
  MU_REQUIRE(NKP % 128 == 0 && NKP >= N, MU_ERR_BAD_SHAPE, "%s: NKP must be a multiple of 128 and >= N (N=%d NKP=%d)",
             fn, N, NKP);
  return 0;
}
#define MU_PTRS(fn, ...)                                                                        \
  do {                                                                                          \
    const void* ptrs_[] = {__VA_ARGS__};                                                        \
    for (size_t i_ = 0; i_ < sizeof(ptrs_) / sizeof(ptrs_[0]); ++i_) {                          \
      MU_REQUIRE(ptrs_[i_] != nullptr, MU_ERR_NULL, "%s: null pointer (argument %zu)", fn, i_); \
      MU_REQUIRE(aligned16(ptrs_[i_]), MU_ERR_MISALIGNED, "%s: pointer %zu not 16-byte aligned", fn, i_); \
    }                                                                                           \
  } while (0)

}  // namespace mu

using namespace mu;

extern "C" {

int mu_version(void) { return 100; }  // 0.1.0

const char* mu_last_error(void) { return g_err; }

int mu_device_supported(void) { return device_cc_major() == 10 ? 1 : 0; }

void mu_tmap_cache_stats(uint64_t* hits, uint64_t* misses, uint64_t* entries) {
  TmapCache& c = tmap_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  if (hits) *hits = c.hits;
  if (misses) *misses = c.misses;
  if (entries) *entries = c.map.size();
}

void mu_set_deterministic(int32_t on) { set_deterministic(on); }

int mu_set_deterministic_scratch(void* scratch, size_t bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("mu_set_deterministic_scratch: no current CUDA device");
    return MU_ERR_DRIVER;
  }
  MU_REQUIRE(scratch == nullptr || (aligned16(scratch) && bytes >= (1u << 20)), MU_ERR_WORKSPACE,
             "mu_set_deterministic_scratch: the scratch must be 16-byte aligned and at least 1 MiB (got %zu bytes)", bytes);
  std::lock_guard<std::mutex> lock(g_det_mu);
  g_det_scratch[dev] = DetScratch{scratch, scratch != nullptr ? bytes : 0};
  return 0;
}
int32_t mu_get_deterministic(void) { return get_deterministic(); }

void mu_tmap_cache_clear(void) {
  TmapCache& c = tmap_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  c.map.clear();
}

int mu_mask_binarize(const int64_t* bits, int32_t B, int32_t N, uint32_t* keep_bits, int32_t* n_keep,
                     int32_t* keep_idx, int32_t* keep_rank, mu_stream_t stream) {
  MU_REQUIRE(B > 0 && N > 0, MU_ERR_BAD_SHAPE, "mu_mask_binarize: B and N must be positive (B=%d N=%d)", B, N);
  MU_PTRS("mu_mask_binarize", bits, keep_bits, n_keep, keep_idx, keep_rank);
  return launch_mask_binarize(bits, B, N, keep_bits, n_keep, keep_idx, keep_rank, (cudaStream_t)stream);
}

int mu_qkv_project(const void* x, const float* w_qkv, const void* w_qkv_lp, const float* b_qkv,
                   const int32_t* keep_rank, const int32_t* n_keep, void* q, void* kc, void* vc, int32_t B, int32_t C,
                   int32_t N, int32_t NKP, int32_t dtype, int32_t x_layout, mu_stream_t stream) {
  int rc;
  if ((rc = check_common("mu_qkv_project", B, C, N, dtype))) return rc;
  if ((rc = check_nkp("mu_qkv_project", N, NKP))) return rc;
  MU_PTRS("mu_qkv_project", x, w_qkv, b_qkv, keep_rank, n_keep, q, kc, vc);
  if (x_layout == MU_X_CHANNEL_MAJOR)
    return launch_qkv_project(x, w_qkv, b_qkv, keep_rank, n_keep, q, kc, vc, B, C, N, NKP, dtype, (cudaStream_t)stream);
  MU_REQUIRE(x_layout == MU_X_TOKEN_MAJOR && dtype == MU_BF16, MU_ERR_BAD_DTYPE,
             "mu_qkv_project: token-major x is supported for MU_BF16 only (tcgen05 path)");
  MU_PTRS("mu_qkv_project", w_qkv_lp);
  MU_REQUIRE(device_cc_major() == 10, MU_ERR_ARCH, "mu_qkv_project: tcgen05 path needs an sm_100 device");
  if ((rc = launch_qkv_project_sm100(x, w_qkv_lp, b_qkv, keep_rank, q, kc, vc, B, C, N, NKP, (cudaStream_t)stream)))
    return rc;
  return launch_zero_pad_rows(n_keep, kc, vc, B, C, NKP, dtype, (cudaStream_t)stream);
}

int mu_attn_fwd_cudacore(const void* q, const void* kc, const void* vc, const int32_t* n_keep, void* o, float* lse,
                         int32_t B, int32_t N, int32_t NKP, int32_t C, int32_t dtype, mu_stream_t stream) {
  int rc;
  if ((rc = check_common("mu_attn_fwd", B, C, N, dtype))) return rc;
  if ((rc = check_nkp("mu_attn_fwd", N, NKP))) return rc;
  MU_PTRS("mu_attn_fwd", q, kc, vc, n_keep, o, lse);
  return launch_attn_fwd_simt(q, kc, vc, n_keep, o, lse, B, N, NKP, C, dtype, (cudaStream_t)stream);
}

int mu_attn_fwd(const void* q, const void* kc, const void* vc, const int32_t* n_keep, void* o, float* lse, int32_t B,
                int32_t N, int32_t NKP, int32_t C, int32_t dtype, mu_stream_t stream) {
  if (dtype != MU_BF16) return mu_attn_fwd_cudacore(q, kc, vc, n_keep, o, lse, B, N, NKP, C, dtype, stream);
  int rc;
  if ((rc = check_common("mu_attn_fwd", B, C, N, dtype))) return rc;
  if ((rc = check_nkp("mu_attn_fwd", N, NKP))) return rc;
  MU_PTRS("mu_attn_fwd", q, kc, vc, n_keep, o, lse);
  MU_REQUIRE(device_cc_major() == 10, MU_ERR_ARCH,
             "mu_attn_fwd: the bf16 path is tcgen05-only and needs an sm_100 device (found cc major %d)",
             device_cc_major());
  return launch_attn_fwd_sm100(q, kc, vc, n_keep, o, lse, B, N, NKP, C, (cudaStream_t)stream);
}

int mu_residual_ln_fwd(const void* o, const void* x, const float* gamma, const float* beta, float eps, void* y,
                       float* mean, float* rstd, int32_t B, int32_t C, int32_t N, int32_t dtype, int32_t x_layout,
                       mu_stream_t stream) {
  int rc;
  if ((rc = check_common("mu_residual_ln_fwd", B, C, N, dtype))) return rc;
  MU_PTRS("mu_residual_ln_fwd", o, x, gamma, beta, y, mean, rstd);
  return launch_residual_ln_fwd(o, x, gamma, beta, eps, y, mean, rstd, B, C, N, dtype,
                                x_layout == MU_X_TOKEN_MAJOR_VIEW ? 2 : x_layout == MU_X_TOKEN_MAJOR, (cudaStream_t)stream);
}

int mu_residual_ln_bwd(const void* dy, const void* o, const void* x, const float* mean, const float* rstd,
                       const float* gamma, void* dz, float* delta, float* dgamma, float* dbeta, int32_t B, int32_t C,
                       int32_t N, int32_t dtype, int32_t x_layout, mu_stream_t stream) {
  int rc;
  if ((rc = check_common("mu_residual_ln_bwd", B, C, N, dtype))) return rc;
  MU_PTRS("mu_residual_ln_bwd", dy, o, x, mean, rstd, gamma, dz, delta, dgamma, dbeta);
  return launch_residual_ln_bwd(dy, o, x, mean, rstd, gamma, dz, delta, dgamma, dbeta, B, C, N, dtype,
                                x_layout == MU_X_TOKEN_MAJOR_VIEW ? 2 : x_layout == MU_X_TOKEN_MAJOR, (cudaStream_t)stream);
}

int mu_attn_bwd_cudacore(const void* q, const void* kc, const void* vc, const int32_t* n_keep,
                         const int32_t* keep_idx, const void* d_o, const float* lse, const float* delta, void* dq,
                         void* dk, void* dv, int32_t B, int32_t N, int32_t NKP, int32_t C, int32_t dtype,
                         mu_stream_t stream) {
  int rc;
  if ((rc = check_common("mu_attn_bwd", B, C, N, dtype))) return rc;
  if ((rc = check_nkp("mu_attn_bwd", N, NKP))) return rc;
  MU_PTRS("mu_attn_bwd", q, kc, vc, n_keep, keep_idx, d_o, lse, delta, dq, dk, dv);
  return launch_attn_bwd_simt(q, kc, vc, n_keep, keep_idx, d_o, lse, delta, dq, dk, dv, B, N, NKP, C, dtype,
                              (cudaStream_t)stream);
}

size_t mu_attn_bwd_workspace_bytes(int32_t B, int32_t N, int32_t C, int32_t dtype) {
  return dtype == MU_BF16 ? attn_bwd_sm100_workspace(B, N, C) : 0;
}

int mu_attn_bwd(const void* q, const void* kc, const void* vc, const int32_t* n_keep, const int32_t* keep_idx,
                const void* d_o, const float* lse, const float* delta, void* dq, void* dkc, void* dvc, void* workspace,
                size_t workspace_bytes, int32_t B, int32_t N, int32_t NKP, int32_t C, int32_t dtype,
                mu_stream_t stream) {
  if (dtype != MU_BF16)
    return mu_attn_bwd_cudacore(q, kc, vc, n_keep, keep_idx, d_o, lse, delta, dq, dkc, dvc, B, N, NKP, C, dtype,
                                stream);
  int rc;
  if ((rc = check_common("mu_attn_bwd", B, C, N, dtype))) return rc;
  if ((rc = check_nkp("mu_attn_bwd", N, NKP))) return rc;
  MU_PTRS("mu_attn_bwd", q, kc, vc, n_keep, keep_idx, d_o, lse, delta, dq, dkc, dvc);
  MU_REQUIRE(device_cc_major() == 10, MU_ERR_ARCH,
             "mu_attn_bwd: the bf16 path is tcgen05-only and needs an sm_100 device (found cc major %d)",
             device_cc_major());
  return launch_attn_bwd_sm100(q, kc, vc, n_keep, keep_idx, d_o, lse, delta, dq, dkc, dvc, workspace, workspace_bytes,
                               B, N, NKP, C, (cudaStream_t)stream);
}

int mu_qkv_project_bwd(const void* x, const void* dz, const void* dq, const void* dk, const void* dv,
                       const float* w_qkv, const void* w_qkv_lp, void* dx, float* dw_qkv, float* db_qkv, int32_t B,
                       int32_t C, int32_t N, int32_t dtype, int32_t x_layout, mu_stream_t stream) {
  int rc;
  if ((rc = check_common("mu_qkv_project_bwd", B, C, N, dtype))) return rc;
  MU_PTRS("mu_qkv_project_bwd", x, dz, dq, dk, dv, w_qkv, dx, dw_qkv, db_qkv);
  if (x_layout == MU_X_CHANNEL_MAJOR)
    return launch_qkv_project_bwd(x, dz, dq, dk, dv, nullptr, w_qkv, dx, dw_qkv, db_qkv, B, C, N, 0, dtype,
                                  (cudaStream_t)stream);
  MU_REQUIRE(x_layout == MU_X_TOKEN_MAJOR && dtype == MU_BF16, MU_ERR_BAD_DTYPE,
             "mu_qkv_project_bwd: token-major x is supported for MU_BF16 only (tcgen05 path)");
  MU_PTRS("mu_qkv_project_bwd", w_qkv_lp);
  MU_REQUIRE(device_cc_major() == 10, MU_ERR_ARCH, "mu_qkv_project_bwd: tcgen05 path needs an sm_100 device");
  return launch_qkv_project_bwd_sm100(x, dz, dq, dk, dv, w_qkv_lp, dx, dw_qkv, db_qkv, B, C, N, (cudaStream_t)stream);
}

int mu_transpose(const void* in, void* out, int32_t batch, int32_t rows, int32_t cols, int32_t elem_bytes,
                 mu_stream_t stream) {
  MU_REQUIRE(batch > 0 && rows > 0 && cols > 0, MU_ERR_BAD_SHAPE, "mu_transpose: bad shape");
  MU_PTRS("mu_transpose", in, out);
  return launch_transpose(in, out, batch, rows, cols, elem_bytes, (cudaStream_t)stream);
}

int mu_bn_act_fwd(const void* x, const void* r, const float* gamma, const float* beta, float* running_mean,
                  float* running_var, float momentum, float eps, void* y, float* mean, float* rstd, float* a,
                  float* b, float* sums, int64_t M, int32_t C, int32_t act, int32_t dtype, mu_stream_t stream) {
  MU_REQUIRE(dtype == MU_F32 || dtype == MU_BF16, MU_ERR_BAD_DTYPE, "mu_bn_act_fwd: unknown dtype code %d", dtype);
  MU_PTRS("mu_bn_act_fwd", x, gamma, beta, y, mean, rstd, a, b, sums);
  return launch_bn_forward(x, r, gamma, beta, running_mean, running_var, momentum, eps, y, mean, rstd, a, b, sums,
                           (long)M, C, act, dtype, (cudaStream_t)stream);
}

int mu_bn_act_apply(const void* x, const void* r, const float* a, const float* b, void* y, int64_t M, int32_t C,
                    int32_t act, int32_t dtype, mu_stream_t stream) {
  MU_REQUIRE(dtype == MU_F32 || dtype == MU_BF16, MU_ERR_BAD_DTYPE, "mu_bn_act_apply: unknown dtype code %d", dtype);
  MU_PTRS("mu_bn_act_apply", x, a, b, y);
  return launch_bn_apply(x, r, a, b, y, (long)M, C, act, dtype, (cudaStream_t)stream);
}

int mu_bn_act_bwd(const void* dy, const void* x, const void* r, const float* a, const float* b, const float* mean,
                  const float* rstd, float* sums, void* dx, void* dr, int64_t M, int32_t C, int32_t act,
                  int32_t dtype, mu_stream_t stream) {
  MU_REQUIRE(dtype == MU_F32 || dtype == MU_BF16, MU_ERR_BAD_DTYPE, "mu_bn_act_bwd: unknown dtype code %d", dtype);
  MU_PTRS("mu_bn_act_bwd", dy, x, a, b, mean, rstd, sums, dx);
  MU_REQUIRE((r == nullptr) == (dr == nullptr), MU_ERR_NULL, "mu_bn_act_bwd: r and dr must both be given or both NULL");
  return launch_bn_backward(dy, x, r, a, b, mean, rstd, sums, dx, dr, (long)M, C, act, dtype, (cudaStream_t)stream);
}

#define MU_BF16_ONLY(fn) \
  MU_REQUIRE(dtype == MU_BF16, MU_ERR_BAD_DTYPE, fn ": this tcgen05 entry point takes MU_BF16 activations (got %d)", dtype)
#define MU_SM100_ONLY(fn) \
  MU_REQUIRE(device_cc_major() == 10, MU_ERR_ARCH, fn ": needs an sm_100 device (tcgen05 / TMEM)")

#define MU_DTYPE_OK(fn) \
  MU_REQUIRE(dtype == MU_F32 || dtype == MU_BF16, MU_ERR_BAD_DTYPE, fn ": unknown dtype code %d", dtype)

int mu_maxpool2(const void* x, const void* dy, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t bwd,
                int32_t dtype, mu_stream_t stream) {
  MU_DTYPE_OK("mu_maxpool2");
  MU_PTRS("mu_maxpool2", x, out);
  if (bwd) MU_PTRS("mu_maxpool2", dy);
  return launch_maxpool2(x, dy, out, B, H, W, C, bwd, dtype, (cudaStream_t)stream);
}

int mu_upsample_concat_fwd(const void* skip, const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t Cs,
                           int32_t Cx, int32_t dtype, mu_stream_t stream) {
  MU_DTYPE_OK("mu_upsample_concat_fwd");
  MU_PTRS("mu_upsample_concat_fwd", skip, x, out);
  return launch_upcat_fwd(skip, x, out, B, H, W, Cs, Cx, dtype, (cudaStream_t)stream);
}

int mu_upsample_concat_bwd(const void* dout, void* dskip, void* dx, int32_t B, int32_t H, int32_t W, int32_t Cs,
                           int32_t Cx, int32_t dtype, mu_stream_t stream) {
  MU_DTYPE_OK("mu_upsample_concat_bwd");
  MU_PTRS("mu_upsample_concat_bwd", dout, dskip, dx);
  return launch_upcat_bwd(dout, dskip, dx, B, H, W, Cs, Cx, dtype, (cudaStream_t)stream);
}

int mu_sample_layernorm_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, float* mean,
                            float* rstd, float* sums, int32_t B, int64_t L, int32_t dtype, mu_stream_t stream) {
  MU_DTYPE_OK("mu_sample_layernorm_fwd");
  MU_PTRS("mu_sample_layernorm_fwd", x, gamma, beta, y, mean, rstd, sums);
  return launch_sample_ln_fwd(x, gamma, beta, eps, y, mean, rstd, sums, B, (long)L, dtype, (cudaStream_t)stream);
}

int mu_sample_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                            float* sums, void* dx, float* dgamma, float* dbeta, int32_t B, int64_t L, int32_t dtype,
                            mu_stream_t stream) {
  MU_DTYPE_OK("mu_sample_layernorm_bwd");
  MU_PTRS("mu_sample_layernorm_bwd", dy, x, gamma, mean, rstd, sums, dx, dgamma, dbeta);
  return launch_sample_ln_bwd(dy, x, gamma, mean, rstd, sums, dx, dgamma, dbeta, B, (long)L, dtype,
                              (cudaStream_t)stream);
}

int mu_cross_entropy_fused(const void* logits, const int64_t* labels, const float* valid_count, int64_t ignore_index,
                           void* dlogits, float* loss_sum, int64_t M, int32_t C, int32_t pitch, int32_t dtype,
                           mu_stream_t stream) {
  MU_DTYPE_OK("mu_cross_entropy_fused");
  MU_PTRS("mu_cross_entropy_fused", logits, labels, valid_count, dlogits, loss_sum);
  return launch_ce_fused(logits, labels, valid_count, (long)ignore_index, dlogits, loss_sum, (long)M, C, pitch, dtype,
                         (cudaStream_t)stream);
}

int mu_argmax_iou(const void* logits, const int64_t* labels, int64_t* pred, int32_t* hist, float* miou, int64_t M,
                  int32_t C, int32_t pitch, float smooth, int32_t dtype, mu_stream_t stream) {
  MU_DTYPE_OK("mu_argmax_iou");
  MU_PTRS("mu_argmax_iou", logits);
  MU_REQUIRE(pred != nullptr || labels != nullptr, MU_ERR_NULL, "mu_argmax_iou: nothing to produce (pred and labels NULL)");
  MU_REQUIRE(labels == nullptr || hist != nullptr, MU_ERR_NULL, "mu_argmax_iou: labels given without a histogram buffer");
  return launch_argmax_iou(logits, labels, pred, hist, miou, (long)M, C, pitch, smooth, dtype, (cudaStream_t)stream);
}

int mu_column_sums(const void* x, float* sums, int64_t M, int32_t C, int32_t dtype, mu_stream_t stream) {
  MU_DTYPE_OK("mu_column_sums");
  MU_PTRS("mu_column_sums", x, sums);
  return launch_column_sums(x, sums, (long)M, C, dtype, (cudaStream_t)stream);
}

int mu_conv1x1_prep(const float* w, const float* bias, void* wf, void* wd, float* bias_p, int32_t Cout, int32_t Cin,
                    int32_t Np, mu_stream_t stream) {
  MU_REQUIRE(Cout > 0 && Cin > 0 && Np >= Cout, MU_ERR_BAD_SHAPE, "mu_conv1x1_prep: bad shape (Cout=%d Cin=%d Np=%d)",
             Cout, Cin, Np);
  MU_PTRS("mu_conv1x1_prep", w, wf, wd, bias_p);
  return launch_conv1x1_prep(w, bias, wf, wd, bias_p, Cout, Cin, Np, (cudaStream_t)stream);
}

int mu_conv1x1_fwd(const void* x, const void* wf, const float* bias_p, void* y, int32_t B, int32_t H, int32_t W,
                   int32_t Cin, int32_t Np, int32_t dtype, mu_stream_t stream) {
  MU_BF16_ONLY("mu_conv1x1_fwd");
  MU_PTRS("mu_conv1x1_fwd", x, wf, y);
  MU_SM100_ONLY("mu_conv1x1_fwd");
  return launch_conv1x1_fprop_sm100(x, wf, bias_p, y, B, H, W, Cin, Np, (cudaStream_t)stream);
}

int mu_conv1x1_fwd_stats(const void* x, const void* wf, const float* bias_p, void* y, float* stats, int32_t B, int32_t H,
                         int32_t W, int32_t Cin, int32_t Np, int32_t dtype, mu_stream_t stream) {
  MU_BF16_ONLY("mu_conv1x1_fwd_stats");
  MU_PTRS("mu_conv1x1_fwd_stats", x, wf, y, stats);
  MU_SM100_ONLY("mu_conv1x1_fwd_stats");
  return launch_conv1x1_fprop_sm100(x, wf, bias_p, y, B, H, W, Cin, Np, (cudaStream_t)stream, stats);
}

int mu_conv1x1_bwd_data(const void* dy, const void* wd, void* dx, int32_t B, int32_t H, int32_t W, int32_t Cin,
                        int32_t Np, int32_t dtype, mu_stream_t stream) {
  MU_BF16_ONLY("mu_conv1x1_bwd_data");
  MU_PTRS("mu_conv1x1_bwd_data", dy, wd, dx);
  MU_SM100_ONLY("mu_conv1x1_bwd_data");
  return launch_conv1x1_fprop_sm100(dy, wd, nullptr, dx, B, H, W, Np, Cin, (cudaStream_t)stream);
}

size_t mu_conv1x1_workspace_bytes(int32_t Cin, int32_t Np) { return (size_t)Cin * Np * sizeof(float); }

int mu_conv1x1_bwd_weight(const void* x, const void* dy, void* workspace, size_t workspace_bytes, float* dw, int32_t B,
                          int32_t H, int32_t W, int32_t Cin, int32_t Np, int32_t dtype, mu_stream_t stream) {
  MU_BF16_ONLY("mu_conv1x1_bwd_weight");
  MU_PTRS("mu_conv1x1_bwd_weight", x, dy, workspace, dw);
  MU_REQUIRE(workspace_bytes >= mu_conv1x1_workspace_bytes(Cin, Np), MU_ERR_WORKSPACE,
             "mu_conv1x1_bwd_weight: workspace too small (%zu < %zu)", workspace_bytes,
             mu_conv1x1_workspace_bytes(Cin, Np));
  MU_SM100_ONLY("mu_conv1x1_bwd_weight");
  return launch_conv1x1_wgrad_sm100(x, dy, (float*)workspace, dw, B, H, W, Cin, Np, (cudaStream_t)stream);
}

int mu_conv1x1_bwd_weight_bias(const void* x, const void* dy, void* workspace, size_t workspace_bytes, float* dw,
                               float* db, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Np, int32_t dtype,
                               mu_stream_t stream) {
  MU_BF16_ONLY("mu_conv1x1_bwd_weight_bias");
  MU_PTRS("mu_conv1x1_bwd_weight_bias", x, dy, workspace, dw, db);
  MU_REQUIRE(workspace_bytes >= mu_conv1x1_workspace_bytes(Cin, Np) && workspace_bytes >= 2 * (size_t)Np * sizeof(float),
             MU_ERR_WORKSPACE, "mu_conv1x1_bwd_weight_bias: workspace too small (%zu < %zu)", workspace_bytes,
             mu_conv1x1_workspace_bytes(Cin, Np));
  MU_SM100_ONLY("mu_conv1x1_bwd_weight_bias");
  int db_done = 0;
  int rc = launch_conv1x1_wgrad_sm100(x, dy, (float*)workspace, dw, B, H, W, Cin, Np, (cudaStream_t)stream, db, &db_done);
  if (rc || db_done) return rc;
  // not carried by the GEMM (deterministic mode, or more than 64 input channels): the ordered column sums of dy, through
  // the workspace (free again: the finish kernel of the weight gradient precedes this in the stream)
  rc = launch_column_sums(dy, (float*)workspace, (long)B * H * W, Np, MU_BF16, (cudaStream_t)stream);
  if (rc) return rc;
  cudaError_t e = cudaMemcpyAsync(db, workspace, (size_t)Np * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("mu_conv1x1_bwd_weight_bias: cudaMemcpyAsync: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int mu_conv_prep_weights(const float* w, void* wf, void* wd, int32_t Cout, int32_t Cin, int32_t taps,
                         mu_stream_t stream) {
  MU_REQUIRE(Cout > 0 && Cin > 0 && (taps == 9 || taps == 1), MU_ERR_BAD_SHAPE,
             "mu_conv_prep_weights: bad shape (Cout=%d Cin=%d taps=%d)", Cout, Cin, taps);
  MU_PTRS("mu_conv_prep_weights", w, wf);
  return launch_conv_prep_weights(w, wf, wd, Cout, Cin, taps, (cudaStream_t)stream);
}

int mu_conv3x3_fwd(const void* x, const void* wf, void* y, float* stats, int32_t B, int32_t H, int32_t W, int32_t Cin,
                   int32_t Cout, int32_t dtype, mu_stream_t stream) {
  MU_BF16_ONLY("mu_conv3x3_fwd");
  MU_PTRS("mu_conv3x3_fwd", x, wf, y);
  MU_SM100_ONLY("mu_conv3x3_fwd");
  return launch_conv_fprop_sm100(x, wf, y, stats, B, H, W, Cin, Cout, 9, (cudaStream_t)stream);
}

int mu_conv3x3_bwd_data(const void* dy, const void* wd, void* dx, int32_t B, int32_t H, int32_t W, int32_t Cin,
                        int32_t Cout, int32_t dtype, mu_stream_t stream) {
  MU_BF16_ONLY("mu_conv3x3_bwd_data");
  MU_PTRS("mu_conv3x3_bwd_data", dy, wd, dx);
  MU_SM100_ONLY("mu_conv3x3_bwd_data");
  return launch_conv_fprop_sm100(dy, wd, dx, nullptr, B, H, W, Cout, Cin, 9, (cudaStream_t)stream);
}

int mu_conv3x3_bwd_data_acc(const void* dy, const void* wd, void* dx, int32_t B, int32_t H, int32_t W, int32_t Cin,
                            int32_t Cout, int32_t dtype, mu_stream_t stream) {
  MU_BF16_ONLY("mu_conv3x3_bwd_data_acc");
  MU_PTRS("mu_conv3x3_bwd_data_acc", dy, wd, dx);
  MU_SM100_ONLY("mu_conv3x3_bwd_data_acc");
  return launch_conv_fprop_sm100(dy, wd, dx, nullptr, B, H, W, Cout, Cin, 9, (cudaStream_t)stream, 1);
}

size_t mu_conv3x3_workspace_bytes(int32_t Cin, int32_t Cout) { return (size_t)9 * Cin * Cout * sizeof(float); }

int mu_conv3x3_bwd_weight(const void* x, const void* dy, void* workspace, size_t workspace_bytes, float* dw, int32_t B,
                          int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t dtype, mu_stream_t stream) {
  MU_BF16_ONLY("mu_conv3x3_bwd_weight");
  MU_PTRS("mu_conv3x3_bwd_weight", x, dy, workspace, dw);
  MU_REQUIRE(workspace_bytes >= mu_conv3x3_workspace_bytes(Cin, Cout), MU_ERR_WORKSPACE,
             "mu_conv3x3_bwd_weight: workspace too small (%zu < %zu)", workspace_bytes,
             mu_conv3x3_workspace_bytes(Cin, Cout));
  MU_SM100_ONLY("mu_conv3x3_bwd_weight");
  return launch_conv_wgrad_sm100(x, dy, (float*)workspace, dw, B, H, W, Cin, Cout, (cudaStream_t)stream);
}

int mu_bn_update_running(float* running_mean, float* running_var, const float* mean, const float* rstd, float momentum,
                         float eps, int64_t M, int32_t C, mu_stream_t stream) {
  MU_REQUIRE(C > 0 && M > 0, MU_ERR_BAD_SHAPE, "mu_bn_update_running: bad shape (M=%ld C=%d)", (long)M, C);
  MU_PTRS("mu_bn_update_running", running_mean, running_var, mean, rstd);
  return launch_bn_update_running(running_mean, running_var, mean, rstd, momentum, eps, (long)M, C, (cudaStream_t)stream);
}

int mu_bn_act_fwd_stats(const void* x, const void* r, const float* gamma, const float* beta, float eps, void* y,
                        float* mean, float* rstd, float* a, float* b, const float* sums, int64_t M, int32_t C,
                        int32_t act, int32_t dtype, mu_stream_t stream) {
  MU_DTYPE_OK("mu_bn_act_fwd_stats");
  MU_PTRS("mu_bn_act_fwd_stats", x, gamma, beta, y, mean, rstd, a, b, sums);
  return launch_bn_forward_stats(x, r, gamma, beta, eps, y, mean, rstd, a, b, sums, (long)M, C, act, dtype,
                                 (cudaStream_t)stream);
}

int mu_instance_triplet_fwd(const void* sem, const int64_t* sem_strides, int32_t B, int32_t C, int32_t H, int32_t W,
                            const int64_t* order, const int64_t* meta, int32_t K, float margin, float eps, int32_t* sel,
                            float* dist, float* loss, int32_t dtype, mu_stream_t stream) {
  MU_DTYPE_OK("mu_instance_triplet_fwd");
  MU_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && K >= 0, MU_ERR_BAD_SHAPE,
             "mu_instance_triplet_fwd: bad shape (B=%d C=%d H=%d W=%d K=%d)", B, C, H, W, K);
  MU_REQUIRE(sem != nullptr && sem_strides != nullptr && loss != nullptr, MU_ERR_NULL, "mu_instance_triplet_fwd: null pointer");
  MU_REQUIRE(K == 0 || (order != nullptr && meta != nullptr && sel != nullptr && dist != nullptr), MU_ERR_NULL,
             "mu_instance_triplet_fwd: null pointer (order / meta / sel / dist)");
  const long st[4] = {(long)sem_strides[0], (long)sem_strides[1], (long)sem_strides[2], (long)sem_strides[3]};
  return launch_instance_triplet_fwd(sem, st, B, C, H, W, order, meta, K, margin, eps, sel, dist, loss, dtype,
                                     (cudaStream_t)stream);
}

int mu_instance_triplet_bwd(const void* sem, const int64_t* sem_strides, int32_t B, int32_t C, const int32_t* sel,
                            int32_t K, float margin, float eps, const float* dist, const float* dloss, float scale,
                            void* dsem, const int64_t* dsem_strides, int32_t dtype, mu_stream_t stream) {
  MU_DTYPE_OK("mu_instance_triplet_bwd");
  MU_REQUIRE(B > 0 && C > 0 && K >= 0, MU_ERR_BAD_SHAPE, "mu_instance_triplet_bwd: bad shape (B=%d C=%d K=%d)", B, C, K);
  MU_REQUIRE(sem != nullptr && sem_strides != nullptr && dsem != nullptr && dsem_strides != nullptr, MU_ERR_NULL,
             "mu_instance_triplet_bwd: null pointer");
  MU_REQUIRE(K == 0 || (sel != nullptr && dist != nullptr), MU_ERR_NULL, "mu_instance_triplet_bwd: null pointer (sel / dist)");
  const long st[4] = {(long)sem_strides[0], (long)sem_strides[1], (long)sem_strides[2], (long)sem_strides[3]};
  const long gst[4] = {(long)dsem_strides[0], (long)dsem_strides[1], (long)dsem_strides[2], (long)dsem_strides[3]};
  return launch_instance_triplet_bwd(sem, st, B, C, sel, K, margin, eps, dist, dloss, scale, dsem, gst, dtype,
                                     (cudaStream_t)stream);
}

int mu_query_mask_bits(const void* qe, const void* feat, int32_t B, int32_t Q, int32_t N, int32_t C, uint32_t* bits,
                       uint32_t* bits_t, int32_t* row_count, float* logits, int32_t dtype, mu_stream_t stream) {
  MU_BF16_ONLY("mu_query_mask_bits");
  MU_REQUIRE(B > 0 && Q > 0 && N > 0, MU_ERR_BAD_SHAPE, "mu_query_mask_bits: bad shape (B=%d Q=%d N=%d)", B, Q, N);
  MU_PTRS("mu_query_mask_bits", qe, feat, bits, bits_t, row_count);
  MU_SM100_ONLY("mu_query_mask_bits");
  return launch_query_mask_bits_sm100(qe, feat, B, Q, N, C, bits, bits_t, row_count, logits, (cudaStream_t)stream);
}

static int check_query_attn(const char* fn, int BH, int heads, int Q, int N, int NKP) {
  MU_REQUIRE(BH > 0 && heads > 0 && BH % heads == 0 && Q > 0 && N > 0, MU_ERR_BAD_SHAPE,
             "%s: bad shape (BH=%d heads=%d Q=%d N=%d)", fn, BH, heads, Q, N);
  return check_nkp(fn, N, NKP);
}

int mu_query_attn_fwd(const void* q, const void* k, const void* v, const uint32_t* bits, void* o, float* lse, int32_t BH,
                      int32_t heads, int32_t Q, int32_t N, int32_t NKP, int32_t D, float scale, int32_t dtype,
                      mu_stream_t stream) {
  MU_BF16_ONLY("mu_query_attn_fwd");
  if (int rc = check_query_attn("mu_query_attn_fwd", BH, heads, Q, N, NKP)) return rc;
  MU_PTRS("mu_query_attn_fwd", q, k, v, bits, o, lse);
  MU_SM100_ONLY("mu_query_attn_fwd");
  return launch_query_attn_fwd_sm100(q, k, v, bits, o, lse, BH, heads, Q, N, NKP, D, scale, (cudaStream_t)stream);
}

size_t mu_query_attn_bwd_workspace_bytes(int32_t BH, int32_t Q, int32_t D) { return attn_bwd_sm100_workspace(BH, Q, D); }

int mu_query_attn_bwd(const void* q, const void* k, const void* v, const uint32_t* bits_t, const void* d_o,
                      const float* lse, const float* delta, void* dq, void* dk, void* dv, void* workspace,
                      size_t workspace_bytes, int32_t BH, int32_t heads, int32_t Q, int32_t N, int32_t NKP, int32_t D,
                      float scale, int32_t dtype, mu_stream_t stream) {
  MU_BF16_ONLY("mu_query_attn_bwd");
  if (int rc = check_query_attn("mu_query_attn_bwd", BH, heads, Q, N, NKP)) return rc;
  MU_PTRS("mu_query_attn_bwd", q, k, v, bits_t, d_o, lse, delta, dq, dk, dv, workspace);
  MU_SM100_ONLY("mu_query_attn_bwd");
  return launch_query_attn_bwd_sm100(q, k, v, bits_t, d_o, lse, delta, dq, dk, dv, workspace, workspace_bytes, BH, heads,
                                     Q, N, NKP, D, scale, (cudaStream_t)stream);
}

int mu_to_tensor_u8(const uint8_t* img, void* out, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cpad,
                    int32_t channels_last, int32_t dtype, mu_stream_t stream) {
  MU_DTYPE_OK("mu_to_tensor_u8");
  MU_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && (!channels_last || Cpad >= Cin), MU_ERR_BAD_SHAPE,
             "mu_to_tensor_u8: bad shape (B=%d H=%d W=%d Cin=%d Cpad=%d)", B, H, W, Cin, Cpad);
  MU_REQUIRE(img != nullptr && out != nullptr, MU_ERR_NULL, "mu_to_tensor_u8: null pointer");
  return launch_to_tensor_u8(img, out, B, H, W, Cin, Cpad, channels_last, dtype, (cudaStream_t)stream);
}

int mu_resize_linear_to_tensor_u8(const uint8_t* img, void* out, int32_t src_h, int32_t src_w, int32_t Cin, int32_t out_h,
                                  int32_t out_w, int32_t Cpad, int32_t channels_last, int32_t normalise, int32_t dtype,
                                  mu_stream_t stream) {
  MU_DTYPE_OK("mu_resize_linear_to_tensor_u8");
  MU_REQUIRE(src_h > 0 && src_w > 0 && out_h > 0 && out_w > 0 && Cin > 0 && Cin <= 4 && Cpad >= Cin, MU_ERR_BAD_SHAPE,
             "mu_resize_linear_to_tensor_u8: bad shape (src %dx%d out %dx%d Cin=%d Cpad=%d)", src_h, src_w, out_h, out_w,
             Cin, Cpad);
  MU_REQUIRE(img != nullptr && out != nullptr, MU_ERR_NULL, "mu_resize_linear_to_tensor_u8: null pointer");
  return launch_resize_linear_to_tensor(img, out, src_h, src_w, Cin, out_h, out_w, Cpad, channels_last, normalise, dtype,
                                        (cudaStream_t)stream);
}

int mu_resize_nearest_u8_i64(const uint8_t* mask, int64_t* out, int32_t src_h, int32_t src_w, int32_t out_h,
                             int32_t out_w, mu_stream_t stream) {
  MU_REQUIRE(src_h > 0 && src_w > 0 && out_h > 0 && out_w > 0, MU_ERR_BAD_SHAPE,
             "mu_resize_nearest_u8_i64: bad shape (src %dx%d out %dx%d)", src_h, src_w, out_h, out_w);
  MU_REQUIRE(mask != nullptr && out != nullptr, MU_ERR_NULL, "mu_resize_nearest_u8_i64: null pointer");
  return launch_resize_nearest_u8_i64(mask, out, src_h, src_w, out_h, out_w, (cudaStream_t)stream);
}

}  // extern "C"
