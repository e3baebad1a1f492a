// Shared host/device helpers for the maskunet_b200 library.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <mutex>
#include <unordered_map>

#include "../../include/maskunet_b200.h"

namespace mu {

// thread-local error text behind mu_last_error()
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // returns 0 or the positive cudaError_t of the last launch

#define MU_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::mu::set_error(__VA_ARGS__);  \
      return (code);                 \
    }                                \
  } while (0)

// One-time cudaFuncSetAttribute(MaxDynamicSharedMemorySize) per (kernel, device): the attribute is sticky, and the call
// takes a driver lock on every launch otherwise.  Idempotent, so a race between two first callers is harmless.
template <typename Kern>
inline cudaError_t set_max_dynamic_smem_once(Kern kern, int bytes) {
  static std::mutex mu;
  static std::unordered_map<uint64_t, int> done;      // (kernel pointer ^ device) -> bytes set
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t key = (uint64_t)reinterpret_cast<uintptr_t>(reinterpret_cast<const void*>(kern)) ^ ((uint64_t)dev << 56);
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return cudaSuccess;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> lock(mu);
    done[key] = bytes;
  }
  return e;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---- element load/store with conversion (T = float or __nv_bfloat16) ----
template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- launchers implemented in the .cu files (dtype-dispatched inside) ----
int launch_mask_binarize(const int64_t* bits, int B, int N, uint32_t* keep_bits, int32_t* n_keep, int32_t* keep_idx,
                         int32_t* keep_rank, cudaStream_t s);
int launch_qkv_project(const void* x, const float* w, const float* b, const int32_t* rank, const int32_t* n_keep,
                       void* q, void* kc, void* vc, int B, int C, int N, int NKP, int dtype, cudaStream_t s);
int launch_attn_fwd_simt(const void* q, const void* kc, const void* vc, const int32_t* n_keep, void* o, float* lse,
                         int B, int N, int NKP, int C, int dtype, cudaStream_t s);
int launch_attn_bwd_simt(const void* q, const void* kc, const void* vc, const int32_t* n_keep,
                         const int32_t* keep_idx, const void* d_o, const float* lse, const float* delta, void* dq,
                         void* dk, void* dv, int B, int N, int NKP, int C, int dtype, cudaStream_t s);
int launch_residual_ln_fwd(const void* o, const void* x, const float* gamma, const float* beta, float eps, void* y,
                           float* mean, float* rstd, int B, int C, int N, int dtype, int tok, cudaStream_t s);
int launch_residual_ln_bwd(const void* dy, const void* o, const void* x, const float* mean, const float* rstd,
                           const float* gamma, void* dz, float* delta, float* dgamma, float* dbeta, int B, int C, int N,
                           int dtype, int tok, cudaStream_t s);
int launch_qkv_project_bwd(const void* x, const void* dz, const void* dq, const void* dkc, const void* dvc,
                           const int32_t* rank, const float* w, void* dx, float* dw, float* db, int B, int C, int N,
                           int NKP, int dtype, cudaStream_t s);
// tcgen05 path (bf16 only)
int launch_attn_fwd_sm100(const void* q, const void* kc, const void* vc, const int32_t* n_keep, void* o, float* lse,
                          int B, int N, int NKP, int C, cudaStream_t s);
int launch_attn_bwd_sm100(const void* q, const void* kc, const void* vc, const int32_t* n_keep,
                          const int32_t* keep_idx, const void* d_o, const float* lse, const float* delta, void* dq,
                          void* dk, void* dv, void* workspace, size_t workspace_bytes, int B, int N, int NKP, int C,
                          cudaStream_t s);
size_t attn_bwd_sm100_workspace(int B, int N, int C);
void set_deterministic(int on);   // fixed-order accumulation in every kernel that sums across CTAs (attn_bwd_sm100.cu)
int get_deterministic();
int launch_qkv_project_sm100(const void* xt, const void* w_bf16, const float* bias, const int32_t* rank, void* q,
                             void* kc, void* vc, int B, int C, int N, int NKP, cudaStream_t s);
int launch_qkv_project_bwd_sm100(const void* xt, const void* dz, const void* dq, const void* dk, const void* dv,
                                 const void* w_bf16, void* dxt, float* dw, float* db, int B, int C, int N,
                                 cudaStream_t s);
int launch_zero_pad_rows(const int32_t* n_keep, void* kc, void* vc, int B, int C, int NKP, int dtype, cudaStream_t s);
int launch_bn_forward(const void* x, const void* r, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, float momentum, float eps, void* y, float* mean, float* rstd, float* a,
                      float* b, float* sums, long M, int C, int act, int dtype, cudaStream_t s);
int launch_bn_apply(const void* x, const void* r, const float* a, const float* b, void* y, long M, int C, int act,
                    int dtype, cudaStream_t s);
int launch_bn_backward(const void* dy, const void* x, const void* r, const float* a, const float* b, const float* mean,
                       const float* rstd, float* sums, void* dx, void* dr, long M, int C, int act, int dtype,
                       cudaStream_t s);
int launch_bn_update_running(float* running_mean, float* running_var, const float* mean, const float* rstd,
                             float momentum, float eps, long M, int C, cudaStream_t s);
int launch_bn_forward_stats(const void* x, const void* r, const float* gamma, const float* beta, float eps, void* y,
                            float* mean, float* rstd, float* a, float* b, const float* sums, long M, int C, int act,
                            int dtype, cudaStream_t s);
// K7 (tcgen05, bf16 channels-last)
int launch_conv_fprop_sm100(const void* x, const void* wt, void* y, float* stats, int B, int H, int W, int K, int N,
                            int taps, cudaStream_t s, int accum = 0);
int launch_conv_wgrad_sm100(const void* x, const void* dy, float* ws, float* dw, int B, int H, int W, int Cin, int Cout,
                            cudaStream_t s);
int launch_conv_prep_weights(const float* w, void* wf, void* wd, int Cout, int Cin, int taps, cudaStream_t s);
int launch_maxpool2(const void* x, const void* dy, void* out, int B, int H, int W, int C, int bwd, int dtype,
                    cudaStream_t s);
int launch_upcat_fwd(const void* skip, const void* x, void* out, int B, int H, int W, int Cs, int Cx, int dtype,
                     cudaStream_t s);
int launch_upcat_bwd(const void* dout, void* dskip, void* dx, int B, int H, int W, int Cs, int Cx, int dtype,
                     cudaStream_t s);
int launch_sample_ln_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, float* mean,
                         float* rstd, float* sums, int B, long L, int dtype, cudaStream_t s);
int launch_sample_ln_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                         float* sums, void* dx, float* dgamma, float* dbeta, int B, long L, int dtype, cudaStream_t s);
int launch_ce_fused(const void* logits, const int64_t* labels, const float* valid_count, long ignore_index,
                    void* dlogits, float* loss_sum, long M, int C, int pitch, int dtype, cudaStream_t s);
int launch_conv1x1_fprop_sm100(const void* x, const void* wt, const float* bias, void* y, int B, int H, int W, int K,
                               int Np, cudaStream_t s, float* stats = nullptr);
int launch_conv1x1_wgrad_sm100(const void* x, const void* dy, float* ws, float* dw, int B, int H, int W, int Cin, int Np,
                               cudaStream_t s, float* db = nullptr, int* db_done = nullptr);
int launch_conv1x1_prep(const float* w, const float* bias, void* wf, void* wd, float* bias_p, int Cout, int Cin, int Np,
                        cudaStream_t s);
int launch_column_sums(const void* x, float* sums, long M, int C, int dtype, cudaStream_t s);
int launch_argmax_iou(const void* logits, const int64_t* labels, int64_t* pred, int32_t* hist, float* miou, long M, int C,
                      int pitch, float smooth, int dtype, cudaStream_t s);
int launch_instance_triplet_fwd(const void* sem, const long* st, int B, int C, int H, int W, const int64_t* order,
                                const int64_t* meta, int K, float margin, float eps, int32_t* sel, float* dist,
                                float* loss, int dtype, cudaStream_t s);
int launch_instance_triplet_bwd(const void* sem, const long* st, int B, int C, const int32_t* sel, int K, float margin,
                                float eps, const float* dist, const float* dloss, float scale, void* dsem,
                                const long* gst, int dtype, cudaStream_t s);
int launch_query_mask_bits_sm100(const void* qe, const void* feat, int B, int Q, int N, int C, uint32_t* bits,
                                 uint32_t* bits_t, int32_t* row_count, float* logits, cudaStream_t s);
int launch_query_attn_fwd_sm100(const void* q, const void* k, const void* v, const uint32_t* qbits, void* o, float* lse,
                                int BH, int heads, int Q, int N, int NKP, int D, float scale, cudaStream_t s);
int launch_query_attn_bwd_sm100(const void* q, const void* k, const void* v, const uint32_t* bits_t, const void* d_o,
                                const float* lse, const float* delta, void* dq, void* dk, void* dv, void* workspace,
                                size_t workspace_bytes, int BH, int heads, int Q, int N, int NKP, int D, float scale,
                                cudaStream_t s);
int launch_to_tensor_u8(const uint8_t* img, void* out, int B, int H, int W, int Cin, int Cpad, int channels_last,
                        int dtype, cudaStream_t s);
int launch_resize_linear_to_tensor(const uint8_t* img, void* out, int sh, int sw, int cin, int oh, int ow, int cpad,
                                   int channels_last, int normalise, int dtype, cudaStream_t s);
int launch_resize_nearest_u8_i64(const uint8_t* mask, int64_t* out, int sh, int sw, int oh, int ow, cudaStream_t s);
int launch_transpose(const void* in, void* out, int batch, int rows, int cols, int elem_bytes, cudaStream_t s);

}  // namespace mu
