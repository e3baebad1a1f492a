// K3-epilogue / K4: residual + LayerNorm over channels, forward and backward.
// Replaces `attention_output + x` and nn.LayerNorm([C]) at
// /root/reference/code/ade20k/ade_semantic.py:187-188 and their autograd (:400).
// x is NCHW ([B, C, N], token-contiguous) while o / y / dz are [B, N, C] (channel-contiguous), so the
// residual needs a transposed read of x: a [C][32 tokens] tile is staged through shared memory with
// coalesced 128-byte rows, then each warp normalises tokens with shuffle reductions.
// HBM-bound: forward moves 3*B*N*C*s bytes (o, x in; y out), backward 4*B*N*C*s (dy, o, x in; dz out).
#include "common.cuh"

namespace mu {

constexpr int kLnTokens = 32;   // tokens per CTA tile
constexpr int kLnThreads = 256; // 8 warps, 4 tokens each
constexpr int kMaxCPerLane = 8; // C <= 256

template <typename T>
__device__ __forceinline__ void load_x_tile(float* xs /*[C][33]*/, const T* xb, int C, int N, int n0) {
  for (int idx = threadIdx.x; idx < C * kLnTokens; idx += kLnThreads) {
    const int c = idx >> 5, t = idx & 31, n = n0 + t;
    xs[c * 33 + t] = (n < N) ? ld_f(xb + (size_t)c * N + n) : 0.f;
  }
}

// TOK = true: x is token-major [B, N, C] (NHWC activations), no transposed staging needed
template <typename T, bool TOK>
__global__ void __launch_bounds__(kLnThreads) residual_ln_fwd_kernel(const T* __restrict__ o, const T* __restrict__ x,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, float eps,
                                                                     T* __restrict__ y, float* __restrict__ mean,
                                                                     float* __restrict__ rstd, int C, int N) {
  extern __shared__ float xs[];
  const int b = blockIdx.y, n0 = blockIdx.x * kLnTokens;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (!TOK) {
    load_x_tile(xs, x + (size_t)b * C * N, C, N, n0);
    __syncthreads();
  }
  const int per = C / 32;
  for (int t = warp; t < kLnTokens; t += kLnThreads / 32) {
    const int n = n0 + t;
    if (n >= N) break;
    const T* orow = o + ((size_t)b * N + n) * C;
    const T* xrow = x + ((size_t)b * N + n) * C;
    float z[kMaxCPerLane];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxCPerLane; ++i) {
      if (i < per) {
        const int c = lane + 32 * i;
        z[i] = ld_f(orow + c) + (TOK ? ld_f(xrow + c) : xs[c * 33 + t]);
        s += z[i];
      }
    }
    const float mu_ = warp_sum(s) / C;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxCPerLane; ++i)
      if (i < per) {
        const float d = z[i] - mu_;
        v += d * d;
      }
    const float rs = rsqrtf(warp_sum(v) / C + eps);
    T* yrow = y + ((size_t)b * N + n) * C;
#pragma unroll
    for (int i = 0; i < kMaxCPerLane; ++i)
      if (i < per) {
        const int c = lane + 32 * i;
        st_f(yrow + c, (z[i] - mu_) * rs * gamma[c] + beta[c]);
      }
    if (lane == 0) {
      mean[(size_t)b * N + n] = mu_;
      rstd[(size_t)b * N + n] = rs;
    }
  }
}

template <typename T, bool TOK>
__global__ void __launch_bounds__(kLnThreads) residual_ln_bwd_kernel(
    const T* __restrict__ dy, const T* __restrict__ o, const T* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, T* __restrict__ dz, float* __restrict__ delta,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int C, int N) {
  extern __shared__ float xs[];  // [C][33] x tile, then reused for the dgamma/dbeta block reduction
  const int b = blockIdx.y, n0 = blockIdx.x * kLnTokens;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (!TOK) {
    load_x_tile(xs, x + (size_t)b * C * N, C, N, n0);
    __syncthreads();
  }
  const int per = C / 32;
  float dg[kMaxCPerLane] = {}, dbt[kMaxCPerLane] = {};
  for (int t = warp; t < kLnTokens; t += kLnThreads / 32) {
    const int n = n0 + t;
    if (n >= N) break;
    const size_t row = ((size_t)b * N + n) * C;
    const float mu_ = mean[(size_t)b * N + n], rs = rstd[(size_t)b * N + n];
    float ov[kMaxCPerLane], zh[kMaxCPerLane], g[kMaxCPerLane];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxCPerLane; ++i)
      if (i < per) {
        const int c = lane + 32 * i;
        ov[i] = ld_f(o + row + c);
        zh[i] = (ov[i] + (TOK ? ld_f(x + row + c) : xs[c * 33 + t]) - mu_) * rs;
        const float d = ld_f(dy + row + c);
        g[i] = d * gamma[c];
        dg[i] += d * zh[i];
        dbt[i] += d;
        s1 += g[i];
        s2 += g[i] * zh[i];
      }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    float dl = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxCPerLane; ++i)
      if (i < per) {
        const int c = lane + 32 * i;
        const float d = rs * (g[i] - s1 - zh[i] * s2);
        // delta uses the value backward will see (rounded to the storage type)
        T tmp;
        st_f(&tmp, d);
        st_f(dz + row + c, d);
        dl += ld_f(&tmp) * ov[i];
      }
    dl = warp_sum(dl);
    if (lane == 0) delta[(size_t)b * N + n] = dl;
  }
  __syncthreads();  // everyone is done with the x tile
  float* red = xs;  // [8 warps][C] x2
#pragma unroll
  for (int i = 0; i < kMaxCPerLane; ++i)
    if (i < per) {
      red[warp * C + lane + 32 * i] = dg[i];
      red[(8 + warp) * C + lane + 32 * i] = dbt[i];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kLnThreads) {
    float a = 0.f, bsum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += red[w * C + c];
      bsum += red[(8 + w) * C + c];
    }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, bsum);
  }
}

static size_t ln_smem(int C) { return sizeof(float) * (size_t)C * 33; }  // >= 16*C for C >= 32

template <typename T>
static int run_fwd(const void* o, const void* x, const float* gamma, const float* beta, float eps, void* y,
                   float* mean, float* rstd, int B, int C, int N, int tok, cudaStream_t s) {
  dim3 grid((N + kLnTokens - 1) / kLnTokens, B);
  if (tok)
    residual_ln_fwd_kernel<T, true><<<grid, kLnThreads, ln_smem(C), s>>>((const T*)o, (const T*)x, gamma, beta, eps,
                                                                        (T*)y, mean, rstd, C, N);
  else
    residual_ln_fwd_kernel<T, false><<<grid, kLnThreads, ln_smem(C), s>>>((const T*)o, (const T*)x, gamma, beta, eps,
                                                                         (T*)y, mean, rstd, C, N);
  return check_launch("residual_ln_fwd");
}
template <typename T>
static int run_bwd(const void* dy, const void* o, const void* x, const float* mean, const float* rstd,
                   const float* gamma, void* dz, float* delta, float* dgamma, float* dbeta, int B, int C, int N,
                   int tok, cudaStream_t s) {
  dim3 grid((N + kLnTokens - 1) / kLnTokens, B);
  if (tok)
    residual_ln_bwd_kernel<T, true><<<grid, kLnThreads, ln_smem(C), s>>>(
        (const T*)dy, (const T*)o, (const T*)x, mean, rstd, gamma, (T*)dz, delta, dgamma, dbeta, C, N);
  else
    residual_ln_bwd_kernel<T, false><<<grid, kLnThreads, ln_smem(C), s>>>(
        (const T*)dy, (const T*)o, (const T*)x, mean, rstd, gamma, (T*)dz, delta, dgamma, dbeta, C, N);
  return check_launch("residual_ln_bwd");
}

int launch_residual_ln_fwd(const void* o, const void* x, const float* gamma, const float* beta, float eps, void* y,
                           float* mean, float* rstd, int B, int C, int N, int dtype, int tok, cudaStream_t s) {
  if (dtype == MU_F32) return run_fwd<float>(o, x, gamma, beta, eps, y, mean, rstd, B, C, N, tok, s);
  return run_fwd<__nv_bfloat16>(o, x, gamma, beta, eps, y, mean, rstd, B, C, N, tok, s);
}
int launch_residual_ln_bwd(const void* dy, const void* o, const void* x, const float* mean, const float* rstd,
                           const float* gamma, void* dz, float* delta, float* dgamma, float* dbeta, int B, int C, int N,
                           int dtype, int tok, cudaStream_t s) {
  if (dtype == MU_F32) return run_bwd<float>(dy, o, x, mean, rstd, gamma, dz, delta, dgamma, dbeta, B, C, N, tok, s);
  return run_bwd<__nv_bfloat16>(dy, o, x, mean, rstd, gamma, dz, delta, dgamma, dbeta, B, C, N, tok, s);
}

}  // namespace mu
