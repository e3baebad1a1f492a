// K3 / K5 in MU_F32 validation mode: flash-style masked attention on CUDA cores, fp32 throughout.
// Replaces /root/reference/code/ade20k/ade_semantic.py:174-186 (QK^T, /sqrt(C), + mask, softmax, PV)
// and its autograd.  The additive -inf bias is realised by only ever visiting the compacted kept keys.
// This path exists to meet the fp32 rel-err <= 1e-4 bar and to cross-check the tcgen05 kernels at sizes the
// CPU oracle cannot reach; the bf16 production path is attn_fwd_sm100.cu / attn_bwd_sm100.cu.
#include "common.cuh"

namespace mu {

constexpr int kQ = 32;         // query rows per CTA
constexpr int kThreadsA = 256; // 8 threads per query row
constexpr float kLog2e = 1.4426950408889634f;

// ------------------------------------------------------------------ forward
// grid (ceil(N/32), B).  smem: Qs[32][D+1] Ks[KT][D+1] Vs[KT][D] Ss[32][KT+1]
template <typename T, int D, int KT>
__global__ void __launch_bounds__(kThreadsA) attn_fwd_simt_kernel(const T* __restrict__ q, const T* __restrict__ kc,
                                                                  const T* __restrict__ vc,
                                                                  const int32_t* __restrict__ n_keep,
                                                                  T* __restrict__ o, float* __restrict__ lse, int N,
                                                                  int NKP, float scale_log2) {
  extern __shared__ float sm[];
  float* Qs = sm;
  float* Ks = Qs + kQ * (D + 1);
  float* Vs = Ks + KT * (D + 1);
  float* Ss = Vs + KT * D;
  const int b = blockIdx.y, q0 = blockIdx.x * kQ, tid = threadIdx.x;
  const int nk = n_keep[b];
  for (int idx = tid; idx < kQ * D; idx += kThreadsA) {
    const int r = idx / D, d = idx - r * D;
    Qs[r * (D + 1) + d] = (q0 + r < N) ? ld_f(q + ((size_t)b * N + q0 + r) * D + d) : 0.f;
  }
  const int row = tid >> 3, sub = tid & 7;   // softmax / PV ownership: 8 threads per query row
  const int sq = tid >> 4, sk = tid & 15;    // S ownership: queries {sq, sq+16}, keys {sk + 16*j}
  constexpr int KJ = KT / 16;
  float acc[D / 8];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) acc[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < nk; k0 += KT) {
    __syncthreads();
    for (int idx = tid; idx < KT * D; idx += kThreadsA) {
      const int r = idx / D, d = idx - r * D;
      const bool ok = k0 + r < nk;
      const size_t off = ((size_t)b * NKP + k0 + r) * D + d;
      Ks[r * (D + 1) + d] = ok ? ld_f(kc + off) : 0.f;
      Vs[r * D + d] = ok ? ld_f(vc + off) : 0.f;
    }
    __syncthreads();
    float s[2][KJ];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < KJ; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      const float a0 = Qs[sq * (D + 1) + d], a1 = Qs[(sq + 16) * (D + 1) + d];
#pragma unroll
      for (int j = 0; j < KJ; ++j) {
        const float kv = Ks[(sk + 16 * j) * (D + 1) + d];
        s[0][j] = fmaf(a0, kv, s[0][j]);
        s[1][j] = fmaf(a1, kv, s[1][j]);
      }
    }
#pragma unroll
    for (int j = 0; j < KJ; ++j) {
      const bool ok = k0 + sk + 16 * j < nk;
      Ss[sq * (KT + 1) + sk + 16 * j] = ok ? s[0][j] : -INFINITY;
      Ss[(sq + 16) * (KT + 1) + sk + 16 * j] = ok ? s[1][j] : -INFINITY;
    }
    __syncthreads();
    // online softmax for `row`, 8 cooperating threads (consecutive lanes)
    float mx = -INFINITY;
    for (int j = sub; j < KT; j += 8) mx = fmaxf(mx, Ss[row * (KT + 1) + j]);
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    const float m_new = fmaxf(m, mx);
    const float alpha = exp2f((m - m_new) * scale_log2);  // m = -inf on the first tile -> 0
    float psum = 0.f;
    for (int j = sub; j < KT; j += 8) {
      const float p = exp2f((Ss[row * (KT + 1) + j] - m_new) * scale_log2);
      Ss[row * (KT + 1) + j] = p;
      psum += p;
    }
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, off);
    l = l * alpha + psum;
    m = m_new;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) acc[i] *= alpha;
    __syncwarp();
    for (int j = 0; j < KT; ++j) {
      const float p = Ss[row * (KT + 1) + j];
#pragma unroll
      for (int i = 0; i < D / 8; ++i) acc[i] = fmaf(p, Vs[j * D + sub + 8 * i], acc[i]);
    }
  }
  if (q0 + row < N) {
    const float inv = 1.f / l;
    T* orow = o + ((size_t)b * N + q0 + row) * D;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) st_f(orow + sub + 8 * i, acc[i] * inv);
    if (sub == 0) lse[(size_t)b * N + q0 + row] = (m * scale_log2 + log2f(l)) / kLog2e;
  }
}

// ------------------------------------------------------------------ backward, dQ (query-stationary)
// grid (ceil(N/32), B).  smem: Qs[32][D+1] dOs[32][D+1] Ks[32][D+1] Vs[32][D+1] dSs[32][33]
template <typename T, int D>
__global__ void __launch_bounds__(kThreadsA) attn_bwd_dq_simt_kernel(
    const T* __restrict__ q, const T* __restrict__ kc, const T* __restrict__ vc, const int32_t* __restrict__ n_keep,
    const T* __restrict__ d_o, const float* __restrict__ lse, const float* __restrict__ delta, T* __restrict__ dq,
    int N, int NKP, float scale) {
  constexpr int KT = 32, LD = D + 1;
  extern __shared__ float sm[];
  float* Qs = sm;
  float* dOs = Qs + kQ * LD;
  float* Ks = dOs + kQ * LD;
  float* Vs = Ks + KT * LD;
  float* dSs = Vs + KT * LD;
  const int b = blockIdx.y, q0 = blockIdx.x * kQ, tid = threadIdx.x;
  const int nk = n_keep[b];
  for (int idx = tid; idx < kQ * D; idx += kThreadsA) {
    const int r = idx / D, d = idx - r * D;
    const bool ok = q0 + r < N;
    const size_t off = ((size_t)b * N + q0 + r) * D + d;
    Qs[r * LD + d] = ok ? ld_f(q + off) : 0.f;
    dOs[r * LD + d] = ok ? ld_f(d_o + off) : 0.f;
  }
  const int row = tid >> 3, sub = tid & 7;
  const int sq = tid >> 3, sk = tid & 7;  // S ownership: query sq, keys sk + 8*j (j < 4)
  const bool qok = q0 + sq < N;
  const float my_lse = qok ? lse[(size_t)b * N + q0 + sq] : 0.f;
  const float my_delta = qok ? delta[(size_t)b * N + q0 + sq] : 0.f;
  float acc[D / 8];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) acc[i] = 0.f;
  for (int k0 = 0; k0 < nk; k0 += KT) {
    __syncthreads();
    for (int idx = tid; idx < KT * D; idx += kThreadsA) {
      const int r = idx / D, d = idx - r * D;
      const bool ok = k0 + r < nk;
      const size_t off = ((size_t)b * NKP + k0 + r) * D + d;
      Ks[r * LD + d] = ok ? ld_f(kc + off) : 0.f;
      Vs[r * LD + d] = ok ? ld_f(vc + off) : 0.f;
    }
    __syncthreads();
    float s[4] = {}, dp[4] = {};
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      const float qv = Qs[sq * LD + d], dov = dOs[sq * LD + d];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] = fmaf(qv, Ks[(sk + 8 * j) * LD + d], s[j]);
        dp[j] = fmaf(dov, Vs[(sk + 8 * j) * LD + d], dp[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = qok && (k0 + sk + 8 * j < nk);
      const float p = ok ? __expf(s[j] * scale - my_lse) : 0.f;
      dSs[sq * 33 + sk + 8 * j] = p * (dp[j] - my_delta) * scale;
    }
    __syncthreads();
    for (int j = 0; j < KT; ++j) {
      const float ds = dSs[row * 33 + j];
#pragma unroll
      for (int i = 0; i < D / 8; ++i) acc[i] = fmaf(ds, Ks[j * LD + sub + 8 * i], acc[i]);
    }
  }
  if (q0 + row < N) {
    T* drow = dq + ((size_t)b * N + q0 + row) * D;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) st_f(drow + sub + 8 * i, acc[i]);
  }
}

// ------------------------------------------------------------------ backward, dK / dV (key-stationary)
// grid (ceil(NKP/32), B); CTAs beyond n_keep exit.  smem: Ks Vs Qs dOs [32][D+1], Ps dSs [32][33] (stored [q][k])
template <typename T, int D>
__global__ void __launch_bounds__(kThreadsA) attn_bwd_dkv_simt_kernel(
    const T* __restrict__ q, const T* __restrict__ kc, const T* __restrict__ vc, const int32_t* __restrict__ n_keep,
    const T* __restrict__ d_o, const float* __restrict__ lse, const float* __restrict__ delta,
    const int32_t* __restrict__ keep_idx, T* __restrict__ dk, T* __restrict__ dv, int N, int NKP, float scale) {
  constexpr int KT = 32, LD = D + 1;
  extern __shared__ float sm[];
  float* Ks = sm;
  float* Vs = Ks + KT * LD;
  float* Qs = Vs + KT * LD;
  float* dOs = Qs + kQ * LD;
  float* Ps = dOs + kQ * LD;
  float* dSs = Ps + kQ * 33;
  const int b = blockIdx.y, k0 = blockIdx.x * KT, tid = threadIdx.x;
  const int nk = n_keep[b];
  if (k0 >= nk) return;
  for (int idx = tid; idx < KT * D; idx += kThreadsA) {
    const int r = idx / D, d = idx - r * D;
    const bool ok = k0 + r < nk;
    const size_t off = ((size_t)b * NKP + k0 + r) * D + d;
    Ks[r * LD + d] = ok ? ld_f(kc + off) : 0.f;
    Vs[r * LD + d] = ok ? ld_f(vc + off) : 0.f;
  }
  const int row = tid >> 3, sub = tid & 7;  // accumulator ownership: key `row`, channels sub + 8*i
  const int sq = tid >> 3, sk = tid & 7;    // S ownership: query sq, keys sk + 8*j
  float accK[D / 8], accV[D / 8];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) accK[i] = accV[i] = 0.f;
  for (int q0 = 0; q0 < N; q0 += kQ) {
    __syncthreads();
    for (int idx = tid; idx < kQ * D; idx += kThreadsA) {
      const int r = idx / D, d = idx - r * D;
      const bool ok = q0 + r < N;
      const size_t off = ((size_t)b * N + q0 + r) * D + d;
      Qs[r * LD + d] = ok ? ld_f(q + off) : 0.f;
      dOs[r * LD + d] = ok ? ld_f(d_o + off) : 0.f;
    }
    __syncthreads();
    const bool qok = q0 + sq < N;
    const float my_lse = qok ? lse[(size_t)b * N + q0 + sq] : 0.f;
    const float my_delta = qok ? delta[(size_t)b * N + q0 + sq] : 0.f;
    float s[4] = {}, dp[4] = {};
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      const float qv = Qs[sq * LD + d], dov = dOs[sq * LD + d];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] = fmaf(qv, Ks[(sk + 8 * j) * LD + d], s[j]);
        dp[j] = fmaf(dov, Vs[(sk + 8 * j) * LD + d], dp[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = qok && (k0 + sk + 8 * j < nk);
      const float p = ok ? __expf(s[j] * scale - my_lse) : 0.f;
      Ps[sq * 33 + sk + 8 * j] = p;
      dSs[sq * 33 + sk + 8 * j] = p * (dp[j] - my_delta) * scale;
    }
    __syncthreads();
    for (int qq = 0; qq < kQ; ++qq) {
      const float p = Ps[qq * 33 + row], ds = dSs[qq * 33 + row];
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        accV[i] = fmaf(p, dOs[qq * LD + sub + 8 * i], accV[i]);
        accK[i] = fmaf(ds, Qs[qq * LD + sub + 8 * i], accK[i]);
      }
    }
  }
  if (k0 + row < nk) {  // scatter the key row back to its token position (dense [B, N, D] outputs)
    const size_t off = ((size_t)b * N + keep_idx[(size_t)b * N + k0 + row]) * D;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      st_f(dk + off + sub + 8 * i, accK[i]);
      st_f(dv + off + sub + 8 * i, accV[i]);
    }
  }
}

// ------------------------------------------------------------------ launchers
template <typename T, int D>
static int run_fwd(const void* q, const void* kc, const void* vc, const int32_t* n_keep, void* o, float* lse, int B,
                   int N, int NKP, cudaStream_t s) {
  constexpr int KT = (D <= 128) ? 64 : 32;
  const size_t smem = sizeof(float) * (kQ * (D + 1) + KT * (D + 1) + KT * D + kQ * (KT + 1));
  auto kern = attn_fwd_simt_kernel<T, D, KT>;
  set_max_dynamic_smem_once(kern, (int)smem);
  dim3 grid((N + kQ - 1) / kQ, B);
  const float scale_log2 = kLog2e / sqrtf((float)D);
  kern<<<grid, kThreadsA, smem, s>>>((const T*)q, (const T*)kc, (const T*)vc, n_keep, (T*)o, lse, N, NKP, scale_log2);
  return check_launch("attn_fwd_simt");
}

template <typename T, int D>
static int run_bwd(const void* q, const void* kc, const void* vc, const int32_t* n_keep, const int32_t* keep_idx,
                   const void* d_o, const float* lse, const float* delta, void* dq, void* dkc, void* dvc, int B, int N,
                   int NKP, cudaStream_t s) {
  cudaMemsetAsync(dkc, 0, (size_t)B * N * D * sizeof(T), s);
  cudaMemsetAsync(dvc, 0, (size_t)B * N * D * sizeof(T), s);
  const float scale = 1.f / sqrtf((float)D);
  {
    const size_t smem = sizeof(float) * (4 * 32 * (D + 1) + 32 * 33);
    auto kern = attn_bwd_dq_simt_kernel<T, D>;
    set_max_dynamic_smem_once(kern, (int)smem);
    dim3 grid((N + kQ - 1) / kQ, B);
    kern<<<grid, kThreadsA, smem, s>>>((const T*)q, (const T*)kc, (const T*)vc, n_keep, (const T*)d_o, lse, delta,
                                       (T*)dq, N, NKP, scale);
  }
  {
    const size_t smem = sizeof(float) * (4 * 32 * (D + 1) + 2 * 32 * 33);
    auto kern = attn_bwd_dkv_simt_kernel<T, D>;
    set_max_dynamic_smem_once(kern, (int)smem);
    dim3 grid((N + 31) / 32, B);
    kern<<<grid, kThreadsA, smem, s>>>((const T*)q, (const T*)kc, (const T*)vc, n_keep, (const T*)d_o, lse, delta,
                                       keep_idx, (T*)dkc, (T*)dvc, N, NKP, scale);
  }
  return check_launch("attn_bwd_simt");
}

#define MU_DISPATCH_D(C, CALL)                                                \
  switch (C) {                                                                \
    case 64: { constexpr int D = 64; return CALL; }                           \
    case 128: { constexpr int D = 128; return CALL; }                         \
    case 256: { constexpr int D = 256; return CALL; }                         \
    default: set_error("channels must be 64, 128 or 256 (got %d)", C); return MU_ERR_BAD_SHAPE; \
  }

int launch_attn_fwd_simt(const void* q, const void* kc, const void* vc, const int32_t* n_keep, void* o, float* lse,
                         int B, int N, int NKP, int C, int dtype, cudaStream_t s) {
  if (dtype == MU_F32) { MU_DISPATCH_D(C, (run_fwd<float, D>(q, kc, vc, n_keep, o, lse, B, N, NKP, s))) }
  MU_DISPATCH_D(C, (run_fwd<__nv_bfloat16, D>(q, kc, vc, n_keep, o, lse, B, N, NKP, s)))
}

int launch_attn_bwd_simt(const void* q, const void* kc, const void* vc, const int32_t* n_keep,
                         const int32_t* keep_idx, const void* d_o, const float* lse, const float* delta, void* dq,
                         void* dk, void* dv, int B, int N, int NKP, int C, int dtype, cudaStream_t s) {
  if (dtype == MU_F32) {
    MU_DISPATCH_D(C, (run_bwd<float, D>(q, kc, vc, n_keep, keep_idx, d_o, lse, delta, dq, dk, dv, B, N, NKP, s)))
  }
  MU_DISPATCH_D(C, (run_bwd<__nv_bfloat16, D>(q, kc, vc, n_keep, keep_idx, d_o, lse, delta, dq, dk, dv, B, N, NKP, s)))
}

}  // namespace mu
