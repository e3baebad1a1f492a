// K1 / K6: Q/K/V projections and their backward, CUDA-core tiled GEMMs (fp32 accumulate).
// Replaces nn.Linear x3 on the permuted token view, /root/reference/code/ade20k/ade_semantic.py:168-172,
// and autograd of the same (:400).  The token permute of :168 is never materialised: the forward reads
// x[b, :, n] straight from NCHW and the backward writes dx back transposed through shared memory.
// K and V rows are written compacted to the kept keys (row keep_rank[b, n]) so attention needs no mask.
#include "common.cuh"

namespace mu {

constexpr int kTile = 64;   // tile edge in M and N
constexpr int kStep = 16;   // K step
constexpr int kThreads = 256;

template <typename T> __device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <> __device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 v;
  v.x = *reinterpret_cast<uint32_t*>(&lo);
  v.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = v;
}

// ------------------------------------------------------------------ forward
// out[b, n, o] = sum_c x[b, c, n] * W[o, c] + bias[o]      grid (ceil(N/64), 3C/64, B)
template <typename T>
__global__ void __launch_bounds__(kThreads) qkv_project_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias,
                                                               const int32_t* __restrict__ rank, T* __restrict__ q,
                                                               T* __restrict__ kc, T* __restrict__ vc, int C, int N,
                                                               int NKP) {
  __shared__ __align__(16) float As[kStep][kTile];  // [k][token]
  __shared__ __align__(16) float Ws[kStep][kTile];  // [k][out]
  const int b = blockIdx.z, n0 = blockIdx.x * kTile, o0 = blockIdx.y * kTile;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const T* xb = x + (size_t)b * C * N;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < C; k0 += kStep) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * kThreads, kk = idx >> 6, mm = idx & 63;
      const int n = n0 + mm;
      As[kk][mm] = (n < N) ? ld_f(xb + (size_t)(k0 + kk) * N + n) : 0.f;
      Ws[kk][mm] = w[(size_t)(o0 + mm) * C + k0 + kk];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kStep; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int which = o0 / C, oc = o0 - which * C + tx * 4;
  const float4 bi = *reinterpret_cast<const float4*>(bias + o0 + tx * 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
    T* dst;
    if (which == 0) {
      dst = q + ((size_t)b * N + n) * C + oc;
    } else {
      const int r = rank[(size_t)b * N + n];
      if (r < 0) continue;
      dst = (which == 1 ? kc : vc) + ((size_t)b * NKP + r) * C + oc;
    }
    store4<T>(dst, acc[i][0] + bi.x, acc[i][1] + bi.y, acc[i][2] + bi.z, acc[i][3] + bi.w);
  }
}

// zero rows [n_keep, roundup(n_keep, 128)) of kc / vc so that padded tile rows are finite
template <typename T>
__global__ void zero_pad_rows_kernel(const int32_t* __restrict__ n_keep, T* __restrict__ kc, T* __restrict__ vc, int C,
                                     int NKP) {
  const int b = blockIdx.x, nk = n_keep[b];
  const int end = min(NKP, (nk + 127) / 128 * 128);
  const int count = (end - nk) * C;
  T* k = kc + ((size_t)b * NKP + nk) * C;
  T* v = vc + ((size_t)b * NKP + nk) * C;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    st_f(k + i, 0.f);
    st_f(v + i, 0.f);
  }
}

template <typename T>
static int run_fwd(const void* x, const float* w, const float* b, const int32_t* rank, const int32_t* n_keep, void* q,
                   void* kc, void* vc, int B, int C, int N, int NKP, cudaStream_t s) {
  dim3 grid((N + kTile - 1) / kTile, 3 * C / kTile, B);
  qkv_project_kernel<T><<<grid, kThreads, 0, s>>>((const T*)x, w, b, rank, (T*)q, (T*)kc, (T*)vc, C, N, NKP);
  zero_pad_rows_kernel<T><<<B, 256, 0, s>>>(n_keep, (T*)kc, (T*)vc, C, NKP);
  return check_launch("qkv_project");
}

int launch_zero_pad_rows(const int32_t* n_keep, void* kc, void* vc, int B, int C, int NKP, int dtype, cudaStream_t s) {
  if (dtype == MU_F32) zero_pad_rows_kernel<float><<<B, 256, 0, s>>>(n_keep, (float*)kc, (float*)vc, C, NKP);
  else zero_pad_rows_kernel<__nv_bfloat16><<<B, 256, 0, s>>>(n_keep, (__nv_bfloat16*)kc, (__nv_bfloat16*)vc, C, NKP);
  return check_launch("zero_pad_rows");
}

int launch_qkv_project(const void* x, const float* w, const float* b, const int32_t* rank, const int32_t* n_keep,
                       void* q, void* kc, void* vc, int B, int C, int N, int NKP, int dtype, cudaStream_t s) {
  if (dtype == MU_F32) return run_fwd<float>(x, w, b, rank, n_keep, q, kc, vc, B, C, N, NKP, s);
  return run_fwd<__nv_bfloat16>(x, w, b, rank, n_keep, q, kc, vc, B, C, N, NKP, s);
}

// ------------------------------------------------------------------ backward: dx
// A[t, o] for o in [0, 3C): dq | dk | dv, all in token space (dk / dv rows of masked keys are zero)
template <typename T>
__device__ __forceinline__ float load_dqkv(const T* dq, const T* dk, const T* dv, const int32_t* rank, int b, int n,
                                           int o, int C, int N, int NKP) {
  // dq, dk, dv are all dense [B, N, C]; rows of masked keys are zero in dk / dv
  const size_t row = ((size_t)b * N + n) * C;
  if (o < C) return ld_f(dq + row + o);
  if (o < 2 * C) return ld_f(dk + row + (o - C));
  return ld_f(dv + row + (o - 2 * C));
}

// dx[b, i, n] = dz[b, n, i] + sum_o A[(b,n), o] * W[o, i]       grid (ceil(N/64), C/64, B)
template <typename T>
__global__ void __launch_bounds__(kThreads) qkv_dx_kernel(const T* __restrict__ dz, const T* __restrict__ dq,
                                                          const T* __restrict__ dkc, const T* __restrict__ dvc,
                                                          const int32_t* __restrict__ rank, const float* __restrict__ w,
                                                          T* __restrict__ dx, int C, int N, int NKP) {
  __shared__ float As[kTile][kStep + 1];            // [token][k]
  __shared__ __align__(16) float Bs[kStep][kTile];  // [k][i]
  __shared__ float Cs[kTile][kTile + 1];            // [i][token] staging for the transposed store
  const int b = blockIdx.z, n0 = blockIdx.x * kTile, i0 = blockIdx.y * kTile;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < 3 * C; k0 += kStep) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * kThreads;
      const int mm = idx >> 4, kk = idx & 15, n = n0 + mm;
      As[mm][kk] = (n < N) ? load_dqkv(dq, dkc, dvc, rank, b, n, k0 + kk, C, N, NKP) : 0.f;
      const int kb = idx >> 6, ii = idx & 63;
      Bs[kb][ii] = w[(size_t)(k0 + kb) * C + i0 + ii];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kStep; ++kk) {
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = As[ty * 4 + i][kk];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a, bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ii = tx * 4 + j;
      float v = acc[i][j];
      if (n < N) v += ld_f(dz + ((size_t)b * N + n) * C + i0 + ii);
      Cs[ii][ty * 4 + i] = v;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < kTile * kTile; idx += kThreads) {
    const int ii = idx >> 6, mm = idx & 63, n = n0 + mm;
    if (n < N) st_f(dx + ((size_t)b * C + i0 + ii) * N + n, Cs[ii][mm]);
  }
}

// ------------------------------------------------------------------ backward: dW, db (split over tokens)
// dW[o, i] += sum_{t in chunk} A[t, o] * x[b(t), i, n(t)]       grid (3C/64, C/64, splits)
template <typename T>
__global__ void __launch_bounds__(kThreads) qkv_dw_kernel(const T* __restrict__ x, const T* __restrict__ dq,
                                                          const T* __restrict__ dkc, const T* __restrict__ dvc,
                                                          const int32_t* __restrict__ rank, float* __restrict__ dw,
                                                          float* __restrict__ db, int C, int N, int NKP, long total,
                                                          int chunk) {
  __shared__ __align__(16) float As[kStep][kTile];      // [t][o]
  __shared__ __align__(16) float Bs[kStep][kTile + 4];  // [t][i]
  const int o0 = blockIdx.x * kTile, i0 = blockIdx.y * kTile;
  const long t_begin = (long)blockIdx.z * chunk;
  const long t_end = t_begin + chunk < total ? t_begin + chunk : total;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  float bsum = 0.f;
  for (long t0 = t_begin; t0 < t_end; t0 += kStep) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * kThreads;
      {
        const int kk = idx >> 6, oo = idx & 63;
        const long t = t0 + kk;
        float v = 0.f;
        if (t < t_end) {
          const int b = (int)(t / N), n = (int)(t - (long)b * N);
          v = load_dqkv(dq, dkc, dvc, rank, b, n, o0 + oo, C, N, NKP);
        }
        As[kk][oo] = v;
      }
      {
        const int ii = idx >> 4, kk = idx & 15;
        const long t = t0 + kk;
        float v = 0.f;
        if (t < t_end) {
          const int b = (int)(t / N), n = (int)(t - (long)b * N);
          v = ld_f(x + ((size_t)b * C + i0 + ii) * N + n);
        }
        Bs[kk][ii] = v;
      }
    }
    __syncthreads();
    if (blockIdx.y == 0 && tid < kTile) {
#pragma unroll
      for (int kk = 0; kk < kStep; ++kk) bsum += As[kk][tid];
    }
#pragma unroll
    for (int kk = 0; kk < kStep; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(dw + (size_t)(o0 + ty * 4 + i) * C + i0 + tx * 4 + j, acc[i][j]);
  if (blockIdx.y == 0 && tid < kTile) atomicAdd(db + o0 + tid, bsum);
}

template <typename T>
static int run_bwd(const void* x, const void* dz, const void* dq, const void* dkc, const void* dvc,
                   const int32_t* rank, const float* w, void* dx, float* dw, float* db, int B, int C, int N, int NKP,
                   cudaStream_t s) {
  dim3 g1((N + kTile - 1) / kTile, C / kTile, B);
  qkv_dx_kernel<T><<<g1, kThreads, 0, s>>>((const T*)dz, (const T*)dq, (const T*)dkc, (const T*)dvc, rank, w, (T*)dx,
                                           C, N, NKP);
  const long total = (long)B * N;
  long chunk = (total + 255) / 256;            // aim for <= 256 splits
  if (chunk < 1024) chunk = 1024;
  chunk = (chunk + kStep - 1) / kStep * kStep;
  const int splits = (int)((total + chunk - 1) / chunk);
  dim3 g2(3 * C / kTile, C / kTile, splits);
  qkv_dw_kernel<T><<<g2, kThreads, 0, s>>>((const T*)x, (const T*)dq, (const T*)dkc, (const T*)dvc, rank, dw, db, C, N,
                                           NKP, total, (int)chunk);
  return check_launch("qkv_project_bwd");
}

int launch_qkv_project_bwd(const void* x, const void* dz, const void* dq, const void* dkc, const void* dvc,
                           const int32_t* rank, const float* w, void* dx, float* dw, float* db, int B, int C, int N,
                           int NKP, int dtype, cudaStream_t s) {
  if (dtype == MU_F32) return run_bwd<float>(x, dz, dq, dkc, dvc, rank, w, dx, dw, db, B, C, N, NKP, s);
  return run_bwd<__nv_bfloat16>(x, dz, dq, dkc, dvc, rank, w, dx, dw, db, B, C, N, NKP, s);
}

}  // namespace mu
