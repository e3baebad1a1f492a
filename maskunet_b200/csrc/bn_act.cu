// K8: training-mode BatchNorm2d fused with its activation and residual, channels-last activations.
// Replaces the nn.BatchNorm2d -> nn.GELU / F.gelu(x + .) / nn.ReLU chains of the reference's DoubleConv blocks,
// /root/reference/code/ade20k/ade_semantic.py:198-210 (ConvBlock), :219, :240 (trailing BN of Down/UpSample) and
// :283-287 (head BN + ReLU), and their autograd.  Activations are [M = B*H*W rows][C channels] (NHWC memory).
//
//   forward   stats   : per-channel sum / sum of squares              (1 read of x)
//             finalize: mean, rstd, running-stat update, a = gamma*rstd, b = beta - mean*a     ([C] work)
//             apply   : y = act(a*x + b [+ r])                         (1 read of x [+ r], 1 write)
//   backward  reduce  : dz = dy*act'(z);  S1 = sum dz, S2 = sum dz*xhat (reads dy, x [, r])
//             apply   : dx = gamma*rstd*(dz - S1/M - xhat*S2/M), dr = dz   (reads dy, x [, r]; writes dx [, dr])
// All HBM-bound: thread <-> fixed channel group (16-byte vectors when C % 8 == 0), rows strided over the grid,
// per-thread fp32 partial sums, shared-memory reduction across row lanes, one atomicAdd per channel per CTA.
#include "common.cuh"
#include "det_reduce.cuh"

namespace mu {

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };
constexpr int kBnThreads = 256;
// 8 CTAs of 256 threads = full occupancy (32 registers per thread).  The streaming kernels track occupancy: the forward
// GELU apply at 40 registers (6 CTAs / SM) ran at 0.76 of the bandwidth of the 32-register variants, and capping it
// gives +19 % (4.27 -> 5.07 TB/s); the backward kernels need their registers (capped, they fell from 3.6 to 2.5 TB/s).
constexpr int kBnMinBlocks = 8;

// Exact-erf GELU (nn.GELU() default) with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. fp32
// round-off level): 2 MUFU (ex2, rcp) + ~12 FMA instead of libdevice erff's branchy ~25 instructions, which made
// the fused kernels instruction-bound rather than HBM-bound.  E = exp(-z^2/2) is shared with the derivative.
__device__ __forceinline__ void gelu_parts(float z, float& cdf, float& e_term) {
  const float x = fabsf(z) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, x, 1.f));
  e_term = __expf(-x * x);
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e_term, 1.f);
  cdf = 0.5f * (1.f + copysignf(erf_abs, z));
}
__device__ __forceinline__ float gelu_f(float z) {
  float cdf, e;
  gelu_parts(z, cdf, e);
  return z * cdf;
}
__device__ __forceinline__ float gelu_grad_f(float z) {
  float cdf, e;
  gelu_parts(z, cdf, e);
  return fmaf(z * 0.3989422804014327f, e, cdf);
}
// bf16 activations: Phi(z) = sigmoid(z (a + b z^2 + c z^4)) fitted to the normal CDF on [-8, 8] (clamped outside):
// |Phi error| <= 1.9e-5, |gelu error| <= 5.5e-5, |gelu' error| <= 1.4e-4 -- 1/70 .. 1/30 of the bf16 rounding of the
// stored result -- at 10 (forward) / 15 (gradient) instructions instead of ~20 / ~24 with the erf form above, which
// kept these HBM kernels at the instruction limit (3.7 TB/s against 5.6 TB/s without GELU).  fp32 activations (the
// validation configuration, tolerance 1e-4) keep the erf form.
constexpr float kGa = 1.59543567f, kGb = 7.35965107e-02f, kGc = -6.31613752e-04f, kNegLog2e = -1.4426950408889634f;
// MU_GELU_TANH = 1: sigmoid(t) = 0.5 + 0.5 tanh(t / 2) -- ONE MUFU op (tanh.approx) instead of two (ex2 + rcp).  The
// GELU kernels run two (forward) / four (backward: reduce and apply pass) MUFU ops per element and sat at 68 % MUFU
// utilisation beside the HBM stream; tanh.approx.f32 has a relative error of 2^-11, i.e. |Phi error| <= 2.5e-4 -- an
// eighth of the bf16 rounding of the stored activation at |z| ~ 1.
#ifndef MU_GELU_TANH
#define MU_GELU_TANH 1
#endif
__device__ __forceinline__ float fast_cdf(float z, float& z2) {
  const float zc = fminf(fmaxf(z, -8.f), 8.f);
  z2 = zc * zc;
#if MU_GELU_TANH
  const float u = zc * fmaf(z2, fmaf(z2, kGc * 0.5f, kGb * 0.5f), kGa * 0.5f);                  // t(z) / 2
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
  return fmaf(0.5f, th, 0.5f);
#else
  const float t = zc * fmaf(z2, fmaf(z2, kGc * kNegLog2e, kGb * kNegLog2e), kGa * kNegLog2e);   // -log2(e) t(z)
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
#endif
}
__device__ __forceinline__ float gelu_fast(float z) {
  float z2;
  return z * fast_cdf(z, z2);
}
__device__ __forceinline__ float gelu_grad_fast(float z) {
  float z2;
  const float cdf = fast_cdf(z, z2);
  const float tp = fmaf(z2, fmaf(z2, 5.f * kGc, 3.f * kGb), kGa);      // t'(z); cdf (1 - cdf) = 0 in the clamped range
  return fmaf(z * cdf * (1.f - cdf), tp, cdf);
}
template <typename T> struct FastAct { static constexpr bool value = false; };
template <> struct FastAct<__nv_bfloat16> { static constexpr bool value = true; };

template <int ACT, bool FAST = false> __device__ __forceinline__ float act_f(float z) {
  if (ACT == ACT_GELU) return FAST ? gelu_fast(z) : gelu_f(z);
  if (ACT == ACT_RELU) return fmaxf(z, 0.f);
  return z;
}
template <int ACT, bool FAST = false> __device__ __forceinline__ float act_grad_f(float z) {
  if (ACT == ACT_GELU) return FAST ? gelu_grad_fast(z) : gelu_grad_f(z);
  if (ACT == ACT_RELU) return z > 0.f ? 1.f : 0.f;
  return 1.f;
}

// ---- VEC-wide loads / stores with conversion -------------------------------------------------------
template <typename T, int VEC> struct Vec;
template <> struct Vec<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] = __uint_as_float(w[e] << 16);
      v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(v[0], v[1]); u.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[2], v[3]); u.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[4], v[5]); u.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[6], v[7]); u.w = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct Vec<float, 8> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <typename T> struct Vec<T, 2> {
  static __device__ __forceinline__ void load(const T* p, float (&v)[2]) { v[0] = ld_f(p); v[1] = ld_f(p + 1); }
  static __device__ __forceinline__ void store(T* p, const float (&v)[2]) { st_f(p, v[0]); st_f(p + 1, v[1]); }
};
template <typename T> struct Vec<T, 1> {
  static __device__ __forceinline__ void load(const T* p, float (&v)[1]) { v[0] = ld_f(p); }
  static __device__ __forceinline__ void store(T* p, const float (&v)[1]) { st_f(p, v[0]); }
};

// thread -> (channel group g, row lane rl); rows handled: rl + RL*(blockIdx.x + k*gridDim.x)
struct RowMap {
  int g, rl, G, RL;
  bool active;
  __device__ RowMap(int C, int VEC) {
    G = C / VEC;
    RL = kBnThreads / G;
    g = threadIdx.x % G;
    rl = threadIdx.x / G;
    active = rl < RL;
  }
};

// block reduction of per-thread [VEC] partials over row lanes, then one atomicAdd per channel
// (deterministic mode: `part` = this CTA's [C] slice of the scratch; the block totals are stored there and det_finish
// adds the slices in CTA order)
template <int VEC>
__device__ __forceinline__ void block_reduce_add(float* red /*[kBnThreads*VEC]*/, const float (&a)[VEC], const RowMap& m,
                                                 float* out, int C, int Cper, float* part = nullptr) {
  __syncthreads();
#pragma unroll
  for (int e = 0; e < VEC; ++e) red[threadIdx.x * VEC + e] = m.active ? a[e] : 0.f;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    const int g = c / VEC, e = c - g * VEC;
    float s = 0.f;
    for (int rl = 0; rl < m.RL; ++rl) s += red[(rl * m.G + g) * VEC + e];
    if (part != nullptr) part[c] = s;
    else atomicAdd(out + (c % Cper), s);      // folded rows: column c of the wide row is channel c mod Cper
  }
}

// ------------------------------------------------------------------ forward: statistics
template <typename T, int VEC>
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const T* __restrict__ x, float* __restrict__ sums, long M,
                                                              int C, int Cper, const DetCtx det) {
  __shared__ float red[kBnThreads * VEC];
  const RowMap m(C, VEC);
  float s1[VEC], s2[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) s1[e] = s2[e] = 0.f;
  if (m.active) {
    for (long r = (long)blockIdx.x * m.RL + m.rl; r < M; r += (long)gridDim.x * m.RL) {
      float v[VEC];
      Vec<T, VEC>::load(x + r * C + m.g * VEC, v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        s1[e] += v[e];
        s2[e] = fmaf(v[e], v[e], s2[e]);
      }
    }
  }
  float* part = det.on() ? det.partial + (size_t)blockIdx.x * 2 * C : nullptr;
  block_reduce_add<VEC>(red, s1, m, sums, C, Cper, part);
  block_reduce_add<VEC>(red, s2, m, sums + Cper, C, Cper, det.on() ? part + C : nullptr);
  if (det.on()) det_finish(det, gridDim.x, gridDim.x, 2, C, Cper, sums, sums + Cper, threadIdx.x, kBnThreads, SyncThreads());
}

// ------------------------------------------------------------------ forward: finalize ([C] work, one CTA)
__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long M, int C, float momentum, float eps,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ a,
                                   float* __restrict__ b) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const double mu_ = (double)sums[c] / (double)M;
    double var = (double)sums[C + c] / (double)M - mu_ * mu_;
    if (var < 0.0) var = 0.0;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    mean[c] = (float)mu_;
    rstd[c] = rs;
    const float ac = gamma[c] * rs;
    a[c] = ac;
    b[c] = beta[c] - (float)mu_ * ac;
    if (running_mean != nullptr) {   // nn.BatchNorm2d: running = (1 - m) * running + m * batch (unbiased variance)
      const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu_;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

// ------------------------------------------------------------------ forward: apply
template <typename T, int VEC, int ACT, bool RES>
__global__ void __launch_bounds__(kBnThreads, (ACT == ACT_GELU && VEC == 8 && !RES) ? kBnMinBlocks : 1) bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ r,
                                                              const float* __restrict__ a, const float* __restrict__ b,
                                                              T* __restrict__ y, long M, int C, int Cper) {
  const RowMap m(C, VEC);
  if (!m.active) return;
  float av[VEC], bv[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    av[e] = a[(m.g * VEC + e) % Cper];
    bv[e] = b[(m.g * VEC + e) % Cper];
  }
  for (long row = (long)blockIdx.x * m.RL + m.rl; row < M; row += (long)gridDim.x * m.RL) {
    const long off = row * C + m.g * VEC;
    float v[VEC], rv[VEC];
    Vec<T, VEC>::load(x + off, v);
    if (RES) Vec<T, VEC>::load(r + off, rv);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float z = fmaf(av[e], v[e], bv[e]);
      if (RES) z += rv[e];
      v[e] = act_f<ACT, FastAct<T>::value>(z);
    }
    Vec<T, VEC>::store(y + off, v);
  }
}

// ------------------------------------------------------------------ backward: reduce
template <typename T, int VEC, int ACT, bool RES>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(
    const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ r, const float* __restrict__ a,
    const float* __restrict__ b, const float* __restrict__ mean, const float* __restrict__ rstd,
    float* __restrict__ sums, long M, int C, int Cper, const DetCtx det) {
  __shared__ float red[kBnThreads * VEC];
  const RowMap m(C, VEC);
  float s1[VEC], s2[VEC], av[VEC], bv[VEC], mv[VEC], rsv[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    s1[e] = s2[e] = 0.f;
    const int c = m.active ? (m.g * VEC + e) % Cper : 0;
    av[e] = a[c]; bv[e] = b[c]; mv[e] = mean[c]; rsv[e] = rstd[c];
  }
  if (m.active) {
    for (long row = (long)blockIdx.x * m.RL + m.rl; row < M; row += (long)gridDim.x * m.RL) {
      const long off = row * C + m.g * VEC;
      float g[VEC], v[VEC], rv[VEC];
      Vec<T, VEC>::load(dy + off, g);
      Vec<T, VEC>::load(x + off, v);
      if (RES) Vec<T, VEC>::load(r + off, rv);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float dz = g[e];
        if (ACT != ACT_NONE) {
          float z = fmaf(av[e], v[e], bv[e]);
          if (RES) z += rv[e];
          dz *= act_grad_f<ACT, FastAct<T>::value>(z);
        }
        s1[e] += dz;
        s2[e] = fmaf(dz, (v[e] - mv[e]) * rsv[e], s2[e]);
      }
    }
  }
  float* part = det.on() ? det.partial + (size_t)blockIdx.x * 2 * C : nullptr;
  block_reduce_add<VEC>(red, s1, m, sums, C, Cper, part);
  block_reduce_add<VEC>(red, s2, m, sums + Cper, C, Cper, det.on() ? part + C : nullptr);
  if (det.on()) det_finish(det, gridDim.x, gridDim.x, 2, C, Cper, sums, sums + Cper, threadIdx.x, kBnThreads, SyncThreads());
}

// ------------------------------------------------------------------ backward: apply
template <typename T, int VEC, int ACT, bool RES>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(
    const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ r, const float* __restrict__ a,
    const float* __restrict__ b, const float* __restrict__ mean, const float* __restrict__ rstd,
    const float* __restrict__ sums, T* __restrict__ dx, T* __restrict__ dr, long M, int C, int Cper) {
  const RowMap m(C, VEC);
  if (!m.active) return;
  float av[VEC], bv[VEC], mv[VEC], rsv[VEC], c1[VEC], c2[VEC];
  const float invM = 1.f / ((float)M * (float)(C / Cper));      // true row count = M wide rows x fold
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    const int c = (m.g * VEC + e) % Cper;
    av[e] = a[c]; bv[e] = b[c]; mv[e] = mean[c]; rsv[e] = rstd[c];
    c1[e] = sums[c] * invM;
    c2[e] = sums[Cper + c] * invM;
  }
  for (long row = (long)blockIdx.x * m.RL + m.rl; row < M; row += (long)gridDim.x * m.RL) {
    const long off = row * C + m.g * VEC;
    float g[VEC], v[VEC], rv[VEC], o[VEC];
    Vec<T, VEC>::load(dy + off, g);
    Vec<T, VEC>::load(x + off, v);
    if (RES) Vec<T, VEC>::load(r + off, rv);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float dz = g[e];
      if (ACT != ACT_NONE) {
        float z = fmaf(av[e], v[e], bv[e]);
        if (RES) z += rv[e];
        dz *= act_grad_f<ACT, FastAct<T>::value>(z);
      }
      g[e] = dz;
      const float xhat = (v[e] - mv[e]) * rsv[e];
      o[e] = av[e] * (dz - c1[e] - xhat * c2[e]);     // a = gamma * rstd
    }
    Vec<T, VEC>::store(dx + off, o);
    if (RES) Vec<T, VEC>::store(dr + off, g);
  }
}

// ------------------------------------------------------------------ backward, bulk-staged (bf16, 8-channel vectors)
// The register-resident backward kernels above sit at 40-48 registers, i.e. 5-6 CTAs per SM with one 16-byte load per
// stream in flight per thread: ~3.6 TB/s.  Here the input rows stream through a shared-memory ring filled by
// cp.async.bulk (the TMA unit, 1-D: whole rows are contiguous), so the bytes in flight are the ring (3 stages x 2-3
// streams x 16 KB per SM) and no longer depend on registers or occupancy; 512 threads read their 16-byte vectors from
// shared memory (conflict-free: consecutive threads, consecutive vectors) and write dx / dr straight to HBM.
constexpr int kBulkThreads = 512;
constexpr int kBulkStages = 3;
constexpr int kBulkChunkBytes = 16384;      // per stream per stage

__device__ __forceinline__ uint32_t bn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(bn_smem_u32(dst)), "l"(src), "r"(bytes), "r"(bn_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bn_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bn_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
                 : "=r"(ok) : "r"(bn_smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void unpack8_bf16(const uint4 u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = __uint_as_float(w[e] << 16);
    v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}

// MODE 0: reduce (sums[c] += sum dz, sums[Cper + c] += sum dz * xhat);  MODE 1: apply (dx, dr)
template <int ACT, bool RES, int MODE>
__global__ void __launch_bounds__(kBulkThreads, 1)
bn_bwd_bulk_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                   const __nv_bfloat16* __restrict__ r, const float* __restrict__ a, const float* __restrict__ b,
                   const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ sums,
                   __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dr, long M, int C, int Cper,
                   int rows_per_chunk, const DetCtx det) {
  constexpr int NS = RES ? 3 : 2;                       // streams: dy, x [, r]
  extern __shared__ __align__(128) uint8_t bulk_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(bulk_smem);                   // [kBulkStages]
  float* red = reinterpret_cast<float*>(bulk_smem + 64);                    // [2][C] block partial sums (MODE 0)
  uint8_t* ring = bulk_smem + 64 + ((2 * C * 4 + 127) / 128) * 128;
  const int G = C / 8, RL = kBulkThreads / G;
  const int g = threadIdx.x % G, rl = threadIdx.x / G;
  const bool active = rl < RL;
  const long n_chunks = (M + rows_per_chunk - 1) / rows_per_chunk;
  const uint32_t row_bytes = (uint32_t)C * 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBulkStages; ++i) bar_init(full + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (MODE == 0) for (int i = threadIdx.x; i < 2 * C; i += kBulkThreads) red[i] = 0.f;
  __syncthreads();

  auto issue = [&](long chunk, int stage) {            // thread 0 only
    const long r0 = chunk * rows_per_chunk;
    const long nrows = min((long)rows_per_chunk, M - r0);
    const uint32_t bytes = (uint32_t)nrows * row_bytes;
    uint8_t* dst = ring + (size_t)stage * NS * kBulkChunkBytes;
    bar_expect(full + stage, bytes * NS);
    bulk_load(dst, dy + r0 * C, bytes, full + stage);
    bulk_load(dst + kBulkChunkBytes, x + r0 * C, bytes, full + stage);
    if (RES) bulk_load(dst + 2 * kBulkChunkBytes, r + r0 * C, bytes, full + stage);
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < kBulkStages; ++i) {
      const long chunk = blockIdx.x + (long)i * gridDim.x;
      if (chunk < n_chunks) issue(chunk, i);
    }

  float av[8], bv[8], mv[8], rsv[8], c1[8], c2[8], s1[8], s2[8];
  const float invM = 1.f / ((float)M * (float)(C / Cper));
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = active ? (g * 8 + e) % Cper : 0;
    av[e] = a[c]; bv[e] = b[c]; mv[e] = mean[c]; rsv[e] = rstd[c];
    s1[e] = s2[e] = 0.f;
    if (MODE == 1) { c1[e] = sums[c] * invM; c2[e] = sums[Cper + c] * invM; }
  }

  int it = 0;
  for (long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++it) {
    const int stage = it % kBulkStages;
    bar_wait(full + stage, (it / kBulkStages) & 1);
    const long r0 = chunk * rows_per_chunk;
    const int nrows = (int)min((long)rows_per_chunk, M - r0);
    const uint8_t* src = ring + (size_t)stage * NS * kBulkChunkBytes;
    if (active) {
      for (int row = rl; row < nrows; row += RL) {
        const size_t off = (size_t)row * row_bytes + g * 16;
        float gy[8], v[8], rv[8];
        unpack8_bf16(*reinterpret_cast<const uint4*>(src + off), gy);
        unpack8_bf16(*reinterpret_cast<const uint4*>(src + kBulkChunkBytes + off), v);
        if (RES) unpack8_bf16(*reinterpret_cast<const uint4*>(src + 2 * kBulkChunkBytes + off), rv);
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float dz = gy[e];
          if (ACT != ACT_NONE) {
            float z = fmaf(av[e], v[e], bv[e]);
            if (RES) z += rv[e];
            dz *= act_grad_f<ACT, true>(z);
          }
          const float xhat = (v[e] - mv[e]) * rsv[e];
          if (MODE == 0) {
            s1[e] += dz;
            s2[e] = fmaf(dz, xhat, s2[e]);
          } else {
            gy[e] = dz;
            o[e] = av[e] * (dz - c1[e] - xhat * c2[e]);
          }
        }
        if (MODE == 1) {
          const size_t goff = (size_t)(r0 + row) * C + g * 8;
          Vec<__nv_bfloat16, 8>::store(dx + goff, o);
          if (RES) Vec<__nv_bfloat16, 8>::store(dr + goff, gy);
        }
      }
    }
    __syncthreads();                                   // everyone is done reading this stage
    const long next = chunk + (long)kBulkStages * gridDim.x;
    if (threadIdx.x == 0 && next < n_chunks) issue(next, stage);
  }
  if (MODE == 0) {
    if (!det.on()) {
      if (active) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          atomicAdd(red + g * 8 + e, s1[e]);
          atomicAdd(red + C + g * 8 + e, s2[e]);
        }
      }
      __syncthreads();
      for (int c = threadIdx.x; c < C; c += kBulkThreads) {
        atomicAdd(sums + (c % Cper), red[c]);
        atomicAdd(sums + Cper + (c % Cper), red[C + c]);
      }
    } else {
      // deterministic mode: the row lanes add into the block totals one after the other (RL rounds), every CTA stores
      // its totals, the last CTA adds the slices in CTA order
      for (int round = 0; round < RL; ++round) {
        if (active && rl == round) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            red[g * 8 + e] += s1[e];
            red[C + g * 8 + e] += s2[e];
          }
        }
        __syncthreads();
      }
      float* part = det.partial + (size_t)blockIdx.x * 2 * C;
      for (int c = threadIdx.x; c < 2 * C; c += kBulkThreads) part[c] = red[c];
      det_finish(det, gridDim.x, gridDim.x, 2, C, Cper, sums, sums + Cper, threadIdx.x, kBulkThreads, SyncThreads());
    }
  }
}

static bool bn_bulk_ok(long M, int C) {      // C = folded channel count (multiple of 8)
  return C % 8 == 0 && C / 8 <= kBulkThreads && (long)C * 2 <= kBulkChunkBytes && M >= 4096;
}

template <int MODE>
static int bn_bwd_bulk(const void* dy, const void* x, const void* r, const float* a, const float* b, const float* mean,
                       const float* rstd, float* sums, void* dx, void* dr, long M, int C, int Cper, int act,
                       cudaStream_t s) {
  const int rows_per_chunk = kBulkChunkBytes / (C * 2);
  const long n_chunks = (M + rows_per_chunk - 1) / rows_per_chunk;
  const int grid = (int)(n_chunks < 148 ? n_chunks : 148);
  const bool res = r != nullptr;
  DetCtx det;
  if (!det_context(kDetSlotBn, MODE == 0 ? (size_t)grid * 2 * C : 0, &det, "bn_backward")) return MU_ERR_WORKSPACE;
  if (MODE != 0) det.partial = nullptr;
  const size_t smem = 64 + ((2 * (size_t)C * 4 + 127) / 128) * 128 + (size_t)kBulkStages * (res ? 3 : 2) * kBulkChunkBytes;
#define MU_BULK(ACTC, RESC)                                                                                          \
  {                                                                                                                  \
    auto kern = bn_bwd_bulk_kernel<ACTC, RESC, MODE>;                                                                \
    set_max_dynamic_smem_once(kern, (int)smem);                              \
    kern<<<grid, kBulkThreads, smem, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)r, \
                                          a, b, mean, rstd, sums, (__nv_bfloat16*)dx, (__nv_bfloat16*)dr, M, C, Cper, \
                                          rows_per_chunk, det);                                                      \
  }
  if (res) {
    if (act == ACT_GELU) MU_BULK(ACT_GELU, true) else if (act == ACT_RELU) MU_BULK(ACT_RELU, true) else MU_BULK(ACT_NONE, true)
  } else {
    if (act == ACT_GELU) MU_BULK(ACT_GELU, false) else if (act == ACT_RELU) MU_BULK(ACT_RELU, false) else MU_BULK(ACT_NONE, false)
  }
#undef MU_BULK
  return check_launch(MODE == 0 ? "bn_bwd_reduce_bulk" : "bn_bwd_apply_bulk");
}

// ------------------------------------------------------------------ dispatch
static int pick_vec(int C) { return (C % 8 == 0) ? 8 : (C % 2 == 0 ? 2 : 1); }
// Channel counts that are not a multiple of 8 (the 150-class head) would fall back to 2- or 4-byte accesses.  Instead
// `fold` consecutive rows are treated as one wide row of fold * C columns (a multiple of 8): every thread still owns
// fixed columns, and column c is channel c mod C for the parameters and the reductions.
static int pick_fold(long M, int C) {
  if (C % 8 == 0) return 1;
  for (int f = 2; f <= 8; f *= 2)
    if ((f * C) % 8 == 0 && M % f == 0 && f * C / 8 <= kBnThreads) return f;
  return 1;
}
static int pick_grid(long M, int C, int VEC, bool reduce = false) {
  const int RL = kBnThreads / (C / VEC);
  long blocks = (M + RL - 1) / RL;
  // deterministic mode: the last CTA of a reduction kernel adds one partial per CTA, so those kernels run one CTA per
  // SM instead of eight (the loops are grid-stride)
  const long cap = (reduce && get_deterministic()) ? 148L : 148L * 8;
  return (int)(blocks < cap ? blocks : cap);
}
static bool bn_shape_ok(long M, int C) {
  if (M <= 0 || C <= 0) return false;
  return C / pick_vec(C) <= kBnThreads;
}

#define MU_BN_VEC(VECV, BODY)                          \
  switch (VECV) {                                      \
    case 8: { constexpr int VEC = 8; BODY; } break;    \
    case 2: { constexpr int VEC = 2; BODY; } break;    \
    default: { constexpr int VEC = 1; BODY; } break;   \
  }
#define MU_BN_ACT(ACTV, RESV, BODY)                                                           \
  if (RESV) {                                                                                 \
    constexpr bool RES = true;                                                                \
    switch (ACTV) {                                                                           \
      case ACT_GELU: { constexpr int ACT = ACT_GELU; BODY; } break;                           \
      case ACT_RELU: { constexpr int ACT = ACT_RELU; BODY; } break;                           \
      default: { constexpr int ACT = ACT_NONE; BODY; } break;                                 \
    }                                                                                         \
  } else {                                                                                    \
    constexpr bool RES = false;                                                               \
    switch (ACTV) {                                                                           \
      case ACT_GELU: { constexpr int ACT = ACT_GELU; BODY; } break;                           \
      case ACT_RELU: { constexpr int ACT = ACT_RELU; BODY; } break;                           \
      default: { constexpr int ACT = ACT_NONE; BODY; } break;                                 \
    }                                                                                         \
  }

template <typename T>
static int bn_stats_t(const void* x, float* sums, long M, int C, cudaStream_t s) {
  const int Cper = C, fold = pick_fold(M, C);
  M /= fold;
  C *= fold;
  const int vec = pick_vec(C), grid = pick_grid(M, C, vec, true);
  DetCtx det;
  if (!det_context(kDetSlotBn, (size_t)grid * 2 * C, &det, "bn_stats")) return MU_ERR_WORKSPACE;
  MU_BN_VEC(vec, (bn_stats_kernel<T, VEC><<<grid, kBnThreads, 0, s>>>((const T*)x, sums, M, C, Cper, det)));
  return check_launch("bn_stats");
}
template <typename T>
static int bn_apply_t(const void* x, const void* r, const float* a, const float* b, void* y, long M, int C, int act,
                      cudaStream_t s) {
  const int Cper = C, fold = pick_fold(M, C);
  M /= fold;
  C *= fold;
  const int vec = pick_vec(C), grid = pick_grid(M, C, vec);
  const bool res = r != nullptr;
  MU_BN_VEC(vec, MU_BN_ACT(act, res, (bn_apply_kernel<T, VEC, ACT, RES><<<grid, kBnThreads, 0, s>>>(
                                         (const T*)x, (const T*)r, a, b, (T*)y, M, C, Cper))));
  return check_launch("bn_apply");
}
template <typename T>
static int bn_bwd_t(const void* dy, const void* x, const void* r, const float* a, const float* b, const float* mean,
                    const float* rstd, float* sums, void* dx, void* dr, long M, int C, int act, cudaStream_t s) {
  const int Cper = C, fold = pick_fold(M, C);
  M /= fold;
  C *= fold;
  const int vec = pick_vec(C), grid = pick_grid(M, C, vec), rgrid = pick_grid(M, C, vec, true);
  const bool res = r != nullptr;
  DetCtx det;
  if (!det_context(kDetSlotBn, (size_t)rgrid * 2 * C, &det, "bn_backward")) return MU_ERR_WORKSPACE;
  MU_BN_VEC(vec, MU_BN_ACT(act, res, (bn_bwd_reduce_kernel<T, VEC, ACT, RES><<<rgrid, kBnThreads, 0, s>>>(
                                         (const T*)dy, (const T*)x, (const T*)r, a, b, mean, rstd, sums, M, C, Cper, det))));
  int rc = check_launch("bn_bwd_reduce");
  if (rc) return rc;
  MU_BN_VEC(vec, MU_BN_ACT(act, res, (bn_bwd_apply_kernel<T, VEC, ACT, RES><<<grid, kBnThreads, 0, s>>>(
                                         (const T*)dy, (const T*)x, (const T*)r, a, b, mean, rstd, sums, (T*)dx,
                                         (T*)dr, M, C, Cper))));
  return check_launch("bn_bwd_apply");
}

// per-channel sum and sum of squares of x [M, C] into sums f32 [2C] (cleared here): the statistics pass on its own
int launch_column_sums(const void* x, float* sums, long M, int C, int dtype, cudaStream_t s) {
  if (!bn_shape_ok(M, C)) {
    set_error("column_sums: unsupported shape M=%ld C=%d", M, C);
    return MU_ERR_BAD_SHAPE;
  }
  cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(float), s);
  return dtype == MU_F32 ? bn_stats_t<float>(x, sums, M, C, s) : bn_stats_t<__nv_bfloat16>(x, sums, M, C, s);
}

int launch_bn_forward(const void* x, const void* r, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, float momentum, float eps, void* y, float* mean, float* rstd, float* a,
                      float* b, float* sums, long M, int C, int act, int dtype, cudaStream_t s) {
  if (!bn_shape_ok(M, C)) {
    set_error("bn_forward: unsupported shape M=%ld C=%d", M, C);
    return MU_ERR_BAD_SHAPE;
  }
  cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(float), s);
  int rc = dtype == MU_F32 ? bn_stats_t<float>(x, sums, M, C, s) : bn_stats_t<__nv_bfloat16>(x, sums, M, C, s);
  if (rc) return rc;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(sums, gamma, beta, running_mean, running_var, M, C, momentum, eps,
                                                    mean, rstd, a, b);
  if ((rc = check_launch("bn_finalize"))) return rc;
  return dtype == MU_F32 ? bn_apply_t<float>(x, r, a, b, y, M, C, act, s)
                         : bn_apply_t<__nv_bfloat16>(x, r, a, b, y, M, C, act, s);
}

__global__ void bn_update_running_kernel(float* __restrict__ running_mean, float* __restrict__ running_var,
                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                         float momentum, float eps, long M, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float rs = rstd[c];
  float var = fmaxf(1.f / (rs * rs) - eps, 0.f);
  if (M > 1) var *= (float)((double)M / (double)(M - 1));
  running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean[c];
  running_var[c] = (1.f - momentum) * running_var[c] + momentum * var;
}

int launch_bn_update_running(float* running_mean, float* running_var, const float* mean, const float* rstd,
                             float momentum, float eps, long M, int C, cudaStream_t s) {
  bn_update_running_kernel<<<(C + 127) / 128, 128, 0, s>>>(running_mean, running_var, mean, rstd, momentum, eps, M, C);
  return check_launch("bn_update_running");
}

// statistics already reduced by the producer (the convolution epilogue): finalize + apply
int launch_bn_forward_stats(const void* x, const void* r, const float* gamma, const float* beta, float eps, void* y,
                            float* mean, float* rstd, float* a, float* b, const float* sums, long M, int C, int act,
                            int dtype, cudaStream_t s) {
  if (!bn_shape_ok(M, C)) {
    set_error("bn_forward_stats: unsupported shape M=%ld C=%d", M, C);
    return MU_ERR_BAD_SHAPE;
  }
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(sums, gamma, beta, nullptr, nullptr, M, C, 0.f, eps, mean, rstd, a,
                                                    b);
  int rc;
  if ((rc = check_launch("bn_finalize"))) return rc;
  return dtype == MU_F32 ? bn_apply_t<float>(x, r, a, b, y, M, C, act, s)
                         : bn_apply_t<__nv_bfloat16>(x, r, a, b, y, M, C, act, s);
}

int launch_bn_apply(const void* x, const void* r, const float* a, const float* b, void* y, long M, int C, int act,
                    int dtype, cudaStream_t s) {
  if (!bn_shape_ok(M, C)) {
    set_error("bn_apply: unsupported shape M=%ld C=%d", M, C);
    return MU_ERR_BAD_SHAPE;
  }
  return dtype == MU_F32 ? bn_apply_t<float>(x, r, a, b, y, M, C, act, s)
                         : bn_apply_t<__nv_bfloat16>(x, r, a, b, y, M, C, act, s);
}

int launch_bn_backward(const void* dy, const void* x, const void* r, const float* a, const float* b, const float* mean,
                       const float* rstd, float* sums, void* dx, void* dr, long M, int C, int act, int dtype,
                       cudaStream_t s) {
  if (!bn_shape_ok(M, C)) {
    set_error("bn_backward: unsupported shape M=%ld C=%d", M, C);
    return MU_ERR_BAD_SHAPE;
  }
  cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(float), s);
  if (dtype == MU_BF16) {
    const int fold = pick_fold(M, C);
    const long Mf = M / fold;
    const int Cf = C * fold;
    if (bn_bulk_ok(Mf, Cf)) {
      int rc = bn_bwd_bulk<0>(dy, x, r, a, b, mean, rstd, sums, dx, dr, Mf, Cf, C, act, s);
      if (rc) return rc;
      return bn_bwd_bulk<1>(dy, x, r, a, b, mean, rstd, sums, dx, dr, Mf, Cf, C, act, s);
    }
  }
  return dtype == MU_F32 ? bn_bwd_t<float>(dy, x, r, a, b, mean, rstd, sums, dx, dr, M, C, act, s)
                         : bn_bwd_t<__nv_bfloat16>(dy, x, r, a, b, mean, rstd, sums, dx, dr, M, C, act, s);
}

}  // namespace mu
