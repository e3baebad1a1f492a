// K3-epilogue / K4: residual + LayerNorm over channels, forward and backward.
// Replaces `attention_output + x` and nn.LayerNorm([C]) at
// /root/reference/code/ade20k/ade_semantic.py:187-188 and their autograd (:400).
// x is NCHW ([B, C, N], token-contiguous) while o / y / dz are [B, N, C] (channel-contiguous), so the
// residual needs a transposed read of x: a [C][32 tokens] tile is staged through shared memory with
// coalesced 128-byte rows, then each warp normalises tokens with shuffle reductions.
// HBM-bound: forward moves 3*B*N*C*s bytes (o, x in; y out), backward 4*B*N*C*s (dy, o, x in; dz out).
#include "common.cuh"
#include "det_reduce.cuh"

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
    float* __restrict__ dgamma, float* __restrict__ dbeta, int C, int N, const DetCtx det) {
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
    if (det.on()) {                                  // deterministic mode: store the block totals, add in CTA order
      float* part = det.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C;
      part[c] = a;
      part[C + c] = bsum;
    } else {
      atomicAdd(dgamma + c, a);
      atomicAdd(dbeta + c, bsum);
    }
  }
  if (det.on()) {
    const int n_cta = gridDim.x * gridDim.y;
    det_finish(det, n_cta, n_cta, 2, C, C, dgamma, dbeta, threadIdx.x, kLnThreads, SyncThreads());
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Production layout (bf16, token-major x): C / 8 lanes own one token, each lane one 16-byte vector of 8 channels,
// so a warp covers 32 / (C / 8) tokens per step; kLnUnroll independent steps are in flight per thread (the kernel
// is a pure HBM stream: 3 or 4 tensor passes, statistics by xor-shuffles inside the token's lane group).
constexpr int kLnUnroll = 4;

__device__ __forceinline__ void unpack8(const uint4 u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = __uint_as_float(w[e] << 16);
    v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack2_bf16(v[0], v[1]), pack2_bf16(v[2], v[3]), pack2_bf16(v[4], v[5]), pack2_bf16(v[6], v[7]));
}
template <int LPT>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPT / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int C>
__global__ void __launch_bounds__(kLnThreads) residual_ln_fwd_tok_kernel(
    const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, float* __restrict__ mean,
    float* __restrict__ rstd, long M) {
  constexpr int LPT = C / 8, TPB = kLnThreads / LPT;          // lanes per token, tokens per CTA step
  const int g = threadIdx.x % LPT, tl = threadIdx.x / LPT;
  float gm[8], bt[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    gm[e] = gamma[g * 8 + e];
    bt[e] = beta[g * 8 + e];
  }
  const long stride = (long)gridDim.x * TPB;
  for (long t0 = (long)blockIdx.x * TPB + tl; t0 < M; t0 += stride * kLnUnroll) {
    uint4 uo[kLnUnroll], ux[kLnUnroll];
#pragma unroll
    for (int u = 0; u < kLnUnroll; ++u) {
      const long t = t0 + u * stride;
      uo[u] = ux[u] = make_uint4(0u, 0u, 0u, 0u);
      if (t < M) {
        uo[u] = *reinterpret_cast<const uint4*>(o + t * C + g * 8);
        ux[u] = *reinterpret_cast<const uint4*>(x + t * C + g * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < kLnUnroll; ++u) {
      const long t = t0 + u * stride;
      const bool ok = t < M;              // no early exit: the shuffles below need every lane of the warp
      float a[8], b[8], z[8];
      unpack8(uo[u], a);
      unpack8(ux[u], b);
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        z[e] = a[e] + b[e];
        s += z[e];
      }
      const float mu_ = group_sum<LPT>(s) * (1.f / C);
      float v = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = z[e] - mu_;
        v = fmaf(d, d, v);
      }
      const float rs = rsqrtf(group_sum<LPT>(v) * (1.f / C) + eps);
#pragma unroll
      for (int e = 0; e < 8; ++e) z[e] = fmaf((z[e] - mu_) * rs, gm[e], bt[e]);
      if (ok) *reinterpret_cast<uint4*>(y + t * C + g * 8) = pack8(z);
      if (ok && g == 0) {
        mean[t] = mu_;
        rstd[t] = rs;
      }
    }
  }
}

// Per-channel parameter gradients of a CTA: lanes with the same channel group inside the warp first, then the CTA, then
// the grid (float atomics, or -- deterministic mode -- one partial per CTA added in CTA order by the last one).
template <int C>
__device__ __forceinline__ void ln_param_grads(float (&dg)[8], float (&db)[8], float* red, float* __restrict__ dgamma,
                                               float* __restrict__ dbeta, const DetCtx& det) {
  constexpr int LPT = C / 8;
  const int g = threadIdx.x % LPT;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
#pragma unroll
    for (int off = 16; off >= LPT; off >>= 1) {
      dg[e] += __shfl_xor_sync(0xffffffffu, dg[e], off);
      db[e] += __shfl_xor_sync(0xffffffffu, db[e], off);
    }
  }
  __syncthreads();
  if (!det.on()) {
    if ((threadIdx.x & 31) < LPT) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(red + g * 8 + e, dg[e]);
        atomicAdd(red + C + g * 8 + e, db[e]);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kLnThreads) {
      atomicAdd(dgamma + c, red[c]);
      atomicAdd(dbeta + c, red[C + c]);
    }
  } else {
    // deterministic mode: the warps add their channel-group totals one after the other (lanes < LPT of a warp own
    // distinct channel groups, or -- when a warp holds several tokens -- the xor-reduction above already merged them),
    // every CTA stores its totals, the last CTA adds the slices in CTA order
    for (int w = 0; w < kLnThreads / 32; ++w) {
      if ((int)(threadIdx.x >> 5) == w && (threadIdx.x & 31) < LPT) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          red[g * 8 + e] += dg[e];
          red[C + g * 8 + e] += db[e];
        }
      }
      __syncthreads();
    }
    float* part = det.partial + (size_t)blockIdx.x * 2 * C;
    for (int c = threadIdx.x; c < 2 * C; c += kLnThreads) part[c] = red[c];
    det_finish(det, gridDim.x, gridDim.x, 2, C, C, dgamma, dbeta, threadIdx.x, kLnThreads, SyncThreads());
  }
}

template <int C>
__global__ void __launch_bounds__(kLnThreads) residual_ln_bwd_tok_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ x,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
    __nv_bfloat16* __restrict__ dz, float* __restrict__ delta, float* __restrict__ dgamma, float* __restrict__ dbeta,
    long M, const DetCtx det) {
  constexpr int LPT = C / 8, TPB = kLnThreads / LPT;
  __shared__ float red[2 * C];
  const int g = threadIdx.x % LPT, tl = threadIdx.x / LPT;
  for (int i = threadIdx.x; i < 2 * C; i += kLnThreads) red[i] = 0.f;
  float gm[8], dg[8], db[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    gm[e] = gamma[g * 8 + e];
    dg[e] = db[e] = 0.f;
  }
  const long stride = (long)gridDim.x * TPB;
  for (long t0 = (long)blockIdx.x * TPB + tl; t0 < M; t0 += stride * kLnUnroll) {
    uint4 ud[kLnUnroll], uo[kLnUnroll], ux[kLnUnroll];
    float mu_[kLnUnroll], rs[kLnUnroll];
#pragma unroll
    for (int u = 0; u < kLnUnroll; ++u) {
      const long t = t0 + u * stride;
      ud[u] = uo[u] = ux[u] = make_uint4(0u, 0u, 0u, 0u);
      mu_[u] = rs[u] = 0.f;
      if (t < M) {
        ud[u] = *reinterpret_cast<const uint4*>(dy + t * C + g * 8);
        uo[u] = *reinterpret_cast<const uint4*>(o + t * C + g * 8);
        ux[u] = *reinterpret_cast<const uint4*>(x + t * C + g * 8);
        mu_[u] = mean[t];
        rs[u] = rstd[t];
      }
    }
#pragma unroll
    for (int u = 0; u < kLnUnroll; ++u) {
      const long t = t0 + u * stride;
      const bool ok = t < M;              // masked tokens carry zeros: they add nothing to dgamma / dbeta
      float d[8], ov[8], xv[8], zh[8], gg[8];
      unpack8(ud[u], d);
      unpack8(uo[u], ov);
      unpack8(ux[u], xv);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        zh[e] = (ov[e] + xv[e] - mu_[u]) * rs[u];
        gg[e] = d[e] * gm[e];
        dg[e] = fmaf(d[e], zh[e], dg[e]);
        db[e] += d[e];
        s1 += gg[e];
        s2 = fmaf(gg[e], zh[e], s2);
      }
      s1 = group_sum<LPT>(s1) * (1.f / C);
      s2 = group_sum<LPT>(s2) * (1.f / C);
      float r[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = rs[u] * (gg[e] - s1 - zh[e] * s2);
      const uint4 packed = pack8(r);
      if (ok) *reinterpret_cast<uint4*>(dz + t * C + g * 8) = packed;
      float rr[8], dl = 0.f;
      unpack8(packed, rr);                                    // delta uses the stored (rounded) gradient
#pragma unroll
      for (int e = 0; e < 8; ++e) dl = fmaf(rr[e], ov[e], dl);
      dl = group_sum<LPT>(dl);
      if (ok && g == 0) delta[t] = dl;
    }
  }
  ln_param_grads<C>(dg, db, red, dgamma, dbeta, det);
}

// ---------------------------------------------------------------------------------------------------------------
// VIEW variants: the module returns its [B, N, C] result RE-VIEWED as [B, C, H, W] (ade_semantic.py:190: .view, not
// .permute), and the channels-last network around it wants that tensor as [B, H, W, C] memory.  Element (token n, channel
// ch) of the result has flat index f = n C + ch in its sample and therefore is channel c = f / N, pixel p = f % N of
// the view.  With N = k C:  n = c k + j,  p = j C + ch  -- the C x C block of tokens {c k + j : c} (fixed j) is the
// TRANSPOSE of the C x C block of pixels {j C + ch : ch} of the channels-last view.  A CTA tile is 64 of those tokens
// (c = c0 .. c0 + 63) x all C channels: normalised per token exactly as above, written to shared memory transposed
// ([channel][64 tokens] bf16, 16-byte chunks XOR-swizzled by the channel group: conflict-free 2-byte stores and 16-byte
// loads), then stored as C rows of 128 contiguous bytes.  This replaces a separate transpose pass over y (and, in the
// backward, over dy): same values bit for bit, two tensor passes fewer per direction.
constexpr int kViewTokens = 64;

template <int C>
__global__ void __launch_bounds__(kLnThreads) residual_ln_fwd_view_kernel(
    const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, float* __restrict__ mean,
    float* __restrict__ rstd, int B, int N) {
  constexpr int LPT = C / 8, TPP = kLnThreads / LPT, PASSES = kViewTokens / TPP;   // tokens per pass, passes per tile
  constexpr int CB = C / 64;                                                       // 64-token blocks along c
  // two tiles (C <= 128): tile t + 1 is written while stragglers still store tile t -- one barrier per tile instead of two
  constexpr int NBUF = C <= 128 ? 2 : 1;
  __shared__ __align__(16) uint16_t tt_all[NBUF * C * kViewTokens];
  const int g = threadIdx.x % LPT, tl = threadIdx.x / LPT;
  float gm[8], bt[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    gm[e] = gamma[g * 8 + e];
    bt[e] = beta[g * 8 + e];
  }
  const int k = N / C, per_sample = k * CB;
  const long total = (long)B * per_sample;
  int parity = 0;
  for (long tile = blockIdx.x; tile < total; tile += gridDim.x, parity ^= (NBUF - 1)) {
    uint16_t* tt = tt_all + parity * (C * kViewTokens);
    const int b = (int)(tile / per_sample), r = (int)(tile - (long)b * per_sample);
    const int j = r / CB, cblk = r - j * CB;
    const long tok0 = (long)b * N + j;
#pragma unroll
    for (int p0 = 0; p0 < PASSES; p0 += 4) {
      uint4 uo[4], ux[4];
#pragma unroll
      for (int u = 0; u < 4 && p0 + u < PASSES; ++u) {
        const long t = tok0 + (long)(cblk * 64 + (p0 + u) * TPP + tl) * k;
        uo[u] = *reinterpret_cast<const uint4*>(o + t * C + g * 8);
        ux[u] = *reinterpret_cast<const uint4*>(x + t * C + g * 8);
      }
#pragma unroll
      for (int u = 0; u < 4 && p0 + u < PASSES; ++u) {
        const int i = (p0 + u) * TPP + tl;
        const long t = tok0 + (long)(cblk * 64 + i) * k;
        float a[8], bb[8], z[8];
        unpack8(uo[u], a);
        unpack8(ux[u], bb);
        float sm = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          z[e] = a[e] + bb[e];
          sm += z[e];
        }
        const float mu_ = group_sum<LPT>(sm) * (1.f / C);
        float v = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = z[e] - mu_;
          v = fmaf(d, d, v);
        }
        const float rs = rsqrtf(group_sum<LPT>(v) * (1.f / C) + eps);
#pragma unroll
        for (int e = 0; e < 8; ++e) z[e] = fmaf((z[e] - mu_) * rs, gm[e], bt[e]);
        const uint4 pk = pack8(z);
        const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
        uint16_t* col = tt + (g * 8) * kViewTokens + ((((i >> 3) ^ (g & 7)) << 3) | (i & 7));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          col[(2 * e) * kViewTokens] = (uint16_t)(w[e] & 0xffffu);
          col[(2 * e + 1) * kViewTokens] = (uint16_t)(w[e] >> 16);
        }
        if (g == 0) {
          mean[t] = mu_;
          rstd[t] = rs;
        }
      }
    }
    __syncthreads();
    __nv_bfloat16* yb = y + ((long)b * N + (long)j * C) * C + cblk * 64;
    for (int idx = threadIdx.x; idx < C * 8; idx += kLnThreads) {
      const int ch = idx >> 3, l = idx & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(tt + ch * kViewTokens + ((l ^ ((ch >> 3) & 7)) << 3));
      *reinterpret_cast<uint4*>(yb + (long)ch * C + l * 8) = v;
    }
    if (NBUF == 1) __syncthreads();
  }
}

template <int C>
__global__ void __launch_bounds__(kLnThreads, C == 64 ? 3 : 2) residual_ln_bwd_view_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ x,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
    __nv_bfloat16* __restrict__ dz, float* __restrict__ delta, float* __restrict__ dgamma, float* __restrict__ dbeta,
    int B, int N, const DetCtx det) {
  constexpr int LPT = C / 8, TPP = kLnThreads / LPT, PASSES = kViewTokens / TPP, CB = C / 64;
  constexpr int NBUF = C <= 128 ? 2 : 1;                  // (see the forward kernel)
  constexpr int G0 = PASSES < 4 ? PASSES : 4;             // passes whose o / x / statistics loads are issued with the dy loads
  __shared__ __align__(16) uint16_t tt_all[NBUF * C * kViewTokens];
  __shared__ float red[2 * C];
  const int g = threadIdx.x % LPT, tl = threadIdx.x / LPT;
  for (int i = threadIdx.x; i < 2 * C; i += kLnThreads) red[i] = 0.f;
  float gm[8], dg[8], db[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    gm[e] = gamma[g * 8 + e];
    dg[e] = db[e] = 0.f;
  }
  const int k = N / C, per_sample = k * CB;
  const long total = (long)B * per_sample;
  int parity = 0;
  for (long tile = blockIdx.x; tile < total; tile += gridDim.x, parity ^= (NBUF - 1)) {
    uint16_t* tt = tt_all + parity * (C * kViewTokens);
    const int b = (int)(tile / per_sample), r = (int)(tile - (long)b * per_sample);
    const int j = r / CB, cblk = r - j * CB;
    const long tok0 = (long)b * N + j;
    // dy of the view, channels-last: C rows (pixels j C + ch) of 64 contiguous channels c -> [ch][64 tokens].
    // All global loads of the tile are issued before the barrier: the dy rows AND the first passes' o / x / statistics,
    // which do not depend on the shared-memory tile (one memory round trip per tile instead of two).
    const __nv_bfloat16* dyb = dy + ((long)b * N + (long)j * C) * C + cblk * 64;
    uint4 dyv[C * 8 / kLnThreads];
#pragma unroll
    for (int it = 0; it < C * 8 / kLnThreads; ++it) {
      const int idx = it * kLnThreads + threadIdx.x, ch = idx >> 3, l = idx & 7;
      dyv[it] = *reinterpret_cast<const uint4*>(dyb + (long)ch * C + l * 8);
    }
    uint4 uo[4], ux[4];
    float mu_[4], rs[4];
#pragma unroll
    for (int u = 0; u < G0; ++u) {
      const long t = tok0 + (long)(cblk * 64 + u * TPP + tl) * k;
      uo[u] = *reinterpret_cast<const uint4*>(o + t * C + g * 8);
      ux[u] = *reinterpret_cast<const uint4*>(x + t * C + g * 8);
      mu_[u] = mean[t];
      rs[u] = rstd[t];
    }
#pragma unroll
    for (int it = 0; it < C * 8 / kLnThreads; ++it) {
      const int idx = it * kLnThreads + threadIdx.x, ch = idx >> 3, l = idx & 7;
      *reinterpret_cast<uint4*>(tt + ch * kViewTokens + ((l ^ ((ch >> 3) & 7)) << 3)) = dyv[it];
    }
    __syncthreads();
#pragma unroll
    for (int p0 = 0; p0 < PASSES; p0 += 4) {
      if (p0 > 0) {
#pragma unroll
        for (int u = 0; u < 4 && p0 + u < PASSES; ++u) {
          const long t = tok0 + (long)(cblk * 64 + (p0 + u) * TPP + tl) * k;
          uo[u] = *reinterpret_cast<const uint4*>(o + t * C + g * 8);
          ux[u] = *reinterpret_cast<const uint4*>(x + t * C + g * 8);
          mu_[u] = mean[t];
          rs[u] = rstd[t];
        }
      }
#pragma unroll
      for (int u = 0; u < 4 && p0 + u < PASSES; ++u) {
        const int i = (p0 + u) * TPP + tl;
        const long t = tok0 + (long)(cblk * 64 + i) * k;
        float d[8], ov[8], xv[8], zh[8], gg[8];
        const uint16_t* col = tt + (g * 8) * kViewTokens + ((((i >> 3) ^ (g & 7)) << 3) | (i & 7));
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] = __uint_as_float((uint32_t)col[e * kViewTokens] << 16);
        unpack8(uo[u], ov);
        unpack8(ux[u], xv);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          zh[e] = (ov[e] + xv[e] - mu_[u]) * rs[u];
          gg[e] = d[e] * gm[e];
          dg[e] = fmaf(d[e], zh[e], dg[e]);
          db[e] += d[e];
          s1 += gg[e];
          s2 = fmaf(gg[e], zh[e], s2);
        }
        s1 = group_sum<LPT>(s1) * (1.f / C);
        s2 = group_sum<LPT>(s2) * (1.f / C);
        float rr_[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) rr_[e] = rs[u] * (gg[e] - s1 - zh[e] * s2);
        const uint4 packed = pack8(rr_);
        *reinterpret_cast<uint4*>(dz + t * C + g * 8) = packed;
        float rr[8], dl = 0.f;
        unpack8(packed, rr);                                    // delta uses the stored (rounded) gradient
#pragma unroll
        for (int e = 0; e < 8; ++e) dl = fmaf(rr[e], ov[e], dl);
        dl = group_sum<LPT>(dl);
        if (g == 0) delta[t] = dl;
      }
    }
    if (NBUF == 1) __syncthreads();
  }
  ln_param_grads<C>(dg, db, red, dgamma, dbeta, det);
}

static int ln_view_grid(int B, int N, int C, bool reduce = false) {
  const long tiles = (long)B * (N / C) * (C / 64);
  const long cap = (reduce && get_deterministic()) ? 148L : 148L * (C == 256 ? 4 : 8);
  return (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}

static int ln_tok_grid(long M, int C, bool reduce = false) {
  const long per = kLnThreads / (C / 8) * kLnUnroll;
  long blocks = (M + per - 1) / per;
  // deterministic mode: one partial per CTA is added by the last CTA, so the backward runs one CTA per SM
  const long cap = (reduce && get_deterministic()) ? 148L : 148L * 8;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
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
  DetCtx det;
  if (!det_context(kDetSlotLn, (size_t)grid.x * grid.y * 2 * C, &det, "residual_ln_bwd")) return MU_ERR_WORKSPACE;
  if (tok)
    residual_ln_bwd_kernel<T, true><<<grid, kLnThreads, ln_smem(C), s>>>(
        (const T*)dy, (const T*)o, (const T*)x, mean, rstd, gamma, (T*)dz, delta, dgamma, dbeta, C, N, det);
  else
    residual_ln_bwd_kernel<T, false><<<grid, kLnThreads, ln_smem(C), s>>>(
        (const T*)dy, (const T*)o, (const T*)x, mean, rstd, gamma, (T*)dz, delta, dgamma, dbeta, C, N, det);
  return check_launch("residual_ln_bwd");
}

int launch_residual_ln_fwd(const void* o, const void* x, const float* gamma, const float* beta, float eps, void* y,
                           float* mean, float* rstd, int B, int C, int N, int dtype, int tok, cudaStream_t s) {
  if (tok == 2) {
    MU_REQUIRE(dtype == MU_BF16 && (C == 64 || C == 128 || C == 256) && N % C == 0, MU_ERR_BAD_SHAPE,
               "residual_ln_fwd: the re-viewed channels-last output needs bf16, C in {64, 128, 256} and N %% C == 0 "
               "(C=%d N=%d dtype=%d)", C, N, dtype);
    const int grid = ln_view_grid(B, N, C);
#define MU_LN_FWDV(CC)                                                                                                 \
  residual_ln_fwd_view_kernel<CC><<<grid, kLnThreads, 0, s>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)x, gamma, \
                                                              beta, eps, (__nv_bfloat16*)y, mean, rstd, B, N)
    if (C == 64) MU_LN_FWDV(64); else if (C == 128) MU_LN_FWDV(128); else MU_LN_FWDV(256);
#undef MU_LN_FWDV
    return check_launch("residual_ln_fwd_view");
  }
  if (dtype == MU_BF16 && tok && (C == 64 || C == 128 || C == 256)) {
    const long M = (long)B * N;
    const int grid = ln_tok_grid(M, C);
#define MU_LN_FWD(CC)                                                                                              \
  residual_ln_fwd_tok_kernel<CC><<<grid, kLnThreads, 0, s>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)x, gamma, \
                                                             beta, eps, (__nv_bfloat16*)y, mean, rstd, M)
    if (C == 64) MU_LN_FWD(64); else if (C == 128) MU_LN_FWD(128); else MU_LN_FWD(256);
#undef MU_LN_FWD
    return check_launch("residual_ln_fwd_tok");
  }
  if (dtype == MU_F32) return run_fwd<float>(o, x, gamma, beta, eps, y, mean, rstd, B, C, N, tok, s);
  return run_fwd<__nv_bfloat16>(o, x, gamma, beta, eps, y, mean, rstd, B, C, N, tok, s);
}
int launch_residual_ln_bwd(const void* dy, const void* o, const void* x, const float* mean, const float* rstd,
                           const float* gamma, void* dz, float* delta, float* dgamma, float* dbeta, int B, int C, int N,
                           int dtype, int tok, cudaStream_t s) {
  if (tok == 2) {
    MU_REQUIRE(dtype == MU_BF16 && (C == 64 || C == 128 || C == 256) && N % C == 0, MU_ERR_BAD_SHAPE,
               "residual_ln_bwd: the re-viewed channels-last gradient needs bf16, C in {64, 128, 256} and N %% C == 0 "
               "(C=%d N=%d dtype=%d)", C, N, dtype);
    const int grid = ln_view_grid(B, N, C, true);
    DetCtx det;
    if (!det_context(kDetSlotLn, (size_t)grid * 2 * C, &det, "residual_ln_bwd")) return MU_ERR_WORKSPACE;
#define MU_LN_BWDV(CC)                                                                                                \
  residual_ln_bwd_view_kernel<CC><<<grid, kLnThreads, 0, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)o,       \
                                                              (const __nv_bfloat16*)x, mean, rstd, gamma,             \
                                                              (__nv_bfloat16*)dz, delta, dgamma, dbeta, B, N, det)
    if (C == 64) MU_LN_BWDV(64); else if (C == 128) MU_LN_BWDV(128); else MU_LN_BWDV(256);
#undef MU_LN_BWDV
    return check_launch("residual_ln_bwd_view");
  }
  if (dtype == MU_BF16 && tok && (C == 64 || C == 128 || C == 256)) {
    const long M = (long)B * N;
    const int grid = ln_tok_grid(M, C, true);
    DetCtx det;
    if (!det_context(kDetSlotLn, (size_t)grid * 2 * C, &det, "residual_ln_bwd")) return MU_ERR_WORKSPACE;
#define MU_LN_BWD(CC)                                                                                               \
  residual_ln_bwd_tok_kernel<CC><<<grid, kLnThreads, 0, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)o,        \
                                                             (const __nv_bfloat16*)x, mean, rstd, gamma,              \
                                                             (__nv_bfloat16*)dz, delta, dgamma, dbeta, M, det)
    if (C == 64) MU_LN_BWD(64); else if (C == 128) MU_LN_BWD(128); else MU_LN_BWD(256);
#undef MU_LN_BWD
    return check_launch("residual_ln_bwd_tok");
  }
  if (dtype == MU_F32) return run_bwd<float>(dy, o, x, mean, rstd, gamma, dz, delta, dgamma, dbeta, B, C, N, tok, s);
  return run_bwd<__nv_bfloat16>(dy, o, x, mean, rstd, gamma, dz, delta, dgamma, dbeta, B, C, N, tok, s);
}

}  // namespace mu
