// K9 / K10 / K11 / A14: the remaining HBM-bound pieces of the U-Net on channels-last activations.
//   K9   MaxPool2d(2)                                   /root/reference/code/ade20k/ade_semantic.py:216
//   K10  Upsample(x2, bilinear, align_corners=True) + cat([skip, up], 1)              :235, :252
//   K11  LayerNorm([C, H, W]) over each sample (1 Mi elements at [64,128,128])        :281, :311
//   A14  CrossEntropyLoss(mean, ignore_index) fused forward + gradient                :377, :399
// Every kernel maps a thread to an 8-channel (16-byte) group of one pixel so that loads and stores are
// coalesced 16-byte vectors along the contiguous channel axis; reductions use fp32 partials, warp shuffles and
// one atomicAdd per CTA.
#include <cstdlib>

#include "common.cuh"
#include "det_reduce.cuh"

namespace mu {

// ---- 8-wide vector helpers (bf16 or fp32 storage, fp32 math) ---------------------------------------
template <typename T> struct V8;
template <> struct V8<__nv_bfloat16> {
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
template <> struct V8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

// ============================================================================ K9: MaxPool2d(2), NHWC
// BWD = 2: `out` already holds another gradient of x (the skip connection's, ade_semantic.py:304/310/312) and the pooled
// gradient is added to it in place (one fp32 add per element, rounded once: what autograd's accumulation pass computes).
template <typename T, int BWD>
__global__ void __launch_bounds__(256) maxpool2_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                       T* __restrict__ out, int B, int H, int W, int C) {
  const int G = C / 8, Ho = H / 2, Wo = W / 2;
  const long total = (long)B * Ho * Wo * G;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    long p = idx / G;
    const int wo = (int)(p % Wo);
    p /= Wo;
    const int ho = (int)(p % Ho), b = (int)(p / Ho);
    const long in0 = (((long)b * H + 2 * ho) * W + 2 * wo) * C + g * 8;
    float v[4][8];
    V8<T>::load(x + in0, v[0]);
    V8<T>::load(x + in0 + C, v[1]);
    V8<T>::load(x + in0 + (long)W * C, v[2]);
    V8<T>::load(x + in0 + (long)W * C + C, v[3]);
    if (!BWD) {
      float m[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = fmaxf(fmaxf(v[0][e], v[1][e]), fmaxf(v[2][e], v[3][e]));
      V8<T>::store(out + (((long)b * Ho + ho) * Wo + wo) * C + g * 8, m);
    } else {  // gradient goes to the first maximum in scan order (0,0) (0,1) (1,0) (1,1), as ATen does
      float g_[8], d[4][8];
      V8<T>::load(dy + (((long)b * Ho + ho) * Wo + wo) * C + g * 8, g_);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int arg = 0;
        float best = v[0][e];
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (v[k][e] > best) { best = v[k][e]; arg = k; }
#pragma unroll
        for (int k = 0; k < 4; ++k) d[k][e] = (k == arg) ? g_[e] : 0.f;
      }
      if (BWD == 2) {
        float o[4][8];
        V8<T>::load(out + in0, o[0]);
        V8<T>::load(out + in0 + C, o[1]);
        V8<T>::load(out + in0 + (long)W * C, o[2]);
        V8<T>::load(out + in0 + (long)W * C + C, o[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int e = 0; e < 8; ++e) d[k][e] += o[k][e];
      }
      V8<T>::store(out + in0, d[0]);
      V8<T>::store(out + in0 + C, d[1]);
      V8<T>::store(out + in0 + (long)W * C, d[2]);
      V8<T>::store(out + in0 + (long)W * C + C, d[3]);
    }
  }
}

// ============================================================================ K10: bilinear x2 + concat, NHWC
// out[b, h2, w2, 0:Cs] = skip;  out[b, h2, w2, Cs:Cs+Cx] = bilinear(x) with align_corners=True.
__device__ __forceinline__ void src_index(int dst, float ratio, int in_size, int& i0, int& ip, float& l0, float& l1) {
  const float r = ratio * dst;          // ATen: area_pixel_compute_source_index(align_corners=True)
  i0 = (int)r;
  ip = (i0 < in_size - 1) ? 1 : 0;
  l1 = r - i0;
  l0 = 1.f - l1;
}

// P2: channel-group count, width and height are powers of two (every U-Net shape) -- the index split is shifts and
// masks; the generic path's 64-bit divisions cost several hundred instructions per 16-byte vector and made these
// kernels instruction-bound (2.4 TB/s).
__device__ __forceinline__ int ilog2_u(unsigned v) { return 31 - __clz(v); }

template <typename T, bool P2>
__global__ void __launch_bounds__(256) upcat_fwd_kernel(const T* __restrict__ skip, const T* __restrict__ x,
                                                        T* __restrict__ out, int B, int H, int W, int Cs, int Cx) {
  const int H2 = 2 * H, W2 = 2 * W, Ct = Cs + Cx, G = Ct / 8, Gs = Cs / 8;
  const float rh = H2 > 1 ? (float)(H - 1) / (H2 - 1) : 0.f, rw = W2 > 1 ? (float)(W - 1) / (W2 - 1) : 0.f;
  const long total = (long)B * H2 * W2 * G;
  const int lg = ilog2_u(G), lw = ilog2_u(W2), lh = ilog2_u(H2);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int g, w2, h2, b;
    if (P2) {
      const unsigned u = (unsigned)idx;      // total < 2^32 is checked by the launcher
      g = u & (G - 1);
      w2 = (u >> lg) & (W2 - 1);
      h2 = (u >> (lg + lw)) & (H2 - 1);
      b = u >> (lg + lw + lh);
    } else {
      g = (int)(idx % G);
      long p = idx / G;
      w2 = (int)(p % W2);
      p /= W2;
      h2 = (int)(p % H2);
      b = (int)(p / H2);
    }
    float o[8];
    if (g < Gs) {
      V8<T>::load(skip + (((long)b * H2 + h2) * W2 + w2) * Cs + g * 8, o);
    } else {
      int h1, hp, w1, wp;
      float h0l, h1l, w0l, w1l;
      src_index(h2, rh, H, h1, hp, h0l, h1l);
      src_index(w2, rw, W, w1, wp, w0l, w1l);
      const long base = (((long)b * H + h1) * W + w1) * Cx + (g - Gs) * 8;
      float a[8], bq[8], c[8], d[8];
      V8<T>::load(x + base, a);
      V8<T>::load(x + base + (long)wp * Cx, bq);
      V8<T>::load(x + base + (long)hp * W * Cx, c);
      V8<T>::load(x + base + (long)hp * W * Cx + (long)wp * Cx, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = h0l * (w0l * a[e] + w1l * bq[e]) + h1l * (w0l * c[e] + w1l * d[e]);
    }
    V8<T>::store(out + (((long)b * H2 + h2) * W2 + w2) * Ct + g * 8, o);
  }
}

// backward: dskip = dout[..., :Cs];  dx[b, h, w, :] = sum over the <= 4x4 output pixels whose stencil touches (h, w)
template <typename T, bool P2>
__global__ void __launch_bounds__(256) upcat_bwd_kernel(const T* __restrict__ dout, T* __restrict__ dskip,
                                                        T* __restrict__ dx, int B, int H, int W, int Cs, int Cx) {
  const int H2 = 2 * H, W2 = 2 * W, Ct = Cs + Cx, Gs = Cs / 8, Gx = Cx / 8;
  const float rh = H2 > 1 ? (float)(H - 1) / (H2 - 1) : 0.f, rw = W2 > 1 ? (float)(W - 1) / (W2 - 1) : 0.f;
  const long n_skip = (long)B * H2 * W2 * Gs, n_x = (long)B * H * W * Gx;
  const int lgs = ilog2_u(Gs), lgx = ilog2_u(Gx), lw = ilog2_u(W), lh = ilog2_u(H);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n_skip + n_x;
       idx += (long)gridDim.x * blockDim.x) {
    float v[8];
    if (idx < n_skip) {
      int g;
      long pix;
      if (P2) {
        g = (unsigned)idx & (Gs - 1);
        pix = (unsigned)idx >> lgs;
      } else {
        g = (int)(idx % Gs);
        pix = idx / Gs;
      }
      V8<T>::load(dout + pix * Ct + g * 8, v);
      V8<T>::store(dskip + pix * Cs + g * 8, v);
      continue;
    }
    int g, w, h, b;
    if (P2) {
      const unsigned u = (unsigned)(idx - n_skip);
      g = u & (Gx - 1);
      w = (u >> lgx) & (W - 1);
      h = (u >> (lgx + lw)) & (H - 1);
      b = u >> (lgx + lw + lh);
    } else {
      long p = idx - n_skip;
      g = (int)(p % Gx);
      p /= Gx;
      w = (int)(p % W);
      p /= W;
      h = (int)(p % H);
      b = (int)(p / H);
    }
    float acc[8] = {};
    // candidate output rows: those with source row h (weight h0l) or h - 1 (weight h1l, when its hp == 1)
    const int h2_lo = max(0, 2 * h - 2), h2_hi = min(H2 - 1, 2 * h + 2);
    const int w2_lo = max(0, 2 * w - 2), w2_hi = min(W2 - 1, 2 * w + 2);
    for (int h2 = h2_lo; h2 <= h2_hi; ++h2) {
      int h1, hp;
      float h0l, h1l;
      src_index(h2, rh, H, h1, hp, h0l, h1l);
      float wh = 0.f;
      if (h1 == h) wh += h0l;
      if (h1 + hp == h) wh += h1l;       // (hp == 0: both taps hit row h1, as in the forward)
      if (wh == 0.f) continue;
      for (int w2 = w2_lo; w2 <= w2_hi; ++w2) {
        int w1, wp;
        float w0l, w1l;
        src_index(w2, rw, W, w1, wp, w0l, w1l);
        float ww = 0.f;
        if (w1 == w) ww += w0l;
        if (w1 + wp == w) ww += w1l;
        if (ww == 0.f) continue;
        V8<T>::load(dout + (((long)b * H2 + h2) * W2 + w2) * Ct + Cs + g * 8, v);
        const float wgt = wh * ww;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, v[e], acc[e]);
      }
    }
    V8<T>::store(dx + (((long)b * H + h) * W + w) * Cx + g * 8, acc);
  }
}

// ---- power-of-two shapes (every U-Net call): the two halves of the concat as separate, warp-uniform index ranges with
// several independent 16-byte loads in flight per thread.  The single-range kernels above mix copy lanes and
// interpolation lanes in every warp and keep one load -> store chain per thread: 3.1-3.3 TB/s at [256, 64 + 64, 128, 128]
// (tools/step_breakdown.py); same arithmetic in the same order here, so results are bit-identical.
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 a;
  __device__ __forceinline__ void ld(const __nv_bfloat16* p) { a = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void st(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = a; }
  __device__ __forceinline__ void f32(float (&v)[8]) const {
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] = __uint_as_float(w[e] << 16);
      v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
    }
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void ld(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void st(float* p) const {
    *reinterpret_cast<float4*>(p) = a;
    *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ __forceinline__ void f32(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// dst[pix, dst_off + g*8 ..] = src[pix, src_off + g*8 ..] for G = 2^lg channel groups per pixel, four vectors in flight
template <typename T>
__device__ __forceinline__ void copy_channel_slab(const T* __restrict__ src, int src_pitch, int src_off, T* __restrict__ dst,
                                                  int dst_pitch, int dst_off, unsigned n, int lg) {
  const unsigned stride = gridDim.x * blockDim.x, mask = (1u << lg) - 1u;
  for (unsigned base = blockIdx.x * blockDim.x + threadIdx.x; base < n; base += 4u * stride) {
    Raw8<T> v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned i = base + u * stride;
      if (i < n) v[u].ld(src + (size_t)(i >> lg) * src_pitch + src_off + (i & mask) * 8);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned i = base + u * stride;
      if (i < n) v[u].st(dst + (size_t)(i >> lg) * dst_pitch + dst_off + (i & mask) * 8);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) upcat_fwd_split_kernel(const T* __restrict__ skip, const T* __restrict__ x,
                                                              T* __restrict__ out, int B, int H, int W, int Cs, int Cx) {
  const int H2 = 2 * H, W2 = 2 * W, Ct = Cs + Cx, Gx = Cx / 8;
  const unsigned n_pix = (unsigned)B * H2 * W2;
  copy_channel_slab<T>(skip, Cs, 0, out, Ct, 0, n_pix * (Cs / 8), ilog2_u(Cs / 8));
  const float rh = H2 > 1 ? (float)(H - 1) / (H2 - 1) : 0.f, rw = W2 > 1 ? (float)(W - 1) / (W2 - 1) : 0.f;
  const int lg = ilog2_u(Gx), lw = ilog2_u(W2), lh = ilog2_u(H2);
  const unsigned n_up = n_pix * Gx, stride = gridDim.x * blockDim.x;
  for (unsigned base = blockIdx.x * blockDim.x + threadIdx.x; base < n_up; base += 2u * stride) {
    Raw8<T> a[2], bq[2], c[2], d[2];
    float h0l[2], h1l[2], w0l[2], w1l[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const unsigned i = base + u * stride;
      if (i < n_up) {
        const int g = i & (Gx - 1), w2 = (i >> lg) & (W2 - 1), h2 = (i >> (lg + lw)) & (H2 - 1), b = i >> (lg + lw + lh);
        int h1, hp, w1, wp;
        src_index(h2, rh, H, h1, hp, h0l[u], h1l[u]);
        src_index(w2, rw, W, w1, wp, w0l[u], w1l[u]);
        const long p0 = (((long)b * H + h1) * W + w1) * Cx + g * 8;
        a[u].ld(x + p0);
        bq[u].ld(x + p0 + (long)wp * Cx);
        c[u].ld(x + p0 + (long)hp * W * Cx);
        d[u].ld(x + p0 + (long)hp * W * Cx + (long)wp * Cx);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const unsigned i = base + u * stride;
      if (i < n_up) {
        float o[8], fa[8], fb[8], fc[8], fd[8];
        a[u].f32(fa);
        bq[u].f32(fb);
        c[u].f32(fc);
        d[u].f32(fd);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          o[e] = h0l[u] * (w0l[u] * fa[e] + w1l[u] * fb[e]) + h1l[u] * (w0l[u] * fc[e] + w1l[u] * fd[e]);
        V8<T>::store(out + (size_t)(i >> lg) * Ct + Cs + (i & (Gx - 1)) * 8, o);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) upcat_bwd_split_kernel(const T* __restrict__ dout, T* __restrict__ dskip,
                                                              T* __restrict__ dx, int B, int H, int W, int Cs, int Cx) {
  const int H2 = 2 * H, W2 = 2 * W, Ct = Cs + Cx, Gx = Cx / 8;
  copy_channel_slab<T>(dout, Ct, 0, dskip, Cs, 0, (unsigned)B * H2 * W2 * (Cs / 8), ilog2_u(Cs / 8));
  const float rh = H2 > 1 ? (float)(H - 1) / (H2 - 1) : 0.f, rw = W2 > 1 ? (float)(W - 1) / (W2 - 1) : 0.f;
  const int lg = ilog2_u(Gx), lw = ilog2_u(W), lh = ilog2_u(H);
  const unsigned n_x = (unsigned)B * H * W * Gx, stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_x; i += stride) {
    const int g = i & (Gx - 1), w = (i >> lg) & (W - 1), h = (i >> (lg + lw)) & (H - 1), b = i >> (lg + lw + lh);
    // weights of the five candidate output rows / columns 2h - 2 .. 2h + 2 (zero: no tap of that row lands on h)
    float wh[5], ww[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int h2 = 2 * h - 2 + k, w2 = 2 * w - 2 + k;
      wh[k] = ww[k] = 0.f;
      if (h2 >= 0 && h2 < H2) {
        int h1, hp;
        float l0, l1;
        src_index(h2, rh, H, h1, hp, l0, l1);
        if (h1 == h) wh[k] += l0;
        if (h1 + hp == h) wh[k] += l1;
      }
      if (w2 >= 0 && w2 < W2) {
        int w1, wp;
        float l0, l1;
        src_index(w2, rw, W, w1, wp, l0, l1);
        if (w1 == w) ww[k] += l0;
        if (w1 + wp == w) ww[k] += l1;
      }
    }
    float acc[8] = {};
    const T* base = dout + (((long)b * H2 + (2 * h - 2)) * W2 + (2 * w - 2)) * Ct + Cs + g * 8;
    // (two candidate rows per round, ten loads in flight, was measured slower: 80 registers, 0.98 against 0.90 ms per step)
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (wh[k] == 0.f) continue;
      Raw8<T> v[5];
#pragma unroll
      for (int l = 0; l < 5; ++l)
        if (ww[l] != 0.f) v[l].ld(base + ((long)k * W2 + l) * Ct);
#pragma unroll
      for (int l = 0; l < 5; ++l) {
        if (ww[l] == 0.f) continue;
        const float wgt = wh[k] * ww[l];
        float f[8];
        v[l].f32(f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, f[e], acc[e]);
      }
    }
    V8<T>::store(dx + (size_t)i * 8, acc);
  }
}

// ============================================================================ K11: LayerNorm over (C, H, W) per sample
// x, y: [B][L] with L = H*W*C in NHWC order; gamma / beta: f32 [L] in the same order.
template <typename T>
__global__ void __launch_bounds__(256) sample_stats_kernel(const T* __restrict__ x, float* __restrict__ sums, long L,
                                                           const DetCtx det) {
  // grid (chunks, B); sums[b] = (sum, sumsq)
  __shared__ float red[2][8];
  const int b = blockIdx.y;
  const T* xb = x + (long)b * L;
  float s1 = 0.f, s2 = 0.f;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < L; i += (long)gridDim.x * blockDim.x * 8) {
    float v[8];
    V8<T>::load(xb + i, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s1 += v[e];
      s2 = fmaf(v[e], v[e], s2);
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    if (det.on()) det.partial[((size_t)blockIdx.x * gridDim.y + b) * 2 + threadIdx.x] = t;   // slice = blockIdx.x
    else atomicAdd(sums + 2 * b + threadIdx.x, t);
  }
  if (det.on())
    det_finish(det, gridDim.x * gridDim.y, gridDim.x, 1, 2 * gridDim.y, 2 * gridDim.y, sums, sums, threadIdx.x, 256,
               SyncThreads());
}

__global__ void sample_finalize_kernel(const float* __restrict__ sums, float* __restrict__ mean,
                                       float* __restrict__ rstd, int B, long L, float eps) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double mu_ = (double)sums[2 * b] / (double)L;
  double var = (double)sums[2 * b + 1] / (double)L - mu_ * mu_;
  if (var < 0.0) var = 0.0;
  mean[b] = (float)mu_;
  rstd[b] = (float)(1.0 / sqrt(var + (double)eps));
}

template <typename T>
__global__ void __launch_bounds__(256) sample_ln_apply_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ rstd, T* __restrict__ y, int B,
                                                              long L) {
  // thread owns 8 consecutive positions and walks over the batch: gamma / beta are read once per thread
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < L; i += (long)gridDim.x * blockDim.x * 8) {
    float g[8], bt[8];
    V8<float>::load(gamma + i, g);
    V8<float>::load(beta + i, bt);
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
      const float mu_ = mean[b], rs = rstd[b];
      float v[8];
      V8<T>::load(x + (long)b * L + i, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (v[e] - mu_) * rs * g[e] + bt[e];
      V8<T>::store(y + (long)b * L + i, v);
    }
  }
}

// backward pass A: per-sample s1 = sum(dy*gamma), s2 = sum(dy*gamma*xhat)
template <typename T>
__global__ void __launch_bounds__(256) sample_ln_bwd_stats_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd,
                                                                  float* __restrict__ sums, long L, const DetCtx det) {
  __shared__ float red[2][8];
  const int b = blockIdx.y;
  const float mu_ = mean[b], rs = rstd[b];
  float s1 = 0.f, s2 = 0.f;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < L; i += (long)gridDim.x * blockDim.x * 8) {
    float d[8], v[8], g[8];
    V8<T>::load(dy + (long)b * L + i, d);
    V8<T>::load(x + (long)b * L + i, v);
    V8<float>::load(gamma + i, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float t = d[e] * g[e];
      s1 += t;
      s2 = fmaf(t, (v[e] - mu_) * rs, s2);
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    if (det.on()) det.partial[((size_t)blockIdx.x * gridDim.y + b) * 2 + threadIdx.x] = t;   // slice = blockIdx.x
    else atomicAdd(sums + 2 * b + threadIdx.x, t);
  }
  if (det.on())
    det_finish(det, gridDim.x * gridDim.y, gridDim.x, 1, 2 * gridDim.y, 2 * gridDim.y, sums, sums, threadIdx.x, 256,
               SyncThreads());
}

// backward pass B: dx, and dgamma / dbeta reduced over the batch by the thread that owns the position
template <typename T>
__global__ void __launch_bounds__(256) sample_ln_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd,
                                                                  const float* __restrict__ sums, T* __restrict__ dx,
                                                                  float* __restrict__ dgamma,
                                                                  float* __restrict__ dbeta, int B, long L) {
  const float invL = 1.f / (float)L;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < L; i += (long)gridDim.x * blockDim.x * 8) {
    float g[8], dg[8] = {}, db[8] = {};
    V8<float>::load(gamma + i, g);
    for (int b = 0; b < B; ++b) {
      const float mu_ = mean[b], rs = rstd[b], c1 = sums[2 * b] * invL, c2 = sums[2 * b + 1] * invL;
      float d[8], v[8], o[8];
      V8<T>::load(dy + (long)b * L + i, d);
      V8<T>::load(x + (long)b * L + i, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (v[e] - mu_) * rs;
        dg[e] = fmaf(d[e], xh, dg[e]);
        db[e] += d[e];
        o[e] = rs * (d[e] * g[e] - c1 - xh * c2);
      }
      V8<T>::store(dx + (long)b * L + i, o);
    }
    V8<float>::store(dgamma + i, dg);
    V8<float>::store(dbeta + i, db);
  }
}

// ============================================================================ A14: fused cross-entropy (mean) fwd + grad
// logits [M rows][C classes] (channels-last), labels int64 [M].  One warp per row.
//   loss_sum += lse(row) - logit[label];   dlogits = (softmax - onehot) * inv_count   (0 for ignored rows)
template <typename T>
__global__ void __launch_bounds__(256) ce_fused_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels,
                                                       const float* __restrict__ valid_count, long ignore_index,
                                                       T* __restrict__ dlogits, float* __restrict__ loss_sum, long M,
                                                       int C, int pitch, const DetCtx det) {
  // pitch >= C: row stride in elements (class-padded logits of the 1x1 head); gradient pad columns are zeroed
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const float inv = 1.f / fmaxf(valid_count[0], 1.f);
  // every row ignored: torch's mean over zero rows is NaN (0 / 0) with zero gradients
  float local = (valid_count[0] == 0.f && blockIdx.x == 0 && threadIdx.x == 0) ? NAN : 0.f;
  for (long row = (long)blockIdx.x * 8 + wib; row < M; row += (long)gridDim.x * 8) {
    const T* lr = logits + row * pitch;
    T* dr = dlogits + row * pitch;
    const long lab = labels[row];
    for (int c = C + lane; c < pitch; c += 32) st_f(dr + c, 0.f);
    if (lab == ignore_index) {
      for (int c = lane; c < C; c += 32) st_f(dr + c, 0.f);
      continue;
    }
    if (lab < 0 || lab >= C) {
      // a label outside [0, C) that is not ignore_index: nn.CrossEntropyLoss raises a device assert for it.  Kernels
      // cannot raise, so the loss and this row's gradient are poisoned with NaN -- a wrong ignore_index (e.g. 255-void
      // Cityscapes labels under the default -100) cannot train silently on garbage.
      for (int c = lane; c < C; c += 32) st_f(dr + c, NAN);
      if (lane == 0) local += NAN;
      continue;
    }
    float v[8];  // C <= 256
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = lane + 32 * k;
      v[k] = (c < C) ? ld_f(lr + c) : -INFINITY;
      mx = fmaxf(mx, v[k]);
    }
    mx = warp_max(mx);
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = __expf(v[k] - mx);       // exp(-inf) = 0 for the padding lanes
      se += v[k];
    }
    se = warp_sum(se);
    const float inv_se = 1.f / se;
    float picked = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = lane + 32 * k;
      if (c < C) {
        const float p = v[k] * inv_se;
        if (c == lab) picked = p;
        st_f(dr + c, (p - (c == lab ? 1.f : 0.f)) * inv);
      }
    }
    picked = warp_sum(picked);        // exactly one lane holds it
    if (lane == 0) local += -__logf(fmaxf(picked, 1e-38f));
  }
  __shared__ float red[8];
  if (lane == 0) red[wib] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    if (det.on()) det.partial[blockIdx.x] = t * inv;
    else atomicAdd(loss_sum, t * inv);
  }
  if (det.on()) det_finish(det, gridDim.x, gridDim.x, 1, 1, 1, loss_sum, loss_sum, threadIdx.x, 256, SyncThreads());
}

__device__ __forceinline__ float fast_exp2_so(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}

// bf16 rows whose pitch is a multiple of 8 (every production shape: 160-pitch padded logits, 32 for the instance
// heads): lane l owns the 16-byte vector of classes [8l, 8l + 8) -- one vector load and one vector store per row and
// lane instead of up to eight 2-byte accesses.
__global__ void __launch_bounds__(256) ce_fused_vec_kernel(const __nv_bfloat16* __restrict__ logits,
                                                           const int64_t* __restrict__ labels,
                                                           const float* __restrict__ valid_count, long ignore_index,
                                                           __nv_bfloat16* __restrict__ dlogits,
                                                           float* __restrict__ loss_sum, long M, int C, int pitch,
                                                           const DetCtx det) {
  constexpr int U = 4;                       // independent rows per warp trip: four vector loads in flight per lane
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const float inv = 1.f / fmaxf(valid_count[0], 1.f);
  const bool in_row = lane * 8 < pitch;
  float local = (valid_count[0] == 0.f && blockIdx.x == 0 && threadIdx.x == 0) ? NAN : 0.f;
  const long stride = (long)gridDim.x * 8;
  for (long row0 = (long)blockIdx.x * 8 + wib; row0 < M; row0 += stride * U) {
    uint4 u[U];
    long lab[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long row = row0 + k * stride;
      u[k] = make_uint4(0u, 0u, 0u, 0u);
      lab[k] = ignore_index;
      if (row < M) {
        lab[k] = labels[row];
        if (in_row && lab[k] != ignore_index) u[k] = *reinterpret_cast<const uint4*>(logits + row * pitch + lane * 8);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long row = row0 + k * stride;          // warp-uniform
      if (row >= M) continue;
      const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
      const bool bad = lab[k] != ignore_index && (lab[k] < 0 || lab[k] >= C);
      float v[8];
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float f = __uint_as_float((e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16));
        v[e] = (lane * 8 + e < C) ? f : -INFINITY;
        mx = fmaxf(mx, v[e]);
      }
      mx = warp_max(mx);
      float se = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] = __expf(v[e] - mx);
        se += v[e];
      }
      se = warp_sum(se);
      const float inv_se = 1.f / se;
      float picked = 0.f, o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = lane * 8 + e;
        const float p = v[e] * inv_se;                       // 0 for the pad classes
        if (c == lab[k]) picked = p;
        o[e] = (lab[k] == ignore_index || c >= C) ? 0.f : (p - (c == lab[k] ? 1.f : 0.f)) * inv;
        if (bad && c < C) o[e] = NAN;                        // label outside [0, C): poisoned like the scalar kernel
      }
      if (in_row) {
        uint4 r;
        r.x = pack2(o[0], o[1]); r.y = pack2(o[2], o[3]); r.z = pack2(o[4], o[5]); r.w = pack2(o[6], o[7]);
        *reinterpret_cast<uint4*>(dlogits + row * pitch + lane * 8) = r;
      }
      picked = warp_sum(picked);
      if (lane == 0 && lab[k] != ignore_index) local += bad ? NAN : -__logf(fmaxf(picked, 1e-38f));
    }
  }
  __shared__ float red[8];
  if (lane == 0) red[wib] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w2 = 0; w2 < 8; ++w2) t += red[w2];
    if (det.on()) det.partial[blockIdx.x] = t * inv;
    else atomicAdd(loss_sum, t * inv);
  }
  if (det.on()) det_finish(det, gridDim.x, gridDim.x, 1, 1, 1, loss_sum, loss_sum, threadIdx.x, 256, SyncThreads());
}

// bf16 rows, FOUR lanes per row (8 rows per warp trip): lane q of a row's quad owns the 16-byte vectors q, q + 4, ...
// (VPL of them: 40 classes at the 160-class pitch).  Against one warp per row -- 20 of 32 lanes busy at pitch 160,
// three 5-step shuffle reductions per row, ~150 warp instructions per 640 bytes: 1.5 TB/s -- the row statistics cost
// two 2-step reductions (the loss needs none: lse - x[label], the owner of the label's class subtracts its logit), the
// warp moves 8 rows per trip in 64-byte pieces, and the kernel is bound by HBM instead of by instruction issue.
template <int LPR, int VPL>
__global__ void __launch_bounds__(256, VPL >= 5 ? 2 : 3) ce_fused_quad_kernel(const __nv_bfloat16* __restrict__ logits,
                                                            const int64_t* __restrict__ labels,
                                                            const float* __restrict__ valid_count, long ignore_index,
                                                            __nv_bfloat16* __restrict__ dlogits,
                                                            float* __restrict__ loss_sum, long M, int C, int pitch,
                                                            const DetCtx det) {
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  constexpr int RPW = 32 / LPR;                            // rows per warp trip
  const int q = lane & (LPR - 1), rq = lane / LPR;         // position in the row's lane group, row of the warp trip
  const int nvec = pitch >> 3;
  const float inv = 1.f / fmaxf(valid_count[0], 1.f);
  float local = (valid_count[0] == 0.f && blockIdx.x == 0 && threadIdx.x == 0) ? NAN : 0.f;
  const long stride = (long)gridDim.x * (8 * RPW);         // 8 warps x RPW rows per block trip
  for (long row = (long)blockIdx.x * (8 * RPW) + wib * RPW + rq; row < M; row += stride) {
    const long lab = labels[row];
    const bool ignored = lab == ignore_index;
    const bool bad = !ignored && (lab < 0 || lab >= C);
    const __nv_bfloat16* src = logits + row * pitch;
    uint4 u[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vec = q + LPR * k;
      u[k] = make_uint4(0u, 0u, 0u, 0u);
      if (vec < nvec && !ignored) u[k] = *reinterpret_cast<const uint4*>(src + vec * 8);
    }
    // (per-element work is kept to unpack, max, fma + exp2 + add, mul, pack: the class-range test runs only in the
    // vector that straddles C, the one-hot term only in the vector that holds the label)
    float v[VPL][8];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
      const int c0 = (q + LPR * k) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[k][e] = __uint_as_float((e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16));
      if (c0 + 8 > C) {                                      // tail vector (or past the row): pad classes -> -inf
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (c0 + e >= C) v[k][e] = -INFINITY;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) mx = fmaxf(mx, v[k][e]);
    }
#pragma unroll
    for (int off = 1; off < LPR; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    const float mx2 = mx * kLog2e;
    float se = 0.f, x_lab = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int rel = (int)lab - (q + LPR * k) * 8;          // label's position in this vector (if in [0, 8))
      if (!ignored && rel >= 0 && rel < 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (e == rel) x_lab = v[k][e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[k][e] = fast_exp2_so(fmaf(v[k][e], kLog2e, -mx2));       // exp2(-inf) = 0 for the pad classes
        se += v[k][e];
      }
    }
#pragma unroll
    for (int off = 1; off < LPR; off <<= 1) se += __shfl_xor_sync(0xffffffffu, se, off);
    // ignored row: zero gradient; label outside [0, C): NaN gradient for the real classes, like the scalar kernel
    const float scale = ignored ? 0.f : (bad ? NAN : inv / se);
    __nv_bfloat16* dst = dlogits + row * pitch;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vec = q + LPR * k, c0 = vec * 8;
      const int rel = (int)lab - c0;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = v[k][e] * scale;
      if (!ignored && rel >= 0 && rel < 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (e == rel) o[e] -= inv;
      }
      if (c0 + 8 > C) {                                      // pad classes stay exactly zero (0 * NaN would not)
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (c0 + e >= C) o[e] = 0.f;
      }
      if (vec < nvec) {
        uint4 r;
        r.x = pack2(o[0], o[1]); r.y = pack2(o[2], o[3]); r.z = pack2(o[4], o[5]); r.w = pack2(o[6], o[7]);
        *reinterpret_cast<uint4*>(dst + c0) = r;
      }
    }
    // loss of the row = max + ln(sum) - x[label]: lane 0 of the quad adds the first two, the label's owner subtracts
    if (!ignored) {
      if (bad) local += (q == 0) ? NAN : 0.f;
      else local += ((q == 0) ? fmaf(__log2f(se), kLn2, mx) : 0.f) - x_lab;
    }
  }
  local = warp_sum(local);
  __shared__ float red[8];
  if (lane == 0) red[wib] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w2 = 0; w2 < 8; ++w2) t += red[w2];
    if (det.on()) det.partial[blockIdx.x] = t * inv;
    else atomicAdd(loss_sum, t * inv);
  }
  if (det.on()) det_finish(det, gridDim.x, gridDim.x, 1, 1, 1, loss_sum, loss_sum, threadIdx.x, 256, SyncThreads());
}

// ============================================================================ logit post-processing
// ade_semantic.py:130-131 `argmax(softmax(y_pred / 0.5, dim=1), dim=1)` and the per-class IoU loop :135-143 that runs
// inside every training step (:403) with a host sync per class.  softmax(x / 0.5) is strictly monotone in x, so the
// class map is argmax(x) with torch's tie rule (lowest index); one pass also histograms prediction, label and match
// counts per class (shared-memory integer atomics), a second tiny kernel forms mean IoU on the device.
//   hist i32 [3][C]: [0] pixels predicted c, [1] pixels labelled c, [2] pixels predicted = labelled = c
template <typename T, bool VEC8>
__global__ void __launch_bounds__(256) argmax_hist_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels,
                                                          int64_t* __restrict__ pred, int32_t* __restrict__ hist,
                                                          long M, int C, int pitch) {
  extern __shared__ int32_t sh[];   // [3][C]
  for (int i = threadIdx.x; i < 3 * C; i += 256) sh[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (long row = (long)blockIdx.x * 8 + wib; row < M; row += (long)gridDim.x * 8) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    if (VEC8) {
      if (lane * 8 < pitch) {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(logits) + row * pitch + lane * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float f = __uint_as_float((e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16));
          const int c = lane * 8 + e;
          if (c < C && (f > best || (f != f && best == best))) { best = f; bi = c; }   // first maximum; NaN wins like torch
        }
      }
    } else {
      for (int c = lane; c < C; c += 32) {
        const float f = ld_f(logits + row * pitch + c);
        if (f > best || (f != f && best == best)) { best = f; bi = c; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const bool take = (ob > best) || (ob != ob && best == best) || (ob == best && oi < bi) || (ob != ob && best != best && oi < bi);
      if (take) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      if (bi == 0x7fffffff) bi = 0;
      if (pred != nullptr) pred[row] = bi;
      if (labels != nullptr) {
        const long lab = labels[row];
        atomicAdd(sh + bi, 1);
        if (lab >= 0 && lab < C) {
          atomicAdd(sh + C + (int)lab, 1);
          if (lab == bi) atomicAdd(sh + 2 * C + bi, 1);
        }
      }
    }
  }
  __syncthreads();
  if (labels != nullptr)
    for (int i = threadIdx.x; i < 3 * C; i += 256)
      if (sh[i] != 0) atomicAdd(hist + i, sh[i]);
}

// mean over the classes with a non-empty union of (intersection + smooth) / (union + smooth)   (:137-146)
__global__ void mean_iou_finalize_kernel(const int32_t* __restrict__ hist, float* __restrict__ out, int C, float smooth) {
  __shared__ float s_sum[32], s_cnt[32];
  float acc = 0.f, cnt = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float inter = (float)hist[2 * C + c];
    const float uni = (float)(hist[c] + hist[C + c] - hist[2 * C + c]);
    if (uni > 0.f) {
      acc += (inter + smooth) / (uni + smooth);
      cnt += 1.f;
    }
  }
  acc = warp_sum(acc);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = acc; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, n = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_sum[w]; n += s_cnt[w]; }
    out[0] = a / n;            // no class present: 0 / 0 = NaN, as torch.mean of an empty stack
  }
}

// ============================================================================ launchers
static int grid_for(long items, int threads = 256) {
  long blocks = (items + threads - 1) / threads;
  const long cap = 148L * 16;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}
#define MU_T(dtype, CALL_F32, CALL_BF16) \
  if ((dtype) == MU_F32) { CALL_F32; } else { CALL_BF16; }

int launch_maxpool2(const void* x, const void* dy, void* out, int B, int H, int W, int C, int bwd, int dtype,
                    cudaStream_t s) {
  if (C % 8 || H % 2 || W % 2) {
    set_error("maxpool2: needs C %% 8 == 0 and even H, W (C=%d H=%d W=%d)", C, H, W);
    return MU_ERR_BAD_SHAPE;
  }
  const int grid = grid_for((long)B * (H / 2) * (W / 2) * (C / 8));
  if (!bwd) {
    MU_T(dtype, (maxpool2_kernel<float, 0><<<grid, 256, 0, s>>>((const float*)x, nullptr, (float*)out, B, H, W, C)),
         (maxpool2_kernel<__nv_bfloat16, 0><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, nullptr,
                                                                (__nv_bfloat16*)out, B, H, W, C)));
  } else if (bwd == 1) {
    MU_T(dtype, (maxpool2_kernel<float, 1><<<grid, 256, 0, s>>>((const float*)x, (const float*)dy, (float*)out, B, H, W, C)),
         (maxpool2_kernel<__nv_bfloat16, 1><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,
                                                                (__nv_bfloat16*)out, B, H, W, C)));
  } else {
    MU_T(dtype, (maxpool2_kernel<float, 2><<<grid, 256, 0, s>>>((const float*)x, (const float*)dy, (float*)out, B, H, W, C)),
         (maxpool2_kernel<__nv_bfloat16, 2><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,
                                                                (__nv_bfloat16*)out, B, H, W, C)));
  }
  return check_launch("maxpool2");
}

static int spatial_split_mode() {        // MU_UPCAT_SPLIT=0 (environment, A/B runs and tests): the single-range kernels
  const char* e = getenv("MU_UPCAT_SPLIT");
  return e != nullptr ? atoi(e) : 1;
}

int launch_upcat_fwd(const void* skip, const void* x, void* out, int B, int H, int W, int Cs, int Cx, int dtype,
                     cudaStream_t s) {
  if (Cs % 8 || Cx % 8) {
    set_error("upsample_concat: channel counts must be multiples of 8 (Cs=%d Cx=%d)", Cs, Cx);
    return MU_ERR_BAD_SHAPE;
  }
  const long total = (long)B * 4 * H * W * ((Cs + Cx) / 8);
  const int grid = grid_for(total);
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  const bool p2 = pow2((Cs + Cx) / 8) && pow2(Cs / 8) && pow2(Cx / 8) && pow2(H) && pow2(W) && total < (1L << 32);
  if (p2 && spatial_split_mode()) {
    MU_T(dtype, (upcat_fwd_split_kernel<float><<<grid, 256, 0, s>>>((const float*)skip, (const float*)x, (float*)out, B, H, W, Cs, Cx)),
         (upcat_fwd_split_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)skip, (const __nv_bfloat16*)x,
                                                                     (__nv_bfloat16*)out, B, H, W, Cs, Cx)));
  } else if (p2) {
    MU_T(dtype, (upcat_fwd_kernel<float, true><<<grid, 256, 0, s>>>((const float*)skip, (const float*)x, (float*)out, B, H, W, Cs, Cx)),
         (upcat_fwd_kernel<__nv_bfloat16, true><<<grid, 256, 0, s>>>((const __nv_bfloat16*)skip, (const __nv_bfloat16*)x,
                                                                     (__nv_bfloat16*)out, B, H, W, Cs, Cx)));
  } else {
    MU_T(dtype, (upcat_fwd_kernel<float, false><<<grid, 256, 0, s>>>((const float*)skip, (const float*)x, (float*)out, B, H, W, Cs, Cx)),
         (upcat_fwd_kernel<__nv_bfloat16, false><<<grid, 256, 0, s>>>((const __nv_bfloat16*)skip, (const __nv_bfloat16*)x,
                                                                      (__nv_bfloat16*)out, B, H, W, Cs, Cx)));
  }
  return check_launch("upsample_concat_fwd");
}

int launch_upcat_bwd(const void* dout, void* dskip, void* dx, int B, int H, int W, int Cs, int Cx, int dtype,
                     cudaStream_t s) {
  if (Cs % 8 || Cx % 8) {
    set_error("upsample_concat: channel counts must be multiples of 8 (Cs=%d Cx=%d)", Cs, Cx);
    return MU_ERR_BAD_SHAPE;
  }
  const long total = (long)B * 4 * H * W * (Cs / 8) + (long)B * H * W * (Cx / 8);
  const int grid = grid_for(total);
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  const bool p2 = pow2(Cs / 8) && pow2(Cx / 8) && pow2(H) && pow2(W) && total < (1L << 32);
  if (p2 && spatial_split_mode()) {
    MU_T(dtype, (upcat_bwd_split_kernel<float><<<grid, 256, 0, s>>>((const float*)dout, (float*)dskip, (float*)dx, B, H, W, Cs, Cx)),
         (upcat_bwd_split_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)dout, (__nv_bfloat16*)dskip,
                                                                     (__nv_bfloat16*)dx, B, H, W, Cs, Cx)));
  } else if (p2) {
    MU_T(dtype, (upcat_bwd_kernel<float, true><<<grid, 256, 0, s>>>((const float*)dout, (float*)dskip, (float*)dx, B, H, W, Cs, Cx)),
         (upcat_bwd_kernel<__nv_bfloat16, true><<<grid, 256, 0, s>>>((const __nv_bfloat16*)dout, (__nv_bfloat16*)dskip,
                                                                     (__nv_bfloat16*)dx, B, H, W, Cs, Cx)));
  } else {
    MU_T(dtype, (upcat_bwd_kernel<float, false><<<grid, 256, 0, s>>>((const float*)dout, (float*)dskip, (float*)dx, B, H, W, Cs, Cx)),
         (upcat_bwd_kernel<__nv_bfloat16, false><<<grid, 256, 0, s>>>((const __nv_bfloat16*)dout, (__nv_bfloat16*)dskip,
                                                                      (__nv_bfloat16*)dx, B, H, W, Cs, Cx)));
  }
  return check_launch("upsample_concat_bwd");
}

int launch_sample_ln_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, float* mean,
                         float* rstd, float* sums, int B, long L, int dtype, cudaStream_t s) {
  if (L % 8) {
    set_error("sample_layernorm: normalized size must be a multiple of 8 (L=%ld)", L);
    return MU_ERR_BAD_SHAPE;
  }
  cudaMemsetAsync(sums, 0, 2 * (size_t)B * sizeof(float), s);
  int chunks = (int)((L / 8 + 255) / 256);
  if (chunks > 64) chunks = 64;
  dim3 g1(chunks, B);
  DetCtx det;
  if (!det_context(kDetSlotSampleLn, (size_t)chunks * 2 * B, &det, "sample_layernorm_fwd")) return MU_ERR_WORKSPACE;
  MU_T(dtype, (sample_stats_kernel<float><<<g1, 256, 0, s>>>((const float*)x, sums, L, det)),
       (sample_stats_kernel<__nv_bfloat16><<<g1, 256, 0, s>>>((const __nv_bfloat16*)x, sums, L, det)));
  sample_finalize_kernel<<<(B + 127) / 128, 128, 0, s>>>(sums, mean, rstd, B, L, eps);
  int pos_blocks = (int)((L / 8 + 255) / 256);
  int by = 1;
  while ((long)pos_blocks * by < 148 * 4 && by < B) by *= 2;
  dim3 g2(pos_blocks, by);
  MU_T(dtype, (sample_ln_apply_kernel<float><<<g2, 256, 0, s>>>((const float*)x, gamma, beta, mean, rstd, (float*)y, B, L)),
       (sample_ln_apply_kernel<__nv_bfloat16><<<g2, 256, 0, s>>>((const __nv_bfloat16*)x, gamma, beta, mean, rstd,
                                                                 (__nv_bfloat16*)y, B, L)));
  return check_launch("sample_layernorm_fwd");
}

int launch_sample_ln_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                         float* sums, void* dx, float* dgamma, float* dbeta, int B, long L, int dtype,
                         cudaStream_t s) {
  if (L % 8) {
    set_error("sample_layernorm: normalized size must be a multiple of 8 (L=%ld)", L);
    return MU_ERR_BAD_SHAPE;
  }
  cudaMemsetAsync(sums, 0, 2 * (size_t)B * sizeof(float), s);
  int chunks = (int)((L / 8 + 255) / 256);
  if (chunks > 64) chunks = 64;
  dim3 g1(chunks, B);
  DetCtx det;
  if (!det_context(kDetSlotSampleLn, (size_t)chunks * 2 * B, &det, "sample_layernorm_bwd")) return MU_ERR_WORKSPACE;
  MU_T(dtype, (sample_ln_bwd_stats_kernel<float><<<g1, 256, 0, s>>>((const float*)dy, (const float*)x, gamma, mean, rstd, sums, L, det)),
       (sample_ln_bwd_stats_kernel<__nv_bfloat16><<<g1, 256, 0, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                                                                     gamma, mean, rstd, sums, L, det)));
  const int pos_blocks = (int)((L / 8 + 255) / 256);
  MU_T(dtype, (sample_ln_bwd_apply_kernel<float><<<pos_blocks, 256, 0, s>>>((const float*)dy, (const float*)x, gamma, mean, rstd,
                                                                         sums, (float*)dx, dgamma, dbeta, B, L)),
       (sample_ln_bwd_apply_kernel<__nv_bfloat16><<<pos_blocks, 256, 0, s>>>(
           (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, gamma, mean, rstd, sums, (__nv_bfloat16*)dx, dgamma, dbeta,
           B, L)));
  return check_launch("sample_layernorm_bwd");
}

int launch_ce_fused(const void* logits, const int64_t* labels, const float* valid_count, long ignore_index,
                    void* dlogits, float* loss_sum, long M, int C, int pitch, int dtype, cudaStream_t s) {
  if (pitch < C) {
    set_error("cross_entropy: row pitch %d smaller than the class count %d", pitch, C);
    return MU_ERR_BAD_SHAPE;
  }
  if (C > 256 || C < 1) {
    set_error("cross_entropy: class count must be in [1, 256] (got %d)", C);
    return MU_ERR_BAD_SHAPE;
  }
  cudaMemsetAsync(loss_sum, 0, sizeof(float), s);
  const int grid = grid_for((M + 7) / 8, 1);
  DetCtx det;
  if (!det_context(kDetSlotCe, (size_t)grid, &det, "cross_entropy")) return MU_ERR_WORKSPACE;
  if (dtype == MU_BF16 && pitch % 8 == 0) {
    // lanes per row: 4 (eight rows per warp trip, 64-byte pieces): 0.78 ms against 0.96 ms with 8 at the 160-class
    // pitch, 256 images (tools/bench_ce.py; MU_CE_LPR overrides for A/B runs)
    static const int lpr_env = [] {
      const char* e = getenv("MU_CE_LPR");
      return e != nullptr ? atoi(e) : 0;
    }();
    const int nvec = pitch / 8;
    const int lpr = lpr_env ? lpr_env : 4;
    const int grid_q = grid_for((M + 8 * (32 / lpr) - 1) / (8 * (32 / lpr)), 1);
    DetCtx det_q;
    if (!det_context(kDetSlotCe, (size_t)grid_q, &det_q, "cross_entropy")) return MU_ERR_WORKSPACE;
#define MU_CE_QUAD(L, V)                                                                                              \
  ce_fused_quad_kernel<L, V><<<grid_q, 256, 0, s>>>((const __nv_bfloat16*)logits, labels, valid_count, ignore_index, \
                                                    (__nv_bfloat16*)dlogits, loss_sum, M, C, pitch, det_q)
    const int vpl = (nvec + lpr - 1) / lpr;
    if (lpr == 8) {
      switch (vpl) {
        case 1: MU_CE_QUAD(8, 1); break;
        case 2: MU_CE_QUAD(8, 2); break;
        case 3: MU_CE_QUAD(8, 3); break;
        default: MU_CE_QUAD(8, 4); break;
      }
    } else {
      switch (vpl) {
        case 1: MU_CE_QUAD(4, 1); break;
        case 2: MU_CE_QUAD(4, 2); break;
        case 3: MU_CE_QUAD(4, 3); break;
        case 4: MU_CE_QUAD(4, 4); break;
        case 5: MU_CE_QUAD(4, 5); break;
        case 6: MU_CE_QUAD(4, 6); break;
        case 7: MU_CE_QUAD(4, 7); break;
        default: MU_CE_QUAD(4, 8); break;
      }
    }
#undef MU_CE_QUAD
    return check_launch("cross_entropy_fused_quad");
  }
  MU_T(dtype, (ce_fused_kernel<float><<<grid, 256, 0, s>>>((const float*)logits, labels, valid_count, ignore_index,
                                                        (float*)dlogits, loss_sum, M, C, pitch, det)),
       (ce_fused_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)logits, labels, valid_count,
                                                            ignore_index, (__nv_bfloat16*)dlogits, loss_sum, M, C, pitch, det)));
  return check_launch("cross_entropy_fused");
}

// pred i64 [M] (optional), hist i32 [3C] scratch (cleared here), miou f32 [1] (optional, needs labels)
int launch_argmax_iou(const void* logits, const int64_t* labels, int64_t* pred, int32_t* hist, float* miou, long M, int C,
                      int pitch, float smooth, int dtype, cudaStream_t s) {
  if (C < 1 || C > 4096 || pitch < C) {
    set_error("argmax_iou: bad class count / pitch (C=%d pitch=%d)", C, pitch);
    return MU_ERR_BAD_SHAPE;
  }
  if (labels != nullptr) cudaMemsetAsync(hist, 0, 3 * (size_t)C * sizeof(int32_t), s);
  const int grid = grid_for((M + 7) / 8, 1);
  const size_t smem = 3 * (size_t)C * sizeof(int32_t);
  if (dtype == MU_BF16 && pitch % 8 == 0 && pitch <= 256)
    argmax_hist_kernel<__nv_bfloat16, true><<<grid, 256, smem, s>>>((const __nv_bfloat16*)logits, labels, pred, hist, M, C, pitch);
  else if (dtype == MU_BF16)
    argmax_hist_kernel<__nv_bfloat16, false><<<grid, 256, smem, s>>>((const __nv_bfloat16*)logits, labels, pred, hist, M, C, pitch);
  else
    argmax_hist_kernel<float, false><<<grid, 256, smem, s>>>((const float*)logits, labels, pred, hist, M, C, pitch);
  int rc = check_launch("argmax_hist");
  if (rc || labels == nullptr || miou == nullptr) return rc;
  mean_iou_finalize_kernel<<<1, 256, 0, s>>>(hist, miou, C, smooth);
  return check_launch("mean_iou_finalize");
}

}  // namespace mu
