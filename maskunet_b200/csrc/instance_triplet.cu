// InstanceContrastiveLoss on the device.
// /root/reference/code/coco/coco_panoptic.py:482-521 (same class: ade20k/ade_panoptic.py:390-, cityscapes/city_panoptic.py:426-,
// cityscapes/city_instance.py:279-307 with a 255 ignore value).  The reference loops in Python over
// torch.unique(instance_mask) with two .nonzero() host syncs per instance; here the host gets ONE small copy (ids and
// pixel counts, needed to draw the negative ranks from the CPU generator exactly as :510 does) and everything else is
// two kernels, one CTA per instance.
//
// Inputs prepared by the caller (maskunet_b200/losses.py):
//   order  int64 [M]     pixel positions (row-major over [B, H, W]) grouped by instance id, ascending inside a group
//                        (a stable sort of the ids) == the order of .nonzero() at :497
//   meta   int64 [3, K]  per qualifying instance: offset of its group in `order`, its pixel count, and the rank k of the
//                        negative among the pixels that do NOT belong to it (the torch.randint draw of :510)
// Selection (:502-503, :511): anchor / positive = first two pixels of the instance, negative = k-th non-member pixel.
// The reference indexes  sem_mask[:, :, idx0, idx1]  with the first two components of the [B, H, W] nonzero tuple, i.e.
// (batch index, row index) of the pixel are used as (h, w) of the logits -- reproduced here.  An index outside
// [0, H) x [0, W) is an IndexError in the reference; here it poisons the loss with NaN (no out-of-bounds access).
#include "common.cuh"

namespace mu {

constexpr int kTripletThreads = 256;

__device__ __forceinline__ float block_sum_256(float v, float* scratch) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  float t = (threadIdx.x < kTripletThreads / 32) ? scratch[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) scratch[0] = t;
  }
  __syncthreads();
  return scratch[0];
}

// sel int32 [K, 6]: (h, w) of anchor, positive, negative; -1 marks an index the reference would reject
__global__ void __launch_bounds__(kTripletThreads)
triplet_select_kernel(const int64_t* __restrict__ order, const int64_t* __restrict__ meta, int K, int H, int W,
                      int32_t* __restrict__ sel) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const int64_t off = meta[i], cnt = meta[K + i], k = meta[2 * (int64_t)K + i];
  const int64_t* P = order + off;
  // k-th (0-based) pixel outside the group: x = k + t with t = number of members below x, i.e. the first t with
  // P[t] - t > k (P[t] - t = number of non-members below member t, non-decreasing in t)
  int64_t lo = 0, hi = cnt;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (P[mid] - mid <= k) lo = mid + 1; else hi = mid;
  }
  const int64_t pos[3] = {P[0], P[1], k + lo};
  const int64_t hw = (int64_t)H * W;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int64_t b = pos[j] / hw, h = (pos[j] / W) % H;   // (dim-0, dim-1) index of the pixel in [B, H, W]
    const bool ok = b < H && h < W;                         // used as (h, w) of sem_mask[:, :, h, w]
    sel[6 * i + 2 * j] = ok ? (int32_t)b : -1;
    sel[6 * i + 2 * j + 1] = ok ? (int32_t)h : -1;
  }
}

// dist f32 [K, 2] = (||a - p + eps||, ||a - n + eps||) over the B*C entries of the three logit columns;
// loss += max(d_ap - d_an + margin, 0) / K
template <typename T>
__global__ void __launch_bounds__(kTripletThreads)
triplet_fwd_kernel(const T* __restrict__ sem, long sb, long sc, long sh, long sw, int B, int C,
                   const int32_t* __restrict__ sel, int K, float margin, float eps, float* __restrict__ dist,
                   float* __restrict__ loss) {
  __shared__ float scratch[kTripletThreads / 32];
  const int i = blockIdx.x;
  const int32_t* s = sel + 6 * i;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) ok = ok && s[j] >= 0;
  if (!ok) {
    if (threadIdx.x == 0) {
      dist[2 * i] = dist[2 * i + 1] = __int_as_float(0x7fc00000);
      atomicAdd(loss, __int_as_float(0x7fc00000));
    }
    return;
  }
  const long oa = s[0] * sh + s[1] * sw, op = s[2] * sh + s[3] * sw, on = s[4] * sh + s[5] * sw;
  float sap = 0.f, san = 0.f;
  for (int e = threadIdx.x; e < B * C; e += kTripletThreads) {
    const T* base = sem + (long)(e / C) * sb + (long)(e % C) * sc;
    const float a = ld_f(base + oa), p = ld_f(base + op), n = ld_f(base + on);
    const float d1 = a - p + eps, d2 = a - n + eps;
    sap = fmaf(d1, d1, sap);
    san = fmaf(d2, d2, san);
  }
  sap = block_sum_256(sap, scratch);
  san = block_sum_256(san, scratch);
  if (threadIdx.x == 0) {
    const float dap = sqrtf(sap), dan = sqrtf(san);
    dist[2 * i] = dap;
    dist[2 * i + 1] = dan;
    atomicAdd(loss, fmaxf(dap - dan + margin, 0.f) / (float)K);
  }
}

template <typename T> __device__ __forceinline__ void atomic_add_t(T* p, float v);
template <> __device__ __forceinline__ void atomic_add_t<float>(float* p, float v) { atomicAdd(p, v); }
template <> __device__ __forceinline__ void atomic_add_t<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  atomicAdd(p, __float2bfloat16_rn(v));
}

// dsem += scale * dloss * d(loss)/d(sem): three logit columns per active instance (atomics: columns may coincide)
template <typename T>
__global__ void __launch_bounds__(kTripletThreads)
triplet_bwd_kernel(const T* __restrict__ sem, long sb, long sc, long sh, long sw, int B, int C,
                   const int32_t* __restrict__ sel, int K, float margin, float eps, const float* __restrict__ dist,
                   const float* __restrict__ dloss, float scale, T* __restrict__ dsem, long gb, long gc, long gh,
                   long gw) {
  const int i = blockIdx.x;
  const int32_t* s = sel + 6 * i;
#pragma unroll
  for (int j = 0; j < 6; ++j)
    if (s[j] < 0) return;
  const float dap = dist[2 * i], dan = dist[2 * i + 1];
  if (!(dap - dan + margin > 0.f)) return;                 // hinge inactive (clamp_min(., 0) has zero gradient)
  const float g = scale * (dloss ? dloss[0] : 1.f) / (float)K;
  const float rap = dap > 0.f ? g / dap : 0.f, ran = dan > 0.f ? g / dan : 0.f;   // the norm's gradient at 0 is 0
  const long oa = s[0] * sh + s[1] * sw, op = s[2] * sh + s[3] * sw, on = s[4] * sh + s[5] * sw;
  const long ga = s[0] * gh + s[1] * gw, gp = s[2] * gh + s[3] * gw, gn = s[4] * gh + s[5] * gw;
  for (int e = threadIdx.x; e < B * C; e += kTripletThreads) {
    const int bb = e / C, c = e % C;
    const T* base = sem + (long)bb * sb + (long)c * sc;
    T* gbase = dsem + (long)bb * gb + (long)c * gc;
    const float a = ld_f(base + oa), p = ld_f(base + op), n = ld_f(base + on);
    const float u = (a - p + eps) * rap, v = (a - n + eps) * ran;
    atomic_add_t(gbase + ga, u - v);
    atomic_add_t(gbase + gp, -u);
    atomic_add_t(gbase + gn, v);
  }
}

int launch_instance_triplet_fwd(const void* sem, const long* st, int B, int C, int H, int W, const int64_t* order,
                                const int64_t* meta, int K, float margin, float eps, int32_t* sel, float* dist,
                                float* loss, int dtype, cudaStream_t s) {
  cudaMemsetAsync(loss, 0, sizeof(float), s);
  if (K == 0) return 0;
  triplet_select_kernel<<<(K + kTripletThreads - 1) / kTripletThreads, kTripletThreads, 0, s>>>(order, meta, K, H, W, sel);
  if (int rc = check_launch("triplet_select")) return rc;
  if (dtype == MU_F32)
    triplet_fwd_kernel<float><<<K, kTripletThreads, 0, s>>>((const float*)sem, st[0], st[1], st[2], st[3], B, C, sel, K,
                                                           margin, eps, dist, loss);
  else
    triplet_fwd_kernel<__nv_bfloat16><<<K, kTripletThreads, 0, s>>>((const __nv_bfloat16*)sem, st[0], st[1], st[2],
                                                                   st[3], B, C, sel, K, margin, eps, dist, loss);
  return check_launch("triplet_fwd");
}

int launch_instance_triplet_bwd(const void* sem, const long* st, int B, int C, const int32_t* sel, int K, float margin,
                                float eps, const float* dist, const float* dloss, float scale, void* dsem,
                                const long* gst, int dtype, cudaStream_t s) {
  if (K == 0) return 0;
  if (dtype == MU_F32)
    triplet_bwd_kernel<float><<<K, kTripletThreads, 0, s>>>((const float*)sem, st[0], st[1], st[2], st[3], B, C, sel, K,
                                                           margin, eps, dist, dloss, scale, (float*)dsem, gst[0],
                                                           gst[1], gst[2], gst[3]);
  else
    triplet_bwd_kernel<__nv_bfloat16><<<K, kTripletThreads, 0, s>>>(
        (const __nv_bfloat16*)sem, st[0], st[1], st[2], st[3], B, C, sel, K, margin, eps, dist, dloss, scale,
        (__nv_bfloat16*)dsem, gst[0], gst[1], gst[2], gst[3]);
  return check_launch("triplet_bwd");
}

}  // namespace mu
