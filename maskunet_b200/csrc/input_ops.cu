// SURVEY 8(f) rank 3 (input pipeline, device half): torchvision's ToTensor() as the scripts apply it to the cv2 image
// (/root/reference/code/ade20k/ade_semantic.py:56-79,97: `transforms=ToTensor()` on an RGB uint8 HWC array):
//     tensor = img.permute(2, 0, 1).float().div(255)          -> float CHW in [0, 1]
// Here the batch stays uint8 HWC through the host->device copy (4x fewer bytes than fp32) and one kernel writes the
// network input directly: NCHW (reference layout) or channels-last with the channel count zero-padded (the stem
// convolution of the tcgen05 path reads 8-channel pixels).  IEEE division by 255.0f: bit-exact with ToTensor.
// The resize itself (cv2.INTER_LINEAR / INTER_NEAREST, :72-73) is further down in this file.
#include "common.cuh"

namespace mu {

template <typename T>
__global__ void to_tensor_u8_kernel(const uint8_t* __restrict__ img, T* __restrict__ out, long pixels, int hw, int cin,
                                    int cpad, int channels_last) {
  const long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (p >= pixels) return;
  const uint8_t* src = img + p * cin;
  if (channels_last) {
    T* dst = out + p * cpad;
    for (int c = 0; c < cpad; ++c) st_f(dst + c, c < cin ? (float)src[c] / 255.0f : 0.f);
  } else {
    const long b = p / hw, s = p % hw;
    for (int c = 0; c < cin; ++c) st_f(out + (b * cin + c) * hw + s, (float)src[c] / 255.0f);
  }
}

int launch_to_tensor_u8(const uint8_t* img, void* out, int B, int H, int W, int Cin, int Cpad, int channels_last,
                        int dtype, cudaStream_t s) {
  const long pixels = (long)B * H * W;
  const int threads = 256;
  const unsigned blocks = (unsigned)((pixels + threads - 1) / threads);
  if (dtype == MU_F32)
    to_tensor_u8_kernel<float><<<blocks, threads, 0, s>>>(img, (float*)out, pixels, H * W, Cin, Cpad, channels_last);
  else
    to_tensor_u8_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>(img, (__nv_bfloat16*)out, pixels, H * W, Cin, Cpad,
                                                                 channels_last);
  return check_launch("to_tensor_u8");
}

// ============================================================================ cv2.resize on the device
// /root/reference/code/ade20k/ade_semantic.py:72-73 (every dataset class has the same two calls):
//     image_rgb = cv2.resize(image_rgb, (W, H), interpolation=cv2.INTER_LINEAR)     uint8 HWC
//     mask      = cv2.resize(mask,      (W, H), interpolation=cv2.INTER_NEAREST)    uint8 HW  -> .long()
// OpenCV's uint8 arithmetic (modules/imgproc/src/resize.cpp; third-party, restated in oracle/resize_oracle.py and
// pinned against cv2 4.13 itself), reproduced bit for bit:
//   taps       f = float((d + 0.5) * scale - 0.5) with the product and the sum in double, s = floor(f), frac = f - s;
//              columns clamp (s < 0 -> 0 / frac 0; s >= sw - 1 -> sw - 1 / frac 0); rows keep the fraction and clip the
//              two row indices; weights = cvRound(w * 2048) saturated to int16 (round half to even)
//   horizontal int32  S[s] * a0 + S[s + 1] * a1
//   vertical   (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
//   exact 2x down-scaling in both directions is INTER_AREA in OpenCV: (a + b + c + d + 2) >> 2
// Every floating-point step uses the explicitly rounded intrinsics: the compiler may not contract them into FMAs,
// which round differently.  The kernel then applies ToTensor (/ 255) and writes the network-input layout directly, so
// an image crosses PCIe once, as the uint8 file content, at its original size.
__device__ __forceinline__ void linear_tap(int d, double scale, int s_len, bool clamp_fraction, int& s, int& a0, int& a1) {
  const double fd = __dadd_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), -0.5);
  const float f = __double2float_rn(fd);
  const float fl = floorf(f);
  s = (int)fl;
  float frac = __fsub_rn(f, fl);
  if (clamp_fraction) {
    if (s < 0) { s = 0; frac = 0.f; }
    if (s >= s_len - 1) { s = s_len - 1; frac = 0.f; }
  }
  int w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, frac), 2048.f));
  int w1 = __float2int_rn(__fmul_rn(frac, 2048.f));
  a0 = min(max(w0, -32768), 32767);
  a1 = min(max(w1, -32768), 32767);
}

template <typename T>
__global__ void resize_linear_to_tensor_kernel(const uint8_t* __restrict__ img, T* __restrict__ out, int sh, int sw,
                                               int cin, int oh, int ow, int cpad, int channels_last, double scale_x,
                                               double scale_y, int area2x, int normalise) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= ow) return;
  int xs = 0, xa0 = 0, xa1 = 0, ys = 0, yb0 = 0, yb1 = 0, x1 = 0, y0 = 0, y1 = 0;
  if (!area2x) {
    linear_tap(dx, scale_x, sw, true, xs, xa0, xa1);
    linear_tap(dy, scale_y, sh, false, ys, yb0, yb1);
    x1 = min(xs + 1, sw - 1);
    y0 = min(max(ys, 0), sh - 1);
    y1 = min(max(ys + 1, 0), sh - 1);
  }
  const long p = (long)dy * ow + dx;
  for (int c = 0; c < (channels_last ? cpad : cin); ++c) {
    float v = 0.f;
    if (c < cin) {
      int q;
      if (area2x) {
        const uint8_t* r0 = img + ((long)(2 * dy) * sw + 2 * dx) * cin + c;
        const uint8_t* r1 = r0 + (long)sw * cin;
        q = ((int)r0[0] + (int)r0[cin] + (int)r1[0] + (int)r1[cin] + 2) >> 2;
      } else {
        const uint8_t* r0 = img + (long)y0 * sw * cin + c;
        const uint8_t* r1 = img + (long)y1 * sw * cin + c;
        const int h0 = (int)r0[(long)xs * cin] * xa0 + (int)r0[(long)x1 * cin] * xa1;
        const int h1 = (int)r1[(long)xs * cin] * xa0 + (int)r1[(long)x1 * cin] * xa1;
        q = (((yb0 * (h0 >> 4)) >> 16) + ((yb1 * (h1 >> 4)) >> 16) + 2) >> 2;
        q = min(max(q, 0), 255);
      }
      v = normalise ? __fdiv_rn((float)q, 255.0f) : (float)q;
    }
    if (channels_last) st_f(out + p * cpad + c, v);
    else st_f(out + (long)c * oh * ow + p, v);
  }
}

__global__ void resize_nearest_u8_i64_kernel(const uint8_t* __restrict__ mask, int64_t* __restrict__ out, int sh, int sw,
                                             int oh, int ow, double ifx, double ify) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= ow) return;
  const int sx = min((int)floor(__dmul_rn((double)dx, ifx)), sw - 1);
  const int sy = min((int)floor(__dmul_rn((double)dy, ify)), sh - 1);
  out[(long)dy * ow + dx] = (int64_t)mask[(long)sy * sw + sx];
}

int launch_resize_linear_to_tensor(const uint8_t* img, void* out, int sh, int sw, int cin, int oh, int ow, int cpad,
                                   int channels_last, int normalise, int dtype, cudaStream_t s) {
  const double scale_x = 1.0 / ((double)ow / (double)sw), scale_y = 1.0 / ((double)oh / (double)sh);
  const int area2x = (sh == 2 * oh && sw == 2 * ow) ? 1 : 0;
  dim3 grid((ow + 127) / 128, oh);
  if (dtype == MU_F32)
    resize_linear_to_tensor_kernel<float><<<grid, 128, 0, s>>>(img, (float*)out, sh, sw, cin, oh, ow, cpad, channels_last,
                                                               scale_x, scale_y, area2x, normalise);
  else
    resize_linear_to_tensor_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(img, (__nv_bfloat16*)out, sh, sw, cin, oh, ow, cpad,
                                                                       channels_last, scale_x, scale_y, area2x, normalise);
  return check_launch("resize_linear_to_tensor");
}

int launch_resize_nearest_u8_i64(const uint8_t* mask, int64_t* out, int sh, int sw, int oh, int ow, cudaStream_t s) {
  const double ifx = 1.0 / ((double)ow / (double)sw), ify = 1.0 / ((double)oh / (double)sh);
  dim3 grid((ow + 127) / 128, oh);
  resize_nearest_u8_i64_kernel<<<grid, 128, 0, s>>>(mask, out, sh, sw, oh, ow, ifx, ify);
  return check_launch("resize_nearest_u8_i64");
}

}  // namespace mu
