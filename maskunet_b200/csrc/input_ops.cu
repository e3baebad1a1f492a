// SURVEY 8(f) rank 3 (input pipeline, device half): torchvision's ToTensor() as the scripts apply it to the cv2 image
// (/root/reference/code/ade20k/ade_semantic.py:56-79,97: `transforms=ToTensor()` on an RGB uint8 HWC array):
//     tensor = img.permute(2, 0, 1).float().div(255)          -> float CHW in [0, 1]
// Here the batch stays uint8 HWC through the host->device copy (4x fewer bytes than fp32) and one kernel writes the
// network input directly: NCHW (reference layout) or channels-last with the channel count zero-padded (the stem
// convolution of the tcgen05 path reads 8-channel pixels).  IEEE division by 255.0f: bit-exact with ToTensor.
// The resize (cv2.INTER_LINEAR fixed-point arithmetic, a third-party dependency absent here) stays on the host.
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

}  // namespace mu
