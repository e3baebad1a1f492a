// Batched 2-D transpose [batch][rows][cols] -> [batch][cols][rows] through a shared-memory tile.
// This is the NCHW <-> NHWC bridge around the Mask Attention Module: the reference returns the [B, N, C]
// result re-viewed as NCHW (/root/reference/code/ade20k/ade_semantic.py:190), while cuDNN's tensor-core
// convolutions want channels-last; PyTorch's generic strided copy is uncoalesced on one side.
// HBM-bound: one read + one write of the tensor, 64x64 tiles, both sides coalesced.
#include "common.cuh"

namespace mu {

template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int rows,
                                                        int cols) {
  __shared__ T tile[64][64 + (4 / sizeof(T) > 0 ? 4 / sizeof(T) : 1)];
  const size_t base = (size_t)blockIdx.z * rows * cols;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = r0 + ty + 4 * i, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + 4 * i][tx] = in[base + (size_t)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = c0 + ty + 4 * i, r = r0 + tx;
    if (r < rows && c < cols) out[base + (size_t)c * rows + r] = tile[tx][ty + 4 * i];
  }
}

// 2-byte elements, rows and cols multiples of 8: 16-byte accesses on both sides.  Load: thread = (row, 8-column chunk);
// store: 8 consecutive lanes write the 8 chunks of one output row segment (a full 128-byte line), gathering their 8
// elements down a tile column.  (The generic kernel's 2-byte accesses ran at 2.7 TB/s.)
__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                                                          int rows, int cols) {
  __shared__ uint16_t tile[64][64 + 2];                     // 132-byte pitch: 33 words, odd
  const size_t base = (size_t)blockIdx.z * rows * cols;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int t = threadIdx.x;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int r = (t >> 3) + 32 * pass, ch = t & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + r < rows && c0 + ch * 8 < cols) v = *reinterpret_cast<const uint4*>(in + base + (size_t)(r0 + r) * cols + c0 + ch * 8);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&tile[r][ch * 8]);   // 4-byte aligned: (132 r + 16 ch) bytes
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int k = t & 7, c = (t >> 3) + 32 * pass;          // output row c0 + c, input rows r0 + 8k .. + 8
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      w[e] = (uint32_t)tile[8 * k + 2 * e][c] | ((uint32_t)tile[8 * k + 2 * e + 1][c] << 16);
    if (c0 + c < cols && r0 + 8 * k < rows)
      *reinterpret_cast<uint4*>(out + base + (size_t)(c0 + c) * rows + r0 + 8 * k) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

int launch_transpose(const void* in, void* out, int batch, int rows, int cols, int elem_bytes, cudaStream_t s) {
  dim3 grid((cols + 63) / 64, (rows + 63) / 64, batch);
  if (elem_bytes == 2 && rows % 8 == 0 && cols % 8 == 0)
    transpose16_kernel<<<grid, 256, 0, s>>>((const uint16_t*)in, (uint16_t*)out, rows, cols);
  else if (elem_bytes == 2)
    transpose_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)in, (uint16_t*)out, rows, cols);
  else if (elem_bytes == 4)
    transpose_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t*)in, (uint32_t*)out, rows, cols);
  else {
    set_error("transpose: element size must be 2 or 4 bytes (got %d)", elem_bytes);
    return MU_ERR_BAD_DTYPE;
  }
  return check_launch("transpose");
}

}  // namespace mu
