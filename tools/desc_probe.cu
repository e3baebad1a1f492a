// Hardware probe (not part of the library): does a tcgen05 shared-memory descriptor whose start address is
// shifted by whole 128-byte rows inside a SWIZZLE_128B tile address the rows TMA wrote there?
//   test K : K-major A operand, start += r * 128 B      (conv3x3 forward: the dx tap shift of an A tile with halo)
//   test MN: MN-major A operand, start += r * 128 B     (conv3x3 weight gradient: pixel shift along the K dimension)
// for base_offset in {0, r & 7}.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/desc_probe tools/desc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../maskunet_b200/csrc/sm100_ptx.cuh"
#include <cuda.h>

using namespace mu;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint64_t desc_bo(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t bo) {
  return make_smem_desc(addr, lbo, sbo) | ((uint64_t)(bo & 7) << 49);
}

// mode 0: K-major A;  mode 1: MN-major A (and MN-major B)
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmG,
                                                    const __grid_constant__ CUtensorMap tmB, float* out, int mode,
                                                    int r, int bo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 320 rows x 128 B = 40960
  uint8_t* sB = smem + 40960;         // 64 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 40960 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc<64>(slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bars, 40960 + 8192);
    tma_load_2d(sA, &tmG, bars, 0, 0);            // rows 0..255
    tma_load_2d(sA + 32768, &tmB, bars, 0, 256);  // placeholder replaced below (see host: tmB maps G too for rows)
    tma_load_2d(sB, &tmB, bars, 0, 320);          // B tile lives in rows 320..383 of the same global matrix
    mbar_wait(bars, 0);
    tc_fence_after();
    const uint32_t a = smem_u32(sA) + r * 128, b = smem_u32(sB);
    if (mode == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      for (int kk = 0; kk < 4; ++kk)
        umma_ss(tmem, desc_bo(a + kk * 32, 0, 1024, bo), make_smem_desc(b + kk * 32, 0, 1024), idesc, kk > 0);
    } else {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      for (int kk = 0; kk < 4; ++kk)
        umma_ss(tmem, desc_bo(a + kk * 2048, 16384, 1024, bo), make_smem_desc(b + kk * 2048, 8192, 1024), idesc, kk > 0);
    }
    umma_commit(bars + 1);
  }
  mbar_wait(bars + 1, 0);
  tc_fence_after();
  uint32_t v[32];
  const int row = warp * 32 + (int)lane_id();
  for (int c = 0; c < 2; ++c) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_wait_ld();
    for (int e = 0; e < 32; ++e) out[row * 64 + c * 32 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

int main() {
  const int R = 384;
  std::vector<__nv_bfloat16> h(R * 64);
  std::vector<float> hf(R * 64);
  srand(1);
  for (int i = 0; i < R * 64; ++i) {
    hf[i] = (float)(rand() % 9 - 4);
    h[i] = __float2bfloat16(hf[i]);
  }
  __nv_bfloat16* d;
  float* dout;
  cudaMalloc(&d, R * 64 * 2);
  cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(d, h.data(), R * 64 * 2, cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_encodeTiled>(p);
  CUtensorMap tmG, tmB;
  cuuint64_t dims[2] = {64, (cuuint64_t)R};
  cuuint64_t strides[1] = {128};
  cuuint32_t boxG[2] = {64, 256}, boxB[2] = {64, 64}, es[2] = {1, 1};
  if (enc(&tmG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, boxG, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ||
      enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) {
    printf("encode failed\n");
    return 1;
  }
  const int smem = 1024 + 40960 + 8192 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> o(128 * 64);
  const float* Bm = hf.data() + 320 * 64;  // [64][64]
  for (int mode = 0; mode < 2; ++mode) {
    const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 17};
    for (int r : shifts) {
      for (int v = 0; v < 2; ++v) {
        const int bo = v ? (r & 7) : 0;
        if (v && bo == 0) continue;
        probe_kernel<<<1, 128, smem>>>(tmG, tmB, dout, mode, r, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d r %d bo %d: CUDA error %s\n", mode, r, bo, cudaGetErrorString(e));
          return 2;
        }
        cudaMemcpy(o.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            if (mode == 0)
              for (int k = 0; k < 64; ++k) ref += hf[(r + m) * 64 + k] * Bm[n * 64 + k];
            else
              for (int k = 0; k < 64; ++k) ref += hf[(r + k + (m >= 64 ? 128 : 0)) * 64 + (m & 63)] * Bm[k * 64 + n];
            double err = fabs(ref - o[m * 64 + n]);
            if (err > maxerr) maxerr = err;
          }
        printf("mode %s shift %2d base_offset %d : max_abs_err %.1f %s\n", mode ? "MN" : "K ", r, bo, maxerr,
               maxerr == 0 ? "OK" : "MISMATCH");
      }
    }
  }
  return 0;
}
