// Throughput of MUFU.EX2 / F2FP.BF16 pack / FFMA per SM sub-partition on sm_100a, as a function of resident warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mufu_probe tools/probes/mufu_probe.cu
// One CTA on one SM, W warps per sub-partition (4 W warps in all); each warp runs ITER iterations of 16 independent ops.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, long long* cycles, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i) * 1e-3f;
  uint32_t pk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
      if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i])); }
    }
    if (MODE == 2 || MODE == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) v[2 * i] = __uint_as_float(pk[i]);
    }
    if (MODE == 5 || MODE == 6) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t w = __float_as_uint(v[i]);
        if (MODE == 5) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(w));
        if (MODE == 6) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(w));
        v[i] = __uint_as_float(w);
      }
    }
    if (MODE == 7) {
      // the forward softmax's instruction mix per pair of columns: 2 FFMA, 2 MUFU.EX2, 2 FADD, 1 F2FP pack, 1 ALU op
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = v[2 * i], b = v[2 * i + 1];
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(1.0001f), "f"(-0.001f));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b) : "f"(1.0001f), "f"(-0.001f));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
        float s2;
        asm volatile("add.f32 %0, %1, %2;" : "=f"(s2) : "f"(a), "f"(b));
        if (i & 1) asm volatile("add.f32 %0, %0, %1;" : "+f"(sum1) : "f"(s2));
        else asm volatile("add.f32 %0, %0, %1;" : "+f"(sum0) : "f"(s2));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(b), "f"(a));
        v[2 * i] = __uint_as_float((pk[i] & 0x3f800000u) ^ 0x3f000000u);       // keeps the chain loop-carried (1 ALU op)
        v[2 * i + 1] = sum0 * 0.f + 0.5f;
      }
      v[1] += sum1 * 0.f;
    }
    if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 4096;
  for (int w = 1; w <= 8; w *= 2) {
    probe<MODE><<<1, 128 * w>>>(out, cyc, iters);
    probe<MODE><<<1, 128 * w>>>(out, cyc, iters);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // per sub-partition: w warps x iters x ops warp-instructions in c cycles
    printf("%-28s warps/SMSP %d: %.2f cycles per warp-instruction per SMSP\n", name, w, (double)c / ((double)w * iters * ops_per_iter));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("MUFU.EX2", 16);
  run<1>("FFMA", 16);
  run<2>("F2FP.BF16.PACK_AB", 8);
  run<3>("MUFU.EX2 + FFMA (pairs)", 16);
  run<4>("16 MUFU.EX2 + 8 F2FP", 24);
  run<7>("softmax mix, per MUFU", 16);
  run<5>("ex2.approx.f16x2", 16);
  run<6>("ex2.approx.ftz.bf16x2", 16);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
