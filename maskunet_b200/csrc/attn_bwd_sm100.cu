// K5: masked attention backward for the bf16 path.
// PLACEHOLDER until the tcgen05 kernel lands: forwards to the CUDA-core kernels (bf16 I/O, fp32 math).
#include "common.cuh"

namespace mu {

int launch_attn_bwd_sm100(const void* q, const void* kc, const void* vc, const int32_t* n_keep, const void* d_o,
                          const float* lse, const float* delta, void* dq, void* dkc, void* dvc, int B, int N, int NKP,
                          int C, cudaStream_t s) {
  return launch_attn_bwd_simt(q, kc, vc, n_keep, d_o, lse, delta, dq, dkc, dvc, B, N, NKP, C, MU_BF16, s);
}

}  // namespace mu
