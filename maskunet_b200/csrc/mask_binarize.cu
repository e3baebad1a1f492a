// K2: mask binarisation + key compaction.
// Replaces /root/reference/code/ade20k/ade_semantic.py:179-181 (binary_mask > 0.5 -> 0 / -inf, expanded
// over queries).  The bias is per key and identical for every query, so instead of an additive [B,N,N]
// tensor we emit keep bits plus the compaction maps that let K/V be stored for kept keys only.
// HBM-bound integer work: 8 B read + ~8.1 B written per token.
#include "common.cuh"

namespace mu {

constexpr int kScanThreads = 1024;

// one CTA per sample; each iteration scans 1024 consecutive tokens (one per thread)
__global__ void __launch_bounds__(kScanThreads) mask_binarize_kernel(const int64_t* __restrict__ bits, int N,
                                                                     uint32_t* __restrict__ keep_bits,
                                                                     int32_t* __restrict__ n_keep,
                                                                     int32_t* __restrict__ keep_idx,
                                                                     int32_t* __restrict__ keep_rank) {
  __shared__ int warp_count[32];
  __shared__ int base_s;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t* row = bits + (size_t)b * N;
  const int words = (N + 31) / 32;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int n0 = 0; n0 < N; n0 += kScanThreads) {
    const int n = n0 + tid;
    // `binary_mask > 0.5` on an int64 tensor: true exactly when the integer is >= 1
    const bool keep = (n < N) && (row[n] > 0);
    const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) {
      warp_count[warp] = __popc(ballot);
      if (n < N) keep_bits[(size_t)b * words + (n >> 5)] = ballot;
    }
    __syncthreads();
    int v = (tid < 32) ? warp_count[tid] : 0;
    if (warp == 0) {  // inclusive scan of the 32 warp counts
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      warp_count[lane] = v;
    }
    __syncthreads();
    const int base = base_s;
    const int warp_excl = (warp == 0) ? 0 : warp_count[warp - 1];
    const int pos = base + warp_excl + __popc(ballot & ((1u << lane) - 1u));
    if (n < N) {
      keep_rank[(size_t)b * N + n] = keep ? pos : -1;
      if (keep) keep_idx[(size_t)b * N + pos] = n;
    }
    __syncthreads();
    if (tid == 0) base_s = base + warp_count[31];
    __syncthreads();
  }
  const int total = base_s;
  if (tid == 0) n_keep[b] = total;
  for (int i = total + tid; i < N; i += kScanThreads) keep_idx[(size_t)b * N + i] = -1;
}

int launch_mask_binarize(const int64_t* bits, int B, int N, uint32_t* keep_bits, int32_t* n_keep, int32_t* keep_idx,
                         int32_t* keep_rank, cudaStream_t s) {
  mask_binarize_kernel<<<B, kScanThreads, 0, s>>>(bits, N, keep_bits, n_keep, keep_idx, keep_rank);
  return check_launch("mask_binarize");
}

}  // namespace mu
