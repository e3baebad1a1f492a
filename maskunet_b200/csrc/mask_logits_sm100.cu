// K13 (generalised mode of BASELINE.json configs[4]): per-query mask logits against pixel
// features, binarised on the fly into the attention bias bits.
//
//   logits[b, q, n] = sum_c qe[b, q, c] * feat[b, n, c]          (einsum 'bqc,bnc->bqn', bf16 in, fp32 accumulate)
//   keep[b, q, n]   = sigmoid(logits) > 0.5                       (bias 0.0 where kept, -inf elsewhere)
//
// The reference has NO such stage (its bias comes from torch.randint, ade_semantic.py:177-181):
// this is the north-star generalisation, used by the kernel sweep only, and its oracle is builder-written
// (oracle/query_attention_oracle.py, "parity unpinned by reference").
//
// The logits never reach HBM.  One CTA = one 128-query x 128-key tile: TMA loads both operands (C = 256: 4 blocks of
// [128 rows][64 channels], 128-byte swizzle), one thread issues the 16 tcgen05.mma of the tile into TMEM, four warps
// read their TMEM lane quadrant (lane = query row) and emit
//   bits   uint32 [B, Q, NKP/32]   bit n%32 of word n/32  <=> query q may attend key n      (forward: row = query)
//   bits_t uint32 [B, NKP, QP/32]  the same relation transposed (warp ballots)               (backward: row = key)
//   row_count int32 [B, Q]         kept keys per query row (a second small kernel opens rows that kept nothing)
// sigmoid(x) > 0.5 in fp32 is NOT x > 0: 1 / (1 + exp(-x)) rounds to exactly 0.5 for 0 < x <= 1.5 * 2^-24.  The rule
// implemented is x > 0x1.8p-24f, bit-identical to torch.sigmoid(x) > 0.5 for every float32 x (verified in the tests
// over the boundary values and on random logits).
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

namespace mu {

constexpr int kMLThreads = 192;     // warp 0: TMA, warp 1: TMEM + MMA issue, warps 2-5: epilogue (one lane quadrant each)
constexpr int kMLTile = 128;
__device__ __forceinline__ bool sigmoid_gt_half(float x) { return x > 0x1.8p-24f; }

template <int C>
struct MLCfg {
  static constexpr int kBlocks = C / 64;
  static constexpr int kOperandBytes = kMLTile * C * 2;
  static constexpr int kSmemBytes = 1024 + 2 * kOperandBytes + 128;
};

template <int C>
__global__ void __launch_bounds__(kMLThreads, 1)
mask_logits_sm100_kernel(const __grid_constant__ CUtensorMap tmap_qe, const __grid_constant__ CUtensorMap tmap_feat,
                         uint32_t* __restrict__ bits, uint32_t* __restrict__ bits_t, int32_t* __restrict__ row_count,
                         float* __restrict__ logits, int Q, int N, int NKP, int QP) {
  using Cfg = MLCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                   // query embeddings [128][C], 64-channel blocks
  uint8_t* sB = sA + Cfg::kOperandBytes;                // pixel features   [128][C]
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + Cfg::kOperandBytes);   // kBlocks
  uint64_t* done = full + Cfg::kBlocks;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int k0 = blockIdx.x * kMLTile, q0 = blockIdx.y * kMLTile, b = blockIdx.z;
  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kBlocks; ++i) mbar_init(full + i, 1);
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc<128>(tmem_slot);
    tmem_relinquish();
  }
  if (warp == 0 && lane_id() == 0) {
    tma_prefetch_desc(&tmap_qe);
    tma_prefetch_desc(&tmap_feat);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane_id() == 0) {
      for (int blk = 0; blk < Cfg::kBlocks; ++blk) {   // rows past Q / N are zero-filled: logit 0 -> not kept
        mbar_expect_tx(full + blk, 2 * kMLTile * 128);
        tma_load_3d(sA + blk * (kMLTile * 128), &tmap_qe, full + blk, blk * 64, q0, b);
        tma_load_3d(sB + blk * (kMLTile * 128), &tmap_feat, full + blk, blk * 64, k0, b);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(kMLTile, kMLTile, 0, 0);
    constexpr uint32_t hi = desc_hi_sbo(1024);
    const uint32_t a_lo = desc_lo(smem_u32(sA)), b_lo = desc_lo(smem_u32(sB));
    for (int blk = 0; blk < Cfg::kBlocks; ++blk) {
      mbar_wait(full + blk, 0);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t off = (blk * (kMLTile * 128) + kk * 32) >> 4;
        if (elect_one()) umma_ss_lo(tmem_base, a_lo + off, b_lo + off, hi, idesc, (blk | kk) ? 1u : 0u);
      }
    }
    if (elect_one()) umma_commit(done);
  } else {
    const int quad = warp & 3;                          // the TMEM lane quadrant this warp may read
    const int lane = (int)lane_id();
    const int r = quad * 32 + lane;                     // query row within the tile
    const bool row_ok = q0 + r < Q;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int wpr = NKP >> 5, wpq = QP >> 5;
    mbar_wait(done, 0);
    tc_fence_after();
    uint32_t words[4];
    uint32_t v[32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      tmem_ld32(lane_base + c * 32, v);
      tmem_wait_ld();
      uint32_t w = 0, tw = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const bool keep = row_ok && sigmoid_gt_half(__uint_as_float(v[i]));
        w |= (uint32_t)keep << i;
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);   // 32 query rows of key column c*32 + i
        if (lane == i) tw = bal;
      }
      words[c] = w;
      bits_t[((size_t)b * NKP + k0 + c * 32 + lane) * wpq + (q0 >> 5) + quad] = tw;
      if (logits != nullptr && row_ok) {                // test hook: the fp32 logits of this tile
        float* dst = logits + ((size_t)b * Q + q0 + r) * N + k0 + c * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (k0 + c * 32 + i < N) dst[i] = __uint_as_float(v[i]);
      }
    }
    if (row_ok) {
      *reinterpret_cast<uint4*>(bits + ((size_t)b * Q + q0 + r) * wpr + (k0 >> 5)) =
          make_uint4(words[0], words[1], words[2], words[3]);
      const int cnt = __popc(words[0]) + __popc(words[1]) + __popc(words[2]) + __popc(words[3]);
      if (cnt) atomicAdd(row_count + (size_t)b * Q + q0 + r, cnt);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// A query whose mask kept nothing attends every key (the Mask2Former rule: a fully masked row would make the softmax
// NaN).  One warp per (b, q) row.
__global__ void mask_rows_open_kernel(uint32_t* __restrict__ bits, uint32_t* __restrict__ bits_t,
                                      const int32_t* __restrict__ row_count, int B, int Q, int N, int NKP, int QP) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B * Q || row_count[row] != 0) return;
  const int b = row / Q, q = row % Q, lane = threadIdx.x & 31;
  const int wpr = NKP >> 5, wpq = QP >> 5;
  for (int w = lane; w < wpr; w += 32) {
    const int first = w * 32;
    bits[(size_t)row * wpr + w] = first + 32 <= N ? 0xffffffffu : (first < N ? (1u << (N - first)) - 1u : 0u);
  }
  for (int n = lane; n < N; n += 32) atomicOr(bits_t + ((size_t)b * NKP + n) * wpq + (q >> 5), 1u << (q & 31));
}

int launch_query_mask_bits_sm100(const void* qe, const void* feat, int B, int Q, int N, int C, uint32_t* bits,
                                 uint32_t* bits_t, int32_t* row_count, float* logits, cudaStream_t s) {
  MU_REQUIRE(C == 256, MU_ERR_BAD_SHAPE, "mu_query_mask_bits: channels must be 256 (got %d)", C);
  const int NKP = round_up(N, kMLTile), QP = round_up(Q, kMLTile);
  using Cfg = MLCfg<256>;
  CUtensorMap tq, tf;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tq, qe, C, Q, B, kMLTile))) return rc;
  if ((rc = make_tmap_bf16_3d(&tf, feat, C, N, B, kMLTile))) return rc;
  auto kern = mask_logits_sm100_kernel<256>;
  cudaError_t e = set_max_dynamic_smem_once(kern, Cfg::kSmemBytes);
  if (e != cudaSuccess) {
    set_error("mask_logits_sm100: cudaFuncSetAttribute(%d bytes): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
    return (int)e;
  }
  cudaMemsetAsync(row_count, 0, (size_t)B * Q * sizeof(int32_t), s);
  dim3 grid(NKP / kMLTile, QP / kMLTile, B);
  kern<<<grid, kMLThreads, Cfg::kSmemBytes, s>>>(tq, tf, bits, bits_t, row_count, logits, Q, N, NKP, QP);
  if ((rc = check_launch("mask_logits_sm100"))) return rc;
  const int rows = B * Q;
  mask_rows_open_kernel<<<(rows + 7) / 8, 256, 0, s>>>(bits, bits_t, row_count, B, Q, N, NKP, QP);
  return check_launch("mask_rows_open");
}

}  // namespace mu
