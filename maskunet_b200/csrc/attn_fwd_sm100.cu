// K3: masked attention forward on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM).
// Replaces /root/reference/code/ade20k/ade_semantic.py:174-186: scores = QK^T / sqrt(C) + mask,
// softmax, PV.  The per-key 0/-inf bias is realised by attending only to the compacted kept keys
// (kc / vc rows [0, n_keep[b])), so no N x N tensor and no in-loop mask exist; only the tail of the last
// key tile is masked.
//
// One CTA = one 128-query tile of one sample.  Warp roles:
//   warp 0   TMA producer: Q once, then K_j / V_j tiles through a ring of shared-memory slots
//   warp 1   allocates TMEM, single thread issues tcgen05.mma:  S_j = Q K_j^T,   O += P_j V_j
//   warp 4-7 softmax: thread <-> query row (TMEM lane); online max / exp2 / row sum in fp32,
//            P_j written to shared memory as bf16 in the UMMA K-major 128B-swizzled layout,
//            O rescaled in TMEM when the running max moves, final O / l and LSE written out.
// Shared-memory operand tiles are [rows][64 bf16] blocks as TMA SWIZZLE_128B writes them
// (see sm100_ptx.cuh for the descriptor conventions).
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

namespace mu {

constexpr int kBM = 128;                 // queries per CTA
constexpr int kFwdThreads = 256;         // warpgroup 0: TMA, MMA, 2 idle warps; warpgroup 1: softmax
constexpr float kLog2eF = 1.4426950408889634f;
// Measured on B200 (N = 16384, d = 64): 1 in 4 exponentials on the FMA pipe = 692 TFLOP/s, all on MUFU = 720: with the
// P tile in TMEM the softmax warps are issue-bound as much as MUFU-bound, so the offload stays off by default.
#ifndef MU_FWD_POLY_EVERY
#define MU_FWD_POLY_EVERY 0
#endif
constexpr int kPolyEvery = MU_FWD_POLY_EVERY;   // 0: all exponentials on MUFU; k: one in k on the FMA pipe
#ifndef MU_FWD_SPECULATE
#define MU_FWD_SPECULATE 1
#endif
// 1: the TMEM loads of S_{j+1} are issued before P_j is published (their latency under the store wait / fence /
// arrive).  Measured slower -- 764 against 813 TFLOP/s at N = 16384, d = 64: the wait for S_{j+1} delays the publish,
// and with it PV_j and the P-tile hand-back of the next tile -- so it stays off.
#ifndef MU_FWD_PREFETCH_S
#define MU_FWD_PREFETCH_S 0
#endif
constexpr bool kPrefetchS = MU_FWD_PREFETCH_S != 0;
#ifndef MU_FWD_EXP_PIPE
#define MU_FWD_EXP_PIPE 1
#endif
// Measured (tools/bench_kernels.py, 64 samples): d = 64, N = 16384: 813 -> 830 TFLOP/s; d = 128, N = 4096: 849 -> 959
// (groups of 4 / 8 / 16 columns: 828 / 820 / 830 and 949 / 949 / 959).
#ifndef MU_FWD_EXP_GROUP
#define MU_FWD_EXP_GROUP 16
#endif
constexpr bool kSpeculate = MU_FWD_SPECULATE != 0;   // exponentials before the tile maximum is known (see the softmax loop)
constexpr float kLazyLog2 = 8.f;         // rescale O only when the row maximum grew by more than 2^8

// -DMU_FWD_TRACE=1: CTA (0, 0) records clock64() at its pipeline events for key tiles [8, 40) (tools/fwd_trace.py).
#ifndef MU_FWD_TRACE
#define MU_FWD_TRACE 0
#endif
#if MU_FWD_TRACE
constexpr int kFTraceTiles = 32, kFTraceFirst = 8, kFTraceEvents = 16;
__device__ long long g_fwd_trace[kFTraceTiles * kFTraceEvents];
#define MU_FTRACE(ev, j)                                                                                   \
  do {                                                                                                     \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (j) >= kFTraceFirst &&            \
        (j) < kFTraceFirst + kFTraceTiles)                                                                 \
      g_fwd_trace[((j) - kFTraceFirst) * kFTraceEvents + (ev)] = clock64();                                \
  } while (0)
extern "C" int mu_debug_fwd_trace(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, g_fwd_trace, sizeof(long long) * (n < kFTraceTiles * kFTraceEvents ? n : kFTraceTiles * kFTraceEvents));
}
#else
#define MU_FTRACE(ev, j) do { } while (0)
#endif

// -DMU_FWD_CTALOG=1: every CTA records (SM id, clock64 at entry, clock64 at exit) for tools/fwd_ctalog.py (how well are the
// SMs filled over the launch, what is the SM clock inside it?); not part of the product build.
#ifndef MU_FWD_CTALOG
#define MU_FWD_CTALOG 0
#endif
#if MU_FWD_CTALOG
__device__ long long g_fwd_ctalog[3 * 65536];
extern "C" int mu_debug_fwd_ctalog(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, g_fwd_ctalog, sizeof(long long) * (n < 3 * 65536 ? n : 3 * 65536));
}
#endif

template <int D, int BN, int SBUFS, int SLOTS>
struct FwdCfg {
  static constexpr int kDBlocks = D / 64;                   // 64-column (128-byte) blocks per row
  static constexpr int kQBytes = kBM * D * 2;
  static constexpr int kKVBytes = BN * D * 2;               // one K or V tile = one ring slot
  static constexpr int kPBytes = kBM * BN * 2;
  static constexpr int kTmemS = 0;
  static constexpr int kTmemO = SBUFS * BN;
  static constexpr int kTmemP = SBUFS * BN + D;              // bf16 P tile, two keys per 32-bit column (A operand of PV)
  static constexpr int kTmemUsed = SBUFS * BN + D + BN / 2;
  static constexpr int kTmemCols = kTmemUsed <= 256 ? 256 : 512;
  static constexpr int kBarBytes = 8 * (1 + 2 * SLOTS + 2 * SBUFS + 2) + 8;
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kQBytes + SLOTS * kKVBytes + 256;
  static_assert(kBarBytes <= 256, "barrier block too small");
  static_assert(kTmemUsed <= 512, "TMEM overflow");
};

// QM = true is the generalised mode of the kernel sweep (no counterpart in the reference):
// a per-(query, key) bias given as bits (K13, mask_logits_sm100.cu: qbits [B / heads, N queries, wpr words], bit set
// <=> may attend) applied in registers before the row maximum, and every key of the tile list is "kept"
// (n_keep == number of keys).  The batch index counts (sample, head) pairs; heads of one sample share the mask.
template <int D, int BN, int SBUFS, int SLOTS, int MINB, bool QM>
__global__ void __launch_bounds__(kFwdThreads, MINB)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, const int32_t* __restrict__ n_keep,
                      __nv_bfloat16* __restrict__ o, float* __restrict__ lse, int N, float scale_log2,
                      const uint32_t* __restrict__ qbits, int heads, int wpr, int nk_all) {
  using Cfg = FwdCfg<D, BN, SBUFS, SLOTS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Cfg::kQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + SLOTS * Cfg::kKVBytes);
  uint64_t* q_full = bars;                  // 1
  uint64_t* kv_full = bars + 1;             // SLOTS
  uint64_t* kv_empty = kv_full + SLOTS;     // SLOTS
  uint64_t* s_full = kv_empty + SLOTS;      // SBUFS
  uint64_t* s_free = s_full + SBUFS;        // SBUFS
  uint64_t* p_full = s_free + SBUFS;        // 1
  uint64_t* o_done = p_full + 1;            // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

#if MU_FWD_CTALOG
  const long long cta_t0 = clock64();
#endif
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int b = blockIdx.y, q0 = blockIdx.x * kBM;
  const int nk = QM ? nk_all : n_keep[b];
  const int T = (nk + BN - 1) / BN;  // key tiles

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < SLOTS; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
    }
    for (int i = 0; i < SBUFS; ++i) {
      mbar_init(s_full + i, 1);
      mbar_init(s_free + i, 128);
    }
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  if (warp == 0 && lane_id() == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // 2 CTAs / SM start with 128 registers per thread; the data-movement warpgroup hands most of its share to the
    // softmax warpgroup, which keeps a whole score row in registers
    if (MINB > 1) reg_dealloc<32>();
  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane_id() == 0 && T > 0) {
      mbar_expect_tx(q_full, Cfg::kQBytes);
      for (int blk = 0; blk < Cfg::kDBlocks; ++blk) tma_load_3d(sQ + blk * (kBM * 128), &tmap_q, q_full, blk * 64, q0, b);
      for (int t = 0; t < 2 * T; ++t) {  // t = 2j: K_j, t = 2j+1: V_j
        const int slot = t % SLOTS, use = t / SLOTS;
        if (use > 0) mbar_wait_relaxed(kv_empty + slot, (use - 1) & 1);
        MU_FTRACE(11 + (t & 1), t >> 1);           // TMA: K_j (11) / V_j (12) issued
        mbar_expect_tx(kv_full + slot, Cfg::kKVBytes);
        const CUtensorMap* tm = (t & 1) ? &tmap_v : &tmap_k;
        uint8_t* dst = sKV + slot * Cfg::kKVBytes;
        for (int blk = 0; blk < Cfg::kDBlocks; ++blk)
          tma_load_3d(dst + blk * (BN * 128), tm, kv_full + slot, blk * 64, (t >> 1) * BN, b);
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer: the whole warp walks the loop, one elected
    // lane issues; descriptors are 32-bit low words + immediates (see umma_ss_lo)
    if (T > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kBM, BN, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(kBM, D, 0, 1);
      constexpr uint32_t hi = desc_hi_sbo(1024);
      const uint32_t q_lo = desc_lo(smem_u32(sQ)), kv_lo = desc_lo(smem_u32(sKV));
      auto issue_s = [&](int j) {
        const int t = 2 * j, slot = t % SLOTS, buf = j % SBUFS;
        mbar_wait(kv_full + slot, (t / SLOTS) & 1);
        MU_FTRACE(0, j);                           // MMA: K_j landed
        tc_fence_after();
        const uint32_t k_lo = kv_lo + slot * (Cfg::kKVBytes >> 4);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t offa = ((kk >> 2) * (kBM * 128) + (kk & 3) * 32) >> 4;
          const uint32_t offb = ((kk >> 2) * (BN * 128) + (kk & 3) * 32) >> 4;
          if (elect_one())
            umma_ss_lo(tmem_base + Cfg::kTmemS + buf * BN, q_lo + offa, k_lo + offb, hi, idesc_s, kk > 0 ? 1u : 0u);
        }
        if (elect_one()) {
          umma_commit(kv_empty + slot);
          umma_commit(s_full + buf);
        }
        MU_FTRACE(1, j);                           // MMA: S_j issued
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) {
          const int use = (j + 1) / SBUFS;
          if (use > 0) mbar_wait(s_free + (j + 1) % SBUFS, (use - 1) & 1);
          issue_s(j + 1);
        }
        const int t = 2 * j + 1, slot = t % SLOTS;
        mbar_wait(kv_full + slot, (t / SLOTS) & 1);
        MU_FTRACE(13, j);                          // MMA: V_j landed
        mbar_wait(p_full, j & 1);
        MU_FTRACE(2, j);                           // MMA: p_full(j)
        tc_fence_after();
        const uint32_t v_lo = desc_lo_lbo(smem_u32(sKV) + slot * Cfg::kKVBytes, BN * 128);
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk)     // A = P straight from TMEM (16 keys = 8 columns per step)
          if (elect_one())
            umma_ts_lo(tmem_base + Cfg::kTmemO, tmem_base + Cfg::kTmemP + kk * 8, v_lo + kk * 128, hi, idesc_o,
                       (j > 0 || kk > 0) ? 1u : 0u);
        if (elect_one()) {
          umma_commit(kv_empty + slot);
          umma_commit(o_done);
        }
        MU_FTRACE(3, j);                           // MMA: PV_j issued
      }
    }
  }
  } else {
    // ===================================================== softmax / correction / epilogue
    if (MINB > 1) reg_alloc<216>();
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may touch
    const int r = quad * 32 + (int)lane_id();       // query row within the tile
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool row_ok = q0 + r < N;
    __nv_bfloat16* orow = o + ((size_t)b * N + q0 + r) * D;
    if (T == 0) {  // every key masked: the reference yields NaN (softmax over all -inf)
      if (row_ok) {
        for (int c = 0; c < D; ++c) orow[c] = __float2bfloat16_rn(__int_as_float(0x7fc00000));
        lse[(size_t)b * N + q0 + r] = -INFINITY;
      }
    } else {
      float m = -INFINITY, l = 0.f;
      uint32_t v[BN / 32][32];                      // the whole S row of this thread, one TMEM pass per tile
      for (int j = 0; j < T; ++j) {
        const int buf = j % SBUFS;
        const uint32_t s_addr = lane_base + Cfg::kTmemS + buf * BN;
        const int limit = nk - j * BN;              // valid key columns in this tile (>= 1)
        if (!kPrefetchS || j == 0) {
          if (warp == 4) MU_FTRACE(4, j);           // softmax: waiting for S_j
          mbar_wait(s_full + buf, (j / SBUFS) & 1);
          if (warp == 4) MU_FTRACE(5, j);           // softmax: s_full(j)
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) tmem_ld32(s_addr + c * 32, v[c]);
        }
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(s_free + buf);                  // S lives in registers now: the next QK^T may overwrite TMEM
        if (warp == 4) MU_FTRACE(6, j);             // softmax: S_j in registers
        if (limit < BN) {                           // tail of the last key tile
#pragma unroll
          for (int c = 0; c < BN / 32; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= limit) v[c][i] = 0xff800000u;   // -inf
        }
        if (QM) {                                   // per-(query, key) bias: -inf where the bit is clear
          const uint32_t* rb = qbits + ((size_t)(b / heads) * N + (row_ok ? q0 + r : 0)) * wpr + j * (BN / 32);
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) {
            const uint32_t w = row_ok ? rb[c] : 0xffffffffu;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (!((w >> i) & 1u)) v[c][i] = 0xff800000u;
          }
        }
        // ---- row maximum: independent FMNMX3 chains (ALU pipe)
        auto row_max = [&]() -> float {
          float mx[BN / 32];
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) {
            mx[c] = max3(__uint_as_float(v[c][0]), __uint_as_float(v[c][1]), __uint_as_float(v[c][2]));
#pragma unroll
            for (int i = 3; i + 1 < 32; i += 2)
              mx[c] = max3(mx[c], __uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1]));
            mx[c] = fmaxf(mx[c], __uint_as_float(v[c][31]));
          }
          float r_ = mx[0];
#pragma unroll
          for (int c = 1; c < BN / 32; ++c) r_ = fmaxf(r_, mx[c]);
          return r_;
        };
        // ---- O *= alpha in TMEM (only when the exponent offset moves)
        auto rescale_o = [&](float a) {
          mbar_wait(o_done, (j - 1) & 1);                              // the previous PV has landed in TMEM
          tc_fence_after();
          uint32_t ov[32];
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            tmem_ld32(lane_base + Cfg::kTmemO + c * 32, ov);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * a);
            tmem_st32(lane_base + Cfg::kTmemO + c * 32, ov);
          }
          tmem_wait_st();
        };
        // ---- p = exp2((s - m) * scale * log2e) -> bf16 P tile in TMEM (it never touches shared memory: 32 keys = 16
        // packed columns); returns the row sum.  The P tile is single-buffered and PV_{j-1} (issued when P_{j-1} was
        // published, one exp phase ago, on a tensor pipe shared with the other CTA of this SM) still reads it ~1400
        // cycles into this tile (tools/fwd_trace.py): stores trail the arithmetic by one 32-key chunk so that the wait
        // comes after 64 exponentials and is normally already satisfied.
#if MU_FWD_EXP_PIPE
        // Software-pipelined in the SOURCE: the FFMA + MUFU.EX2 of group G + 1 (MU_FWD_EXP_GROUP columns) are written
        // before the additions and bf16 packs that consume group G.  With producer and consumer adjacent in the source,
        // ptxas kept them ~16 issue cycles apart against a MUFU latency of ~23+: a warp on its own ran at 11.4 cycles
        // per exponential instead of the 8 the MUFU pipe allows (profiles/r02_attn_fwd_timeline_1cta_2cta.txt), so the
        // pipe idled whenever the other CTA's warp was between tiles.
        auto exp_pass = [&](float m_off) -> float {
          const float mb = (QM && m_off == -INFINITY) ? 0.f : m_off * scale_log2;   // all -inf so far: p = exp2(-inf) = 0
          constexpr int GE = MU_FWD_EXP_GROUP, NG = BN / GE, GPC = 32 / GE;
          float sum[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t pk_prev[16], pk[16];
          float t[2][GE];
#pragma unroll
          for (int e = 0; e < GE; ++e) t[0][e] = fast_exp2(fmaf(__uint_as_float(v[0][e]), scale_log2, -mb));
#pragma unroll
          for (int G = 0; G < NG; ++G) {
            if (G + 1 < NG) {
#pragma unroll
              for (int e = 0; e < GE; ++e) {
                const int col = (G + 1) * GE + e;
                t[(G + 1) & 1][e] = fast_exp2(fmaf(__uint_as_float(v[col >> 5][col & 31]), scale_log2, -mb));
              }
            }
#pragma unroll
            for (int e = 0; e < GE / 2; ++e) {
              const float p0 = t[G & 1][2 * e], p1 = t[G & 1][2 * e + 1];
              sum[e & 3] += p0 + p1;
              pk[(G % GPC) * (GE / 2) + e] = pack_bf16(p0, p1);
            }
            if ((G % GPC) == GPC - 1) {            // a 32-column chunk of P is complete
              const int c = G / GPC;
              if (c == 1 && j > 0) {
                if (warp == 4) MU_FTRACE(8, j);         // softmax: first 64 exponentials done
                mbar_wait(o_done, (j - 1) & 1);
                if (warp == 4) MU_FTRACE(9, j);         // softmax: PV_{j-1} done, P tile free
                tc_fence_after();
              }
              if (c > 0) tmem_st16(lane_base + Cfg::kTmemP + (c - 1) * 16, pk_prev);
              if (c + 1 < BN / 32) {
#pragma unroll
                for (int i = 0; i < 16; ++i) pk_prev[i] = pk[i];
              } else {
                tmem_st16(lane_base + Cfg::kTmemP + c * 16, pk);
              }
            }
          }
          return (sum[0] + sum[1]) + (sum[2] + sum[3]);
        };
#else
        auto exp_pass = [&](float m_off) -> float {
          const float mb = (QM && m_off == -INFINITY) ? 0.f : m_off * scale_log2;   // all -inf so far: p = exp2(-inf) = 0
          float sum[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t pk_prev[16];
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float p0 = fast_exp2(fmaf(__uint_as_float(v[c][2 * i]), scale_log2, -mb));
              const float x1 = fmaf(__uint_as_float(v[c][2 * i + 1]), scale_log2, -mb);
              // every kPolyEvery-th exponential goes to the FMA pipe (MUFU.EX2 is the forward pass's busiest pipe)
              const float p1 = (kPolyEvery > 0 && (i % (kPolyEvery / 2 > 0 ? kPolyEvery / 2 : 1)) == 0) ? poly_exp2(x1)
                                                                                                    : fast_exp2(x1);
              sum[i & 3] += p0 + p1;
              pk[i] = pack_bf16(p0, p1);
            }
            if (c == 1 && j > 0) {
              if (warp == 4) MU_FTRACE(8, j);         // softmax: first 64 exponentials done
              mbar_wait(o_done, (j - 1) & 1);
              if (warp == 4) MU_FTRACE(9, j);         // softmax: PV_{j-1} done, P tile free
              tc_fence_after();
            }
            if (c > 0) tmem_st16(lane_base + Cfg::kTmemP + (c - 1) * 16, pk_prev);
            if (c + 1 < BN / 32) {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk_prev[i] = pk[i];
            } else {
              tmem_st16(lane_base + Cfg::kTmemP + c * 16, pk);
            }
          }
          return (sum[0] + sum[1]) + (sum[2] + sum[3]);
        };
#endif
        // ---- lazy rescale: the exponent offset m only follows the running maximum when it has moved by more than
        // 2^kLazyLog2 (P then stays below 2^kLazyLog2, harmless in bf16 / fp32), so after the first tiles the
        // softmax warps neither wait for the previous PV nor touch O in TMEM.  O / l and the LSE use the same m.
        float alpha = 1.f, tile_sum;
        if (kSpeculate && j > 0) {
          // Speculative order (MU_FWD_SPECULATE): the offset almost never moves after the first tiles, so the
          // exponentials start at once with the current m while the tile maximum is reduced on the ALU pipe beside
          // them -- the 400-cycle max phase leaves the serial chain of the tile (tools/fwd_trace.py).  If some row
          // did outgrow m by more than 2^kLazyLog2 the tile is redone with the new offset (its first pass may have
          // overflowed: every P value and the row sum are rewritten).
          tile_sum = exp_pass(m);
          const float m_new = fmaxf(m, row_max());
          const bool grow = (m_new - m) * scale_log2 > kLazyLog2;
          if (__any_sync(0xffffffffu, grow)) {
            alpha = (QM && m_new == -INFINITY) ? 1.f : fast_exp2((m - m_new) * scale_log2);
            m = m_new;
            tmem_wait_st();
            rescale_o(alpha);
            tile_sum = exp_pass(m);
          }
        } else {
          const float m_new = fmaxf(m, row_max());
          const bool grow = (m_new - m) * scale_log2 > kLazyLog2;       // true on the first tile (m = -inf)
          if (__any_sync(0xffffffffu, grow)) {
            alpha = (QM && m_new == -INFINITY) ? 1.f : fast_exp2((m - m_new) * scale_log2);   // row masked so far
            m = m_new;
            if (j > 0) rescale_o(alpha);
          }
          if (warp == 4) MU_FTRACE(7, j);             // softmax: row maximum (and rescale) done
          tile_sum = exp_pass(m);
        }
        if (warp == 4) MU_FTRACE(10, j);            // softmax: all exponentials done
        if (kPrefetchS && j + 1 < T) {
          // S_j is dead (exponentials and the maximum check are done): the TMEM loads of S_{j+1} -- computed while this
          // tile's exponentials ran -- are issued BEFORE the publish sequence of P_j, so their latency (and the MIO queue
          // they share with the MUFU stream) overlaps the store wait, the fence and the arrive instead of adding ~170
          // cycles without exponentials to every tile
          const int nbuf = (j + 1) % SBUFS;
          if (warp == 4) MU_FTRACE(4, j + 1);
          mbar_wait(s_full + nbuf, ((j + 1) / SBUFS) & 1);
          if (warp == 4) MU_FTRACE(5, j + 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) tmem_ld32(lane_base + Cfg::kTmemS + nbuf * BN + c * 32, v[c]);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(p_full);
        if (warp == 4) MU_FTRACE(14, j);            // softmax: P_j published
        l = l * alpha + tile_sum;
      }
      // ---- epilogue: O / l, LSE
      mbar_wait(o_done, (T - 1) & 1);
      tc_fence_after();
      const float inv = 1.f / l;
      uint32_t(&vo)[32] = v[0];
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        tmem_ld32(lane_base + Cfg::kTmemO + c * 32, vo);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(vo[8 * g + 0]) * inv, __uint_as_float(vo[8 * g + 1]) * inv);
            w.y = pack_bf16(__uint_as_float(vo[8 * g + 2]) * inv, __uint_as_float(vo[8 * g + 3]) * inv);
            w.z = pack_bf16(__uint_as_float(vo[8 * g + 4]) * inv, __uint_as_float(vo[8 * g + 5]) * inv);
            w.w = pack_bf16(__uint_as_float(vo[8 * g + 6]) * inv, __uint_as_float(vo[8 * g + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = w;
          }
        }
      }
      if (row_ok) lse[(size_t)b * N + q0 + r] = (m * scale_log2 + log2f(l)) / kLog2eF;
    }
  }
  tc_fence_before();
  __syncthreads();
#if MU_FWD_CTALOG
  if (threadIdx.x == 0) {
    const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x;
    if (lin < 65536u) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      g_fwd_ctalog[3 * lin] = smid;
      g_fwd_ctalog[3 * lin + 1] = cta_t0;
      g_fwd_ctalog[3 * lin + 2] = clock64();
    }
  }
#endif
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

template <int D, int BN, int SBUFS, int SLOTS, int MINB, bool QM = false>
static int run(const void* q, const void* kc, const void* vc, const int32_t* n_keep, void* o, float* lse, int B, int N,
               int NKP, cudaStream_t s, float scale = 0.f, const uint32_t* qbits = nullptr, int heads = 1,
               int nk_all = 0) {
  using Cfg = FwdCfg<D, BN, SBUFS, SLOTS>;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tq, q, D, N, B, kBM))) return rc;
  if ((rc = make_tmap_bf16_3d(&tk, kc, D, NKP, B, BN))) return rc;
  if ((rc = make_tmap_bf16_3d(&tv, vc, D, NKP, B, BN))) return rc;
  auto kern = attn_fwd_sm100_kernel<D, BN, SBUFS, SLOTS, MINB, QM>;
  cudaError_t e = set_max_dynamic_smem_once(kern, Cfg::kSmemBytes);
  if (e != cudaSuccess) {
    set_error("attn_fwd_sm100: cudaFuncSetAttribute(%d bytes): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
    return (int)e;
  }
  dim3 grid((N + kBM - 1) / kBM, B);
  const float scale_log2 = kLog2eF * (scale > 0.f ? scale : 1.f / sqrtf((float)D));
  kern<<<grid, kFwdThreads, Cfg::kSmemBytes, s>>>(tq, tk, tv, n_keep, (__nv_bfloat16*)o, lse, N, scale_log2, qbits, heads,
                                                  NKP / 32, nk_all);
  return check_launch("attn_fwd_sm100");
}

// Generalised mode: q [BH, Q, 64], k / v [BH, NKP, 64] (rows >= N zero), qbits [BH / heads, Q, NKP / 32]
int launch_query_attn_fwd_sm100(const void* q, const void* k, const void* v, const uint32_t* qbits, void* o, float* lse,
                                int BH, int heads, int Q, int N, int NKP, int D, float scale, cudaStream_t s) {
  MU_REQUIRE(D == 64, MU_ERR_BAD_SHAPE, "mu_query_attn_fwd: head dim must be 64 (pad 32 to 64), got %d", D);
  return run<64, 128, 1, 3, 2, true>(q, k, v, nullptr, o, lse, BH, Q, NKP, s, scale, qbits, heads, N);
}

int launch_attn_fwd_sm100(const void* q, const void* kc, const void* vc, const int32_t* n_keep, void* o, float* lse,
                          int B, int N, int NKP, int C, cudaStream_t s) {
  switch (C) {
    case 64:  // 2 CTAs / SM: TMEM 256 cols each, 96 KB of tiles each
      return run<64, 128, 1, 3, 2>(q, kc, vc, n_keep, o, lse, B, N, NKP, s);
    case 128:
      return run<128, 128, 2, 4, 1>(q, kc, vc, n_keep, o, lse, B, N, NKP, s);
    case 256:
      return run<256, 64, 2, 4, 1>(q, kc, vc, n_keep, o, lse, B, N, NKP, s);
    default:
      set_error("attn_fwd_sm100: channels must be 64, 128 or 256 (got %d)", C);
      return MU_ERR_BAD_SHAPE;
  }
}

}  // namespace mu
