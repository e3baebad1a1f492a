// K5: masked attention backward on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM).
// Autograd of /root/reference/code/ade20k/ade_semantic.py:174-186 (implicit at :400).  P is recomputed from
// Q, Kc and the saved LSE, so no N x N tensor is ever stored (the reference keeps P, 1 GiB per image at N=16384).
//
// One CTA owns one 128-key tile j of one sample (compacted kept keys) and walks over all query tiles i:
//   MMA1  S^T  = K_j  Q_i^T            [128 keys x BM queries]   TMEM
//   MMA2  dP^T = V_j  dO_i^T           [128 x BM]                TMEM
//         (D = 64: one more K = 16 step each adds  ones . aug^T, so that the accumulators hold S - lse / scale and
//          dP - delta: the per-query statistics enter through the tensor core, see MU_BWD_FOLD_STATS)
//   threads (row = key): P^T = exp2(S^T c - lse) -> TMEM (bf16), dS^T = P^T (dP^T - delta) -> bf16 tile in shared memory
//   MMA3  dV_j += P^T  dO_i            [128 x DH]  TMEM, accumulates over i (A operand from TMEM)
//   MMA4  dK_j += dS^T Q_i             [128 x DH]  TMEM, accumulates over i
//   MMA5  dQ_i  = dS K_j  (D = 64: [BM queries x 64];  D >= 128: transposed, [128 channels x BM queries])
//         read back by a second warpgroup, staged in shared memory and added by the TMA unit (cp.reduce.async.bulk.tensor)
//         into the bf16 dq itself (D = 64) or an fp32 accumulator (D >= 128)
// D = 256 splits the accumulator width over blockIdx.z (DH = 128 channels each) because dK + dV alone
// would need 512 TMEM columns.  The 1/sqrt(C) factor of dS is applied when dK / dQ leave the chip.
#include <atomic>
#include <cstring>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

namespace mu {

constexpr int kBK = 128;            // keys per CTA
constexpr int kBwdThreads = 512;    // warps 0-3: TMA, MMA, 2 idle; 4-11: softmax (2 warpgroups); 12-15: dQ reduction
constexpr float kLog2eB = 1.4426950408889634f;
// Q_i / dO_i ring depth.  With 2 stages the loads of tile i+2 can only be issued when every MMA of tile i has
// completed, and the ~1250-cycle TMA round trip sat on the critical path of every tile (timeline of one CTA:
// tools/bwd_trace.py, profiles/r01_attn_bwd_timeline.txt): 3 stages = 686 -> 742 TFLOP/s at N = 16384, d = 64.
// 3 stages still fit: 231,680 (d = 64) and 230,656 (d = 128) of 232,448 bytes.
#ifndef MU_BWD_STAGES_64
#define MU_BWD_STAGES_64 3
#endif
#ifndef MU_BWD_STAGES_128
#define MU_BWD_STAGES_128 3
#endif
// 1: K_j / V_j as TMEM-resident A operands of the S^T / dP^T MMAs (d = 64).  Measured on B200 at N = 16384:
// 715 TFLOP/s against 741 with shared-memory A operands (parity-green either way), so it stays off: the TS-mode MMAs
// and the extra dq_free dependency of the shared P^T / dQ columns cost more than 32 KB per tile of operand reads save.
#ifndef MU_BWD_KV_TMEM
#define MU_BWD_KV_TMEM 0
#endif
#ifndef MU_BWD_DQ_RED
#define MU_BWD_DQ_RED 0     // 1: dQ tiles leave through red.global.add.v4.f32 from registers instead of the TMA reduce-add
#endif
// 1: the key-tile CTAs of a sample start their walk over the query tiles at staggered offsets (CTA j of J starts at
// tile floor(j T / J) and wraps), so that at any moment they reduce-add into DIFFERENT dQ tiles instead of all hitting
// the same 32 KB of the accumulator at once.  Measured on B200 (N = 16384, d = 64, 128 samples): 14.30 ms against
// 12.10 ms with every CTA of a sample on the same query tile (profiles/r02_attn_bwd_stagger_det_ab.log) -- the shared
// Q_i / dO_i tile is what the L2 serves best -- so the walk stays in lock step in free-running mode.  Deterministic
// mode (fixed accumulation order, below) turns the stagger on at run time: in lock step the ordered adds of a tile form a
// chain of J completions, staggered the CTA that is next in the order passed the tile two periods earlier.
#ifndef MU_BWD_STAGGER
#define MU_BWD_STAGGER 0
#endif
// Deterministic mode: 1 = publish a tile's semaphore one step late (after the NEXT reduce-add was issued), 0 = wait
// for the reduce-add to land and publish at once.  The successor in a tile's order arrives two tile periods after its
// predecessor, so the hand-over has to fit in that: publishing at once keeps the slack.
#ifndef MU_BWD_DET_LAG
#define MU_BWD_DET_LAG 0
#endif
// 1 (d = 64): the dQ partial tiles leave as bf16 and the TMA unit adds them straight into the bf16 dq output
// (cp.reduce.async.bulk.tensor .add on a bf16 tensor map): half the staging traffic on the shared-memory port (16 KB
// written + 16 KB read per tile instead of 32 + 32), half the L2 reduction traffic, no 1 GiB fp32 workspace, no clear of
// it, no convert kernel.  Cost: the up to N / 128 partial sums of a query tile are accumulated in bf16 -- a random walk
// of ~0.6 % relative error at 64 key tiles, beside the ~0.4 % the bf16 P / dS operands already carry (bar: 2e-2).
// 0 keeps the fp32 accumulator (and is what d >= 128 uses).
#ifndef MU_BWD_DQ_BF16
#define MU_BWD_DQ_BF16 1
#endif
// 1 (d = 64): the per-query statistics enter through the tensor core instead of 516 broadcast LDS.128 wavefronts per tile.
// A 16-wide "augmentation" slab per query, aug[q] = [-lse/scale split into three bf16 | 0 .. | delta split into three,
// negated | 0 ..], travels with Q_i / dO_i (two 2 KB TMA boxes per stage, canonical no-swizzle K-major core matrices);
// one extra K = 16 MMA step adds  ones_S . aug^T  to S^T  (S' = S - lse / scale, so P = exp2(c S')) and one adds
// ones_dP . aug^T  to dP^T  (dP' = dP - delta): the softmax warps load nothing from shared memory.  The shared memory
// comes from the bf16 dQ staging (16 KB less than the fp32 tile).
#ifndef MU_BWD_FOLD_STATS
#define MU_BWD_FOLD_STATS 1
#endif
#ifndef MU_BWD_PROBE
#define MU_BWD_PROBE 0      // 1 / 2: performance probes that skip work (wrong results), see DESIGN.md
#endif

// -DMU_BWD_TRACE=1: CTA (0, 0, 0) records clock64() at its pipeline events for query tiles [8, 40) into a global
// array read back by tools/bwd_trace.py (a timeline of one CTA; not part of the product build).
#ifndef MU_BWD_TRACE
#define MU_BWD_TRACE 0
#endif
#if MU_BWD_TRACE
constexpr int kTraceTiles = 32, kTraceFirst = 8, kTraceEvents = 24;
__device__ long long g_bwd_trace[kTraceTiles * kTraceEvents];
#define MU_TRACE(ev, i)                                                                                 \
  do {                                                                                                  \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0 &&             \
        (i) >= kTraceFirst && (i) < kTraceFirst + kTraceTiles)                                          \
      g_bwd_trace[((i) - kTraceFirst) * kTraceEvents + (ev)] = clock64();                               \
  } while (0)
#else
#define MU_TRACE(ev, i) do { } while (0)
#endif

template <int D, int BM, int DH, int STAGES, int PB>
struct BwdCfg {
  static constexpr bool kDQT = (D >= 128);               // dQ tile computed transposed (dQ^T = K_j^T dS)
  static constexpr int kDQM = 128;
  static constexpr int kKBytes = kBK * D * 2;            // K_j or V_j
  static constexpr int kQBytes = BM * D * 2;             // Q_i or dO_i
  static constexpr int kPBytes = kBK * BM * 2;           // P^T or dS^T
  // Optional (MU_BWD_KV_TMEM, off: measured slower), d = 64: K_j and V_j, the A operands of the S^T / dP^T MMAs of
  // EVERY query tile, live in TMEM for the whole CTA (32 KB less shared-memory operand traffic per tile).  The columns
  // come from letting the dQ accumulator share the P^T tile's columns: the tensor pipe runs dV_i (last reader of
  // P^T_i) before dQ_i, and the softmax warps wait for dq_free(i) before they store P^T_{i+1}.
  static constexpr bool kKVT = (MU_BWD_KV_TMEM != 0) && (D == 64) && !kDQT;
  static constexpr int kTmS = 0, kTmDP = BM, kTmDV = 2 * BM, kTmDK = 2 * BM + DH;
  static constexpr int kDQCols = kDQT ? BM : DH;
  static constexpr int kTmK = 2 * BM + 2 * DH, kTmV = kTmK + D / 2;          // packed bf16: two channels per column
  static constexpr int kTmDQ = kKVT ? kTmV + D / 2 : 2 * BM + 2 * DH;
  static constexpr int kTmP = kKVT ? kTmDQ : kTmDQ + kDQCols;   // bf16 P^T tile, two queries per 32-bit column (A of the dV MMA)
  static constexpr int kTmemUsed = kTmP + (kKVT ? (BM / 2 > kDQCols ? BM / 2 : kDQCols) : BM / 2);
  static_assert(kTmemUsed <= 512, "TMEM overflow");
  // PB = 2 double-buffers the P^T / dS^T tiles so that the exp / dS work of query tile i+1 overlaps the
  // dV / dK / dQ MMAs of tile i.  Dynamic shared memory is declared __align__(1024), no alignment slack.
  static constexpr bool kDQTma = (D <= 128);             // dQ tile leaves through a TMA reduce-add (smem permitting)
  static constexpr bool kDQBf16 = (MU_BWD_DQ_BF16 != 0) && (D == 64) && !kDQT && (MU_BWD_DQ_RED == 0);
  static constexpr int kDQStageBytes = kDQTma ? BM * DH * (kDQBf16 ? 2 : 4) : 0;   // dQ tile staged for the TMA reduce-add
  static constexpr bool kFold = (MU_BWD_FOLD_STATS != 0) && kDQBf16;     // statistics folded into the accumulators (d = 64)
  static constexpr int kAugBytes = kFold ? BM * 16 * 2 : 0;            // per stage: [2 K-chunks][BM queries][8 bf16]
  static constexpr int kOnesBytes = kFold ? 3 * kBK * 16 : 0;          // [zeros | ones | zeros] 16-byte chunk tiles
  static constexpr int kStatBytes = kFold ? STAGES * kAugBytes + kOnesBytes : 2 * 2 * BM * 4;   // or lse2, delta (x2 tiles)
  static constexpr int kSmemBytes =
      2 * kKBytes + 2 * STAGES * kQBytes + PB * kPBytes + kDQStageBytes + kStatBytes + 256;
  static_assert(kSmemBytes <= 232448, "shared memory overflow");
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Ordered accumulation (deterministic mode): sem counts the partial tiles already added into one dQ tile.
__device__ __forceinline__ void sem_wait_turn(const int32_t* sem, int turn) {
  const long long start = clock64();
  int v;
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(sem) : "memory");
    if (v == turn) break;
    __nanosleep(32);
#if MU_SPIN_CYCLES > 0
    if (clock64() - start > MU_SPIN_CYCLES) {
      printf("attn_bwd_sm100: dQ order semaphore timeout: block (%d,%d,%d) sees %d, waits for %d\n", blockIdx.x,
             blockIdx.y, blockIdx.z, v, turn);
      __trap();
    }
#endif
  } while (true);
  asm volatile("fence.proxy.async.global;" ::: "memory");   // the acquire above orders the TMA reduce-add issued next
}
__device__ __forceinline__ void sem_publish(int32_t* sem) {
  __threadfence();
  asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(sem) : "memory");
}

// QM = true: generalised mode of the kernel sweep (see attn_fwd_sm100.cu): a per-(query, key) bias as transposed bits
// (bits_t [B / heads, NKP keys, wpq words over queries], K13), every key kept (no compaction: n_keep == nk_all and key
// row r of tile j is token k0 + r), N queries against NT key tokens.
// DET = true: deterministic mode (staggered walk + order semaphores); a separate instantiation, so that the
// free-running kernel carries none of its address arithmetic (as a run-time switch it cost the softmax warps registers:
// 224 bytes of spill loads and 14 % of the kernel's speed).
template <int D, int BM, int DH, int STAGES, int PB, bool QM, bool DET = false>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                      const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                      const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ CUtensorMap tmap_aug,
                      const int32_t* __restrict__ n_keep,
                      const int32_t* __restrict__ keep_idx,
                      const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dq_acc,
                      __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int N, int NKP, float scale,
                      int NT, const uint32_t* __restrict__ bits_t, int heads, int wpq, int nk_all,
                      int32_t* __restrict__ sem) {
  using Cfg = BwdCfg<D, BM, DH, STAGES, PB>;
  constexpr bool DQT = Cfg::kDQT;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {   // the 128B-swizzled tiles need a 1024-byte aligned base
    if (threadIdx.x == 0) printf("attn_bwd_sm100: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sK = smem;
  uint8_t* sV = sK + Cfg::kKBytes;
  uint8_t* sQ = sV + Cfg::kKBytes;                       // STAGES x Q_i
  uint8_t* sDO = sQ + STAGES * Cfg::kQBytes;             // STAGES x dO_i
  uint8_t* sDS = sDO + STAGES * Cfg::kQBytes;             // PB x dS^T (P^T lives in TMEM: only the dV MMA reads it)
  uint8_t* sDQ = sDS + PB * Cfg::kPBytes;                 // fp32 dQ staging (D = 64 only)
  float* sLse = reinterpret_cast<float*>(sDQ + Cfg::kDQStageBytes);   // [2][BM], already * log2e (not kFold)
  float* sDelta = sLse + 2 * BM;
  uint8_t* sAug = sDQ + Cfg::kDQStageBytes;                            // kFold: STAGES x [2][BM][16 B], then the ones tiles
  uint8_t* sOnes = sAug + STAGES * Cfg::kAugBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDQ + Cfg::kDQStageBytes + Cfg::kStatBytes);
  uint64_t* kv_full = bars;                 // 1
  uint64_t* qdo_full = bars + 1;            // STAGES
  uint64_t* qdo_empty = qdo_full + STAGES;  // STAGES
  uint64_t* s_full = qdo_empty + STAGES;    // 1
  uint64_t* pds_full = s_full + 1;          // 1 (256 arrivals)
  uint64_t* pds_free = pds_full + 1;        // PB
  uint64_t* dq_full = pds_free + PB;        // 1
  uint64_t* dq_free = dq_full + 1;          // 1 (128 arrivals)
  uint64_t* sdp_free = dq_free + 1;         // 1 (256 arrivals): S^T / dP^T of this tile are in registers
  uint64_t* p_free = sdp_free + 1;          // 1: the dV MMAs have consumed P^T
  uint64_t* kv_tm = p_free + 1;             // 1 (256 arrivals): K_j / V_j copied into TMEM (kKVT)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_tm + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int b = blockIdx.y, k0 = blockIdx.x * kBK, half = blockIdx.z;
  const int nk = QM ? nk_all : n_keep[b];
  if (!QM && nk == 0) {
    // no key kept in this sample (probability 2^-N for the reference's masks, but masks can be injected): nobody
    // scatters a row, so the first CTA clears this sample's dk / dv rows (its half of the channels)
    if (blockIdx.x == 0) {
      const int vec_per_row = DH / 8;
      for (long idx = threadIdx.x; idx < (long)NT * vec_per_row; idx += kBwdThreads) {
        const long row = idx / vec_per_row;
        const int v = (int)(idx % vec_per_row);
        const size_t off = ((size_t)b * NT + row) * D + half * DH + v * 8;
        *reinterpret_cast<uint4*>(dk + off) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dv + off) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    return;
  }
  if (k0 >= nk) return;                      // whole CTA: nothing kept in this tile
  const int T = (N + BM - 1) / BM;           // query tiles
  // Walk order over the query tiles.  J = key tiles of this sample that hold kept keys (this CTA is tile j < J).
  const int J = QM ? (nk + kBK - 1) / kBK : (nk + kBK - 1) / kBK;
  constexpr bool kMayStagger = ((MU_BWD_STAGGER != 0) || DET) && !QM;
  const bool stagger = kMayStagger && J <= T;
  const int i0 = stagger ? (int)(((long)blockIdx.x * T) / J) : 0;
  auto tile_of = [&](int step) {             // query tile processed at `step` (0 <= step < T)
    if (!kMayStagger) return step;
    const int t = step + i0;
    return t >= T ? t - T : t;
  };
  // Deterministic mode (sem != nullptr): the J partial dQ tiles of query tile qt are added in a FIXED order -- the
  // order in which the staggered walks reach qt: CTA j is number turn_of(qt) in it.  sem[b][qt] counts finished adds.
  auto turn_of = [&](int qt) {
    if (!stagger) return (int)blockIdx.x;
    int c = (int)((((long)qt + 1) * J + T - 1) / T);           // CTAs whose start tile is <= qt
    if (c > J) c = J;
    int t = c - 1 - (int)blockIdx.x;
    return t < 0 ? t + J : t;
  };

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(qdo_full + i, 1);
      mbar_init(qdo_empty + i, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(pds_full, 256);
    for (int i = 0; i < PB; ++i) mbar_init(pds_free + i, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 128);
    mbar_init(sdp_free, 256);
    mbar_init(p_free, 1);
    mbar_init(kv_tm, 256);
    mbar_fence_init();
  }
  if (Cfg::kFold) {
    // constant A operands of the two augmentation MMAs: three [128 rows][8 bf16] chunk tiles  zeros | ones | zeros
    // (ones = 1, 1, 1, 0, 0, 0, 0, 0 in every row).  ones_S = chunks (ones, zeros) starts at tile 1, ones_dP = (zeros, ones)
    // at tile 0; both with a K-chunk stride (LBO) of one tile.
    for (int idx = threadIdx.x; idx < 3 * kBK; idx += kBwdThreads) {
      const bool ones = idx >= kBK && idx < 2 * kBK;
      st_shared_v4(smem_u32(sOnes) + idx * 16, ones ? 0x3F803F80u : 0u, ones ? 0x00003F80u : 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  if (warp == 0 && lane_id() == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // 512 threads start with 128 registers each; data-movement and reduction warpgroups hand registers to the two
  // softmax warpgroups, which hold a whole S^T / dP^T half-row (the TMEM buffers are released right after the
  // loads, so the next tile's S^T / dP^T MMAs run underneath the exp / dS arithmetic of this one).
  if (warp < 4) {
  reg_dealloc<56>();
  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane_id() == 0) {
      mbar_expect_tx(kv_full, 2 * Cfg::kKBytes);
      for (int blk = 0; blk < D / 64; ++blk) {
        tma_load_3d(sK + blk * (kBK * 128), &tmap_k, kv_full, blk * 64, k0, b);
        tma_load_3d(sV + blk * (kBK * 128), &tmap_v, kv_full, blk * 64, k0, b);
      }
      for (int i = 0; i < T; ++i) {
        const int st = i % STAGES, use = i / STAGES;
        if (use > 0) mbar_wait_relaxed(qdo_empty + st, (use - 1) & 1);
        MU_TRACE(0, i);                          // TMA: stage free, loads of tile i issued
        mbar_expect_tx(qdo_full + st, 2 * Cfg::kQBytes + Cfg::kAugBytes);
        if (Cfg::kFold) {          // the statistics slab of this query tile: K-chunk 0 (lse) and K-chunk 1 (delta)
          tma_load_3d(sAug + st * Cfg::kAugBytes, &tmap_aug, qdo_full + st, 0, tile_of(i) * BM, b);
          tma_load_3d(sAug + st * Cfg::kAugBytes + BM * 16, &tmap_aug, qdo_full + st, 8, tile_of(i) * BM, b);
        }
        for (int blk = 0; blk < D / 64; ++blk) {
          tma_load_3d(sQ + st * Cfg::kQBytes + blk * (BM * 128), &tmap_q, qdo_full + st, blk * 64, tile_of(i) * BM, b);
          tma_load_3d(sDO + st * Cfg::kQBytes + blk * (BM * 128), &tmap_do, qdo_full + st, blk * 64, tile_of(i) * BM, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer: the whole warp walks the loop, one elected lane
    // issues; every descriptor is a loop-invariant 32-bit low word plus an immediate (see umma_ss_lo): the single
    // issuing thread feeds 32 MMAs per query tile
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(kBK, BM, 0, 0);     // S^T, dP^T
      constexpr uint32_t idesc_acc = make_idesc_bf16(kBK, DH, 0, 1);   // dV, dK: A K-major, B MN-major
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, DQT ? BM : DH, 1, 1);
      constexpr uint32_t hi = desc_hi_sbo(1024);
      const uint32_t k_lo = desc_lo(smem_u32(sK)), v_lo = desc_lo(smem_u32(sV));
      const uint32_t q_lo0 = desc_lo(smem_u32(sQ)), do_lo0 = desc_lo(smem_u32(sDO)), ds_lo0 = desc_lo(smem_u32(sDS));
      constexpr uint32_t kLboQ = (((uint32_t)(BM * 128) >> 4) & 0x3FFFu) << 16;     // MN-major tiles of BM rows
      constexpr uint32_t kLboK = (((uint32_t)(kBK * 128) >> 4) & 0x3FFFu) << 16;    // MN-major tiles of 128 rows
      auto issue_s_dp = [&](int i) {
        const int st = i % STAGES;
        mbar_wait(qdo_full + st, (i / STAGES) & 1);
        MU_TRACE(1, i);                          // MMA: Q_i / dO_i landed
        tc_fence_after();
        const uint32_t qa = q_lo0 + st * (Cfg::kQBytes >> 4), da = do_lo0 + st * (Cfg::kQBytes >> 4);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t offa = ((kk >> 2) * (kBK * 128) + (kk & 3) * 32) >> 4, offb = ((kk >> 2) * (BM * 128) + (kk & 3) * 32) >> 4;
          if (Cfg::kKVT) {
            if (elect_one()) umma_ts_lo(tmem_base + Cfg::kTmS, tmem_base + Cfg::kTmK + kk * 8, qa + offb, hi, idesc_s, kk > 0 ? 1u : 0u);
          } else {
            if (elect_one()) umma_ss_lo(tmem_base + Cfg::kTmS, k_lo + offa, qa + offb, hi, idesc_s, kk > 0 ? 1u : 0u);
          }
        }
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t offa = ((kk >> 2) * (kBK * 128) + (kk & 3) * 32) >> 4, offb = ((kk >> 2) * (BM * 128) + (kk & 3) * 32) >> 4;
          if (Cfg::kKVT) {
            if (elect_one()) umma_ts_lo(tmem_base + Cfg::kTmDP, tmem_base + Cfg::kTmV + kk * 8, da + offb, hi, idesc_s, kk > 0 ? 1u : 0u);
          } else {
            if (elect_one()) umma_ss_lo(tmem_base + Cfg::kTmDP, v_lo + offa, da + offb, hi, idesc_s, kk > 0 ? 1u : 0u);
          }
        }
        if (Cfg::kFold) {
          // canonical no-swizzle K-major operands: 8-row core matrices 128 B apart (SBO), K-chunks one tile apart (LBO)
          const uint64_t nosw = ((uint64_t)((128u >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
          const uint64_t aug_d = nosw | (uint64_t)(((uint32_t)(BM * 16) >> 4) & 0x3FFFu) << 16 |
                                 (uint64_t)desc_lo(smem_u32(sAug) + st * Cfg::kAugBytes);
          const uint64_t ones_lbo = (uint64_t)(((uint32_t)(kBK * 16) >> 4) & 0x3FFFu) << 16;
          const uint64_t ones_s = nosw | ones_lbo | (uint64_t)desc_lo(smem_u32(sOnes) + kBK * 16);
          const uint64_t ones_dp = nosw | ones_lbo | (uint64_t)desc_lo(smem_u32(sOnes));
          if (elect_one()) {
            umma_ss(tmem_base + Cfg::kTmS, ones_s, aug_d, idesc_s, 1u);
            umma_ss(tmem_base + Cfg::kTmDP, ones_dp, aug_d, idesc_s, 1u);
          }
        }
        if (elect_one()) umma_commit(s_full);
      };
      auto issue_acc = [&](int i) {
        const int st = i % STAGES;
        const uint32_t ds_lo = ds_lo0 + (i % PB) * (Cfg::kPBytes >> 4);
        const uint32_t qa = q_lo0 + st * (Cfg::kQBytes >> 4) + ((half * 2 * (BM * 128)) >> 4);
        const uint32_t da = do_lo0 + st * (Cfg::kQBytes >> 4) + ((half * 2 * (BM * 128)) >> 4);
        const uint32_t acc = i > 0 ? 1u : 0u;
#pragma unroll
        for (int kk = 0; kk < BM / 16; ++kk)     // dV += P^T dO_i, A = P^T from TMEM (16 queries = 8 columns)
          if (elect_one())
            umma_ts_lo(tmem_base + Cfg::kTmDV, tmem_base + Cfg::kTmP + kk * 8, (da + kk * 128) | kLboQ, hi, idesc_acc,
                       kk > 0 ? 1u : acc);
        if (elect_one()) umma_commit(p_free);
#pragma unroll
        for (int kk = 0; kk < BM / 16; ++kk) {   // dK += dS^T Q_i
          const uint32_t offa = ((kk >> 2) * (kBK * 128) + (kk & 3) * 32) >> 4;
          if (MU_BWD_PROBE != 3 && elect_one())
            umma_ss_lo(tmem_base + Cfg::kTmDK, ds_lo + offa, (qa + kk * 128) | kLboQ, hi, idesc_acc, kk > 0 ? 1u : acc);
        }
        MU_TRACE(4, i);                          // MMA: dV / dK issued
        if (i > 0) {
          mbar_wait(dq_free, (i - 1) & 1);
          tc_fence_after();
        }
        MU_TRACE(5, i);                          // MMA: dQ accumulator free
#pragma unroll
        for (int kk = 0; kk < kBK / 16; ++kk) {  // dQ tile: contraction over the 128 keys of this CTA
          uint32_t a, bd;
          if (DQT) {  // dQ^T [channels x queries] = K_j^T (MN-major A) . dS (MN-major B)
            a = (k_lo + ((half * 2 * (kBK * 128)) >> 4) + kk * 128) | kLboK;
            bd = (ds_lo + kk * 128) | kLboK;
          } else {    // dQ [queries x channels] = dS (MN-major A over queries) . K_j (MN-major B)
            a = (ds_lo + kk * 128) | kLboK;
            bd = (k_lo + kk * 128) | kLboK;
          }
          if (MU_BWD_PROBE != 3 && MU_BWD_PROBE != 4 && elect_one())
            umma_ss_lo(tmem_base + Cfg::kTmDQ, a, bd, hi, idesc_dq, kk > 0 ? 1u : 0u);
        }
        if (elect_one()) {
          umma_commit(qdo_empty + st);
          umma_commit(pds_free + (i % PB));
          umma_commit(dq_full);
        }
      };
      mbar_wait(kv_full, 0);
      if (Cfg::kKVT) {
        mbar_wait(kv_tm, 0);                 // the softmax warps have copied K_j / V_j into TMEM
        tc_fence_after();
      }
      issue_s_dp(0);
      for (int i = 0; i < T; ++i) {
        if (STAGES >= 2) {
          mbar_wait(sdp_free, i & 1);        // S^T / dP^T of tile i drained into registers
          MU_TRACE(2, i);                    // MMA: sdp_free(i) seen
          tc_fence_after();
          if (i + 1 < T) issue_s_dp(i + 1);  // runs underneath the softmax arithmetic of tile i
          mbar_wait(pds_full, i & 1);        // P^T / dS^T of tile i are in shared memory
          MU_TRACE(3, i);                    // MMA: pds_full(i) seen
          tc_fence_after();
          issue_acc(i);
          MU_TRACE(6, i);                    // MMA: acc(i) issued
        } else {                             // single Q/dO stage: tile i+1 can only load after acc(i) released it
          mbar_wait(pds_full, i & 1);
          tc_fence_after();
          issue_acc(i);
          if (i + 1 < T) issue_s_dp(i + 1);
        }
      }
    }
  }
  } else if (warp < 12) {
    reg_alloc<184>();
    // ===================================================== softmax-backward warps: thread <-> key row.
    // Two warpgroups split the query columns of every tile (no row reduction is needed in backward, so the
    // halves are independent); two warps per SM sub-partition hide each other's MUFU / FMA latencies.
    const int quad = warp & 3;
    const int hcol = (warp - 4) >> 2;                    // 0: first half of the columns, 1: second half
    const int r = quad * 32 + (int)lane_id();
    const int t = threadIdx.x - 128;                     // 0..255 over both warpgroups
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool key_ok = k0 + r < nk;
    const float scale_log2 = scale * kLog2eB;
    const float* lse_b = lse + (size_t)b * N;
    const float* delta_b = delta + (size_t)b * N;
    const uint32_t lse_addr = smem_u32(sLse), delta_addr = smem_u32(sDelta);
    const uint32_t ds_base = smem_u32(sDS);
    auto fetch = [&](int i, float& l2, float& dl) {
      const int qi = tile_of(i) * BM + t;
      const bool ok = (t < BM) && (qi < N);
      l2 = ok ? lse_b[qi] : INFINITY;                    // +inf -> p = 0 for rows past N (scaled by log2e when staged)
      dl = ok ? delta_b[qi] : 0.f;
    };
    float nl2 = 0.f, ndl = 0.f;
    if (!Cfg::kFold) fetch(0, nl2, ndl);
    if (Cfg::kKVT) {
      // thread <-> key row r: its 64 channels (128 bytes, 8 swizzled 16-byte chunks) become 32 packed TMEM columns;
      // the first warpgroup copies K_j, the second V_j
      mbar_wait(kv_full, 0);
      const uint32_t src = smem_u32(hcol == 0 ? sK : sV) + r * 128;
      uint32_t w[32];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const float4 f = ld_shared_v4f(src + ((uint32_t)(ch ^ (r & 7))) * 16);
        w[4 * ch] = __float_as_uint(f.x);
        w[4 * ch + 1] = __float_as_uint(f.y);
        w[4 * ch + 2] = __float_as_uint(f.z);
        w[4 * ch + 3] = __float_as_uint(f.w);
      }
      tmem_st32(lane_base + (hcol == 0 ? Cfg::kTmK : Cfg::kTmV), w);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(kv_tm);
    }
    constexpr int kChunksPerThread = BM / 64;            // 32-column chunks per thread
    uint32_t s[kChunksPerThread][32], dp[kChunksPerThread][32];
    const bool tile_partial = k0 + kBK > nk;
    for (int i = 0; i < T; ++i) {
      const uint32_t my_lse = lse_addr + (i & 1) * BM * 4, my_delta = delta_addr + (i & 1) * BM * 4;
      if (!Cfg::kFold) {
        if (t < BM) {
          st_shared_f32(my_lse + t * 4, nl2 * kLog2eB);
          st_shared_f32(my_delta + t * 4, ndl);
        }
        if (i + 1 < T) fetch(i + 1, nl2, ndl);
        named_bar_sync(1, 256);
      }
      if (warp == 4) MU_TRACE(7, i);                     // softmax: waiting for S^T / dP^T
      mbar_wait(s_full, i & 1);
      if (warp == 4) MU_TRACE(8, i);                     // softmax: s_full(i)
      if (warp == 8) MU_TRACE(19, i);                    // second warpgroup: s_full(i)
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < kChunksPerThread; ++cc) {
        const int c = hcol * kChunksPerThread + cc;      // 32-column chunk index within the tile
        tmem_ld32(lane_base + Cfg::kTmS + c * 32, s[cc]);
        tmem_ld32(lane_base + Cfg::kTmDP + c * 32, dp[cc]);
      }
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(sdp_free);
      if (warp == 4) MU_TRACE(9, i);                     // softmax: tile in registers
      if (tile_partial && !key_ok) {                     // rows past n_keep in the last key tile: p = exp2(-inf) = 0
#pragma unroll
        for (int cc = 0; cc < kChunksPerThread; ++cc)
#pragma unroll
          for (int e = 0; e < 32; ++e) s[cc][e] = 0xff800000u;
      }
      if (QM) {                                          // per-(query, key) bias: p = exp2(-inf) = 0 where the bit is clear
        const uint32_t* kb = bits_t + ((size_t)(b / heads) * NKP + k0 + r) * wpq + tile_of(i) * (BM / 32) + hcol * kChunksPerThread;
#pragma unroll
        for (int cc = 0; cc < kChunksPerThread; ++cc) {
          const uint32_t w = kb[cc];
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (!((w >> e) & 1u)) s[cc][e] = 0xff800000u;
        }
      }
      // the MMAs that last read this P^T / dS^T buffer (tile i - PB) must have completed
      if (i >= PB) mbar_wait(pds_free + (i % PB), ((i / PB) - 1) & 1);
      if (warp == 4) MU_TRACE(10, i);                    // softmax: output buffers free
      if (warp == 8) MU_TRACE(20, i);                    // second warpgroup: output buffers free
      const uint32_t ds_addr = ds_base + (i % PB) * Cfg::kPBytes;
#pragma unroll
      for (int cc = 0; cc < kChunksPerThread; ++cc) {
        const int c = hcol * kChunksPerThread + cc;
        uint32_t pk[16], dk[16];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float p0, p1, p2, p3, d0, d1, d2, d3;
          if (Cfg::kFold) {
            // S' = S - lse / scale and dP' = dP - delta came out of the tensor core
            p0 = fast_exp2(__uint_as_float(s[cc][4 * e + 0]) * scale_log2);
            p1 = fast_exp2(__uint_as_float(s[cc][4 * e + 1]) * scale_log2);
            p2 = fast_exp2(__uint_as_float(s[cc][4 * e + 2]) * scale_log2);
            p3 = fast_exp2(__uint_as_float(s[cc][4 * e + 3]) * scale_log2);
            d0 = p0 * __uint_as_float(dp[cc][4 * e + 0]);
            d1 = p1 * __uint_as_float(dp[cc][4 * e + 1]);
            d2 = p2 * __uint_as_float(dp[cc][4 * e + 2]);
            d3 = p3 * __uint_as_float(dp[cc][4 * e + 3]);
          } else {
#if MU_BWD_PROBE == 5    // no per-column statistics loads (wrong results): what do the broadcast LDS.128 cost?
            const float4 l2 = make_float4(scale, scale, scale, scale), dl = l2;
#else
            const float4 l2 = ld_shared_v4f(my_lse + (c * 32 + 4 * e) * 4);
            const float4 dl = ld_shared_v4f(my_delta + (c * 32 + 4 * e) * 4);
#endif
            p0 = fast_exp2(fmaf(__uint_as_float(s[cc][4 * e + 0]), scale_log2, -l2.x));
            p1 = fast_exp2(fmaf(__uint_as_float(s[cc][4 * e + 1]), scale_log2, -l2.y));
            p2 = fast_exp2(fmaf(__uint_as_float(s[cc][4 * e + 2]), scale_log2, -l2.z));
            p3 = fast_exp2(fmaf(__uint_as_float(s[cc][4 * e + 3]), scale_log2, -l2.w));
            d0 = p0 * (__uint_as_float(dp[cc][4 * e + 0]) - dl.x);
            d1 = p1 * (__uint_as_float(dp[cc][4 * e + 1]) - dl.y);
            d2 = p2 * (__uint_as_float(dp[cc][4 * e + 2]) - dl.z);
            d3 = p3 * (__uint_as_float(dp[cc][4 * e + 3]) - dl.w);
          }
          pk[2 * e] = pack_bf16(p0, p1);
          pk[2 * e + 1] = pack_bf16(p2, p3);
          dk[2 * e] = pack_bf16(d0, d1);
          dk[2 * e + 1] = pack_bf16(d2, d3);
        }
        const uint32_t row_off = (c >> 1) * (kBK * 128) + r * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const uint32_t chunk = (((c & 1) * 4 + ch) ^ (r & 7)) * 16;
#if MU_BWD_PROBE == 2
          if (scale == 12345.f)
#endif
          st_shared_v4(ds_addr + row_off + chunk, dk[4 * ch], dk[4 * ch + 1], dk[4 * ch + 2], dk[4 * ch + 3]);
        }
        // (Writing the exponentials of the next 16 columns ahead of the dS arithmetic of the current ones in the source, which
        // pays in the forward kernel, was measured slower here: 780 against 810 TFLOP/s at N = 16384 -- this kernel is bound
        // by the shared-memory port, not by MUFU latency.)
        // P^T is single-buffered in TMEM: the dV MMAs of tile i-1 must be done with it.  They run right after S^T /
        // dP^T of this tile on the tensor pipe, so waiting here -- after the first chunk's arithmetic -- instead of
        // before it takes ~400 cycles of stall off the critical chain of every tile (tools/bwd_trace.py).
        if (cc == 0 && i > 0) {
          mbar_wait(p_free, (i - 1) & 1);
          if (Cfg::kKVT) mbar_wait(dq_free, (i - 1) & 1);   // dQ_{i-1} shares these columns and has been read out
          tc_fence_after();
        }
        tmem_st16(lane_base + Cfg::kTmP + c * 16, pk);   // 32 queries = 16 packed columns of the TMEM P^T tile
        if (warp == 4) MU_TRACE(15 + cc, i);             // softmax: chunk cc computed and stored
      }
      tmem_wait_st();
      if (warp == 4) MU_TRACE(17, i);                    // softmax: tcgen05.wait::st
      tc_fence_before();
      fence_proxy_async_smem();
      if (warp == 4) MU_TRACE(18, i);                    // softmax: proxy fence
      mbar_arrive(pds_full);
      if (warp == 4) MU_TRACE(11, i);                    // softmax: P^T / dS^T published
      if (warp == 8) MU_TRACE(21, i);                    // second warpgroup: published
    }
    // ---- epilogue: dV_j (first warpgroup) and dK_j (second) out of TMEM, rows of kept keys only
    mbar_wait(pds_free + ((T - 1) % PB), ((T - 1) / PB) & 1);   // last tile's MMAs (and all before) are done
    tc_fence_after();
    // dense token-space outputs: key row r of this tile goes back to token keep_idx[k0 + r]
    const int tok = key_ok ? (QM ? k0 + r : keep_idx[(size_t)b * NT + k0 + r]) : 0;
    const size_t row_off = ((size_t)b * NT + tok) * D + half * DH;
    {
      const int which = hcol;
      if (!QM && key_ok) {
        // Rows of masked keys must read zero (token-space outputs).  Instead of clearing the whole tensor first (two
        // 0.5 GiB memsets per launch at the 16384-token site), the thread that owns kept key k also clears the masked
        // tokens between the previous kept key and its own -- and, for the last kept key, the tail of the sample.
        __nv_bfloat16* base = (which == 0 ? dv : dk) + (size_t)b * NT * D + half * DH;
        const int prev = (k0 + r == 0) ? -1 : keep_idx[(size_t)b * NT + k0 + r - 1];
        for (int z = prev + 1; z < tok; ++z)
#pragma unroll
          for (int g = 0; g < DH / 8; ++g) *reinterpret_cast<uint4*>(base + (size_t)z * D + g * 8) = make_uint4(0u, 0u, 0u, 0u);
        if (k0 + r == nk - 1)
          for (int z = tok + 1; z < NT; ++z)
#pragma unroll
            for (int g = 0; g < DH / 8; ++g) *reinterpret_cast<uint4*>(base + (size_t)z * D + g * 8) = make_uint4(0u, 0u, 0u, 0u);
      }
      __nv_bfloat16* dst = (which == 0 ? dv : dk) + row_off;
      const float mul = which == 0 ? 1.f : scale;
      const int col0 = which == 0 ? Cfg::kTmDV : Cfg::kTmDK;
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t(&sv)[32] = s[0];
        tmem_ld32(lane_base + col0 + c * 32, sv);
        tmem_wait_ld();
        if (key_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(sv[8 * g + 0]) * mul, __uint_as_float(sv[8 * g + 1]) * mul);
            w.y = pack_bf16(__uint_as_float(sv[8 * g + 2]) * mul, __uint_as_float(sv[8 * g + 3]) * mul);
            w.z = pack_bf16(__uint_as_float(sv[8 * g + 4]) * mul, __uint_as_float(sv[8 * g + 5]) * mul);
            w.w = pack_bf16(__uint_as_float(sv[8 * g + 6]) * mul, __uint_as_float(sv[8 * g + 7]) * mul);
            *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = w;
          }
        }
      }
    }
  } else {
    reg_dealloc<88>();
    // ===================================================== dQ reduction warps
    const int quad = warp & 3;
    const int r = quad * 32 + (int)lane_id();
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t v[32];
    int32_t* sem_b = (DET && sem != nullptr) ? sem + ((size_t)(b * (int)gridDim.z + half)) * T : nullptr;
    int prev_qt = -1;                                    // tile whose reduce-add is still in flight (TMA paths)
    for (int i = 0; i < T; ++i) {
      const int qt = tile_of(i);                         // query tile of this step
      mbar_wait_relaxed(dq_full, i & 1);
      if (warp == 12) MU_TRACE(12, i);                   // dQ warps: dq_full(i)
      tc_fence_after();
      if (DQT && Cfg::kDQTma) {
        // dQ^T tile: lanes = channels, columns = queries.  Staged transposed ([query][channel], 128-byte rows,
        // TMA swizzle; a warp's 32 lanes fill one row segment per store) and added by the TMA unit.
        const uint32_t stage = smem_u32(sDQ);
        const bool issuer = (threadIdx.x == kBwdThreads - 128);
        if (issuer) tma_store_wait_read<0>();
        named_bar_sync(2, 128);
        const uint32_t sub = stage + (r >> 5) * (BM * 128);      // sub-tile = 32 channels
        const uint32_t cj = (uint32_t)((r & 31) >> 2), cw = (uint32_t)(r & 3) * 4;
#pragma unroll
        for (int c = 0; c < BM / 32; ++c) {
          tmem_ld32(lane_base + Cfg::kTmDQ + c * 32, v);
          tmem_wait_ld();
          if (c == BM / 32 - 1) {
            tc_fence_before();
            mbar_arrive(dq_free);
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const uint32_t qrow = (uint32_t)(c * 32 + e);
            st_shared_f32(sub + qrow * 128 + ((cj ^ (qrow & 7)) * 16) + cw, __uint_as_float(v[e]) * scale);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (issuer) {
          if (sem_b != nullptr) sem_wait_turn(sem_b + qt, turn_of(qt));
#pragma unroll
          for (int c = 0; c < DH / 32; ++c)
            tma_reduce_add_3d(&tmap_dq, stage + c * (BM * 128), half * DH + c * 32, qt * BM, b);
          tma_store_commit();
          if (sem_b != nullptr) {
#if MU_BWD_DET_LAG
            if (prev_qt >= 0) {                          // the PREVIOUS tile's adds have landed: let its next CTA in
              tma_store_wait<1>();
              sem_publish(sem_b + prev_qt);
            }
            prev_qt = qt;
#else
            tma_store_wait<0>();                         // this tile's adds have landed: let the next CTA of its order in
            sem_publish(sem_b + qt);
#endif
          }
        }
      } else if (DQT) {
        // lanes = channel (half * 128 + r), columns = queries of tile i: coalesced scalar reductions
        float* base = dq_acc + ((size_t)b * N + (size_t)qt * BM) * D + half * DH + r;
        if (sem_b != nullptr) {                          // ordered: nobody adds before this CTA's turn
          if (threadIdx.x == kBwdThreads - 128) sem_wait_turn(sem_b + qt, turn_of(qt));
          named_bar_sync(2, 128);
        }
#pragma unroll
        for (int c = 0; c < BM / 32; ++c) {
          tmem_ld32(lane_base + Cfg::kTmDQ + c * 32, v);
          tmem_wait_ld();
          if (c == BM / 32 - 1) {
            tc_fence_before();
            mbar_arrive(dq_free);
          }
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (qt * BM + c * 32 + e < N) atomicAdd(base + (size_t)(c * 32 + e) * D, __uint_as_float(v[e]) * scale);
        }
        if (sem_b != nullptr) {
          __threadfence();
          named_bar_sync(2, 128);
          if (threadIdx.x == kBwdThreads - 128) sem_publish(sem_b + qt);
        }
      } else if (MU_BWD_DQ_RED) {
        // A/B option: lanes = query row, 16-byte vector reductions straight from registers (no shared-memory staging:
        // 64 KB less port traffic per tile, but 64 RED.128 warp instructions through the LSU instead)
        const int qrow = qt * BM + r;
        float* dst = dq_acc + ((size_t)b * N + qrow) * D;
#pragma unroll
        for (int c = 0; c < DH / 32; ++c) {
          tmem_ld32(lane_base + Cfg::kTmDQ + c * 32, v);
          tmem_wait_ld();
          if (c == DH / 32 - 1) {
            tc_fence_before();
            mbar_arrive(dq_free);
          }
          if (qrow < N) {
#pragma unroll
            for (int g4 = 0; g4 < 8; ++g4)
              red_add_v4(dst + c * 32 + 4 * g4, __uint_as_float(v[4 * g4]) * scale, __uint_as_float(v[4 * g4 + 1]) * scale,
                         __uint_as_float(v[4 * g4 + 2]) * scale, __uint_as_float(v[4 * g4 + 3]) * scale);
          }
        }
      } else {
        // lanes = query row, columns = channels.  The fp32 tile is staged in shared memory (128-byte rows,
        // TMA swizzle) and added into the dQ accumulator by the TMA unit (cp.reduce.async.bulk.tensor .add):
        // no LSU atomics, which were the dominant load on the L1 data pipe.
        const uint32_t stage = smem_u32(sDQ);
        const bool issuer = (threadIdx.x == kBwdThreads - 128);
        if (issuer) tma_store_wait_read<0>();          // the previous tile's reduce has finished reading the stage
        named_bar_sync(2, 128);
        if (warp == 12) MU_TRACE(13, i);                 // dQ warps: staging buffer free
#pragma unroll
        for (int c = 0; c < DH / 32; ++c) {
          tmem_ld32(lane_base + Cfg::kTmDQ + c * 32, v);
          tmem_wait_ld();
          if (c == DH / 32 - 1) {
            tc_fence_before();
            mbar_arrive(dq_free);
          }
          if (Cfg::kDQBf16) {
            // one [BM rows][64 channels] bf16 tile (128-byte rows, TMA swizzle); this pass fills chunks 4c .. 4c + 3
            const uint32_t row = stage + r * 128;
#if MU_BWD_PROBE == 1
            if (scale == 12345.f)
#endif
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const uint32_t chunk = ((uint32_t)((c * 4 + ch) ^ (r & 7))) * 16;
              st_shared_v4(row + chunk,
                           pack_bf16(__uint_as_float(v[8 * ch]) * scale, __uint_as_float(v[8 * ch + 1]) * scale),
                           pack_bf16(__uint_as_float(v[8 * ch + 2]) * scale, __uint_as_float(v[8 * ch + 3]) * scale),
                           pack_bf16(__uint_as_float(v[8 * ch + 4]) * scale, __uint_as_float(v[8 * ch + 5]) * scale),
                           pack_bf16(__uint_as_float(v[8 * ch + 6]) * scale, __uint_as_float(v[8 * ch + 7]) * scale));
            }
            continue;
          }
          const uint32_t row = stage + c * (BM * 128) + r * 128;   // sub-tile c = channels [32c, 32c+32)
#if MU_BWD_PROBE == 1
          if (scale == 12345.f)      // never true: the staging stores and the reduce are compiled but skipped
#endif
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint32_t chunk = ((uint32_t)(ch ^ (r & 7))) * 16;
            st_shared_v4(row + chunk, __float_as_uint(__uint_as_float(v[4 * ch]) * scale),
                         __float_as_uint(__uint_as_float(v[4 * ch + 1]) * scale),
                         __float_as_uint(__uint_as_float(v[4 * ch + 2]) * scale),
                         __float_as_uint(__uint_as_float(v[4 * ch + 3]) * scale));
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
#if MU_BWD_PROBE == 1
        if (scale == 12345.f)
#endif
        if (issuer) {
          if (sem_b != nullptr) sem_wait_turn(sem_b + qt, turn_of(qt));
          if (Cfg::kDQBf16) {
            tma_reduce_add_3d(&tmap_dq, stage, 0, qt * BM, b);
          } else {
#pragma unroll
            for (int c = 0; c < DH / 32; ++c) tma_reduce_add_3d(&tmap_dq, stage + c * (BM * 128), c * 32, qt * BM, b);
          }
          tma_store_commit();
          if (sem_b != nullptr) {
#if MU_BWD_DET_LAG
            if (prev_qt >= 0) {                          // the PREVIOUS tile's adds have landed: let its next CTA in
              tma_store_wait<1>();
              sem_publish(sem_b + prev_qt);
            }
            prev_qt = qt;
#else
            tma_store_wait<0>();                         // this tile's adds have landed: let the next CTA of its order in
            sem_publish(sem_b + qt);
#endif
          }
        }
        if (warp == 12) MU_TRACE(14, i);                 // dQ warps: reduce issued
      }
    }
    if (Cfg::kDQTma && threadIdx.x == kBwdThreads - 128) {
      tma_store_wait<0>();                               // all reduce-adds landed before exit
      if (sem_b != nullptr && prev_qt >= 0) sem_publish(sem_b + prev_qt);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<512>(tmem_base);
  }
}

// fp32 accumulator -> bf16 dq
__global__ void dq_convert_kernel(const float4* __restrict__ acc, uint2* __restrict__ out, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 a = acc[i];
    uint2 o;
    o.x = pack_bf16(a.x, a.y);
    o.y = pack_bf16(a.z, a.w);
    out[i] = o;
  }
}

// aug bf16 [B][Npad][16] for MU_BWD_FOLD_STATS: columns 0-2 = -lse / scale, columns 8-10 = -delta, each split into three
// bf16 terms (hi + mid + lo carries 24 bits); rows [N, Npad) get -3e4 in the lse slot, so that P = exp2(c S') = 0 for
// query columns past N (a finite value: the slab also meets the zero rows of the other constant operand, and 0 * inf
// would be NaN).
__device__ __forceinline__ void split3_bf16(float a, uint32_t& w01, uint32_t& w2) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(a);
  const float r1 = a - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
  w01 = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16);
  w2 = (uint32_t)__bfloat16_as_ushort(lo);
}
__global__ void attn_bwd_aug_kernel(const float* __restrict__ lse, const float* __restrict__ delta, uint4* __restrict__ aug,
                                    int N, int Npad, long total, float inv_scale) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / Npad;
    const int n = (int)(i - b * Npad);
    float a = -3.0e4f, d = 0.f;
    if (n < N) {
      a = fminf(fmaxf(-lse[b * N + n] * inv_scale, -3.0e4f), 3.0e4f);   // lse = -inf (a query without keys): finite, 0 * inf = NaN in dP'
      d = -delta[b * N + n];
    }
    uint4 w0 = make_uint4(0u, 0u, 0u, 0u), w1 = make_uint4(0u, 0u, 0u, 0u);
    split3_bf16(a, w0.x, w0.y);
    split3_bf16(d, w1.x, w1.y);
    aug[2 * i] = w0;
    aug[2 * i + 1] = w1;
  }
}

template <int D, int BM, int DH, int STAGES, int PB, bool QM = false>
static int run(const void* q, const void* kc, const void* vc, const int32_t* n_keep, const int32_t* keep_idx,
               const void* d_o, const float* lse, const float* delta, void* dq, void* dkc, void* dvc, float* dq_acc,
               int B, int N, int NKP, cudaStream_t s, int NT = 0, float scale_in = 0.f, const uint32_t* bits_t = nullptr,
               int heads = 1, int nk_all = 0, int32_t* sem = nullptr) {
  if (NT == 0) NT = N;                     // self-attention: as many key tokens as queries
  using Cfg = BwdCfg<D, BM, DH, STAGES, PB>;
  CUtensorMap tq, tdo, tk, tv, tdq, taug;
  int rc;
  std::memset(&taug, 0, sizeof(taug));
  const float scale_run = scale_in > 0.f ? scale_in : 1.f / sqrtf((float)D);
  if (Cfg::kFold) {           // the statistics slab lives at the start of the workspace (where the fp32 dQ tile used to)
    const int Npad = round_up(N, BM);
    const long total = (long)B * Npad;
    attn_bwd_aug_kernel<<<(int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8), 256, 0, s>>>(
        lse, delta, reinterpret_cast<uint4*>(dq_acc), N, Npad, total, 1.f / scale_run);
    if ((rc = check_launch("attn_bwd_aug"))) return rc;
    if ((rc = make_tmap_bf16_3d_w16(&taug, dq_acc, Npad, B, BM))) return rc;
  }
  if (Cfg::kDQBf16) {
    if ((rc = make_tmap_bf16_3d(&tdq, dq, D, N, B, BM))) return rc;      // partial tiles are added straight into dq
  } else {
    if ((rc = make_tmap_f32_3d(&tdq, dq_acc, D, N, B, BM))) return rc;
  }
  if ((rc = make_tmap_bf16_3d(&tq, q, D, N, B, BM))) return rc;
  if ((rc = make_tmap_bf16_3d(&tdo, d_o, D, N, B, BM))) return rc;
  if ((rc = make_tmap_bf16_3d(&tk, kc, D, NKP, B, kBK))) return rc;
  if ((rc = make_tmap_bf16_3d(&tv, vc, D, NKP, B, kBK))) return rc;
  auto kern = (sem != nullptr && !QM) ? attn_bwd_sm100_kernel<D, BM, DH, STAGES, PB, QM, !QM>
                                      : attn_bwd_sm100_kernel<D, BM, DH, STAGES, PB, QM, false>;
  if (QM) sem = nullptr;                      // generalised mode has no ordered accumulation
  cudaError_t e = set_max_dynamic_smem_once(kern, Cfg::kSmemBytes);
  if (e != cudaSuccess) {
    set_error("attn_bwd_sm100: cudaFuncSetAttribute(%d bytes): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
    return (int)e;
  }
  const size_t n = (size_t)B * N * D, nt = (size_t)B * NT * D;
  // one memset clears the fp32 dQ accumulator and, in deterministic mode, the order semaphores right behind it
  const size_t sem_bytes = sem != nullptr ? (size_t)B * (D / DH) * ((N + BM - 1) / BM) * sizeof(int32_t) : 0;
  if (Cfg::kDQBf16) {
    cudaMemsetAsync(dq, 0, n * sizeof(__nv_bfloat16), s);
    if (sem_bytes) cudaMemsetAsync(sem, 0, sem_bytes, s);
  } else {
    cudaMemsetAsync(dq_acc, 0, n * sizeof(float) + sem_bytes, s);
  }
  (void)nt;   // dk / dv need no clearing: the kernel writes every row (kept keys: gradients, masked keys: zeros)
  dim3 grid(NKP / kBK, B, D / DH);
  const float scale = scale_in > 0.f ? scale_in : 1.f / sqrtf((float)D);
  kern<<<grid, kBwdThreads, Cfg::kSmemBytes, s>>>(tq, tdo, tk, tv, tdq, taug, n_keep, keep_idx, lse, delta, dq_acc, (__nv_bfloat16*)dkc,
                                                  (__nv_bfloat16*)dvc, N, NKP, scale, NT, bits_t, heads,
                                                  round_up(N, 128) / 32, nk_all, sem);
  if ((rc = check_launch("attn_bwd_sm100"))) return rc;
  if (Cfg::kDQBf16) return 0;
  const size_t n4 = n / 4;
  const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  dq_convert_kernel<<<blocks, 256, 0, s>>>((const float4*)dq_acc, (uint2*)dq, n4);
  return check_launch("dq_convert");
}

#if MU_BWD_TRACE
extern "C" int mu_debug_bwd_trace(long long* host, int n) {   // tools/bwd_trace.py only (trace builds)
  return (int)cudaMemcpyFromSymbol(host, g_bwd_trace, sizeof(long long) * (n < kTraceTiles * kTraceEvents ? n : kTraceTiles * kTraceEvents));
}
#endif

// fp32 dQ accumulator [B, N, C] + order semaphores int32 [B, C / DH, ceil(N / 64)] (sized for the smallest query tile)
static size_t dq_acc_bytes(int B, int N, int C) {
  if (MU_BWD_DQ_BF16 != 0 && MU_BWD_DQ_RED == 0 && C == 64)              // partial tiles are added into dq itself;
    return MU_BWD_FOLD_STATS != 0 ? (size_t)B * round_up(N, 128) * 32 : 0;   // what is left is the statistics slab
  return (size_t)B * N * C * sizeof(float);
}
size_t attn_bwd_sm100_workspace(int B, int N, int C) {
  return dq_acc_bytes(B, N, C) + (size_t)B * 2 * ((N + 63) / 64) * sizeof(int32_t) + 16;
}

static std::atomic<int> g_deterministic{0};
void set_deterministic(int on) { g_deterministic.store(on ? 1 : 0); }
int get_deterministic() { return g_deterministic.load(); }

int launch_attn_bwd_sm100(const void* q, const void* kc, const void* vc, const int32_t* n_keep,
                          const int32_t* keep_idx, const void* d_o, const float* lse, const float* delta, void* dq,
                          void* dkc, void* dvc, void* workspace, size_t workspace_bytes, int B, int N, int NKP, int C,
                          cudaStream_t s) {
  MU_REQUIRE(workspace != nullptr && workspace_bytes >= attn_bwd_sm100_workspace(B, N, C), MU_ERR_WORKSPACE,
             "mu_attn_bwd: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
             attn_bwd_sm100_workspace(B, N, C));
  float* acc = (float*)workspace;
  // deterministic mode: semaphores live right behind the accumulator (same memset); nullptr = free-running adds
  int32_t* sem = get_deterministic() ? reinterpret_cast<int32_t*>((char*)workspace + dq_acc_bytes(B, N, C)) : nullptr;
  switch (C) {
    case 64:
      return run<64, 128, 64, MU_BWD_STAGES_64, 2>(q, kc, vc, n_keep, keep_idx, d_o, lse, delta, dq, dkc, dvc, acc, B, N, NKP, s,
                                                   0, 0.f, nullptr, 1, 0, sem);
    case 128:
      return run<128, 64, 128, MU_BWD_STAGES_128, 2>(q, kc, vc, n_keep, keep_idx, d_o, lse, delta, dq, dkc, dvc, acc, B, N, NKP, s,
                                                     0, 0.f, nullptr, 1, 0, sem);
    case 256:
      return run<256, 64, 128, 1, 1>(q, kc, vc, n_keep, keep_idx, d_o, lse, delta, dq, dkc, dvc, acc, B, N, NKP, s,
                                     0, 0.f, nullptr, 1, 0, sem);
    default:
      set_error("attn_bwd_sm100: channels must be 64, 128 or 256 (got %d)", C);
      return MU_ERR_BAD_SHAPE;
  }
}

// Generalised mode: q, d_o [BH, Q, 64], k / v [BH, NKP, 64], bits_t [BH / heads, NKP, roundup(Q, 128) / 32];
// dq [BH, Q, 64], dk / dv [BH, N, 64]; workspace = fp32 dQ accumulator.
int launch_query_attn_bwd_sm100(const void* q, const void* k, const void* v, const uint32_t* bits_t, const void* d_o,
                                const float* lse, const float* delta, void* dq, void* dk, void* dv, void* workspace,
                                size_t workspace_bytes, int BH, int heads, int Q, int N, int NKP, int D, float scale,
                                cudaStream_t s) {
  MU_REQUIRE(D == 64, MU_ERR_BAD_SHAPE, "mu_query_attn_bwd: head dim must be 64 (pad 32 to 64), got %d", D);
  MU_REQUIRE(workspace != nullptr && workspace_bytes >= attn_bwd_sm100_workspace(BH, Q, D), MU_ERR_WORKSPACE,
             "mu_query_attn_bwd: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
             attn_bwd_sm100_workspace(BH, Q, D));
  return run<64, 128, 64, MU_BWD_STAGES_64, 2, true>(q, k, v, nullptr, nullptr, d_o, lse, delta, dq, dk, dv, (float*)workspace, BH, Q,
                                      NKP, s, N, scale, bits_t, heads, N);
}

}  // namespace mu
