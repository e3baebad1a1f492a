// K1 / K6 on tcgen05 tensor cores for token-major (channels-last) bf16 activations.
// Replaces nn.Linear x3 on the token view and its autograd, /root/reference/code/ade20k/ade_semantic.py:168-172.
// With NHWC activations the reference's permute(0, 2, 1) view IS the memory layout: xt = [B*N, C] row-major.
//
//   P1  [Q | K | V] = xt Wqkv^T + b        M = tokens, N = 3C in chunks, K = C;  K / V rows written compacted
//   P2  dxt = dz + [dq | dk | dv] Wqkv     M = tokens, N = C, K = 3C in 64-wide chunks
//   P3  dWqkv += [dq | dk | dv]^T xt       M = 128 output channels, N = C, K = tokens (split over CTAs, fp32 red.add)
//   db  column sums of dq | dk | dv        (HBM-bound reduction, 16-byte loads)
// All three GEMMs are HBM/L2-bound (K <= 768): the point of tensor cores here is to keep the math off the
// critical path, the layout work (TMA 128B-swizzled tiles, MN-major operands for the transposed products) is
// what removes every explicit transpose.
#include "common.cuh"
#include "det_reduce.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

// Measured at B = 256, N = 16384, C = 64 (tools/bench_qkv.py): one accumulator buffer (64 TMEM columns, more of these
// short latency-bound CTAs resident) 0.572 ms against 0.597 ms with two.
#ifndef MU_P1_ACC_BUFS_64
#define MU_P1_ACC_BUFS_64 1
#endif
// C = 256: the 64 KB token tile dominates shared memory; 64-column chunks in a 1-deep weight ring fit two CTAs per SM
// instead of one: 0.362 -> 0.259 ms at B = 256, N = 1024 (128-column chunks, 2-deep ring: MU_P1_NC_256=128
// MU_P1_WSTAGES_256=2).
#ifndef MU_P1_STAGED_EPILOGUE
#define MU_P1_STAGED_EPILOGUE 1
#endif
#ifndef MU_P2_STAGED_EPILOGUE
#define MU_P2_STAGED_EPILOGUE 1
#endif
#ifndef MU_P3_FOLD_DB
#define MU_P3_FOLD_DB 1
#endif
#ifndef MU_P1_NC_256
#define MU_P1_NC_256 64
#endif
#ifndef MU_P1_WSTAGES_256
#define MU_P1_WSTAGES_256 1
#endif
#ifndef MU_P1_NC_128
#define MU_P1_NC_128 64     // output columns per chunk at C = 128: 64-column chunks (smaller weight tiles, 3 CTAs per SM
                            // instead of 2) 0.378 ms against 0.503 ms with 128-column chunks at B = 256, N = 4096
#endif

namespace mu {

constexpr int kGemmThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

// ============================================================================ P1: forward projection
template <int C>
struct P1Cfg {
  static constexpr int NC = (C == 64) ? 64 : (C == 128 ? MU_P1_NC_128 : MU_P1_NC_256);   // output columns per chunk
  static constexpr int kWStages = (C == 256) ? MU_P1_WSTAGES_256 : 2;                  // weight-chunk ring depth
  static constexpr int kChunks = 3 * C / NC;
  static constexpr int kXBytes = 128 * C * 2;
  static constexpr int kWBytes = NC * C * 2;
  // accumulator buffers in TMEM: 2 lets the MMA of chunk j+1 run under the epilogue of chunk j; 1 halves the TMEM
  // columns per CTA (64 at C = 64), so that twice as many of these short, latency-bound CTAs are resident
  static constexpr int kAcc = (C == 64) ? MU_P1_ACC_BUFS_64 : 2;
  static constexpr int kTmemCols = kAcc * NC;           // 64, 128 or 256 (power of two)
  // Epilogue staging: 2 KB per epilogue warp (32 token rows x 32 bf16 columns).  A thread owns a token ROW of the
  // accumulator; storing its row straight to global memory makes every warp store touch 32 different 128-byte lines
  // with 16 bytes each.  Through the staging tile the warp writes 8 rows x 64 contiguous bytes per instruction
  // (MU_P1_STAGED_EPILOGUE=0: the direct stores).
  static constexpr int kStageBytes = MU_P1_STAGED_EPILOGUE ? 4 * 2048 : 0;
  static constexpr int kSmemBytes = 1024 + kXBytes + kWStages * kWBytes + kStageBytes + 256;
};

template <int C>
__global__ void __launch_bounds__(kGemmThreads, C == 256 ? 2 : (C == 128 ? 4 : 5))   // measured (tools/bench_qkv.py): 64 registers / five CTAs
                                              // per SM at C = 64 (0.486 -> 0.464 ms), uncapped at C = 128 (80 registers: 0.288 ms, 0.349 capped)
qkv_project_sm100_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         const float* __restrict__ bias, const int32_t* __restrict__ rank,
                         __nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ kc,
                         __nv_bfloat16* __restrict__ vc, int Mtot, int N, int NKP) {
  using Cfg = P1Cfg<C>;
  constexpr int NC = Cfg::NC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;
  uint8_t* sW = sX + Cfg::kXBytes;
  uint8_t* sStage = sW + Cfg::kWStages * Cfg::kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + Cfg::kStageBytes);
  uint64_t* x_full = bars;
  uint64_t* w_full = bars + 1;      // 2
  uint64_t* w_empty = bars + 3;     // 2
  uint64_t* acc_full = bars + 5;    // 2
  uint64_t* acc_free = bars + 7;    // 2 (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int t0 = blockIdx.x * 128;

  if (threadIdx.x == 0) {
    mbar_init(x_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(w_full + i, 1);
      mbar_init(w_empty + i, 1);
      mbar_init(acc_full + i, 1);
      mbar_init(acc_free + i, 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane_id() == 0) {
      mbar_expect_tx(x_full, Cfg::kXBytes);
      for (int blk = 0; blk < C / 64; ++blk) tma_load_3d(sX + blk * 16384, &tmap_x, x_full, blk * 64, t0, 0);
      for (int j = 0; j < Cfg::kChunks; ++j) {
        const int st = j % Cfg::kWStages, use = j / Cfg::kWStages;
        if (use > 0) mbar_wait(w_empty + st, (use - 1) & 1);
        mbar_expect_tx(w_full + st, Cfg::kWBytes);
        for (int blk = 0; blk < C / 64; ++blk)
          tma_load_3d(sW + st * Cfg::kWBytes + blk * (NC * 128), &tmap_w, w_full + st, blk * 64, j * NC, 0);
      }
    }
  } else if (warp == 1) {
    {   // whole warp walks the loop, one elected lane issues; 32-bit descriptor words (see umma_ss_lo)
      constexpr uint32_t idesc = make_idesc_bf16(128, NC, 0, 0);
      constexpr uint32_t hi = desc_hi_sbo(1024);
      const uint32_t x_lo = desc_lo(smem_u32(sX)), w_lo0 = desc_lo(smem_u32(sW));
      mbar_wait(x_full, 0);
      for (int j = 0; j < Cfg::kChunks; ++j) {
        const int st = j % Cfg::kWStages, use = j / Cfg::kWStages;                  // weight ring slot / its use count
        const int ast = Cfg::kAcc == 2 ? (j & 1) : 0, ause = Cfg::kAcc == 2 ? (j >> 1) : j;   // accumulator buffer / use count
        mbar_wait(w_full + st, use & 1);
        if (ause > 0) mbar_wait(acc_free + ast, (ause - 1) & 1);
        tc_fence_after();
        const uint32_t w_lo = w_lo0 + st * (Cfg::kWBytes >> 4);
#pragma unroll
        for (int kk = 0; kk < C / 16; ++kk)
          if (elect_one())
            umma_ss_lo(tmem_base + ast * NC, x_lo + (((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                       w_lo + (((kk >> 2) * (NC * 128) + (kk & 3) * 32) >> 4), hi, idesc, kk > 0 ? 1u : 0u);
        if (elect_one()) {
          umma_commit(w_empty + st);
          umma_commit(acc_full + ast);
        }
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + (int)lane_id();
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int tok = t0 + r;
    const bool tok_ok = tok < Mtot;
    int rk = -1, bidx = 0;
    if (tok_ok) {
      rk = rank[tok];
      bidx = tok / N;
    }
    uint32_t v[32];
#if MU_P1_STAGED_EPILOGUE
    // destination rows of the four store rounds: round `it` writes rows it * 8 + lane / 4, 16-byte piece lane % 4
    const int lane = (int)lane_id();
    const uint32_t stage = smem_u32(sStage) + quad * 2048;
    int kv_row[4];                                       // compacted K / V row of the token (-1: masked or past the end)
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 8 + (lane >> 2);
      const int rkr = __shfl_sync(0xffffffffu, tok_ok ? rk : -1, row);
      const int br = __shfl_sync(0xffffffffu, bidx, row);
      kv_row[it] = rkr >= 0 ? br * NKP + rkr : -1;
    }
    for (int j = 0; j < Cfg::kChunks; ++j) {
      const int st = Cfg::kAcc == 2 ? (j & 1) : 0, use = Cfg::kAcc == 2 ? (j >> 1) : j;
      const int which = (j * NC) / C, col0 = j * NC - which * C;
      __nv_bfloat16* const base = which == 0 ? q : (which == 1 ? kc : vc);
      mbar_wait(acc_full + st, use & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < NC / 32; ++c) {
        tmem_ld32(lane_base + st * NC + c * 32, v);
        tmem_wait_ld();
        if (c == NC / 32 - 1) {
          tc_fence_before();
          mbar_arrive(acc_free + st);
        }
        const float* bptr = bias + j * NC + c * 32;
        __syncwarp();                                    // the previous round's reads of the staging tile are done
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          st_shared_v4(stage + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4),
                       pack_bf16(__uint_as_float(v[8 * g + 0]) + __ldg(bptr + 8 * g + 0), __uint_as_float(v[8 * g + 1]) + __ldg(bptr + 8 * g + 1)),
                       pack_bf16(__uint_as_float(v[8 * g + 2]) + __ldg(bptr + 8 * g + 2), __uint_as_float(v[8 * g + 3]) + __ldg(bptr + 8 * g + 3)),
                       pack_bf16(__uint_as_float(v[8 * g + 4]) + __ldg(bptr + 8 * g + 4), __uint_as_float(v[8 * g + 5]) + __ldg(bptr + 8 * g + 5)),
                       pack_bf16(__uint_as_float(v[8 * g + 6]) + __ldg(bptr + 8 * g + 6), __uint_as_float(v[8 * g + 7]) + __ldg(bptr + 8 * g + 7)));
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int row = it * 8 + (lane >> 2), g = lane & 3;
          const int qrow = t0 + quad * 32 + row;
          const int drow = which == 0 ? (qrow < Mtot ? qrow : -1) : kv_row[it];
          const float4 w = ld_shared_v4f(stage + row * 64 + ((g ^ ((row >> 1) & 3)) << 4));
          if (drow >= 0) *reinterpret_cast<float4*>(base + (size_t)drow * C + col0 + c * 32 + g * 8) = w;
        }
      }
    }
#else
    for (int j = 0; j < Cfg::kChunks; ++j) {
      const int st = Cfg::kAcc == 2 ? (j & 1) : 0, use = Cfg::kAcc == 2 ? (j >> 1) : j;
      const int which = (j * NC) / C, col0 = j * NC - which * C;
      __nv_bfloat16* dst = nullptr;
      if (tok_ok) {
        if (which == 0) dst = q + (size_t)tok * C + col0;
        else if (rk >= 0) dst = (which == 1 ? kc : vc) + ((size_t)bidx * NKP + rk) * C + col0;
      }
      mbar_wait(acc_full + st, use & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < NC / 32; ++c) {
        tmem_ld32(lane_base + st * NC + c * 32, v);
        tmem_wait_ld();
        if (c == NC / 32 - 1) {
          tc_fence_before();
          mbar_arrive(acc_free + st);
        }
        if (dst != nullptr) {
          const float* bptr = bias + j * NC + c * 32;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(v[8 * g + 0]) + __ldg(bptr + 8 * g + 0), __uint_as_float(v[8 * g + 1]) + __ldg(bptr + 8 * g + 1));
            w.y = pack_bf16(__uint_as_float(v[8 * g + 2]) + __ldg(bptr + 8 * g + 2), __uint_as_float(v[8 * g + 3]) + __ldg(bptr + 8 * g + 3));
            w.z = pack_bf16(__uint_as_float(v[8 * g + 4]) + __ldg(bptr + 8 * g + 4), __uint_as_float(v[8 * g + 5]) + __ldg(bptr + 8 * g + 5));
            w.w = pack_bf16(__uint_as_float(v[8 * g + 6]) + __ldg(bptr + 8 * g + 6), __uint_as_float(v[8 * g + 7]) + __ldg(bptr + 8 * g + 7));
            *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = w;
          }
        }
      }
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ============================================================================ P2: dxt = dz + [dq|dk|dv] Wqkv
template <int C>
struct P2Cfg {
  static constexpr int kStages = 3;
  static constexpr int kKChunks = 3 * C / 64;
  static constexpr int kABytes = 128 * 128;              // [128 tokens x 64 channels]
  static constexpr int kWBytes = (C / 64) * 64 * 128;    // [64 K-rows x C] as C/64 blocks of [64 x 128 B]
  static constexpr int kStageBytes = kABytes + kWBytes;
  static constexpr int kTmemCols = C < 32 ? 32 : C;      // 64 / 128 / 256
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256;
};

template <int C>
__global__ void __launch_bounds__(kGemmThreads)
qkv_dx_sm100_kernel(const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ CUtensorMap tmap_dk,
                    const __grid_constant__ CUtensorMap tmap_dv, const __grid_constant__ CUtensorMap tmap_w,
                    const __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ dxt, int Mtot) {
  using Cfg = P2Cfg<C>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* full = bars;          // S
  uint64_t* empty = bars + S;     // S
  uint64_t* acc_full = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int t0 = blockIdx.x * 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane_id() == 0) {
      for (int j = 0; j < Cfg::kKChunks; ++j) {
        const int st = j % S, use = j / S;
        if (use > 0) mbar_wait(empty + st, (use - 1) & 1);
        uint8_t* sA = smem + st * Cfg::kStageBytes;
        uint8_t* sW = sA + Cfg::kABytes;
        mbar_expect_tx(full + st, Cfg::kStageBytes);
        const int src = j / (C / 64), col = (j % (C / 64)) * 64;
        const CUtensorMap* tm = src == 0 ? &tmap_dq : (src == 1 ? &tmap_dk : &tmap_dv);
        tma_load_3d(sA, tm, full + st, col, t0, 0);
        for (int nb = 0; nb < C / 64; ++nb) tma_load_3d(sW + nb * 8192, &tmap_w, full + st, nb * 64, j * 64, 0);
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, C, 0, 1);   // A K-major, B (= W rows) MN-major
      constexpr uint32_t hi = desc_hi_sbo(1024);
      constexpr uint32_t kLboW = ((8192u >> 4) & 0x3FFFu) << 16;
      const uint32_t base_lo = desc_lo(smem_u32(smem));
      for (int j = 0; j < Cfg::kKChunks; ++j) {
        const int st = j % S;
        mbar_wait(full + st, (j / S) & 1);
        tc_fence_after();
        const uint32_t a_lo = base_lo + st * (Cfg::kStageBytes >> 4), w_lo = (a_lo + (Cfg::kABytes >> 4)) | kLboW;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (elect_one()) umma_ss_lo(tmem_base, a_lo + kk * 2, w_lo + kk * 128, hi, idesc, (j > 0 || kk > 0) ? 1u : 0u);
        if (elect_one()) umma_commit(empty + st);
      }
      if (elect_one()) umma_commit(acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + (int)lane_id();
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int tok = t0 + r;
    const bool ok = tok < Mtot;
    const __nv_bfloat16* zrow = dz + (size_t)tok * C;
    __nv_bfloat16* orow = dxt + (size_t)tok * C;
#if MU_P2_STAGED_EPILOGUE
    // Coalesced epilogue (see P1): every MMA has completed when acc_full fires, so the operand ring is free and its first
    // 8 KB stage the tile, 2 KB per warp.  Per 32-channel piece: dz rows arrive as 8 rows x 64 contiguous bytes per warp
    // load, the row's owner adds its accumulator in fp32 and rounds once, the bf16 result leaves the same way.
    const int lane = (int)lane_id();
    const uint32_t stage = smem_u32(smem) + quad * 2048;
    const uint32_t mine = stage + lane * 64, sw = (lane >> 1) & 3;
    (void)zrow; (void)orow; (void)ok;
    uint32_t v[32];
    mbar_wait_relaxed(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < C / 32; ++c) {
      tmem_ld32(lane_base + c * 32, v);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int row = it * 8 + (lane >> 2), g = lane & 3;
        const int trow = t0 + quad * 32 + row;
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (trow < Mtot) z = *reinterpret_cast<const float4*>(dz + (size_t)trow * C + c * 32 + g * 8);
        st_shared_v4(stage + row * 64 + ((g ^ ((row >> 1) & 3)) << 4), __float_as_uint(z.x), __float_as_uint(z.y),
                     __float_as_uint(z.z), __float_as_uint(z.w));
      }
      tmem_wait_ld();
      __syncwarp();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t a = mine + ((g ^ sw) << 4);
        const float4 zf = ld_shared_v4f(a);
        const uint32_t zz[4] = {__float_as_uint(zf.x), __float_as_uint(zf.y), __float_as_uint(zf.z), __float_as_uint(zf.w)};
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float lo = __uint_as_float(zz[e] << 16), hi = __uint_as_float(zz[e] & 0xffff0000u);
          w[e] = pack_bf16(__uint_as_float(v[8 * g + 2 * e]) + lo, __uint_as_float(v[8 * g + 2 * e + 1]) + hi);
        }
        st_shared_v4(a, w[0], w[1], w[2], w[3]);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int row = it * 8 + (lane >> 2), g = lane & 3;
        const int trow = t0 + quad * 32 + row;
        const float4 w = ld_shared_v4f(stage + row * 64 + ((g ^ ((row >> 1) & 3)) << 4));
        if (trow < Mtot) *reinterpret_cast<float4*>(dxt + (size_t)trow * C + c * 32 + g * 8) = w;
      }
      __syncwarp();
    }
#else
    uint32_t v[32];
    mbar_wait_relaxed(acc_full, 0);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < C / 32; ++c) {
      tmem_ld32(lane_base + c * 32, v);
      tmem_wait_ld();
      if (ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 z = *reinterpret_cast<const uint4*>(zrow + c * 32 + g * 8);
          const uint32_t zz[4] = {z.x, z.y, z.z, z.w};
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lo = __uint_as_float(zz[e] << 16), hi = __uint_as_float(zz[e] & 0xffff0000u);
            w[e] = pack_bf16(__uint_as_float(v[8 * g + 2 * e]) + lo, __uint_as_float(v[8 * g + 2 * e + 1]) + hi);
          }
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  #endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ============================================================================ P3: dW += [dq|dk|dv]^T xt
template <int C>
struct P3Cfg {
  static constexpr int kStages = (C == 256) ? 2 : 3;
  static constexpr int kMTiles = (C == 64) ? 2 : 3 * C / 128;
  static constexpr int kABytes = 2 * 16384;               // two [128 tokens x 64 out-channels] blocks
  static constexpr int kXBytes = (C / 64) * 16384;        // [128 tokens x C]
  static constexpr int kStageBytes = kABytes + kXBytes;
  // MU_P3_FOLD_DB: the bias gradient db = column sums of dq | dk | dv is one more output COLUMN of this GEMM:
  // [dq|dk|dv]^T . 1.  A constant all-ones B tile (16 token rows x 128 bytes: all ones is its own swizzle image) and one
  // N = 16 MMA per K step put it into TMEM columns [C, C + 16); the separate pass over dq, dk, dv (qkv_db_kernel, 1.6 GB
  // at the 16384-token site) disappears.  Free-running mode only: deterministic mode keeps the ordered db kernel.
  static constexpr int kOnesBytes = MU_P3_FOLD_DB ? 2048 : 0;
  static constexpr int kTmemCols = MU_P3_FOLD_DB ? 2 * C : C;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kOnesBytes + 256;
};

template <int C>
__global__ void __launch_bounds__(kGemmThreads)
qkv_dw_sm100_kernel(const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ CUtensorMap tmap_dk,
                    const __grid_constant__ CUtensorMap tmap_dv, const __grid_constant__ CUtensorMap tmap_x,
                    float* __restrict__ dw, int Mtot, int chunks_per_cta, float* __restrict__ det_slices,
                    float* __restrict__ db) {
  using Cfg = P3Cfg<C>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sOnes = smem + S * Cfg::kStageBytes;             // (1024-byte aligned: the stages are multiples of 16 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + Cfg::kOnesBytes);
  const bool fold_db = MU_P3_FOLD_DB != 0 && db != nullptr;
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* acc_full = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int m = blockIdx.y;
  const int total_chunks = (Mtot + 127) / 128;
  const int c_begin = blockIdx.x * chunks_per_cta;
  const int c_end = min(total_chunks, c_begin + chunks_per_cta);
  const int nchunks = c_end - c_begin;
  if (nchunks <= 0) return;
  // the two 64-wide output-channel blocks of this M tile: (source tensor, column offset)
  int src[2], col[2];
  if (C == 64) {
    src[0] = (m == 0) ? 0 : 2; col[0] = 0;
    src[1] = (m == 0) ? 1 : 2; col[1] = 0;          // m == 1: second block duplicates dv, its rows are ignored
  } else {
    src[0] = src[1] = (m * 128) / C;
    col[0] = (m * 128) % C; col[1] = col[0] + 64;
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  if (fold_db) {
    for (int i = threadIdx.x; i < Cfg::kOnesBytes / 16; i += kGemmThreads)
      st_shared_v4(smem_u32(sOnes) + i * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);   // bf16 1.0 pairs
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane_id() == 0) {
      for (int j = 0; j < nchunks; ++j) {
        const int st = j % S, use = j / S;
        if (use > 0) mbar_wait(empty + st, (use - 1) & 1);
        uint8_t* sA = smem + st * Cfg::kStageBytes;
        uint8_t* sX = sA + Cfg::kABytes;
        const int t0 = (c_begin + j) * 128;
        mbar_expect_tx(full + st, Cfg::kStageBytes);
        for (int h = 0; h < 2; ++h) {
          const CUtensorMap* tm = src[h] == 0 ? &tmap_dq : (src[h] == 1 ? &tmap_dk : &tmap_dv);
          tma_load_3d(sA + h * 16384, tm, full + st, col[h], t0, 0);
        }
        for (int nb = 0; nb < C / 64; ++nb) tma_load_3d(sX + nb * 16384, &tmap_x, full + st, nb * 64, t0, 0);
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, C, 1, 1);   // both operands MN-major (contraction over tokens)
      constexpr uint32_t hi = desc_hi_sbo(1024);
      constexpr uint32_t kLbo = ((16384u >> 4) & 0x3FFFu) << 16;
      const uint32_t base_lo = desc_lo(smem_u32(smem));
      for (int j = 0; j < nchunks; ++j) {
        const int st = j % S;
        mbar_wait(full + st, (j / S) & 1);
        tc_fence_after();
        const uint32_t a_lo = (base_lo + st * (Cfg::kStageBytes >> 4)) | kLbo;
        const uint32_t x_lo = (base_lo + st * (Cfg::kStageBytes >> 4) + (Cfg::kABytes >> 4)) | kLbo;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          if (elect_one()) umma_ss_lo(tmem_base, a_lo + kk * 128, x_lo + kk * 128, hi, idesc, (j > 0 || kk > 0) ? 1u : 0u);
        if (fold_db) {        // db column: the same A tiles against the all-ones B tile (every K step reads the same 16 rows)
          constexpr uint32_t idesc1 = make_idesc_bf16(128, 16, 1, 1);
          const uint32_t ones_lo = desc_lo(smem_u32(sOnes));
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            if (elect_one()) umma_ss_lo(tmem_base + C, a_lo + kk * 128, ones_lo, hi, idesc1, (j > 0 || kk > 0) ? 1u : 0u);
        }
        if (elect_one()) umma_commit(empty + st);
      }
      if (elect_one()) umma_commit(acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + (int)lane_id();
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int h = r >> 6;
    const bool valid = !(C == 64 && m == 1 && h == 1);
    // deterministic mode: every token split (blockIdx.x) stores its [3C, C] partial into its own slice, a small kernel
    // adds the slices in order; free-running: float atomics into dw
    float* wrow = (det_slices != nullptr ? det_slices + (size_t)blockIdx.x * 3 * C * C : dw) +
                  (size_t)(src[h] * C + col[h] + (r & 63)) * C;
    uint32_t v[32];
    mbar_wait_relaxed(acc_full, 0);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < C / 32; ++c) {
      tmem_ld32(lane_base + c * 32, v);
      tmem_wait_ld();
      if (valid) {
        if (det_slices != nullptr) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            *reinterpret_cast<float4*>(wrow + c * 32 + 4 * e) =
                make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                            __uint_as_float(v[4 * e + 3]));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) atomicAdd(wrow + c * 32 + e, __uint_as_float(v[e]));
        }
      }
    }
    if (fold_db) {            // column C of the accumulator: this CTA's share of db for the row's output channel
      tmem_ld32(lane_base + C, v);
      tmem_wait_ld();
      if (valid) atomicAdd(db + src[h] * C + col[h] + (r & 63), __uint_as_float(v[0]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// deterministic mode: out[i] = slices[0][i] + slices[1][i] + ... in order
__global__ void sum_slices_kernel(const float* __restrict__ slices, float* __restrict__ out, int n, int n_slices) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = slices[i];
  for (int k = 1; k < n_slices; ++k) acc += slices[(size_t)k * n + i];
  out[i] = acc;
}

// ============================================================================ db: column sums of dq | dk | dv
// grid (row blocks, 3).  Each thread owns 8 channels (one 16-byte load per row).
template <int C>
__global__ void __launch_bounds__(256) qkv_db_kernel(const __nv_bfloat16* __restrict__ dq,
                                                     const __nv_bfloat16* __restrict__ dk,
                                                     const __nv_bfloat16* __restrict__ dv, float* __restrict__ db,
                                                     int Mtot, int rows_per_block, const DetCtx det) {
  constexpr int G = C / 8;          // column groups
  constexpr int RL = 256 / G;       // row lanes
  __shared__ float red[RL][C + 1];
  const __nv_bfloat16* src = blockIdx.y == 0 ? dq : (blockIdx.y == 1 ? dk : dv);
  const int g = threadIdx.x % G, rl = threadIdx.x / G;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(Mtot, r0 + rows_per_block);
  float acc[8] = {};
  for (int r = r0 + rl; r < r1; r += RL) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + (size_t)r * C + g * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[2 * e] += __uint_as_float(w[e] << 16);
      acc[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][g * 8 + e] = acc[e];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float s = 0.f;
#pragma unroll 4
    for (int i = 0; i < RL; ++i) s += red[i][c];
    if (det.on()) det.partial[(size_t)blockIdx.x * 3 * C + blockIdx.y * C + c] = s;     // slice = row block
    else atomicAdd(db + blockIdx.y * C + c, s);
  }
  if (det.on())
    det_finish(det, gridDim.x * gridDim.y, gridDim.x, 1, 3 * C, 3 * C, db, db, threadIdx.x, 256, SyncThreads());
}

// ============================================================================ launchers
template <int C>
static int run_p1(const void* xt, const void* w_bf16, const float* bias, const int32_t* rank, void* q, void* kc,
                  void* vc, int B, int N, int NKP, cudaStream_t s) {
  using Cfg = P1Cfg<C>;
  const int Mtot = B * N;
  CUtensorMap tx, tw;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tx, xt, C, Mtot, 1, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tw, w_bf16, C, 3 * C, 1, Cfg::NC))) return rc;
  auto kern = qkv_project_sm100_kernel<C>;
  set_max_dynamic_smem_once(kern, Cfg::kSmemBytes);
  kern<<<(Mtot + 127) / 128, kGemmThreads, Cfg::kSmemBytes, s>>>(tx, tw, bias, rank, (__nv_bfloat16*)q,
                                                                (__nv_bfloat16*)kc, (__nv_bfloat16*)vc, Mtot, N, NKP);
  return check_launch("qkv_project_sm100");
}

template <int C>
static int run_bwd(const void* xt, const void* dz, const void* dq, const void* dk, const void* dv, const void* w_bf16,
                   void* dxt, float* dw, float* db, int B, int N, cudaStream_t s) {
  const int Mtot = B * N;
  CUtensorMap tdq, tdk, tdv, tw, tx;
  int rc;
  bool db_folded = false;
  if ((rc = make_tmap_bf16_3d(&tdq, dq, C, Mtot, 1, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tdk, dk, C, Mtot, 1, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tdv, dv, C, Mtot, 1, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tw, w_bf16, C, 3 * C, 1, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&tx, xt, C, Mtot, 1, 128))) return rc;
  {
    using Cfg = P2Cfg<C>;
    auto kern = qkv_dx_sm100_kernel<C>;
    set_max_dynamic_smem_once(kern, Cfg::kSmemBytes);
    kern<<<(Mtot + 127) / 128, kGemmThreads, Cfg::kSmemBytes, s>>>(tdq, tdk, tdv, tw, (const __nv_bfloat16*)dz,
                                                                  (__nv_bfloat16*)dxt, Mtot);
    if ((rc = check_launch("qkv_dx_sm100"))) return rc;
  }
  {
    using Cfg = P3Cfg<C>;
    auto kern = qkv_dw_sm100_kernel<C>;
    set_max_dynamic_smem_once(kern, Cfg::kSmemBytes);
    const int total_chunks = (Mtot + 127) / 128;
    int per = (total_chunks + 147) / 148;          // about one wave of CTAs per M tile
    if (per < 8) per = 8;
    dim3 grid((total_chunks + per - 1) / per, Cfg::kMTiles);
    DetCtx det;
    if (!det_context(kDetSlotMisc, (size_t)grid.x * 3 * C * C, &det, "qkv_project_bwd (dW)")) return MU_ERR_WORKSPACE;
    db_folded = MU_P3_FOLD_DB != 0 && !det.on();       // free-running: db comes out of the same GEMM (its ones column)
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, s>>>(tdq, tdk, tdv, tx, dw, Mtot, per, det.partial,
                                                    db_folded ? db : nullptr);
    if ((rc = check_launch("qkv_dw_sm100"))) return rc;
    if (det.on()) {
      const int n = 3 * C * C;
      sum_slices_kernel<<<(n + 255) / 256, 256, 0, s>>>(det.partial, dw, n, (int)grid.x);
      if ((rc = check_launch("qkv_dw_sum_slices"))) return rc;
    }
  }
  if (!db_folded) {
    int rows = (Mtot + 295) / 296;
    if (rows < 256) rows = 256;
    dim3 grid((Mtot + rows - 1) / rows, 3);
    DetCtx det;
    if (!det_context(kDetSlotQkvDb, (size_t)grid.x * 3 * C, &det, "qkv_project_bwd")) return MU_ERR_WORKSPACE;
    qkv_db_kernel<C><<<grid, 256, 0, s>>>((const __nv_bfloat16*)dq, (const __nv_bfloat16*)dk,
                                          (const __nv_bfloat16*)dv, db, Mtot, rows, det);
    if ((rc = check_launch("qkv_db"))) return rc;
  }
  return 0;
}

int launch_qkv_project_sm100(const void* xt, const void* w_bf16, const float* bias, const int32_t* rank, void* q,
                             void* kc, void* vc, int B, int C, int N, int NKP, cudaStream_t s) {
  switch (C) {
    case 64: return run_p1<64>(xt, w_bf16, bias, rank, q, kc, vc, B, N, NKP, s);
    case 128: return run_p1<128>(xt, w_bf16, bias, rank, q, kc, vc, B, N, NKP, s);
    case 256: return run_p1<256>(xt, w_bf16, bias, rank, q, kc, vc, B, N, NKP, s);
  }
  set_error("qkv_project_sm100: channels must be 64, 128 or 256 (got %d)", C);
  return MU_ERR_BAD_SHAPE;
}

int launch_qkv_project_bwd_sm100(const void* xt, const void* dz, const void* dq, const void* dk, const void* dv,
                                 const void* w_bf16, void* dxt, float* dw, float* db, int B, int C, int N,
                                 cudaStream_t s) {
  switch (C) {
    case 64: return run_bwd<64>(xt, dz, dq, dk, dv, w_bf16, dxt, dw, db, B, N, s);
    case 128: return run_bwd<128>(xt, dz, dq, dk, dv, w_bf16, dxt, dw, db, B, N, s);
    case 256: return run_bwd<256>(xt, dz, dq, dk, dv, w_bf16, dxt, dw, db, B, N, s);
  }
  set_error("qkv_project_bwd_sm100: channels must be 64, 128 or 256 (got %d)", C);
  return MU_ERR_BAD_SHAPE;
}

}  // namespace mu
