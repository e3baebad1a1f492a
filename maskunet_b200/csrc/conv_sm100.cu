// K7: 3x3 convolution (stride 1, zero padding 1, no bias) of the reference's DoubleConv blocks on tcgen05 tensor
// cores, channels-last bf16 activations, fp32 accumulation.  Replaces nn.Conv2d(.., kernel_size=3, padding=1,
// bias=False) and its autograd, /root/reference/code/ade20k/ade_semantic.py:199, :202 (ConvBlock) as used by
// DownSample :212-229, UpSample :231-256 and the bottom blocks :268-270.
//
// Implicit GEMM, no im2col buffer:  y[p, co] = sum_{tap, ci} x[p + off(tap), ci] * w[co, ci, tap]
//   forward / data gradient  (one kernel; dX is the same convolution of dY with flipped, transposed weights):
//     M = 128 pixels (TH image rows x the full width W), N = up to 256 output channels, K = 9 taps x Cin.
//     The A operand is ONE TMA box with halo per 64-channel chunk -- the 4-D tensor map zero-fills what falls
//     outside the image, which is the convolution's padding -- and each tap is the same shared-memory tile
//     addressed through a UMMA descriptor whose start is shifted by whole 128-byte pixel rows
//     (SWIZZLE_128B is a function of the absolute shared-memory address, so a row-shifted start reads exactly
//     what TMA wrote; tools/desc_probe.cu is the hardware check).  W = 128: one [3 x 130 pixel] box serves all 9
//     taps; W < 128: one [(TH + 2) x W] box per horizontal tap serves the 3 vertical taps.  So the activation
//     tile crosses L2 -> SMEM 1.0 / 1.5 - 3 times instead of 9.
//     Weights stream as [N x 64] K-major tiles per (tap, chunk).  Persistent CTAs, double-buffered TMEM
//     accumulator: the epilogue of tile i (TMEM -> bf16 -> swizzled staging -> TMA store, BatchNorm partial sums)
//     overlaps the main loop of tile i + 1.
//   weight gradient:  dw[tap, ci, co] = sum_p x[p + off(tap), ci] * dy[p, co]
//     contraction over pixels: both operands MN-major (pixel rows = K), the tap again a row shift of one x tile
//     with halo; 3 taps x 128 ci x 128 co accumulate in TMEM over the CTA's pixel range (split-K), then
//     red.global.add.v4.f32 into an fp32 [9][Cin][Cout] workspace; a small kernel permutes to the parameter layout.
#include <cstdlib>

#include "common.cuh"
#include "det_reduce.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

namespace mu {

constexpr int kConvThreads = 192;   // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue
constexpr int kMaxRing = 8;
constexpr int kSmemLimit = 232448;  // 227 KB

enum { CONV_W128 = 0, CONV_ROWS = 1, CONV_1X1 = 2 };

struct ConvArgs {
  int B, H, W, K, N;       // K = contraction channels, N = output channels of this GEMM
  int TH, tiles_y, m_tiles, n_tiles;
  int mode, groups, taps;  // A loads per 64-channel chunk, taps served by each
  int a_bytes, a_stride, SA, SB, SO;  // ring geometry; SO = output staging buffers (1 or 2)
  int accum;               // 1: the output tiles are ADDED into y by the TMA unit (y already holds an addend)
};

__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void red_add_v4f(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ============================================================================ forward / data gradient
// One CTA step ("super tile") = MT vertically adjacent 128-pixel tiles of one image x NT output channels: every
// weight tile fetched from L2 feeds MT MMAs, and the activation box with halo is shared by the MT tiles.
//   TMEM: NBUF accumulator sets of MT * NT fp32 columns (NBUF = 2 when 2 * MT * NT <= 512).
// PAIR = true: CTA pairs (cta_group::2).  The two CTAs of a cluster work on two adjacent super tiles of the same output
// channel block with ONE MMA of M = 256 per step: each CTA loads its own activation box and only HALF of every weight
// tile (NT / 2 rows), the tensor cores read the other half from the partner's shared memory.  The forward / data
// gradient kernels were bound by the L2 -> SM stream (6.5 TB/s, 69 % of it weights at 128 -> 128 channels:
// profiles/r01_conv_128to128_128x128_ncu.txt) and, for NT <= 128, by the shared-memory port (A 4 KB + B 4 KB per
// 64-cycle MMA); the pair halves the weight bytes on both.  The leader (cluster rank 0) issues the MMAs and owns the
// full / accumulator-free barriers; TMA loads of both CTAs complete on the leader's barriers, MMA completion is
// committed to the barriers of both.
// RES = true: the weight tiles of ALL taps and channel chunks stay resident in shared memory (one output channel block,
// 9 x ceil(K / 64) tiles of NT x 64: 72 KB for 64 -> 64 channels) instead of streaming through the ring once per super
// tile: at 64 output channels the weights were half of the L2 -> SM bytes of a super tile and the kernel sat on that
// stream (23.6 B/clk/SM = 6.4 TB/s, the same ceiling the 128 -> 128 capture shows).
template <int NT, int MT, int MODE, bool PAIR = false, bool RES = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_fprop_sm100_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                        const __grid_constant__ CUtensorMap tmap_y, float* __restrict__ stats,
                        const float* __restrict__ bias, const ConvArgs p, const DetCtx det) {
  constexpr int kBRows = PAIR ? NT / 2 : NT;       // weight rows of a tile in THIS CTA's shared memory
  constexpr int kBBytes = kBRows * 128;
  constexpr int GROUPS = MODE == CONV_ROWS ? 3 : 1;
  constexpr int TAPS = MODE == CONV_W128 ? 9 : (MODE == CONV_ROWS ? 3 : 1);
  constexpr int NBUF = (2 * MT * NT <= 512) ? 2 : 1;
  constexpr int kTmemUsed = NBUF * MT * NT;
  constexpr int kTmemCols = kTmemUsed <= 32 ? 32 : kTmemUsed <= 64 ? 64 : kTmemUsed <= 128 ? 128 : kTmemUsed <= 256 ? 256 : 512;
  constexpr int kBlocks = (NT + 63) / 64;          // 64-channel output blocks per tile (the last may be half)
  static_assert(kTmemUsed <= 512, "TMEM overflow");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + p.SA * p.a_stride;
  uint8_t* sO = sB + p.SB * kBBytes;
  float* s_stats = reinterpret_cast<float*>(sO + p.SO * 16384);       // [2][512]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stats + 1024);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + kMaxRing;
  uint64_t* b_full = bars + 2 * kMaxRing;
  uint64_t* b_empty = bars + 3 * kMaxRing;
  uint64_t* acc_full = bars + 4 * kMaxRing;       // 2
  uint64_t* acc_empty = acc_full + 2;             // 2 (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  // warp index through a shuffle: the compiler then knows the role branches are warp-uniform and keeps the MMA
  // issuer's address arithmetic on the uniform datapath (no per-MMA ELECT / R2UR.BROADCAST loops)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const bool leader = rank == 0;
  // work units: a super tile (PAIR: two adjacent super tiles, one per CTA of the pair) x an output channel block
  const int m_units = PAIR ? p.m_tiles / 2 : p.m_tiles;   // m_tiles counts super tiles
  const int total = m_units * p.n_tiles;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int kchunks = (p.K + 63) / 64;            // a partial last chunk is zero-filled by TMA (1x1 heads)

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxRing; ++i) {
      mbar_init(a_full + i, 1);
      mbar_init(a_empty + i, 1);
      mbar_init(b_full + i, 1);
      mbar_init(b_empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, PAIR ? 256 : 128);   // PAIR: the epilogue warps of both CTAs arrive on the leader's
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_y);
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_2sm<kTmemCols>(tmem_slot);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc<kTmemCols>(tmem_slot);
      tmem_relinquish();
    }
  }
  for (int i = threadIdx.x; i < 1024; i += kConvThreads) s_stats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                   // the partner's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Ring protocol: consumer waits full[s] with its phase bit, producer waits empty[s] with phase ^ 1 (a fresh
  // barrier passes a parity-1 wait), both flip the bit when the stage index wraps.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane_id() == 0) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      const int SA = p.SA, SB = p.SB, a_stride = p.a_stride, a_bytes = p.a_bytes;
      if (RES && unit0 < total) {                        // every weight tile once: [chunk][tap] -> slot ch * 9 + tap
        const int ntaps = GROUPS * TAPS;
        if (!PAIR || leader) mbar_expect_tx(b_full, (PAIR ? 2 : 1) * kchunks * ntaps * kBBytes);
        for (int ch = 0; ch < kchunks; ++ch)
          for (int tap = 0; tap < ntaps; ++tap) {
            if (PAIR) tma_load_3d_2sm(sB + (ch * ntaps + tap) * kBBytes, &tmap_w, b_full, ch * 64, rank * kBRows, tap);
            else tma_load_3d(sB + (ch * ntaps + tap) * kBBytes, &tmap_w, b_full, ch * 64, 0, tap);
          }
      }
      for (int unit = unit0; unit < total; unit += unit_step) {
        const int nt = unit / m_units, mt = (unit - nt * m_units) * (PAIR ? 2 : 1) + rank;
        const int b = mt / p.tiles_y, y0 = (mt - b * p.tiles_y) * (MT * p.TH);
        const int n0 = nt * NT + rank * kBRows;          // PAIR: this CTA's half of the weight rows
        for (int ch = 0; ch < kchunks; ++ch) {
#pragma unroll
          for (int g = 0; g < GROUPS; ++g) {
            mbar_wait_relaxed(a_empty + sa, pa ^ 1);
            const int cx = MODE == CONV_W128 ? -1 : (MODE == CONV_ROWS ? g - 1 : 0);
            const int cy = MODE == CONV_1X1 ? y0 : y0 - 1;
            if (PAIR) {   // both loads complete on the leader's barrier, which expects the bytes of both
              if (leader) mbar_expect_tx(a_full + sa, 2 * a_bytes);
              tma_load_4d_2sm(sA + sa * a_stride, &tmap_x, a_full + sa, ch * 64, cx, cy, b);
            } else {
              mbar_expect_tx(a_full + sa, a_bytes);
              tma_load_4d(sA + sa * a_stride, &tmap_x, a_full + sa, ch * 64, cx, cy, b);
            }
            if (++sa == SA) { sa = 0; pa ^= 1; }
#pragma unroll
            for (int j = 0; j < TAPS; ++j) {
              if (RES) continue;
              const int tap = MODE == CONV_ROWS ? j * 3 + g : j;
              mbar_wait_relaxed(b_empty + sb, pb ^ 1);
              if (PAIR) {
                if (leader) mbar_expect_tx(b_full + sb, 2 * kBBytes);
                tma_load_3d_2sm(sB + sb * kBBytes, &tmap_w, b_full + sb, ch * 64, n0, tap);
              } else {
                mbar_expect_tx(b_full + sb, kBBytes);
                tma_load_3d(sB + sb * kBBytes, &tmap_w, b_full + sb, ch * 64, n0, tap);
              }
              if (++sb == SB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp walks the loop,
    // one elected lane issues)
    if (!PAIR || leader) {
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, NT, 0, 0);
      constexpr uint32_t hi = desc_hi_sbo(1024);
      const uint32_t a_base = desc_lo(smem_u32(sA)), b_base = desc_lo(smem_u32(sB));
      const int SA = p.SA, SB = p.SB;
      const uint32_t a_stride16 = p.a_stride >> 4;
      constexpr uint32_t tile_pitch16 = (MODE == CONV_W128 ? 130 : 128) * 8;   // (bytes >> 4) between the MT tiles' rows
      const uint32_t row_w16 = p.W * 8;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, buf = 0, pacc = 0;
      if (RES && unit0 < total) {
        mbar_wait(b_full, 0);                            // the resident weight tiles have landed
        tc_fence_after();
      }
      for (int unit = unit0; unit < total; unit += unit_step) {
        mbar_wait(acc_empty + buf, pacc ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (MT * NT);
        uint32_t accumulate = 0;
        for (int ch = 0; ch < kchunks; ++ch) {
          const int kk_n = min(4, (p.K - ch * 64 + 15) / 16);     // 16-channel MMA steps with data in this chunk
#pragma unroll
          for (int g = 0; g < GROUPS; ++g) {
            mbar_wait(a_full + sa, pa);
            if (RES) tc_fence_after();
            const uint32_t a_tile = a_base + sa * a_stride16;
#pragma unroll
            for (int j = 0; j < TAPS; ++j) {
              const uint32_t row_off = MODE == CONV_W128 ? ((j / 3) * 130 + (j % 3)) * 8
                                                         : (MODE == CONV_ROWS ? j * row_w16 : 0);
              if (!RES) {
                mbar_wait(b_full + sb, pb);
                tc_fence_after();
              }
              const int res_tap = MODE == CONV_ROWS ? j * 3 + g : j;
              const uint32_t b_lo = b_base + (RES ? (uint32_t)(ch * (GROUPS * TAPS) + res_tap) : sb) * (kBBytes >> 4);
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                const uint32_t a_lo = a_tile + row_off + m * tile_pitch16;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  if (kk < kk_n && elect_one()) {
                    if (PAIR) umma_ss_lo_2sm(d_tmem + m * NT, a_lo + 2 * kk, b_lo + 2 * kk, hi, idesc, kk > 0 ? 1u : accumulate);
                    else umma_ss_lo(d_tmem + m * NT, a_lo + 2 * kk, b_lo + 2 * kk, hi, idesc, kk > 0 ? 1u : accumulate);
                  }
              }
              accumulate = 1;
              if (!RES) {
                if (elect_one()) {
                  if (PAIR) umma_commit_2sm(b_empty + sb, 3); else umma_commit(b_empty + sb);
                }
                if (++sb == SB) { sb = 0; pb ^= 1; }
              }
            }
            if (elect_one()) {
              if (PAIR) umma_commit_2sm(a_empty + sa, 3); else umma_commit(a_empty + sa);
            }
            if (++sa == SA) { sa = 0; pa ^= 1; }
          }
        }
        if (elect_one()) {
          if (PAIR) umma_commit_2sm(acc_full + buf, 3); else umma_commit(acc_full + buf);
        }
        if (++buf == NBUF) { buf = 0; pacc ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int e = threadIdx.x - 64;
    const int quad = warp & 3;
    const int r = quad * 32 + (int)lane_id();              // pixel row of the tile == TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t so_base = smem_u32(sO);
    // (NT = 32 / 160, the 1x1 heads: the last 64-channel block is half wide -- the channel pairs past NT are skipped)
    const int sq = e >> 5, cp = e & 31;                    // statistics: rows sq*32.., channel pair cp
    const int SO = p.SO;
    uint32_t buf = 0, pacc = 0, so_idx = 0;
    uint32_t v[32];
    // BatchNorm partial sums of this thread's channel pair (cp) and row group (sq), one set per 64-channel block
    float acc[kBlocks][4];
#pragma unroll
    for (int i = 0; i < kBlocks; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    int acc_nt = -1;
    auto flush_stats = [&](int nt_done) {
      if (nt_done < 0) return;
      if (det.on()) {
        // deterministic mode: the four row groups add into the CTA totals one after the other
        for (int round = 0; round < 4; ++round) {
          if (sq == round) {
#pragma unroll
            for (int i = 0; i < kBlocks; ++i) {
              if (i * 64 + 2 * cp >= NT) continue;
              const int chn = nt_done * NT + i * 64 + 2 * cp;
              s_stats[chn] += acc[i][0];
              s_stats[chn + 1] += acc[i][1];
              s_stats[512 + chn] += acc[i][2];
              s_stats[512 + chn + 1] += acc[i][3];
              acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            }
          }
          named_bar_sync(1, 128);
        }
        return;
      }
#pragma unroll
      for (int i = 0; i < kBlocks; ++i) {
        if (i * 64 + 2 * cp >= NT) continue;
        const int chn = nt_done * NT + i * 64 + 2 * cp;
        atomicAdd(s_stats + chn, acc[i][0]);
        atomicAdd(s_stats + chn + 1, acc[i][1]);
        atomicAdd(s_stats + 512 + chn, acc[i][2]);
        atomicAdd(s_stats + 512 + chn + 1, acc[i][3]);
        acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      }
    };
    for (int unit = unit0; unit < total; unit += unit_step) {
      const int nt = unit / m_units, mt = (unit - nt * m_units) * (PAIR ? 2 : 1) + rank;
      const int b = mt / p.tiles_y, y0 = (mt - b * p.tiles_y) * (MT * p.TH);
      mbar_wait_relaxed(acc_full + buf, pacc);   // (a main loop long: sleep between polls, the step is power-capped)
      tc_fence_after();
      if (stats != nullptr && nt != acc_nt) {             // (rare: at most once per CTA) new channel range
        flush_stats(acc_nt);
        acc_nt = nt;
      }
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
#pragma unroll
      for (int cb = 0; cb < kBlocks; ++cb) {
        const int blk = m * kBlocks + cb;
        const uint32_t so = so_base + so_idx * 16384;
        if (e == 0) {                                      // the store that last read this staging buffer is done
          if (SO == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
        }
        named_bar_sync(1, 128);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (cb * 64 + h * 32 < NT) {                     // NT = 32 / 160: the last block is only half wide
            tmem_ld32(lane_base + buf * (MT * NT) + m * NT + cb * 64 + h * 32, v);
            tmem_wait_ld();
            if (bias != nullptr) {
              const float* bp = bias + nt * NT + cb * 64 + h * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __ldg(bp + i));
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t c = h * 4 + g;
              st_shared_v4(so + r * 128 + ((c ^ (r & 7)) << 4),
                           pack_bf16(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1])),
                           pack_bf16(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                           pack_bf16(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                           pack_bf16(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
            }
          }
        }
        if (blk == MT * kBlocks - 1) {                     // accumulator drained: the next tile's MMAs may start
          tc_fence_before();
          if (PAIR) mbar_arrive_leader(acc_empty + buf); else mbar_arrive(acc_empty + buf);
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (e == 0) {
          if (p.accum) tma_reduce_add_4d(&tmap_y, so, nt * NT + cb * 64, 0, y0 + m * p.TH, b);
          else tma_store_4d(&tmap_y, so, nt * NT + cb * 64, 0, y0 + m * p.TH, b);
          tma_store_commit();
        }
        if (stats != nullptr && cb * 64 + 2 * cp < NT) {
          // BatchNorm partial sums of the ROUNDED outputs (what the normalisation will read back)
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const int rr = sq * 32 + i;
            const uint32_t w = ld_shared_u32(so + rr * 128 + ((((uint32_t)cp >> 2) ^ (rr & 7)) << 4) + (cp & 3) * 4);
            const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
            s0 += lo;
            s1 += hi;
            q0 = fmaf(lo, lo, q0);
            q1 = fmaf(hi, hi, q1);
          }
          // registers, not shared-memory atomics: float atomics on shared memory are CAS loops, and four row
          // groups hitting every address made this block cost ~1900 cycles per 64 channels
          acc[cb][0] += s0;
          acc[cb][1] += s1;
          acc[cb][2] += q0;
          acc[cb][3] += q1;
        }
        if (++so_idx == (uint32_t)SO) so_idx = 0;
      }
      }
      if (++buf == NBUF) { buf = 0; pacc ^= 1; }
    }
    if (e == 0) tma_store_wait<0>();
    if (stats != nullptr) {
      flush_stats(acc_nt);
      named_bar_sync(1, 128);
      if (det.on()) {                        // every CTA stores its totals, the last one adds them in CTA order
        float* part = det.partial + (size_t)blockIdx.x * 2 * p.N;
        for (int i = e; i < p.N; i += 128) {
          part[i] = s_stats[i];
          part[p.N + i] = s_stats[512 + i];
        }
        det_finish(det, gridDim.x, gridDim.x, 2, p.N, p.N, stats, stats + p.N, e, 128, SyncNamed<1, 128>());
      } else {
        for (int i = e; i < p.N; i += 128) {
          const float a = s_stats[i], q = s_stats[512 + i];
          if (a != 0.f || q != 0.f) {
            atomicAdd(stats + i, a);
            atomicAdd(stats + p.N + i, q);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                   // nobody leaves while the partner may still signal its barriers
  if (warp == 1) {
    __syncwarp();
    if (PAIR) tmem_dealloc_2sm<kTmemCols>(tmem_base); else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ============================================================================ weight gradient
struct WgradArgs {
  int B, H, W, Cin, Cout;
  int TH, tiles_y, m_tiles;
  int mode;                 // CONV_W128: tap group = ky, taps along kx (row step 1); CONV_ROWS: group = kx, step W rows
  int x_bytes, x_stride;    // one 64-channel x tile with halo
  int ci_blocks;            // 64-channel x blocks per CTA: 2 (Cin % 128 == 0) or 1 (two taps share an M tile)
  int n_cb, n_nb;           // ci / co blocks over the grid
  int S, chunks_per_cta;
  long slice_floats;        // deterministic mode: every pixel split (blockIdx.y) STORES into its own [taps][Cin][Cout] slice
  int ones_off;             // 1x1 convolution with 64 input channels: byte offset of an all-ones [128 px][128 B] tile (0: none)
};

template <int NB>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_wgrad_sm100_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                        float* __restrict__ ws, const WgradArgs p, float* __restrict__ db) {
  // db != NULL (1x1 heads with 64 input channels, free-running mode): the bias gradient, the column sums of dy, comes out
  // of the same GEMM.  The M = 128 tile of such a convolution has only 64 real rows (input channels); the second 64-row
  // block of the MN-major A operand is addressed through the descriptor's LBO, which here points at a constant all-ones
  // tile: rows 64..127 of the accumulator are then  sum over pixels of 1 * dy[p, co] = db[co]  -- no extra MMA, and
  // the separate pass over dy (1.3 GB for the 160-channel logits gradient of 256 images) disappears.
  constexpr int kDyBlocks = (NB + 63) / 64;
  constexpr int kDyBytes = kDyBlocks * 16384;
  constexpr int kTmemCols = NB == 128 ? 512 : 256;       // 3 taps x 128 | 3 x 64 | 1 x 160 (1x1 mode) | 1 x 32
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int x_tile_bytes = p.ci_blocks * p.x_stride;
  const int stage_bytes = x_tile_bytes + kDyBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.S * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxRing;
  uint64_t* acc_full = bars + 2 * kMaxRing;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler

  int combo = blockIdx.x;
  const int g = p.mode == CONV_1X1 ? 0 : combo % 3;
  if (p.mode != CONV_1X1) combo /= 3;
  const int cb = combo % p.n_cb, nb = combo / p.n_cb;
  const int c_begin = blockIdx.y * p.chunks_per_cta;
  const int c_end = min(p.m_tiles, c_begin + p.chunks_per_cta);
  const int nchunks = c_end - c_begin;
  if (nchunks <= 0) return;
  const int row_step = p.mode == CONV_W128 ? 1 : (p.mode == CONV_ROWS ? p.W : 0);
  const int ci0 = cb * p.ci_blocks * 64;
  // M tiles of this CTA: 3 taps (128 input channels each), 2 (64 channels: taps 0|1 paired, tap 2), 1 for a 1x1 conv
  const int m_tiles_cta = p.mode == CONV_1X1 ? 1 : (p.ci_blocks == 2 ? 3 : 2);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxRing; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dy);
  }
  if (warp == 1) {
    tmem_alloc<kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  const bool fold_db = db != nullptr && p.ones_off != 0;
  if (fold_db) {
    const uint32_t ones = smem_u32(smem) + p.ones_off;
    for (int i = threadIdx.x; i < 16384 / 16; i += kConvThreads)
      st_shared_v4(ones + i * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);      // bf16 1.0 pairs
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane_id() == 0) {
      for (int j = 0; j < nchunks; ++j) {
        const int st = j % p.S, use = j / p.S;
        if (use > 0) mbar_wait_relaxed(empty + st, (use - 1) & 1);
        const int mt = c_begin + j;
        const int b = mt / p.tiles_y, y0 = (mt - b * p.tiles_y) * p.TH;
        uint8_t* sX = smem + st * stage_bytes;
        uint8_t* sD = sX + x_tile_bytes;
        mbar_expect_tx(full + st, p.ci_blocks * p.x_bytes + kDyBytes);
        const int cx = p.mode == CONV_W128 ? -1 : (p.mode == CONV_ROWS ? g - 1 : 0);
        const int cy = p.mode == CONV_W128 ? y0 + g - 1 : (p.mode == CONV_ROWS ? y0 - 1 : y0);
        for (int blk = 0; blk < p.ci_blocks; ++blk)
          tma_load_4d(sX + blk * p.x_stride, &tmap_x, full + st, ci0 + blk * 64, cx, cy, b);
        for (int blk = 0; blk < kDyBlocks; ++blk)     // channels past Cout (NB = 160: 160..191) arrive as zeros
          tma_load_4d(sD + blk * 16384, &tmap_dy, full + st, nb * NB + blk * 64, 0, y0, b);
      }
    }
  } else if (warp == 1) {
    {   // whole warp walks the loop, one elected lane issues (see conv_fprop_sm100_kernel)
      constexpr uint32_t idesc = make_idesc_bf16(128, NB, 1, 1);   // both operands MN-major (contraction over pixels)
      constexpr uint32_t hi = desc_hi_sbo(1024);
      const int S = p.S;
      const uint32_t x_base = smem_u32(smem), stage16 = (uint32_t)stage_bytes >> 4;
      // per M tile: start offset (>> 4) inside the x tile and LBO of the A operand; all loop-invariant
      uint32_t a_off[3], a_lbo[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        if (p.ci_blocks == 2) {
          a_off[m] = (uint32_t)(m * row_step * 128) >> 4;        // tap m, 128 input channels = two 64-wide blocks
          a_lbo[m] = (uint32_t)p.x_stride;
        } else if (m == 0) {
          a_off[m] = 0;                                          // taps 0 and 1 of 64 input channels side by side
          a_lbo[m] = (uint32_t)(row_step * 128);                 // (1x1: row_step = 0, rows 64..127 duplicate, ignored)
        } else {
          a_off[m] = (uint32_t)(2 * row_step * 128) >> 4;        // tap 2; rows 64..127 of the tile are ignored
          a_lbo[m] = 0;
        }
        a_lbo[m] = ((a_lbo[m] >> 4) & 0x3FFFu) << 16;
      }
      const uint32_t b_lbo = ((16384u >> 4) & 0x3FFFu) << 16;
      uint32_t st = 0, ph = 0;
      for (int j = 0; j < nchunks; ++j) {
        mbar_wait(full + st, ph);
        tc_fence_after();
        const uint32_t x_lo = desc_lo(x_base) + st * stage16;
        const uint32_t d_lo = (x_lo + ((uint32_t)x_tile_bytes >> 4)) | b_lbo;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          if (m < m_tiles_cta) {
            uint32_t a_lo = (x_lo + a_off[m]) | a_lbo[m];
            if (fold_db)          // second 64-row block of the A operand = the ones tile (distance from THIS stage's x tile)
              a_lo = x_lo | (((((uint32_t)p.ones_off - st * (uint32_t)stage_bytes) >> 4) & 0x3FFFu) << 16);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (elect_one())
                umma_ss_lo(tmem_base + m * NB, a_lo + kk * 128, d_lo + kk * 128, hi, idesc, (j > 0 || kk > 0) ? 1u : 0u);
          }
        }
        if (elect_one()) umma_commit(empty + st);
        if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + (int)lane_id();
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t v[32];
    mbar_wait_relaxed(acc_full, 0);      // the whole contraction long: sleep between polls
    tc_fence_after();
    for (int m = 0; m < m_tiles_cta; ++m) {
      int j, ci;
      bool valid = true;
      if (p.ci_blocks == 2) {
        j = m;
        ci = ci0 + r;
      } else {
        j = m == 0 ? (r >> 6) : 2;
        ci = ci0 + (r & 63);
        valid = (m == 0 && p.mode != CONV_1X1) || r < 64;
      }
      valid = valid && ci < p.Cin;                            // stem: 8 (3 + padding) of the 64 tile rows exist
      const int tap = p.mode == CONV_1X1 ? 0 : (p.mode == CONV_W128 ? g * 3 + j : j * 3 + g);
      float* dst = ws + (size_t)blockIdx.y * p.slice_floats + ((size_t)tap * p.Cin + ci) * p.Cout + nb * NB;
#pragma unroll
      for (int c = 0; c < NB / 32; ++c) {
        tmem_ld32(lane_base + m * NB + c * 32, v);
        tmem_wait_ld();
        if (valid) {
          if (p.slice_floats != 0) {           // deterministic mode: plain stores, the finish kernel adds the slices in order
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(dst + c * 32 + q * 4) =
                  make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                              __uint_as_float(v[4 * q + 3]));
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              red_add_v4f(dst + c * 32 + q * 4, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                          __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
          }
        } else if (fold_db && r == 64) {           // rows 64..127 all hold this CTA's share of db: row 64 adds it
#pragma unroll
          for (int q = 0; q < 8; ++q)
            red_add_v4f(db + nb * NB + c * 32 + q * 4, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                        __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ============================================================================ small helpers
// w f32 [Cout][Cin][taps] (the nn.Conv2d parameter) -> wf bf16 [taps][Cout][Kp] (forward operand; Kp = Cin rounded
// up to 64, extra columns zero) and wd bf16 [taps][Cin][Cout] with the taps reversed (the data-gradient operand).
__global__ void conv_prep_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                         __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int Kp, int taps) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * Kp) return;
  const int co = idx / Kp, ci = idx - co * Kp;
  const bool real = ci < Cin;
  const float* src = w + ((size_t)co * Cin + (real ? ci : 0)) * taps;
  for (int t = 0; t < taps; ++t) {
    const __nv_bfloat16 v = __float2bfloat16_rn(real ? src[t] : 0.f);
    wf[((size_t)t * Cout + co) * Kp + ci] = v;
    if (wd != nullptr && real) wd[((size_t)(taps - 1 - t) * Cin + ci) * Cout + co] = v;
  }
}

// ws f32 [n_slices][taps][Cin][Cout] -> dw f32 [Cout][Cin][taps], slices added in order (free-running mode: one slice
// that the CTAs reduced into with red.global.add; deterministic mode: one slice per pixel split)
__global__ void conv_wgrad_finish_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cout, int Cin,
                                         int taps, int n_slices) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // co fastest: coalesced reads
  if (idx >= Cout * Cin) return;
  const int ci = idx / Cout, co = idx - ci * Cout;
  float* dst = dw + ((size_t)co * Cin + ci) * taps;
  const size_t slice = (size_t)taps * Cin * Cout;
  for (int t = 0; t < taps; ++t) {
    const float* src = ws + ((size_t)t * Cin + ci) * Cout + co;
    float acc = src[0];
    for (int k = 1; k < n_slices; ++k) acc += src[(size_t)k * slice];
    dst[t] = acc;
  }
}

// ============================================================================ launchers
static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int conv_geometry_ok(const char* fn, int B, int H, int W, int K, int N) {
  MU_REQUIRE(B > 0 && H > 0, MU_ERR_BAD_SHAPE, "%s: bad batch / height (B=%d H=%d)", fn, B, H);
  MU_REQUIRE(W == 16 || W == 32 || W == 64 || W == 128, MU_ERR_BAD_SHAPE,
             "%s: width must be 16, 32, 64 or 128 (got %d)", fn, W);
  MU_REQUIRE(H % (128 / W) == 0, MU_ERR_BAD_SHAPE, "%s: height %d must be a multiple of %d", fn, H, 128 / W);
  MU_REQUIRE((K % 64 == 0 || (K < 64 && K % 8 == 0)) && N % 64 == 0 && K >= 8 && N >= 64 && K <= 512 && N <= 512,
             MU_ERR_BAD_SHAPE,
             "%s: channel counts must be multiples of 64 in [64, 512] (input side: also 8..56 in steps of 8; got %d -> %d)",
             fn, K, N);
  return 0;
}

// MU_CONV_PAIR (environment, A/B runs): 0 = never pair CTAs, 1 = pair when the geometry allows and the problem fills
// the pairs (default), 2 = pair whenever the geometry allows
static int conv_pair_mode() {           // (read per call: the GPU tests switch it inside one process)
  const char* e = getenv("MU_CONV_PAIR");
  return e != nullptr ? atoi(e) : 1;
}

static int conv_res_mode() {           // MU_CONV_RES=0 (environment): no resident weights (A/B runs, tests)
  const char* e = getenv("MU_CONV_RES");
  return e != nullptr ? atoi(e) : 1;
}

template <int NT, int MT, int MODE, bool PAIR = false, bool RES = false>
static int run_fprop(const void* x, const void* wt, void* y, float* stats, const float* bias, int B, int H, int W, int K,
                     int N, int taps, cudaStream_t s, int accum = 0) {
  ConvArgs p;
  p.B = B; p.H = H; p.W = W; p.K = K; p.N = N;
  p.accum = accum;
  p.TH = 128 / W;
  p.tiles_y = H / (p.TH * MT);          // super tiles per image
  p.m_tiles = B * p.tiles_y;
  p.n_tiles = N / NT;
  p.mode = MODE;
  int box_w, box_h;
  if (MODE == CONV_1X1) {
    p.groups = 1; p.taps = 1; box_w = W; box_h = p.TH * MT;
  } else if (MODE == CONV_W128) {
    p.groups = 1; p.taps = 9; box_w = 130; box_h = MT + 2;
  } else {
    p.groups = 3; p.taps = 3; box_w = W; box_h = p.TH * MT + 2;
  }
  p.a_bytes = box_w * box_h * 128;
  p.a_stride = round_up(p.a_bytes, 1024);
  constexpr int kBRows = PAIR ? NT / 2 : NT;     // weight rows per tile in one CTA's shared memory
  const int fixed = 1024 + 4096 + 512;
  p.SA = 2;
  p.SO = 2;
  p.SB = (kSmemLimit - fixed - p.SO * 16384 - p.SA * p.a_stride) / (kBRows * 128);
  if (p.SB < 3) {                        // a deep enough weight ring matters more than a second staging buffer
    p.SO = 1;
    p.SB = (kSmemLimit - fixed - p.SO * 16384 - p.SA * p.a_stride) / (kBRows * 128);
  }
  if (p.SB > kMaxRing) p.SB = kMaxRing;
  if (RES) {                             // resident weights: one slot per (chunk, tap)
    p.SB = 9 * ((K + 63) / 64);
    if (fixed + 2 * 16384 + p.SA * p.a_stride + p.SB * kBRows * 128 > kSmemLimit) p.SO = 1;
    else p.SO = 2;
    if (fixed + p.SO * 16384 + p.SA * p.a_stride + p.SB * kBRows * 128 > kSmemLimit) {
      set_error("conv_fprop: resident weights do not fit (K %d NT %d)", K, NT);
      return MU_ERR_BAD_SHAPE;
    }
  }
  if (p.SB < 2) {
    set_error("conv_fprop: shared memory budget exceeded (a_stride %d NT %d)", p.a_stride, NT);
    return MU_ERR_BAD_SHAPE;
  }
  const int smem = fixed + p.SO * 16384 + p.SA * p.a_stride + p.SB * kBRows * 128;
  CUtensorMap tx, tw, ty;
  int rc;
  if ((rc = make_tmap_bf16_nhwc(&tx, x, K, W, H, B, box_w, box_h))) return rc;
  if ((rc = make_tmap_bf16_3d(&tw, wt, round_up(K, 64), N, taps, kBRows))) return rc;   // weight rows are padded to 64 K
  if ((rc = make_tmap_bf16_nhwc(&ty, y, N, W, H, B, W, p.TH))) return rc;
  auto kern = conv_fprop_sm100_kernel<NT, MT, MODE, PAIR, RES>;
  set_max_dynamic_smem_once(kern, smem);
  int grid;
  if (PAIR) {
    const int units = (p.m_tiles / 2) * p.n_tiles, pairs = sm_count() / 2;
    grid = 2 * (units < pairs ? units : pairs);
  } else {
    const int total = p.m_tiles * p.n_tiles;
    grid = total < sm_count() ? total : sm_count();
  }
  DetCtx det{nullptr, nullptr, 0};
  if (stats != nullptr && !det_context(kDetSlotConvStats, (size_t)grid * 2 * N, &det, "conv_fprop_sm100")) return MU_ERR_WORKSPACE;
  if (PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tx, tw, ty, stats, bias, p, det);
    if (e != cudaSuccess) {
      set_error("conv_fprop_sm100 (CTA pairs): cudaLaunchKernelEx: %s", cudaGetErrorString(e));
      return (int)e;
    }
  } else {
    kern<<<grid, kConvThreads, smem, s>>>(tx, tw, ty, stats, bias, p, det);
  }
  return check_launch("conv_fprop_sm100");
}

template <int NT, int MT>
static int run_fprop_mode(const void* x, const void* wt, void* y, float* stats, int B, int H, int W, int K, int N,
                          int taps, cudaStream_t s, int accum) {
  if (taps == 1) return run_fprop<NT, MT, CONV_1X1>(x, wt, y, stats, nullptr, B, H, W, K, N, taps, s, accum);
  // CTA pairs: an even number of super tiles (adjacent ones share a pair) and enough of them to fill the pairs
  const int m_tiles = B * (H / ((128 / W) * MT));
  // (measured per layer shape at 256 images, tools/bench_conv_ours.py: pairs gain 3-13 % at widths 64 and 128, lose
  // 5-10 % at 16 x 16, mixed at 32 x 32)
  const bool pair = conv_pair_mode() != 0 && m_tiles % 2 == 0 &&
                    (conv_pair_mode() == 2 || (W >= 64 && (m_tiles / 2) * (N / NT) >= 148));   // 2: force (small test shapes)
  // resident weights: a single 64-channel output block whose 9 x K weights fit beside the activation ring
  if (NT == 64 && N == 64 && conv_res_mode() != 0) {
    const int TH = 128 / W, box_h = (W == 128) ? MT + 2 : TH * MT + 2, box_w = (W == 128) ? 130 : W;
    const int a_stride = round_up(box_w * box_h * 128, 1024);
    const int b_rows = pair ? 32 : 64;
    if (1024 + 4096 + 512 + 16384 + 2 * a_stride + 9 * ((K + 63) / 64) * b_rows * 128 <= kSmemLimit) {
      if (W == 128)
        return pair ? run_fprop<64, MT, CONV_W128, true, true>(x, wt, y, stats, nullptr, B, H, W, K, N, taps, s, accum)
                    : run_fprop<64, MT, CONV_W128, false, true>(x, wt, y, stats, nullptr, B, H, W, K, N, taps, s, accum);
      return pair ? run_fprop<64, MT, CONV_ROWS, true, true>(x, wt, y, stats, nullptr, B, H, W, K, N, taps, s, accum)
                  : run_fprop<64, MT, CONV_ROWS, false, true>(x, wt, y, stats, nullptr, B, H, W, K, N, taps, s, accum);
    }
  }
  if (W == 128)
    return pair ? run_fprop<NT, MT, CONV_W128, true>(x, wt, y, stats, nullptr, B, H, W, K, N, taps, s, accum)
                : run_fprop<NT, MT, CONV_W128>(x, wt, y, stats, nullptr, B, H, W, K, N, taps, s, accum);
  return pair ? run_fprop<NT, MT, CONV_ROWS, true>(x, wt, y, stats, nullptr, B, H, W, K, N, taps, s, accum)
              : run_fprop<NT, MT, CONV_ROWS>(x, wt, y, stats, nullptr, B, H, W, K, N, taps, s, accum);
}

// y [B,H,W,N] = conv(x [B,H,W,K], wt [taps][N][K]); stats (optional) f32 [2N] += (sum, sum of squares) per channel
int launch_conv_fprop_sm100(const void* x, const void* wt, void* y, float* stats, int B, int H, int W, int K, int N,
                            int taps, cudaStream_t s, int accum) {
  int rc;
  if ((rc = conv_geometry_ok("conv_fprop_sm100", B, H, W, K, N))) return rc;
  MU_REQUIRE(taps == 9 || taps == 1, MU_ERR_BAD_SHAPE, "conv_fprop_sm100: taps must be 9 or 1 (got %d)", taps);
  const bool pair = H % (2 * (128 / W)) == 0;       // two vertically adjacent tiles per CTA step
  if (N % 256 == 0)
    return pair ? run_fprop_mode<256, 2>(x, wt, y, stats, B, H, W, K, N, taps, s, accum)
                : run_fprop_mode<256, 1>(x, wt, y, stats, B, H, W, K, N, taps, s, accum);
  if (N % 128 == 0)
    return pair ? run_fprop_mode<128, 2>(x, wt, y, stats, B, H, W, K, N, taps, s, accum)
                : run_fprop_mode<128, 1>(x, wt, y, stats, B, H, W, K, N, taps, s, accum);
  return pair ? run_fprop_mode<64, 2>(x, wt, y, stats, B, H, W, K, N, taps, s, accum)
              : run_fprop_mode<64, 1>(x, wt, y, stats, B, H, W, K, N, taps, s, accum);
}

// ---- K12: 1x1 convolution heads (nn.Conv2d(64, c_out, kernel_size=1), ade_semantic.py:284; the embedding head
// city_instance.py:248).  Output channels are padded to Np in {32, 64, 128, 160, 256} (a single N tile; the
// activation buffers carry Np channels, the extra ones are zero), K may be any multiple of 8 (TMA zero-fills the
// last 64-channel chunk): this is what lets the 150-class logits live in a 320-byte-pitch, TMA-addressable buffer.
static int conv1x1_geometry_ok(const char* fn, int B, int H, int W, int K, int Np) {
  MU_REQUIRE(B > 0 && H > 0, MU_ERR_BAD_SHAPE, "%s: bad batch / height (B=%d H=%d)", fn, B, H);
  MU_REQUIRE(W == 16 || W == 32 || W == 64 || W == 128, MU_ERR_BAD_SHAPE,
             "%s: width must be 16, 32, 64 or 128 (got %d)", fn, W);
  MU_REQUIRE(H % (128 / W) == 0, MU_ERR_BAD_SHAPE, "%s: height %d must be a multiple of %d", fn, H, 128 / W);
  MU_REQUIRE(K % 8 == 0 && K >= 8 && K <= 512, MU_ERR_BAD_SHAPE, "%s: input channels must be a multiple of 8 in [8, 512] (got %d)", fn, K);
  MU_REQUIRE(Np == 32 || Np == 64 || Np == 128 || Np == 160 || Np == 256, MU_ERR_BAD_SHAPE,
             "%s: padded output channels must be 32, 64, 128, 160 or 256 (got %d)", fn, Np);
  return 0;
}

template <int NT>
static int run_conv1x1(const void* x, const void* wt, const float* bias, void* y, float* stats, int B, int H, int W, int K,
                       cudaStream_t s) {
  const bool pair = H % (2 * (128 / W)) == 0;
  return pair ? run_fprop<NT, 2, CONV_1X1>(x, wt, y, stats, bias, B, H, W, K, NT, 1, s)
              : run_fprop<NT, 1, CONV_1X1>(x, wt, y, stats, bias, B, H, W, K, NT, 1, s);
}

// y [B,H,W,Np] = x [B,H,W,K] . wt[Np][Kpad]^T + bias[Np]
// stats (optional): f32 [2 Np] += per-channel (sum, sum of squares) of the rounded outputs, as in launch_conv_fprop_sm100
int launch_conv1x1_fprop_sm100(const void* x, const void* wt, const float* bias, void* y, int B, int H, int W, int K,
                               int Np, cudaStream_t s, float* stats) {
  int rc;
  if ((rc = conv1x1_geometry_ok("conv1x1_fprop_sm100", B, H, W, K, Np))) return rc;
  switch (Np) {
    case 32: return run_conv1x1<32>(x, wt, bias, y, stats, B, H, W, K, s);
    case 64: return run_conv1x1<64>(x, wt, bias, y, stats, B, H, W, K, s);
    case 128: return run_conv1x1<128>(x, wt, bias, y, stats, B, H, W, K, s);
    case 160: return run_conv1x1<160>(x, wt, bias, y, stats, B, H, W, K, s);
    default: return run_conv1x1<256>(x, wt, bias, y, stats, B, H, W, K, s);
  }
}

// *ws_io: in = the caller's [taps][Cin][Cout] workspace; out = where the partials are (deterministic mode: the
// registered scratch, *n_slices of them; free-running: the caller's workspace, cleared here, one slice)
template <int NB>
static int run_wgrad(const void* x, const void* dy, float** ws_io, int* n_slices, int B, int H, int W, int Cin, int Cout,
                     int taps, cudaStream_t s, float* db = nullptr, int* db_done = nullptr) {
  WgradArgs p;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.TH = 128 / W;
  p.tiles_y = H / p.TH;
  p.m_tiles = B * p.tiles_y;
  int box_w, box_h;
  if (taps == 1) {
    p.mode = CONV_1X1; box_w = W; box_h = p.TH;
  } else if (W == 128) {
    p.mode = CONV_W128; box_w = 130; box_h = 1;
  } else {
    p.mode = CONV_ROWS; box_w = W; box_h = p.TH + 2;
  }
  p.x_bytes = box_w * box_h * 128;
  p.x_stride = round_up(p.x_bytes, 1024);
  p.ci_blocks = Cin % 128 == 0 ? 2 : 1;
  p.n_cb = (Cin + 64 * p.ci_blocks - 1) / (64 * p.ci_blocks);
  p.n_nb = Cout / NB;
  const int stage = p.ci_blocks * p.x_stride + ((NB + 63) / 64) * 16384;
  // bias gradient as rows 64..127 of the accumulator (see the kernel): 1x1, 64 input channels, caller wants db
  const bool db_room = db != nullptr && taps == 1 && Cin == 64 && p.ci_blocks == 1;
  const int ones_bytes = db_room ? 1024 + 16384 : 0;     // (the ones tile starts 1024 bytes past the barriers' 512)
  p.S = (kSmemLimit - 1024 - 512 - ones_bytes) / stage;
  if (p.S > kMaxRing) p.S = kMaxRing;
  if (p.S < 2) {
    set_error("conv_wgrad: shared memory budget exceeded (stage %d)", stage);
    return MU_ERR_BAD_SHAPE;
  }
  const int smem = 1024 + 512 + p.S * stage + ones_bytes;
  const int combos = (taps == 1 ? 1 : 3) * p.n_cb * p.n_nb;
  // split-K: whole waves only.  One CTA per SM is resident (shared memory), so combos * splits must not exceed a
  // multiple of the SM count by a few CTAs (297 CTAs on 148 SMs ran as three waves, the last with one CTA).
  const int sms = sm_count();
  int waves = 2;
  int splits = waves * sms / combos;
  if (splits < 1) splits = 1;
  if (splits > p.m_tiles) splits = p.m_tiles;
  p.chunks_per_cta = (p.m_tiles + splits - 1) / splits;
  splits = (p.m_tiles + p.chunks_per_cta - 1) / p.chunks_per_cta;
  const size_t slice = (size_t)taps * Cin * Cout;
  DetCtx det;
  if (!det_context(kDetSlotMisc, slice * splits, &det, "conv_wgrad_sm100")) return MU_ERR_WORKSPACE;
  float* ws = *ws_io;
  if (det.on()) {
    ws = det.partial;
    p.slice_floats = (long)slice;
    *n_slices = splits;
  } else {
    cudaMemsetAsync(ws, 0, slice * sizeof(float), s);
    p.slice_floats = 0;
    *n_slices = 1;
  }
  *ws_io = ws;
  const bool fold_db = db_room && !det.on();            // deterministic mode keeps the ordered column sums
  p.ones_off = fold_db ? p.S * stage + 1024 : 0;
  if (db_done != nullptr) *db_done = fold_db ? 1 : 0;
  if (fold_db) cudaMemsetAsync(db, 0, (size_t)Cout * sizeof(float), s);
  CUtensorMap tx, td;
  int rc;
  if ((rc = make_tmap_bf16_nhwc(&tx, x, Cin, W, H, B, box_w, box_h))) return rc;
  if ((rc = make_tmap_bf16_nhwc(&td, dy, Cout, W, H, B, W, p.TH))) return rc;
  auto kern = conv_wgrad_sm100_kernel<NB>;
  set_max_dynamic_smem_once(kern, smem);
  kern<<<dim3(combos, splits), kConvThreads, smem, s>>>(tx, td, ws, p, fold_db ? db : nullptr);
  return check_launch("conv_wgrad_sm100");
}

// dw f32 [Cout][Cin][3][3] = sum over pixels; ws f32 [9][Cin][Cout] scratch (cleared here)
int launch_conv_wgrad_sm100(const void* x, const void* dy, float* ws, float* dw, int B, int H, int W, int Cin, int Cout,
                            cudaStream_t s) {
  int rc;
  if ((rc = conv_geometry_ok("conv_wgrad_sm100", B, H, W, Cin, Cout))) return rc;
  int n_slices = 1;
  rc = Cout % 128 == 0 ? run_wgrad<128>(x, dy, &ws, &n_slices, B, H, W, Cin, Cout, 9, s)
                       : run_wgrad<64>(x, dy, &ws, &n_slices, B, H, W, Cin, Cout, 9, s);
  if (rc) return rc;
  const int n = Cin * Cout;
  conv_wgrad_finish_kernel<<<(n + 255) / 256, 256, 0, s>>>(ws, dw, Cout, Cin, 9, n_slices);
  return check_launch("conv_wgrad_finish");
}

// dw f32 [Np][Cin] of a 1x1 convolution whose output gradient dy carries Np (padded) channels; ws f32 [Cin][Np]
// db (optional) f32 [Np]: the bias gradient (column sums of dy); *db_done = 1 when this call produced it (the weight-
// gradient GEMM carries it for 64 input channels in free-running mode), 0 when the caller still has to reduce dy itself
int launch_conv1x1_wgrad_sm100(const void* x, const void* dy, float* ws, float* dw, int B, int H, int W, int Cin, int Np,
                               cudaStream_t s, float* db, int* db_done) {
  int rc;
  if ((rc = conv1x1_geometry_ok("conv1x1_wgrad_sm100", B, H, W, Cin, Np))) return rc;
  MU_REQUIRE(Cin % 64 == 0, MU_ERR_BAD_SHAPE, "conv1x1_wgrad_sm100: input channels must be a multiple of 64 (got %d)", Cin);
  int n_slices = 1;
  if (db_done != nullptr) *db_done = 0;
  switch (Np) {
    case 32: rc = run_wgrad<32>(x, dy, &ws, &n_slices, B, H, W, Cin, Np, 1, s, db, db_done); break;
    case 64: rc = run_wgrad<64>(x, dy, &ws, &n_slices, B, H, W, Cin, Np, 1, s, db, db_done); break;
    case 128: rc = run_wgrad<128>(x, dy, &ws, &n_slices, B, H, W, Cin, Np, 1, s, db, db_done); break;
    case 160: rc = run_wgrad<160>(x, dy, &ws, &n_slices, B, H, W, Cin, Np, 1, s, db, db_done); break;
    default: rc = run_wgrad<128>(x, dy, &ws, &n_slices, B, H, W, Cin, Np, 1, s); break;   // 256 = two 128-channel blocks
  }
  if (rc) return rc;
  const int n = Cin * Np;
  conv_wgrad_finish_kernel<<<(n + 255) / 256, 256, 0, s>>>(ws, dw, Np, Cin, 1, n_slices);
  return check_launch("conv1x1_wgrad_finish");
}

// w f32 [Cout][Cin] -> wf bf16 [Np][Kp = roundup(Cin, 64)] (rows >= Cout and columns >= Cin zero) and
// wd bf16 [Cin][Dp = roundup(Np, 64)] (the data-gradient operand: wd[ci][co] = w[co][ci], zero padded);
// bias f32 [Cout] -> bias_p f32 [Np] (zero padded).
__global__ void conv1x1_prep_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                    __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                                    float* __restrict__ bias_p, int Cout, int Cin, int Np, int Kp, int Dp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < Np * Kp) {
    const int co = idx / Kp, ci = idx - co * Kp;
    wf[idx] = __float2bfloat16_rn((co < Cout && ci < Cin) ? w[(size_t)co * Cin + ci] : 0.f);
  }
  if (idx < Cin * Dp) {
    const int ci = idx / Dp, co = idx - ci * Dp;
    wd[idx] = __float2bfloat16_rn(co < Cout ? w[(size_t)co * Cin + ci] : 0.f);
  }
  if (idx < Np) bias_p[idx] = (bias != nullptr && idx < Cout) ? bias[idx] : 0.f;
}

int launch_conv1x1_prep(const float* w, const float* bias, void* wf, void* wd, float* bias_p, int Cout, int Cin, int Np,
                        cudaStream_t s) {
  const int Kp = round_up(Cin, 64), Dp = round_up(Np, 64);
  int n = Np * Kp > Cin * Dp ? Np * Kp : Cin * Dp;
  if (n < Np) n = Np;
  conv1x1_prep_kernel<<<(n + 255) / 256, 256, 0, s>>>(w, bias, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd, bias_p, Cout, Cin,
                                                     Np, Kp, Dp);
  return check_launch("conv1x1_prep");
}

int launch_conv_prep_weights(const float* w, void* wf, void* wd, int Cout, int Cin, int taps, cudaStream_t s) {
  const int Kp = round_up(Cin, 64), n = Kp * Cout;
  conv_prep_weights_kernel<<<(n + 255) / 256, 256, 0, s>>>(w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd, Cout, Cin, Kp, taps);
  return check_launch("conv_prep_weights");
}

}  // namespace mu
