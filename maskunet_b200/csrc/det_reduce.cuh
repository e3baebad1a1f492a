// Deterministic mode (mu_set_deterministic): fixed-order replacements for the float atomics that sum across CTAs.
//
// Free-running, a kernel that reduces over the whole tensor (BatchNorm statistics, parameter gradients, the loss)
// lets every CTA add its partial result with atomicAdd / red.global.add: fast, but the order of the additions -- and
// with it the last bits of the result -- changes from run to run, and 39 batch-statistics BatchNorms amplify those
// bits.  In deterministic mode every CTA STORES its partial into a scratch slice, and the last
// CTA to arrive (a counter in the same scratch) adds the slices in CTA-index order.  Same kernel, one extra pass over
// n_cta x n_values floats by one CTA; no second launch.
//
// The scratch is registered once per device by the caller (mu_set_deterministic_scratch; the library never allocates
// device memory): 1 KiB of zero-initialised counters followed by the partial area.  Kernels of one stream run one
// after the other, so they share it; deterministic mode therefore assumes ONE compute stream per device.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mu {

struct DetCtx {
  float* partial;        // nullptr: free-running mode (atomics)
  unsigned* counter;     // zero on entry, reset to zero by the last CTA
  size_t floats;         // capacity of `partial`
  __host__ __device__ bool on() const { return partial != nullptr; }
};

// host (capi.cu): the registered scratch of the current device, or an all-null context when deterministic mode is off.
// `slot` selects one of 256 counters (one per kernel kind, so that back-to-back kernels never share a live counter).
// Returns false (and sets the error text) when the mode is on but no / too small a scratch is registered.
bool det_context(int slot, size_t floats_needed, DetCtx* out, const char* who);

constexpr int kDetSlotConvStats = 1, kDetSlotBn = 2, kDetSlotLn = 3, kDetSlotSampleLn = 4, kDetSlotCe = 5,
              kDetSlotQkvDb = 6, kDetSlotMisc = 7;

// Device side.  Called by ALL `nthreads` threads of the group (tid = 0 .. nthreads - 1) after each of them wrote its
// share of the partials into ctx.partial (layout [n_slices][n_rows][row_len]; a slice may be filled by several CTAs).
// sync() is the group's barrier (__syncthreads or a named barrier).  n_arrive = CTAs (groups) that call this.  In the
// last one to arrive, for row r (r = 0 -> out0, r = 1 -> out1):
//     out_r[o] = sum over slices k = 0 .. n_slices - 1, in order, of
//                sum over i = o, o + out_len, o + 2 out_len, ... < row_len of partial[k][r][i]
// (out_len < row_len folds several columns onto one output: the 150-class BatchNorm folds rows).  Plain stores: the
// outputs need no clearing.
template <typename Sync>
__device__ __forceinline__ void det_finish(const DetCtx& ctx, int n_arrive, int n_slices, int n_rows, int row_len,
                                           int out_len, float* __restrict__ out0, float* __restrict__ out1, int tid,
                                           int nthreads, Sync sync) {
  __shared__ unsigned s_det_last;
  __threadfence();                                   // this thread's partial stores are visible device-wide
  sync();
  if (tid == 0) s_det_last = (atomicAdd(ctx.counter, 1u) == (unsigned)(n_arrive - 1)) ? 1u : 0u;
  sync();
  if (s_det_last == 0u) return;
  __threadfence();                                   // acquire: every other CTA's partial is visible
  const int n_vals = n_rows * row_len;
  const volatile float* part = ctx.partial;
  for (int idx = tid; idx < n_rows * out_len; idx += nthreads) {
    const int r = idx / out_len, o = idx - r * out_len;
    float s = 0.f;
    for (int k = 0; k < n_slices; ++k) {
      const volatile float* row = part + (size_t)k * n_vals + (size_t)r * row_len;
      for (int i = o; i < row_len; i += out_len) s += row[i];
    }
    (r == 0 ? out0 : out1)[o] = s;
  }
  if (tid == 0) *ctx.counter = 0u;                   // ready for the next kernel that uses this slot
}

struct SyncThreads {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
template <int kId, int kCount>
struct SyncNamed {
  __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"n"(kId), "n"(kCount) : "memory"); }
};

}  // namespace mu
