/* maskunet_b200 -- C ABI of the B200-native Mask Attention hot path.
 *
 * One shared library (libmaskunet_b200.so, built for sm_100a only).  Plain C
 * symbols, raw device pointers and sizes, no torch types.  This is the
 * drop-in boundary described in SURVEY.md section 8(b): the reference has no
 * FFI of its own (it is pure PyTorch), so each entry point cites the reference
 * lines whose arithmetic it replaces.  All line numbers are into
 * /root/reference/code/ade20k/ade_semantic.py unless noted.
 *
 * Conventions
 *   - Caller owns every buffer (inputs, outputs, workspace); the library never
 *     allocates, frees or retains device memory.  Inputs are never modified.
 *   - All tensors contiguous in the documented layout, 16-byte aligned.
 *   - Work is enqueued on `stream`; no device synchronisation inside.
 *   - Return 0 on success, a negative mu_status for argument errors, a positive
 *     cudaError_t for launch failures.  mu_last_error() gives a thread-local message.
 *   - dtype codes: MU_F32 computes every contraction in fp32 on CUDA cores
 *     (validation mode, tolerance 1e-4 vs the reference); MU_BF16 stores
 *     activations as bf16 and runs the attention contractions on tcgen05 tensor
 *     cores with fp32 accumulation (tolerance 2e-2).  Parameters, statistics
 *     (lse, mean, rstd, delta) and parameter gradients are always fp32.
 *   - B batch, C channels (= head dim, single head), N = H*W tokens,
 *     NKP = N rounded up to a multiple of 128 (row pitch of compacted K/V).
 */
#ifndef MASKUNET_B200_H_
#define MASKUNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mu_stream_t; /* == cudaStream_t */

enum mu_dtype { MU_F32 = 0, MU_BF16 = 1 };

/* Layout of the module input x (and of dx): channel-major = NCHW as the reference holds it ([B, C, N]);
 * token-major = channels-last / NHWC ([B, N, C]), which is the reference's permuted token view (:168) in memory.
 * The CUDA-core kernels take channel-major x; the tcgen05 projection kernels take token-major bf16 x. */
enum mu_layout { MU_X_CHANNEL_MAJOR = 0, MU_X_TOKEN_MAJOR = 1,
                 /* mu_residual_ln_fwd / _bwd only: x token-major AND y / dy in the layout the channels-last network
                  * holds the module's re-viewed result in (see mu_residual_ln_fwd) */
                 MU_X_TOKEN_MAJOR_VIEW = 2 };

enum mu_status {
  MU_OK = 0,
  MU_ERR_BAD_SHAPE = -1,
  MU_ERR_BAD_DTYPE = -2,
  MU_ERR_MISALIGNED = -3,
  MU_ERR_ARCH = -4,
  MU_ERR_NULL = -5,
  MU_ERR_WORKSPACE = -6,
  MU_ERR_DRIVER = -7
};

int mu_version(void);
const char* mu_last_error(void);
/* 1 when the current device is compute capability 10.x; tcgen05 entry points refuse to run otherwise. */
int mu_device_supported(void);
/* TMA descriptor cache: tensor maps are kept per (base pointer, shape, box) -- a pure function of the
 * key, so a hit is never stale -- behind a mutex; bounded (cleared when 4096 entries are reached).  Counters since load. */
void mu_tmap_cache_stats(uint64_t* hits, uint64_t* misses, uint64_t* entries);
void mu_tmap_cache_clear(void);
/* Deterministic mode (process-wide switch, default off).  On: every kernel that sums partial results across CTAs does it
 * in a fixed order -- mu_attn_bwd adds the per-key-tile dQ partials through per-query-tile order semaphores in its
 * workspace -- so that two runs on the same inputs return the same bits. */
void mu_set_deterministic(int32_t on);
int32_t mu_get_deterministic(void);
/* Scratch of deterministic mode for the CURRENT device: caller-owned device memory (the library never allocates), at
 * least 1 MiB -- 128 MiB covers every shape of the U-Net -- whose first 1 KiB must be zero when it is registered (the
 * kernels keep it zero).  The reductions that free-running mode does with float atomics (BatchNorm statistics and
 * gradients, LayerNorm / projection / convolution weight gradients, the loss) store one partial per CTA here and add
 * them in CTA order.  One compute stream per device: the kernels of a stream share the scratch.  NULL unregisters. */
int mu_set_deterministic_scratch(void* scratch, size_t bytes);

/* K2. Mask binarisation (:179-180  `binary_mask > 0.5` -> 0 / -inf bias, per key).
 *   bits      int64 [B, N]   the torch.randint(0, 2, ...) draw of :178 (stays a torch call)
 *   keep_bits uint32 [B, ceil(N/32)]  bit n%32 of word n/32 set  <=>  key n kept (bias 0.0)
 *   n_keep    int32 [B]      number of kept keys
 *   keep_idx  int32 [B, N]   compacted position -> token index (entries >= n_keep are -1)
 *   keep_rank int32 [B, N]   token index -> compacted position, -1 when the key is masked
 * Bit-exact: keep <=> (bits > 0.5). */
int mu_mask_binarize(const int64_t* bits, int32_t B, int32_t N, uint32_t* keep_bits, int32_t* n_keep,
                     int32_t* keep_idx, int32_t* keep_rank, mu_stream_t stream);

/* K1. Q/K/V projections (:168-172).  Tokens are x[b, :, n] (the permute of :168 is never materialised).
 *   x      T   [B, C, N] (MU_X_CHANNEL_MAJOR, CUDA cores) or [B, N, C] (MU_X_TOKEN_MAJOR, MU_BF16 only, tcgen05)
 *   w_qkv  f32 [3C, C] = cat(query.weight, key.weight, value.weight);  b_qkv f32 [3C]
 *   w_qkv_lp   bf16 copy of w_qkv, required by the token-major path (the tensor-core operand), else may be NULL
 *   q      T   [B, N, C]      kc, vc  T [B, NKP, C]: rows of kept keys only, in token order
 *                             (row keep_rank[b, n]); rows [n_keep, roundup(n_keep, 128)) are zero-filled. */
int mu_qkv_project(const void* x, const float* w_qkv, const void* w_qkv_lp, const float* b_qkv,
                   const int32_t* keep_rank, const int32_t* n_keep, void* q, void* kc, void* vc, int32_t B, int32_t C,
                   int32_t N, int32_t NKP, int32_t dtype, int32_t x_layout, mu_stream_t stream);

/* K3. Masked attention forward (:174-186): O = softmax(Q Kc^T / sqrt(C)) Vc over the kept keys only, which
 * equals the reference's softmax(QK^T/sqrt(C) + mask) V exactly (masked keys contribute exp(-inf) = 0).
 *   o   T   [B, N, C]      lse f32 [B, N] = log sum_j exp(s_ij / sqrt(C))
 * MU_BF16: tcgen05/TMEM tiles fed by TMA, online softmax; MU_F32: CUDA-core fp32. */
int mu_attn_fwd(const void* q, const void* kc, const void* vc, const int32_t* n_keep, void* o, float* lse, int32_t B,
                int32_t N, int32_t NKP, int32_t C, int32_t dtype, mu_stream_t stream);

/* Diagnostic twins of mu_attn_fwd / mu_attn_bwd that always run the CUDA-core kernels, for either dtype.
 * Used by the GPU tests to cross-check the tensor-core kernels on-device at sizes the CPU oracle cannot reach. */
int mu_attn_fwd_cudacore(const void* q, const void* kc, const void* vc, const int32_t* n_keep, void* o, float* lse,
                         int32_t B, int32_t N, int32_t NKP, int32_t C, int32_t dtype, mu_stream_t stream);
int mu_attn_bwd_cudacore(const void* q, const void* kc, const void* vc, const int32_t* n_keep,
                         const int32_t* keep_idx, const void* d_o, const float* lse, const float* delta, void* dq,
                         void* dk, void* dv, int32_t B, int32_t N, int32_t NKP, int32_t C, int32_t dtype,
                         mu_stream_t stream);

/* K3 epilogue. Residual + LayerNorm over channels (:187-188), output in the [B, N, C] layout that the
 * module returns re-viewed as [B, C, H, W] (:190).
 *   y T [B, N, C] = LN_C(o + x^T) * gamma + beta;  mean, rstd f32 [B, N] saved for backward.
 * x_layout = MU_X_TOKEN_MAJOR_VIEW (MU_BF16, C in {64, 128, 256}, N % C == 0): y is written as the channels-last memory
 *   [B, H*W, C] of that re-viewed [B, C, H, W] tensor, i.e. y_view[b, p, c] = y[b].flat[c N + p] -- the transpose pass
 *   between the module and the next convolution is done inside the kernel; mu_residual_ln_bwd takes dy in the same
 *   layout.  Values are bit-identical to MU_X_TOKEN_MAJOR followed by mu_transpose. */
int mu_residual_ln_fwd(const void* o, const void* x, const float* gamma, const float* beta, float eps, void* y,
                       float* mean, float* rstd, int32_t B, int32_t C, int32_t N, int32_t dtype, int32_t x_layout,
                       mu_stream_t stream);

/* K4. Backward of the residual + LayerNorm (autograd of :187-188, implicit at :400).
 *   dy T [B, N, C];  dz T [B, N, C] = dL/d(o + x^T)  (this is both dO and the residual branch of dX^T)
 *   delta f32 [B, N] = sum_c dz * o    (softmax-backward row term)
 *   dgamma, dbeta f32 [C]: ACCUMULATED with atomics, caller zero-fills. */
int mu_residual_ln_bwd(const void* dy, const void* o, const void* x, const float* mean, const float* rstd,
                       const float* gamma, void* dz, float* delta, float* dgamma, float* dbeta, int32_t B, int32_t C,
                       int32_t N, int32_t dtype, int32_t x_layout, mu_stream_t stream);

/* K5. Masked attention backward (autograd of :174-186).  Recomputes P from q, kc, lse.
 *   dq, dk, dv T [B, N, C], all in TOKEN space: the gradient of compacted key row r is scattered back to
 *   token keep_idx[b, r]; rows of masked keys are zero (the entry point clears dk / dv itself).
 *   workspace: mu_attn_bwd_workspace_bytes(...) bytes of scratch: the order semaphores of deterministic mode and,
 *   for C >= 128, the fp32 dQ accumulator the key tiles add into (C = 64 adds bf16 partial tiles straight into dq
 *   with the TMA reduce-add and needs no accumulator); 0 bytes / NULL allowed for MU_F32. */
size_t mu_attn_bwd_workspace_bytes(int32_t B, int32_t N, int32_t C, int32_t dtype);
int mu_attn_bwd(const void* q, const void* kc, const void* vc, const int32_t* n_keep, const int32_t* keep_idx,
                const void* d_o, const float* lse, const float* delta, void* dq, void* dk, void* dv, void* workspace,
                size_t workspace_bytes, int32_t B, int32_t N, int32_t NKP, int32_t C, int32_t dtype,
                mu_stream_t stream);

/* K6. Backward of the projections and the residual branch (autograd of :168-172, :187).
 *   dq, dk, dv T [B, N, C] as produced by mu_attn_bwd (token space).
 *   dx     T   in the layout of x:  channel-major [B, C, N] = (dz + dq Wq + dk Wk + dv Wv)^T, token-major [B, N, C]
 *   dw_qkv f32 [3C, C], db_qkv f32 [3C]: ACCUMULATED with atomics, caller zero-fills. */
int mu_qkv_project_bwd(const void* x, const void* dz, const void* dq, const void* dk, const void* dv,
                       const float* w_qkv, const void* w_qkv_lp, void* dx, float* dw_qkv, float* db_qkv, int32_t B,
                       int32_t C, int32_t N, int32_t dtype, int32_t x_layout, mu_stream_t stream);

/* Batched 2-D transpose in[batch][rows][cols] -> out[batch][cols][rows] (elem_bytes 2 or 4): the NCHW <-> NHWC bridge
 * between the module's re-viewed output (:190) and channels-last convolutions. */
int mu_transpose(const void* in, void* out, int32_t batch, int32_t rows, int32_t cols, int32_t elem_bytes,
                 mu_stream_t stream);

/* K8. Training-mode BatchNorm2d fused with activation and residual on channels-last activations
 * (ConvBlock :198-210, trailing BN :219 / :240, head BN + ReLU :283-287).  x, r, y are [M = B*H*W, C] rows (NHWC).
 *   y = act(gamma * (x - mean) * rstd + beta [+ r]),  act: 0 none, 1 GELU (erf), 2 ReLU;  r may be NULL
 *   mean, rstd, a = gamma*rstd, b = beta - mean*a: f32 [C] outputs (saved for backward)
 *   running_mean / running_var (f32 [C], may be NULL) updated in place with `momentum`, unbiased variance
 *   sums: f32 [2C] scratch.  Any C with C / vec <= 256 where vec = 8 (C % 8 == 0), 2 (C even) or 1. */
int mu_bn_act_fwd(const void* x, const void* r, const float* gamma, const float* beta, float* running_mean,
                  float* running_var, float momentum, float eps, void* y, float* mean, float* rstd, float* a,
                  float* b, float* sums, int64_t M, int32_t C, int32_t act, int32_t dtype, mu_stream_t stream);
/* Inference-mode / precomputed-affine variant: y = act(a * x + b [+ r]). */
int mu_bn_act_apply(const void* x, const void* r, const float* a, const float* b, void* y, int64_t M, int32_t C,
                    int32_t act, int32_t dtype, mu_stream_t stream);
/* Backward of mu_bn_act_fwd.  sums f32 [2C] receives (dbeta, dgamma) = (sum dz, sum dz * xhat);
 * dx [M, C]; dr [M, C] = dz when r was given (else NULL). */
int mu_bn_act_bwd(const void* dy, const void* x, const void* r, const float* a, const float* b, const float* mean,
                  const float* rstd, float* sums, void* dx, void* dr, int64_t M, int32_t C, int32_t act,
                  int32_t dtype, mu_stream_t stream);

/* K9. MaxPool2d(2) on channels-last [B, H, W, C] (:216).  bwd = 0: out [B, H/2, W/2, C] = max over 2x2;
 * bwd = 1: dy [B, H/2, W/2, C] -> out = dx [B, H, W, C], gradient to the first maximum in scan order;
 * bwd = 2: the same gradient ADDED to out, which already holds the skip connection's gradient of x (:304-312). */
int mu_maxpool2(const void* x, const void* dy, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t bwd,
                int32_t dtype, mu_stream_t stream);

/* K10. Upsample(scale 2, bilinear, align_corners=True) of x [B, H, W, Cx] fused with cat([skip, up], channel)
 * (:235, :252): out [B, 2H, 2W, Cs + Cx].  Backward splits dout into dskip [B, 2H, 2W, Cs] and dx [B, H, W, Cx]. */
int mu_upsample_concat_fwd(const void* skip, const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t Cs,
                           int32_t Cx, int32_t dtype, mu_stream_t stream);
int mu_upsample_concat_bwd(const void* dout, void* dskip, void* dx, int32_t B, int32_t H, int32_t W, int32_t Cs,
                           int32_t Cx, int32_t dtype, mu_stream_t stream);

/* K11. LayerNorm over all L = C*H*W elements of each sample (nn.LayerNorm([64, 128, 128]), :281, :311).
 * x, y [B, L] in channels-last element order; gamma, beta f32 [L] in the SAME order; sums f32 [2B] scratch. */
int mu_sample_layernorm_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y, float* mean,
                            float* rstd, float* sums, int32_t B, int64_t L, int32_t dtype, mu_stream_t stream);
int mu_sample_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                            float* sums, void* dx, float* dgamma, float* dbeta, int32_t B, int64_t L, int32_t dtype,
                            mu_stream_t stream);

/* A14. nn.CrossEntropyLoss(mean, ignore_index) forward fused with its gradient (:377, :399): logits [M, C]
 * (channels-last rows), labels int64 [M], valid_count f32 [1] = number of labels != ignore_index.
 * loss_sum f32 [1] receives the mean loss; dlogits [M, C] = (softmax - onehot) / valid_count. */
int mu_cross_entropy_fused(const void* logits, const int64_t* labels, const float* valid_count, int64_t ignore_index,
                           void* dlogits, float* loss_sum, int64_t M, int32_t C, int32_t pitch, int32_t dtype,
                           mu_stream_t stream);
/* (pitch >= C: row stride of logits / dlogits in elements -- the class-padded buffers of mu_conv1x1_fwd; the
 * pad columns of dlogits are written as zeros.) */

/* K7. 3x3 convolution, stride 1, zero padding 1, no bias (nn.Conv2d(cin, cout, kernel_size=3, padding=1,
 * bias=False), :199 and :202, inside ConvBlock :192-210) on tcgen05 tensor cores.  bf16 channels-last
 * activations [B, H, W, C], fp32 accumulation; W in {16, 32, 64, 128}, H a multiple of 128 / W, channel counts
 * multiples of 64 in [64, 512].  MU_BF16 only (fp32 validation mode keeps the stock NCHW convolution).
 *
 * The 3-channel stem (:259 -> ConvBlock(3, 64)) runs on the same kernels with its input zero-padded to 8 channels
 * (Cin < 64 must be a multiple of 8; TMA zero-fills the rest of the 64-channel operand tile).
 * mu_conv_prep_weights: w f32 [Cout, Cin, taps] (the parameter as nn.Conv2d holds it, taps = 9 or 1) ->
 *   wf bf16 [taps, Cout, roundup(Cin, 64)] (forward operand, zero padded) and wd bf16 [taps, Cin, Cout] with the tap
 *   order reversed (data-gradient operand; wd may be NULL).
 * mu_conv3x3_fwd: y [B, H, W, Cout] = conv(x [B, H, W, Cin], wf).  stats (may be NULL): f32 [2 * Cout],
 *   ADDED to: per-channel sum and sum of squares of the (rounded) outputs -- the statistics pass of the
 *   training-mode BatchNorm2d that always follows (:200, :203), so the caller zeroes it first and passes it to
 *   mu_bn_act_fwd_stats.
 * mu_conv3x3_bwd_data: dx [B, H, W, Cin] = conv(dy [B, H, W, Cout], wd).
 * mu_conv3x3_bwd_data_acc: dx += conv(dy, wd): dx already holds another gradient of the same activation (the
 *   residual branch of a ConvBlock, ade_semantic.py:207: gelu(x + block(x))); the TMA unit adds every output tile into
 *   it (bf16 reduce-add, each element added exactly once: bit-reproducible), which replaces autograd's accumulation pass.
 * mu_conv3x3_bwd_weight: dw f32 [Cout, Cin, 3, 3] = sum over pixels of dy (x) shifted x; workspace of
 *   mu_conv3x3_workspace_bytes(Cin, Cout) bytes (f32 [9, Cin, Cout] split-K accumulator, cleared inside). */
int mu_conv_prep_weights(const float* w, void* wf, void* wd, int32_t Cout, int32_t Cin, int32_t taps,
                         mu_stream_t stream);
int mu_conv3x3_fwd(const void* x, const void* wf, void* y, float* stats, int32_t B, int32_t H, int32_t W, int32_t Cin,
                   int32_t Cout, int32_t dtype, mu_stream_t stream);
int mu_conv3x3_bwd_data(const void* dy, const void* wd, void* dx, int32_t B, int32_t H, int32_t W, int32_t Cin,
                        int32_t Cout, int32_t dtype, mu_stream_t stream);
int mu_conv3x3_bwd_data_acc(const void* dy, const void* wd, void* dx, int32_t B, int32_t H, int32_t W, int32_t Cin,
                            int32_t Cout, int32_t dtype, mu_stream_t stream);
size_t mu_conv3x3_workspace_bytes(int32_t Cin, int32_t Cout);
int mu_conv3x3_bwd_weight(const void* x, const void* dy, void* workspace, size_t workspace_bytes, float* dw, int32_t B,
                          int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t dtype, mu_stream_t stream);
/* nn.BatchNorm2d running-statistics update from the batch statistics mu_bn_act_fwd produced:
 * running_mean = (1 - momentum) running_mean + momentum mean; running_var likewise with the unbiased batch
 * variance (1 / rstd^2 - eps) M / (M - 1).  In place on running_mean / running_var (f32 [C]). */
int mu_bn_update_running(float* running_mean, float* running_var, const float* mean, const float* rstd, float momentum,
                         float eps, int64_t M, int32_t C, mu_stream_t stream);
/* mu_bn_act_fwd with the statistics pass already done (sums f32 [2C] = per-channel sum, sum of squares over the
 * M rows, e.g. from mu_conv3x3_fwd): finalize + apply only. */
int mu_bn_act_fwd_stats(const void* x, const void* r, const float* gamma, const float* beta, float eps, void* y,
                        float* mean, float* rstd, float* a, float* b, const float* sums, int64_t M, int32_t C,
                        int32_t act, int32_t dtype, mu_stream_t stream);

/* K12. 1x1 convolution heads (nn.Conv2d(64, c_out, kernel_size=1) with bias, :284; the embedding head
 * cityscapes/city_instance.py:248) on the same tcgen05 implicit-GEMM kernels as K7.  Output channels are padded
 * to Np in {32, 64, 128, 160, 256} so that the activation rows are TMA-addressable (150 classes -> 160: 320-byte
 * pitch): y, dy are [B, H, W, Np] with channels >= Cout equal to zero; the caller exposes the first Cout channels.
 *   mu_conv1x1_prep:       w f32 [Cout, Cin], bias f32 [Cout] or NULL -> wf bf16 [Np, roundup(Cin, 64)],
 *                          wd bf16 [Cin, roundup(Np, 64)] (data-gradient operand), bias_p f32 [Np], all zero padded
 *   mu_conv1x1_fwd:        y = x [B, H, W, Cin] . wf^T + bias_p
 *   mu_conv1x1_bwd_data:   dx [B, H, W, Cin] = dy [B, H, W, Np] . wd^T    (Cin in {32, 64, 128, 160, 256})
 *   mu_conv1x1_bwd_weight: dw f32 [Np, Cin] = dy^T x   (Cin a multiple of 64); workspace f32 [Cin, Np]
 * mu_column_sums: per-channel sum and sum of squares of x [M, C] -> sums f32 [2C] (the bias gradient of the head). */
int mu_conv1x1_prep(const float* w, const float* bias, void* wf, void* wd, float* bias_p, int32_t Cout, int32_t Cin,
                    int32_t Np, mu_stream_t stream);
int mu_conv1x1_fwd(const void* x, const void* wf, const float* bias_p, void* y, int32_t B, int32_t H, int32_t W,
                   int32_t Cin, int32_t Np, int32_t dtype, mu_stream_t stream);
/* mu_conv1x1_fwd that also ADDS the per-channel (sum, sum of squares) of its rounded outputs to stats f32 [2 * Np] (the
 * caller zeroes it): the statistics pass of the BatchNorm2d that follows the head convolution (:284-286). */
int mu_conv1x1_fwd_stats(const void* x, const void* wf, const float* bias_p, void* y, float* stats, int32_t B, int32_t H,
                         int32_t W, int32_t Cin, int32_t Np, int32_t dtype, mu_stream_t stream);
int mu_conv1x1_bwd_data(const void* dy, const void* wd, void* dx, int32_t B, int32_t H, int32_t W, int32_t Cin,
                        int32_t Np, int32_t dtype, mu_stream_t stream);
size_t mu_conv1x1_workspace_bytes(int32_t Cin, int32_t Np);
int mu_conv1x1_bwd_weight(const void* x, const void* dy, void* workspace, size_t workspace_bytes, float* dw, int32_t B,
                          int32_t H, int32_t W, int32_t Cin, int32_t Np, int32_t dtype, mu_stream_t stream);
/* mu_conv1x1_bwd_weight that also returns the bias gradient db f32 [Np] = sum over pixels of dy.  With 64 input channels
 * (every head of the U-Net) in free-running mode it is carried by the weight-gradient GEMM itself (the unused half of the
 * 128-row accumulator tile, through an all-ones operand tile); otherwise it is reduced from dy by the column-sums kernel.
 * workspace: max(mu_conv1x1_workspace_bytes(Cin, Np), 2 * Np floats). */
int mu_conv1x1_bwd_weight_bias(const void* x, const void* dy, void* workspace, size_t workspace_bytes, float* dw,
                               float* db, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Np, int32_t dtype,
                               mu_stream_t stream);
int mu_column_sums(const void* x, float* sums, int64_t M, int32_t C, int32_t dtype, mu_stream_t stream);

/* Logit post-processing on the device: the class map `argmax(softmax(y_pred / 0.5, dim=1), dim=1)`
 * (:130-131; cityscapes/city_instance.py:461-462) -- softmax is monotone, so argmax of the logits with torch's tie
 * rule (lowest index) -- and mean_iou (:128-146) without its per-class host syncs.
 * logits [M, pitch >= C] rows (channels-last, class-padded allowed); labels int64 [M] or NULL; pred int64 [M] or NULL;
 * hist int32 [3C] scratch (predicted / labelled / matching pixels per class, cleared inside); miou f32 [1] or NULL:
 * mean over classes with a non-empty union of (intersection + smooth) / (union + smooth). */
int mu_argmax_iou(const void* logits, const int64_t* labels, int64_t* pred, int32_t* hist, float* miou, int64_t M,
                  int32_t C, int32_t pitch, float smooth, int32_t dtype, mu_stream_t stream);

/* InstanceContrastiveLoss (coco/coco_panoptic.py:482-521; cityscapes/city_instance.py:279-307).
 * The caller groups the pixels by instance id (stable sort) and draws the negative ranks from the CPU generator as
 * :510 does; the device selects anchor / positive (first two pixels of the instance) and the k-th non-member pixel
 * as the negative, gathers the three logit columns sem[:, :, i0, i1] -- (i0, i1) = (batch index, row index) of the
 * pixel, the reference's own indexing at :502-503 -- and accumulates the mean TripletMarginLoss(margin, p=2, eps).
 *   sem     [B, C, H, W] with ELEMENT strides sem_strides[4], a HOST array (any layout, e.g. the class-padded
 *           channels-last view the 1x1 head returns)
 *   order   int64 [M = B*H*W]  pixel positions grouped by instance id, ascending inside a group
 *   meta    int64 [3, K]       per instance: group offset in order, pixel count (>= 2), negative rank k in [0, M - count)
 *   sel     int32 [K, 6] out   (h, w) of anchor, positive, negative; -1 where the reference would raise IndexError
 *   dist    f32 [K, 2] out     (d(a,p), d(a,n));   loss f32 [1] out: mean over the K instances (NaN if any -1)
 * Backward ACCUMULATES  scale * dloss[0] * dloss/dsem  into dsem (strides dsem_strides[4], same dtype as sem); dloss
 * may be NULL (= 1). */
int mu_instance_triplet_fwd(const void* sem, const int64_t* sem_strides, int32_t B, int32_t C, int32_t H, int32_t W,
                            const int64_t* order, const int64_t* meta, int32_t K, float margin, float eps, int32_t* sel,
                            float* dist, float* loss, int32_t dtype, mu_stream_t stream);
int mu_instance_triplet_bwd(const void* sem, const int64_t* sem_strides, int32_t B, int32_t C, const int32_t* sel,
                            int32_t K, float margin, float eps, const float* dist, const float* dloss, float scale,
                            void* dsem, const int64_t* dsem_strides, int32_t dtype, mu_stream_t stream);

/* ---- Generalised mode of the kernel sweep (BASELINE.json configs[4]) -------------------------
 * The north star names a per-query mask-logit einsum and a sigmoid > 0.5 binarisation; the reference has neither (its
 * bias is torch.randint, :177-181).  These entry points implement that generalisation for the
 * sweep; their oracle is builder-written (oracle/query_attention_oracle.py): "parity unpinned by reference".
 *
 * K13. bits = sigmoid(einsum('bqc,bnc->bqn', qe, feat)) > 0.5 on tcgen05 (bf16 in, fp32 accumulate), logits never stored.
 *   qe [B, Q, C] bf16, feat [B, N, C] bf16, C = 256;  NKP = roundup(N, 128), QP = roundup(Q, 128)
 *   bits   uint32 [B, Q, NKP/32]   bit n%32 of word n/32 set <=> query q may attend key n (bias 0.0, else -inf)
 *   bits_t uint32 [B, NKP, QP/32]  the same relation transposed (what the backward kernel reads)
 *   row_count int32 [B, Q]         kept keys per query BEFORE the rule below
 *   logits f32 [B, Q, N] or NULL   test hook
 * Binarisation is bit-exact to torch.sigmoid(x) > 0.5 in fp32 (x > 1.5 * 2^-24, not x > 0).  A query that kept no key
 * attends every key (Mask2Former's rule; a fully masked row is NaN otherwise). */
int mu_query_mask_bits(const void* qe, const void* feat, int32_t B, int32_t Q, int32_t N, int32_t C, uint32_t* bits,
                       uint32_t* bits_t, int32_t* row_count, float* logits, int32_t dtype, mu_stream_t stream);
/* Masked multi-head cross attention of Q queries over N keys, (sample, head) pairs as the batch BH = B * heads:
 *   q, o, d_o, dq [BH, Q, D]; k, v [BH, NKP, D] (rows >= N zero); dk, dv [BH, N, D]; lse, delta f32 [BH, Q]
 *   o = softmax(q k^T * scale + bias(bits[bh / heads])) v;  D = 64 (pad 32-wide heads with zeros and pass their scale).
 * tcgen05 kernels of K3 / K5 with the bias applied in registers; bf16 only. */
int mu_query_attn_fwd(const void* q, const void* k, const void* v, const uint32_t* bits, void* o, float* lse, int32_t BH,
                      int32_t heads, int32_t Q, int32_t N, int32_t NKP, int32_t D, float scale, int32_t dtype,
                      mu_stream_t stream);
size_t mu_query_attn_bwd_workspace_bytes(int32_t BH, int32_t Q, int32_t D);
int mu_query_attn_bwd(const void* q, const void* k, const void* v, const uint32_t* bits_t, const void* d_o,
                      const float* lse, const float* delta, void* dq, void* dk, void* dv, void* workspace,
                      size_t workspace_bytes, int32_t BH, int32_t heads, int32_t Q, int32_t N, int32_t NKP, int32_t D,
                      float scale, int32_t dtype, mu_stream_t stream);

/* SURVEY 8(f) rank 3, device half of the input pipeline: torchvision ToTensor() (:56-79, :97) on a uint8 batch.
 *   img uint8 [B, H, W, Cin] (HWC as cv2 / PIL deliver it);  out = img / 255 (IEEE division, bit-exact with ToTensor)
 *   channels_last = 0: out [B, Cin, H, W] (the reference's layout; Cpad ignored)
 *   channels_last = 1: out [B, H, W, Cpad], channels >= Cin zero (Cpad = 8 feeds the tcgen05 stem convolution) */
int mu_to_tensor_u8(const uint8_t* img, void* out, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cpad,
                    int32_t channels_last, int32_t dtype, mu_stream_t stream);

/* SURVEY 8(f) rank 3, the resize of the dataset classes (ade_semantic.py:72-73), bit-exact with OpenCV's uint8
 * arithmetic (opencv-python, third party: restated in oracle/resize_oracle.py, pinned against cv2 4.13 itself):
 *   mu_resize_linear_to_tensor_u8: cv2.resize(img, (out_w, out_h), INTER_LINEAR) of ONE uint8 image [src_h, src_w, Cin]
 *     followed by ToTensor when normalise = 1 (/ 255; normalise = 0 keeps the resized bytes as numbers 0..255);
 *     out is one image slot of the network input: [Cin, out_h, out_w] (channels_last = 0) or [out_h, out_w, Cpad]
 *     with channels >= Cin zero (channels_last = 1), dtype f32 / bf16.  Call once per image of a batch with the slot's
 *     pointer: source sizes differ per image.
 *   mu_resize_nearest_u8_i64: cv2.resize(mask, (out_w, out_h), INTER_NEAREST) of one uint8 label map, written as the
 *     int64 labels the loss takes (`torch.from_numpy(mask).long()`, :78). */
int mu_resize_linear_to_tensor_u8(const uint8_t* img, void* out, int32_t src_h, int32_t src_w, int32_t Cin, int32_t out_h,
                                  int32_t out_w, int32_t Cpad, int32_t channels_last, int32_t normalise, int32_t dtype,
                                  mu_stream_t stream);
int mu_resize_nearest_u8_i64(const uint8_t* mask, int64_t* out, int32_t src_h, int32_t src_w, int32_t out_h,
                             int32_t out_w, mu_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MASKUNET_B200_H_ */
