/*
 * edtr_b200 — C-ABI of the B200 (sm_100a) kernels behind the EDTR ControlLDM
 * restore path (SpacedSampler loop -> ControlNet + SD-2.1 UNet -> VAE decoder).
 *
 * The reference (JaehaKim97/EDTR) has no FFI layer: its hot path is a tree of
 * torch.nn modules calling ATen/cuDNN/cuBLAS.  Each entry point below replaces
 * one family of those library call sites; the citation after "replaces:" is the
 * reference file:line whose arithmetic the kernel reproduces.
 *
 * Conventions (all entry points):
 *   - raw device pointers + sizes only, no torch types; activations are bf16
 *     channels-last ("NHWC": a [rows, channels] matrix with a row stride `ld`
 *     given in ELEMENTS), parameters of norms / biases are fp32;
 *   - stream-ordered on `stream` (a cudaStream_t passed as void*), no allocation,
 *     no synchronisation, CUDA-Graph capturable;
 *   - return 0 on success or a negative EDTR_ERR_* code; edtr_last_error() gives
 *     a thread-local human readable message.
 */
#ifndef EDTR_B200_H_
#define EDTR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDTR_OK 0
#define EDTR_ERR_INVALID (-1) /* bad argument (shape / alignment / mode) */
#define EDTR_ERR_CUDA (-2)    /* CUDA runtime / driver error */
#define EDTR_ERR_DEVICE (-3)  /* not a compute-capability 10.x device */

/* Epilogue activation of edtr_gemm_bf16 / edtr_conv3x3_bf16. */
#define EDTR_ACT_NONE 0
#define EDTR_ACT_SILU 1  /* x * sigmoid(x)            replaces: model/unet.py:166-172 (emb SiLU) */
#define EDTR_ACT_GEGLU 2 /* x * gelu_erf(gate)         replaces: model/attention.py:20-27 */
#define EDTR_ACT_GELU 3      /* gelu_erf(x)            replaces: nn.GELU, model/swinir.py:19-31 */
#define EDTR_ACT_LRELU_02 4  /* LeakyReLU(0.2)         replaces: model/swinir.py:803, 874-880 */
#define EDTR_ACT_LRELU_001 5 /* LeakyReLU(0.01)        replaces: model/swinir.py:766-769 (nn.LeakyReLU default) */

/* Output addressing of the GEMM epilogue. */
#define EDTR_OUT_BF16 0      /* out[row*ldc + col], bf16                         */
#define EDTR_OUT_F32 1       /* out[row*ldc + col], fp32                         */
#define EDTR_OUT_NCHW_F32 2  /* out[((row/hw)*N + col)*hw + row%hw], fp32 (NCHW) */
#define EDTR_OUT_NCHW_BF16 3 /* same addressing, bf16                            */

/*
 * Fused epilogue applied to the fp32 accumulator acc[row, col]:
 *   v = alpha*acc + bias[col] + rowvec[(row / rows_per_group)*rowvec_ld + col]
 *         + residual[row*ldr + col];   v = act(v);   store(v)
 * Any of bias / rowvec / residual may be NULL.  For EDTR_ACT_GEGLU the weight
 * rows must be pre-interleaved per N-tile (first half of every tile = value
 * columns, second half = gate columns; see edtr_gemm_tile_n) and the stored
 * matrix has N/2 columns.
 *   rowvec  replaces: h + emb_out[:, :, None, None]        model/unet.py:221
 *   residual replaces: skip_connection(x) + h               model/unet.py:223,
 *            attn(x) + x / ff(x) + x                        model/attention.py:231-233,
 *            x + x_in                                       model/attention.py:302,
 *            hs.pop() + control.pop(), h += control.pop()   model/controlnet.py:31,37
 */
typedef struct EdtrEpilogue {
  const float* bias;
  const float* rowvec;
  int32_t rowvec_ld;
  int32_t rows_per_group;
  const void* residual; /* bf16 */
  int32_t ldr;
  void* out;
  int32_t ldc;
  int32_t act;
  int32_t out_mode;
  int32_t hw; /* rows per image, for the NCHW output modes */
  float alpha;
  /* ---- per-call scratch: the library keeps NO state between calls (re-entrant per device / stream) ----
   * workspace: device scratch for split-K partial tiles (256-byte aligned, >= edtr_gemm_workspace_size(...) bytes to
   * enable every split the planner may pick; smaller buffers just limit the split factor; NULL disables split-K).
   * One workspace must not be used by two launches that may run concurrently (one per stream).
   * max_clusters: CTA pairs this launch may occupy (0 = the whole GPU, 74 pairs on a B200). */
  void* workspace;
  uint64_t workspace_bytes;
  int32_t max_clusters;
  /* ---- LayerNorm folded into this GEMM (replaces nn.LayerNorm in front of to_q/k/v and the GEGLU projection,
   * model/attention.py:222-224,231-233): A holds the UN-normalised rows x, Wt must be pre-scaled by the LayerNorm
   * gain (Wt[n,k] * gamma[k], rounded to bf16), and the epilogue evaluates
   *     acc' = rstd[row] * (acc - mean[row] * ln_colsum[n]),      ln_colsum[n] = sum_k Wt[n,k] (of the bf16 values)
   * before alpha / bias (the bias must already contain sum_k beta[k] * W[n,k]).  mean / rstd come from
   * ln_stats = fp32 [M][ln_parts][2] partial (sum, sum of squares) pairs as written by a producer's row_stats
   * (below), over ln_c = K channels, eps inside the sqrt.  NULL = no LayerNorm. */
  const float* ln_stats;
  int32_t ln_parts;
  int32_t ln_c;
  float ln_eps;
  const float* ln_colsum;
  /* ---- row statistics of the STORED matrix (the producer side of the fold above): row_stats = fp32
   * [M][row_stats_cap][2]; for every row the launch writes P = edtr_gemm_row_stats_parts(...) <= row_stats_cap
   * (sum, sum of squares) pairs — one per column tile and epilogue warp group, in-lane sums of the fp32 values before
   * bf16 rounding — into entries [m][0..P); a consumer passes ln_parts = P with a row stride of P pairs, so allocate
   * exactly [M][P][2].  EDTR_OUT_BF16, act != GEGLU.  NULL = not wanted. */
  float* row_stats;
  int32_t row_stats_cap;
  /* ---- GroupNorm statistics of the STORED matrix, produced by the epilogue (replaces the statistics pass of the
   * GroupNorm that consumes this launch's output: model/vae.py:103-113 norm1 / norm2, :279-283, :553): gn_partial =
   * fp32 [images][gn_slabs][N/u][2], u = gn_unit (4, or 2 for group widths that are not multiples of 4: 320 / 32 = 10).
   * Output rows are grouped in 32-row slabs (one epilogue warp; gn_hw = output rows
   * per image of THIS launch, gn_hw % 32 == 0, gn_hw | M) and columns in u-channel units; the launch writes, for every
   * (slab, unit) it covers, the (sum, sum of squares) of the fp32 values before bf16 rounding to
   * gn_partial[((image * gn_slabs + gn_slab0 + slab_in_image) * (N/u) + unit) * 2] — every entry exactly once, no
   * atomics, fixed summation order.  gn_slabs >= gn_slab0 + gn_hw / 32 is the slab count per image of the buffer
   * (edtr_conv3x3_up2x_bf16 fills 4 * H*W/32 slabs per image: the library offsets gn_slab0 per phase itself).
   * edtr_groupnorm_fold turns the buffer into (mean, variance) per (image, group) for edtr_groupnorm_apply_stats.
   * CTA-pair kernel only (M >= 256, N % 64 == 0), EDTR_OUT_BF16, act != GEGLU; the launch does not split K.
   * NULL = not wanted. */
  float* gn_partial;
  int32_t gn_hw;
  int32_t gn_slabs;
  int32_t gn_slab0;
  int32_t gn_unit; /* channels per unit: 0 (= 4), 2 or 4 */
} EdtrEpilogue;


/* ---- library ------------------------------------------------------------ */
const char* edtr_last_error(void);
int edtr_version(void);
/* Makes `device` current for this library's runtime instance (one process per
 * GPU: call once with the rank's device before edtr_init). */
int edtr_set_device(int device);
/* Verifies the current device is sm_100-class and primes kernel attributes. */
int edtr_init(void);
/* Bytes of EdtrEpilogue.workspace with which the GEMM / convolution planner is free to pick any split-K factor for an
 * M x N x K problem (fp32 partial tiles); a smaller workspace only limits the split.  There is no
 * library-global scratch: the workspace (and the share of the GPU a launch may take, EdtrEpilogue.max_clusters) are
 * per-call arguments, so calls on different devices / streams / host threads do not interact. */
size_t edtr_gemm_workspace_size(int M, int N, int K);
/* Pairs per row that edtr_gemm_bf16(M, N, K, ep) writes to ep->row_stats (depends on the tile plan, i.e. on the shape
 * and on ep->workspace_bytes / ep->max_clusters / ep->act; ep->row_stats itself may still be NULL). */
int edtr_gemm_row_stats_parts(int M, int N, int K, const EdtrEpilogue* ep);
/* N-tile width the GEMM will use for an N-column problem (for GEGLU weight
 * interleaving at plan time). */
int edtr_gemm_tile_n(int M, int N, int K, int act);

/* ---- tensor-core kernels (tcgen05 / TMEM / TMA) ------------------------- */
/* out = epilogue(A[M,K] * Wt[N,K]^T).  A, Wt bf16, K-contiguous, lda/ldw in
 * elements (multiples of 8), K a multiple of 64.
 * replaces: nn.Linear / 1x1 nn.Conv2d call sites — model/attention.py:23,43,
 * 170-174,266,280; model/unet.py:168-171,189,476-480; model/controlnet.py:129-133,
 * 260-261; model/vae.py:97-101,265-284,689-690. */
int edtr_gemm_bf16(const void* A, int lda, const void* Wt, int ldw, int M, int N, int K,
                   const EdtrEpilogue* ep, void* stream);

/* 3x3 / stride 1 / zero-pad 1 convolution as an implicit GEMM.  X is bf16
 * channels-last [B, H, W, Cin] with pixel stride ldx; Wt is bf16 [Cout, 3, 3, Cin]
 * (tap-major, channel-minor).  Cin % 64 == 0; W in {8,16,32,64} or W % 128 == 0.
 * Output rows are pixels in (b, y, x) order.
 * replaces: conv_nd(dims=2, ..., 3, padding=1) — model/unet.py:152,178,76-78,
 * 675-679; model/controlnet.py:137; model/vae.py:74-88,36-38,477-481,523-527. */
int edtr_conv3x3_bf16(const void* X, int ldx, int B, int H, int W, int Cin, const void* Wt,
                      int Cout, const EdtrEpilogue* ep, void* stream);

/* Nearest 2x up-sampling followed by a 3x3 / pad-1 convolution, without materialising the up-sampled tensor:
 * four 2x2-tap implicit GEMMs on the low-resolution input, one per output phase (py, px).  X is bf16
 * channels-last [B, H, W, Cin]; Wt4 is bf16 [2(py), 2(px), Cout, 2(dy), 2(dx), Cin] holding the phase filters
 * (sums of the 3x3 taps that fall on the same source pixel); the output is bf16 [B, 2H, 2W, Cout] with pixel
 * stride ep->ldc.  Epilogue: bias and optional SiLU only.  W a power of two >= 8, B*H*W >= 256.
 * replaces: Upsample.forward = F.interpolate(nearest, x2) + conv — model/unet.py:69-79, model/vae.py:36-38. */
int edtr_conv3x3_up2x_bf16(const void* X, int ldx, int B, int H, int W, int Cin, const void* Wt4, int Cout,
                           const EdtrEpilogue* ep, void* stream);

/* Flash-style softmax(Q K^T * scale) V, head dim 64, no mask.
 * Q [B, Lq, heads*64] (row stride ldq), K/V [B, Lk, heads*64] (strides ldk/ldv),
 * O [B, Lq, heads*64] (stride ldo); all bf16; head h occupies columns
 * [64h, 64h+64).  Lk is arbitrary (tail keys are masked).
 * replaces: F.scaled_dot_product_attention + the head split/merge permutes —
 * model/attention.py:176-203. */
int edtr_attention_bf16(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv,
                        void* O, int ldo, int B, int heads, int Lq, int Lk, float scale,
                        void* stream);

/* ---- memory-bound kernels ------------------------------------------------ */
/* GroupNorm statistics, pass 1: per-(image, pixel chunk, channel slab, group) partial sum and
 * sum of squares of X (bf16 [B, HW, C], row stride ldx) written to `stats` (fp32,
 * edtr_groupnorm_partial_size(B, HW, C, groups) elements; no zero-fill needed, no atomics, so the
 * result is bit-reproducible).  replaces: first half of GroupNorm32 — model/util.py:161-163,
 * model/attention.py:50-51, model/vae.py:22-23. */
size_t edtr_groupnorm_partial_size(int B, int HW, int C, int groups);
int edtr_groupnorm_stats(const void* X, int ldx, int B, int HW, int C, int groups, float* stats,
                         void* stream);
/* Pass 2: folds the partial sums of edtr_groupnorm_stats (same B, HW, C, groups) and writes
 * Y = (X - mean) * rstd * gamma + beta, optionally followed by SiLU; biased
 * variance, eps inside the sqrt.  Y bf16 [B, HW, C] with row stride ldy.
 * replaces: GroupNorm32 + nn.SiLU — model/unet.py:149-151,173-175,675-677;
 * model/vae.py:103-113,553-554 (nonlinearity). */
int edtr_groupnorm_apply(const void* X, int ldx, void* Y, int ldy, int B, int HW, int C,
                         int groups, const float* stats, const float* gamma, const float* beta,
                         float eps, int silu, void* stream);

/* Tiled VAE (utils/tilevae/tilevae.py:232-304): every GroupNorm of a tiled decode uses statistics pooled over
 * the tiles.  edtr_groupnorm_pool folds the partial sums edtr_groupnorm_stats wrote for ONE tile (same B, HW, C,
 * groups) into that tile's mean / biased variance per (image, group) and adds weight * (mean, var) into
 * acc[B][groups][2] (fp32, zeroed by the caller before the first tile; stream-ordered, deterministic).
 * edtr_groupnorm_apply_stats normalises with given statistics mean_var[B][groups][2] instead of the tensor's own
 * (custom_group_norm, :188-215). */
int edtr_groupnorm_pool(const float* stats, int B, int HW, int C, int groups, float weight, float* acc, void* stream);
int edtr_groupnorm_apply_stats(const void* X, int ldx, void* Y, int ldy, int B, int HW, int C, int groups,
                               const float* mean_var, const float* gamma, const float* beta, float eps, int silu,
                               void* stream);
/* Folds the partial sums a GEMM / convolution epilogue wrote through EdtrEpilogue.gn_partial (fp32
 * [B][slabs][C/unit][2], every slab = 32 rows of the image, unit = 2 or 4 channels) into mean_var[B][groups][2] =
 * (mean, biased variance) per (image, group), the input of edtr_groupnorm_apply_stats: the GroupNorm of a tensor that
 * a convolution just produced needs no pass over the tensor for its statistics.  (C / groups) % unit == 0; fixed
 * reduction order (deterministic).
 * replaces: the statistics half of GroupNorm — model/vae.py:26-28 (Normalize), model/util.py:161-163. */
int edtr_groupnorm_fold(const float* gn_partial, int B, int slabs, int C, int groups, int unit, float* mean_var,
                        void* stream);

/* Single-launch GroupNorm (+SiLU) for small L2-resident tensors (<= 1 MB per image): a thread-block cluster per
 * image, statistics exchanged through distributed shared memory, deterministic.  Same arithmetic as
 * edtr_groupnorm_stats + edtr_groupnorm_apply (reference: model/util.py:161-163, model/attention.py:50-51).
 * edtr_groupnorm_fused_supported returns 1 when the shape is eligible. */
int edtr_groupnorm_fused_supported(int B, int HW, int C, int groups);
int edtr_groupnorm_fused(const void* X, int ldx, void* Y, int ldy, int B, int HW, int C, int groups,
                         const float* gamma, const float* beta, float eps, int silu, void* stream);
/* Row-wise LayerNorm over C (biased variance, eps 1e-5 typical), bf16 in/out.
 * replaces: nn.LayerNorm — model/attention.py:222-224. */
int edtr_layernorm_bf16(const void* X, int ldx, void* Y, int ldy, int M, int C,
                        const float* gamma, const float* beta, float eps, void* stream);
/* Same over rows that are padded to C channels: statistics over the first C_real channels, the pad channels hold
 * zeros on input and (gamma = beta = 0 there) on output.  replaces: nn.LayerNorm(180) in 192-wide rows —
 * model/swinir.py:205,212,531-533,760. */
int edtr_layernorm_padded_bf16(const void* X, int ldx, void* Y, int ldy, int M, int C, int C_real,
                               const float* gamma, const float* beta, float eps, void* stream);

/* ---- SwinIR pre-restoration network (model/swinir.py) --------------------- */
/* Y[b, y, x, (c*r + dy)*r + dx] = (X[b, c, y*r + dy, x*r + dx] - mean3[c]) * scale: nn.PixelUnshuffle(8) of the
 * mean-shifted fp32 NCHW image into a bf16 channels-last tensor with pixel stride ldy.  mean3 is a HOST pointer to
 * C floats (or NULL).  replaces: model/swinir.py:859-860 + conv_first[0] (:700-704). */
int edtr_pixel_unshuffle_f32_to_nhwc_bf16(const float* X, void* Y, int ldy, int B, int C, int H, int W, int r,
                                          const float* mean3, float scale, void* stream);
/* Window attention of one SwinTransformerBlock on a [B, H, W] token grid (H, W multiples of 8): QKV is the fused
 * projection output, rows = tokens in (b, y, x) order with stride ld, columns q | k | v each heads*32 wide (head h at
 * [32h, 32h+32), the reference's 30 channels + 2 zero pads); O likewise heads*32 wide with stride ldo.  `shift` is the
 * cyclic shift (0 or 4), `bias` fp32 [heads, 64, 64] the gathered relative-position bias, `mask` fp32
 * [(H/8)*(W/8), 64, 64] the 0 / -100 SW-MSA mask (NULL when shift == 0).  The roll, the window partition and their
 * inverses are index arithmetic.  replaces: model/swinir.py:120-151, 253-281. */
int edtr_window_attention_bf16(const void* QKV, int ld, void* O, int ldo, int B, int H, int W, int heads,
                               int shift, float scale, const float* bias, const float* mask, void* stream);
/* Row softmax of fp32 S[M, N] (stride lds) scaled by `scale`, bf16 output P.
 * replaces: the softmax inside SDPA for the single-head d=512 VAE attention —
 * model/vae.py:298. */
int edtr_softmax_rows(const float* S, int lds, void* P, int ldp, int M, int N, float scale,
                      void* stream);

/* Y[b, 2y+dy, 2x+dx, :] = X[b, y, x, :]  (nearest, x2), bf16 channels-last.
 * replaces: F.interpolate(scale_factor=2, mode="nearest") — model/unet.py:76,
 * model/vae.py:36. */
int edtr_upsample2x_bf16(const void* X, int ldx, void* Y, int ldy, int B, int H, int W, int C,
                         void* stream);
/* Generic im2col for the convolutions the TMA path does not cover (stride 2,
 * asymmetric padding, odd sizes): Y[(b,oy,ox), (ky,kx,c)] = X[b, oy*s+ky-pt, ox*s+kx-pl, c]
 * (zero outside).  replaces: the patch gather inside nn.Conv2d(stride=2) —
 * model/unet.py:99-101, model/vae.py:54-58. */
int edtr_im2col_bf16(const void* X, int ldx, void* Y, int B, int H, int W, int C, int KH, int KW,
                     int stride, int pad_top, int pad_left, int Ho, int Wo, void* stream);

/* NCHW fp32 -> channels-last bf16: Y[(b, p), coff + c] = X[b, c, p].
 * replaces: x.type(self.dtype) + torch.cat((x, hint), 1) — model/controlnet.py:266-269. */
int edtr_nchw_f32_to_nhwc_bf16(const float* X, void* Y, int ldy, int coff, int B, int C, int HW,
                               void* stream);
/* 1x1 convolution over a few channels fused with the layout change:
 * Y[(b, p), coff + co] = bias[co] + sum_ci W[co, ci] * (scale * X[b, ci, p]), Cin, Cout <= 16,
 * X NCHW fp32, Y channels-last bf16, fp32 math.
 * replaces: z / self.scale_factor (model/cldm.py:156) + post_quant_conv (model/vae.py:732). */
int edtr_pointwise_nchw_f32_to_nhwc_bf16(const float* X, const float* W, const float* bias, float scale,
                                         void* Y, int ldy, int coff, int B, int Cin, int Cout, int HW,
                                         void* stream);
/* channels-last bf16 -> NCHW (fp32 if out_f32 else bf16): Y[b, c, p] = X[(b,p), c]. */
int edtr_nhwc_bf16_to_nchw(const void* X, int ldx, void* Y, int B, int C, int HW, int out_f32,
                           void* stream);
/* fp32 -> bf16 element cast (n elements). */
int edtr_cast_f32_to_bf16(const float* X, void* Y, size_t n, void* stream);

/* Numerator of the gaussian tile blend of the cldm-tiled path: out[bc, y, x] = sum over the `ntiles` given
 * tiles covering (y, x) of weight[y-hi, x-wi] * tiles[t, bc, y-hi, x-wi]; coords = int32 [ntiles, 2] (hi, wi),
 * all fp32.  The caller divides by the (data-independent) weight sum, after the all-reduce when the tiles are
 * spread over ranks.  replaces: out[..] += fn(tile) * weights — utils/common.py:414-424. */
int edtr_tile_blend(const float* tiles, const int32_t* coords, int ntiles, const float* weight, float* out,
                    int BC, int H, int W, int th, int tw, void* stream);

/* Sinusoidal timestep embedding [B, dim] = [cos(t f_k) | sin(t f_k)],
 * f_k = exp(-ln(max_period) k / (dim/2)), computed in fp32, stored bf16.
 * replaces: timestep_embedding — model/util.py:98-118. */
int edtr_timestep_embedding(const int64_t* t, void* Y, int B, int dim, float max_period,
                            void* stream);

/* One spaced-DDPM update for every element of x [n_per_image * B]:
 *   pred_x0 = sqrt_recip[idx]*x - sqrt_recipm1[idx]*eps
 *   mean    = coef1[idx]*pred_x0 + coef2[idx]*x
 *   x_prev  = mean + (idx != 0) * sqrt(var[idx]) * noise
 * idx is the per-image coefficient index (int64 [B]); tables are fp32 device
 * arrays.  pred_x0 may be NULL.
 * replaces: SpacedSampler.p_sample arithmetic — utils/sampler.py:150-164,195-203. */
int edtr_sampler_update(const float* x, const float* eps, const float* noise,
                        const int64_t* index, const float* sqrt_recip, const float* sqrt_recipm1,
                        const float* coef1, const float* coef2, const float* var, float* x_prev,
                        float* pred_x0, int B, int n_per_image, void* stream);

/* One level of the wavelet colour fix applied after the decode (utils/common.py:99-147, wavelet_blur /
 * wavelet_decomposition / wavelet_reconstruction): 3x3 binomial blur with dilation `radius`, replicate padding, on
 * every plane of a contiguous fp32 [planes, H, W] tensor.  mode 0: out = low; mode 1: out = low and
 * high (+)= in - low (`first` != 0 stores); mode 2: out = high + low.  `in` and `out` must not alias. */
int edtr_wavelet_level(const float* in, float* out, float* high, int planes, int H, int W, int radius, int mode,
                       int first, void* stream);

/* ---- fp32 mode (BASELINE.json: per-step latent max-rel error <= 1e-4) ------------------------------------------
 * Every tensor stays fp32 in HBM and every contraction accumulates fp32 products on the CUDA cores: the accuracy mode
 * of the path (the bf16 tensor-core kernels above are the throughput mode).  One generic implicit GEMM:
 *   C = act(alpha * A * W^T + bias[col] + rowvec[(row / rows_per_group) * rowvec_ld + col] + residual[row * ldr + col])
 * A [M, K] row-major (lda), or — conv = 1 — the 3x3 im2col view of an NHWC tensor [B, H, W_in, Cin] with pixel stride
 * lda: K = 9 * Cin tap-major / channel-minor, rows = output pixels (b, yo, xo) of an Ho x Wo grid, input pixel
 * (yo * conv_stride + ty - pad_top, xo * conv_stride + tx - pad_left), zero outside; up2x = 1: 3x3 / pad 1 on the
 * nearest-2x up-sampled grid (Ho = 2H, Wo = 2W_in).  W [N, K] row-major (ldw), or [K, N] when w_kn = 1.  batch1 x
 * batch2 problems with element strides *_stride1 / *_stride2 (attention: samples x heads).  out_nchw = 1 stores
 * C[((row / hw) * N + col) * hw + row % hw].
 * replaces (fp32 mode): every nn.Conv2d / nn.Linear / einsum call site listed for edtr_gemm_bf16 / edtr_conv3x3_bf16
 * and the attention products of model/attention.py:176-203, model/vae.py:279-308. */
typedef struct EdtrF32Gemm {
  const float* A;
  const float* W;
  float* C;
  int32_t M, N, K;
  int64_t lda, ldw, ldc;
  int32_t w_kn;
  int32_t batch1, batch2;
  int64_t a_stride1, a_stride2, w_stride1, w_stride2, c_stride1, c_stride2;
  int32_t conv, H, W_in, Cin, Ho, Wo, conv_stride, pad_top, pad_left, up2x;
  float alpha;
  const float* bias;
  const float* rowvec;
  int64_t rowvec_ld;
  int32_t rows_per_group;
  const float* residual;
  int64_t ldr;
  int32_t act;      /* EDTR_ACT_NONE or EDTR_ACT_SILU */
  int32_t out_nchw;
  int32_t hw;
} EdtrF32Gemm;
int edtr_f32_gemm(const EdtrF32Gemm* g, void* stream);
/* GroupNorm (+SiLU) on fp32 [B, HW, C] rows (row strides ldx / ldy), statistics accumulated in double precision;
 * scratch >= edtr_f32_groupnorm_scratch_bytes(B, HW, C) bytes, 8-byte aligned.
 * replaces (fp32 mode): GroupNorm32 / Normalize — model/util.py:146-163, model/attention.py:50-51, model/vae.py:26-28. */
size_t edtr_f32_groupnorm_scratch_bytes(int B, int HW, int C);
int edtr_f32_groupnorm(const float* X, long long ldx, float* Y, long long ldy, int B, int HW, int C, int groups,
                       const float* gamma, const float* beta, float eps, int silu, void* scratch, void* stream);
/* nn.LayerNorm over the last dimension (model/attention.py:222-224). */
int edtr_f32_layernorm(const float* X, long long ldx, float* Y, long long ldy, int M, int C, const float* gamma,
                       const float* beta, float eps, void* stream);
/* In place: S[r, :N] = softmax(scale * S[r, :N]) (model/attention.py:196-199, model/vae.py:298-301). */
int edtr_f32_softmax_rows(float* S, long long lds, long long rows, int N, float scale, void* stream);
/* Y[m, n] = X[m, n] * gelu_erf(X[m, N + n]) (model/attention.py:20-27). */
int edtr_f32_geglu(const float* X, long long ldx, float* Y, long long ldy, long long M, int N, void* stream);
/* Y = X * sigmoid(X) (model/unet.py:166-172: the SiLU in front of the embedding projections). */
int edtr_f32_silu(const float* X, float* Y, long long n, void* stream);
/* X [B, C, HW] fp32 (NCHW) times `scale` -> channels [coff, coff + C) of Y [B, HW, ldy] fp32 (channels-last). */
int edtr_f32_nchw_to_nhwc(const float* X, float* Y, long long ldy, int B, int C, int HW, int coff, float scale,
                          void* stream);
/* out[b] = [cos(t_b f_k) | sin(t_b f_k)], f_k = exp(-ln(max_period) k / (dim/2)), fp32 (model/util.py:98-118). */
int edtr_f32_timestep_embedding(const long long* t, float* out, int B, int dim, float max_period, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EDTR_B200_H_ */
