/*
 * wxformer_b200.h — C ABI of the B200-native WXFormer/CrossFormer forecast forward step.
 *
 * The reference (NCAR/miles-credit) has no native code: its boundary for this path is the Python
 * nn.Module `credit.models.crossformer.CrossFormer` called as `y = model(x)`
 * (credit/trainers/rollout_utils.py:281).  The entry points below are what a binding for that path
 * binds: one per fused block of `CrossFormer.forward` (credit/models/crossformer.py:593-644).
 *
 * Conventions
 *  - plain C: raw device pointers, ints, a cudaStream_t passed as void*; no torch types.
 *  - every function only enqueues work on `stream` (no sync, no allocation, graph-capturable),
 *    returns 0 on success, a positive cudaError_t, or a negative WXF_E* code for bad arguments.
 *  - activations are fp32 "pixel-major" (NHWC): element (b, y, x, c) of a [B,H,W,C] field lives at
 *    ((b*H + y)*W + x)*ld + c, where ld >= C is the pixel stride in elements (so a producer can
 *    write straight into a channel slice of a concatenated tensor).
 *  - the caller owns every buffer.
 */
#ifndef WXFORMER_B200_H
#define WXFORMER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WXF_ABI_VERSION 17

#define WXF_EINVAL (-1)      /* bad argument / unsupported geometry */
#define WXF_EALIGN (-2)      /* pointer or stride not aligned as the kernel requires */
#define WXF_EUNSUPPORTED (-3)

#define WXF_PAD_EARTH 0
#define WXF_PAD_MIRROR 1

#define WXF_ACT_NONE 0
#define WXF_ACT_GELU_ERF 1

#define WXF_ATTN_SHORT 0
#define WXF_ATTN_LONG 1

/* ABI version of the loaded library (WXF_ABI_VERSION). */
int wxf_abi_version(void);

/* Last error text of the calling thread (static storage; never NULL). */
const char* wxf_last_error(void);

/*
 * Boundary padding fused with the NCHW -> pixel-major transpose.
 * Replaces TensorPadding.pad (credit/boundary_padding.py:20-33: _earth_padding :50-72,
 * _mirror_padding :98-117) and the frame flatten (credit/models/crossformer.py:604-609).
 *   x  : [B, C, T, H, W] fp32 contiguous
 *   xp : [B, H+pt+pb, W+pl+pr, ld] fp32, channel index c*T + t; channels [C*T, ld) are zero-filled.
 */
int wxf_pad_to_pixel_major(const float* x, float* xp, int B, int C, int T, int H, int W,
                           int pt, int pb, int pl, int pr, int mode, int ld, int row0, int nrows, void* stream);
/* Same pass, result written as fp16 hi/lo operand planes [B, Hp, Wp, ld] (input of the stage-0 cross-embed).
 * Both variants write rows [row0, row0 + nrows) of the padded image only (0, Hp = everything): a rank of the lat-band
 * decomposition pads just the rows its stage-0 band reads (the reference pads the full grid on every rank,
 * credit/trainers/trainer_gen2.py:211-213). */
int wxf_pad_to_pixel_major_f16x2(const float* x, void* xp_hi, void* xp_lo, int B, int C, int T, int H, int W,
                                 int pt, int pb, int pl, int pr, int mode, int ld, int row0, int nrows, void* stream);

/*
 * The per-step pre-blocks fused into the padding pass: instead of one [B, C, T, H, W] tensor the source is a
 * device table of per-channel planes, chan_planes[b*C + c] -> [T, H, W] fp32 (the variables in ConcatToTensor's sorted
 * order, credit/preblock/concat.py:101-207: the torch.cat along the channel axis never materialises), and every value is
 * z-scored on the way, (x - mean[c]) / max(std[c], 1e-12) (ERA5Normalizer._normalize_tensor, credit/preblock/norm.py:80-98;
 * pass-through variables carry mean 0 / std 1).  Output exactly as wxf_pad_to_pixel_major (xp) or its _f16x2 variant
 * (xp_hi / xp_lo); give one of the two.  ld % 8 == 0.
 */
int wxf_preblock_pad_to_pixel_major(const void* const* chan_planes, const float* mean, const float* stdv, float* xp,
                                    void* xp_hi, void* xp_lo, int B, int C, int T, int H, int W, int pt, int pb, int pl,
                                    int pr, int mode, int ld, int row0, int nrows, void* stream);

/*
 * Channel LayerNorm at every pixel (credit/models/crossformer.py:182-192):
 *   y = (x - mean) / sqrt(var_biased + eps) * g + b   over the d channels of each of M pixels.
 */
int wxf_layernorm(const float* x, int ldx, float* y, int ldy, const float* g, const float* b,
                  int64_t M, int d, float eps, void* stream);

/*
 * Same LayerNorm, result written as the two fp16 operand planes of the tensor-core GEMM
 * (hi = fp16(y), lo = fp16(y - hi); see wxf_gemm_f16x2_tc).  y_hi / y_lo : [M, ldh] fp16.
 */
int wxf_layernorm_f16x2(const float* x, int ldx, void* y_hi, void* y_lo, int ldh, const float* g, const float* b,
                        int64_t M, int d, float eps, void* stream);

/*
 * Implicit-GEMM convolution, exact fp32 FMA.  One descriptor covers
 *   - Conv2d kxk stride s zero-pad p        (CrossEmbedLayer, crossformer.py:139-152; UpBlock 3x3, :96-99)
 *   - 1x1 Conv2d                           (to_qkv/to_out :229-230, FeedForward :198-204)
 *   - ConvTranspose2d k2 s2 and k4 s2 p1   (crossformer.py:92, :572-574) as `phases` = 4 output-parity
 *     classes, each an ordinary small convolution scattered to (oy*out_scale + phase/2, ox*out_scale + phase%2).
 *   - Conv2d 3x3 -> 4C channels + PixelShuffle(2) (wxformer/crossformer.py:143-159, 813-830): phase (dy, dx) owns the
 *     channels 4c + 2dy + dx, same scatter, per-phase bias (bias_phase_stride = C).
 * For GEMM row m = (b, oy, ox) over the [B, Ho, Wo] grid and phase z:
 *   acc[n] = sum_{t<T} sum_{c<Cin} in[b, oy*stride + taps[z][t].dy, ox*stride + taps[z][t].dx, c]
 *                                  * w[z][n][t*Cin + c]          (out-of-range pixels read as 0)
 *   v = acc[n] + bias[n];  v = act(v);  v += res[pixel, n] (if res);  out[pixel, c_off + n] = v
 */
typedef struct WxfConvDesc {
  const float* in;   /* [B, Hi, Wi, lda] */
  const float* w;    /* [phases, N, T*Cin] */
  const int32_t* taps; /* [phases, T, 2] = (dy, dx), padding already subtracted */
  const float* bias; /* [N] or NULL */
  const float* res;  /* residual, same pixel grid as out, or NULL */
  float* out;        /* [B, Ho*out_scale, Wo*out_scale, ldc] */
  int32_t B, Hi, Wi, lda, Cin;
  int32_t N, T, stride;
  int32_t Ho, Wo;
  int32_t phases, out_scale;
  int32_t ldc, c_off;
  int32_t ldr, r_off;
  int32_t act;
  int32_t bias_phase_stride; /* bias index = phase*bias_phase_stride + n (0: shared; N: sub-pixel conv + PixelShuffle) */
} WxfConvDesc;

int wxf_conv_igemm_f32(const WxfConvDesc* desc, void* stream);

/*
 * Cross-scale window attention core (credit/models/crossformer.py:261-296, 301-314) for all windows
 * and heads of one Attention block, on the pixel-major output of the to_qkv 1x1 conv:
 *   qkv   : [B, H, W, ldq] with q = [0,d), k = [d,2d), v = [2d,3d); head h = channels h*dh..h*dh+dh-1
 *   biasT : [L, L] transposed position bias, biasT[j*L + i] = bias[i][j]  (L = wsz*wsz)
 *   out   : [B, H, W, ldo] channel h*dh + e   (input of to_out)
 * S = (q*scale) k^T + bias; P = softmax(S); out = P v.   kind: WXF_ATTN_SHORT tiles wsz x wsz blocks,
 * WXF_ATTN_LONG groups the dilated tokens (l1*(H/wsz)+gh, l2*(W/wsz)+gw).   dh must be 32, L <= 128.
 */
int wxf_window_attention_f32(const float* qkv, int ldq, const float* biasT, float* out, int ldo,
                             int B, int H, int W, int d, int dh, int wsz, int kind, float scale, void* stream);

/* Same attention core; the result is written as fp16 hi/lo operand planes [B*H*W, ldh] for the to_out GEMM. */
int wxf_window_attention_f16x2(const float* qkv, int ldq, const float* biasT, void* out_hi, void* out_lo, int ldh,
                               int B, int H, int W, int d, int dh, int wsz, int kind, float scale, void* stream);

/*
 * The same attention core on the tcgen05 tensor cores: Q, K, V gathered by TMA from the fp16 hi/lo planes of the
 * to_qkv output ([B, H, W, ldq]; short windows as 4-D boxes, dilated long groups as 5-D boxes), S = QK^T and O = PV as
 * f16x2 three-pass MMAs with TMEM accumulators, windows of the same head packed block-diagonally into 128-row tiles,
 * softmax in fp32 registers.  Output: fp16 hi/lo planes [B*H*W, ldh].  dh must be 32, L <= 128, ldq % 8 == 0.
 * bias_tile: the [128 x 128] fp32 tile wxf_attention_bias_tile() builds once per layer from the transposed position
 * bias (bias * log2 e inside a row's own window, -1e30 elsewhere; the kernel keeps it in TMEM).
 */
int wxf_attention_bias_tile(const float* biasT, float* tile, int W, int wsz, int kind, void* stream);
int wxf_window_attention_tc(const void* qkv_hi, const void* qkv_lo, int ldq, const float* bias_tile, void* out_hi,
                            void* out_lo, int ldh, int B, int H, int W, int d, int dh, int wsz, int kind, float scale,
                            void* stream);

/*
 * Pointwise (1x1 conv) GEMM on the tcgen05 tensor cores with TMA-staged operands
 * (to_qkv / to_out, crossformer.py:229-230, 268, 297; FeedForward 1x1 convs, :198-204):
 *     acc[m, n] = sum_k A[m, k] * W[n, k]
 * with every fp32 operand carried as two fp16 planes (hi = fp16(x), lo = fp16(x - hi)) and three MMA passes
 * (hi*lo + lo*hi + hi*hi) into one fp32 TMEM accumulator: 22-bit operands, fp32 accumulation.
 *   a_hi, a_lo : [M, lda] fp16;  w_hi, w_lo : [N, K] fp16 planes of W * 2^w_scale_log2
 *   v = acc * 2^-w_scale_log2 + bias[n];  v = act(v);  v += res[m*ldr + r_off + n]
 *   out (fp32, optional): out[m*ldc + c_off + n] = v;  out_hi/out_lo (fp16 planes, optional): [M, ldh]
 * K and lda must be multiples of 8; all planes 16-byte aligned.
 */
typedef struct WxfGemmDesc {
  const void* a_hi;
  const void* a_lo;
  const void* w_hi;
  const void* w_lo;
  const float* bias;
  const float* res;
  float* out;
  void* out_hi;
  void* out_lo;
  int64_t M;
  int32_t N, K, lda;
  int32_t ldc, c_off, ldr, r_off, ldh;
  int32_t act, w_scale_log2;
} WxfGemmDesc;

int wxf_gemm_f16x2_tc(const WxfGemmDesc* desc, void* stream);


/*
 * Convolution as an implicit GEMM on the tcgen05 tensor cores (same f16x2 scheme and epilogue as
 * wxf_gemm_f16x2_tc).  Covers the same operators as wxf_conv_igemm_f32 (Conv2d k x k stride 1/2 with zero
 * padding; ConvTranspose2d k2 s2 / k4 s2 p1 as 4 output-parity phases): CrossEmbedLayer stages 1-3
 * (crossformer.py:139-152), UpBlock (crossformer.py:92-116), up_block4 (crossformer.py:572-574).
 * The A tile of one K-step (one tap, 64 channels) is a single 4-D TMA box of the pixel-major fp16 planes;
 * out-of-image pixels are TMA zero fill.
 *   in_hi, in_lo : [B, Hi, Wi, lda] fp16 planes;   w_hi, w_lo : [phases, N, T*cin_pad] planes of W * 2^w_scale_log2
 *   taps         : HOST pointer, [phases, T, 2] = (dy, dx) with padding already subtracted, T <= 64
 *   epilogue     : v = acc*2^-w_scale_log2 + bias[n]; act; + res[pixel*ldr + r_off + n];
 *                  out[pixel*ldc + c_off + n] (fp32, optional), out_hi/lo[pixel*ldh + h_off + n] (optional)
 * N % 4 == 0, lda % 8 == 0, cin_pad % 64 == 0 (weights zero-padded per tap), stride in {1, 2, 4} (4: the patch
 * embedding of FuXi's CubeEmbedding, Conv3d kernel = stride = (T, 4, 4) with the frames folded into the channels,
 * credit/models/fuxi.py:107).
 */
typedef struct WxfConvTcDesc {
  const void* in_hi;
  const void* in_lo;
  const void* w_hi;
  const void* w_lo;
  const int32_t* taps;
  const float* bias;
  const float* res;
  float* out;
  void* out_hi;
  void* out_lo;
  int32_t B, Hi, Wi, lda, Cin, cin_pad;
  int32_t N, T, stride;
  int32_t Ho, Wo;
  int32_t phases, out_scale;
  int32_t ldc, c_off, ldr, r_off, ldh, h_off;
  int32_t act, w_scale_log2;
  int32_t bias_phase_stride; /* bias index = phase*bias_phase_stride + n */
} WxfConvTcDesc;

int wxf_conv_f16x2_tc(const WxfConvTcDesc* desc, void* stream);

/*
 * Stage-0 CrossEmbedLayer branch (Conv2d k x k, stride 2, zero pad p = (k-2)/2, few output channels;
 * crossformer.py:139-152) as a Toeplitz-lifted implicit GEMM on the tensor cores: the kernel column
 * kx = 2j + r is split and j is folded into the GEMM's N dimension,
 *     P[oy, m, (j, c)] = sum_{ky, r, ci} in[2 oy + ky - p, 2 m + r - p, ci] * W[c, ci, ky, 2j + r]
 *     out[oy, ox, c]   = bias[c] + sum_j P[oy, ox + j, (j, c)]
 * so N = (k/2)*ch instead of ch.  in planes: [B, Hi, Wi, lda] fp16 (Cin <= 64, zero-padded to 64 by TMA);
 * w planes: [(k/2)*ch, 2k*64] fp16 with row (j, c), column (ky, r, ci), pre-scaled by 2^w_scale_log2;
 * out: fp32 out[pixel*ldc + c_off + c].  Requires k even, (k/2)*ch <= 256, ch % 4 == 0.
 */
typedef struct WxfToeplitzDesc {
  const void* in_hi;
  const void* in_lo;
  const void* w_hi;
  const void* w_lo;
  const float* bias;
  float* out;
  int32_t B, Hi, Wi, lda, Cin, cin_pad;
  int32_t ch, kernel, pad;
  int32_t Ho, Wo;
  int32_t ldc, c_off;
  int32_t w_scale_log2;
  int32_t oy_off; /* domain decomposition: local output row oy stands for global row oy + oy_off of the input image */
} WxfToeplitzDesc;

int wxf_cross_embed_toeplitz_tc(const WxfToeplitzDesc* desc, void* stream);

/* Split an fp32 [M, ldx] matrix (first d columns) into fp16 hi/lo planes [M, ldh] (test/utility pass). */
int wxf_split_f16x2(const float* x, int ldx, void* hi, void* lo, int ldh, int64_t M, int d, void* stream);

/*
 * GroupNorm + SiLU on a pixel-major field (UpBlock residual stack, crossformer.py:96-116).
 * Two calls: statistics (deterministic two-level reduction, fp64 final combine) then apply.
 *   stats : [B, G, 2] fp32 (mean, rstd);   scratch: >= wxf_groupnorm_scratch_bytes(...) bytes.
 *   apply : y = silu((x - mean) * rstd * gamma + beta) (+ res if res != NULL)
 */
int64_t wxf_groupnorm_scratch_bytes(int B, int64_t HW, int C);
int wxf_groupnorm_stats(const float* x, int ldx, float* stats, void* scratch, int B, int64_t HW, int C, int G,
                        float eps, void* stream);
int wxf_groupnorm_silu(const float* x, int ldx, const float* stats, const float* gamma, const float* beta,
                       const float* res, int ldr, float* y, int ldy, int B, int64_t HW, int C, int G, void* stream);
/* Same apply pass, result written as fp16 hi/lo operand planes: y_hi/y_lo[pixel*ldh + h_off + c]. */
int wxf_groupnorm_silu_f16x2(const float* x, int ldx, const float* stats, const float* gamma, const float* beta,
                             const float* res, int ldr, void* y_hi, void* y_lo, int ldh, int h_off, int B, int64_t HW,
                             int C, int G, void* stream);

/*
 * GroupNorm statistics in two halves for the lat-band domain decomposition (the reference's DomainParallelGroupNorm
 * all-reduces (sum, sum of squares), credit/domain_parallel/layers.py:507-518): raw fp64 sums [B, G, 2] of the local
 * band, then - after the caller's all-reduce - (mean, rstd) from the global sums.
 */
int wxf_groupnorm_sums(const float* x, int ldx, double* sums, void* scratch, int B, int64_t HW, int C, int G, void* stream);
int wxf_groupnorm_stats_from_sums(const double* sums, float* stats, int B, int G, double count, float eps, void* stream);

/*
 * Row gather dst[i, 0:d] = src[idx[i], 0:d] (fp32 rows, idx int32 on the device): packs / unpacks the all-to-all buffers
 * that move the residual stream between the lat-band layout and the attention-unit layout (multi-GPU).
 */
int wxf_gather_rows(const float* src, int ld_src, const int32_t* idx, float* dst, int ld_dst, int64_t n, int d, void* stream);

/*
 * Un-pad + bilinear resize (align_corners = False) + pixel-major -> NCHW
 * (TensorPadding.unpad, boundary_padding.py:35-48; F.interpolate, crossformer.py:628-635).
 *   y   : [B, Hd, Wd, ld]; the crop is rows [top, top+Hc), cols [left, left+Wc)
 *   out : [B, C, Ho, Wo] fp32 contiguous (C = base_output_channels*output_frames); only output rows
 *         [o0, o0 + n_out) are written (0, Ho = everything; a lat-band rank writes its own rows, the per-shard
 *         un-pad of credit/parallel/domain.py:37-64)
 */
int wxf_unpad_resize_to_nchw(const float* y, int ld, float* out, int B, int C, int Hd, int Wd, int top, int left,
                             int Hc, int Wc, int Ho, int Wo, int o0, int n_out, void* stream);

/*
 * The same pass with the per-step post-blocks fused into its epilogue, per output channel c:
 *   v = v * scale[c] + shift[c]      inverse scaling (y * std + mean, applications/rollout_to_netcdf.py:287; the gen2
 *                                    bridgescaler inverse transform), two roundings like torch
 *   v = min(max(v, clamp_lo[c]), clamp_hi[c])   TracerFixer (credit/postblock/conservation.py:88-115); -inf / +inf = no clamp
 */
int wxf_unpad_resize_post_to_nchw(const float* y, int ld, float* out, int B, int C, int Hd, int Wd, int top, int left,
                                  int Hc, int Wc, int Ho, int Wo, int o0, int n_out, const float* scale, const float* shift,
                                  const float* clamp_lo, const float* clamp_hi, void* stream);

/*
 * GlobalMassFixer (credit/postblock/conservation.py:118-176; hybrid-sigma grid, midpoint quantities): the two global sums
 *   sums[b, 0] = sum_p area[p] * sum_l da[l] * (1 - q[b, l, p])
 *   sums[b, 1] = sum_p area[p] * sp[b, p] * sum_l db[l] * (1 - q[b, l, p])
 * over the pixels p in [p0, p0 + np) (a latitude band of a decomposed forecast; the caller adds the bands).  q is addressed
 * as q[b*q_bstride + l*q_lstride + p], sp as sp[b*sp_bstride + p] (channel views of an NCHW state).  Column sums in fp32 in
 * level order, pixel sums in fp64 (deterministic two-level reduction).  scratch: wxf_dry_mass_scratch_bytes(B), zeroed once.
 * wxf_scale_planes: x[b, 0:n] *= ratio[b] (the corrected surface pressure).
 */
int64_t wxf_dry_mass_scratch_bytes(int B);
int wxf_dry_mass_sums(const float* q, int64_t q_bstride, int64_t q_lstride, const float* sp, int64_t sp_bstride,
                      const float* area, const float* da, const float* db, int B, int L, int64_t p0, int64_t np, double* sums,
                      void* scratch, void* stream);
int wxf_scale_planes(float* x, int64_t bstride, int64_t n, const float* ratio, int B, void* stream);

/*
 * GlobalWaterFixer (credit/postblock/conservation.py:179-236) and GlobalEnergyFixerUpDown (:239-376) on the hybrid-sigma grid
 * with midpoint quantities.  Fields are channel views of NCHW tensors: a 3-D field x is addressed x[b*bs + l*ls + p], a 2-D
 * field x[b*bs + p]; "pred" fields are the prediction (t1), "in" fields the last frame of the input state (t0); pixels
 * p in [p0, p0 + np) (a latitude band of a decomposed forecast: the caller adds the bands' sums).  coef_a / coef_b: the L + 1
 * interface coefficients.  The per-pixel terms are formed in fp32 in the reference's order, the area-weighted sums in fp64
 * (deterministic two-level reduction).  scratch: wxf_budget_scratch_bytes(B), zeroed once.
 *   wxf_water_budget_sums : sums[b] = ( sum area dTWC/dt, sum area evaporation flux, sum area precipitation flux );
 *                           the caller forms ratio = (-TWC - E) / P and rescales precipitation with wxf_scale_planes.
 *   wxf_energy_budget_sums: sums[b] = ( sum area R_T, sum area F_S, sum area TE(t0), sum area TE(t1) );
 *                           ratio = (N (R_T - F_S) + TE0) / TE1.
 *   wxf_energy_fix_temperature: T_pred <- (E_level(t1) * ratio[b] - E_qgk(t1)) / CP(t1), in place.
 */
typedef struct {
  const float* q_pred;  int64_t q_pred_bs, q_pred_ls;
  const float* sp_pred; int64_t sp_pred_bs;
  const float* q_in;    int64_t q_in_bs, q_in_ls;
  const float* sp_in;   int64_t sp_in_bs;
  const float* precip;  int64_t precip_bs;
  const float* evapor;  int64_t evapor_bs;
  const float* area;
  const float* coef_a;
  const float* coef_b;
  int64_t p0, np;
  int32_t B, L;
  float n_seconds;
} WxfWaterDesc;

typedef struct {
  float* t_pred;                 /* 3-D prediction fields share pred3_bs / pred3_ls (channels of one NCHW tensor) */
  const float* q_pred;
  const float* u_pred;
  const float* v_pred;
  int64_t pred3_bs, pred3_ls;
  const float* sp_pred;          /* 2-D prediction fields share pred2_bs */
  const float* toa_up_solar;
  const float* toa_up_olr;
  const float* surf_down_solar;
  const float* surf_up_solar;
  const float* surf_down_lw;
  const float* surf_up_lw;
  const float* surf_sh;
  const float* surf_lh;
  int64_t pred2_bs;
  const float* t_in;             /* 3-D input fields share in3_bs / in3_ls */
  const float* q_in;
  const float* u_in;
  const float* v_in;
  int64_t in3_bs, in3_ls;
  const float* sp_in;       int64_t sp_in_bs;
  const float* toa_down_in; int64_t toa_down_bs;
  const float* gph_surf;         /* [pixels] surface geopotential */
  const float* area;
  const float* coef_a;
  const float* coef_b;
  int64_t p0, np;
  int32_t B, L;
  float n_seconds;
} WxfEnergyDesc;

int64_t wxf_budget_scratch_bytes(int B);
int wxf_water_budget_sums(const WxfWaterDesc* desc, double* sums, void* scratch, void* stream);
int wxf_energy_budget_sums(const WxfEnergyDesc* desc, double* sums, void* scratch, void* stream);
int wxf_energy_fix_temperature(const WxfEnergyDesc* desc, const float* ratio, void* stream);

/*
 * Autoregressive state update (update_x, credit/datasets/gen_2/channel_utils.py:253-291):
 * for every group g < n_groups: dst[:, dst_c0[g] : dst_c0[g]+len[g]] = src[:, src_c0[g] : src_c0[g]+len[g]]
 * on [B, C, plane] tensors (plane = T*H*W elements per channel).
 */
int wxf_copy_channels(float* dst, int dst_C, const float* src, int src_C, int B, int64_t plane,
                      const int32_t* dst_c0, const int32_t* src_c0, const int32_t* len, int n_groups, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * FuXi forecast step (credit/models/fuxi.py:454-506).  Its contractions run on the entry points above: CubeEmbedding
 * (fuxi.py:82-143) = wxf_conv_f16x2_tc with a 4x4 stride-4 kernel on the frame-folded padded input + wxf_layernorm;
 * DownBlock / UpBlock (fuxi.py:146-201) = wxf_conv_f16x2_tc + wxf_groupnorm_*; the Linear layers of the Swin-V2 stage
 * (timm SwinTransformerV2Stage, fuxi.py:250-260) and the dense head (fuxi.py:420, 484) = wxf_gemm_f16x2_tc.
 * The four entry points below are what FuXi needs on top.
 */

/*
 * Res-post-norm of a Swin-V2 block (timm SwinTransformerV2Block.forward: x = x + norm1(attn(x)); x = x + norm2(mlp(x))):
 *   o = res[m, :] + LayerNorm(x[m, :]) * g + b   (biased variance, eps inside the root, nn.LayerNorm)
 * written as fp32 (out, optional, may alias res) and / or fp16 hi/lo operand planes (out_hi/out_lo, optional).
 */
int wxf_layernorm_residual(const float* x, int ldx, const float* res, int ldr, float* out, int ldo, void* out_hi,
                           void* out_lo, int ldh, const float* g, const float* b, int64_t M, int d, float eps, void* stream);

/*
 * Swin-V2 window attention core (timm WindowAttention.forward + the shift / mask logic of SwinTransformerV2Block) for all
 * windows and heads of one block, exact fp32:
 *   qkv  : [B, H, W, ldq] fp32 (output of the qkv Linear incl. q_bias / v_bias), q = [0,d), k = [d,2d), v = [2d,3d),
 *          head h = channels h*dh .. h*dh+dh-1, dh = d / heads
 *   window (wy, wx), token (ty, tx) = source pixel ((wy*ws_h + ty + shift_h) mod H, (wx*ws_w + tx + shift_w) mod W)
 *          [torch.roll(x, (-shift_h, -shift_w)) then window_partition]; the output row returns to the same pixel
 *   S[i][j] = <q_i / max(|q_i|, 1e-12), k_j / max(|k_j|, 1e-12)> * logit_scale[h] + bias[h][i][j]
 *             + (region(i) != region(j) ? -100 : 0),   P = softmax_j(S),   out_i = sum_j P[i][j] v_j
 *   region = the 3 x 3 partition of the rolled image by the slices (0, -ws), (-ws, -shift), (-shift, end) per axis
 *            (only for a non-zero shift on that axis)
 *   bias        : [heads, L, L] fp32 = 16 * sigmoid(cpb_mlp(log-spaced relative coordinates)), L = ws_h*ws_w <= 64
 *   logit_scale : [heads] fp32 = exp(min(logit_scale, ln 100))
 *   output: fp16 hi/lo planes [B*H*W, ldh] (input of the proj GEMM) or fp32 [B*H*W, ldh] (exactly one of the two).
 *   mask_shift_h: row shift the -100 mask is built for; < 0 = shift_h.  A latitude band of a decomposed forecast does the
 *            row roll itself (its buffer starts shift rows into the band: shift_h = 0) and passes the true shift here on the
 *            one band that holds the wrapped window row, 0 elsewhere.
 */
int wxf_swin_window_attention(const float* qkv, int ldq, const float* bias, const float* logit_scale, void* out_hi,
                              void* out_lo, float* out_f32, int ldh, int B, int H, int W, int d, int heads, int ws_h,
                              int ws_w, int shift_h, int shift_w, int mask_shift_h, void* stream);

/*
 * Row gather with zero fill: dst[i, 0:d] = idx[i] >= 0 ? src[idx[i], 0:d] : 0, written as fp32 (dst, optional) and / or
 * fp16 hi/lo planes at column offset h_off (optional).  ZeroPad2d to a window multiple and its crop (fuxi.py:67-79,
 * 281-283, 288-289) and the channel concat with the shortcut (fuxi.py:292) are index lists for this pass.
 */
int wxf_gather_rows_ex(const float* src, int ld_src, const int32_t* idx, float* dst, int ld_dst, void* hi, void* lo, int ldh,
                       int h_off, int64_t n, int d, void* stream);

/*
 * Dense-head output -> prediction (fuxi.py:484-498): y is token-major [B, Lat, Lon, ph*pw*cp] where pixel (py, px) of a
 * patch owns the columns (py*pw + px)*cp + c, c < C <= cp (the head's weight rows re-ordered and padded at load);
 * un-patchify, crop rows [top, top+Hc) x cols [left, left+Wc), bilinear resize (align_corners = False) to Ho x Wo,
 * write NCHW [B, C, Ho, Wo]; only output rows [o0, o0 + n_out).  The buffer holds the patch rows [lat0, lat0 + Lat) of
 * the grid (lat0 = 0: all of it; a latitude band with its halo rows otherwise; top / Hc stay those of the whole grid).
 */
int wxf_unpatchify_unpad_resize_to_nchw(const float* y, float* out, int B, int C, int cp, int Lat, int Lon, int ph, int pw,
                                        int top, int left, int Hc, int Wc, int Ho, int Wo, int o0, int n_out, int lat0,
                                        void* stream);

/*
 * Autoregressive state update with a history window of T input frames (the gen2 rollout's slide,
 * credit/trainers/rollout_utils.py:288-311: drop the oldest time step, append the newest; the newest step is update_x,
 * credit/datasets/gen_2/channel_utils.py:253-291), in place on x [B, C, T, plane]:
 *   t < T-1 : x[b, c, t] = x[b, c, t+1]
 *   t = T-1 : c < n_prog: y[b, c, 0] (y is [B, Cy, Ty, plane]);  n_prog <= c < n_prog + n_dyn: forcing[b, c - n_prog]
 *             (forcing [B, n_dyn, plane], NULL = carried);  other channels (static) carried.
 * T = 1 is the plain update_x.
 */
int wxf_history_update(float* x, const float* y, const float* forcing, int B, int C, int T, int n_prog, int n_dyn, int Cy,
                       int Ty, int64_t plane, void* stream);

/*
 * Noise injection of the ensemble variant CrossFormerWithNoise (credit/models/wxformer/crossformer_ensemble.py:110-177;
 * StochasticDecompositionLayer.forward, credit/models/wxformer/stochastic_decomposition_layer.py:21-42):
 *     feature + noise_factor * eps * Linear(latent)[b, c] * modulation[c],   eps ~ N(0, 1) per element, latent ~ N(0, 1)^D
 * wxf_noise_coef   : coef[b, c] = factor[0] * (W[c, :] . latent[b, :] + bias[c]) * modulation[c]; latent == NULL draws it
 *                    (Philox, keyed by seed / *step_counter / site / b)
 * wxf_noise_inject : out[p, c] = x[p, c] + eps[p, c] * coef[b, c] on a pixel-major field ([B*HW, ld]), written as fp32 (out,
 *                    optional, may alias x) and / or fp16 hi/lo planes at column h_off; eps == NULL draws it (Philox).
 *                    eps, when given, is pixel-major [B*HW, C] (the tests feed the reference's recorded draws).
 * wxf_noise_step_advance : *step_counter += 1 (uint64, once per forward: a replayed CUDA graph draws fresh noise).
 */
int wxf_noise_coef(const float* latent, const float* W, const float* bias, const float* modulation, const float* factor,
                   float* coef, int B, int C, int D, uint64_t seed, const void* step_counter, int site, void* stream);
int wxf_noise_inject(const float* x, int ldx, float* out, int ldo, void* out_hi, void* out_lo, int ldh, int h_off,
                     const float* coef, const float* eps, int B, int64_t HW, int C, uint64_t seed, const void* step_counter,
                     int site, void* stream);
int wxf_noise_step_advance(void* step_counter, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Exchanges of the lat-lon domain decomposition over NVLink peer memory (csrc/wxf_peer.cu): what the reference does with
 * batch_isend_irecv per convolution (credit/domain_parallel/halo_exchange.py:56-67) and two all-reduces per GroupNorm
 * (credit/domain_parallel/layers.py:507-518) becomes stores into the consumer's memory plus an arrival counter.
 *
 * wxf_peer_alloc / _free      : one arena per rank (cudaMalloc, zero-filled)
 * wxf_peer_export / _open / _close : 64-byte cudaIpc handle of an arena; peers map it (lazy peer access)
 * wxf_peer_epoch_advance      : *epoch += 1 on the stream (once per forward)
 * wxf_peer_put                : copy nseg <= 8 segments (16-byte multiples) to peer (or local) addresses, then add 1 to
 *                               nsig <= 16 counters with system-scope release.  done_counter: a zeroed uint32 of this rank
 * wxf_peer_scatter_rows       : row i: src[src_idx[i], 0:d] -> dst_base[dst_rank[i]] + dst_idx[i]*ld_dst, then add 1 to
 *                               signals[r] for every rank r (world <= 8): the band <-> attention-unit re-layout
 * wxf_peer_wait               : block the stream until every counter >= *epoch (acquire; traps after ~seconds)
 * wxf_sum_rank_slots          : sums[i] = sum_r slots[r*n + i] in rank order (GroupNorm sums, bit-identical on all ranks)
 */
int wxf_peer_alloc(void** ptr, int64_t bytes);
int wxf_peer_free(void* ptr);
int wxf_peer_export(const void* ptr, void* handle64);
int wxf_peer_open(const void* handle64, void** ptr);
int wxf_peer_close(void* ptr);
int wxf_peer_epoch_advance(void* epoch, void* stream);
int wxf_peer_put(const void* const* src, void* const* dst, const int64_t* bytes, int nseg, void* const* signals, int nsig,
                 void* done_counter, void* stream);
int wxf_peer_scatter_rows(const float* src, int ld_src, const int32_t* src_idx, const int32_t* dst_rank, const int32_t* dst_idx,
                          void* const* dst_base, void* const* signals, int world, int ld_dst, int64_t n, int d,
                          void* done_counter, void* stream);
int wxf_peer_wait(void* const* signals, int nsig, const void* epoch, void* stream);
int wxf_sum_rank_slots(const double* slots, double* sums, int world, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WXFORMER_B200_H */
