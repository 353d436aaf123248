/* freefine_b200.h -- C ABI of libfreefine_b200.so (hand-written sm_100a CUDA; no torch types, no CPU fallback).
 *
 * Drop-in boundary for the denoising hot path of CIawevy/FreeFine.  The reference has no
 * FFI -- its hot path is eager PyTorch -- so each entry point names the reference Python function(s) whose
 * arithmetic it replaces (file:line relative to the upstream repo root).  INTEGRATION.md shows the ctypes binding
 * a maintainer adds inside those functions.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with h_ (host);
 *   - the caller owns all buffers; the library allocates nothing on the device and never synchronises;
 *   - `stream` is a cudaStream_t passed as void* (e.g. torch.cuda.current_stream().cuda_stream): every call is
 *     asynchronous on it and CUDA-graph capturable;
 *   - return value 0 = success, negative = error (FF_E_*); the message is in ff_last_error() (thread-local);
 *   - unsupported shapes / dtypes / misaligned pointers are errors, never a silent fallback.
 */
#ifndef FREEFINE_B200_H_
#define FREEFINE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FF_VERSION 100 /* 0.1.0 */

enum {
  FF_OK = 0,
  FF_E_INVALID = -1,     /* bad argument (null pointer, shape, alignment, enum)  */
  FF_E_UNSUPPORTED = -2, /* legal request outside what the kernels implement     */
  FF_E_CUDA = -3,        /* a CUDA runtime / driver call failed                  */
  FF_E_ARCH = -4         /* device is not sm_100 (tcgen05 / TMEM / TMA required) */
};

enum { FF_DT_F32 = 0, FF_DT_BF16 = 1, FF_DT_F16 = 2 };

int ff_version(void);
const char* ff_last_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * (a) masked, KV-injected flash attention  --  tcgen05 / TMEM / TMA
 *
 * Replaces, per attention-layer call: Attention_Modulator.Temporal_contextal_attention (src/utils/attention.py:
 * 1043-1091), _bg (:1284-1324), _compose (:1092-1140), style_align_share_attention (:1142-1192), the plain branch
 * of ca_forward (:395-404), get_cross_hidden_state (:808-837), together with the mask builders they call
 * (prepare_various_attention_mask :862-889, _compose :891-906, _for_bggen :907-925, prepare_sdsa_mask :940-951) --
 * whose [B*heads,S,S] additive masks are never materialised: the kernel reads bit-vectors of S bits.
 *
 * q,k,v are the UN-SPLIT projections the controller receives, bf16, logical shape [streams, S, heads*d] with the
 * channel index contiguous (a dense PyTorch [B,S,C] tensor).  Head h of a row is channels [h*d, (h+1)*d).
 *
 * Every (stream, head) gets a *plan*: out(q) = sum_p  weight_p * roww_p(q) * softmax_{k in allowed_p(q)}(q.k*scale) V
 * where pass p reads K,V of stream kv_stream (and optionally, under the SAME softmax, of kv_stream2 -- the
 * doubled [K_self;K_ref] context of the SSA/SDSA variants) and
 *     kb(k)       = key_mask < 0 ? 1 : bit k of bitmask row key_mask
 *     rb(q)       = row_mask < 0 ? 0 : bit q of bitmask row row_mask
 *     allowed(q,k)= key_mask < 0 ? 1 : kb(k) ^ KEY_INVERT ^ (ROW_XOR & rb(q))
 *                   (a segment WITHOUT a key mask admits every key: KEY_INVERT / ROW_XOR act on masked segments only)
 *     roww(q)     = ROW_WEIGHT ? rb(q) : 1
 * A row whose allowed set is EMPTY attends uniformly to every key (reference quirk Q4: the additive fill is
 * finfo.min, not -inf, attention.py:857); the kernel derives emptiness from mask_popcount, no host sync needed
 * (single-segment passes; with a second segment the plan builders never produce an empty allowed set).
 * The host-side controller (freefine_b200/attention.py) encodes quirk Q0 (head-parity mask tiling,
 * attention.py:859,881) simply by which plan it gives to which (stream, head).
 * ------------------------------------------------------------------------------------------------------------ */
#define FF_MAX_PASS 4

enum {
  FF_PASS_KEY_INVERT = 1u,  /* allowed = NOT keybit                                               */
  FF_PASS_ROW_XOR = 2u,     /* allowed ^= rowbit  (rows inside the region read the complement set) */
  FF_PASS_ROW_WEIGHT = 4u,  /* multiply the pass output by rowbit (compose: sum_i tgt_i(q) * O_i)   */
  FF_PASS_KEY2_INVERT = 8u, /* KEY_INVERT for the second segment                                  */
  FF_PASS_KEY_PREFIX = 16u, /* the caller SORTED the keys of kv_stream so that the popcount[key_mask] set keys   */
                            /* come first: keybit(k) = k < popcount[key_mask] (no bit-vector read; whole K/V     */
                            /* tiles become all-in / all-out and are skipped or processed without mask work)     */
  FF_PASS_KEY2_PREFIX = 32u /* same for the second segment                                        */
};

typedef struct FFAttnPass {
  int32_t kv_stream;  /* stream whose K,V this pass reads                          */
  int32_t key_mask;   /* bitmask row for the keys, -1 = every key                  */
  int32_t row_mask;   /* bitmask row for the query rows, -1 = none                 */
  uint32_t flags;     /* FF_PASS_*                                                 */
  float weight;       /* e.g. context_guidance, 1-context_guidance                 */
  int32_t kv_stream2; /* second KV segment under the same softmax, -1 = none       */
  int32_t key_mask2;  /* key bitmask row of the second segment, -1 = every key     */
  int32_t reserved;
} FFAttnPass;

typedef struct FFAttnHeadPlan {
  int32_t n_pass; /* 1..FF_MAX_PASS */
  int32_t reserved[3];
  FFAttnPass pass[FF_MAX_PASS];
} FFAttnHeadPlan;

typedef struct FFAttnArgs {
  const void* q;              /* bf16 [n_streams, s_q,  heads*head_dim]                                   */
  const void* k;              /* bf16 [n_kv_streams, s_kv, heads*head_dim]                                */
  const void* v;              /* v_dtype [n_kv_streams, s_kv, heads, v_head_stride]: the staging of ff_kv_gather_cast */
  void* out;                  /* out_dtype [n_streams, s_q, heads*head_dim]                               */
  const FFAttnHeadPlan* plan; /* DEVICE array [n_streams*heads], index stream*heads+head                  */
  const uint32_t* bitmasks;   /* DEVICE [n_masks, mask_words] bit i of word i/32 <=> token i; may be NULL */
  const int32_t* mask_popcount; /* DEVICE [n_masks] number of set bits among the first s_kv tokens          */
  int32_t n_streams, n_kv_streams, heads, head_dim; /* head_dim % 8 == 0, <= 160                          */
  int32_t s_q, s_kv;
  int32_t n_masks, mask_words;                      /* mask_words >= ceil(max(s_q,s_kv)/32)               */
  int32_t out_dtype;                                /* FF_DT_BF16 or FF_DT_F32                            */
  float scale;                                      /* softmax scale (head_dim^-0.5)                      */
  int32_t v_dtype;   /* FF_DT_F16: P = ONE fp16 tensor-core operand (fast path, ~3e-4 max-abs vs fp32);          */
                     /* FF_DT_BF16: P = hi+lo bf16 operand pair (2x PV tensor work, ~3e-5)                         */
  int32_t v_head_stride; /* must equal ff_attn_v_head_stride(head_dim)                                            */
} FFAttnArgs;

int ff_attn_masked_kv(const FFAttnArgs* h_args, void* stream);

/* Plain attention out = softmax(q k^T * scale) v for SHORT key sequences: the plain branch of ca_forward
 * (src/utils/attention.py:395-404) as it runs for the text cross-attention of every transformer block (77 keys) and the
 * plain 8 x 8 / 16 x 16 self-attention (64 / 256 keys).  q [n_streams, s_q, heads*head_dim], k / v [n_streams, s_kv,
 * heads*head_dim] bf16 (the un-split projections, stream i attends to K/V stream i; v is read as is -- no
 * ff_kv_gather_cast staging), out [n_streams, s_q, heads*head_dim] bf16 or f32.  s_kv <= 256, head_dim in {40, 80, 160}
 * (and 8 with s_kv <= 128: the reference's golden vectors), scale > 0.  K and V of a (stream, head) stay in shared memory
 * while Q streams through: the launch is bound by Q in + O out, where the tcgen05 kernel below pays ~7 us of per-CTA
 * pipeline set-up per 128-row tile (csrc/attn_smallkv.cu).  Numerics as ff_attn_masked_kv with v_dtype = FF_DT_F16 (P and V
 * fp16 tensor-core operands, |v| >= 65504 saturates).                                                                  */
int ff_attn_plain_smallkv(const void* q, const void* k, const void* v, void* out, int32_t n_streams, int32_t heads,
                          int32_t head_dim, int32_t s_q, int32_t s_kv, float scale, int32_t out_dtype, void* stream);

/* Channels per head of the V staging for a given head_dim (48 for 40, 96 for 80, 176 for 160). */
int ff_attn_v_head_stride(int32_t head_dim);

/* K/V staging for ff_attn_masked_kv in one pass over HBM: optional row gather (row_index[r] = source token row of
 * output row r, over the flattened [n_kv_streams*s_kv] rows; NULL = identity) of K (bf16 copy; k and k_out may both
 * be NULL to skip it) and V, with V (bf16 in) written as v_out_dtype -- FF_DT_F16 (saturating at +-65504) or
 * FF_DT_BF16 -- into the layout the kernel reads: v_out [rows, heads, ff_attn_v_head_stride(head_dim)], channels
 * [0,head_dim) = V, channel head_dim = 1.0 (its P.V column is the softmax denominator), channels above = 0.
 * The gather is how the host sorts the keys of a masked stream "source-mask keys first" (FF_PASS_KEY_PREFIX); the
 * reference has no counterpart (it materialises [B*H,S,S] masks instead, attention.py:856-889).               */
int ff_kv_gather_cast(const void* k, const void* v, const int64_t* row_index, void* k_out, void* v_out,
                      int32_t v_out_dtype, int64_t rows, int32_t heads, int32_t head_dim, void* stream);

/* Nearest-neighbour down-sample of n full-resolution uint8 masks [n,H,W] to [h,w] and bit-packing.
 * Replaces process_mask_before_attention (attention.py:841-855) + the flatten that follows: index
 * floor(i*H/h) in fp32 as ATen's `nearest`; if a mask's max is > 1 the reference rescales it by its max and
 * truncates back to uint8, so only pixels EQUAL to the max survive (kept).  bit = pixel != 0 afterwards.
 * bits: [n, words] (words >= ceil(h*w/32), padding bits written 0); popcount: [n].                          */
int ff_mask_downsample_pack(const uint8_t* masks, int32_t n, int32_t H, int32_t W, int32_t h, int32_t w,
                            uint32_t* bits, int32_t words, int32_t* popcount, void* stream);

/* Mask preparation of E edits in one launch.  Replaces FreeFinePipeline.prepare_various_mask (src/demo/model.py:1432-1512)
 * incl. its dilate_mask calls (:927-934: cv2.dilate with a 15 x 15 / 30 x 30 ones kernel, anchor k/2), prepare_tensor_mask
 * (:1622-1639, `> 0` binarisation) and the nearest down-sampling of the completion / local-variance masks to the latent
 * grid (:1505-1511).  All inputs uint8 [E,H,W], any non-zero = set: shifted (target mask), ori (source-object mask), draw
 * (user completion region; may be NULL with use_auto_draw), cons (constraint area; may be NULL when neither flag is set).
 * Outputs uint8: fg / shifted_out / ori_out [E,H,W] in {0,1} (the controller's fg_retain / fg_retain_st2 / fg_ref masks),
 * comp_lat / lvar_lat [E,h,w] in {0,1,2} -- the uint8 wrap-around of the reference's `cons - ori` and `1 - x` (quirk Q1)
 * is reproduced bit for bit.  H % h == 0, W % w == 0, W % 4 == 0.                                                    */
int ff_mask_prep(const uint8_t* shifted, const uint8_t* ori, const uint8_t* draw, const uint8_t* cons, int32_t E,
                 int32_t H, int32_t W, int32_t h, int32_t w, int32_t use_auto_draw, int32_t reduce_inp_artifacts,
                 uint8_t* fg, uint8_t* shifted_out, uint8_t* ori_out, uint8_t* comp_lat, uint8_t* lvar_lat, void* stream);

/* out[n] = cv2.dilate(mask[n] != 0, ones(k, k)) (anchor k/2, outside = 0), uint8 [N,H,W] -> {0,1}: dilate_mask
 * (model.py:927-934, src/utils/vis_utils.py:340-347).  W % 4 == 0.                                                    */
int ff_dilate_mask(const uint8_t* mask, uint8_t* out, int32_t N, int32_t H, int32_t W, int32_t k, void* stream);

/* y[n, 2h+dy, 2w+dx, c] = x[n, h, w, c] (nearest 2x up-sampling, F.interpolate(scale_factor=2, mode="nearest") of diffusers'
 * Upsample2D) and out[m, :] = [a[m, :Ca], b[m, :Cb]] (torch.cat([hidden_states, res], dim=1) of the up blocks) on dense bf16
 * NHWC tensors, any 16-bit type really: pure 128-bit copies.  C, Ca, Cb % 8 == 0.                                          */
int ff_upsample2x_nhwc(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, void* stream);
int ff_concat_nhwc(const void* a, const void* b, void* out, int64_t M, int32_t Ca, int32_t Cb, void* stream);

/* out[M,N] = x[M,K] . w[N,K]^T (+ bias[N]) (+ res[M,N]), bf16 row-major, fp32 accumulation, ONE cuBLASLt GEMM (bias epilogue +
 * beta * C): a plain library GEMM behind the C ABI.  Replaces `Linear(h) + hidden_states` at the end of every sub-block of
 * diffusers' BasicTransformerBlock / Transformer2DModel as walked by override_forward (src/utils/attention.py:13-223), i.e. the
 * GEMM + the separate elementwise add.  bias / res may be NULL; out may alias res (measured: cuBLASLt then rounds the GEMM
 * result to bf16 before adding res, like the eager pair; out of place it rounds once).  N % 8 == 0, K % 8 == 0.  `workspace`
 * (ws_bytes, 16-byte aligned, may be 0) is caller-owned cuBLASLt scratch.  Links libcublasLt.so.12.                          */
int ff_linear_bias_residual(const void* x, const void* w, const void* bias, const void* res, void* out, int64_t M, int32_t N,
                            int32_t K, void* workspace, int64_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (b) fused affine warp + resample + mask-guided blend
 *
 * Replaces wrapAffine_tensor (src/utils/geo_utils.py:304-341: F.affine_grid + F.grid_sample, padding 'zeros',
 * align_corners=False) for the image (bilinear) and the mask (nearest), and the np.where blend of re_edit_2d
 * (src/utils/vis_utils.py:252-256,272):   out = warped_mask != 0 ? warped_src : bg.
 * src [N,C,H,W], bg/out [N,C,dH,dW] (dtype f32 or bf16), theta [N,2,3] f32 in normalised coordinates (the output
 * of param2theta, geo_utils.py:292-302), mask_src [N,H,W] u8 or NULL (then out = warped_src, `bg` ignored),
 * mask_out [N,dH,dW] u8 (0/1) or NULL.  mode: 0 bilinear, 1 nearest (for src).                                  */
int ff_warp_affine_blend(const void* src, const float* theta, const uint8_t* mask_src, const void* bg, void* out,
                         uint8_t* mask_out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t dH, int32_t dW,
                         int32_t mode, int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * (c) fused classifier-free guidance + DDIM / local-DDPM step
 *
 * ff_ddim_cfg_step replaces model.py:605-611 (local CFG) + ctrl_step (model.py:134-198) for n_edits edits whose
 * UNet output is laid out [n_edits, 4, C, h, w] = [u_e, u_r, c_e, c_r] per edit (model.py:594) and whose latents
 * are [n_edits, 2, C, h, w] = [edit, ref].  fp32, reference operation order, no FMA contraction (bit-exact with
 * the CPU oracle).  cfg_mask / var_mask: u8 [n_edits, h, w] with uint8 wrap-around arithmetic (quirk Q1);
 * cfg_mask NULL = plain CFG (model.py:608).  noise: [n_edits,2,C,h,w] (the randn_tensor draw of model.py:186,
 * consumed only when sigma != 0); pred_x0 may be NULL.  Scalars are computed by the host in fp32 exactly as the
 * reference does from alphas_cumprod:
 *   sqrt_1m_at = (1-a_t)**.5, sqrt_at = a_t**.5, sqrt_ap = a_prev**.5, c_ddim = (1-a_prev)**.5,
 *   c_ddpm = (1-a_prev-sigma**2)**.5, sigma = eta*variance**.5  (edit stream; the ref stream uses sigma 0).       */
int ff_ddim_cfg_step(const float* eps4, const float* x, const float* noise, const uint8_t* cfg_mask,
                     const uint8_t* var_mask, float guidance_scale, float sqrt_1m_at, float sqrt_at, float sqrt_ap,
                     float c_ddim, float c_ddpm, float sigma, float* x_prev, float* pred_x0, int32_t n_edits,
                     int32_t C, int32_t h, int32_t w, void* stream);

/* ff_ddim_step = ctrl_step (model.py:134-198) alone, for callers that already combined the guidance: eps2
 * [n_edits, 2, C, h, w] = [edit, ref].  Same arithmetic as above from pred_x0 on.                                 */
int ff_ddim_step(const float* eps2, const float* x, const float* noise, const uint8_t* var_mask, float sqrt_1m_at,
                 float sqrt_at, float sqrt_ap, float c_ddim, float c_ddpm, float sigma, float* x_prev,
                 float* pred_x0, int32_t n_edits, int32_t C, int32_t h, int32_t w, void* stream);

/* Composition form (forward_sampling_compose, model.py:407-431): the UNet batch of an edit is [e, r_1..r_N, c_e]
 * (streams_per_edit = N+2), guidance uses streams 0 and N+1, and ONLY the edit stream is stepped (single-stream
 * ctrl_step with the local-DDPM mask): x / noise / x_prev [n_edits, C, h, w].                                       */
int ff_ddim_cfg_step_compose(const float* eps, int32_t streams_per_edit, const float* x, const float* noise,
                             const uint8_t* cfg_mask, const uint8_t* var_mask, float guidance_scale, float sqrt_1m_at,
                             float sqrt_at, float sqrt_ap, float c_ddim, float c_ddpm, float sigma, float* x_prev,
                             float* pred_x0, int32_t n_edits, int32_t C, int32_t h, int32_t w, void* stream);

/* ff_ddim_inv_step replaces inv_step (model.py:109-132): x_next = sqrt_an*((x - sqrt_1m_at*eps)/sqrt_at) +
 * c_next*eps over n elements; pred_x0 may be NULL.                                                               */
int ff_ddim_inv_step(const float* eps, const float* x, float sqrt_1m_at, float sqrt_at, float sqrt_an, float c_next,
                     float* x_next, float* pred_x0, int64_t n, void* stream);

/* Local cross-attention blend (modulate_local_cross_attn, attention.py:1381-1383, after the 77-key attention):
 * hs [n_edits,4,S,C] (f32 or bf16) in place -> [u_e, u_r, region*c_e + (1-region)*u_e, u_r]; region: bit-vector
 * row `region_mask` of `bitmasks` per edit ([n_edits] indices).                                                  */
int ff_cross_region_blend(void* hs, const uint32_t* bitmasks, int32_t mask_words, const int32_t* region_mask,
                          int32_t n_edits, int32_t S, int32_t C, int32_t dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * UNet-body glue in channels-last layout ("bf16/channels_last UNet").  HBM-bound.
 *
 * The reference runs diffusers' UNet blocks eagerly (src/utils/attention.py:13-223 `override_forward` walks
 * down_blocks / mid_block / up_blocks; diffusers ResnetBlock2D, Transformer2DModel, BasicTransformerBlock, GEGLU):
 * conv -> +bias -> +temb -> GroupNorm -> SiLU, permute copies around every transformer, LayerNorm x3, gelu * x.
 * These entry points fuse that elementwise / normalisation traffic; every activation is bf16 [N, H*W, C] with the
 * channel index contiguous (torch channels_last == the [B, S, C] token layout of the attention layers).
 * ------------------------------------------------------------------------------------------------------------ */

/* y = act(GroupNorm_G(x + add_nc[n, c])) : x, y bf16 [N, HW, C]; add_nc fp32 [N, C] with row stride add_ld >= C elements
 * (a column block of a wider matrix: all time-embedding projections of a UNet call come out of one GEMM), or NULL (conv
 * bias + projected time embedding of ResnetBlock2D); gamma, beta bf16 [C]; fp32 statistics over (HW, C/G) per (n, group), biased
 * variance, eps inside the square root (torch.nn.GroupNorm); silu != 0 fuses x*sigmoid(x).  `workspace`: at least
 * ff_group_norm_ws_bytes(N, G) bytes, 16-byte aligned, ZERO-FILLED before its first use (cudaMemset once); every call
 * leaves it zero-filled where that matters (it holds the inter-CTA barrier counters of the single-read kernel, which
 * reads x from HBM exactly once), so one buffer serves all calls of a stream.  Not to be shared by calls that may run
 * concurrently (different streams: one workspace each).  C % 8 == 0, C % G == 0.                                  */
int64_t ff_group_norm_ws_bytes(int32_t N, int32_t G);
int ff_group_norm_nhwc(const void* x, const float* add_nc, int64_t add_ld, const void* gamma, const void* beta, void* y,
                       void* workspace, int32_t N, int32_t HW, int32_t C, int32_t G, float eps, int32_t silu, void* stream);

/* out[m, c] = h[m, c] + bias[c] + res[m, c] : bf16 [M, C]; bias bf16 [C] or NULL, res bf16 [M, C] or NULL (conv
 * bias + skip connection of ResnetBlock2D / Down- / Upsample2D in one pass); out may alias h or res.            */
int ff_bias_residual_nhwc(const void* h, const void* bias, const void* res, void* out, int64_t M, int32_t C,
                          void* stream);

/* GEGLU (diffusers attention.py GEGLU.forward: hidden, gate = proj(x).chunk(2, -1); hidden * gelu(gate)):
 * h bf16 [M, 2F] -> out bf16 [M, F] = h[:, :F] * gelu_erf(h[:, F:]); gelu(gate) is rounded to bf16 before the
 * multiply like the eager pair of kernels does; erf is evaluated as 1 - erfc with the Abramowitz-Stegun 7.1.26
 * rational approximation (absolute error of gelu < 5e-7, far below the bf16 rounding).  F % 8 == 0.             */
int ff_geglu(const void* h, void* out, int64_t M, int32_t F, void* stream);

/* y = LayerNorm(x) over the last dim: x, y bf16 [M, C], gamma / beta bf16 [C], fp32 statistics, biased variance
 * (torch.nn.LayerNorm).  C % 8 == 0, C <= 2048.                                                                 */
int ff_layer_norm(const void* x, const void* gamma, const void* beta, void* y, int64_t M, int32_t C, float eps,
                  void* stream);

/* Debugging aid, not used by the product path: progress trace of ff_attn_masked_kv into a device-visible buffer of
 * 8 uint32 per CTA (see csrc/attn_tcgen05.cu); NULL switches it off.                                              */
int ff_debug_set_trace(void* device_visible_ptr);
/* Debugging aid (library built with -DFF_TIMELINE only): per-tile clock64 stamps of two CTAs, 2 x 2 x 64 x 8 uint64. */
int ff_debug_set_timeline(void* device_ptr);

#ifdef __cplusplus
}
#endif
#endif /* FREEFINE_B200_H_ */
