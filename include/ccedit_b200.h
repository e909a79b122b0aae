/*
 * ccedit_b200 — C ABI of the B200-native (sm_100a) kernels behind CCEdit's denoising hot path.
 *
 * The reference (RuoyuFeng/CCEdit) is 100 % Python/PyTorch and has no FFI of its own; every entry point below
 * replaces one family of ATen call sites inside
 *   OpenAIWrapperControlLDM3DTV2V.forward   sgm/modules/diffusionmodules/wrappers.py:155-207
 *   ControlNet2D.forward                    sgm/modules/diffusionmodules/controlmodel.py:252-317
 *   ControlledUNetModel3DTV2V.forward       sgm/modules/diffusionmodules/controlmodel.py:471-550
 * The host-side mirror of those classes (ccedit_b200/*.py) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless stated; no torch types.
 *  - activations are fp16, channels-last: a video tensor is [B][T][H][W][C] (C contiguous), a token matrix [M][C].
 *  - every function is asynchronous on `stream` (a cudaStream_t passed as void*), returns 0 on success and a
 *    non-zero code otherwise; ccedit_last_error() returns the message of the last failure on the calling thread.
 *  - there is NO CPU fallback: without a CUDA device / sm_100a the calls fail with CCEDIT_ERR_CUDA.
 */
#ifndef CCEDIT_B200_H_
#define CCEDIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCEDIT_OK 0
#define CCEDIT_ERR_INVALID 1
#define CCEDIT_ERR_CUDA 2

const char* ccedit_last_error(void);
/* ABI version of this library (bumped when a struct below changes). */
int ccedit_abi_version(void);
/* Number of kernels launched by this library on the calling process since load (bench.py's gpu_launches). */
int64_t ccedit_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Tap-GEMM on tcgen05 tensor cores (TMA-staged operands, TMEM accumulators).
 *
 *   out[p, n] = epilogue( sum_{tap} sum_{c} A[p + tap, c] * W[n, tap, c] )
 *
 * A is a 5-D fp16 tensor (C, d1, d2, d3, d4) with C contiguous; p ranges over the 4-D output grid out_dims[0..3]
 * (same axes d1..d4); a tap is a 4-D integer offset; reads outside A return zero (TMA out-of-bounds fill), which is
 * the convolution zero padding.  This one primitive is
 *   nn.Linear / 1x1 conv / Conv1d k=1  (1 tap)             attention.py:383-388,118,136 ; openaimodel.py:706-709
 *   3x3 conv, stride 1, pad 1          (9 taps on d1,d2)   openaimodel.py:444-448,479-492,612-616,660-673
 *   3x3 conv, stride 2 (Downsample)    (9 taps over 4 parity planes on d3)  openaimodel.py:308-315,361-368
 *   temporal Conv1d k=3, pad 1         (3 taps on the T axis)               openaimodel.py:617-629,674-687
 * W is fp16 [N][ntaps][kpad] (kpad = C rounded up to 64, zero padded).
 * Epilogue (fp32): + bias[n] + rowbias[coord[rb_dim]/rb_div][n] ; optional SiLU ; optional GEGLU (value*gelu(gate),
 * attention.py:115-122, W rows interleaved per BN tile: BN/2 value rows then BN/2 gate rows) ; + res1 + res2 ;
 * fp16 store at out + sum_i coord_i*out_strides[i] + n.
 * ------------------------------------------------------------------------------------------------------------------ */
#define CCEDIT_GEMM_SILU 1
#define CCEDIT_GEMM_GEGLU 2
#define CCEDIT_MAX_TAPS 9

typedef struct ccedit_gemm_desc {
  const void* a;          /* fp16 A base                                                            */
  int32_t a_dims[5];      /* extents (C, d1, d2, d3, d4)                                            */
  int64_t a_strides[4];   /* element strides of d1..d4 (C stride is 1); each *2 bytes % 16 == 0     */
  int32_t box[4];         /* tile extents along d1..d4, product must be 128                         */
  int32_t out_dims[4];    /* output grid extents along d1..d4                                       */
  int32_t ntaps;          /* 1..9                                                                   */
  int32_t taps[CCEDIT_MAX_TAPS][4]; /* offsets along d1..d4 added to the output coordinate          */
  const void* w;          /* fp16 [N][ntaps*kpad]                                                   */
  int32_t n;              /* output channels of W (GEGLU: value+gate rows, out has n/2 channels)    */
  int32_t kpad;           /* per-tap padded K, multiple of 64, >= a_dims[0]                         */
  int32_t bn;             /* N tile: multiple of 16, 16..256, divides n                             */
  void* out;              /* fp16                                                                   */
  int64_t out_strides[4]; /* element strides of d1..d4 in out (channel stride 1)                    */
  const float* bias;      /* [n] or NULL                                                            */
  const float* rowbias;   /* [R][rb_ld] or NULL                                                     */
  int32_t rb_dim;         /* which of d1..d4 (0..3) indexes rowbias                                 */
  int32_t rb_div;         /* rowbias row = coord[rb_dim] / rb_div                                   */
  int32_t rb_ld;          /* rowbias row stride in floats (0 => n_out)                              */
  const void* res1;       /* fp16 or NULL                                                           */
  int64_t res1_strides[4];
  const void* res2;       /* fp16 or NULL                                                           */
  int64_t res2_strides[4];
  int32_t flags;          /* CCEDIT_GEMM_*                                                          */
  /* LayerNorm folded into this GEMM (attention.py:667-669 feeding to_q / to_k / to_v / ff.net.0): A holds the
   * UN-normalised tokens, W is gamma o W, bias is bias + W beta, and the epilogue applies
   *   out = rstd[m] * (acc - mean[m] * colsum[n]) + bias[n]        (before GEGLU / SiLU)
   * rowstats: fp32 (mean, rstd) pairs from ccedit_layernorm_stats, indexed by the d1 coordinate (out_dims[1..3] must be 1);
   * colsum: fp32 [n] = sum_k fp16(W[n][k]).  Both NULL => no fold.  Not combinable with rowbias / residuals. */
  const float* rowstats;
  const float* colsum;
  /* Statistics of THIS GEMM's output for the LayerNorm that consumes it (the next layer's rowstats without another
   * pass over the activation): fp32 [M][P][2] with P = 2 * n / bn; every epilogue thread writes (sum, sum of squares) of
   * the final fp32 values of its row over its column range.  ccedit_layernorm_stats_combine turns them into (mean, rstd).
   * Plain 2-D GEMMs only (no GEGLU / SiLU).  NULL => off. */
  float* stats_out;
  /* rowstats_slots > 1: rowstats is not (mean, rstd) but the [M][rowstats_slots][2] partial sums another GEMM wrote
   * through stats_out; the epilogue finishes them itself with eps = ln_eps over a_dims[0] channels (no combine kernel). */
  int32_t rowstats_slots;
  float ln_eps;
} ccedit_gemm_desc;

int ccedit_gemm(const ccedit_gemm_desc* d, void* stream);
/* Diagnostics: while device_buf (int64 [64][16], device memory) is set, CTA 0 of every ccedit_gemm launch records the
 * SM clock at its per-tile phases: [0] epilogue tile start, [1] epilogue global reads issued, [2] accumulator ready,
 * [3] stores issued, [4] accumulator released, [5] MMA warp owns the accumulator, [6] MMAs issued, [7] first operands
 * landed, [8..11] first 32-column block: TMEM load issued / returned / math done / stored.  NULL switches it off.  Used by tools/dev_gemm.py; never set on the product path. */
int ccedit_gemm_trace(int64_t* device_buf);

/* ------------------------------------------------------------------------------------------------------------------
 * Normalisation (fp32 statistics, fp16 in/out).
 * GroupNorm(32, C): spatial = per frame over (C/32)*HW  (util.py:296-302, attention.py:153-156; openaimodel.py:147)
 *                   temporal = per pixel over (C/32)*T   (openaimodel.py:157)
 * LayerNorm(C) per token (attention.py:667-669, 749-750).
 * ------------------------------------------------------------------------------------------------------------------ */
/* x,y: [F][HW][C]; gamma/beta fp32 [C]; silu!=0 fuses SiLU.  partial: fp32 scratch of 8192 + F*32*64 elements, private to
 * the stream: 8192 int32 arrival counters (2 per frame) that must be ZERO before the first call - the single-pass kernel,
 * used whenever all of its CTAs can be resident at once, re-arms them itself - followed by [F][32][32][2] partial sums. */
int ccedit_groupnorm_spatial(const void* x, void* y, const float* gamma, const float* beta, float* partial,
                             int32_t F, int32_t HW, int32_t C, float eps, int32_t silu, void* stream);
/* x,y: [B][T][HW][C]; statistics over (C/32, T) for every (b, hw). */
int ccedit_groupnorm_temporal(const void* x, void* y, const float* gamma, const float* beta, int32_t B, int32_t T,
                              int32_t HW, int32_t C, float eps, int32_t silu, void* stream);
/* x: [M][C] with row stride ldx (elements); y: [M][C] contiguous. */
int ccedit_layernorm(const void* x, int64_t ldx, void* y, const float* gamma, const float* beta, int64_t M, int32_t C,
                     float eps, void* stream);
/* Row statistics only: stats[m] = (mean, 1/sqrt(var + eps)) as fp32 pairs.  Used when the LayerNorm is folded into the
 * GEMM that consumes it (ccedit_gemm_desc.rowstats): LN(x) W^T = rstd (x (gamma o W)^T - mean colsum(gamma o W)) + beta W^T. */
int ccedit_layernorm_stats(const void* x, int64_t ldx, float* stats, int64_t M, int32_t C, float eps, void* stream);
/* partial: fp32 [M][P][2] written by ccedit_gemm (stats_out) for an activation of C channels -> stats [M][2]. */
int ccedit_layernorm_stats_combine(const float* partial, int32_t P, float* stats, int64_t M, int32_t C, float eps,
                                   void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Attention (F.scaled_dot_product_attention call site attention.py:444-448; scale = d^-1/2; 8 heads typical).
 * ------------------------------------------------------------------------------------------------------------------ */
/* Flash-style softmax(q k^T * scale) v over up to two concatenated key/value segments.
 * q: [Fq][Lq][*] rows of stride ldq, head h at column h*d;  o: [Fq][Lq][heads*d] with stride ldo.
 * segment s: k,v rows [Fkv][Lkv_s] with strides ldk/ldv; query frame f reads kv frame (f / kv_div_s) * kv_mul_s + kv_add_s.
 * d in {40, 80, 160} (C/8 of the SD-1.5 widths) or any multiple of 8 up to 160. */
typedef struct ccedit_attn_desc {
  const void* q; int64_t ldq; int64_t q_frame_stride;
  void* o; int64_t ldo; int64_t o_frame_stride;
  int32_t nseg;
  const void* k[2]; const void* v[2];
  int64_t ldk[2]; int64_t ldv[2]; int64_t kv_frame_stride[2];
  int32_t lkv[2]; int32_t kv_div[2]; int32_t kv_mul[2]; int32_t kv_add[2];
  int32_t frames; int32_t lq; int32_t heads; int32_t d;
  float scale;
} ccedit_attn_desc;
int ccedit_attention(const ccedit_attn_desc* d, void* stream);

/* Temporal attention: for every (b, pixel, head) attend over the T frames (attention.py:1182-1194, 758-761).
 * q: [B][T][HW][*] row stride ldq; k, v likewise (ldk, ldv); o: [B][T][HW][heads*d] stride ldo. T <= 64. */
int ccedit_temporal_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                              void* o, int64_t ldo, int32_t B, int32_t T, int32_t HW, int32_t heads, int32_t d,
                              float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Small / layout kernels.
 * ------------------------------------------------------------------------------------------------------------------ */
/* [B][Cin][T][H][W] (fp32 if src_f32 else fp16) -> channels-last fp16 [B][T][H][W][Cpad], channels >= Cin zeroed,
 * value = fp16( fma( fl32(v + pre), mul, add ) ).  The hint transform 1 - (h + 1) / 2 of wrappers.py:160-162 is
 * pre = 1, mul = -0.5, add = 1: the same two fp32 roundings torch makes, so the fused transform is bit-identical to
 * transforming in PyTorch and converting afterwards. */
int ccedit_ncthw_to_cl(const void* src, int32_t src_f32, void* dst, int32_t B, int32_t Cin, int32_t T, int32_t H,
                       int32_t W, int32_t Cpad, float pre, float mul, float add, void* stream);
/* UNet tail (controlmodel.py:550; openaimodel.py:1627-1632): y is the spatial out conv result, channels-last fp16
 * [B][T][HW][ldy] (first Cout channels valid); dst[b][c][t][hw] = y + bias_t[c] + sum_{dt,c'} wt[c][c'][dt] *
 * silu(y[b][t+dt-1][hw][c']) with zero padding in t; dst is fp32 if dst_f32 else fp16, layout [B][Cout][T][HW]. */
int ccedit_out_temporal(const void* y, int32_t ldy, const float* wt, const float* bias_t, void* dst, int32_t dst_f32,
                        int32_t B, int32_t Cout, int32_t T, int32_t HW, void* stream);
/* Sinusoidal timestep embedding (util.py:244-268): out fp32 [B][dim] = [cos(t f_i), sin(t f_i)]. */
int ccedit_timestep_embedding(const float* t, float* out, int32_t B, int32_t dim, float max_period, void* stream);
/* Small-M linear in fp32 (time_embed MLP openaimodel.py:1216-1223; ResBlock emb_layers :471-477,:652-658):
 * out[m][n] = act_out( sum_k act_in(x[m][k]) * w[n][k] + b[n] ), w fp16 [N][K], M <= 8; act: 0 none, 1 SiLU. */
int ccedit_linear_small(const float* x, const void* w, const float* b, float* out, int32_t M, int32_t N, int32_t K,
                        int32_t act_in, int32_t act_out, void* stream);
/* Split [F][H][W][C] into 4 parity planes [F][4][H/2][W/2][C] (plane = (h%2)*2 + w%2) for the stride-2 conv. */
int ccedit_parity_split(const void* x, void* y, int32_t F, int32_t H, int32_t W, int32_t C, void* stream);
/* Nearest x2 upsample of [F][H][W][C] -> [F][2H][2W][C] (openaimodel.py:256-260). */
int ccedit_upsample_nearest2x(const void* x, void* y, int32_t F, int32_t H, int32_t W, int32_t C, void* stream);
/* dst[.., dst_off + c] (row stride ldd) = a[.., c] (row stride lda) + b[.., c] (ldb; b may be NULL) for c < C. */
int ccedit_add_rows(const void* a, int64_t lda, const void* b, int64_t ldb, void* dst, int64_t ldd, int64_t M,
                    int32_t C, void* stream);
/* x[b][T/2][hw][c] += y[b][hw][c]  (img_control on the centre frame, controlmodel.py:529-535). */
int ccedit_add_center_frame(void* x, const void* y, int32_t B, int32_t T, int32_t HW, int32_t C, void* stream);

/* dst fp16 [n] = (half) src fp32 [n]  (the cast of c["crossattn"] to the model dtype, wrappers.py:164-166). */
int ccedit_to_half(const float* src, void* dst, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * ControlNet hint stem, first two layers fused (controlmodel.py:215-219: conv3x3(hint_channels->16) + SiLU +
 * conv3x3(16->16) + SiLU at the full hint resolution).  x: [F][H][W][8] fp16 (hint channels zero-padded to 8, as written by
 * ccedit_ncthw_to_cl), y: [F][H][W][16] fp16.  w0: fp16 [16][80], k = tap*8 + channel (tap = kh*3+kw), zero padded;
 * w1: fp16 [16][144], k = tap*16 + channel; b0, b1: fp32 [16].
 * ------------------------------------------------------------------------------------------------------------------ */
int ccedit_hint_stem01(const void* x, void* y, const void* w0, const float* b0, const void* w1, const float* b1,
                       int32_t F, int32_t H, int32_t W, void* stream);
/* Layers 2 and 3 of the same stem fused (controlmodel.py:220-223: conv3x3(16->32, stride 2) + SiLU + conv3x3(32->32) + SiLU).
 * x: [F][H][W][16] fp16 with even H, W (the output of ccedit_hint_stem01), y: [F][H/2][W/2][32] fp16.
 * w2: fp16 [32][144], k = tap*16 + channel; w3: fp16 [32][288], k = tap*32 + channel (16-byte aligned); b2, b3: fp32 [32]. */
int ccedit_hint_stem23(const void* x, void* y, const void* w2, const float* b2, const void* w3, const float* b3,
                       int32_t F, int32_t H, int32_t W, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * First stage (VAE): helpers next to ccedit_gemm / ccedit_groupnorm_spatial / ccedit_upsample_nearest2x.
 * ------------------------------------------------------------------------------------------------------------------ */
/* In-place softmax over the N columns of each of the M rows of a row-major fp16 matrix (row stride ld elements), fp32
 * arithmetic: the middle of AttnBlock's single-head d = 512 attention (model.py:161-201), run as GEMM -> softmax -> GEMM. */
int ccedit_softmax_rows(void* x, int64_t ld, int64_t M, int32_t N, void* stream);
/* channels-last fp16 [B][T][HW][ld] (first C channels valid) -> [B][C][T][HW], fp32 if dst_f32 else fp16: the decoded
 * image in the reference's "b c t h w" layout (autoencoder.py:341-342). */
int ccedit_cl_to_ncthw(const void* src, int32_t ld, void* dst, int32_t dst_f32, int32_t B, int32_t C, int32_t T, int64_t HW,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Text conditioner: FrozenCLIPEmbedder (sgm/modules/encoders/modules.py:358-420) wraps HuggingFace
 * transformers' CLIPTextModel (pinned 4.19.1 in the reference's requirements.txt:34; not vendored).  Its linears and
 * LayerNorms run on ccedit_gemm (LayerNorm folded) / ccedit_layernorm; these three entry points cover the rest.
 * ------------------------------------------------------------------------------------------------------------------ */
/* out fp16 [B*L][D] = token_emb[ids[b][l]] + pos_emb[l]  (CLIPTextEmbeddings; ids int64 [B][L], tables fp16 [V][D], [L][D]). */
int ccedit_embed_tokens(const int64_t* ids, const void* token_emb, const void* pos_emb, void* out, int32_t B, int32_t L,
                        int32_t D, int32_t V, void* stream);
/* x = x * sigmoid(1.702 x) in place, fp16 [n], n % 8 == 0  (CLIP's hidden_act "quick_gelu"). */
int ccedit_quick_gelu(void* x, int64_t n, void* stream);
/* Causal self-attention of a short sequence (CLIPAttention with the causal mask of CLIPTextTransformer): q, k, v fp16
 * [B][L][*] with row stride ld (heads * 64 channels each), o [B][L][heads*64] row stride ldo; L <= 128, d == 64. */
int ccedit_causal_attention_small(const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo, int32_t B,
                                  int32_t L, int32_t heads, int32_t d, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Sampler step, fused: the elementwise math of DPMPP2SAncestralSampler.sampler_step
 * (sampling.py:385-407) + DiscreteDenoiser / EpsScaling (denoiser.py:22-40, denoiser_scaling.py:16-22) + VanillaCFG
 * (guiders.py:25-29, 56-67) around the two network calls of a step.  x, x2, x_euler, noise, x_out: fp32 [n] latents
 * (n = B*4*T*h*w); xin2 / eps2: fp32 [2n] network input / output at CFG batch 2B (uncond half first); t2: int64 [2B]
 * timestep indices of the next network call.  sc: fp32 device table [steps][CCEDIT_SAMPLER_ROW] of per-step scalars
 * (columns below), step: device int32 selecting the row - so a captured CUDA graph of a whole step replays for every step.
 * Every operation is an individually rounded fp32 op in the reference's order: results are bit-identical to the unfused
 * PyTorch formulas given the same scalars.
 * ------------------------------------------------------------------------------------------------------------------ */
#define CCEDIT_SAMPLER_ROW 16
#define CCEDIT_SC_CIN1 0        /* 1 / sqrt(sigma_hat^2 + 1) of call 1 (sigma_hat = nearest entry of the 1000-sigma table) */
#define CCEDIT_SC_COUT1 1       /* -sigma_hat of call 1 */
#define CCEDIT_SC_IDX1 2        /* timestep index of call 1 */
#define CCEDIT_SC_SIGMA 3       /* sigma of the step */
#define CCEDIT_SC_DSIGMA 4      /* sigma_down - sigma */
#define CCEDIT_SC_M1 5          /* get_mult: to_sigma(s) / to_sigma(t) */
#define CCEDIT_SC_M2 6          /* expm1(-h / 2) */
#define CCEDIT_SC_CIN2 7
#define CCEDIT_SC_COUT2 8
#define CCEDIT_SC_IDX2 9
#define CCEDIT_SC_M3 10         /* to_sigma(t_next) / to_sigma(t) */
#define CCEDIT_SC_M4 11         /* expm1(-h) */
#define CCEDIT_SC_SIGMA_DOWN 12
#define CCEDIT_SC_NEXT_SIGMA 13
#define CCEDIT_SC_SIGMA_UP 14
#define CCEDIT_SC_EULER_ONLY 15 /* 1 when sum(sigma_down) < 1e-14 (sampling.py:390): no second network call */
/* xin2[0:n] = xin2[n:2n] = x * c_in1; t2 = idx1. */
int ccedit_sampler_prepare(const float* x, float* xin2, int64_t* t2, const float* sc, const int32_t* step, int64_t n,
                           int32_t B, void* stream);
/* eps2 = network output of call 1: d = CFG(eps2 * c_out1 + x); x_euler = x + (x - d) / sigma * (sigma_down - sigma);
 * x2 = m1 x - m2 d; xin2 = cat([x2 * c_in2] * 2); t2 = idx2. */
int ccedit_sampler_mid(const float* x, const float* eps2, float* x_euler, float* x2, float* xin2, int64_t* t2,
                       const float* sc, const int32_t* step, float cfg_scale, int64_t n, int32_t B, void* stream);
/* eps2 = network output of call 2 (ignored when the row is Euler-only): d2 = CFG(eps2 * c_out2 + x2);
 * x' = sigma_down > 0 ? m3 x - m4 d2 : x_euler; x_out = next_sigma > 0 ? x' + noise * s_noise * sigma_up : x'. */
int ccedit_sampler_final(const float* x, const float* x2, const float* x_euler, const float* eps2, const float* noise,
                         float* x_out, const float* sc, const int32_t* step, float cfg_scale, float s_noise, int64_t n,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CCEDIT_B200_H_ */
