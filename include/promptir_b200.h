/* promptir_b200 -- C ABI of the B200 (sm_100a) kernels behind the PromptIR restoration forward.
 *
 * The reference (kongwanbianjinyu/PromptIR) is pure PyTorch: its "operator API" for this path is the
 * nn.Module net/model.py:PromptIR, which dispatches ~1800 ATen/cuDNN/cuBLAS kernels per forward.  This
 * library is what a maintainer binds instead (ctypes stub in INTEGRATION.md): plain pointers, sizes and a
 * cudaStream_t -- no torch types.  Each entry point names the reference lines it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or a negative PIR_ERR_* code; pir_last_error() returns a
 *     thread-local human-readable message.  No exceptions cross the boundary, nothing is allocated,
 *     nothing synchronises: kernels are enqueued on the stream that is passed in.
 *   - activations are NHWC ("channels last") 16-bit: element (b, y, x, c) of a tensor with `pitch`
 *     elements per pixel lives at base + b*bstride + (y*W + x)*pitch + c.  A tensor may be a channel slice
 *     of a wider buffer (pitch > C): that is how torch.cat is folded away (model.py:341,347,353,359,365,370).
 *   - dtype selects the 16-bit storage/operand type: PIR_DTYPE_BF16 or PIR_DTYPE_FP16; all accumulation,
 *     LayerNorm statistics, softmax and norms are fp32.
 *   - channel counts and pitches of 16-bit tensors are multiples of 8 (16-byte rows for TMA).
 */
#ifndef PROMPTIR_B200_H_
#define PROMPTIR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIR_ABI_VERSION 1

enum { PIR_DTYPE_FP16 = 0, PIR_DTYPE_BF16 = 1 };
enum { PIR_OK = 0, PIR_ERR_ARG = -1, PIR_ERR_CUDA = -2, PIR_ERR_DRIVER = -3, PIR_ERR_UNSUPPORTED = -4 };

/* epilogue store modes of pir_gemm */
enum {
  PIR_OUT_NHWC16 = 0,       /* 16-bit NHWC rows (optional residual add)                                  */
  PIR_OUT_UNSHUFFLE16 = 1,  /* nn.PixelUnshuffle(2) folded into the store        (model.py:165)          */
  PIR_OUT_SHUFFLE16 = 2,    /* nn.PixelShuffle(2) folded into the store          (model.py:175)          */
  PIR_OUT_FINAL_NCHW32 = 3, /* fp32 NCHW, "+ inp_img"                            (model.py:377)          */
  PIR_OUT_NHWC32 = 4        /* fp32 NHWC rows                                                            */
};
enum { PIR_LN_NONE = 0, PIR_LN_WITHBIAS = 1, PIR_LN_BIASFREE = 2 };

int pir_abi_version(void);
const char* pir_last_error(void);
/* device sanity: returns 0 iff the current device is compute capability 10.x */
int pir_check_device(void);

/* ---- pointwise / dense-3x3 convolution as a tcgen05 GEMM ----------------------------------------
 * Replaces nn.Conv2d(k=1) (model.py:88,92,111,113,294-313), the preceding LayerNorm (model.py:39-41,
 * 60-63; folded: weights carry gamma, ln_s/vec_t carry the mean/bias terms), nn.Conv2d(k=3)
 * (model.py:164,174,223,320) and the residual adds (model.py:193-194,377).
 *   out[b,p,n] = epi( sum_{tap,k} a[b, p+tap, k] * w[b?, n, tap*Kpad + k] )
 * w is packed [w_batched ? B : 1][N][taps*Kpad] 16-bit, Kpad = ceil(K/64)*64, zero padded.         */
typedef struct PirGemm {
  int32_t dtype;
  int32_t B, H, W;          /* batch and spatial size of the A tensor                                     */
  int32_t K, N, taps;       /* input channels, output channels, 1 or 9                                    */
  int32_t w_batched;        /* 1: one weight matrix per image (attention fold)                            */
  int32_t out_mode, ln_mode;
  const void* a; int64_t a_pitch, a_bstride;
  const void* w;
  void* out; int64_t out_pitch, out_bstride;      /* bstride in elements of the output type             */
  const void* res; int64_t res_pitch, res_bstride; /* optional residual (NHWC16 only), may alias out     */
  const float* ln_s;        /* [N] row sums of the 16-bit gamma-scaled weights (ln_mode != 0)             */
  const float* vec_t;       /* [N] additive per-channel vector (W.beta and/or conv bias) or NULL          */
  const float* img;         /* fp32 NCHW network input (PIR_OUT_FINAL_NCHW32)                             */
} PirGemm;
int pir_gemm(const PirGemm* d, void* stream);

/* ---- depthwise 3x3 (pad 1), optionally fused with the GDFN gate ----------------------------------
 * gate == 0: out[b,p,c] = dw(in)[c] (+bias[c]),                c < C      (model.py:112,120  qkv_dwconv)
 * gate == 1: out[b,p,c] = gelu_erf(dw(in)[c]) * dw(in)[C + c], c < C      (model.py:90,96-97 GDFN)
 *            (in has 2*C channels).  w is [3][3][Cin] 16-bit (tap-major), bias fp32 [Cin] or NULL.
 * gate == 2: backward of gate == 1 with the stencil recomputed: y = dw(in) (fp32, never stored),
 *            out[c] = dg[c] y2 (Phi(y1) + y1 phi(y1)),  out[C + c] = dg[c] y1 Phi(y1)    (out has 2*C channels; training).
 *            Phi comes from the forward gate's logistic-polynomial fit (|dPhi| <= 5.1e-5, a tenth of a 16-bit ulp of the result). */
typedef struct PirDwConv {
  int32_t dtype, gate;
  int32_t B, H, W, C;       /* C = output channels; input channels = C (gate 0) or 2*C (gate 1)           */
  const void* in; int64_t in_pitch, in_bstride;
  const void* w; const float* bias;
  void* out; int64_t out_pitch, out_bstride;
  const void* dg; int64_t dg_pitch, dg_bstride;   /* gate == 2 only */
} PirDwConv;
int pir_dwconv3x3(const PirDwConv* d, void* stream);

/* ---- fused LayerNorm -> 1x1 conv -> depthwise 3x3 (-> GELU gate) -----------------------------------
 * gate == 0: out[b,p,n] = dw3x3( W . LN(x) )[n],                                   n < N      (model.py:60-63,111-112,120)
 * gate == 1: out[b,p,n] = gelu_erf(dw3x3(W . LN(x))[n]) * dw3x3(W . LN(x))[N + n], n < N      (model.py:60-63,88-90,96-97)
 * Same function as pir_gemm (ln fold) followed by pir_dwconv3x3, but the pre-conv tensor (N resp. 2N channels per
 * pixel) never reaches HBM: it is converted from the fp32 accumulators to fp16 for the stencil (values saturate at +-65504;
 * shared-memory tile in pwdw.cu, straight from tensor memory in pwdwt.cu) and the LayerNorm is applied to the x tile
 * before the GEMM (normalised rows rounded to the 16-bit type; gamma lives in w, beta in vec_t).  w: packed
 * [N or 2N][Kpad] 16-bit like pir_gemm; vec_t: fp32 [N or 2N] additive vector (W.beta + conv bias) or NULL;
 * dw_w: [3][3][N or 2N] fp16 (always IEEE half, also when dtype is bf16); dw_bias fp32 or NULL.
 * pir_pwdw_supported() tells whether (C, N, gate) fits the shared-memory plan; otherwise use the two calls.     */
typedef struct PirPwDw {
  int32_t dtype, gate, ln_mode;
  int32_t B, H, W, C;       /* input tensor                                                                */
  int32_t N;                /* output channels (the 1x1 conv produces N, or 2N when gate == 1)             */
  const void* a; int64_t a_pitch, a_bstride;
  const void* w;
  const float* vec_t;
  const void* dw_w; const float* dw_bias;
  void* out; int64_t out_pitch, out_bstride;
  /* gate == 0 only, optional: channels n >= split are written to out2[..., n - split] instead (MDTA: q|k and v as two dense
   * tensors, model.py:121, so the Gram and attn.v kernels stream contiguous rows).  out2 == NULL or split <= 0: one tensor.   */
  void* out2; int64_t out2_pitch, out2_bstride;
  int32_t split;
} PirPwDw;
int pir_pwdw_supported(int32_t C, int32_t N, int32_t gate);
int pir_pwdw_split_supported(int32_t C, int32_t N);   /* gate == 0 with out2 / split */
int pir_pwdw(const PirPwDw* d, void* stream);

/* ---- MDTA: transposed (channel) attention -----------------------------------------------------------
 * Step 1, pir_mdta_gram: split-K tcgen05 Gram of q and k over the pixels of each image plus the squared
 * L2 norms of every q/k channel (model.py:123-130).  qkv is the NHWC 16-bit output of the qkv dwconv
 * (q = channels [0,C), k = [C,2C), v = [2C,3C)).  Partials go to `ws` (fp32, pir_mdta_ws_floats()).
 * Step 2, pir_mdta_finalize: deterministic reduction of the partials, cosine logits * temperature,
 * softmax over the key channels of each head (model.py:130-131), then the attention matrix is folded into
 * project_out:  wfold[b][o][h*c + j] = sum_i Wo[o][h*c + i] * A[b,h][i][j]   (model.py:133-137), so that
 * attn@v followed by project_out is ONE pir_gemm on v with per-image weights.  wo is fp32 [C][C].
 * wfold is 16-bit [B][C][Kpad], Kpad = ceil(C/64)*64, zero padded.                                      */
typedef struct PirMdta {
  int32_t dtype;
  int32_t B, HW, C, heads;
  int32_t splits;           /* split-K factor chosen by pir_mdta_splits()                                 */
  const void* qkv; int64_t qkv_pitch, qkv_bstride;
  float* ws;                /* workspace                                                                   */
  const float* temperature; /* [heads]                                                                     */
  const float* wo;          /* [C][C] fp32 project_out weight                                              */
  void* wfold;              /* [B][C][Kpad] 16-bit                                                         */
} PirMdta;
int pir_mdta_splits(int32_t B, int32_t HW, int32_t C);
int64_t pir_mdta_ws_floats(int32_t B, int32_t C, int32_t splits);
int pir_mdta_gram(const PirMdta* d, void* stream);
int pir_mdta_finalize(const PirMdta* d, void* stream);
/* kernels pir_mdta_finalize enqueues for this shape: 1 (fused softmax + fold, head dim <= 192) or 2 */
int pir_mdta_finalize_kernels(int32_t C, int32_t heads);

/* ---- PromptGenBlock (model.py:226-232): pool -> linear -> softmax -> weighted prompt sum -> bilinear --
 * x: NHWC 16-bit [B,H,W,C] feature; prompt: fp32 [L][S][S][D] (repacked prompt_param); lin_w fp32 [L][C],
 * lin_b fp32 [L]; out: NHWC 16-bit [B,H,W,D] (the conv3x3 that follows is a pir_gemm with taps = 9).
 * ws: fp32 workspace of pir_prompt_ws_floats() elements (per-chunk pooled sums; pir_prompt_bwd reads them).  weights_out (optional)
 * receives softmax [B][L]. */
typedef struct PirPrompt {
  int32_t dtype;
  int32_t B, H, W, C;       /* feature map                                                                 */
  int32_t L, D, S;          /* prompt components, channels, stored size                                    */
  const void* x; int64_t x_pitch, x_bstride;
  const float* prompt; const float* lin_w; const float* lin_b;
  void* out; int64_t out_pitch, out_bstride;
  float* ws; float* weights_out;
  int32_t align_corners;    /* bilinear rule: 0 = net/model.py:231 (PromptGenBlock), 1 = prompt_xrestormer.py:350 (PromptBlock)  */
  int32_t* sync;            /* optional: two zero-initialised int32 in device memory owned by this call site.  Given: the whole block is
                             * ONE launch (a persistent grid with a device-wide barrier between the pool and the mix); NULL: two launches.
                             * pir_prompt_gen_kernels() tells which.                                                                  */
} PirPrompt;
int64_t pir_prompt_ws_floats(int32_t B, int32_t HW, int32_t C);
int pir_prompt_gen(const PirPrompt* d, void* stream);
int pir_prompt_gen_kernels(const PirPrompt* d);   /* launches pir_prompt_gen enqueues for this descriptor: 1 or 2 */

/* ---- OverlapPatchEmbed (model.py:206): dense 3x3, fp32 NCHW image -> 16-bit NHWC ---------------------
 * w: fp32 [Cout][Cin][3][3] (the parameter itself), bias fp32 [Cout] or NULL.  Cout % 8 == 0.            */
typedef struct PirPatchEmbed {
  int32_t dtype;
  int32_t B, H, W, Cin, Cout;
  const float* img; const float* w; const float* bias;
  void* out; int64_t out_pitch, out_bstride;
} PirPatchEmbed;
int pir_patch_embed(const PirPatchEmbed* d, void* stream);

/* ---- tile blend for tiled inference (demo.py:36-47) ----------------------------------------------------
 * tiles: fp32 [ny*nx][C][th][tw] restored tiles in the reference's loop order (rows outer, columns inner);
 * ys[ny], xs[nx]: DEVICE int32 arrays of tile origins (demo.py:32-34).  out[c,y,x] = clamp(sum of the
 * tiles covering (y,x) / their count, 0, 1), fp32 [C][H][W].  Deterministic (gather, no atomics).          */
int pir_tile_blend(const float* tiles, int32_t ny, int32_t nx, const int32_t* ys, const int32_t* xs, int32_t C,
                   int32_t th, int32_t tw, float* out, int32_t H, int32_t W, void* stream);

/* ==================================================================================================================
 * Training: forward extras and the backward kernels (reference: autograd of net/model.py under train.py:37-46).
 * Gradients of activations are NHWC 16-bit like the activations; parameter gradients are fp32 in the parameter's own
 * layout.  `inv_scale` undoes a static loss scale the caller applied to dL/d(out).
 * ================================================================================================================== */

/* ---- LayerNorm over channels, split from the GEMM for training (model.py:39-41, 60-63) ---------------------------
 * pir_ln_fwd: xhat = (x - mu) * rstd (WithBias) or x * rstd (BiasFree), rounded to 16 bits; rstd[b*H*W + p] fp32.
 *             (gamma/beta live in the following 1x1 conv's packed weights.)
 * pir_ln_bwd: g += rstd * (d - mean_c(d) - xhat * mean_c(d * xhat))                    (WithBias)
 *             g += rstd * (d - (xhat - mean_c(xhat)) * mean_c(d * xhat))               (BiasFree)
 *             where d = dL/d(xhat) is passed in `x`.                                                            */
typedef struct PirLn {
  int32_t dtype, ln_mode;
  int32_t B, H, W, C;
  const void* x; int64_t x_pitch, x_bstride;
  void* xhat; int64_t xh_pitch, xh_bstride;
  float* rstd;
  void* g; int64_t g_pitch, g_bstride;
} PirLn;
int pir_ln_fwd(const PirLn* d, void* stream);
int pir_ln_bwd(const PirLn* d, void* stream);

/* ---- weight gradient of a 1x1 / 3x3 convolution: split-K tcgen05 Gram over the pixels ----------------------------
 * ws[(img*splits + s)][tap][m][n] = sum over the pixels p of split s of  a[p, m] * b[p + off(tap), n]
 * (tap = ky*3 + kx, off = (ky-1, kx-1), zero outside the image; taps = 1 or 9).  per_image = 0: the splits cover the
 * whole batch (P = splits partials); per_image = 1: every image keeps its own P = B*splits partials (MDTA backward).
 * colsum (optional) receives the matching partial column sums of a: colsum[part][m].  fp32, deterministic.          */
typedef struct PirWgrad {
  int32_t dtype;
  int32_t B, H, W;
  int32_t M, N, taps;
  int32_t per_image, splits;
  const void* a; int64_t a_pitch, a_bstride;
  const void* b; int64_t b_pitch, b_bstride;
  float* ws;
  float* colsum;
} PirWgrad;
int pir_wgrad_splits(int32_t B, int32_t HW, int32_t M, int32_t N, int32_t taps, int32_t per_image);
int pir_wgrad(const PirWgrad* d, void* stream);

/* Reduce P partials into the parameter gradient dst_w[R][Cc][taps] (a conv weight [R, Cc, 3, 3] or [R, Cc, 1, 1]).
 * Parameter row r reads partial row  r < half ? r : r - half + half_pad  (GDFN's padded [x1 | x2] row space).
 * With gamma (taps == 1; the conv input was xhat, the packed weight W*gamma, the additive vector W.beta + bias):
 *   G = sum_P ws,  s = sum_P colsum:   dW[r][k] = gamma[k] G[r][k] + beta[k] s[r],
 *   dgamma[k] = sum_r W[r][k] G[r][k],  dbeta[k] = sum_r W[r][k] s[r],  dbias[r] = s[r].     (model.py:60-63 backward) */
typedef struct PirWgradFin {
  int32_t P, M, N, taps;
  int32_t R, Cc, half, half_pad;
  float* ws; float* colsum;                /* consumed: the LayerNorm path reduces them in place into partial 0 */
  float inv_scale;
  const float* gamma; const float* beta; const float* w;
  float* dst_w; float* dst_gamma; float* dst_beta; float* dst_bias;
} PirWgradFin;
int pir_wgrad_finalize(const PirWgradFin* d, void* stream);

/* ---- weight gradient of a depthwise 3x3: dw[c][tap] = sum_p dy[p, c] * x[p + off(tap), c]  (model.py:90,112) -------
 * ws: fp32 [parts][10][C] scratch (9 taps + bias); dst_w [R][9], dst_bias [R] or NULL; row map as above.            */
typedef struct PirDwWgrad {
  int32_t dtype;
  int32_t B, H, W, C;
  int32_t parts, R, half, half_pad;
  const void* x; int64_t x_pitch, x_bstride;
  const void* dy; int64_t dy_pitch, dy_bstride;
  float* ws; float inv_scale;
  float* dst_w; float* dst_bias;
} PirDwWgrad;
int pir_dw_wgrad_parts(int32_t B, int32_t H, int32_t W, int32_t C);
int pir_dw_wgrad(const PirDwWgrad* d, void* stream);

/* ---- GDFN gate backward (model.py:96-97): y = [y1 | y2] (2C channels) is overwritten by its gradient --------------
 * dy1 = dg * y2 * (Phi(y1) + y1 phi(y1)),  dy2 = dg * y1 * Phi(y1)      (exact erf GELU)                            */
typedef struct PirGateBwd {
  int32_t dtype;
  int32_t B, H, W, C;
  void* y; int64_t y_pitch, y_bstride;
  const void* dg; int64_t dg_pitch, dg_bstride;
} PirGateBwd;
int pir_gate_bwd(const PirGateBwd* d, void* stream);

/* ---- MDTA backward on the c x c matrices (model.py:123-137) --------------------------------------------------------
 * Inputs: the forward workspace of pir_mdta_gram/finalize (Gram partials, norms, attention), and the per-image partial
 * weight gradient dWfold[b] = sum_p g[p,:]^T v[p,:] produced by pir_wgrad(per_image = 1) in ws_b (+ colsum_b).
 * Outputs: dst_wo (+=0, assigned) [C][C], dst_temp [heads], dst_bias [C] or NULL, and two per-image 16-bit weight sets
 *   wft[b][j][o] = Wfold[b][o][j]                   -> dv     = pir_gemm(g, wft, batched)
 *   wqk[b][2C][Kpad(2C)]                            -> d[q|k] = pir_gemm([q|k], wqk, batched)
 * where wqk carries softmax, temperature, cosine and L2-normalisation backward.  scratch: pir_mdta_bwd_ws_floats().   */
typedef struct PirMdtaBwd {
  int32_t dtype;
  int32_t B, C, heads;
  int32_t splits_f, splits_b;
  const float* ws_f; const float* ws_b; const float* colsum_b;
  const float* temperature; const float* wo;
  float inv_scale;
  float* scratch;
  void* wft; void* wqk;
  float* dst_wo; float* dst_temp; float* dst_bias;
} PirMdtaBwd;
int64_t pir_mdta_bwd_ws_floats(int32_t B, int32_t C, int32_t heads);
int pir_mdta_bwd(const PirMdtaBwd* d, void* stream);

/* ---- PixelShuffle(2) / PixelUnshuffle(2) as a permutation (gradients of model.py:165,175) --------------------------
 * up = 1: in [B,H,W,4c] -> out [B,2H,2W,c];  up = 0: in [B,2H,2W,c] -> out [B,H,W,4c].  H, W, C describe the 4c side. */
typedef struct PirShuffle {
  int32_t dtype, up;
  int32_t B, H, W, C;
  const void* in; int64_t in_pitch, in_bstride;
  void* out; int64_t out_pitch, out_bstride;
} PirShuffle;
int pir_pixel_shuffle(const PirShuffle* d, void* stream);

/* ---- PromptGenBlock backward (model.py:226-232): bilinear^T, component mix, softmax, linear, mean pool --------------
 * dup: gradient of the resized prompt [B,H,W,D]; pool_ws / weights: what pir_prompt_gen left in ws / weights_out.
 * demb[b][c] = dL/d(mean-pooled feature) / (H*W): add it to every pixel of the feature gradient with pir_bcast_add.     */
typedef struct PirPromptBwd {
  int32_t dtype;
  int32_t B, H, W, C, L, D, S;
  const void* dup; int64_t dup_pitch, dup_bstride;
  const float* prompt; const float* weights; const float* pool_ws; const float* lin_w;
  float inv_scale;
  float* scratch;
  float* demb;
  float* dst_prompt; float* dst_lin_w; float* dst_lin_b;
  int32_t align_corners;    /* bilinear rule of the forward (see PirPrompt)                                                  */
} PirPromptBwd;
int64_t pir_prompt_bwd_ws_floats(int32_t B, int32_t L, int32_t D, int32_t S);
int pir_prompt_bwd(const PirPromptBwd* d, void* stream);

/* g[b,p,c] += v[b][c] */
typedef struct PirBcastAdd {
  int32_t dtype;
  int32_t B, H, W, C;
  void* g; int64_t g_pitch, g_bstride;
  const float* v;
} PirBcastAdd;
int pir_bcast_add(const PirBcastAdd* d, void* stream);

/* fp32 NCHW [B,C,H,W] -> 16-bit NHWC [B,H,W,Cpad] (* scale), channels C..Cpad-1 zero */
typedef struct PirToNhwc16 {
  int32_t dtype;
  int32_t B, C, H, W, Cpad;
  const float* src;
  void* out; int64_t out_pitch, out_bstride;
  float scale;
} PirToNhwc16;
int pir_nchw32_to_nhwc16(const PirToNhwc16* d, void* stream);

/* ---- OCAB: overlapping cross-attention of PromptXRestormer (net/prompt_xrestormer.py:189-235, RelPosEmb :25-73) -----------
 * qkv: NHWC 16-bit [B,H,W,3*inner] from the 1x1 qkv conv (inner = heads*dim_head; q = [0,inner), k = [inner,2 inner), v = rest;
 * head h = channels [h*dim_head, (h+1)*dim_head)).  For every ws x ws query window and head:
 *   out = softmax( qs.k^T + qs.rel_w[kc - y + ows - 1] + qs.rel_h[kr - x + ows - 1] ) . v,    qs = q * dim_head^-0.5
 * over the ows x ows key window centred on it; keys outside the image are ZERO vectors that still take part in the softmax
 * (nn.Unfold padding).  rel_h / rel_w: fp32 [2*ows - 1][dim_head].  out: NHWC 16-bit [B,H,W,inner] (project_out follows as a
 * pir_gemm).  Built for ws = 8, ows = 12, dim_head = 16 (the only values the reference constructs).                           */
typedef struct PirOcab {
  int32_t dtype;
  int32_t B, H, W;
  int32_t heads, dim_head, ws, ows;
  const void* qkv; int64_t qkv_pitch, qkv_bstride;
  const float* rel_h; const float* rel_w;
  void* out; int64_t out_pitch, out_bstride;
} PirOcab;
int pir_ocab(const PirOcab* d, void* stream);

/* ---- evaluation I/O on the device (test.py:100-114, utils/val_utils.py:50-66, utils/dataset_utils.py:195-198) ----------------
 * pir_mirror_pad: out[p][y][x] = in[p][y < H ? y : 2H-1-y][x < W ? x : 2W-1-x] for `planes` = B*C fp32 planes (the reference's
 *   torch.cat([x, flip(x)])[:Hp] padding); H <= Hp <= 2H.
 * pir_psnr_ssim: per image, on clip(., 0, 1) fp32 NCHW tensors: out[2b] = skimage peak_signal_noise_ratio(data_range=1),
 *   out[2b+1] = skimage structural_similarity(win_size=7 uniform, sample covariance, K1=.01, K2=.03, data_range=1, channel mean).
 *   ws: pir_psnr_ssim_ws_bytes() bytes of scratch.  H, W >= 7.
 * pir_add_noise: out = floor(clip(clean255 + sigma * N(0,1), 0, 255)) / 255 (clean255 holds 0..255 values); Philox stream (seed, index). */
int pir_mirror_pad(const float* in, float* out, int32_t planes, int32_t H, int32_t W, int32_t Hp, int32_t Wp, void* stream);
int64_t pir_psnr_ssim_ws_bytes(int32_t B, int32_t C, int32_t H, int32_t W);
int pir_psnr_ssim(const float* restored, const float* clean, int32_t B, int32_t C, int32_t H, int32_t W, void* ws, double* out, void* stream);
int pir_add_noise(const float* clean255, float* out, int64_t n, float sigma, uint64_t seed, void* stream);

/* ---- OCAB backward (autograd of prompt_xrestormer.py:215-232 + RelPosEmb :25-73) ----------------------------------------------
 * Given qkv (as for pir_ocab) and dout = dL/d(out) [B,H,W,inner], recomputes the attention of every (window, head) and writes
 *   dqkv [B,H,W,3*inner] 16-bit: dq directly (query windows do not overlap); dk, dv of a pixel are the sum over the (up to four)
 *        overlapping key windows that contain it, gathered deterministically from per-window partials in `ws`,
 *   dst_rel_h, dst_rel_w: fp32 [2*ows-1][dim_head] parameter gradients (* inv_scale), reduced over all windows / images / heads.
 * ws: pir_ocab_bwd_ws_floats() floats of scratch.                                                                              */
typedef struct PirOcabBwd {
  int32_t dtype;
  int32_t B, H, W;
  int32_t heads, dim_head, ws_, ows;
  const void* qkv; int64_t qkv_pitch, qkv_bstride;
  const void* dout; int64_t dout_pitch, dout_bstride;
  const float* rel_h; const float* rel_w;
  void* dqkv; int64_t dqkv_pitch, dqkv_bstride;
  float* ws; float inv_scale;
  float* dst_rel_h; float* dst_rel_w;
} PirOcabBwd;
int64_t pir_ocab_bwd_ws_floats(int32_t B, int32_t H, int32_t W, int32_t heads);
int pir_ocab_bwd(const PirOcabBwd* d, void* stream);

/* ---- derived weight caches rebuilt on the device (host spec: promptir_b200/packing.py) -----------------------------------------
 * The reference keeps fp32 nn.Parameters and lets cuDNN/cuBLAS read them directly (model.py:88-92,111-113,164,174,223); the sm_100a
 * kernels read 16-bit K-major copies (LayerNorm gamma/beta of model.py:60-63 folded in, zero padding, GDFN [x1 | x2] padded channel
 * space, tap-major 3x3 layouts, transposed / tap-flipped variants for the backward).  pir_repack rebuilds ALL of them from the live
 * parameters in one launch: after optimizer.step() (train.py:52-56), after load_state_dict, at engine construction.
 *   jobs_dev      n_jobs records in DEVICE memory (every pointer inside is a device pointer)
 *   first_row_dev n_jobs + 1 int32 in device memory: exclusive prefix sum of pir_repack_rows(job); n_rows = first_row[n_jobs]
 * inverse channel map of a job axis (split, hp): split == 0 -> identity (indices >= source length are zero padding); else indices
 * [0, split) map to themselves, [split, hp) are padding and hp + j maps to source split + j.                                    */
enum { PIR_PACK_POINTWISE = 0, PIR_PACK_CONV3X3 = 1, PIR_PACK_DEPTHWISE = 2, PIR_PACK_VEC = 3, PIR_PACK_PROMPT = 4 };
typedef struct PirPackJob {
  int32_t kind;
  int32_t dst_dtype;               /* 16-bit kinds: PIR_DTYPE_*; VEC / PROMPT write fp32                                          */
  int32_t n, k;                    /* logical source rows / columns (POINTWISE: N x K; CONV3X3: N x Cin; DEPTHWISE / VEC: n = length; PROMPT: L x D) */
  int32_t transpose;               /* POINTWISE: source stored [k][n]; CONV3X3: source stored [Cin][N][3][3]                      */
  int32_t flip;                    /* CONV3X3 / DEPTHWISE: taps reversed (tap -> 8 - tap), i.e. the kernel rotated by 180 degrees  */
  int32_t n_total, k_pad;          /* destination rows / row length (PROMPT: n_total = S*S; VEC / DEPTHWISE: n_total = length)    */
  int32_t row_split, row_hp;       /* inverse map of the destination row (DEPTHWISE / VEC: channel) axis                          */
  int32_t col_split, col_hp;       /* inverse map of the destination column axis (POINTWISE)                                      */
  int32_t gamma_axis;              /* POINTWISE: 0 none, 1 gamma[source column], 2 gamma[source row]                              */
  int32_t reserved;
  const float* src; const float* gamma; const float* beta; const float* bias;
  void* dst; float* ln_s; float* vec_t;
} PirPackJob;
int64_t pir_repack_rows(const PirPackJob* job);
int pir_repack(const PirPackJob* jobs_dev, const int32_t* first_row_dev, int32_t n_jobs, int32_t n_rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROMPTIR_B200_H_ */
