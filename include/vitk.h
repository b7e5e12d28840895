/*
 * vitk.h -- C-ABI of libvitk.so: the B200 (sm_100a) kernels behind the ViT/DeiT training path.
 *
 * The reference (gogolB/thyroid-vit-cnn-comparison) is pure Python and has no FFI of its own
 * (SURVEY.md section 8b): its "boundary" is the set of eager ATen call sites inside
 * src/models/vit/vision_transformer_base.py, src/models/vit/deit_models.py and
 * src/training/lightning_modules.py.  Each entry point below replaces one (or one fused group)
 * of those call sites; the file:line it replaces is cited next to it.  INTEGRATION.md shows the
 * ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer into caller-owned memory
 *     (torch storage in practice) unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs.
 *   - return value: 0 on success, negative vitk_status on failure; vitk_last_error() gives text.
 *   - bf16 buffers are passed as void* (uint16 storage, torch.bfloat16).
 *   - the library never falls back to a CPU or library path: a missing GPU is an error.
 *
 * Numerics contract
 *   - tensor-core operands are 16-bit (fp16 by default in the engine, bf16 supported everywhere), accumulation,
 *     softmax/LayerNorm statistics, the residual stream and all PARAMETER gradients are fp32.
 *   - activation gradients (dY tensors) are stored in 16 bits multiplied by a loss scale S that lives on the
 *     device (amp_state[0]); kernels that emit parameter gradients multiply by 1/S (amp_state[1], passed as
 *     `grad_unscale`) while accumulating, so the fp32 gradient buffer always holds TRUE gradients.
 *   - amp_state fp32[8] = {S, 1/S, good_steps, skipped_steps, last_step_overflowed, -, -, -}.
 */
#ifndef VITK_H_
#define VITK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITK_ABI_VERSION 25

typedef enum {
  VITK_OK = 0,
  VITK_ERR_INVALID = -1,   /* bad shape / alignment / null pointer            */
  VITK_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed             */
  VITK_ERR_UNSUPPORTED = -3 /* shape outside what the sm_100a kernels cover    */
} vitk_status;

int vitk_abi_version(void);
const char* vitk_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py "gpu_launches"). */
int64_t vitk_launch_count(void);
void vitk_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * GEMM on tcgen05 / TMEM fed by TMA.   D[M,N] = A[M,K] * B[N,K]^T  (+ fused epilogue)
 * replaces nn.Linear / nn.Conv2d(k=16,s=16) forward, dgrad and wgrad:
 *   vision_transformer_base.py:95-101,136-138 (patch proj), :166,178 (qkv), :168,192 (proj),
 *   :212-222 (fc1/fc2) and their autograd backward.
 * Operand storage: a "K-major" operand is row-major [rows, K] (K contiguous, `ld` = row pitch
 * in elements); an "MN-major" operand is row-major [K, rows] (rows contiguous).  The second
 * form lets dgrad read W[N,K] and wgrad read dY[M,N] / X[M,K] without any transpose copy.
 * ------------------------------------------------------------------------------------------ */
typedef enum { VITK_BF16 = 0, VITK_FP32 = 1, VITK_FP16 = 2 } vitk_dtype;

typedef enum {
  VITK_EPI_STORE = 0,      /* out = acc*alpha + bias + residual                               */
  VITK_EPI_GELU = 1,       /* pre = acc*alpha + bias: out = gelu_erf'(pre), out2 = gelu_erf(pre) (16-bit): the    */
                           /* derivative is all backward needs, and one erf evaluation yields both               */
  VITK_EPI_DGELU = 2,      /* out = acc * aux  (aux = the derivative saved by VITK_EPI_GELU, 16-bit)              */
  VITK_EPI_ATOMIC_ADD = 3, /* out(fp32) += acc*alpha  (split-K wgrad, red.global.add.f32)     */
  VITK_EPI_TOKENS = 4      /* patch rows -> token rows: out[b, prefix+p, :] = acc+bias+pos    */
} vitk_epilogue;

/* nn.Dropout with drop_rate > 0 -- pos_drop (vision_transformer_base.py:371,452), Attention.proj_drop (:169,193),
 * Mlp.drop after the activation and after fc2 (:213,219-222).  Masks are never stored: the keep/drop decision of element
 * (row, col) of one Dropout call ("site") is a pure function of (*seed, site, row * cols + col) (Philox4x32-7, 16-bit
 * uniforms, dropped when uniform < round(p * 65536), kept values scaled by 1/(1-p)), so the forward epilogue and the
 * backward kernel that needs the same mask both recompute it.  `seed` is a DEVICE scalar the caller redraws every
 * training step (a captured CUDA graph therefore draws fresh masks on every replay).  cols must be a multiple of 8. */
typedef struct {
  const uint64_t* seed; /* device pointer; NULL (or p == 0) disables the dropout */
  float p;              /* drop probability, 0 <= p < 1 */
  int32_t site;         /* any id that is unique per Dropout call within one step */
} vitk_dropout;
/* factors fp32 [rows, cols] = 0 or 1/(1-p): the mask the fused kernels derive for `drop` (tests / debugging). */
int vitk_dropout_mask(const vitk_dropout* drop, float* factors, int64_t rows, int32_t cols, void* stream);

typedef struct {
  const void* A;      /* bf16 */
  const void* B;      /* bf16 */
  int64_t lda, ldb;   /* row pitch in elements of the row-major storage described above */
  int32_t a_mn_major; /* 0: A stored [M,K]; 1: A stored [K,M] */
  int32_t b_mn_major; /* 0: B stored [N,K]; 1: B stored [K,N] */
  int32_t M, N, K;
  int32_t split_k;    /* >=1 (>1 requires VITK_EPI_ATOMIC_ADD), or 0: chosen by the library */
  int32_t epilogue;   /* vitk_epilogue */
  int32_t out_dtype;  /* vitk_dtype of out (and out2) */
  int32_t a_dtype;    /* VITK_BF16 or VITK_FP16: element type of A (both are 16-bit tensor-core operands and */
  int32_t b_dtype;    /* may be mixed: e.g. dY in bf16 (range) times saved activations in fp16 (precision))   */
  int32_t aux_dtype;  /* element type of aux */
  float alpha;
  const float* alpha_dev; /* optional DEVICE scalar multiplied into alpha (1/loss-scale for wgrad) */
  const float* bias;      /* [N] fp32 or NULL */
  const float* residual;  /* [M, ldr] fp32 or NULL (may alias out) */
  int64_t ldr;
  void* out;
  int64_t ldo;
  void* out2;             /* [M, ldo2], same dtype as out (GELU: the activation) */
  int64_t ldo2;
  const void* aux;        /* 16-bit [M, ldaux] (DGELU: saved gelu'(pre)) */
  int64_t ldaux;
  /* VITK_EPI_TOKENS: input row r = b*rows_per_img + p  ->  output row b*tokens_per_img + prefix + p,
   * pos is fp32 [tokens_per_img, N] (pos_embed), added to the row it lands on. */
  int32_t rows_per_img, tokens_per_img, prefix;
  const float* pos;
  /* VITK_EPI_ATOMIC_ADD only (optional): colsum_out[m] += alpha * sum_k A[m,k], fp32 [M].  In wgrad (A = dY read
   * MN-major) this is the bias gradient of the same nn.Linear; it falls out of one extra N=16 tensor-core MMA per
   * k-step against a tile of ones, so dY is not read a second time. */
  float* colsum_out;
  /* VITK_EPI_STORE with an fp32 residual (optional): out = residual + row_scale[m] * (acc*alpha + bias), fp32 [M] --
   * the per-sample stochastic-depth factor floor(keep + u) / keep of DropPath (vision_transformer_base.py:56-64, :283-284). */
  const float* row_scale;
  /* Optional dropout of the epilogue result (see vitk_dropout above; element index = out_row * N + col):
   *   fp32 VITK_EPI_STORE with residual: out = residual + row_scale * mask * (acc*alpha + bias)   (proj_drop / Mlp.drop #2)
   *   VITK_EPI_GELU: out = mask * gelu'(pre), out2 = mask * gelu(pre)   (Mlp.drop #1; the saved derivative carries the mask
   *                  into the backward VITK_EPI_DGELU for free)
   *   VITK_EPI_TOKENS: out = mask * (acc + bias + pos)   (pos_drop; vitk_prefix_tokens_fwd covers the cls/dist rows) */
  const uint64_t* drop_seed;
  float drop_p;
  int32_t drop_site;
} vitk_gemm_args;

int vitk_gemm(const vitk_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm (eps 1e-5, affine) -- vision_transformer_base.py:263,273,377 (nn.LayerNorm)
 * fwd: x fp32 [rows, dim] -> y (fp16|bf16) [rows, dim], mean/rstd fp32 [rows]
 * bwd: dx = (dres or 0) + LN'(dy) in fp32 (+ optional 16-bit copy dx16 for the next dgrad/wgrad GEMM);
 *      dgamma/dbeta/dcolsum are ACCUMULATED (+=) into fp32 [dim], each multiplied by *grad_unscale when given.
 *      dcolsum (optional) receives the column sum of the emitted dx -- the bias gradient of the nn.Linear
 *      that produced the residual branch feeding this LayerNorm's input.
 *      branch_scale (optional, fp32 [rows]): stochastic-depth factor of that residual branch; dx16 and dcolsum are the
 *      gradient ENTERING the branch, i.e. dx * branch_scale[row] (dx itself, the residual-stream gradient, is unscaled).
 *      branch_drop (optional): the dropout applied to that branch's output in forward; dx16 / dcolsum are additionally
 *      multiplied by its mask.
 * ------------------------------------------------------------------------------------------ */
int vitk_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int32_t y_dtype,
                       float* mean, float* rstd, int64_t rows, int32_t dim, float eps, void* stream);
int vitk_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* mean,
                       const float* rstd, const float* gamma, const float* dres, float* dx, void* dx16,
                       int32_t dx16_dtype, float* dgamma, float* dbeta, float* dcolsum,
                       const float* grad_unscale, const float* branch_scale, const vitk_dropout* branch_drop,
                       int64_t rows, int32_t dim, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused softmax attention, dh = 64 -- vision_transformer_base.py:174-191 (Attention.forward:
 * q@k^T * scale, softmax, attn@v) without materialising [B,H,N,N].  One element type (`dtype`:
 * fp16|bf16) for qkv / out / dout / dqkv.
 * qkv  [B, N, 3, H, 64]  (exactly the layout nn.Linear(D,3D) emits, :178)
 * out  [B, N, H, 64]     (== (attn@v).transpose(1,2).reshape(B,N,C), :191)
 * lse  fp32 [B, H, N]    natural-log sum-exp of the scaled scores (saved for backward)
 * bwd recomputes P from q,k and lse; delta = rowsum(dout*out) is computed internally into
 * `delta` (fp32 [B,H,N] scratch).
 * probs (optional, eval only): fp32 [B,H,N,N] attention maps (:186-188 `attention_maps`).
 * Any sequence length: up to 256 (forward) / 240 (backward) tokens one CTA holds the whole key range, longer sequences
 *   (384x384 images: 577 tokens; patch 8: 785 / 1025) walk key / query tiles; every product runs on tcgen05.
 * q_rows (0 = every row): the caller needs only query rows 0..q_rows-1 of `out` / `lse` (forward; the other rows may be left
 *   unwritten) resp. guarantees that `dout` is zero from row q_rows on (backward; dqkv is still complete: dQ is zero there and
 *   dK / dV sum over the leading rows).  The last block of a class-token model: the classifier reads x[:, 0] (and x[:, 1]) only
 *   (vision_transformer_base.py:474-479, deit_models.py:224-235).  Ignored when `probs` is requested.
 * ------------------------------------------------------------------------------------------ */
int vitk_attention_fwd(const void* qkv, void* out, int32_t dtype, float* lse, float* probs, int32_t B,
                       int32_t N, int32_t H, float scale, int32_t q_rows, void* stream);
/* The maps alone, from the lse a preceding vitk_attention_fwd wrote: probs + b * probs_batch_stride holds image b's
 * [H,N,N] maps (stride in elements; H*N*N = the contiguous [B,H,N,N] layout). */
int vitk_attention_probs(const void* qkv, int32_t dtype, const float* lse, float* probs, int64_t probs_batch_stride,
                         int32_t B, int32_t N, int32_t H, float scale, void* stream);
int vitk_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                       float* delta, void* dqkv, int32_t dtype, int32_t B, int32_t N, int32_t H,
                       float scale, int32_t q_rows, void* stream);
/* The same pair with nn.Dropout on the softmax output (Attention.attn_drop, vision_transformer_base.py:184), training mode
 * only: out = (P o M) V, M[b,h,q,key] = 0 | 1/(1-p) from the counter-based generator at element
 * ((b*H + h)*N + q)*Npad + key, Npad = N rounded up to 8 -- i.e. vitk_dropout_mask(seed, p, site, B*H*N, Npad)[:, :N];
 * lse is that of the unmasked P; the backward re-derives M (nothing is stored).  0 < p < 1 required. */
int vitk_attention_dropout_fwd(const void* qkv, void* out, int32_t dtype, float* lse, int32_t B, int32_t N,
                               int32_t H, float scale, const vitk_dropout* attn_drop, void* stream);
int vitk_attention_dropout_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                               float* delta, void* dqkv, int32_t dtype, int32_t B, int32_t N, int32_t H,
                               float scale, const vitk_dropout* attn_drop, void* stream);

/* ------------------------------------------------------------------------------------------
 * Patch / token plumbing -- vision_transformer_base.py:120-143 (PatchEmbed.forward),
 * deit_models.py:200-211 and vision_transformer_base.py:446-452 (cls/dist tokens + pos_embed).
 * ------------------------------------------------------------------------------------------ */
/* images fp32 NCHW [B,C,H,W] -> fp16|bf16 patch matrix [B*gh*gw, C*P*P] (k = c*P*P + ky*P + kx,
 * the flattening order of Conv2d.weight[D,C,P,P]) */
int vitk_patchify(const float* images, void* patches, int32_t patches_dtype, int32_t B, int32_t C,
                  int32_t H, int32_t W, int32_t P, void* stream);
/* the same gather with channel-last patch vectors, k = (ky*P + kx)*C + c: PatchEmbed(projection_type='linear'),
 * einops 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' followed by nn.Linear (vision_transformer_base.py:102-107, :140) */
int vitk_patchify_hwc(const float* images, void* patches, int32_t patches_dtype, int32_t B, int32_t C,
                      int32_t H, int32_t W, int32_t P, void* stream);
/* x[b, t, :] = drop(tok_t + pos[t, :]) for t < n_prefix (cls, dist); drop = pos_drop (optional) */
int vitk_prefix_tokens_fwd(float* x, const float* cls_tok, const float* dist_tok, const float* pos,
                           int32_t B, int32_t tokens_per_img, int32_t dim, int32_t n_prefix,
                           const vitk_dropout* drop, void* stream);
/* dpos[t,:] += u*sum_b dx[b,t,:] ; dcls += u*sum_b dx[b,0,:] ; ddist += u*sum_b dx[b,1,:]  (u = *grad_unscale or 1);
 * dpatch16 [B*rows_per_img, dim] = 16-bit copy of dx[b, n_prefix+p, :] (the dY of the patch GEMM, still scaled);
 * dbias_patch[dim] += u * column sum over patch rows.  With `drop` (pos_drop) every dx element is first multiplied by the
 * mask the forward applied to that token element. */
int vitk_tokens_bwd(const float* dx, float* dpos, float* dcls, float* ddist, void* dpatch16,
                    int32_t dpatch_dtype, float* dbias_patch, const float* grad_unscale, int32_t B,
                    int32_t tokens_per_img, int32_t dim, int32_t n_prefix, const vitk_dropout* drop, void* stream);

/* Leading rows of every image, compact <-> dense.  The classifier reads x[:, 0] (and x[:, 1] for the distillation head) of the
 * last block's output only (vision_transformer_base.py:474-479, deit_models.py:224-235), so that block's attn.proj, norm2 and Mlp
 * (vision_transformer_base.py:274-285) need the first n tokens of each image and nothing else: the engine runs them on B*n rows.
 * gather: dst[b, j, :] = src[b, j, :] for j < n (src is [B, T, row_bytes], dst is [B, n, row_bytes]);
 * expand: dst[b, j, :] = j < n ? src[b, j, :] : 0 (src [B, n, row_bytes], dst [B, T, row_bytes]) -- the gradient of those rows
 * placed back into the dense tensor the rest of the backward continues from.  row_bytes % 16 == 0, 16-byte aligned bases. */
int vitk_gather_rows(const void* src, void* dst, int32_t B, int32_t tokens_per_img, int32_t n, int64_t row_bytes, void* stream);
int vitk_expand_rows(const void* src, void* dst, int32_t B, int32_t tokens_per_img, int32_t n, int64_t row_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Final norm + classification heads on the pooled rows only --
 * deit_models.py:217,224-235 / vision_transformer_base.py:468-486 (norm, x[:,0], head, head_dist)
 * x fp32 [B, tokens_per_img, dim]; for head h in [0,n_heads): row h of every image is
 * LayerNorm'd and multiplied by W_h [C, dim] (+ b_h [C]) -> logits_h [B, C].
 * bwd takes TRUE dlogits, writes dx / dx16 = S * d(loss)/dx (S = *loss_scale or 1; ZERO outside the pooled
 * rows), accumulates every (true) parameter gradient and (optional) dcolsum[dim] += column sum of the true dx
 * (bias gradient of the last block's fc2).
 * ------------------------------------------------------------------------------------------ */
int vitk_head_fwd(const float* x, const float* gamma, const float* beta, const float* W0,
                  const float* b0, const float* W1, const float* b1, float* logits0,
                  float* logits1, float* xhat, float* rstd, float* pooled, int32_t B,
                  int32_t tokens_per_img, int32_t dim, int32_t C, int32_t n_heads, float eps,
                  void* stream);
int vitk_head_bwd(const float* dlogits0, const float* dlogits1, const float* xhat,
                  const float* rstd, const float* gamma, const float* beta, const float* W0,
                  const float* W1, float* dx, void* dx16, int32_t dx16_dtype, float* dgamma, float* dbeta,
                  float* dW0, float* db0, float* dW1, float* db1, float* dcolsum,
                  const float* loss_scale, const float* branch_scale, const vitk_dropout* branch_drop, int32_t B,
                  int32_t tokens_per_img, int32_t dim, int32_t C, int32_t n_heads, void* stream);

/* General classification tail -- vision_transformer_base.py:468-486 for the constructor options the fused head kernels above
 * do not cover: pool_type 'gap' (mean of norm(x)[:, 1:], or of all tokens without a class token, :471-474), and
 * `pre_logits` = Linear + Tanh when representation_size is set (:380-386, :477).
 * pool_norm_fwd: pooled[b, :] = mean over tokens t in [t0, t1) of LayerNorm(x[b, t, :]) (fp32 [B, dim]); mean / rstd
 *   (fp32 [B, tokens_per_img]) are written for those rows.
 * pool_norm_bwd: takes the TRUE dpooled; writes dx / dx16 = S * d(loss)/dx for every row (zero outside the range) with the
 *   same loss_scale / branch_scale / branch_drop meaning as vitk_head_bwd; accumulates dgamma, dbeta, dcolsum.
 * dense_fwd: y = act(x W^T + bias), x fp32 [B, in_dim], W [out_dim, in_dim]; act 0 identity, 1 tanh.
 * dense_bwd: dz = dy * act'(y) (scratch fp32 [B, out_dim]); dx = dz W (optional); dW += dz^T x; db += column sums (optional). */
int vitk_pool_norm_fwd(const float* x, const float* gamma, const float* beta, float* pooled, float* mean, float* rstd,
                       int32_t B, int32_t tokens_per_img, int32_t dim, int32_t t0, int32_t t1, float eps, void* stream);
int vitk_pool_norm_bwd(const float* dpooled, const float* x, const float* mean, const float* rstd, const float* gamma,
                       float* dx, void* dx16, int32_t dx16_dtype, float* dgamma, float* dbeta, float* dcolsum,
                       const float* loss_scale, const float* branch_scale, const vitk_dropout* branch_drop, int32_t B,
                       int32_t tokens_per_img, int32_t dim, int32_t t0, int32_t t1, void* stream);
int vitk_dense_fwd(const float* x, const float* W, const float* bias, float* y, int32_t B, int32_t in_dim, int32_t out_dim,
                   int32_t act, void* stream);
int vitk_dense_bwd(const float* dy, const float* y, const float* x, const float* W, float* dz, float* dx, float* dW,
                   float* db, int32_t B, int32_t in_dim, int32_t out_dim, int32_t act, void* stream);

/* Frozen-teacher fast path (lightning_modules.py:943-947: `self.teacher(images)`, eval mode, no_grad): eval-mode BatchNorm
 * + ReLU over the first C channels of an NHWC feature buffer whose pixels are x_ld elements apart (the dense block's
 * preallocated concatenation buffer) -> y (pixel pitch y_ld): y[p, c] = max(0, x[p, c] * scale[c] + shift[c]) (relu = 0: affine
 * only), scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale.  Replaces torch.cat + batch_norm + relu of
 * torchvision's _DenseLayer / _Transition (three read+write passes) by one.  C, x_ld, y_ld multiples of 8; 16-byte aligned. */
int vitk_affine_relu_nhwc(const void* x, int64_t x_ld, void* y, int64_t y_ld, const float* scale, const float* shift,
                          int64_t pixels, int32_t C, int32_t dtype, int32_t relu, void* stream);
/* The 1x1 bottleneck convolution of a frozen dense layer with everything around it (torchvision densenet.py _DenseLayer:
 * norm1 -> relu1 -> conv1 -> norm2 -> relu2, eval mode, norm2 folded into conv1):
 *   out[p, 0:128] = relu( relu(x[p, 0:C] * scale + shift) @ w^T + bias )
 * x: 16-bit NHWC concatenation buffer, pixel pitch x_ld elements; w: 16-bit [128, C] (row-major); scale / shift fp32 [C],
 * bias fp32 [128]; out: 16-bit [pixels, 128].  tcgen05 GEMM whose A tile gets the affine + ReLU in shared memory between
 * the TMA load and the MMA: the concatenation is read once per layer (vitk_affine_relu_nhwc + a library convolution read
 * and wrote it three times). */
int vitk_dense_bottleneck(const void* x, int64_t x_ld, const float* scale, const float* shift, const void* w,
                          const float* bias, void* out, int64_t pixels, int32_t C, int32_t dtype, void* stream);
/* Patch rows of a k x k / stride / zero-padded convolution over a 16-bit NHWC tensor with few channels (the teacher's 7x7
 * 3-channel stem, densenet.py features.conv0): patches[(b,oy,ox), (ky*k + kx)*C + c], row pitch ld (multiple of 8, the tail
 * is zero-filled) -- the A operand of vitk_gemm against the filters reshaped to [Cout, ky, kx, c]. */
int vitk_im2col_rows(const void* x, void* patches, int32_t B, int32_t H, int32_t W, int32_t C, int32_t kernel,
                     int32_t stride, int32_t pad, int64_t ld, void* stream);
/* The same stem without the patch matrix: 7x7 / stride 2 / padding 3 convolution of a 3-channel NHWC batch as an implicit GEMM
 * (torchvision densenet.py features.conv0 with norm0 folded into w / bias; relu != 0 applies relu0 -- ReLU and the max pool
 * behind it commute).  x [B,H,W,3], out [B,H/2,W/2,64], one 16-bit type for x / w / out; w [64,192] with
 * w[o][ky*24 + 1 + kx*3 + c] = filter[o][c][ky][kx] and zeros in slots 0, 22, 23 of every ky segment and from column 168 on;
 * bias fp32 [64].  H even, W % 8 == 0. */
int vitk_stem_conv7(const void* x, const void* w, const float* bias, void* out, int32_t B, int32_t H, int32_t W,
                    int32_t relu, int32_t dtype, void* stream);
/* MaxPool2d(kernel, stride, pad) (is_max = 1, -inf padding) / AvgPool2d(kernel, stride, pad) (is_max = 0, count_include_pad) of a
 * compact NHWC tensor x [B,H,W,C], ceil_mode False, written with pixel pitch y_ld into y [B,OH,OW,y_ld]: the DenseNet stem's
 * pool0 and the transitions' pool (torchvision densenet.py) store straight into the next dense block's buffer. */
int vitk_pool_nhwc(const void* x, void* y, int64_t y_ld, int32_t B, int32_t H, int32_t W, int32_t C, int32_t kernel,
                   int32_t stride, int32_t pad, int32_t is_max, int32_t dtype, void* stream);

/* Stochastic depth (DropPath.forward, vision_transformer_base.py:56-64): scale[br, b*T + t] = floor(keep + u[br,b]) / keep
 * with keep = 1 - drop_prob[br]; `uniform` fp32 [branches, B] in [0,1), `scale` fp32 [branches, B*T]. */
int vitk_droppath_scale(const float* uniform, const float* drop_prob, float* scale, int32_t branches, int32_t B,
                        int32_t tokens_per_img, void* stream);

/* ------------------------------------------------------------------------------------------
 * GPU-side input pipeline (the step right before the hot path) for raw single-channel uint16
 * tiles already in device memory:
 *   vitk_resize_u16        dataset.py:533-551 `_preprocess_image`: cv2.resize(INTER_LINEAR) on uint16 (skipped when the size
 *                          already matches), result rounded to the uint16 grid, / 65535 -> gray fp32 [B, H, W]
 *   vitk_percentile_bounds quality_preprocessing.py:306-318 AdaptiveNormalization('percentile'): per image
 *                          (torch.quantile(x, q_lo), torch.quantile(x, q_hi)), default linear interpolation, exact order
 *                          statistics by radix select -> bounds fp32 [B, 2]
 *   vitk_finish_tiles      clamp to the bounds and (x - lo) / (hi - lo + 1e-8) (:314-316, optional); gray -> C channels
 *                          and T.Normalize(mean, std) (vit_transforms.py:381-393; mean/std are HOST arrays [C], NULL = none);
 *                          then MixUp (cutmix = 0: lam*a + (1-lam)*b, b = image perm[b]) or CutMix (cutmix = 1: image perm[b]
 *                          inside the box rows [y1,y2) x columns [x1,x2)) -- vit_transforms.py:396-462; perm NULL = no mix.
 *                          out fp32 [B, C, H, W], W % 4 == 0, out must not alias gray.
 * ------------------------------------------------------------------------------------------ */
int vitk_resize_u16(const uint16_t* raw, float* gray, int32_t B, int32_t Hs, int32_t Ws, int32_t H, int32_t W, void* stream);
int vitk_percentile_bounds(const float* x, int32_t B, int64_t n, float q_lo, float q_hi, float* bounds, void* stream);
int vitk_finish_tiles(const float* gray, const float* bounds, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                      const float* mean, const float* stdv, const int32_t* perm, int32_t cutmix, float lam, int32_t x1,
                      int32_t y1, int32_t x2, int32_t y2, void* stream);
/* Single-channel tiles straight to the 16-bit patch matrix the patch-embedding GEMM reads -- the input side of
 * PatchEmbed.proj (vision_transformer_base.py:95-101,136-138) fused with `x.repeat(3,1,1)` + T.Normalize
 * (vit_transforms.py:381-393) and, optionally, the percentile clamp-normalise (:314-316):
 *   patches[(b,py,px), c*P*P + ky*P + kx] = 16-bit((norm(tile[b, py*P+ky, px*P+kx]) - mean[c]) / std[c])
 * tiles [B,H,W]: tiles_kind 0 = fp32 in [0,1], 1 = raw uint16 (value / 65535, dataset.py:549), 2 = fp16, 3 = bf16.
 * The fp32 value before the 16-bit rounding is bit-identical to vitk_finish_tiles' output, so the result equals
 * vitk_finish_tiles + vitk_patchify; the fp32 [B,C,H,W] batch is never materialised.  mean/std: HOST arrays [C] or NULL. */
int vitk_tiles_to_patches(const void* tiles, int32_t tiles_kind, const float* bounds, const float* mean,
                          const float* stdv, void* patches, int32_t patches_dtype, int32_t B, int32_t C,
                          int32_t H, int32_t W, int32_t P, void* stream);

/* ------------------------------------------------------------------------------------------
 * On-device classification metrics -- the torchmetrics objects of lightning_modules.py:358-374 and their per-step
 * updates (:496-516 validation, :542-560 test, :902-920 distillation module).
 * update: confusion[label * C + argmax(logits)] += 1 (int64 [C*C + 1]; the last slot counts labels outside [0, C));
 *   for C == 2 and scores != NULL the positive-class probability softmax(logits)[:, 1] (:493-495) and the label are
 *   appended at position *count (device scalar, advanced by B; samples past `capacity` are dropped but still counted).
 *   Accuracy / F1 / specificity / sensitivity / PPV / NPV / StatScores are functions of the four binary counters.
 * binary_auroc: AUROC(task='binary', thresholds=None) over the first min(*count, capacity) samples = (pairs with
 *   score_pos > score_neg + ties / 2) / (P * N), counted in integers; out (float64 [4]) = {auroc (0 when a class is
 *   absent, as torchmetrics), P, N, ties}; scratch int64 [4].
 * ------------------------------------------------------------------------------------------ */
int vitk_metrics_update(const float* logits, const int64_t* labels, int32_t B, int32_t C, int64_t* confusion,
                        float* scores, uint8_t* score_labels, int64_t* count, int64_t capacity, void* stream);
int vitk_binary_auroc(const float* scores, const uint8_t* score_labels, const int64_t* count, int64_t capacity,
                      int64_t* scratch, double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused classification / distillation loss with gradient --
 * lightning_modules.py:455-465 (0.5*CE(cls)+0.5*CE(dist) or CE), :959-974 (distillation:
 * (1-alpha)*CE(cls,y) + alpha*KL(log_softmax(d/T) || softmax(t/T))*T^2 'batchmean', or hard CE
 * on the teacher argmax), deit_models.py:461-480 (DistillationLoss), nn.CrossEntropyLoss
 * label_smoothing (lightning_modules.py:345-350).
 * mode 0: w_cls*CE(cls,y) [+ w_dist*CE(dist,y) if dist given]
 * mode 1: w_cls*CE(cls,y) + w_dist*KL_soft(dist, teacher, T)
 * mode 2: w_cls*CE(cls,y) + w_dist*CE(dist, argmax teacher)
 * out_scalars fp32[8]: {loss, cls_loss, dist_loss, n_correct(cls vs y), n_agree(cls vs teacher), B, n_invalid_labels, 0}.
 * A label outside [0, C) never indexes a row (torch's CE device-asserts): it is counted in out[6], the loss and that
 * row's gradient become NaN, so the optimizer skips the step.
 * dcls/ddist: fp32 [B,C] gradient of `loss` (already divided by B*grad_div; grad_div = world size
 * under data parallel so that an all-reduce SUM yields the global-batch mean).
 * ------------------------------------------------------------------------------------------ */
int vitk_loss_fwd_bwd(const float* cls_logits, const float* dist_logits, const float* teacher_logits,
                      const int64_t* labels, float* out_scalars, float* dcls, float* ddist,
                      int32_t B, int32_t C, int32_t mode, float w_cls, float w_dist, float T,
                      float label_smoothing, float grad_div, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer -- torch.optim.AdamW as configured at lightning_modules.py:599-604,1108-1113 plus
 * Lightning's clip_grad_norm_(1.0) (configs/trainer/default.yaml:21,54).
 * The parameters live in ONE flat fp32 buffer cut into chunks; chunk c covers
 * [chunk_off[c], chunk_off[c]+chunk_len[c]) and uses lr*lr_scale[c], wd[c].
 * state fp32[4] on device: {step, lr, grad_sqnorm, clip_coef}.
 * ------------------------------------------------------------------------------------------ */
/* state[2] += ||grads||^2, summed in a FIXED order (bit-identical from launch to launch and across data-parallel ranks that
 * hold identical gradients).  scratch: vitk_sqnorm_scratch_floats() floats, zero-initialised once by the caller. */
int vitk_sqnorm_scratch_floats(void);
int vitk_grad_sqnorm(const float* grads, int64_t n, float* state, float* scratch, void* stream);
/* If state[2] (the squared norm) is not finite the step is SKIPPED (fp16 overflow): nothing is written, and --
 * when amp_state is given -- S is halved; otherwise S doubles after `growth_interval` clean steps. */
int vitk_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                    void* params_bf16, void* params_fp16, const int64_t* chunk_off,
                    const int32_t* chunk_len, const float* chunk_lr_scale, const float* chunk_wd,
                    int32_t n_chunks, float* state, float* amp_state, float beta1, float beta2, float eps,
                    float max_grad_norm, int32_t growth_interval, void* stream);
/* Loss-scale bookkeeping for callers that run their OWN optimizer on the fp32 gradients (torch.optim.*):
 * checks grads for inf/nan; on overflow zeroes them and halves S, else counts towards the next doubling. */
int vitk_amp_update(float* grads, int64_t n, float* amp_state, float* scratch4, float* sqnorm_scratch,
                    int32_t growth_interval, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small memory-bound helpers
 * ------------------------------------------------------------------------------------------ */
/* fp32 -> 16-bit shadow(s): dst_bf16 and/or dst_fp16 (either may be NULL) */
int vitk_cast_f32_to_16(const float* src, void* dst_bf16, void* dst_fp16, int64_t n, void* stream);
/* out[dim] += u * column sums of a 16-bit [rows, dim] matrix (bias gradients of qkv / fc1), u = *grad_unscale or 1 */
int vitk_colsum16(const void* x, int32_t dtype, float* out, const float* grad_unscale, int64_t rows,
                  int32_t dim, void* stream);

/* ------------------------------------------------------------------------------------------
 * Ensemble + attention rollout (config 5) --
 * scripts/run_ensemble_kfold_evaluation.py:127-152 (sum_m w_m*softmax(logits_m) -> argmax),
 * src/models/vit/attention_utils.py:129-145 (rollout; the reference body is `pass`: spec is
 * Abnar & Zuidema 2020, see oracle/vit_oracle.py::attention_rollout).
 * ------------------------------------------------------------------------------------------ */
int vitk_ensemble_probs(const float* logits /*[F,B,C]*/, const float* weights /*[F]*/,
                        float* probs /*[B,C]*/, int64_t* pred /*[B]*/, int32_t F, int32_t B,
                        int32_t C, void* stream);
/* probs fp32 [L,B,H,N,N] -> rollout fp32 [B,N,N]; fusion 0 mean, 1 max, 2 min */
int vitk_attention_rollout(const float* probs, float* rollout, float* scratch, int32_t L,
                           int32_t B, int32_t H, int32_t N, int32_t fusion, void* stream);
/* Row `row` of the same rollout matrix (row 0 = the class token's map, the only row config 5 / the visualisation code
 * uses, attention_utils.py:49-62): L vector-matrix products per image, the maps are read once, no [B,N,N] scratch.
 * The [H,N,N] maps of (layer l, image b) start at probs + l * layer_stride + b * batch_stride (elements): [L,B,H,N,N] is
 * (B*H*N*N, H*N*N); the image-major layout [B,L,H,N,N] = (H*N*N, L*H*N*N) keeps one image's maps contiguous, which is what
 * the one-CTA-per-image walk wants (measured: 1.31 ms -> see profiles/ for DeiT-tiny, batch 256).  out fp32 [B,N]. */
int vitk_attention_rollout_row(const float* probs, float* out, int64_t layer_stride, int64_t batch_stride, int32_t L,
                               int32_t B, int32_t H, int32_t N, int32_t row, int32_t fusion, void* stream);

/* Class-token heat map of the visualisation code (src/models/vit/attention_utils.py:50-67: `attn.mean(dim=0)[0, 1:]` ->
 * sqrt grid -> F.interpolate(size=image, mode='bilinear', align_corners=False)), for every image of the batch.
 * Element j of head h of image b is src[b * batch_stride + h * head_stride + j]: one layer's maps [B,H,N,N] are
 * (H*N*N, N*N) -- row 0 of each head --, a rollout row [B,N] or a grid [B,g*g] is H = 1 with batch_stride = N / g*g.
 * The g*g patch columns start at n_prefix (1 class token; 2 for a distilled DeiT).  out fp32 [B,out_h,out_w]. */
int vitk_cls_attention_heatmap(const float* src, float* out, int64_t batch_stride, int64_t head_stride,
                               int32_t B, int32_t H, int32_t n_prefix, int32_t grid, int32_t out_h,
                               int32_t out_w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITK_H_ */
