/*
 * b200dino.h -- C-ABI of libb200dino.so: the B200 (sm_100a) kernels behind the DINOv2 training step of
 * lightly-train.  Plain pointers and sizes only; no torch types.  All pointers are DEVICE pointers unless
 * stated otherwise; `stream` is a cudaStream_t passed as void*.  Every entry point enqueues work on `stream`
 * and returns immediately (no device synchronisation), returning 0 on success or a negative error code
 * (B200_ERR_*) -- it never throws and never exits.  Memory is owned by the caller for the duration of the
 * enqueued work.
 *
 * Each function names the reference code (paths relative to the lightly-train source tree,
 * LT = src/lightly_train) whose arithmetic it replaces.
 */
#ifndef B200DINO_H_
#define B200DINO_H_

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_INVALID_ARG (-1)
#define B200_ERR_UNSUPPORTED (-2)
#define B200_ERR_CUDA (-3) /* -3 - 16*cudaError_t */
#define B200_ERR_DRIVER (-4)

/* Library / build information. Returns a static string "b200dino <version> sm_100a". */
const char* b200_version(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM on tcgen05 tensor cores: C[M,N] = epilogue(alpha * A[M,K] * B[N,K]^T).
 * Replaces torch.nn.functional.linear / conv2d-as-GEMM under bf16 autocast and their backward:
 *   LT/_models/dinov2_vit/dinov2_vit_src/layers/attention.py:44-63 (qkv, proj)
 *   LT/_models/dinov2_vit/dinov2_vit_src/layers/mlp.py:31-42       (fc1 + GELU, fc2)
 *   LT/_models/dinov2_vit/dinov2_vit_src/layers/patch_embed.py:77-110 (proj conv, k=s=patch)
 *   LT/_models/dinov2_vit/dinov2_vit_src/layers/layer_scale.py:27-28 + layers/block.py:109-114
 *   LT/_methods/dinov2/dinov2_head.py:66-95 (projection-head linears)
 * A and B are bf16.  x_mn = 0: operand stored row-major [rows, K] (K contiguous, "K-major");
 * x_mn = 1: operand stored [K, rows] (rows contiguous, "MN-major") -- used by wgrad, where the
 * contraction runs over tokens.  ld* are leading dimensions in ELEMENTS.
 */
enum {
  B200_EPI_BF16 = 0,       /* C(bf16) = alpha*acc + bias                                           */
  B200_EPI_F32 = 1,        /* C(f32)  = alpha*acc + bias                                           */
  B200_EPI_F32_ATOMIC = 2, /* C(f32) += alpha*acc   (split-K allowed; wgrad accumulation)          */
  B200_EPI_BIAS_GELU = 3,  /* u = bf16(acc+bias); C2(bf16)=u (optional); C(bf16) = gelu_erf(u)      */
  B200_EPI_RESIDUAL = 4,   /* o = bf16(acc+bias); C2(bf16)=o (optional);
                              C(f32) = aux(f32) + gamma[n]*o*rowscale[m / rows_per_scale]          */
  B200_EPI_DGELU = 5,      /* C(bf16) = bf16(acc) * gelu_erf'(aux(bf16))                           */
  B200_EPI_BIAS_GELU_DG = 6, /* u = bf16(acc+bias); C(bf16) = gelu_erf(u); C2(bf16) = gelu_erf'(u)  */
  B200_EPI_MUL_AUX = 7     /* C(bf16) = bf16(acc) * aux(bf16)   (dGELU with the saved derivative)  */
};

typedef struct b200_gemm_args {
  const void* A; long long lda; int a_mn;
  const void* B; long long ldb; int b_mn;
  int M, N, K;
  int splits;            /* split-K factor (>1 only with B200_EPI_F32_ATOMIC; 0 there = choose automatically) */
  int epi;               /* B200_EPI_*                                                              */
  int block_n;           /* 0 = auto, or 128 / 192 / 256                                            */
  int ws_mode;           /* schedule: 0 = auto, 1 = force weight-stationary (K <= 384), 2 = force the generic  */
                         /* single-CTA ring, 3 = force the CTA-pair (cta_group::2, 256-row tile) kernel      */
  float alpha;
  void* C; long long ldc;
  void* C2; long long ldc2;
  const void* aux; long long ldaux;
  const float* bias;     /* [N] or NULL                                                             */
  const float* gamma;    /* [N] LayerScale or NULL                                                  */
  const float* rowscale; /* [ceil(M/rows_per_scale)] per-sample DropPath scale or NULL              */
  int rows_per_scale;
} b200_gemm_args;

int b200_gemm(const b200_gemm_args* args, void* stream);

/* LayerNorm as the A-operand prologue of a GEMM (north_star: "LayerNorm fused into the adjacent GEMM"):
 *   C = epilogue( LayerNorm(x) @ B^T )   replaces  norm1 -> attn.qkv  and  norm2 -> mlp.fc1  of Block.forward,
 *   LT/_models/dinov2_vit/dinov2_vit_src/layers/block.py:60-78,90-115 (nn.LayerNorm + nn.Linear under bf16 autocast).
 * x: f32 [M, K] row pitch ldx (the residual stream); weight / bias: f32 [K]; the normalised rows are rounded to bf16 (what
 * autocast feeds the Linear) and written straight into the swizzled shared-memory tiles the MMA reads.  Optional outputs:
 * xn_out bf16 [M, K] (the wgrad GEMM of the backward reads it), mean / rstd f32 [M] (LayerNorm backward).
 * args->A is ignored; K % 64 == 0 and K <= 384 (ViT-T / ViT-S); epi in {B200_EPI_BF16, B200_EPI_BIAS_GELU,
 * B200_EPI_BIAS_GELU_DG}; b_mn = 0; block_n 0 (auto) | 128 | 192. */
typedef struct b200_ln_args {
  const float* x; long long ldx;
  const float* weight; const float* bias; float eps;
  void* xn_out; long long ld_xn;
  float* mean; float* rstd;
} b200_ln_args;
int b200_ln_gemm(const b200_gemm_args* args, const b200_ln_args* ln, void* stream);

/* Attention on the 5th-generation tensor cores (tcgen05.mma, fp32 accumulators in TMEM, TMA-staged operands):
 * same contracts as b200_attention_fwd / b200_attention_bwd below, for the global-crop sequence lengths
 * (forward N <= 256, backward N <= 208; lightly_train_b200/csrc/attention_tc.cu).  Replaces
 * LT/_models/dinov2_vit/dinov2_vit_src/layers/attention.py:49-66 and its autograd backward.  P is rounded to bf16
 * before the 1/l normalisation (flash-style); the other rounding points are those of the warp-level kernels. */
int b200_attention_fwd_tc(const void* qkv, long long ld_tok, int B, int N, int h, int head_dim, float scale, void* out,
                          long long ld_out, float* lse, void* stream);
int b200_attention_bwd_tc(const void* qkv, long long ld_tok, const void* out, const void* dout, long long ld_out,
                          const float* lse, int B, int N, int h, int head_dim, float scale, void* dqkv,
                          long long ld_dtok, float* dqkv_colsum, void* stream);
/* ------------------------------------------------------------------------------------------------
 * Distillation path (BASELINE cfg4): DINOv3 teacher RoPE and the DistillationV3 KL terms.
 *
 * b200_rope_apply: axial RoPE on q and k of the patch tokens (tokens >= prefix of every image), in place in the bf16
 *   qkv buffer [B*N, 3*h*64]; sin/cos: f32 [N - prefix, 64].  out = x*cos + rotate_half(x)*sin in fp32, cast to bf16.
 *   Replaces SelfAttention.apply_rope, LT/_models/dinov3/dinov3_src/layers/attention.py:21-33,79-100.
 * b200_kl_rows: rows of KLDivLoss(reduction="batchmean")(log_softmax(s*inv_temp), softmax(t*inv_temp)):
 *   loss_rows[r] = sum_k p_t (log p_t - log p_s); ds (optional) = gscale * d loss_rows[r] / d s[r, :].
 *   Replaces the softmax / log_softmax / KLDivLoss chains of DistillationV3Loss.forward,
 *   LT/_methods/distillationv3/distillationv3_loss.py:60-84 (queue logits) and :86-115 (token-token logits). */
int b200_rope_apply(void* qkv, long long ld, int B, int N, int prefix, int h, int head_dim, const float* sin_tab,
                    const float* cos_tab, void* stream);
int b200_kl_rows(const float* s, long long lds, const float* t, long long ldt, int R, int K, float inv_temp, float gscale,
                 float* loss_rows, float* ds, long long ldds, void* stream);
/* out[0] += mean((teacher - student)^2); ds (optional) = d/d student.  DistillationV2Loss.forward,
 * LT/_methods/distillationv2/distillationv2_loss.py:27-44 (MSELoss, reduction "mean") + its autograd backward. */
int b200_mse(const float* teacher, const float* student, long long n, float* out, float* ds, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention core for short sequences (head_dim 64), forward and backward.
 * Replaces LT/_models/dinov2_vit/dinov2_vit_src/layers/attention.py:55-63 (q*scale @ k^T, softmax, @ v) and
 * its autograd backward.  qkv: bf16 [B, N, 3, h, 64] with token pitch ld_tok; out/dout: bf16 [B, N, h*64]
 * with token pitch ld_out; lse: f32 [B*h, N] (log-sum-exp of the scaled, bf16-rounded scores).
 * dqkv has the layout of qkv (token pitch ld_dtok).  N <= 272.
 */
int b200_attention_fwd(const void* qkv, long long ld_tok, int B, int N, int h, int head_dim, float scale,
                       void* out, long long ld_out, float* lse, void* stream);
int b200_attention_bwd(const void* qkv, long long ld_tok, const void* out, const void* dout, long long ld_out,
                       const float* lse, int B, int N, int h, int head_dim, float scale, void* dqkv,
                       long long ld_dtok, float* dqkv_colsum /* optional [3*h*64] f32: += column sums of dqkv (= the qkv bias gradient); NULL to skip */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-wise kernels around the GEMMs.
 */
/* nn.LayerNorm(eps) forward (vision_transformer.py:138,377; block.py:92,95). x f32 [T,D] -> y bf16|f32. */
int b200_layernorm_fwd(const float* x, long long ldx, int T, int D, const float* w, const float* b, float eps,
                       void* y, long long ldy, int out_bf16, float* mean, float* rstd, void* stream);
/* LayerNorm backward: dx (+)= LN'(dy); dw/db += column sums (may be NULL). dy bf16|f32. */
int b200_layernorm_bwd(const void* dy, long long lddy, int dy_bf16, const float* x, long long ldx, int T, int D,
                       const float* w, const float* mean, const float* rstd, float* dx, long long lddx,
                       int accumulate, float* dw, float* db, void* stream);
/* LayerNorm backward fused with the LayerScale(+DropPath) backward that follows it in the block schedule:
 * dx (+)= LN'(dy); dout = bf16(dx*rowscale*gamma); dgamma += sum dx*rowscale*o; dbias += sum dout. */
int b200_layernorm_bwd_ls(const void* dy, long long lddy, int dy_bf16, const float* x, long long ldx, int T, int D,
                          const float* w, const float* mean, const float* rstd, float* dx, long long lddx,
                          int accumulate, float* dw, float* db, const void* o, long long ldo, const float* gamma,
                          const float* rowscale, int rows_per_scale, void* dout, long long lddo, float* dgamma,
                          float* dbias, void* stream);
/* im2col for PatchEmbed.proj = Conv2d(k=s=p) (patch_embed.py:77-79,108): x f32 [B,C,H,W] -> bf16 [B*Np, C*p*p]. */
int b200_im2col(const float* x, int B, int C, int H, int W, int p, void* cols, long long ldc, void* stream);
/* prepare_tokens_with_masks (vision_transformer.py:307-329): mask-token select, cls/register concat, +pos. */
int b200_assemble_tokens(const void* tok, long long ldt, const unsigned char* masks, const float* mask_token,
                         const float* cls, const float* reg, const float* pos, int B, int Np, int R, int D,
                         float* x, void* stream);
int b200_assemble_tokens_bwd(const float* dx, const unsigned char* masks, int B, int Np, int R, int D, void* dtok,
                             long long lddt, float* dpos, float* dcls, float* dreg, float* dmask_token,
                             void* stream);
/* LayerScale (+DropPath) backward (layer_scale.py:27-28, block.py:109-114): do = bf16(dx*rowscale*gamma),
 * dgamma += sum dx*rowscale*o, dbias += sum do. */
int b200_layerscale_bwd(const float* dx, long long lddx, const void* o, long long ldo, const float* gamma,
                        const float* rowscale, int rows_per_scale, int T, int D, void* dout, long long lddo,
                        float* dgamma, float* dbias, void* stream);
/* SwiGLU FFN gate (SwiGLUFFN.forward, LT/_models/dinov2_vit/dinov2_vit_src/layers/swiglu_ffn.py:31-35):
 * x12 bf16 [T, 2H] = w12(x); hidden[t, j] = bf16( bf16(silu(x12[t, j])) * x12[t, H + j] ) -- the two roundings of the
 * bf16-autocast reference.  Backward: dx12[:, :H] = bf16( bf16(dh * x2) * silu'(x1) ), dx12[:, H:] = bf16( dh * bf16(silu(x1)) ).
 * H % 8 == 0, row pitches % 8 == 0. */
int b200_swiglu_fwd(const void* x12, long long ld12, int T, int H, void* hidden, long long ldh, void* stream);
int b200_swiglu_bwd(const void* x12, long long ld12, const void* dhidden, long long lddh, int T, int H, void* dx12,
                    long long lddx12, void* stream);
/* Device-side block-wise mask generation (the input side of the step).  Same algorithm and
 * distribution as MaskingGenerator.__call__ / _mask (LT/_methods/dinov2/utils.py:41-113) and the collation of
 * create_collated_masks (:116-152), counter-based device RNG instead of python's `random` (statistical, not bit, parity).
 * b200_block_masks: targets int32 [B] (per-crop mask counts, 0 = unmasked crop) -> masks u8 [B, H*W]; one RNG stream per
 *   (seed, *step_dev, crop).
 * b200_collate_masks: masks -> idx int64 [cap] (mask_indices_list), weight f32 [cap] (masks_weight), row_w (1/M), pad
 *   (0 | -1e30 for rows >= M), m_valid int32 [1]; rows >= M are inert (idx 0, weight 0). */
int b200_block_masks(const int* targets, int B, int H, int W, int min_patches, int max_patches, float min_aspect,
                     float max_aspect, long long seed, const int* step_dev, unsigned char* masks, void* stream);
int b200_collate_masks(const unsigned char* masks, int B, int Np, int cap, long long* idx, float* weight, float* row_w,
                       float* pad, int* m_valid, void* stream);
/* Random batch subset of stochastic depth (drop_add_residual_stochastic_depth, LT/_models/dinov2_vit/dinov2_vit_src/
 * layers/block.py:118-141: `torch.randperm(b)[:sample_subset_size]`): idx_out int64 [k] = a uniformly random k-subset of
 * {0..n-1} in random order (the k smallest of n counter-based random keys; statistical parity, not torch's stream).
 * *counter_dev (int64 in device memory) is read and incremented by the launch: graph replays draw fresh subsets.  n <= 2048. */
int b200_random_subset(int n, int k, long long seed, long long* counter_dev, long long* idx_out, void* stream);
/* Batch-subset stochastic depth (drop_add_residual_stochastic_depth, LT/_models/dinov2_vit/dinov2_vit_src/layers/block.py:118-141):
 * sample-granular copies of the fp32 residual stream, rows of row_elems floats (% 4 == 0).
 * scatter = 0: dst[j, :] = src[idx[j], :] (x[brange]);  scatter = 1: dst[idx[j], :] = src[j, :] (the index_add target rows). */
int b200_copy_samples(const float* src, float* dst, const long long* idx, int n_idx, long long row_elems, int scatter,
                      void* stream);
/* torch.index_select of token rows (dinov2.py:427-431,496-500) and its backward (scatter).
 * Row index = Np>0 ? (idx/Np)*N + off + idx%Np : idx. */
int b200_gather_rows(const float* src, long long lds, const long long* idx, int M, int D, int Np, int N, int off,
                     void* out, long long ldo, int out_bf16, void* stream);
int b200_scatter_rows(const void* in, long long ldi, int in_bf16, const long long* idx, int M, int D, int Np,
                      int N, int off, float* dst, long long ldd, int accumulate, const int* count_dev,
                      void* stream); /* count_dev: optional device int; rows m >= *count_dev are skipped */
/* F.normalize(p=2, eps) on bf16 rows (dinov2_head.py:68-69) and backward. */
int b200_l2norm_fwd(const void* x, long long ldx, int R, int D, float eps, void* y, long long ldy, float* nrm,
                    void* stream);
int b200_l2norm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const float* nrm, int R, int D,
                    void* dx, long long lddx, void* stream);
/* parametrizations.weight_norm (dinov2_head.py:54-58): w bf16 [O,I] = g*v/||v||; backward dg, dv +=. */
int b200_weightnorm_fwd(const float* g, const float* v, int O, int I, void* w, float* vnorm, void* stream);
int b200_weightnorm_bwd(const float* dW, const float* g, const float* v, int O, int I, float* dg, float* dv,
                        void* stream);
/* tiny fp32 matmul C (+)= op(A) * B, B row-major [K,N]; positional-embedding resampling operator
 * (interpolate_pos_encoding, vision_transformer.py:251-305, as a fixed linear map and its transpose). */
int b200_small_matmul(const float* A, long long lda, int a_trans, const float* B, long long ldb, int M, int N,
                      int K, float* C, long long ldc, int accumulate, void* stream);
int b200_cast_bf16(const float* x, void* y, long long n, void* stream);
int b200_fill_f32(float* x, long long n, float v, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DINO / iBOT loss path (LT/_methods/dinov2/dinov2_loss.py). Teacher probabilities are represented as
 *   p[b,k] = exp(t[b,k]*t_scale + colterm[k] + rowterm[b])   and never materialised.
 * `*_dev` arguments: optional DEVICE scalar overriding the by-value scalar of the same name (lets a captured
 * CUDA graph follow the teacher-temperature schedule without re-capture); NULL = use the by-value argument.
 */
/* rowterm[r] = -logsumexp_k(x[r,k]*scale + colterm[k]); x bf16 [R,K]. (softmax_center_teacher :76-82,
 * last column normalisation of sinkhorn_knopp_teacher :109-113) */
int b200_row_lse(const void* x, long long ld, int R, int K, const float* colterm, float scale,
                 const float* scale_dev, float* rowterm, void* stream);
/* out[k] += sum_r x[r,k]*rowvec[r] (mode 0; center batch sums :135-145, bias grads) or
 * out[k] += sum_r exp(x[r,k]*scale + rowvec[r]) (mode 1; Sinkhorn prototype sums :100-104). */
int b200_col_reduce(const void* x, long long ld, int R, int K, const float* rowvec, float scale,
                    const float* scale_dev, int mode, float* out, void* stream);
/* K-vector helpers: op0 y=a*x+b*y (center EMA :148-160), op1 y=-a*x (colterm from center),
 * op2 y=-log(x)-a (Sinkhorn log-scaling). */
int b200_vec_op(float* y, const float* x, int n, float a, float b, int op, const float* a_dev, void* stream);
/* Fused CE forward+backward, one student row per CTA (DINOLoss.forward :117-133, forward_masked :246-268):
 * loss_rows[r] = w*(n_t*LSE_s - s_scale*sum_k p_t[k]*s[k]); ds = w*s_scale*(n_t*softmax_s - p_t)*gscale. */
int b200_dino_ce(const void* s, long long lds, int Rs, int K, const void* t, long long ldt, const float* colterm,
                 const float* t_rowterm, const int* t_idx0, const int* t_idx1, const float* weight,
                 float s_scale, float t_scale, const float* t_scale_dev, float gscale, float* loss_rows, void* ds,
                 long long ldds, void* stream);
int b200_segment_sum(const float* x, const int* offsets, int n_segments, const float* scale, float* out,
                     void* stream);
/* KoLeoLoss forward+backward (lightly.loss.KoLeoLoss; call site dinov2.py:377-380), `groups` independent
 * sets of n rows. dx += gscale * dloss/dx.  Groups whose 2*n*D fp32 values exceed one CTA's shared memory (or n > 256)
 * run tiled over rows and need `scratch` (>= groups*n*(2*D + 1) floats); smaller groups ignore it (may be NULL). */
int b200_koleo(const float* x, long long ldx, int groups, int n, int D, float eps, int bf16_sim, float gscale,
               float* loss_out, float* dx, long long lddx, int* nn_out, float* scratch, long long scratch_elems,
               void* stream);

/* ------------------------------------------------------------------------------------------------
 * Parameter sweeps over flat fp32 arenas.
 */
/* update_momentum (LT/_torch_helpers.py:75-96): teacher = teacher*m + student*(1-m), in place. */
int b200_ema(float* teacher, const float* student, long long n, float m, void* teacher_bf16, void* stream);
int b200_sumsq(const float* x, long long n, float* out, void* stream);

typedef struct b200_adamw_args {
  float* p; const float* g; float* m; float* v;
  float* t;                   /* teacher (EMA) arena or NULL                                         */
  void* p_bf16; void* t_bf16; /* optional bf16 shadows refreshed in the same pass                    */
  long long n; int chunk;     /* hyper-parameters are constant per `chunk` elements                  */
  const float* lr_scale;      /* [n/chunk] layer-wise lr decay * patch-embed multiplier              */
  const float* wd_scale;      /* [n/chunk] 1 = decayed, 0 = bias/norm/gamma                          */
  const unsigned char* flags; /* [n/chunk] bit0 last_layer, bit1 backbone                            */
  float lr, wd, beta1, beta2, eps;
  int step;                   /* 1-based                                                             */
  float ema_m;
  const float* gradnorm_sq;   /* device scalar (sum of squares of g) or NULL = no clipping           */
  const float* dyn;           /* optional device [7]: lr, wd, 1-b1^t, sqrt(1-b2^t), ema_m, freeze_last,
                                 freeze_backbone -- overrides the by-value fields (CUDA-graph replay)  */
  float max_norm;
  float grad_scale;           /* multiplies g (e.g. 1/world_size)                                    */
  int freeze_last_layer, freeze_backbone;
} b200_adamw_args;
/* clip_grad_norm_ + AdamW.step + update_momentum in one sweep (dinov2.py:588-660, utils.py:191-273). */
int b200_adamw_ema(const b200_adamw_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DINO_H_ */
