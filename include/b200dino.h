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
  B200_EPI_DGELU = 5       /* C(bf16) = bf16(acc) * gelu_erf'(aux(bf16))                           */
};

typedef struct b200_gemm_args {
  const void* A; long long lda; int a_mn;
  const void* B; long long ldb; int b_mn;
  int M, N, K;
  int splits;            /* split-K factor (>1 only with B200_EPI_F32_ATOMIC)                       */
  int epi;               /* B200_EPI_*                                                              */
  int block_n;           /* 0 = auto, or 128 / 192 / 256                                            */
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

#ifdef __cplusplus
}
#endif
#endif /* B200DINO_H_ */
