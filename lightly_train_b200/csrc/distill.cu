// Distillation path (BASELINE cfg4): the kernels the DINOv3 teacher forward and the
// DistillationV3 loss need on top of the ViT kernels (GEMM / attention / LayerNorm are shared with the DINOv2 path).
//
//   rope_apply_kernel      axial RoPE on q and k of the PATCH tokens, in place in the bf16 qkv buffer
//                          (LT/_models/dinov3/dinov3_src/layers/attention.py:21-33,79-100: x*cos + rotate_half(x)*sin in the
//                          rope dtype (fp32 for the hub models), cls + storage tokens untouched, result cast back to bf16)
//   kl_rows_kernel         KLDivLoss(batchmean)(log_softmax(s/T), softmax(t/T)) over rows of two logit matrices, forward
//                          value per row + analytic gradient wrt the student logits in one pass
//                          (LT/_methods/distillationv3/distillationv3_loss.py:60-84 global term, :86-115 local term)
// HBM-bound row kernels: one warp per row for short rows (token-token similarities, 196..1024 columns), one CTA per row
// for the queue logits (8192 columns).
#include "common.cuh"
#include "../../include/b200dino.h"

namespace b200 {

// qkv: bf16 [B*N, 3*h*64]; sin / cos: f32 [N - prefix, 64]; one thread handles 8 consecutive dims of one (token, q|k, head)
// together with its rotate-half partner chunk (dims j and j +- 32), i.e. one thread per pair of 16-byte chunks.
__global__ void rope_apply_kernel(__nv_bfloat16* __restrict__ qkv, long long ld, int B, int N, int prefix, int h,
                                  const float* __restrict__ sn, const float* __restrict__ cs) {
  B200_PDL_SYNC();
  const int P = N - prefix;
  const long long total = (long long)B * P * 2 * h * 4;  // 4 chunk pairs (8 dims each, first half) per head vector
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 3);
    long long r = i >> 2;
    const int head = (int)(r % h); r /= h;
    const int which = (int)(r & 1); r >>= 1;  // 0: q, 1: k
    const int p = (int)(r % P);
    const int b = (int)(r / P);
    __nv_bfloat16* v = qkv + ((long long)b * N + prefix + p) * ld + (long long)which * h * 64 + head * 64 + c * 8;
    const uint4 lo = *reinterpret_cast<const uint4*>(v), hi = *reinterpret_cast<const uint4*>(v + 32);
    const uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w}, hw[4] = {hi.x, hi.y, hi.z, hi.w};
    const float* s0 = sn + (long long)p * 64 + c * 8;
    const float* c0 = cs + (long long)p * 64 + c * 8;
    uint32_t ol[4], oh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = unpack_bf16x2(lw[k]), bq = unpack_bf16x2(hw[k]);
      const float sl0 = s0[2 * k], sl1 = s0[2 * k + 1], cl0 = c0[2 * k], cl1 = c0[2 * k + 1];
      const float sh0 = s0[32 + 2 * k], sh1 = s0[32 + 2 * k + 1], ch0 = c0[32 + 2 * k], ch1 = c0[32 + 2 * k + 1];
      // out[j] = x[j] cos[j] - x[j+32] sin[j]   (j < 32) ;   out[j+32] = x[j+32] cos[j+32] + x[j] sin[j+32]
      ol[k] = pack_bf16x2(fmaf(a.x, cl0, -bq.x * sl0), fmaf(a.y, cl1, -bq.y * sl1));
      oh[k] = pack_bf16x2(fmaf(bq.x, ch0, a.x * sh0), fmaf(bq.y, ch1, a.y * sh1));
    }
    *reinterpret_cast<uint4*>(v) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    *reinterpret_cast<uint4*>(v + 32) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
  }
}

// One row per warp (WARP_ROWS = true) or per CTA.  s, t: f32 [R, K] similarity logits; inv_temp = 1 / temperature.
//   loss_row = sum_k p_t (log p_t - log p_s),   ds[k] = gscale * (p_s - p_t) * inv_temp      (d loss_row / d s[k])
template <bool WARP_ROWS>
__global__ void __launch_bounds__(256) kl_rows_kernel(const float* __restrict__ s, long long lds, const float* __restrict__ t,
                                                      long long ldt, int R, int K, float inv_temp, float gscale,
                                                      float* __restrict__ loss_rows, float* __restrict__ ds, long long ldds) {
  B200_PDL_SYNC();
  __shared__ float red[32];
  const int lane = threadIdx.x & 31;
  const int row = WARP_ROWS ? blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : blockIdx.x;
  if (WARP_ROWS && row >= R) return;
  const int tid = WARP_ROWS ? lane : threadIdx.x, nt = WARP_ROWS ? 32 : blockDim.x;
  const float* sr = s + (long long)row * lds;
  const float* tr = t + (long long)row * ldt;
  const float k2 = inv_temp * 1.4426950408889634f;
  float ms = -INFINITY, mt = -INFINITY;
  for (int k = tid; k < K; k += nt) {
    ms = fmaxf(ms, sr[k]);
    mt = fmaxf(mt, tr[k]);
  }
  ms = WARP_ROWS ? warp_max(ms) : block_max(ms, red);
  mt = WARP_ROWS ? warp_max(mt) : block_max(mt, red);
  float zs = 0.f, zt = 0.f, cross = 0.f;  // cross = sum_k e_t[k] * (t[k] - s[k])
  for (int k = tid; k < K; k += nt) {
    const float a = sr[k], b = tr[k];
    const float et = ex2_ftz((b - mt) * k2);
    zs += ex2_ftz((a - ms) * k2);
    zt += et;
    cross = fmaf(et, b - a, cross);
  }
  zs = WARP_ROWS ? warp_sum(zs) : block_sum(zs, red);
  zt = WARP_ROWS ? warp_sum(zt) : block_sum(zt, red);
  cross = WARP_ROWS ? warp_sum(cross) : block_sum(cross, red);
  // sum_k p_t (log p_t - log p_s) = inv_temp * E_t[t - s] - (lse_t - lse_s)
  const float lse_s = ms * inv_temp + __logf(zs), lse_t = mt * inv_temp + __logf(zt);
  if (tid == 0) loss_rows[row] = inv_temp * cross / zt - (lse_t - lse_s);
  if (ds) {
    float* dr = ds + (long long)row * ldds;
    const float izs = 1.f / zs, izt = 1.f / zt, g = gscale * inv_temp;
    for (int k = tid; k < K; k += nt)
      dr[k] = g * (ex2_ftz((sr[k] - ms) * k2) * izs - ex2_ftz((tr[k] - mt) * k2) * izt);
  }
}

// out[0] += mean((t - s)^2) ; ds = 2 (s - t) / n   (MSELoss(reduction="mean") and its gradient wrt s, one pass)
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ t, const float* __restrict__ s, long long n,
                                                  float* __restrict__ out, float* __restrict__ ds) {
  B200_PDL_SYNC();
  __shared__ float red[32];
  const float inv_n = 1.f / (float)n;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = s[i] - t[i];
    acc = fmaf(d, d, acc);
    if (ds) ds[i] = 2.f * d * inv_n;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc * inv_n);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_mse(const float* teacher, const float* student, long long n, float* out, float* ds, void* stream) {
  if (!teacher || !student || !out || n <= 0) return B200_ERR_INVALID_ARG;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_kernel(mse_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, teacher, student, n, out, ds);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_rope_apply(void* qkv, long long ld, int B, int N, int prefix, int h, int head_dim, const float* sin_tab,
                               const float* cos_tab, void* stream) {
  if (!qkv || !sin_tab || !cos_tab || B <= 0 || N <= 0 || prefix < 0 || prefix > N || h <= 0) return B200_ERR_INVALID_ARG;
  if (head_dim != 64 || (ld % 8) || ((uintptr_t)qkv & 15)) return B200_ERR_UNSUPPORTED;
  if (prefix == N) return B200_OK;
  const long long total = (long long)B * (N - prefix) * 2 * h * 4;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  launch_kernel(rope_apply_kernel, grid, 256, 0, (cudaStream_t)stream, (__nv_bfloat16*)qkv, ld, B, N, prefix, h, sin_tab, cos_tab);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_kl_rows(const float* s, long long lds, const float* t, long long ldt, int R, int K, float inv_temp,
                            float gscale, float* loss_rows, float* ds, long long ldds, void* stream) {
  if (!s || !t || !loss_rows || R <= 0 || K <= 0 || !(inv_temp > 0.f)) return B200_ERR_INVALID_ARG;
  if (K <= 1024)
    launch_kernel(kl_rows_kernel<true>, (R + 7) / 8, 256, 0, (cudaStream_t)stream, s, lds, t, ldt, R, K, inv_temp, gscale, loss_rows, ds, ldds);
  else
    launch_kernel(kl_rows_kernel<false>, R, 256, 0, (cudaStream_t)stream, s, lds, t, ldt, R, K, inv_temp, gscale, loss_rows, ds, ldds);
  B200_CHECK_LAUNCH();
  return B200_OK;
}
