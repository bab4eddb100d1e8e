// Parameter-sweep kernels over flat fp32 arenas: EMA teacher update, gradient-norm, fused
// clip + AdamW + EMA + bf16 shadow refresh.  HBM-bound, 128-bit coalesced access, one pass.
//
// Reference arithmetic replaced (LT = src/lightly_train):
//   update_momentum / update_ema_tensors       LT/_torch_helpers.py:75-96   (2 foreach passes, 20 B/param)
//   clip_gradients(norm=3.0)                   LT/_methods/dinov2/dinov2.py:588-598 (torch clip_grad_norm_)
//   AdamW (foreach) with per-group lr / wd     LT/_optim/adamw_args.py:33-36, LT/_methods/dinov2/utils.py:191-273
//   last-layer / backbone lr freeze, wd sched  LT/_methods/dinov2/dinov2.py:600-639
#include "common.cuh"
#include "../../include/b200dino.h"

namespace b200 {

__global__ void __launch_bounds__(256) ema_kernel(float* __restrict__ t, const float* __restrict__ s, long long n, float m,
                                                  __nv_bfloat16* __restrict__ t_bf16) {
  B200_PDL_SYNC();
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 a = *reinterpret_cast<const float4*>(t + i);
      const float4 b = *reinterpret_cast<const float4*>(s + i);
      // torch: ema.mul_(m).add_(p, alpha=1-m)
      a.x = a.x * m + b.x * (1.f - m); a.y = a.y * m + b.y * (1.f - m);
      a.z = a.z * m + b.z * (1.f - m); a.w = a.w * m + b.w * (1.f - m);
      *reinterpret_cast<float4*>(t + i) = a;
      if (t_bf16) {
        uint2 p; p.x = pack_bf16x2(a.x, a.y); p.y = pack_bf16x2(a.z, a.w);
        *reinterpret_cast<uint2*>(t_bf16 + i) = p;
      }
    } else {
      for (long long j = i; j < n; ++j) {
        float a = t[j] * m + s[j] * (1.f - m);
        t[j] = a;
        if (t_bf16) t_bf16[j] = __float2bfloat16_rn(a);
      }
    }
  }
}

// Sum of squares, DETERMINISTIC: every block writes its partial sum to a fixed slot, the last block to finish (ticket
// counter) adds the slots in index order with a fixed reduction tree.  Data-parallel replicas compute the gradient norm
// from bit-identical all-reduced gradients, so the clip coefficient -- and therefore the weights -- stay bit-identical
// across ranks (a float atomicAdd per block would make the sum depend on block retirement order).
// One launch in flight per device (the scratch below is per device, not per stream): the optimizer sweep is the only user.
static constexpr int SUMSQ_MAX_BLOCKS = 4096;
__device__ float g_sumsq_partials[SUMSQ_MAX_BLOCKS];
__device__ unsigned int g_sumsq_ticket = 0;

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  B200_PDL_SYNC();
  __shared__ float red[32];
  __shared__ bool last;
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 a = *reinterpret_cast<const float4*>(x + i);
      acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    } else {
      for (long long j = i; j < n; ++j) acc += x[j] * x[j];
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    g_sumsq_partials[blockIdx.x] = acc;
    __threadfence();
    last = atomicAdd(&g_sumsq_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float tot = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) tot += __ldcg(&g_sumsq_partials[i]);
  tot = block_sum(tot, red);
  if (threadIdx.x == 0) {
    *out += tot;
    g_sumsq_ticket = 0;
  }
}

struct AdamDev {
  float* p; const float* g; float* m; float* v; float* t;
  __nv_bfloat16* p_bf16; __nv_bfloat16* t_bf16;
  long long n; int chunk;
  const float* lr_scale; const float* wd_scale; const unsigned char* flags;
  float lr, wd, beta1, beta2, eps, bc1, bc2_sqrt, ema_m, max_norm, grad_scale;
  const float* gradnorm_sq;
  const float* dyn;
  int freeze_last_layer, freeze_backbone;
};

__global__ void __launch_bounds__(256) adamw_ema_kernel(AdamDev a) {
  B200_PDL_SYNC();
  if (a.dyn) {  // per-step scalars from device memory (CUDA-graph replay): see b200_adamw_args.dyn
    a.lr = a.dyn[0]; a.wd = a.dyn[1]; a.bc1 = a.dyn[2]; a.bc2_sqrt = a.dyn[3]; a.ema_m = a.dyn[4];
    a.freeze_last_layer = a.dyn[5] != 0.f; a.freeze_backbone = a.dyn[6] != 0.f;
  }
  float coef = a.grad_scale;
  if (a.gradnorm_sq) {
    // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float total = sqrtf(*a.gradnorm_sq) * a.grad_scale;
    coef *= fminf(a.max_norm / (total + 1e-6f), 1.0f);
  }
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < a.n; i += stride) {
    const long long c = i / a.chunk;
    const unsigned char fl = a.flags ? a.flags[c] : 0;
    float lr = a.lr * (a.lr_scale ? a.lr_scale[c] : 1.f);
    if ((a.freeze_last_layer && (fl & 1)) || (a.freeze_backbone && (fl & 2))) lr = 0.f;
    const float wd = a.wd * (a.wd_scale ? a.wd_scale[c] : 1.f);
    float4 p = *reinterpret_cast<const float4*>(a.p + i);
    float4 g = *reinterpret_cast<const float4*>(a.g + i);
    float4 m = *reinterpret_cast<const float4*>(a.m + i);
    float4 v = *reinterpret_cast<const float4*>(a.v + i);
    float* pp = &p.x; float* gp = &g.x; float* mp = &m.x; float* vp = &v.x;
    const float step_size = lr / a.bc1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gg = gp[k] * coef;
      pp[k] *= (1.f - lr * wd);
      mp[k] = mp[k] + (gg - mp[k]) * (1.f - a.beta1);          // exp_avg.lerp_(grad, 1-beta1)
      vp[k] = vp[k] * a.beta2 + gg * gg * (1.f - a.beta2);     // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
      const float denom = sqrtf(vp[k]) / a.bc2_sqrt + a.eps;
      pp[k] -= step_size * (mp[k] / denom);
    }
    *reinterpret_cast<float4*>(a.p + i) = p;
    *reinterpret_cast<float4*>(a.m + i) = m;
    *reinterpret_cast<float4*>(a.v + i) = v;
    if (a.p_bf16) {
      uint2 o; o.x = pack_bf16x2(p.x, p.y); o.y = pack_bf16x2(p.z, p.w);
      *reinterpret_cast<uint2*>(a.p_bf16 + i) = o;
    }
    if (a.t) {
      float4 t = *reinterpret_cast<const float4*>(a.t + i);
      t.x = t.x * a.ema_m + p.x * (1.f - a.ema_m); t.y = t.y * a.ema_m + p.y * (1.f - a.ema_m);
      t.z = t.z * a.ema_m + p.z * (1.f - a.ema_m); t.w = t.w * a.ema_m + p.w * (1.f - a.ema_m);
      *reinterpret_cast<float4*>(a.t + i) = t;
      if (a.t_bf16) {
        uint2 o; o.x = pack_bf16x2(t.x, t.y); o.y = pack_bf16x2(t.z, t.w);
        *reinterpret_cast<uint2*>(a.t_bf16 + i) = o;
      }
    }
  }
}

__global__ void fill_kernel(float* __restrict__ x, long long n, float v) {
  B200_PDL_SYNC();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = v;
}

}  // namespace b200

using namespace b200;

static inline int sweep_grid(long long n) {
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

extern "C" int b200_ema(float* teacher, const float* student, long long n, float m, void* teacher_bf16, void* stream) {
  if (!teacher || !student || n <= 0) return B200_ERR_INVALID_ARG;
  if (((uintptr_t)teacher & 15) || ((uintptr_t)student & 15)) return B200_ERR_UNSUPPORTED;
  launch_kernel(ema_kernel, sweep_grid(n), 256, 0, (cudaStream_t)stream, teacher, student, n, m, (__nv_bfloat16*)teacher_bf16);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_sumsq(const float* x, long long n, float* out, void* stream) {
  if (!x || !out || n <= 0 || ((uintptr_t)x & 15)) return B200_ERR_INVALID_ARG;
  int grid = sweep_grid(n);
  if (grid > SUMSQ_MAX_BLOCKS) grid = SUMSQ_MAX_BLOCKS;
  launch_kernel(sumsq_kernel, grid, 256, 0, (cudaStream_t)stream, x, n, out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_adamw_ema(const b200_adamw_args* a, void* stream) {
  if (!a || !a->p || !a->g || !a->m || !a->v || a->n <= 0 || a->step < 1) return B200_ERR_INVALID_ARG;
  if (a->chunk <= 0 || (a->chunk % 4) || (a->n % 4)) return B200_ERR_INVALID_ARG;
  AdamDev d;
  d.p = a->p; d.g = a->g; d.m = a->m; d.v = a->v; d.t = a->t;
  d.p_bf16 = (__nv_bfloat16*)a->p_bf16; d.t_bf16 = (__nv_bfloat16*)a->t_bf16;
  d.n = a->n; d.chunk = a->chunk;
  d.lr_scale = a->lr_scale; d.wd_scale = a->wd_scale; d.flags = a->flags;
  d.lr = a->lr; d.wd = a->wd; d.beta1 = a->beta1; d.beta2 = a->beta2; d.eps = a->eps;
  d.bc1 = (float)(1.0 - pow((double)a->beta1, (double)a->step));
  d.bc2_sqrt = (float)sqrt(1.0 - pow((double)a->beta2, (double)a->step));
  d.ema_m = a->ema_m; d.max_norm = a->max_norm; d.grad_scale = a->grad_scale;
  d.gradnorm_sq = a->gradnorm_sq;
  d.dyn = a->dyn;
  d.freeze_last_layer = a->freeze_last_layer; d.freeze_backbone = a->freeze_backbone;
  launch_kernel(adamw_ema_kernel, sweep_grid(a->n), 256, 0, (cudaStream_t)stream, d);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_fill_f32(float* x, long long n, float v, void* stream) {
  if (!x || n <= 0) return B200_ERR_INVALID_ARG;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_kernel(fill_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, x, n, v);
  B200_CHECK_LAUNCH();
  return B200_OK;
}
