// DINO / iBOT loss path: teacher centering (softmax or Sinkhorn-Knopp), student cross-entropy forward +
// analytic backward, KoLeo regulariser.  HBM-bound kernels: every logit is read from HBM once.
//
// Reference arithmetic replaced (LT = src/lightly_train):
//   LT/_methods/dinov2/dinov2_loss.py:76-82,178-186   softmax_center_teacher
//   LT/_methods/dinov2/dinov2_loss.py:84-115,188-224  sinkhorn_knopp_teacher
//   LT/_methods/dinov2/dinov2_loss.py:117-133         DINOLoss.forward
//   LT/_methods/dinov2/dinov2_loss.py:246-268,55-56   IBOTPatchLoss.forward_masked / lossfunc
//   LT/_methods/dinov2/dinov2_loss.py:135-160,270-297 center update
//   lightly.loss.KoLeoLoss (call site LT/_methods/dinov2/dinov2.py:377-380)
//
// Representation: teacher probabilities are never materialised.  For both centering methods
//     p[b,k] = exp(t[b,k]*t_scale + colterm[k] + rowterm[b])
//   softmax-centering: colterm = -center*t_scale,           rowterm = -LSE_k(t*t_scale + colterm)
//   Sinkhorn-Knopp   : colterm = log u (3 scaling iters),    rowterm = -LSE_k(t*t_scale + colterm)
// (Sinkhorn is a diagonal scaling Q = diag(u) E diag(v); the last column normalisation *is* the softmax
//  normalisation, so both methods share the CE kernel.)
#include "common.cuh"
#include "../../include/b200dino.h"

namespace b200 {

static constexpr int LOSS_THREADS = 512;

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Online (max, sum-exp) merge helpers
struct MaxSum {
  float m, s;
};
__device__ __forceinline__ MaxSum ms_merge(MaxSum a, MaxSum b) {
  float m = fmaxf(a.m, b.m);
  if (m == -INFINITY) return {m, 0.f};
  return {m, a.s * __expf(a.m - m) + b.s * __expf(b.m - m)};
}
__device__ __forceinline__ MaxSum block_maxsum(MaxSum v, MaxSum* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxSum w{__shfl_xor_sync(0xffffffffu, v.m, o), __shfl_xor_sync(0xffffffffu, v.s, o)};
    v = ms_merge(v, w);
  }
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  MaxSum r = (lane < nw) ? red[lane] : MaxSum{-INFINITY, 0.f};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxSum w{__shfl_xor_sync(0xffffffffu, r.m, o), __shfl_xor_sync(0xffffffffu, r.s, o)};
    r = ms_merge(r, w);
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// rowterm[r] = -LSE_k( x[r,k]*scale + colterm[k] )      one CTA per row
__global__ void __launch_bounds__(LOSS_THREADS) row_lse_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                                int R, int K, const float* __restrict__ colterm,
                                                                float scale, const float* __restrict__ scale_dev,
                                                                float* __restrict__ rowterm) {
  B200_PDL_SYNC();
  __shared__ MaxSum red[32];
  if (scale_dev) scale = __ldg(scale_dev);
  const int r = blockIdx.x;
  const __nv_bfloat16* xr = x + (size_t)r * ld;
  MaxSum acc{-INFINITY, 0.f};
  for (int k = threadIdx.x * 8; k < K; k += blockDim.x * 8) {
    float v[8], c[8];
    load8(xr + k, v);
    if (colterm) load8f(colterm + k, c);
    float z[8];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      z[i] = v[i] * scale + (colterm ? c[i] : 0.f);
      m = fmaxf(m, z[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __expf(z[i] - m);
    acc = ms_merge(acc, MaxSum{m, s});
  }
  MaxSum t = block_maxsum(acc, red);
  if (threadIdx.x == 0) rowterm[r] = -(t.m + __logf(t.s));
}

// ------------------------------------------------------------------------------------------------
// Column reductions over rows.  mode 0: out[k] += sum_r x[r,k] * rowweight[r]   (rowweight NULL -> 1)
//                               mode 1: out[k] += sum_r exp(x[r,k]*scale + rowterm[r])
// A CTA owns a slab of `cg` column groups (8 columns = one 16-byte load each) and a contiguous range of rows;
// its threads are arranged [row lanes][column groups] so every warp load is a run of consecutive 16-byte
// pieces of one or two rows (coalesced), 4 rows in flight per thread.  Row lanes are combined through smem,
// then one atomicAdd per column per CTA.
static constexpr int CR_THREADS = 512;
__global__ void __launch_bounds__(CR_THREADS) col_reduce_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int R, int K,
                                                                int cg, int rows_per_cta, const float* __restrict__ rowvec,
                                                                float scale, const float* __restrict__ scale_dev, int mode,
                                                                float* __restrict__ out) {
  B200_PDL_SYNC();
  extern __shared__ float cr_sm[];  // [lanes][cg*8]
  if (scale_dev) scale = __ldg(scale_dev);
  const int lanes = CR_THREADS / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int k = (blockIdx.x * cg + g) * 8;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes && k < K) {
    int r = r0 + lane;
    for (; r + 3 * lanes < r1; r += 4 * lanes) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8(x + (size_t)(r + u * lanes) * ld + k, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float w = rowvec ? __ldg(rowvec + r + u * lanes) : (mode == 0 ? 1.f : 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += (mode == 0) ? v[u][i] * w : __expf(v[u][i] * scale + w);
      }
    }
    for (; r < r1; r += lanes) {
      float v[8];
      load8(x + (size_t)r * ld + k, v);
      const float w = rowvec ? __ldg(rowvec + r) : (mode == 0 ? 1.f : 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += (mode == 0) ? v[i] * w : __expf(v[i] * scale + w);
    }
  }
  if (lane < lanes) {
#pragma unroll
    for (int i = 0; i < 8; ++i) cr_sm[(size_t)lane * cg * 8 + g * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cg * 8; c += CR_THREADS) {
    const int kk = blockIdx.x * cg * 8 + c;
    if (kk < K) {
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += cr_sm[(size_t)l * cg * 8 + c];
      atomicAdd(out + kk, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Small vector helpers for the centering state (K-length fp32 vectors).
// op 0: y = a*x + b*y          (center EMA: center = center*m + mean*(1-m))
// op 1: y = -a * x             (colterm = -center * t_scale)
// op 2: y = -log(x) - a        (Sinkhorn: log u = -log(sum) - log K)
__global__ void vec_op_kernel(float* __restrict__ y, const float* __restrict__ x, int n, float a, float b, int op,
                              const float* __restrict__ a_dev) {
  B200_PDL_SYNC();
  if (a_dev) a = __ldg(a_dev);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (op == 0) y[i] = a * x[i] + b * y[i];
  else if (op == 1) y[i] = -a * x[i];
  else y[i] = -__logf(x[i]) - a;
}

// ------------------------------------------------------------------------------------------------
// Fused cross-entropy forward + backward, one CTA per student row.
//   n_t   = number of paired teacher rows (1 or 2)
//   loss  = w * ( n_t * LSE_s - s_scale * sum_k p_t[k] * s[k] )
//   dS[k] = w * s_scale * ( n_t * softmax_s[k] - p_t[k] ) * gscale     (bf16, optional)
// Pass 1 streams the student row (HBM) for its log-sum-exp -- and, when t_rowterm is NULL (single-teacher rows: the iBOT
// term), the paired teacher row for ITS log-sum-exp, so that the teacher logits are read from HBM once per step instead
// of once by row_lse and once here.  Pass 2 re-reads the row(s) (L2-resident: 128 KB each), and writes the gradient row.
// All exponentials are raw ex2 with log2(e) folded into the scales (MUFU time ~ HBM time for this kernel: 4.25 ex2 per
// element against 6 bytes); two 16-byte loads per tensor are in flight per thread and iteration.
struct MaxSum2 {  // running (max, sum 2^(v - max)) in the base-2 domain
  float m, s;
};
__device__ __forceinline__ void ms2_add8(MaxSum2& a, const float (&z)[8]) {
  float cm = fmaxf(fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3])), fmaxf(fmaxf(z[4], z[5]), fmaxf(z[6], z[7])));
  const float mn = fmaxf(a.m, cm);
  float sum = a.s * ex2_ftz(a.m - mn);  // a.m = -inf on the first chunk: 2^-inf = 0
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += ex2_ftz(z[i] - mn);
  a.m = mn;
  a.s = sum;
}
__device__ __forceinline__ MaxSum2 ms2_merge(MaxSum2 a, MaxSum2 b) {
  const float m = fmaxf(a.m, b.m);
  if (m == -INFINITY) return {m, 0.f};
  return {m, a.s * ex2_ftz(a.m - m) + b.s * ex2_ftz(b.m - m)};
}
__device__ __forceinline__ MaxSum2 block_maxsum2(MaxSum2 v, MaxSum2* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxSum2 w{__shfl_xor_sync(0xffffffffu, v.m, o), __shfl_xor_sync(0xffffffffu, v.s, o)};
    v = ms2_merge(v, w);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  MaxSum2 r = (lane < nw) ? red[lane] : MaxSum2{-INFINITY, 0.f};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxSum2 w{__shfl_xor_sync(0xffffffffu, r.m, o), __shfl_xor_sync(0xffffffffu, r.s, o)};
    r = ms2_merge(r, w);
  }
  return r;
}

// THREADS / UNROLL2 were tuning parameters (profiles/r02_loss_bench.log; only <512, false> is instantiated): 512 threads with one
// 16-byte load per tensor in flight measured fastest (masked rows, fused LSE: 445 us; 256 threads x 4 rows per SM and / or two
// loads in flight: 486-545 us -- the kernel is bound by MUFU + issue, more registers per thread only cost occupancy).
template <bool FUSE_T_LSE, int THREADS, bool UNROLL2>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
dino_ce_kernel(const __nv_bfloat16* __restrict__ s, long long lds, int Rs, int K, const __nv_bfloat16* __restrict__ t,
               long long ldt, const float* __restrict__ colterm, const float* __restrict__ t_rowterm,
               const int* __restrict__ t_idx0, const int* __restrict__ t_idx1, const float* __restrict__ weight,
               float s_scale, float t_scale, const float* __restrict__ t_scale_dev, float gscale,
               float* __restrict__ loss_rows, __nv_bfloat16* __restrict__ ds, long long ldds) {
  B200_PDL_SYNC();
  __shared__ MaxSum2 red[32];
  __shared__ MaxSum2 red_t[32];
  __shared__ float redf[32];
  constexpr float kL2e = 1.4426950408889634f;
  if (t_scale_dev) t_scale = __ldg(t_scale_dev);
  const int r = blockIdx.x;
  const float w = weight ? __ldg(weight + r) : 1.f;
  const int i0 = t_idx0[r];
  const int i1 = (!FUSE_T_LSE && t_idx1) ? t_idx1[r] : -1;
  const float n_t = (i1 >= 0) ? 2.f : 1.f;
  const __nv_bfloat16* sr = s + (size_t)r * lds;
  const __nv_bfloat16* t0 = t + (size_t)i0 * ldt;
  const __nv_bfloat16* t1 = (i1 >= 0) ? t + (size_t)i1 * ldt : nullptr;
  const float ss2 = s_scale * kL2e, ts2 = t_scale * kL2e;
  const int step = blockDim.x * 8;

  // pass 1: log-sum-exp of the student row (base 2), and of the teacher row when fused
  MaxSum2 acc{-INFINITY, 0.f}, acc_t{-INFINITY, 0.f};
  for (int k = threadIdx.x * 8; k < K; k += (UNROLL2 ? 2 : 1) * step) {
    const bool two = UNROLL2 && k + step < K;
    float v0[8], v1[8], u0[8], u1[8], c0[8], c1[8];
    load8(sr + k, v0);
    if (two) load8(sr + k + step, v1);
    if (FUSE_T_LSE) {
      load8(t0 + k, u0);
      if (two) load8(t0 + k + step, u1);
      if (colterm) {
        load8f(colterm + k, c0);
        if (two) load8f(colterm + k + step, c1);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v0[i] *= ss2;
    ms2_add8(acc, v0);
    if (two) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v1[i] *= ss2;
      ms2_add8(acc, v1);
    }
    if (FUSE_T_LSE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) u0[i] = fmaf(u0[i], ts2, colterm ? c0[i] * kL2e : 0.f);
      ms2_add8(acc_t, u0);
      if (two) {
#pragma unroll
        for (int i = 0; i < 8; ++i) u1[i] = fmaf(u1[i], ts2, colterm ? c1[i] * kL2e : 0.f);
        ms2_add8(acc_t, u1);
      }
    }
  }
  const MaxSum2 st = block_maxsum2(acc, red);
  const float lse2 = st.m + __log2f(st.s);           // base-2 LSE of s * s_scale
  const float lse = lse2 * 0.6931471805599453f;
  float rt0, rt1 = 0.f;                               // base-2 row terms: p = 2^(t*ts2 + c*log2e + rt)
  if (FUSE_T_LSE) {
    const MaxSum2 tt = block_maxsum2(acc_t, red_t);
    rt0 = -(tt.m + __log2f(tt.s));
  } else {
    rt0 = __ldg(t_rowterm + i0) * kL2e;
    if (i1 >= 0) rt1 = __ldg(t_rowterm + i1) * kL2e;
  }

  // pass 2: dot(p_t, s) and gradient
  float dot = 0.f;
  const float gw = w * s_scale * gscale;
  for (int k0 = threadIdx.x * 8; k0 < K; k0 += (UNROLL2 ? 2 : 1) * step) {
#pragma unroll
    for (int u = 0; u < (UNROLL2 ? 2 : 1); ++u) {
      const int k = k0 + u * step;
      if (k >= K) break;
      float sv[8], tv[8], c[8], p[8];
      load8(sr + k, sv);
      load8(t0 + k, tv);
      if (colterm) load8f(colterm + k, c);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        c[i] = colterm ? c[i] * kL2e : 0.f;
        p[i] = ex2_ftz(fmaf(tv[i], ts2, c[i] + rt0));
      }
      if (t1) {
        load8(t1 + k, tv);
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] += ex2_ftz(fmaf(tv[i], ts2, c[i] + rt1));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) dot = fmaf(p[i], sv[i], dot);
      if (ds) {
        float g[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = gw * (n_t * ex2_ftz(fmaf(sv[i], ss2, -lse2)) - p[i]);
        uint4 o;
        o.x = pack_bf16x2(g[0], g[1]); o.y = pack_bf16x2(g[2], g[3]);
        o.z = pack_bf16x2(g[4], g[5]); o.w = pack_bf16x2(g[6], g[7]);
        *reinterpret_cast<uint4*>(ds + (size_t)r * ldds + k) = o;
      }
    }
  }
  dot = block_sum(dot, redf);
  if (threadIdx.x == 0) loss_rows[r] = w * (n_t * lse - s_scale * dot);
}

// ------------------------------------------------------------------------------------------------
// out[j] = sum_{i in segment j} x[i];  segments given by offsets[j]..offsets[j+1]. One CTA per segment,
// fixed summation order (deterministic).
__global__ void __launch_bounds__(256) segment_sum_kernel(const float* __restrict__ x, const int* __restrict__ offsets,
                                                          float* __restrict__ out, const float* __restrict__ scale) {
  B200_PDL_SYNC();
  __shared__ float red[32];
  const int j = blockIdx.x;
  const int a = offsets[j], b = offsets[j + 1];
  float acc = 0.f;
  for (int i = a + threadIdx.x; i < b; i += blockDim.x) acc += x[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[j] = acc * (scale ? scale[j] : 1.f);
}

// ------------------------------------------------------------------------------------------------
// KoLeo forward + backward for one group of n <= 256 rows of dimension D (single CTA per group).
//   xn = x / max(||x||, eps); sim = round_bf16?(xn xn^T), diag = -2; j(i) = argmax_k sim[i,k] (first index)
//   d_i = || xn_i - xn_j + eps ||_2 ; loss = -(1/n) sum_i log(d_i + eps)
// dx (fp32, accumulated with +=) = gscale * dloss/dx, argmax treated as constant (as autograd does).
template <int DL>
__global__ void __launch_bounds__(256) koleo_kernel(const float* __restrict__ x, long long ldx, int n, int D, float eps,
                                                    int bf16_sim, float gscale, float* __restrict__ loss_out,
                                                    float* __restrict__ dx, long long lddx, int* __restrict__ nn_out) {
  B200_PDL_SYNC();
  extern __shared__ float sm[];
  float* xn = sm;                        // [n, D]
  float* gxn = xn + (size_t)n * D;       // [n, D] grad wrt xn
  float* nrm = gxn + (size_t)n * D;      // [n]
  int* nn = reinterpret_cast<int*>(nrm + n);  // [n]
  float* dist = reinterpret_cast<float*>(nn + n);  // [n]
  const int g = blockIdx.x;
  x += (size_t)g * n * ldx;
  if (dx) dx += (size_t)g * n * lddx;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

  for (int i = warp; i < n; i += nwarps) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { float v = x[(size_t)i * ldx + d]; ss += v * v; }
    ss = warp_sum(ss);
    float nr = fmaxf(sqrtf(ss), eps);
    if (lane == 0) nrm[i] = nr;
    for (int d = lane; d < D; d += 32) { xn[i * D + d] = x[(size_t)i * ldx + d] / nr; gxn[i * D + d] = 0.f; }
  }
  __syncthreads();
  // optional bf16 rounding of the similarity operands (the autocast reference runs x @ x.T in bf16): done once, in place
  // in a copy that reuses gxn (zeroed again below)
  float* xs = gxn;
  for (int i = threadIdx.x; i < n * D; i += blockDim.x) xs[i] = bf16_sim ? bf16_round(xn[i]) : xn[i];
  __syncthreads();
  // nearest neighbour: a warp owns up to KR rows at a time (kept in registers) and streams every candidate row k once
  constexpr int KR = 4;     // DL = D / 32 elements of a row per lane (template parameter -> registers)
  for (int i0 = warp * KR; i0 < n; i0 += nwarps * KR) {
    float r[KR][DL];
#pragma unroll
    for (int q = 0; q < KR; ++q)
#pragma unroll
      for (int e = 0; e < DL; ++e) r[q][e] = (i0 + q < n) ? xs[(i0 + q) * D + lane + 32 * e] : 0.f;
    float best[KR]; int bi[KR];
#pragma unroll
    for (int q = 0; q < KR; ++q) { best[q] = -INFINITY; bi[q] = 0; }
    for (int k = 0; k < n; ++k) {
      float dot[KR];
#pragma unroll
      for (int q = 0; q < KR; ++q) dot[q] = 0.f;
#pragma unroll
      for (int e = 0; e < DL; ++e) {
        const float xk = xs[k * D + lane + 32 * e];
#pragma unroll
        for (int q = 0; q < KR; ++q) dot[q] = fmaf(r[q][e], xk, dot[q]);
      }
#pragma unroll
      for (int q = 0; q < KR; ++q) {
        float dsum = warp_sum(dot[q]);
        if (bf16_sim) dsum = bf16_round(dsum);
        if (k == i0 + q) dsum = -2.f;
        if (dsum > best[q]) { best[q] = dsum; bi[q] = k; }
      }
    }
#pragma unroll
    for (int q = 0; q < KR; ++q) {
      const int i = i0 + q;
      if (i >= n) continue;
      float ss = 0.f;
      for (int d = lane; d < D; d += 32) { float df = xn[i * D + d] - xn[bi[q] * D + d] + eps; ss += df * df; }
      ss = warp_sum(ss);
      if (lane == 0) { nn[i] = bi[q]; dist[i] = sqrtf(ss); if (nn_out) nn_out[g * n + i] = bi[q]; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n * D; i += blockDim.x) gxn[i] = 0.f;
  __syncthreads();
  if (warp == 0) {
    float l = 0.f;
    for (int i = lane; i < n; i += 32) l += -__logf(dist[i] + eps);
    l = warp_sum(l);
    if (lane == 0) loss_out[g] = l / n;
  }
  if (!dx) return;
  // grad wrt xn: i gets +c_i * diff, j(i) gets -c_i * diff, c_i = -(1/n)/(d_i+eps)/d_i ; serial over i per column
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    for (int i = 0; i < n; ++i) {
      const int j = nn[i];
      const float di = dist[i];
      const float c = -(1.f / n) / (di + eps) / fmaxf(di, 1e-30f);
      const float df = xn[i * D + d] - xn[j * D + d] + eps;
      gxn[i * D + d] += c * df;
      gxn[j * D + d] -= c * df;
    }
  }
  __syncthreads();
  // through the normalisation: dx = (g - xn * <xn, g>) / max(||x||, eps)
  for (int i = warp; i < n; i += nwarps) {
    float dot = 0.f;
    for (int d = lane; d < D; d += 32) dot += xn[i * D + d] * gxn[i * D + d];
    dot = warp_sum(dot);
    for (int d = lane; d < D; d += 32)
      dx[(size_t)i * lddx + d] += gscale * (gxn[i * D + d] - xn[i * D + d] * dot) / nrm[i];
  }
}

// ------------------------------------------------------------------------------------------------
// KoLeo for groups that do not fit one CTA's shared memory (n * D * 8 B > 220 KB: e.g. ViT-B at 64 images per GPU):
// the same arithmetic as koleo_kernel, tiled over rows, with the normalised rows / gradient accumulator / per-row
// results in a global scratch (a few hundred KB: L2-resident).  Three launches:
//   koleo_prep   : xn = x / max(||x||, eps) ; G = 0 ; loss = 0
//   koleo_search : 32 query rows per CTA against all n candidates -> nn, d, loss += -log(d + eps)/n, and the two
//                  gradient contributions of row i (to itself and to its neighbour) with fp32 atomics on G
//   koleo_apply  : dx += gscale * (G - xn <xn, G>) / ||x||
__global__ void __launch_bounds__(256) koleo_prep_kernel(const float* __restrict__ x, long long ldx, int rows, int n, int D, float eps,
                                                         float* __restrict__ xn, float* __restrict__ G, float* __restrict__ nrm,
                                                         float* __restrict__ loss_out) {
  B200_PDL_SYNC();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) { const float v = x[(size_t)r * ldx + d]; ss += v * v; }
  ss = warp_sum(ss);
  const float nr = fmaxf(sqrtf(ss), eps);
  for (int d = lane; d < D; d += 32) { xn[(size_t)r * D + d] = x[(size_t)r * ldx + d] / nr; G[(size_t)r * D + d] = 0.f; }
  if (lane == 0) {
    nrm[r] = nr;
    if (r % n == 0) loss_out[r / n] = 0.f;
  }
}

template <int DL>
__global__ void __launch_bounds__(256) koleo_search_kernel(const float* __restrict__ xn, int n, int D, float eps, int bf16_sim,
                                                           float* __restrict__ G, float* __restrict__ loss_out,
                                                           int* __restrict__ nn_out, bool want_grad) {
  B200_PDL_SYNC();
  constexpr int KR = 4;
  const int g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xg = xn + (size_t)g * n * D;
  float* Gg = G + (size_t)g * n * D;
  const int i0 = blockIdx.x * 32 + warp * KR;
  if (i0 >= n) return;
  float r[KR][DL];
#pragma unroll
  for (int q = 0; q < KR; ++q)
#pragma unroll
    for (int e = 0; e < DL; ++e) {
      const float v = (i0 + q < n) ? xg[(size_t)(i0 + q) * D + lane + 32 * e] : 0.f;
      r[q][e] = bf16_sim ? bf16_round(v) : v;
    }
  float best[KR]; int bi[KR];
#pragma unroll
  for (int q = 0; q < KR; ++q) { best[q] = -INFINITY; bi[q] = 0; }
  for (int k = 0; k < n; ++k) {
    float dot[KR];
#pragma unroll
    for (int q = 0; q < KR; ++q) dot[q] = 0.f;
#pragma unroll
    for (int e = 0; e < DL; ++e) {
      float xk = xg[(size_t)k * D + lane + 32 * e];
      if (bf16_sim) xk = bf16_round(xk);
#pragma unroll
      for (int q = 0; q < KR; ++q) dot[q] = fmaf(r[q][e], xk, dot[q]);
    }
#pragma unroll
    for (int q = 0; q < KR; ++q) {
      float dsum = warp_sum(dot[q]);
      if (bf16_sim) dsum = bf16_round(dsum);
      if (k == i0 + q) dsum = -2.f;
      if (dsum > best[q]) { best[q] = dsum; bi[q] = k; }
    }
  }
#pragma unroll
  for (int q = 0; q < KR; ++q) {
    const int i = i0 + q;
    if (i >= n) continue;
    const int j = bi[q];
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float df = xg[(size_t)i * D + d] - xg[(size_t)j * D + d] + eps; ss += df * df; }
    ss = warp_sum(ss);
    const float di = sqrtf(ss);
    if (lane == 0) {
      atomicAdd(loss_out + g, -__logf(di + eps) / n);
      if (nn_out) nn_out[g * n + i] = j;
    }
    if (want_grad) {
      const float c = -(1.f / n) / (di + eps) / fmaxf(di, 1e-30f);
      for (int d = lane; d < D; d += 32) {
        const float v = c * (xg[(size_t)i * D + d] - xg[(size_t)j * D + d] + eps);
        atomicAdd(Gg + (size_t)i * D + d, v);
        atomicAdd(Gg + (size_t)j * D + d, -v);
      }
    }
  }
}

__global__ void __launch_bounds__(256) koleo_apply_kernel(const float* __restrict__ xn, const float* __restrict__ G,
                                                          const float* __restrict__ nrm, int rows, int D, float gscale,
                                                          float* __restrict__ dx, long long lddx) {
  B200_PDL_SYNC();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) dot += xn[(size_t)r * D + d] * G[(size_t)r * D + d];
  dot = warp_sum(dot);
  const float inv = gscale / nrm[r];
  for (int d = lane; d < D; d += 32) dx[(size_t)r * lddx + d] += inv * (G[(size_t)r * D + d] - xn[(size_t)r * D + d] * dot);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_row_lse(const void* x, long long ld, int R, int K, const float* colterm, float scale,
                            const float* scale_dev, float* rowterm, void* stream) {
  if (!x || !rowterm || R <= 0 || K <= 0 || (K % 8) || (ld % 8)) return B200_ERR_INVALID_ARG;
  launch_kernel(row_lse_kernel, R, LOSS_THREADS, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, ld, R, K, colterm, scale, scale_dev, rowterm);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_col_reduce(const void* x, long long ld, int R, int K, const float* rowvec, float scale,
                               const float* scale_dev, int mode, float* out, void* stream) {
  if (!x || !out || R <= 0 || K <= 0 || (K % 8) || (ld % 8) || mode < 0 || mode > 1) return B200_ERR_INVALID_ARG;
  const int groups = K / 8;
  const int gx = (groups + 127) / 128;            // <= 128 column groups (1024 columns) per CTA, evenly split
  const int cg = (groups + gx - 1) / gx;
  int gy = (148 * 2 + gx - 1) / gx;                // ~2 CTAs per SM in total
  const int lanes = CR_THREADS / cg;
  const int min_rows = lanes * 8;
  if ((long long)gy * min_rows > R) gy = (R + min_rows - 1) / min_rows;
  if (gy < 1) gy = 1;
  const int rows_per_cta = (R + gy - 1) / gy;
  gy = (R + rows_per_cta - 1) / rows_per_cta;
  const size_t smem = (size_t)lanes * cg * 8 * sizeof(float);
  launch_kernel(col_reduce_kernel, dim3(gx, gy), CR_THREADS, smem, (cudaStream_t)stream, (const __nv_bfloat16*)x, ld, R, K, cg, rows_per_cta,
                                                                                rowvec, scale, scale_dev, mode, out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_vec_op(float* y, const float* x, int n, float a, float b, int op, const float* a_dev, void* stream) {
  if (!y || !x || n <= 0 || op < 0 || op > 2) return B200_ERR_INVALID_ARG;
  launch_kernel(vec_op_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, y, x, n, a, b, op, a_dev);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_dino_ce(const void* s, long long lds, int Rs, int K, const void* t, long long ldt,
                            const float* colterm, const float* t_rowterm, const int* t_idx0, const int* t_idx1,
                            const float* weight, float s_scale, float t_scale, const float* t_scale_dev, float gscale,
                            float* loss_rows, void* ds, long long ldds, void* stream) {
  if (!s || !t || !t_idx0 || !loss_rows || Rs <= 0 || K <= 0) return B200_ERR_INVALID_ARG;
  if (!t_rowterm && t_idx1) return B200_ERR_INVALID_ARG;  // the fused teacher LSE covers single-teacher rows only
  if ((K % 8) || (lds % 8) || (ldt % 8) || (ds && (ldds % 8))) return B200_ERR_INVALID_ARG;
  // 512 threads per row, one 16-byte load per tensor in flight: the fastest of the measured launch shapes (256 threads x 4 rows
  // per SM, two loads in flight, a persistent variant staging the student row in shared memory: profiles/r02_loss_bench*.log)
  const __nv_bfloat16* sp = (const __nv_bfloat16*)s;
  const __nv_bfloat16* tp = (const __nv_bfloat16*)t;
  __nv_bfloat16* dsp = (__nv_bfloat16*)ds;
  cudaStream_t st = (cudaStream_t)stream;
  if (t_rowterm)
    launch_kernel(dino_ce_kernel<false, 512, false>, Rs, 512, 0, st, sp, lds, Rs, K, tp, ldt, colterm, t_rowterm, t_idx0, t_idx1, weight,
                  s_scale, t_scale, t_scale_dev, gscale, loss_rows, dsp, ldds);
  else
    launch_kernel(dino_ce_kernel<true, 512, false>, Rs, 512, 0, st, sp, lds, Rs, K, tp, ldt, colterm, (const float*)nullptr, t_idx0,
                  (const int*)nullptr, weight, s_scale, t_scale, t_scale_dev, gscale, loss_rows, dsp, ldds);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_segment_sum(const float* x, const int* offsets, int n_segments, const float* scale, float* out,
                                void* stream) {
  if (!x || !offsets || !out || n_segments <= 0) return B200_ERR_INVALID_ARG;
  launch_kernel(segment_sum_kernel, n_segments, 256, 0, (cudaStream_t)stream, x, offsets, out, scale);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_koleo(const float* x, long long ldx, int groups, int n, int D, float eps, int bf16_sim, float gscale,
                          float* loss_out, float* dx, long long lddx, int* nn_out, float* scratch, long long scratch_elems,
                          void* stream) {
  if (!x || !loss_out || groups <= 0 || n <= 1 || D <= 0 || (D % 32) || D > 1024) return B200_ERR_INVALID_ARG;
  const size_t smem = (size_t)(2 * n * D + 3 * n) * sizeof(float);
  const bool tiled = smem > 220 * 1024 || n > 256;
  if (tiled) {
    // rows tiled over CTAs, state in the caller's scratch: xn [R, D] | G [R, D] | nrm [R]   (R = groups * n)
    const long long R = (long long)groups * n;
    if (!scratch || scratch_elems < 2 * R * D + R) return B200_ERR_INVALID_ARG;
    float* xn = scratch;
    float* G = xn + R * D;
    float* nrm = G + R * D;
    cudaStream_t st = (cudaStream_t)stream;
    launch_kernel(koleo_prep_kernel, (int)((R + 7) / 8), 256, 0, st, x, ldx, (int)R, n, D, eps, xn, G, nrm, loss_out);
#define B200_KOLEO_T(DLV) \
    launch_kernel(koleo_search_kernel<DLV>, dim3((n + 31) / 32, groups), 256, 0, st, (const float*)xn, n, D, eps, bf16_sim, G, loss_out, nn_out, dx != nullptr)
    switch (D / 32) {
      case 4: B200_KOLEO_T(4); break;
      case 6: B200_KOLEO_T(6); break;
      case 12: B200_KOLEO_T(12); break;
      case 24: B200_KOLEO_T(24); break;
      case 32: B200_KOLEO_T(32); break;
      default: return B200_ERR_UNSUPPORTED;
    }
#undef B200_KOLEO_T
    if (dx) launch_kernel(koleo_apply_kernel, (int)((R + 7) / 8), 256, 0, st, (const float*)xn, (const float*)G, (const float*)nrm, (int)R, D, gscale, dx, lddx);
    B200_CHECK_LAUNCH();
    return B200_OK;
  }
#define B200_KOLEO(DLV)                                                                                              \
  do {                                                                                                               \
    static bool attr = false;                                                                                        \
    if (!attr) {                                                                                                     \
      if (cudaFuncSetAttribute(koleo_kernel<DLV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) \
        return B200_ERR_CUDA;                                                                                        \
      attr = true;                                                                                                   \
    }                                                                                                                \
    launch_kernel(koleo_kernel<DLV>, groups, 256, smem, (cudaStream_t)stream, x, ldx, n, D, eps, bf16_sim, gscale, loss_out, dx, lddx, nn_out); \
  } while (0)
  switch (D / 32) {
    case 4: B200_KOLEO(4); break;
    case 6: B200_KOLEO(6); break;
    case 12: B200_KOLEO(12); break;
    case 24: B200_KOLEO(24); break;
    case 32: B200_KOLEO(32); break;
    default: return B200_ERR_UNSUPPORTED;
  }
#undef B200_KOLEO
  B200_CHECK_LAUNCH();
  return B200_OK;
}
