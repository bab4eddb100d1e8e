// Row-wise / gather / layout kernels around the GEMMs (all HBM-bound, one warp per token row, 128-bit access).
//
// Reference arithmetic replaced (LT = src/lightly_train):
//   nn.LayerNorm(eps=1e-6) fwd/bwd            LT/_models/dinov2_vit/dinov2_vit_src/models/vision_transformer.py:138,377
//                                             LT/_models/dinov2_vit/dinov2_vit_src/layers/block.py:60,74,92,95
//   PatchEmbed im2col + token assembly        layers/patch_embed.py:92-113, vision_transformer.py:307-329
//   LayerScale backward                       layers/layer_scale.py:27-28
//   torch.index_select of masked tokens       LT/_methods/dinov2/dinov2.py:427-431,496-500
//   F.normalize + weight_norm                 LT/_methods/dinov2/dinov2_head.py:54-58,66-71
#include "common.cuh"
#include "../../include/b200dino.h"

namespace b200 {

static constexpr int ROW_THREADS = 256;  // 8 warps = 8 rows per CTA
static constexpr int MAXV = 8;           // float4 chunks per lane -> D <= 1024

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: x f32 [T,D] -> y (bf16 or f32) ; saves mean, rstd.
// Per-lane column partials (VMAX float4 per lane, column c = lane + 32 j) -> one [D] vector in shared memory.  The warps
// of the CTA take turns (plain read-modify-write, a __syncthreads between turns): shared-memory atomicAdd(float) is a
// compare-and-swap loop on this architecture, and 8 warps hammering the same addresses made the tails of the LayerNorm
// backward kernels cost more instructions than their row loops (ncu: 63 % of ln_bwd_ls's executed instructions).
template <int VMAX>
__device__ __forceinline__ void add_cols(float* sm_vec, const float4 (&a)[VMAX], int nv, int lane) {
#pragma unroll
  for (int j = 0; j < VMAX; ++j) {
    const int c = lane + 32 * j;
    if (c < nv) {
      float4* p = reinterpret_cast<float4*>(sm_vec) + c;
      float4 t = *p;
      t.x += a[j].x; t.y += a[j].y; t.z += a[j].z; t.w += a[j].w;
      *p = t;
    }
  }
}

template <bool OUT_BF16, int VMAX>
__global__ void __launch_bounds__(ROW_THREADS) ln_fwd_kernel(const float* __restrict__ x, long long ldx, int T, int D,
                                                             const float* __restrict__ w, const float* __restrict__ b,
                                                             float eps, void* __restrict__ y, long long ldy,
                                                             float* __restrict__ mean, float* __restrict__ rstd) {
  B200_PDL_SYNC();
  const int row = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= T) return;
  const int lane = threadIdx.x & 31;
  const int nv = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
  float4 v[VMAX];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VMAX; ++j) {
    int c = lane + 32 * j;
    if (c < nv) { v[j] = xr[c]; s += v[j].x + v[j].y + v[j].z + v[j].w; }
  }
  const float mu = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VMAX; ++j) {
    int c = lane + 32 * j;
    if (c < nv) {
      float a = v[j].x - mu, bb = v[j].y - mu, cc = v[j].z - mu, d = v[j].w - mu;
      q += a * a + bb * bb + cc * cc + d * d;
    }
  }
  const float rs = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) { if (mean) mean[row] = mu; if (rstd) rstd[row] = rs; }
#pragma unroll
  for (int j = 0; j < VMAX; ++j) {
    int c = lane + 32 * j;
    if (c < nv) {
      float4 ww = __ldg(reinterpret_cast<const float4*>(w) + c);
      float4 bb = __ldg(reinterpret_cast<const float4*>(b) + c);
      float4 o;
      o.x = (v[j].x - mu) * rs * ww.x + bb.x; o.y = (v[j].y - mu) * rs * ww.y + bb.y;
      o.z = (v[j].z - mu) * rs * ww.z + bb.z; o.w = (v[j].w - mu) * rs * ww.w + bb.w;
      if (OUT_BF16) {
        uint2 p; p.x = pack_bf16x2(o.x, o.y); p.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + (size_t)row * ldy)[c] = p;
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (size_t)row * ldy)[c] = o;
      }
    }
  }
}

// LayerNorm backward.  dy (bf16 or f32) = grad wrt LN output.  dx: f32, accumulate (+=) or assign.
// dw/db: per-CTA partials reduced through smem, then atomicAdd into global [D].
template <bool DY_BF16, int VMAX>
__global__ void __launch_bounds__(ROW_THREADS) ln_bwd_kernel(const void* __restrict__ dy, long long lddy,
                                                             const float* __restrict__ x, long long ldx, int T, int D,
                                                             const float* __restrict__ w, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, float* __restrict__ dx,
                                                             long long lddx, int accumulate, float* __restrict__ dw,
                                                             float* __restrict__ db) {
  B200_PDL_SYNC();
  extern __shared__ float sm[];  // [2][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = D >> 2;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  float4 aw[VMAX], ab[VMAX];
#pragma unroll
  for (int j = 0; j < VMAX; ++j) { aw[j] = make_float4(0, 0, 0, 0); ab[j] = make_float4(0, 0, 0, 0); }
  for (int row = blockIdx.x * (ROW_THREADS / 32) + warp; row < T; row += gridDim.x * (ROW_THREADS / 32)) {
    const float mu = mean[row], rs = rstd[row];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
    float4 g[VMAX], xh[VMAX];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VMAX; ++j) {
      int c = lane + 32 * j;
      if (c < nv) {
        float4 d;
        if (DY_BF16) {
          uint2 p = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)row * lddy)[c];
          float2 a = unpack_bf16x2(p.x), b2 = unpack_bf16x2(p.y);
          d = make_float4(a.x, a.y, b2.x, b2.y);
        } else {
          d = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + (size_t)row * lddy)[c];
        }
        float4 xv = xr[c];
        float4 ww = __ldg(reinterpret_cast<const float4*>(w) + c);
        xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[j] = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
        s1 += g[j].x + g[j].y + g[j].z + g[j].w;
        s2 += g[j].x * xh[j].x + g[j].y * xh[j].y + g[j].z * xh[j].z + g[j].w * xh[j].w;
        aw[j].x += d.x * xh[j].x; aw[j].y += d.y * xh[j].y; aw[j].z += d.z * xh[j].z; aw[j].w += d.w * xh[j].w;
        ab[j].x += d.x; ab[j].y += d.y; ab[j].z += d.z; ab[j].w += d.w;
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    float4* dxr = reinterpret_cast<float4*>(dx + (size_t)row * lddx);
#pragma unroll
    for (int j = 0; j < VMAX; ++j) {
      int c = lane + 32 * j;
      if (c < nv) {
        float4 o;
        o.x = rs * (g[j].x - s1 - xh[j].x * s2); o.y = rs * (g[j].y - s1 - xh[j].y * s2);
        o.z = rs * (g[j].z - s1 - xh[j].z * s2); o.w = rs * (g[j].w - s1 - xh[j].w * s2);
        if (accumulate) { float4 p = dxr[c]; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
        dxr[c] = o;
      }
    }
  }
  if (dw) {
    for (int w = 0; w < ROW_THREADS / 32; ++w) {
      if (warp == w) { add_cols<VMAX>(sm, aw, nv, lane); add_cols<VMAX>(sm + D, ab, nv, lane); }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) { atomicAdd(dw + i, sm[i]); atomicAdd(db + i, sm[D + i]); }
  }
}


// ------------------------------------------------------------------------------------------------
// LayerNorm backward fused with the LayerScale backward that always follows it in the block schedule:
//   dx_new = dx_old + LN'(dy)                                   (residual-stream gradient, fp32, in place)
//   dout   = bf16(dx_new * rowscale * gamma)                    (gradient of the NEXT branch output, toward its GEMM)
//   dgamma += sum dx_new*rowscale*o ; dbias += sum dout ; dw_ln/db_ln += LayerNorm parameter gradients
// One read of dx less and one launch less than ln_bwd + layerscale_bwd (18 instead of 22 bytes per element).
// EXACT: D == 128 * VMAX, i.e. every lane owns VMAX full float4 columns (no per-column bounds tests in the row loop).
template <bool DY_BF16, int VMAX, bool EXACT>
__global__ void __launch_bounds__(ROW_THREADS) ln_bwd_ls_kernel(
    const void* __restrict__ dy, long long lddy, const float* __restrict__ x, long long ldx, int T, int D,
    const float* __restrict__ w, const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx,
    long long lddx, int accumulate, float* __restrict__ dw, float* __restrict__ db, const __nv_bfloat16* __restrict__ o,
    long long ldo, const float* __restrict__ gamma, const float* __restrict__ rowscale, int rows_per_scale,
    __nv_bfloat16* __restrict__ dout, long long lddo, float* __restrict__ dgamma, float* __restrict__ dbias) {
  B200_PDL_SYNC();
  extern __shared__ float sm[];  // [4][D]: dw, db, dgamma, dbias
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = D >> 2;
  for (int i = threadIdx.x; i < 4 * D; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  float4 aw[VMAX], ab[VMAX], ag[VMAX], ao[VMAX];
#pragma unroll
  for (int j = 0; j < VMAX; ++j) { aw[j] = ab[j] = ag[j] = ao[j] = make_float4(0, 0, 0, 0); }
  for (int row = blockIdx.x * (ROW_THREADS / 32) + warp; row < T; row += gridDim.x * (ROW_THREADS / 32)) {
    {  // this warp's next row -> L2 (each row is read exactly once; the loads below are the only ones in flight)
      const int nrow = row + gridDim.x * (ROW_THREADS / 32);
      if (nrow < T) {
        const int b4 = lane * 128, b2 = lane * 128;
        if (b4 < D * 4) {
          prefetch_l2(reinterpret_cast<const char*>(x + (size_t)nrow * ldx) + b4);
          if (accumulate) prefetch_l2(reinterpret_cast<const char*>(dx + (size_t)nrow * lddx) + b4);
        }
        if (b2 < D * 2) {
          prefetch_l2(reinterpret_cast<const char*>(o + (size_t)nrow * ldo) + b2);
          if (DY_BF16) prefetch_l2(reinterpret_cast<const char*>(reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)nrow * lddy) + b2);
        }
      }
    }
    const float mu = mean[row], rs = rstd[row];
    const float dps = rowscale ? __ldg(rowscale + row / rows_per_scale) : 1.f;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
    float4* dxr = reinterpret_cast<float4*>(dx + (size_t)row * lddx);
    float4 g[VMAX], xh[VMAX], dold[VMAX];
    uint2 ob[VMAX];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VMAX; ++j) {
      int c = lane + 32 * j;
      if (EXACT || c < nv) {
        float4 d;
        if (DY_BF16) {
          uint2 p = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)row * lddy)[c];
          float2 a = unpack_bf16x2(p.x), b2 = unpack_bf16x2(p.y);
          d = make_float4(a.x, a.y, b2.x, b2.y);
        } else {
          d = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + (size_t)row * lddy)[c];
        }
        float4 xv = xr[c];
        dold[j] = accumulate ? dxr[c] : make_float4(0, 0, 0, 0);
        ob[j] = reinterpret_cast<const uint2*>(o + (size_t)row * ldo)[c];
        float4 ww = __ldg(reinterpret_cast<const float4*>(w) + c);
        xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[j] = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
        s1 += g[j].x + g[j].y + g[j].z + g[j].w;
        s2 += g[j].x * xh[j].x + g[j].y * xh[j].y + g[j].z * xh[j].z + g[j].w * xh[j].w;
        aw[j].x += d.x * xh[j].x; aw[j].y += d.y * xh[j].y; aw[j].z += d.z * xh[j].z; aw[j].w += d.w * xh[j].w;
        ab[j].x += d.x; ab[j].y += d.y; ab[j].z += d.z; ab[j].w += d.w;
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int j = 0; j < VMAX; ++j) {
      int c = lane + 32 * j;
      if (EXACT || c < nv) {
        float4 nx;
        nx.x = dold[j].x + rs * (g[j].x - s1 - xh[j].x * s2); nx.y = dold[j].y + rs * (g[j].y - s1 - xh[j].y * s2);
        nx.z = dold[j].z + rs * (g[j].z - s1 - xh[j].z * s2); nx.w = dold[j].w + rs * (g[j].w - s1 - xh[j].w * s2);
        dxr[c] = nx;
        const float4 gs = make_float4(nx.x * dps, nx.y * dps, nx.z * dps, nx.w * dps);
        const float4 gm = gamma ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(1, 1, 1, 1);
        const float2 o0 = unpack_bf16x2(ob[j].x), o1 = unpack_bf16x2(ob[j].y);
        ag[j].x += gs.x * o0.x; ag[j].y += gs.y * o0.y; ag[j].z += gs.z * o1.x; ag[j].w += gs.w * o1.y;
        uint2 pk; pk.x = pack_bf16x2(gs.x * gm.x, gs.y * gm.y); pk.y = pack_bf16x2(gs.z * gm.z, gs.w * gm.w);
        const float2 d01 = unpack_bf16x2(pk.x), d23 = unpack_bf16x2(pk.y);  // the bf16-rounded values that are stored
        ao[j].x += d01.x; ao[j].y += d01.y; ao[j].z += d23.x; ao[j].w += d23.y;
        reinterpret_cast<uint2*>(dout + (size_t)row * lddo)[c] = pk;
      }
    }
  }
  for (int w = 0; w < ROW_THREADS / 32; ++w) {
    if (warp == w) {
      add_cols<VMAX>(sm, aw, nv, lane); add_cols<VMAX>(sm + D, ab, nv, lane);
      add_cols<VMAX>(sm + 2 * D, ag, nv, lane); add_cols<VMAX>(sm + 3 * D, ao, nv, lane);
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dw) { atomicAdd(dw + i, sm[i]); atomicAdd(db + i, sm[D + i]); }
    if (dgamma) atomicAdd(dgamma + i, sm[2 * D + i]);
    if (dbias) atomicAdd(dbias + i, sm[3 * D + i]);
  }
}

// ------------------------------------------------------------------------------------------------
// im2col for Conv2d(k=s=p): x f32 [B,C,H,W] -> cols bf16 [B*Np, C*p*p], column order (c, ky, kx).
__global__ void im2col_kernel(const float* __restrict__ x, int B, int C, int H, int W, int p,
                              __nv_bfloat16* __restrict__ cols, long long ldc) {
  B200_PDL_SYNC();
  const int gw = W / p, gh = H / p;
  const long long total = (long long)B * gh * gw * C * p;  // one work item = p contiguous pixels
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int ky = (int)(i % p);
    long long r = i / p;
    int c = (int)(r % C); r /= C;
    int px = (int)(r % gw); r /= gw;
    int py = (int)(r % gh);
    int b = (int)(r / gh);
    const float* src = x + (((size_t)b * C + c) * H + (size_t)py * p + ky) * W + (size_t)px * p;
    __nv_bfloat16* dst = cols + ((size_t)(b * gh + py) * gw + px) * ldc + ((size_t)c * p + ky) * p;
    if ((p & 3) == 0 && (W & 3) == 0) {
      for (int k = 0; k < p; k += 4) {
        float4 v = *reinterpret_cast<const float4*>(src + k);
        uint2 o; o.x = pack_bf16x2(v.x, v.y); o.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + k) = o;
      }
    } else {
      for (int k = 0; k < p; ++k) dst[k] = __float2bfloat16_rn(src[k]);
    }
  }
}

// Token assembly (prepare_tokens_with_masks): tok bf16 [B*Np, D] -> x f32 [B, N=1+R+Np, D]
//   x[b,0]       = cls + pos[0]
//   x[b,1..R]    = reg[r]                      (register tokens get no positional embedding)
//   x[b,1+R+p]   = (mask ? bf16(mask_token) : tok) + pos[1+p]
__global__ void __launch_bounds__(ROW_THREADS) assemble_tokens_kernel(const __nv_bfloat16* __restrict__ tok, long long ldt,
                                                                      const unsigned char* __restrict__ masks,
                                                                      const float* __restrict__ mask_token,
                                                                      const float* __restrict__ cls,
                                                                      const float* __restrict__ reg,
                                                                      const float* __restrict__ pos, int B, int Np, int R,
                                                                      int D, float* __restrict__ x) {
  B200_PDL_SYNC();
  const int N = 1 + R + Np;
  const long long row = (long long)blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= (long long)B * N) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(row / N), n = (int)(row % N);
  float* xr = x + (size_t)row * D;
  for (int d = lane * 4; d < D; d += 128) {
    float4 o;
    if (n == 0) {
      float4 c = *reinterpret_cast<const float4*>(cls + d), pp = *reinterpret_cast<const float4*>(pos + d);
      o = make_float4(c.x + pp.x, c.y + pp.y, c.z + pp.z, c.w + pp.w);
    } else if (n <= R) {
      o = *reinterpret_cast<const float4*>(reg + (size_t)(n - 1) * D + d);
    } else {
      const int p = n - 1 - R;
      float4 t;
      if (masks && masks[(size_t)b * Np + p]) {
        float4 m = *reinterpret_cast<const float4*>(mask_token + d);
        t = make_float4(bf16_round(m.x), bf16_round(m.y), bf16_round(m.z), bf16_round(m.w));
      } else {
        uint2 u = *reinterpret_cast<const uint2*>(tok + ((size_t)b * Np + p) * ldt + d);
        float2 a = unpack_bf16x2(u.x), c2 = unpack_bf16x2(u.y);
        t = make_float4(a.x, a.y, c2.x, c2.y);
      }
      float4 pp = *reinterpret_cast<const float4*>(pos + (size_t)(1 + p) * D + d);
      o = make_float4(t.x + pp.x, t.y + pp.y, t.z + pp.z, t.w + pp.w);
    }
    *reinterpret_cast<float4*>(xr + d) = o;
  }
}

// Backward of token assembly. dx f32 [B,N,D]:
//   dtok bf16 [B*Np, D] = masked ? 0 : dx[b,1+R+p]
//   dpos[n'] += sum_b dx (cls & patch rows), dcls += sum_b dx[b,0], dreg[r] += sum_b dx[b,1+r],
//   dmask_token += sum over masked tokens of dx.      One thread per (n, 4 columns), loop over b.
__global__ void assemble_tokens_bwd_kernel(const float* __restrict__ dx, const unsigned char* __restrict__ masks, int B,
                                           int Np, int R, int D, __nv_bfloat16* __restrict__ dtok, long long lddt,
                                           float* __restrict__ dpos, float* __restrict__ dcls, float* __restrict__ dreg,
                                           float* __restrict__ dmask_token) {
  B200_PDL_SYNC();
  const int N = 1 + R + Np;
  const int nd4 = D >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * nd4) return;
  const int n = (int)(i / nd4), d = (int)(i % nd4) * 4;
  float4 acc = make_float4(0, 0, 0, 0), accm = make_float4(0, 0, 0, 0);
  const int per = (B + gridDim.y - 1) / gridDim.y;  // the batch is split over blockIdx.y (partial sums meet in atomics)
  const int b0 = blockIdx.y * per, b1 = min(B, b0 + per);
  for (int b = b0; b < b1; ++b) {
    float4 g = *reinterpret_cast<const float4*>(dx + ((size_t)b * N + n) * D + d);
    if (n > R) {
      const int p = n - 1 - R;
      const bool m = masks && masks[(size_t)b * Np + p];
      if (m) { accm.x += g.x; accm.y += g.y; accm.z += g.z; accm.w += g.w; }
      uint2 o;
      o.x = m ? 0u : pack_bf16x2(g.x, g.y);
      o.y = m ? 0u : pack_bf16x2(g.z, g.w);
      *reinterpret_cast<uint2*>(dtok + ((size_t)b * Np + p) * lddt + d) = o;
    }
    acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
  }
  if (n == 0) {
    atomicAdd(dcls + d, acc.x); atomicAdd(dcls + d + 1, acc.y); atomicAdd(dcls + d + 2, acc.z); atomicAdd(dcls + d + 3, acc.w);
    atomicAdd(dpos + d, acc.x); atomicAdd(dpos + d + 1, acc.y); atomicAdd(dpos + d + 2, acc.z); atomicAdd(dpos + d + 3, acc.w);
  } else if (n <= R) {
    float* q = dreg + (size_t)(n - 1) * D + d;
    atomicAdd(q, acc.x); atomicAdd(q + 1, acc.y); atomicAdd(q + 2, acc.z); atomicAdd(q + 3, acc.w);
  } else {
    float* q = dpos + (size_t)(n - R) * D + d;
    atomicAdd(q, acc.x); atomicAdd(q + 1, acc.y); atomicAdd(q + 2, acc.z); atomicAdd(q + 3, acc.w);
    if (masks) {
      atomicAdd(dmask_token + d, accm.x); atomicAdd(dmask_token + d + 1, accm.y);
      atomicAdd(dmask_token + d + 2, accm.z); atomicAdd(dmask_token + d + 3, accm.w);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerScale backward: o bf16 [T,D] (branch output saved by the forward), dx f32 [T,D] (grad of the residual
// stream), gamma [D], rowscale (per-sample DropPath scale) ->
//   do bf16 = bf16(dx * rowscale * gamma) ; dgamma[d] += sum_t dx*rowscale*o ; dbias[d] += sum_t do
template <int VMAX>
__global__ void __launch_bounds__(ROW_THREADS) layerscale_bwd_kernel(const float* __restrict__ dx, long long lddx,
                                                                     const __nv_bfloat16* __restrict__ o, long long ldo,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ rowscale, int rows_per_scale,
                                                                     int T, int D, __nv_bfloat16* __restrict__ dout,
                                                                     long long lddo, float* __restrict__ dgamma,
                                                                     float* __restrict__ dbias) {
  B200_PDL_SYNC();
  extern __shared__ float sm[];  // [2][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = D >> 2;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  float4 ag[VMAX], ab[VMAX];
#pragma unroll
  for (int j = 0; j < VMAX; ++j) { ag[j] = make_float4(0, 0, 0, 0); ab[j] = make_float4(0, 0, 0, 0); }
  for (int row = blockIdx.x * (ROW_THREADS / 32) + warp; row < T; row += gridDim.x * (ROW_THREADS / 32)) {
    const float rs = rowscale ? __ldg(rowscale + row / rows_per_scale) : 1.f;
#pragma unroll
    for (int j = 0; j < VMAX; ++j) {
      int c = lane + 32 * j;
      if (c < nv) {
        float4 g = reinterpret_cast<const float4*>(dx + (size_t)row * lddx)[c];
        g.x *= rs; g.y *= rs; g.z *= rs; g.w *= rs;
        float4 gm = gamma ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(1, 1, 1, 1);
        uint2 ob = reinterpret_cast<const uint2*>(o + (size_t)row * ldo)[c];
        float2 o0 = unpack_bf16x2(ob.x), o1 = unpack_bf16x2(ob.y);
        ag[j].x += g.x * o0.x; ag[j].y += g.y * o0.y; ag[j].z += g.z * o1.x; ag[j].w += g.w * o1.y;
        float4 d = make_float4(bf16_round(g.x * gm.x), bf16_round(g.y * gm.y), bf16_round(g.z * gm.z), bf16_round(g.w * gm.w));
        ab[j].x += d.x; ab[j].y += d.y; ab[j].z += d.z; ab[j].w += d.w;
        uint2 p; p.x = pack_bf16x2(d.x, d.y); p.y = pack_bf16x2(d.z, d.w);
        reinterpret_cast<uint2*>(dout + (size_t)row * lddo)[c] = p;
      }
    }
  }
  for (int w = 0; w < ROW_THREADS / 32; ++w) {
    if (warp == w) { add_cols<VMAX>(sm, ag, nv, lane); add_cols<VMAX>(sm + D, ab, nv, lane); }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, sm[i]);
    if (dbias) atomicAdd(dbias + i, sm[D + i]);
  }
}

// ------------------------------------------------------------------------------------------------
// SwiGLU gate (layers/swiglu_ffn.py:31-35): 8 bf16 columns per thread, x1 = x12[:, :H], x2 = x12[:, H:].
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_ftz(1.0f + ex2_ftz(x * -1.4426950408889634f)); }

__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ x12, long long ld12, int T, int H,
                                  __nv_bfloat16* __restrict__ hid, long long ldh) {
  B200_PDL_SYNC();
  const int hv = H >> 3;
  const long long total = (long long)T * hv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / hv;
    const int c = (int)(i % hv) * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(x12 + row * ld12 + c);
    const uint4 b = *reinterpret_cast<const uint4*>(x12 + row * ld12 + H + c);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
    uint32_t ov[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x1 = unpack_bf16x2(av[k]), x2 = unpack_bf16x2(bv[k]);
      const float2 s = unpack_bf16x2(pack_bf16x2(x1.x * sigmoid_fast(x1.x), x1.y * sigmoid_fast(x1.y)));  // bf16 silu
      ov[k] = pack_bf16x2(s.x * x2.x, s.y * x2.y);
    }
    *reinterpret_cast<uint4*>(hid + row * ldh + c) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
  }
}

__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ x12, long long ld12, const __nv_bfloat16* __restrict__ dh,
                                  long long lddh, int T, int H, __nv_bfloat16* __restrict__ dx12, long long lddx) {
  B200_PDL_SYNC();
  const int hv = H >> 3;
  const long long total = (long long)T * hv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / hv;
    const int c = (int)(i % hv) * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(x12 + row * ld12 + c);
    const uint4 b = *reinterpret_cast<const uint4*>(x12 + row * ld12 + H + c);
    const uint4 g = *reinterpret_cast<const uint4*>(dh + row * lddh + c);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, gv[4] = {g.x, g.y, g.z, g.w};
    uint32_t o1[4], o2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x1 = unpack_bf16x2(av[k]), x2 = unpack_bf16x2(bv[k]), d = unpack_bf16x2(gv[k]);
      const float sx = sigmoid_fast(x1.x), sy = sigmoid_fast(x1.y);
      const float2 s = unpack_bf16x2(pack_bf16x2(x1.x * sx, x1.y * sy));             // bf16 silu(x1), as saved by autograd
      const float2 da = unpack_bf16x2(pack_bf16x2(d.x * x2.x, d.y * x2.y));          // grad wrt silu output (bf16)
      o1[k] = pack_bf16x2(da.x * (sx * (1.0f + x1.x * (1.0f - sx))), da.y * (sy * (1.0f + x1.y * (1.0f - sy))));
      o2[k] = pack_bf16x2(d.x * s.x, d.y * s.y);
    }
    *reinterpret_cast<uint4*>(dx12 + row * lddx + c) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    *reinterpret_cast<uint4*>(dx12 + row * lddx + H + c) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// Row gather / scatter of f32 rows.  src row index = map(idx[m]) with map(i) = (i / Np) * N + off + (i % Np)
// (token index inside [B, N, D] from an index into the flattened patch grid; Np = 0 -> identity map).
//   gather : out[m] = src[map(idx[m])]  (out f32 or bf16)
//   scatter: dst[map(idx[m])] = in[m]   (f32 <- f32|bf16; indices unique, no atomics)
template <bool OUT_BF16>
__global__ void __launch_bounds__(ROW_THREADS) gather_rows_kernel(const float* __restrict__ src, long long lds,
                                                                  const long long* __restrict__ idx, int M, int D, int Np,
                                                                  int N, int off, void* __restrict__ out, long long ldo) {
  B200_PDL_SYNC();
  const int m = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  long long i = idx[m];
  long long row = Np > 0 ? (i / Np) * N + off + (i % Np) : i;
  for (int d = lane * 4; d < D; d += 128) {
    float4 v = *reinterpret_cast<const float4*>(src + (size_t)row * lds + d);
    if (OUT_BF16) {
      uint2 p; p.x = pack_bf16x2(v.x, v.y); p.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)m * ldo + d) = p;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)m * ldo + d) = v;
    }
  }
}
template <bool IN_BF16>
__global__ void __launch_bounds__(ROW_THREADS) scatter_rows_kernel(const void* __restrict__ in, long long ldi,
                                                                   const long long* __restrict__ idx, int M, int D, int Np,
                                                                   int N, int off, float* __restrict__ dst, long long ldd,
                                                                   int accumulate, const int* __restrict__ count_dev) {
  B200_PDL_SYNC();
  const int m = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (m >= M) return;
  if (count_dev && m >= __ldg(count_dev)) return;  // padded rows of a static-shape (CUDA graph) step
  const int lane = threadIdx.x & 31;
  long long i = idx ? idx[m] : m;
  long long row = Np > 0 ? (i / Np) * N + off + (i % Np) : i;
  for (int d = lane * 4; d < D; d += 128) {
    float4 v;
    if (IN_BF16) {
      uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(in) + (size_t)m * ldi + d);
      float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
      v = make_float4(a.x, a.y, b.x, b.y);
    } else {
      v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + (size_t)m * ldi + d);
    }
    float4* q = reinterpret_cast<float4*>(dst + (size_t)row * ldd + d);
    if (accumulate) { float4 p = *q; v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
    *q = v;
  }
}

// ------------------------------------------------------------------------------------------------
// F.normalize(p=2, eps) forward: x bf16 [R,D] -> y bf16 = bf16(x / max(||x||, eps)); saves the fp32 norm.
// (reference: normalize output is fp32 and is cast to bf16 by the following autocast linear)
__global__ void __launch_bounds__(ROW_THREADS) l2norm_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int R,
                                                                 int D, float eps, __nv_bfloat16* __restrict__ y,
                                                                 long long ldy, float* __restrict__ nrm) {
  B200_PDL_SYNC();
  const int row = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  float ss = 0.f;
  for (int d = lane * 2; d < D; d += 64) {
    float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + (size_t)row * ldx + d));
    ss += v.x * v.x + v.y * v.y;
  }
  const float n = fmaxf(sqrtf(warp_sum(ss)), eps);
  if (lane == 0) nrm[row] = n;
  for (int d = lane * 2; d < D; d += 64) {
    float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + (size_t)row * ldx + d));
    *reinterpret_cast<uint32_t*>(y + (size_t)row * ldy + d) = pack_bf16x2(v.x / n, v.y / n);
  }
}
// backward: dy bf16 (grad wrt normalized output), x bf16, nrm -> dx bf16 = (dy - xn*<xn,dy>)/nrm
__global__ void __launch_bounds__(ROW_THREADS) l2norm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                                 const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                 const float* __restrict__ nrm, int R, int D,
                                                                 __nv_bfloat16* __restrict__ dx, long long lddx) {
  B200_PDL_SYNC();
  const int row = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  const float n = nrm[row];
  float dot = 0.f;
  for (int d = lane * 2; d < D; d += 64) {
    float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + (size_t)row * ldx + d));
    float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dy + (size_t)row * lddy + d));
    dot += (v.x / n) * g.x + (v.y / n) * g.y;
  }
  dot = warp_sum(dot);
  for (int d = lane * 2; d < D; d += 64) {
    float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + (size_t)row * ldx + d));
    float2 g = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dy + (size_t)row * lddy + d));
    *reinterpret_cast<uint32_t*>(dx + (size_t)row * lddx + d) =
        pack_bf16x2((g.x - (v.x / n) * dot) / n, (g.y - (v.y / n) * dot) / n);
  }
}

// weight_norm (dim=0): W[o,:] = g[o] * v[o,:] / ||v[o,:]||  -> bf16 weight for the GEMM; saves ||v||.
__global__ void __launch_bounds__(ROW_THREADS) weightnorm_fwd_kernel(const float* __restrict__ g, const float* __restrict__ v,
                                                                     int O, int I, __nv_bfloat16* __restrict__ w,
                                                                     float* __restrict__ vnorm) {
  B200_PDL_SYNC();
  const int row = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= O) return;
  const int lane = threadIdx.x & 31;
  float ss = 0.f;
  for (int d = lane * 4; d < I; d += 128) {
    float4 a = *reinterpret_cast<const float4*>(v + (size_t)row * I + d);
    ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  const float n = sqrtf(warp_sum(ss));
  if (lane == 0 && vnorm) vnorm[row] = n;
  const float s = g[row] / n;
  for (int d = lane * 4; d < I; d += 128) {
    float4 a = *reinterpret_cast<const float4*>(v + (size_t)row * I + d);
    uint2 p; p.x = pack_bf16x2(a.x * s, a.y * s); p.y = pack_bf16x2(a.z * s, a.w * s);
    *reinterpret_cast<uint2*>(w + (size_t)row * I + d) = p;
  }
}
// backward: dW f32 [O,I] -> dg[o] += <dW, v>/||v|| ; dv += g/||v|| * (dW - <dW,v>/||v||^2 * v)
__global__ void __launch_bounds__(ROW_THREADS) weightnorm_bwd_kernel(const float* __restrict__ dW, const float* __restrict__ g,
                                                                     const float* __restrict__ v, int O, int I,
                                                                     float* __restrict__ dg, float* __restrict__ dv) {
  B200_PDL_SYNC();
  const int row = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= O) return;
  const int lane = threadIdx.x & 31;
  float ss = 0.f, dot = 0.f;
  for (int d = lane * 4; d < I; d += 128) {
    float4 a = *reinterpret_cast<const float4*>(v + (size_t)row * I + d);
    float4 b = *reinterpret_cast<const float4*>(dW + (size_t)row * I + d);
    ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float n = sqrtf(ss);
  if (lane == 0) dg[row] += dot / n;
  const float s = g[row] / n, c = dot / ss;
  for (int d = lane * 4; d < I; d += 128) {
    float4 a = *reinterpret_cast<const float4*>(v + (size_t)row * I + d);
    float4 b = *reinterpret_cast<const float4*>(dW + (size_t)row * I + d);
    float4* q = reinterpret_cast<float4*>(dv + (size_t)row * I + d);
    float4 o = *q;
    o.x += s * (b.x - c * a.x); o.y += s * (b.y - c * a.y); o.z += s * (b.z - c * a.z); o.w += s * (b.w - c * a.w);
    *q = o;
  }
}

// ------------------------------------------------------------------------------------------------
// Tiny fp32 matmul C[M,N] (+)= A[M,K] * B (b_trans ? B[N,K]^T : B[K,N]); used for the positional-embedding
// resampling (a fixed [36,196] bicubic operator applied to pos_embed and its transpose in the backward).
__global__ void small_matmul_kernel(const float* __restrict__ A, long long lda, int a_trans, const float* __restrict__ B,
                                    long long ldb, int M, int N, int K, float* __restrict__ Cm, long long ldc, int accumulate) {
  B200_PDL_SYNC();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float a = a_trans ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k];
    acc += a * B[(size_t)k * ldb + n];
  }
  float* c = Cm + (size_t)m * ldc + n;
  *c = accumulate ? *c + acc : acc;
}

// f32 -> bf16 cast of a flat buffer (weights for the GEMMs)
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  B200_PDL_SYNC();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(x + i);
    uint2 p; p.x = pack_bf16x2(v.x, v.y); p.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(y + i) = p;
  } else {
    for (; i < n; ++i) y[i] = __float2bfloat16_rn(x[i]);
  }
}

}  // namespace b200

using namespace b200;

static inline bool ok_dim(int D) { return D > 0 && (D % 4) == 0 && D <= 128 * MAXV; }

extern "C" int b200_layernorm_fwd(const float* x, long long ldx, int T, int D, const float* w, const float* b, float eps,
                                  void* y, long long ldy, int out_bf16, float* mean, float* rstd, void* stream) {
  if (!x || !w || !b || !y || T <= 0 || !ok_dim(D) || (ldx % 4) || (ldy % 4)) return B200_ERR_INVALID_ARG;
  dim3 grid((T + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
#define B200_LN_FWD(V)                                                                                      \
  do {                                                                                                      \
    if (out_bf16) launch_kernel(ln_fwd_kernel<true, V>, grid, ROW_THREADS, 0, st, x, ldx, T, D, w, b, eps, y, ldy, mean, rstd); \
    else launch_kernel(ln_fwd_kernel<false, V>, grid, ROW_THREADS, 0, st, x, ldx, T, D, w, b, eps, y, ldy, mean, rstd);         \
  } while (0)
  if (D <= 384) B200_LN_FWD(3); else if (D <= 768) B200_LN_FWD(6); else B200_LN_FWD(8);
#undef B200_LN_FWD
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_layernorm_bwd(const void* dy, long long lddy, int dy_bf16, const float* x, long long ldx, int T, int D,
                                  const float* w, const float* mean, const float* rstd, float* dx, long long lddx,
                                  int accumulate, float* dw, float* db, void* stream) {
  if (!dy || !x || !w || !mean || !rstd || !dx || T <= 0 || !ok_dim(D)) return B200_ERR_INVALID_ARG;
  if ((ldx % 4) || (lddx % 4) || (lddy % 4) || (dw && !db)) return B200_ERR_INVALID_ARG;
  int grid = (T + 7) / 8;
  if (grid > 148 * 3) grid = 148 * 3;  // resident CTAs only: every CTA ends with 2*D global atomics
  size_t smem = 2 * D * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define B200_LN_BWD(V)                                                                                                           \
  do {                                                                                                                           \
    if (dy_bf16) launch_kernel(ln_bwd_kernel<true, V>, grid, ROW_THREADS, smem, st, dy, lddy, x, ldx, T, D, w, mean, rstd, dx, lddx, accumulate, dw, db); \
    else launch_kernel(ln_bwd_kernel<false, V>, grid, ROW_THREADS, smem, st, dy, lddy, x, ldx, T, D, w, mean, rstd, dx, lddx, accumulate, dw, db);        \
  } while (0)
  if (D <= 384) B200_LN_BWD(3); else if (D <= 768) B200_LN_BWD(6); else B200_LN_BWD(8);
#undef B200_LN_BWD
  B200_CHECK_LAUNCH();
  return B200_OK;
}


extern "C" int b200_layernorm_bwd_ls(const void* dy, long long lddy, int dy_bf16, const float* x, long long ldx, int T, int D,
                                     const float* w, const float* mean, const float* rstd, float* dx, long long lddx,
                                     int accumulate, float* dw, float* db, const void* o, long long ldo, const float* gamma,
                                     const float* rowscale, int rows_per_scale, void* dout, long long lddo, float* dgamma,
                                     float* dbias, void* stream) {
  if (!dy || !x || !w || !mean || !rstd || !dx || !o || !dout || T <= 0 || !ok_dim(D)) return B200_ERR_INVALID_ARG;
  if ((ldx % 4) || (lddx % 4) || (lddy % 4) || (ldo % 4) || (lddo % 4) || (dw && !db)) return B200_ERR_INVALID_ARG;
  int grid = (T + 7) / 8;
  if (grid > 148 * 3) grid = 148 * 3;
  const size_t smem = 4 * D * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  const int rps = rows_per_scale > 0 ? rows_per_scale : 1;
#define B200_LN_BWD_LS_(V, E)                                                                                               \
  do {                                                                                                                      \
    if (dy_bf16)                                                                                                            \
      launch_kernel(ln_bwd_ls_kernel<true, V, E>, grid, ROW_THREADS, smem, st, dy, lddy, x, ldx, T, D, w, mean, rstd, dx, lddx, accumulate, dw, db, \
                                                                    (const __nv_bfloat16*)o, ldo, gamma, rowscale, rps,    \
                                                                    (__nv_bfloat16*)dout, lddo, dgamma, dbias);            \
    else                                                                                                                    \
      launch_kernel(ln_bwd_ls_kernel<false, V, false>, grid, ROW_THREADS, smem, st, dy, lddy, x, ldx, T, D, w, mean, rstd, dx, lddx, accumulate, dw, db, \
                                                                  (const __nv_bfloat16*)o, ldo, gamma, rowscale, rps,      \
                                                                  (__nv_bfloat16*)dout, lddo, dgamma, dbias);              \
  } while (0)
#define B200_LN_BWD_LS(V)                                   \
  do {                                                      \
    if (D == 128 * V) B200_LN_BWD_LS_(V, true);             \
    else B200_LN_BWD_LS_(V, false);                         \
  } while (0)
  if (D <= 384) B200_LN_BWD_LS(3); else if (D <= 768) B200_LN_BWD_LS(6); else B200_LN_BWD_LS(8);
#undef B200_LN_BWD_LS_
#undef B200_LN_BWD_LS
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_im2col(const float* x, int B, int C, int H, int W, int p, void* cols, long long ldc, void* stream) {
  if (!x || !cols || B <= 0 || C <= 0 || p <= 0 || (H % p) || (W % p)) return B200_ERR_INVALID_ARG;
  long long total = (long long)B * (H / p) * (W / p) * C * p;
  int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  launch_kernel(im2col_kernel, grid, 256, 0, (cudaStream_t)stream, x, B, C, H, W, p, (__nv_bfloat16*)cols, ldc);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_assemble_tokens(const void* tok, long long ldt, const unsigned char* masks, const float* mask_token,
                                    const float* cls, const float* reg, const float* pos, int B, int Np, int R, int D,
                                    float* x, void* stream) {
  if (!tok || !cls || !pos || !x || B <= 0 || Np <= 0 || R < 0 || (D % 4) || (ldt % 4)) return B200_ERR_INVALID_ARG;
  if ((masks && !mask_token) || (R > 0 && !reg)) return B200_ERR_INVALID_ARG;
  long long rows = (long long)B * (1 + R + Np);
  launch_kernel(assemble_tokens_kernel, (unsigned)((rows + 7) / 8), ROW_THREADS, 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)tok, ldt, masks, mask_token, cls, reg, pos, B, Np, R, D, x);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_assemble_tokens_bwd(const float* dx, const unsigned char* masks, int B, int Np, int R, int D, void* dtok,
                                        long long lddt, float* dpos, float* dcls, float* dreg, float* dmask_token,
                                        void* stream) {
  if (!dx || !dtok || !dpos || !dcls || B <= 0 || Np <= 0 || (D % 4) || (lddt % 4)) return B200_ERR_INVALID_ARG;
  if ((masks && !dmask_token) || (R > 0 && !dreg)) return B200_ERR_INVALID_ARG;
  long long n = (long long)(1 + R + Np) * (D / 4);
  const int ysplit = B >= 64 ? 8 : (B >= 8 ? 4 : 1);
  launch_kernel(assemble_tokens_bwd_kernel, dim3((unsigned)((n + 127) / 128), ysplit), 128, 0, (cudaStream_t)stream, 
      dx, masks, B, Np, R, D, (__nv_bfloat16*)dtok, lddt, dpos, dcls, dreg, dmask_token);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_layerscale_bwd(const float* dx, long long lddx, const void* o, long long ldo, const float* gamma,
                                   const float* rowscale, int rows_per_scale, int T, int D, void* dout, long long lddo,
                                   float* dgamma, float* dbias, void* stream) {
  if (!dx || !o || !dout || T <= 0 || !ok_dim(D) || (lddx % 4) || (ldo % 4) || (lddo % 4)) return B200_ERR_INVALID_ARG;
  int grid = (T + 7) / 8;
  if (grid > 148 * 4) grid = 148 * 4;
  cudaStream_t st = (cudaStream_t)stream;
  const int rps = rows_per_scale > 0 ? rows_per_scale : 1;
#define B200_LS_BWD(V)                                                                                          \
  launch_kernel(layerscale_bwd_kernel<V>, grid, ROW_THREADS, 2 * D * sizeof(float), st, dx, lddx, (const __nv_bfloat16*)o, ldo, gamma, \
                                                                            rowscale, rps, T, D, (__nv_bfloat16*)dout, lddo, \
                                                                            dgamma, dbias)
  if (D <= 384) B200_LS_BWD(3); else if (D <= 768) B200_LS_BWD(6); else B200_LS_BWD(8);
#undef B200_LS_BWD
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// Sample-granular copies for batch-subset stochastic depth (layers/block.py:118-141): rows of `row_elems` floats.
//   gather : dst[j, :] = src[idx[j], :]      scatter: dst[idx[j], :] = src[j, :]
__global__ void copy_samples_kernel(const float* __restrict__ src, float* __restrict__ dst, const long long* __restrict__ idx,
                                    long long row_elems, int scatter) {
  B200_PDL_SYNC();
  const long long s = idx[blockIdx.y];
  const float4* in = reinterpret_cast<const float4*>(src + (scatter ? (long long)blockIdx.y : s) * row_elems);
  float4* out = reinterpret_cast<float4*>(dst + (scatter ? s : (long long)blockIdx.y) * row_elems);
  const long long nv = row_elems >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
}

extern "C" int b200_copy_samples(const float* src, float* dst, const long long* idx, int n_idx, long long row_elems, int scatter,
                                 void* stream) {
  if (!src || !dst || !idx || n_idx <= 0 || row_elems <= 0 || (row_elems % 4)) return B200_ERR_INVALID_ARG;
  if (((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return B200_ERR_UNSUPPORTED;
  long long per = (row_elems / 4 + 255) / 256;
  int gx = (int)(per < 64 ? per : 64);
  launch_kernel(copy_samples_kernel, dim3(gx, n_idx), 256, 0, (cudaStream_t)stream, src, dst, idx, row_elems, scatter);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_swiglu_fwd(const void* x12, long long ld12, int T, int H, void* hidden, long long ldh, void* stream) {
  if (!x12 || !hidden || T <= 0 || H <= 0) return B200_ERR_INVALID_ARG;
  if ((H % 8) || (ld12 % 8) || (ldh % 8) || ((uintptr_t)x12 & 15) || ((uintptr_t)hidden & 15)) return B200_ERR_UNSUPPORTED;
  const long long total = (long long)T * (H / 8);
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  launch_kernel(swiglu_fwd_kernel, grid, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x12, ld12, T, H, (__nv_bfloat16*)hidden, ldh);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_swiglu_bwd(const void* x12, long long ld12, const void* dhidden, long long lddh, int T, int H, void* dx12,
                               long long lddx12, void* stream) {
  if (!x12 || !dhidden || !dx12 || T <= 0 || H <= 0) return B200_ERR_INVALID_ARG;
  if ((H % 8) || (ld12 % 8) || (lddh % 8) || (lddx12 % 8) || ((uintptr_t)x12 & 15) || ((uintptr_t)dhidden & 15) ||
      ((uintptr_t)dx12 & 15))
    return B200_ERR_UNSUPPORTED;
  const long long total = (long long)T * (H / 8);
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  launch_kernel(swiglu_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x12, ld12, (const __nv_bfloat16*)dhidden, lddh, T, H,
                                                            (__nv_bfloat16*)dx12, lddx12);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_gather_rows(const float* src, long long lds, const long long* idx, int M, int D, int Np, int N, int off,
                                void* out, long long ldo, int out_bf16, void* stream) {
  if (!src || !idx || !out || M < 0 || (D % 4) || (lds % 4) || (ldo % 4)) return B200_ERR_INVALID_ARG;
  if (M == 0) return B200_OK;
  dim3 grid((M + 7) / 8);
  if (out_bf16) launch_kernel(gather_rows_kernel<true>, grid, ROW_THREADS, 0, (cudaStream_t)stream, src, lds, idx, M, D, Np, N, off, out, ldo);
  else launch_kernel(gather_rows_kernel<false>, grid, ROW_THREADS, 0, (cudaStream_t)stream, src, lds, idx, M, D, Np, N, off, out, ldo);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_scatter_rows(const void* in, long long ldi, int in_bf16, const long long* idx, int M, int D, int Np, int N,
                                 int off, float* dst, long long ldd, int accumulate, const int* count_dev, void* stream) {
  if (!in || !dst || M < 0 || (D % 4) || (ldi % 4) || (ldd % 4)) return B200_ERR_INVALID_ARG;
  if (M == 0) return B200_OK;
  dim3 grid((M + 7) / 8);
  if (in_bf16) launch_kernel(scatter_rows_kernel<true>, grid, ROW_THREADS, 0, (cudaStream_t)stream, in, ldi, idx, M, D, Np, N, off, dst, ldd, accumulate, count_dev);
  else launch_kernel(scatter_rows_kernel<false>, grid, ROW_THREADS, 0, (cudaStream_t)stream, in, ldi, idx, M, D, Np, N, off, dst, ldd, accumulate, count_dev);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_l2norm_fwd(const void* x, long long ldx, int R, int D, float eps, void* y, long long ldy, float* nrm,
                               void* stream) {
  if (!x || !y || !nrm || R <= 0 || (D % 2) || (ldx % 2) || (ldy % 2)) return B200_ERR_INVALID_ARG;
  launch_kernel(l2norm_fwd_kernel, (R + 7) / 8, ROW_THREADS, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, ldx, R, D, eps,
                                                                           (__nv_bfloat16*)y, ldy, nrm);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_l2norm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const float* nrm, int R, int D,
                               void* dx, long long lddx, void* stream) {
  if (!dy || !x || !nrm || !dx || R <= 0 || (D % 2) || (ldx % 2) || (lddy % 2) || (lddx % 2)) return B200_ERR_INVALID_ARG;
  launch_kernel(l2norm_bwd_kernel, (R + 7) / 8, ROW_THREADS, 0, (cudaStream_t)stream, (const __nv_bfloat16*)dy, lddy,
                                                                           (const __nv_bfloat16*)x, ldx, nrm, R, D,
                                                                           (__nv_bfloat16*)dx, lddx);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_weightnorm_fwd(const float* g, const float* v, int O, int I, void* w, float* vnorm, void* stream) {
  if (!g || !v || !w || O <= 0 || (I % 4)) return B200_ERR_INVALID_ARG;
  launch_kernel(weightnorm_fwd_kernel, (O + 7) / 8, ROW_THREADS, 0, (cudaStream_t)stream, g, v, O, I, (__nv_bfloat16*)w, vnorm);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_weightnorm_bwd(const float* dW, const float* g, const float* v, int O, int I, float* dg, float* dv,
                                   void* stream) {
  if (!dW || !g || !v || !dg || !dv || O <= 0 || (I % 4)) return B200_ERR_INVALID_ARG;
  launch_kernel(weightnorm_bwd_kernel, (O + 7) / 8, ROW_THREADS, 0, (cudaStream_t)stream, dW, g, v, O, I, dg, dv);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_small_matmul(const float* A, long long lda, int a_trans, const float* B, long long ldb, int M, int N, int K,
                                 float* C, long long ldc, int accumulate, void* stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return B200_ERR_INVALID_ARG;
  dim3 grid((N + 127) / 128, M);
  launch_kernel(small_matmul_kernel, grid, 128, 0, (cudaStream_t)stream, A, lda, a_trans, B, ldb, M, N, K, C, ldc, accumulate);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_cast_bf16(const float* x, void* y, long long n, void* stream) {
  if (!x || !y || n <= 0) return B200_ERR_INVALID_ARG;
  long long threads = (n + 3) / 4;
  launch_kernel(cast_bf16_kernel, (unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, n);
  B200_CHECK_LAUNCH();
  return B200_OK;
}
