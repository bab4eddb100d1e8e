// Short-sequence multi-head attention core, forward + backward, head_dim = 64.
//
// Replaces the non-xformers path of Attention.forward
//   LT/_models/dinov2_vit/dinov2_vit_src/layers/attention.py:49-66  (q*scale @ k^T -> softmax -> @ v)
// and its autograd backward.  DINOv2 sequences are tiny (37..261 tokens), so one CTA owns one (image, head)
// pair with the whole Q/K/V (and dO in the backward) resident in shared memory; scores never touch HBM
// (the reference materialises [B,h,N,N] scores and probabilities).
//
// Round-1 implementation: warp-level mma.sync (m16n8k16, bf16 -> fp32) with ldmatrix from XOR-swizzled
// smem.  The score/probability tiles live in registers.  bf16 rounding points of the autocast reference are
// reproduced: S is rounded to bf16 before the fp32 softmax, P is rounded to bf16 before P@V.
//
// qkv layout: bf16 [B, N, 3, h, 64] (row pitch ld_tok elements); out / dout: bf16 [B, N, h*64].
#include "common.cuh"
#include "../../include/b200dino.h"

namespace b200 {

static constexpr int HD = 64;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk `chunk` of row `row` inside a [rows][64] bf16 tile (128 B rows, XOR swizzle)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}
// ldmatrix address helpers (see file header of the derivation in DESIGN.md)
//  A operand (row-major [m][k]) 16x16 block at (row0, kstep):     m0:(r0-7,klo) m1:(r8-15,klo) m2:(r0-7,khi) m3:(r8-15,khi)
__device__ __forceinline__ uint32_t addr_A(uint32_t tile, int row0, int ks, int lane) {
  const int mi = lane >> 3;
  return tile + tile_off(row0 + (mi & 1) * 8 + (lane & 7), ks * 2 + (mi >> 1));
}
//  B operand from a [n][k] row-major tile (non-transposed): two n-tiles of 8 at (n0, kstep):
//     m0:(n0-7,klo) m1:(n0-7,khi) m2:(n8-15,klo) m3:(n8-15,khi)  -> {b0,b1} of n-tile 0, {b0,b1} of n-tile 1
__device__ __forceinline__ uint32_t addr_B(uint32_t tile, int n0, int ks, int lane) {
  const int mi = lane >> 3;
  return tile + tile_off(n0 + (mi >> 1) * 8 + (lane & 7), ks * 2 + (mi & 1));
}
//  B operand from a [k][n] row-major tile (transposed load): k block of 16 at k0, n-tile pair dp (16 columns):
//     m0:(k0-7,nlo) m1:(k8-15,nlo) m2:(k0-7,nhi) m3:(k8-15,nhi)  -> {b0,b1} of n-tile 2dp, {b0,b1} of n-tile 2dp+1
__device__ __forceinline__ uint32_t addr_Bt(uint32_t tile, int k0, int dp, int lane) {
  const int mi = lane >> 3;
  return tile + tile_off(k0 + (mi & 1) * 8 + (lane & 7), dp * 2 + (mi >> 1));
}

// load a [N x 64] bf16 panel (row pitch ld) into a swizzled smem tile of NP rows, zero-filling rows >= N
__device__ __forceinline__ void load_panel(uint32_t tile, const __nv_bfloat16* src, long long ld, int N, int NP) {
  for (int i = threadIdx.x; i < NP * 8; i += blockDim.x) {
    const int row = i >> 3, chunk = i & 7;
    const bool ok = row < N;
    cp_async16(tile + tile_off(row, chunk), ok ? (const void*)(src + (size_t)row * ld + chunk * 8) : (const void*)src, ok ? 16 : 0);
  }
}

// write a 16 x 64 fp32 accumulator tile (mma C layout, 8 n-tiles) as bf16 rows [row0, row0+16) of a global
// [N x 64] panel, through a per-warp smem staging buffer for 16-byte coalesced stores
__device__ __forceinline__ void store_tile(const float (&o)[8][4], float mul, uint8_t* stage, __nv_bfloat16* dst,
                                           long long ld, int row0, int N, int lane) {
  const int r = lane >> 2, c = (lane & 3) * 2;
  __syncwarp();
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    *reinterpret_cast<uint32_t*>(stage + r * 128 + (t * 8 + c) * 2) = pack_bf16x2(o[t][0] * mul, o[t][1] * mul);
    *reinterpret_cast<uint32_t*>(stage + (r + 8) * 128 + (t * 8 + c) * 2) = pack_bf16x2(o[t][2] * mul, o[t][3] * mul);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = lane + 32 * i;  // 128 chunks of 16 B
    const int row = idx >> 3, chunk = idx & 7;
    if (row0 + row < N)
      *reinterpret_cast<uint4*>(dst + (size_t)(row0 + row) * ld + chunk * 8) =
          *reinterpret_cast<const uint4*>(stage + row * 128 + chunk * 16);
  }
}

// ------------------------------------------------------------------------------------------------
template <int NKV16, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld_tok, int N, int h, float scale,
                __nv_bfloat16* __restrict__ out, long long ld_out, float* __restrict__ lse) {
  constexpr int NP = NKV16 * 16;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem), sK = sQ + NP * 128, sV = sK + NP * 128;
  uint8_t* stage_base = smem + 3 * NP * 128;
  const int bh = blockIdx.x, b = bh / h, head = bh % h;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* base = qkv + (size_t)b * N * ld_tok + head * HD;
  load_panel(sQ, base, ld_tok, N, NP);
  load_panel(sK, base + (size_t)h * HD, ld_tok, N, NP);
  load_panel(sV, base + (size_t)2 * h * HD, ld_tok, N, NP);
  cp_async_wait_all();
  __syncthreads();

  const int nqt = (N + 15) / 16;
  for (int qt = warp; qt < nqt; qt += WARPS) {
    uint32_t aq[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(aq[ks], addr_A(sQ, qt * 16, ks, lane));
    float s[NKV16 * 2][4];
#pragma unroll
    for (int j = 0; j < NKV16; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[2 * j][e] = 0.f; s[2 * j + 1][e] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bk[4];
        ldsm_x4(bk, addr_B(sK, j * 16, ks, lane));
        mma16816(s[2 * j], aq[ks], bk[0], bk[1]);
        mma16816(s[2 * j + 1], aq[ks], bk[2], bk[3]);
      }
    }
    // softmax over the key axis; rows r0 = lane/4 (elements 0,1) and r0+8 (elements 2,3)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int t = 0; t < NKV16 * 2; ++t) {
      const int col = t * 8 + (lane & 3) * 2;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = bf16_round(s[t][e] * scale);
        if (col + (e & 1) >= N) v = -INFINITY;
        s[t][e] = v;
      }
      m0 = fmaxf(m0, fmaxf(s[t][0], s[t][1]));
      m1 = fmaxf(m1, fmaxf(s[t][2], s[t][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int t = 0; t < NKV16 * 2; ++t) {
      s[t][0] = __expf(s[t][0] - m0); s[t][1] = __expf(s[t][1] - m0);
      s[t][2] = __expf(s[t][2] - m1); s[t][3] = __expf(s[t][3] - m1);
      l0 += s[t][0] + s[t][1];
      l1 += s[t][2] + s[t][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    if (lse && (lane & 3) == 0) {
      const int r0 = qt * 16 + (lane >> 2);
      if (r0 < N) lse[(size_t)bh * N + r0] = m0 + __logf(l0);
      if (r0 + 8 < N) lse[(size_t)bh * N + r0 + 8] = m1 + __logf(l1);
    }
    float o[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) { o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < NKV16; ++j) {
      uint32_t ap[4];
      ap[0] = pack_bf16x2(s[2 * j][0] * inv0, s[2 * j][1] * inv0);
      ap[1] = pack_bf16x2(s[2 * j][2] * inv1, s[2 * j][3] * inv1);
      ap[2] = pack_bf16x2(s[2 * j + 1][0] * inv0, s[2 * j + 1][1] * inv0);
      ap[3] = pack_bf16x2(s[2 * j + 1][2] * inv1, s[2 * j + 1][3] * inv1);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t bv[4];
        ldsm_x4_t(bv, addr_Bt(sV, j * 16, dp, lane));
        mma16816(o[2 * dp], ap, bv[0], bv[1]);
        mma16816(o[2 * dp + 1], ap, bv[2], bv[3]);
      }
    }
    store_tile(o, 1.f, stage_base + warp * 2048, out + (size_t)b * N * ld_out + head * HD, ld_out, qt * 16, N, lane);
  }
}

// D[bh, n] = sum_d dO[b,n,h,d] * O[b,n,h,d]  (== sum_j dP_ij P_ij), one warp per token, all heads.
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                            long long ld_out, int B, int N, int h, float* __restrict__ dvec) {
  const long long tok = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= (long long)B * N) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(tok / N), n = (int)(tok % N);
  for (int head = 0; head < h; ++head) {
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + (size_t)tok * ld_out + head * HD + lane * 2));
    const float2 c = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + (size_t)tok * ld_out + head * HD + lane * 2));
    const float acc = warp_sum(a.x * c.x + a.y * c.y);
    if (lane == 0) dvec[((size_t)b * h + head) * N + n] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
template <int NKV16, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld_tok, const float* __restrict__ dvec,
                const __nv_bfloat16* __restrict__ dout, long long ld_out, const float* __restrict__ lse, int N, int h,
                float scale, __nv_bfloat16* __restrict__ dqkv, long long ld_dtok) {
  constexpr int NP = NKV16 * 16;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem), sK = sQ + NP * 128, sV = sK + NP * 128, sDO = sV + NP * 128;
  uint8_t* stage_base = smem + 4 * NP * 128;
  float* sL = reinterpret_cast<float*>(stage_base + WARPS * 2048);
  float* sD = sL + NP;
  const int bh = blockIdx.x, b = bh / h, head = bh % h;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* base = qkv + (size_t)b * N * ld_tok + head * HD;
  const __nv_bfloat16* dob = dout + (size_t)b * N * ld_out + head * HD;
  load_panel(sQ, base, ld_tok, N, NP);
  load_panel(sK, base + (size_t)h * HD, ld_tok, N, NP);
  load_panel(sV, base + (size_t)2 * h * HD, ld_tok, N, NP);
  load_panel(sDO, dob, ld_out, N, NP);
  for (int r = threadIdx.x; r < NP; r += blockDim.x) {
    const bool ok = r < N;
    sD[r] = ok ? dvec[(size_t)bh * N + r] : 0.f;
    sL[r] = ok ? lse[(size_t)bh * N + r] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  const int nt = (N + 15) / 16;
  __nv_bfloat16* dq_dst = dqkv + (size_t)b * N * ld_dtok + head * HD;
  __nv_bfloat16* dk_dst = dq_dst + (size_t)h * HD;
  __nv_bfloat16* dv_dst = dq_dst + (size_t)2 * h * HD;
  uint8_t* stage = stage_base + warp * 2048;

  // ---------------- phase A: dQ, one 16-row query tile per warp iteration ----------------
  for (int qt = warp; qt < nt; qt += WARPS) {
    uint32_t aq[4][4], ado[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ldsm_x4(aq[ks], addr_A(sQ, qt * 16, ks, lane));
      ldsm_x4(ado[ks], addr_A(sDO, qt * 16, ks, lane));
    }
    const int r0 = qt * 16 + (lane >> 2);
    const float L0 = sL[r0], L1 = sL[r0 + 8], D0 = sD[r0], D1 = sD[r0 + 8];
    float dq[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) { dq[t][0] = dq[t][1] = dq[t][2] = dq[t][3] = 0.f; }
#pragma unroll 1
    for (int j = 0; j < nt; ++j) {
      float s[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, dp[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bk[4], bv[4];
        ldsm_x4(bk, addr_B(sK, j * 16, ks, lane));
        ldsm_x4(bv, addr_B(sV, j * 16, ks, lane));
        mma16816(s[0], aq[ks], bk[0], bk[1]);
        mma16816(s[1], aq[ks], bk[2], bk[3]);
        mma16816(dp[0], ado[ks], bv[0], bv[1]);
        mma16816(dp[1], ado[ks], bv[2], bv[3]);
      }
      uint32_t ads[4];
      float ds[2][4];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int col = j * 16 + t * 8 + (lane & 3) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float L = (e < 2) ? L0 : L1, Dv = (e < 2) ? D0 : D1;
          float p = (col + (e & 1) < N) ? __expf(bf16_round(s[t][e] * scale) - L) : 0.f;
          ds[t][e] = p * (bf16_round(dp[t][e]) - Dv);
        }
      }
      ads[0] = pack_bf16x2(ds[0][0], ds[0][1]); ads[1] = pack_bf16x2(ds[0][2], ds[0][3]);
      ads[2] = pack_bf16x2(ds[1][0], ds[1][1]); ads[3] = pack_bf16x2(ds[1][2], ds[1][3]);
#pragma unroll
      for (int d2 = 0; d2 < 4; ++d2) {
        uint32_t bkt[4];
        ldsm_x4_t(bkt, addr_Bt(sK, j * 16, d2, lane));
        mma16816(dq[2 * d2], ads, bkt[0], bkt[1]);
        mma16816(dq[2 * d2 + 1], ads, bkt[2], bkt[3]);
      }
    }
    store_tile(dq, scale, stage, dq_dst, ld_dtok, qt * 16, N, lane);
  }

  // ---------------- phase B: dK, dV, one 16-row key tile per warp iteration ----------------
  for (int kt = warp; kt < nt; kt += WARPS) {
    uint32_t ak[4][4], av[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      ldsm_x4(ak[ks], addr_A(sK, kt * 16, ks, lane));
      ldsm_x4(av[ks], addr_A(sV, kt * 16, ks, lane));
    }
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) { dk[t][0] = dk[t][1] = dk[t][2] = dk[t][3] = 0.f; dv[t][0] = dv[t][1] = dv[t][2] = dv[t][3] = 0.f; }
    const int kr0 = kt * 16 + (lane >> 2);
#pragma unroll 1
    for (int i = 0; i < nt; ++i) {
      float st[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, dpt[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
      float2 Lv[2], Dv2[2];  // per-query-column log-sum-exp and D, fetched ahead of the MMAs
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int qc = i * 16 + t * 8 + (lane & 3) * 2;
        Lv[t] = *reinterpret_cast<const float2*>(sL + qc);
        Dv2[t] = *reinterpret_cast<const float2*>(sD + qc);
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bq[4], bdo[4];
        ldsm_x4(bq, addr_B(sQ, i * 16, ks, lane));
        ldsm_x4(bdo, addr_B(sDO, i * 16, ks, lane));
        mma16816(st[0], ak[ks], bq[0], bq[1]);
        mma16816(st[1], ak[ks], bq[2], bq[3]);
        mma16816(dpt[0], av[ks], bdo[0], bdo[1]);
        mma16816(dpt[1], av[ks], bdo[2], bdo[3]);
      }
      uint32_t apt[4], adst[4];
      float pt[2][4], dst_[2][4];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int qc = i * 16 + t * 8 + (lane & 3) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = qc + (e & 1);
          const int kr = kr0 + ((e >> 1) ? 8 : 0);
          const bool valid = (q < N) && (kr < N);
          const float Lq = (e & 1) ? Lv[t].y : Lv[t].x, Dq = (e & 1) ? Dv2[t].y : Dv2[t].x;
          const float p = valid ? __expf(bf16_round(st[t][e] * scale) - Lq) : 0.f;
          pt[t][e] = p;
          dst_[t][e] = p * (bf16_round(dpt[t][e]) - Dq);
        }
      }
      apt[0] = pack_bf16x2(pt[0][0], pt[0][1]); apt[1] = pack_bf16x2(pt[0][2], pt[0][3]);
      apt[2] = pack_bf16x2(pt[1][0], pt[1][1]); apt[3] = pack_bf16x2(pt[1][2], pt[1][3]);
      adst[0] = pack_bf16x2(dst_[0][0], dst_[0][1]); adst[1] = pack_bf16x2(dst_[0][2], dst_[0][3]);
      adst[2] = pack_bf16x2(dst_[1][0], dst_[1][1]); adst[3] = pack_bf16x2(dst_[1][2], dst_[1][3]);
#pragma unroll
      for (int d2 = 0; d2 < 4; ++d2) {
        uint32_t bqt[4], bdot[4];
        ldsm_x4_t(bqt, addr_Bt(sQ, i * 16, d2, lane));
        ldsm_x4_t(bdot, addr_Bt(sDO, i * 16, d2, lane));
        mma16816(dk[2 * d2], adst, bqt[0], bqt[1]);
        mma16816(dk[2 * d2 + 1], adst, bqt[2], bqt[3]);
        mma16816(dv[2 * d2], apt, bdot[0], bdot[1]);
        mma16816(dv[2 * d2 + 1], apt, bdot[2], bdot[3]);
      }
    }
    store_tile(dk, scale, stage, dk_dst, ld_dtok, kt * 16, N, lane);
    store_tile(dv, 1.f, stage, dv_dst, ld_dtok, kt * 16, N, lane);
  }
}

template <int NKV16, int WARPS>
static int launch_fwd(const void* qkv, long long ld_tok, int B, int N, int h, float scale, void* out, long long ld_out,
                      float* lse, cudaStream_t s) {
  constexpr int NP = NKV16 * 16;
  const int smem = 3 * NP * 128 + WARPS * 2048;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_fwd_kernel<NKV16, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return B200_ERR_CUDA;
    attr = true;
  }
  attn_fwd_kernel<NKV16, WARPS><<<B * h, WARPS * 32, smem, s>>>((const __nv_bfloat16*)qkv, ld_tok, N, h, scale,
                                                          (__nv_bfloat16*)out, ld_out, lse);
  B200_CHECK_LAUNCH();
  return B200_OK;
}
template <int NKV16, int WARPS>
static int launch_bwd(const void* qkv, long long ld_tok, const float* dvec, const void* dout, long long ld_out, const float* lse,
                      int B, int N, int h, float scale, void* dqkv, long long ld_dtok, cudaStream_t s) {
  constexpr int NP = NKV16 * 16;
  const int smem = 4 * NP * 128 + WARPS * 2048 + 2 * NP * 4;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_bwd_kernel<NKV16, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return B200_ERR_CUDA;
    attr = true;
  }
  attn_bwd_kernel<NKV16, WARPS><<<B * h, WARPS * 32, smem, s>>>((const __nv_bfloat16*)qkv, ld_tok, dvec,
                                                                 (const __nv_bfloat16*)dout, ld_out, lse, N, h, scale,
                                                          (__nv_bfloat16*)dqkv, ld_dtok);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attention_fwd(const void* qkv, long long ld_tok, int B, int N, int h, int head_dim, float scale, void* out,
                                  long long ld_out, float* lse, void* stream) {
  if (!qkv || !out || B <= 0 || N <= 0 || h <= 0) return B200_ERR_INVALID_ARG;
  if (head_dim != HD || (ld_tok % 8) || (ld_out % 8)) return B200_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = (N + 15) / 16;
  if (nb <= 3) return launch_fwd<3, 4>(qkv, ld_tok, B, N, h, scale, out, ld_out, lse, s);
  if (nb <= 4) return launch_fwd<4, 4>(qkv, ld_tok, B, N, h, scale, out, ld_out, lse, s);
  if (nb <= 13) return launch_fwd<13, 4>(qkv, ld_tok, B, N, h, scale, out, ld_out, lse, s);
  if (nb <= 17) return launch_fwd<17, 4>(qkv, ld_tok, B, N, h, scale, out, ld_out, lse, s);
  return B200_ERR_UNSUPPORTED;
}

extern "C" int b200_attention_bwd(const void* qkv, long long ld_tok, const void* out, const void* dout, long long ld_out,
                                  const float* lse, int B, int N, int h, int head_dim, float scale, void* dqkv,
                                  long long ld_dtok, float* dvec_ws, void* stream) {
  if (!qkv || !out || !dout || !lse || !dqkv || !dvec_ws || B <= 0 || N <= 0 || h <= 0) return B200_ERR_INVALID_ARG;
  if (head_dim != HD || (ld_tok % 8) || (ld_out % 8) || (ld_dtok % 8)) return B200_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = (N + 15) / 16;
  if (nb > 17) return B200_ERR_UNSUPPORTED;
  attn_bwd_prep_kernel<<<(unsigned)(((long long)B * N + 7) / 8), 256, 0, s>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, ld_out, B, N, h, dvec_ws);
  B200_CHECK_LAUNCH();
  if (nb <= 3) return launch_bwd<3, 4>(qkv, ld_tok, dvec_ws, dout, ld_out, lse, B, N, h, scale, dqkv, ld_dtok, s);
  if (nb <= 4) return launch_bwd<4, 4>(qkv, ld_tok, dvec_ws, dout, ld_out, lse, B, N, h, scale, dqkv, ld_dtok, s);
  if (nb <= 13) return launch_bwd<13, 13>(qkv, ld_tok, dvec_ws, dout, ld_out, lse, B, N, h, scale, dqkv, ld_dtok, s);
  if (nb <= 17) return launch_bwd<17, 9>(qkv, ld_tok, dvec_ws, dout, ld_out, lse, B, N, h, scale, dqkv, ld_dtok, s);
  return B200_ERR_UNSUPPORTED;
}
