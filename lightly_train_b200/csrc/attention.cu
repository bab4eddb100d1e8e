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
// first MMA of an accumulation chain: C = 0 (ptxas maps the zero inputs to RZ, no accumulator clearing moves)
__device__ __forceinline__ void mma16816_z(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
// round two fp32 values to bf16 precision (one F2FP + two bit ops instead of two F2F + two shifts)
__device__ __forceinline__ void bf16_round2(float& a, float& b) {
  const uint32_t u = pack_bf16x2(a, b);
  a = __uint_as_float(u << 16);
  b = __uint_as_float(u & 0xffff0000u);
}
static constexpr float kLog2e = 1.4426950408889634f;

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
// EXACT: N > (NKV16 - 1) * 16, i.e. only the last 16-column block can hold padded key columns (mask nothing else).
// S is rounded to bf16 before the scale is applied: identical to the reference order for power-of-two scales
// (head_dim 64 -> 0.125), and the scale then folds into the exponent FMA.
template <int NKV16, int WARPS, bool EXACT>
__global__ void __launch_bounds__(WARPS * 32)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld_tok, int N, int h, float scale,
                __nv_bfloat16* __restrict__ out, long long ld_out, float* __restrict__ lse) {
  B200_PDL_SYNC();
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

  // per-lane ldmatrix offsets: the XOR swizzle only touches address bits 4-6, so a k-step is `^ (ks << 5)` and a
  // 16-row block is `+ 2048` (folded into the instruction's immediate by ptxas)
  const int mi = lane >> 3;
  const uint32_t offA = tile_off((mi & 1) * 8 + (lane & 7), mi >> 1);  // A operand and transposed-B operand
  const uint32_t offB = tile_off((mi >> 1) * 8 + (lane & 7), mi & 1);  // B operand from an [n][k] tile
  uint32_t kaddr[4], vaddr[4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) { kaddr[ks] = sK + (offB ^ (ks << 5)); vaddr[ks] = sV + (offA ^ (ks << 5)); }
  const float sl2 = scale * kLog2e;

  const int nqt = (N + 15) / 16;
  for (int qt = warp; qt < nqt; qt += WARPS) {
    uint32_t aq[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(aq[ks], sQ + qt * 2048 + (offA ^ (ks << 5)));
    float s[NKV16 * 2][4];
#pragma unroll
    for (int j = 0; j < NKV16; ++j) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bk[4];
        ldsm_x4(bk, kaddr[ks] + j * 2048);
        if (ks == 0) {
          mma16816_z(s[2 * j], aq[ks], bk[0], bk[1]);
          mma16816_z(s[2 * j + 1], aq[ks], bk[2], bk[3]);
        } else {
          mma16816(s[2 * j], aq[ks], bk[0], bk[1]);
          mma16816(s[2 * j + 1], aq[ks], bk[2], bk[3]);
        }
      }
    }
    // softmax over the key axis; rows r0 = lane/4 (elements 0,1) and r0+8 (elements 2,3)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int t = 0; t < NKV16 * 2; ++t) {
      float v0 = s[t][0], v1 = s[t][1], v2 = s[t][2], v3 = s[t][3];
      bf16_round2(v0, v1);
      bf16_round2(v2, v3);
      if (!EXACT || t >= NKV16 * 2 - 2) {  // compile-time: only blocks that can hold padded key columns pay for the mask
        const int col = t * 8 + (lane & 3) * 2;
        if (col >= N) { v0 = -INFINITY; v2 = -INFINITY; }
        if (col + 1 >= N) { v1 = -INFINITY; v3 = -INFINITY; }
      }
      s[t][0] = v0; s[t][1] = v1; s[t][2] = v2; s[t][3] = v3;
      m0 = fmaxf(m0, fmaxf(v0, v1));
      m1 = fmaxf(m1, fmaxf(v2, v3));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float mb0 = -m0 * sl2, mb1 = -m1 * sl2;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int t = 0; t < NKV16 * 2; ++t) {
      s[t][0] = ex2_ftz(fmaf(s[t][0], sl2, mb0)); s[t][1] = ex2_ftz(fmaf(s[t][1], sl2, mb0));
      s[t][2] = ex2_ftz(fmaf(s[t][2], sl2, mb1)); s[t][3] = ex2_ftz(fmaf(s[t][3], sl2, mb1));
      l0 += s[t][0] + s[t][1];
      l1 += s[t][2] + s[t][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    if (lse && (lane & 3) == 0) {
      const int r0 = qt * 16 + (lane >> 2);
      if (r0 < N) lse[(size_t)bh * N + r0] = m0 * scale + __logf(l0);
      if (r0 + 8 < N) lse[(size_t)bh * N + r0 + 8] = m1 * scale + __logf(l1);
    }
    float o[8][4];
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
        ldsm_x4_t(bv, vaddr[dp] + j * 2048);
        if (j == 0) {
          mma16816_z(o[2 * dp], ap, bv[0], bv[1]);
          mma16816_z(o[2 * dp + 1], ap, bv[2], bv[3]);
        } else {
          mma16816(o[2 * dp], ap, bv[0], bv[1]);
          mma16816(o[2 * dp + 1], ap, bv[2], bv[3]);
        }
      }
    }
    store_tile(o, 1.f, stage_base + warp * 2048, out + (size_t)b * N * ld_out + head * HD, ld_out, qt * 16, N, lane);
  }
}

// Column sums of one 16 x 64 accumulator tile (mma C layout) into this warp's 64-float slot: the bias gradient of the
// qkv projection is the column sum of dqkv, taken here from registers instead of a second pass over dqkv in HBM.
// Values are rounded to bf16 first (the reference sums the bf16 dqkv tensor); rows flagged invalid are skipped.
__device__ __forceinline__ void colsum_tile(const float (&o)[8][4], float mul, bool v0, bool v1, float* slot, int lane) {
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    float a0 = o[t][0] * mul, a1 = o[t][1] * mul, b0 = o[t][2] * mul, b1 = o[t][3] * mul;
    bf16_round2(a0, a1);
    bf16_round2(b0, b1);
    float s0 = (v0 ? a0 : 0.f) + (v1 ? b0 : 0.f), s1 = (v0 ? a1 : 0.f) + (v1 ? b1 : 0.f);
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, off);
      s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    }
    if (lane < 4) {
      slot[t * 8 + lane * 2] += s0;
      slot[t * 8 + lane * 2 + 1] += s1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <int NKV16, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (WARPS <= 4 ? 4 : 1))
attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld_tok, const __nv_bfloat16* __restrict__ outp,
                const __nv_bfloat16* __restrict__ dout, long long ld_out, const float* __restrict__ lse, int N, int h,
                float scale, __nv_bfloat16* __restrict__ dqkv, long long ld_dtok, float* __restrict__ colsum) {
  B200_PDL_SYNC();
  constexpr int NP = NKV16 * 16;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem), sK = sQ + NP * 128, sV = sK + NP * 128, sDO = sV + NP * 128;
  const uint32_t sO = sDO + NP * 128;  // forward output panel, only needed for D_i = sum_d dO_id * O_id
  uint8_t* stage_base = smem + 5 * NP * 128;
  float* sL = reinterpret_cast<float*>(stage_base + WARPS * 2048);
  float* sD = sL + NP;
  float* sC = sD + NP;  // [WARPS][3][64] per-warp column sums of dq | dk | dv
  const int bh = blockIdx.x, b = bh / h, head = bh % h;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* base = qkv + (size_t)b * N * ld_tok + head * HD;
  const __nv_bfloat16* dob = dout + (size_t)b * N * ld_out + head * HD;
  load_panel(sQ, base, ld_tok, N, NP);
  load_panel(sK, base + (size_t)h * HD, ld_tok, N, NP);
  load_panel(sV, base + (size_t)2 * h * HD, ld_tok, N, NP);
  load_panel(sDO, dob, ld_out, N, NP);
  load_panel(sO, outp + (size_t)b * N * ld_out + head * HD, ld_out, N, NP);
  for (int r = threadIdx.x; r < NP; r += blockDim.x)
    sL[r] = (r < N) ? lse[(size_t)bh * N + r] * kLog2e : 0.f;  // base-2 log-sum-exp: p = 2^(s*log2e - L)
  if (colsum)
    for (int i = threadIdx.x; i < WARPS * 192; i += blockDim.x) sC[i] = 0.f;
  cp_async_wait_all();
  __syncthreads();
  // D_r = sum_d dO[r,d] * O[r,d] (== sum_j dP_rj P_rj), one thread per row straight from the two smem panels
  // (replaces a separate prep kernel: 24 launches and a second pass over O / dO per step); padded rows give 0
  for (int r = threadIdx.x; r < NP; r += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      const uint32_t off = tile_off(r, ch);
      uint32_t a[4], c[4];
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(sDO + off));
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]) : "r"(sO + off));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = unpack_bf16x2(a[i]), y = unpack_bf16x2(c[i]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
    }
    sD[r] = acc;
  }
  __syncthreads();

  // No masks are needed below: padded Q/K/V/dO rows are zero in smem, so a padded key column multiplies a zero K row
  // (phase A) and a padded query column a zero Q / dO row (phase B); padded output rows are never stored.  The
  // exponent is clamped so that such a dead element stays finite (0 * inf would poison the accumulators).
  const int nt = (N + 15) / 16;
  __nv_bfloat16* dq_dst = dqkv + (size_t)b * N * ld_dtok + head * HD;
  __nv_bfloat16* dk_dst = dq_dst + (size_t)h * HD;
  __nv_bfloat16* dv_dst = dq_dst + (size_t)2 * h * HD;
  uint8_t* stage = stage_base + warp * 2048;
  const int mi = lane >> 3;
  const uint32_t offA = tile_off((mi & 1) * 8 + (lane & 7), mi >> 1);  // A operand and transposed-B operand
  const uint32_t offB = tile_off((mi >> 1) * 8 + (lane & 7), mi & 1);  // B operand from an [n][k] tile
  constexpr uint32_t kPanel = NP * 128;                                // sQ, sK, sV, sDO are kPanel bytes apart
  constexpr float kClamp = 64.f;
  const float sl2 = scale * kLog2e;

  // ---------------- phase A: dQ, one 16-row query tile per warp iteration ----------------
  for (int qt = warp; qt < nt; qt += WARPS) {
    uint32_t aq[4][4], ado[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t a = sQ + qt * 2048 + (offA ^ (ks << 5));
      ldsm_x4(aq[ks], a);
      ldsm_x4(ado[ks], a + 3 * kPanel);
    }
    const int r0 = qt * 16 + (lane >> 2);
    const float L0 = sL[r0], L1 = sL[r0 + 8], D0 = sD[r0], D1 = sD[r0 + 8];
    float dq[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) { dq[t][0] = dq[t][1] = dq[t][2] = dq[t][3] = 0.f; }
    uint32_t kb = sK + offB, kt = sK + offA;
#pragma unroll 1
    for (int j = 0; j < nt; ++j, kb += 2048, kt += 2048) {
      float s[2][4], dp[2][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bk[4], bv[4];
        ldsm_x4(bk, kb ^ (ks << 5));
        ldsm_x4(bv, (kb ^ (ks << 5)) + kPanel);
        if (ks == 0) {
          mma16816_z(s[0], aq[ks], bk[0], bk[1]);
          mma16816_z(s[1], aq[ks], bk[2], bk[3]);
          mma16816_z(dp[0], ado[ks], bv[0], bv[1]);
          mma16816_z(dp[1], ado[ks], bv[2], bv[3]);
        } else {
          mma16816(s[0], aq[ks], bk[0], bk[1]);
          mma16816(s[1], aq[ks], bk[2], bk[3]);
          mma16816(dp[0], ado[ks], bv[0], bv[1]);
          mma16816(dp[1], ado[ks], bv[2], bv[3]);
        }
      }
      uint32_t ads[4];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float s0 = s[t][0], s1 = s[t][1], s2 = s[t][2], s3 = s[t][3];
        bf16_round2(s0, s1);
        bf16_round2(s2, s3);
        bf16_round2(dp[t][0], dp[t][1]);
        bf16_round2(dp[t][2], dp[t][3]);
        const float p0 = ex2_ftz(fminf(fmaf(s0, sl2, -L0), kClamp)), p1 = ex2_ftz(fminf(fmaf(s1, sl2, -L0), kClamp));
        const float p2 = ex2_ftz(fminf(fmaf(s2, sl2, -L1), kClamp)), p3 = ex2_ftz(fminf(fmaf(s3, sl2, -L1), kClamp));
        ads[2 * t] = pack_bf16x2(p0 * (dp[t][0] - D0), p1 * (dp[t][1] - D0));
        ads[2 * t + 1] = pack_bf16x2(p2 * (dp[t][2] - D1), p3 * (dp[t][3] - D1));
      }
#pragma unroll
      for (int d2 = 0; d2 < 4; ++d2) {
        uint32_t bkt[4];
        ldsm_x4_t(bkt, kt ^ (d2 << 5));
        mma16816(dq[2 * d2], ads, bkt[0], bkt[1]);
        mma16816(dq[2 * d2 + 1], ads, bkt[2], bkt[3]);
      }
    }
    if (colsum) colsum_tile(dq, scale, qt * 16 + (lane >> 2) < N, qt * 16 + (lane >> 2) + 8 < N, sC + warp * 192, lane);
    store_tile(dq, scale, stage, dq_dst, ld_dtok, qt * 16, N, lane);
  }

  // ---------------- phase B: dK, dV, one 16-row key tile per warp iteration ----------------
  for (int kt = warp; kt < nt; kt += WARPS) {
    uint32_t ak[4][4], av[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t a = sK + kt * 2048 + (offA ^ (ks << 5));
      ldsm_x4(ak[ks], a);
      ldsm_x4(av[ks], a + kPanel);
    }
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) { dk[t][0] = dk[t][1] = dk[t][2] = dk[t][3] = 0.f; dv[t][0] = dv[t][1] = dv[t][2] = dv[t][3] = 0.f; }
    const int kr0 = kt * 16 + (lane >> 2);  // this lane's key rows: kr0 (elements 0,1) and kr0 + 8 (elements 2,3)
    uint32_t qb = sQ + offB, qt_ = sQ + offA;
    const float* pL = sL + (lane & 3) * 2;
#pragma unroll 1
    for (int i = 0; i < nt; ++i, qb += 2048, qt_ += 2048, pL += 16) {
      float st[2][4], dpt[2][4];
      float2 Lv[2], Dv2[2];  // per-query-column log-sum-exp and D, fetched ahead of the MMAs
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        Lv[t] = *reinterpret_cast<const float2*>(pL + t * 8);
        Dv2[t] = *reinterpret_cast<const float2*>(pL + NP + t * 8);
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bq[4], bdo[4];
        ldsm_x4(bq, qb ^ (ks << 5));
        ldsm_x4(bdo, (qb ^ (ks << 5)) + 3 * kPanel);
        if (ks == 0) {
          mma16816_z(st[0], ak[ks], bq[0], bq[1]);
          mma16816_z(st[1], ak[ks], bq[2], bq[3]);
          mma16816_z(dpt[0], av[ks], bdo[0], bdo[1]);
          mma16816_z(dpt[1], av[ks], bdo[2], bdo[3]);
        } else {
          mma16816(st[0], ak[ks], bq[0], bq[1]);
          mma16816(st[1], ak[ks], bq[2], bq[3]);
          mma16816(dpt[0], av[ks], bdo[0], bdo[1]);
          mma16816(dpt[1], av[ks], bdo[2], bdo[3]);
        }
      }
      uint32_t apt[4], adst[4];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float s0 = st[t][0], s1 = st[t][1], s2 = st[t][2], s3 = st[t][3];
        bf16_round2(s0, s1);
        bf16_round2(s2, s3);
        bf16_round2(dpt[t][0], dpt[t][1]);
        bf16_round2(dpt[t][2], dpt[t][3]);
        const float p0 = ex2_ftz(fminf(fmaf(s0, sl2, -Lv[t].x), kClamp)), p1 = ex2_ftz(fminf(fmaf(s1, sl2, -Lv[t].y), kClamp));
        const float p2 = ex2_ftz(fminf(fmaf(s2, sl2, -Lv[t].x), kClamp)), p3 = ex2_ftz(fminf(fmaf(s3, sl2, -Lv[t].y), kClamp));
        apt[2 * t] = pack_bf16x2(p0, p1);
        apt[2 * t + 1] = pack_bf16x2(p2, p3);
        adst[2 * t] = pack_bf16x2(p0 * (dpt[t][0] - Dv2[t].x), p1 * (dpt[t][1] - Dv2[t].y));
        adst[2 * t + 1] = pack_bf16x2(p2 * (dpt[t][2] - Dv2[t].x), p3 * (dpt[t][3] - Dv2[t].y));
      }
#pragma unroll
      for (int d2 = 0; d2 < 4; ++d2) {
        uint32_t bqt[4], bdot[4];
        ldsm_x4_t(bqt, qt_ ^ (d2 << 5));
        ldsm_x4_t(bdot, (qt_ ^ (d2 << 5)) + 3 * kPanel);
        mma16816(dk[2 * d2], adst, bqt[0], bqt[1]);
        mma16816(dk[2 * d2 + 1], adst, bqt[2], bqt[3]);
        mma16816(dv[2 * d2], apt, bdot[0], bdot[1]);
        mma16816(dv[2 * d2 + 1], apt, bdot[2], bdot[3]);
      }
    }
    if (colsum) {  // padded key rows hold garbage (their scores are exp(-L), not 0): masked by the row flags
      colsum_tile(dk, scale, kr0 < N, kr0 + 8 < N, sC + warp * 192 + 64, lane);
      colsum_tile(dv, 1.f, kr0 < N, kr0 + 8 < N, sC + warp * 192 + 128, lane);
    }
    store_tile(dk, scale, stage, dk_dst, ld_dtok, kt * 16, N, lane);
    store_tile(dv, 1.f, stage, dv_dst, ld_dtok, kt * 16, N, lane);
  }
  if (colsum) {
    __syncthreads();
    for (int i = threadIdx.x; i < 192; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) t += sC[w * 192 + i];
      atomicAdd(colsum + (size_t)(i >> 6) * h * HD + head * HD + (i & 63), t);
    }
  }
}

template <int NKV16, int WARPS, bool EXACT>
static int launch_fwd(const void* qkv, long long ld_tok, int B, int N, int h, float scale, void* out, long long ld_out,
                      float* lse, cudaStream_t s) {
  constexpr int NP = NKV16 * 16;
  const int smem = 3 * NP * 128 + WARPS * 2048;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_fwd_kernel<NKV16, WARPS, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return B200_ERR_CUDA;
    attr = true;
  }
  launch_kernel(attn_fwd_kernel<NKV16, WARPS, EXACT>, B * h, WARPS * 32, smem, s, (const __nv_bfloat16*)qkv, ld_tok, N, h, scale,
                                                          (__nv_bfloat16*)out, ld_out, lse);
  B200_CHECK_LAUNCH();
  return B200_OK;
}
template <int NKV16, int WARPS>
static int launch_bwd(const void* qkv, long long ld_tok, const void* outp, const void* dout, long long ld_out, const float* lse,
                      int B, int N, int h, float scale, void* dqkv, long long ld_dtok, float* colsum, cudaStream_t s) {
  constexpr int NP = NKV16 * 16;
  const int smem = 5 * NP * 128 + WARPS * 2048 + 2 * NP * 4 + WARPS * 192 * 4;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_bwd_kernel<NKV16, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return B200_ERR_CUDA;
    attr = true;
  }
  launch_kernel(attn_bwd_kernel<NKV16, WARPS>, B * h, WARPS * 32, smem, s, (const __nv_bfloat16*)qkv, ld_tok, (const __nv_bfloat16*)outp,
                                                                 (const __nv_bfloat16*)dout, ld_out, lse, N, h, scale,
                                                          (__nv_bfloat16*)dqkv, ld_dtok, colsum);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attention_fwd(const void* qkv, long long ld_tok, int B, int N, int h, int head_dim, float scale, void* out,
                                  long long ld_out, float* lse, void* stream) {
  if (!qkv || !out || B <= 0 || N <= 0 || h <= 0 || !(scale > 0.f)) return B200_ERR_INVALID_ARG;
  if (head_dim != HD || (ld_tok % 8) || (ld_out % 8)) return B200_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = (N + 15) / 16;
#define B200_FWD(NB)                                                                                      \
  if (nb == NB) return launch_fwd<NB, 4, true>(qkv, ld_tok, B, N, h, scale, out, ld_out, lse, s);         \
  if (nb < NB) return launch_fwd<NB, 4, false>(qkv, ld_tok, B, N, h, scale, out, ld_out, lse, s);
  B200_FWD(3) B200_FWD(4) B200_FWD(13) B200_FWD(17)
#undef B200_FWD
  return B200_ERR_UNSUPPORTED;
}

extern "C" int b200_attention_bwd(const void* qkv, long long ld_tok, const void* out, const void* dout, long long ld_out,
                                  const float* lse, int B, int N, int h, int head_dim, float scale, void* dqkv,
                                  long long ld_dtok, float* dqkv_colsum, void* stream) {
  if (!qkv || !out || !dout || !lse || !dqkv || B <= 0 || N <= 0 || h <= 0 || !(scale > 0.f)) return B200_ERR_INVALID_ARG;
  if (head_dim != HD || (ld_tok % 8) || (ld_out % 8) || (ld_dtok % 8)) return B200_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = (N + 15) / 16;
  if (nb > 17) return B200_ERR_UNSUPPORTED;
  if (nb <= 3) return launch_bwd<3, 4>(qkv, ld_tok, out, dout, ld_out, lse, B, N, h, scale, dqkv, ld_dtok, dqkv_colsum, s);
  if (nb <= 4) return launch_bwd<4, 4>(qkv, ld_tok, out, dout, ld_out, lse, B, N, h, scale, dqkv, ld_dtok, dqkv_colsum, s);
  if (nb <= 13) return launch_bwd<13, 13>(qkv, ld_tok, out, dout, ld_out, lse, B, N, h, scale, dqkv, ld_dtok, dqkv_colsum, s);
  if (nb <= 17) return launch_bwd<17, 9>(qkv, ld_tok, out, dout, ld_out, lse, B, N, h, scale, dqkv, ld_dtok, dqkv_colsum, s);
  return B200_ERR_UNSUPPORTED;
}
