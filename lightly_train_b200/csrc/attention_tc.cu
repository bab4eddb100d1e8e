// Short-sequence attention on the 5th-generation tensor cores: tcgen05.mma with fp32 accumulators in TMEM, operands
// staged by TMA into 128B-swizzled shared memory.  Forward and backward for 128 < N <= 256 (forward) / N <= 208
// (backward) tokens per image, head_dim 64 -- the global-crop sequences of the ViT (N = 197 / 201); shorter and longer
// sequences stay on the warp-level kernels of attention.cu (b200_attention_fwd / _bwd dispatch).
//
// Replaces Attention.forward of LT/_models/dinov2_vit/dinov2_vit_src/layers/attention.py:49-66 and its autograd backward.
// Rounding points of the bf16-autocast reference are kept: S = q k^T is rounded to bf16 before the fp32 softmax, P is
// rounded to bf16 before P V (here BEFORE the 1/l normalisation, flash-style: same relative rounding error), dP and dS
// are rounded to bf16 (autocast matmul outputs / operands), dq dk dv are written in bf16.
//
// One CTA per (image, head) -- or, for N <= 128, per (G = 128 / N consecutive images, head): the G sequences share one
// 128-row MMA tile and a block-diagonal mask keeps every query on its own image's keys (local crops: 3 x 37 tokens).
// qkv: bf16 [B*N, 3*h*64] (row pitch ld_tok); out / dout: bf16 [B*N, h*64].
// Rows past the group's tokens inside a TMA box belong to the next image (or are zero-filled past the tensor): finite
// data that is always MASKED (keys outside a query's own image get probability 0, query rows past the group are zeroed
// before they are contracted over and never stored).
//
// FORWARD  (9 warps): warp 8 = control (TMA loads, all tcgen05.mma issues); warps 0-3 / 4-7 = the two 128-row query
//   tiles, one query row per thread (TMEM lane = row, so row max / row sum are thread-local: no shuffles).
//     S_t = Q_t K^T          M128 x N=NKV x K16 x4, fp32 -> TMEM columns [256 t, 256 t + NKV)
//     P   = 2^((s - m) * scale * log2e) -> bf16 -> swizzled smem tile (A operand of the next MMA), l = sum p
//     O_t = P V              M128 x N64 x K16 x NKV/16 -> TMEM columns [256 t, 256 t + 64) (S_t is dead by then)
//     out = O / l, lse = m * scale + ln l
// BACKWARD (9 warps): warp 8 = control; warps 0-3 / 4-7 = the key columns [0, NKV/2) / [NKV/2, NKV) of the current
//   128-row query tile (LSE and D_i are known, so the row can be split across two threads without a reduction).
//   Per query tile t:  S = Q_t K^T -> P (smem) ; dV += P^T dO_t ; dP = dO_t V^T (reuses S's TMEM columns) ->
//   dS = P (dP - D) (smem, in place of P) ; dK += dS^T Q_t ; dQ_t = dS K (reuses the same TMEM columns) -> global.
//   dK / dV accumulate in TMEM over the two query tiles (keys on the lanes: two 128-row M tiles x 64 columns each).
//   The qkv-bias gradient (column sums of the bf16 dq | dk | dv) is accumulated from registers like attention.cu does.
#include <cuda.h>
#include "common.cuh"
#include "../../include/b200dino.h"

namespace b200 {
namespace attn_tc {

static constexpr int HD = 64;
static constexpr int BLOCK_Q = 128;                 // query rows per MMA tile (= TMEM lanes)
static constexpr int TILE_BYTES = BLOCK_Q * 128;    // one [128 x 64] bf16 tile (also one 64-key slab of a P tile)
static constexpr float kLog2e = 1.4426950408889634f;

// same encodings as gemm_tcgen05.cu (hardware-verified there)
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t instr_desc(int m, int n, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                    // D = f32
  d |= 1u << 7;                    // A = bf16
  d |= 1u << 10;                   // B = bf16
  d |= (uint32_t)(a_mn ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn ? 1 : 0) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 64-thread barrier of warp pair q (0..3) with a compile-time id each, so that ptxas reserves 5 barriers, not all 16
__device__ __forceinline__ void pair_bar_sync4(int q) {
  switch (q) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}
// 16-byte chunk `chunk` (8 bf16) of row `row` of a [rows][64] bf16 tile with 128-byte rows, 128B swizzle
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// ====================================================================================================================
// forward
// ====================================================================================================================
template <int NKV16>
struct FwdCfg {
  static constexpr int NKV = NKV16 * 16;                   // padded key count (MMA N of S, K of P V)
  static constexpr int KV_BYTES = NKV * 128;
  static constexpr int P_SLABS = (NKV + 63) / 64;
  static constexpr int OFF_K = 2 * TILE_BYTES;             // after the two query tiles
  static constexpr int OFF_V = OFF_K + KV_BYTES;
  static constexpr int OFF_P = (OFF_V + KV_BYTES + 1023) / 1024 * 1024;
  static constexpr int P_BYTES = P_SLABS * TILE_BYTES;     // per query tile
  static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;  // barriers + TMEM slot, alignment slack
  static constexpr int THREADS = 9 * 32;
  static_assert(NKV <= 256, "one S tile per 256 TMEM columns");
  static_assert(SMEM_BYTES <= 232448, "shared memory");
};

template <int NKV16>
__global__ void __launch_bounds__(FwdCfg<NKV16>::THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, int B, int N, int G, int h,
                   float scale, __nv_bfloat16* __restrict__ out, long long ld_out, float* __restrict__ lse) {
  pdl_launch_dependents();  // the wait follows the barrier / TMEM set-up below
  using C = FwdCfg<NKV16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* bar_load = bars + 0;  // TMA bytes landed
  uint64_t* bar_s = bars + 1;     // [2] S_t in TMEM
  uint64_t* bar_p = bars + 3;     // [2] P_t in smem (4 worker warps each)
  uint64_t* bar_o = bars + 5;     // [2] O_t in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // one CTA = G consecutive images x one head (G > 1: short sequences packed into one 128-row tile, block-diagonal mask)
  const int bg = blockIdx.x / h, head = blockIdx.x % h;
  const int b0 = bg * G, n_img = min(G, B - b0);
  const int rows_valid = n_img * N;
  const int n_tiles = (rows_valid + BLOCK_Q - 1) / BLOCK_Q;  // 1 or 2 query tiles

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_s + t, 1);
      mbar_init(bar_p + t, 4);
      mbar_init(bar_o + t, 1);
    }
    fence_barrier_init();
  }
  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
    }
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ===================== control warp: TMA + MMA issue =====================
    const int row0 = b0 * N;  // first token row of this group in the [B*N, 3*h*64] qkv matrix
    if (lane == 0) {
      mbar_expect_tx(bar_load, 2 * TILE_BYTES + 2 * C::KV_BYTES);
      tma_load_2d(smem, &tmQ, bar_load, head * HD, row0);
      tma_load_2d(smem + TILE_BYTES, &tmQ, bar_load, head * HD, row0 + BLOCK_Q);
      tma_load_2d(smem + C::OFF_K, &tmKV, bar_load, (h + head) * HD, row0);
      tma_load_2d(smem + C::OFF_V, &tmKV, bar_load, (2 * h + head) * HD, row0);
    }
    __syncwarp();
    mbar_wait(bar_load, 0);
    const uint32_t s0 = smem_u32(smem);
    const uint64_t dk = smem_desc(s0 + C::OFF_K, 16, 1024);        // B of S: K rows, K-major (N = keys)
    const uint64_t dv = smem_desc(s0 + C::OFF_V, 64 * 128, 1024);  // B of O: V rows, MN-major (N = head dim)
    const uint32_t idesc_s = instr_desc(BLOCK_Q, C::NKV, 0, 0);
    const uint32_t idesc_o = instr_desc(BLOCK_Q, HD, 0, 1);
    if (elect_one_sync()) {
      for (int t = 0; t < n_tiles; ++t) {
        const uint64_t dq = smem_desc(s0 + t * TILE_BYTES, 16, 1024);  // A of S: Q tile, K-major
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_f16(tmem_base + t * 256, dq + (uint64_t)((ks * 32) >> 4), dk + (uint64_t)((ks * 32) >> 4), idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(bar_s + t);
      }
    }
    __syncwarp();
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(bar_p + t, 0);  // P_t complete in smem (workers fenced generic->async proxy before arriving)
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t dp = smem_desc(s0 + C::OFF_P + t * C::P_BYTES, 16, 1024);  // A of O: P tile, K-major (K = keys)
#pragma unroll 1
        for (int ks = 0; ks < NKV16; ++ks)
          umma_f16(tmem_base + t * 256, dp + (uint64_t)(((ks >> 2) * TILE_BYTES + (ks & 3) * 32) >> 4),
                   dv + (uint64_t)((ks * 16 * 128) >> 4), idesc_o, ks > 0 ? 1u : 0u);
        umma_commit(bar_o + t);
      }
      __syncwarp();
    }
  } else if ((warp >> 2) < n_tiles) {
    // ===================== worker warps: tile t = warp / 4, one query row per thread =====================
    const int t = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;        // row inside the 128-row tile == TMEM lane
    const int m = t * BLOCK_Q + r;      // query row inside the group
    const int img = min(m / N, n_img - 1);
    const int klo = img * N, khi = klo + N;  // this row's keys: its own image's tokens
    const uint32_t tS = tmem_base + ((uint32_t)(q * 32) << 16) + t * 256;
    const float sl2 = scale * kLog2e;
    uint8_t* sP = smem + C::OFF_P + t * C::P_BYTES;
    mbar_wait(bar_s + t, 0);
    tc_fence_after();
    // pass 1: row max of the bf16-rounded scores over the valid keys
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < NKV16; ++c) {
      uint32_t v[16];
      tmem_ld_32x16(tS + c * 16, v);
      tmem_ld_wait();
      if (c * 16 >= klo && c * 16 + 16 <= khi) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float2 rr = unpack_bf16x2(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          mx = fmaxf(mx, fmaxf(rr.x, rr.y));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float2 rr = unpack_bf16x2(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          const int k = c * 16 + i;
          if (k >= klo && k < khi) mx = fmaxf(mx, rr.x);
          if (k + 1 >= klo && k + 1 < khi) mx = fmaxf(mx, rr.y);
        }
      }
    }
    const float mb = -mx * sl2;
    // pass 2: p = 2^(s*sl2 - m*sl2), row sum in fp32, P (unnormalised, bf16) -> swizzled smem A tile
    float l = 0.f;
#pragma unroll 1
    for (int c = 0; c < NKV16; ++c) {
      uint32_t v[16];
      tmem_ld_32x16(tS + c * 16, v);
      tmem_ld_wait();
      uint32_t pw[8];
      if (c * 16 >= klo && c * 16 + 16 <= khi) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float2 rr = unpack_bf16x2(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          const float p0 = ex2_ftz(fmaf(rr.x, sl2, mb)), p1 = ex2_ftz(fmaf(rr.y, sl2, mb));
          l += p0 + p1;
          pw[i >> 1] = pack_bf16x2(p0, p1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float2 rr = unpack_bf16x2(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          const int k = c * 16 + i;
          const float p0 = (k >= klo && k < khi) ? ex2_ftz(fmaf(rr.x, sl2, mb)) : 0.f;
          const float p1 = (k + 1 >= klo && k + 1 < khi) ? ex2_ftz(fmaf(rr.y, sl2, mb)) : 0.f;
          l += p0 + p1;
          pw[i >> 1] = pack_bf16x2(p0, p1);
        }
      }
      // keys [c*16, c*16+16) = 16-byte chunks (c&3)*2, +1 of slab c/4
      uint8_t* slab = sP + (c >> 2) * TILE_BYTES;
      const int ch = (c & 3) * 2;
      *reinterpret_cast<uint4*>(slab + swz(r, ch)) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
      *reinterpret_cast<uint4*>(slab + swz(r, ch + 1)) = make_uint4(pw[4], pw[5], pw[6], pw[7]);
    }
    fence_proxy_async();  // the MMA (async proxy) reads what this thread just wrote through the generic proxy
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p + t);
    if (lse && m < rows_valid) lse[((size_t)(b0 + img) * h + head) * N + (m - klo)] = mx * scale + __logf(l);
    // O row: scale by 1/l, round to bf16, 128 contiguous bytes per row
    mbar_wait(bar_o + t, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    __nv_bfloat16* orow = out + ((size_t)b0 * N + m) * ld_out + head * HD;
#pragma unroll
    for (int c = 0; c < HD / 16; ++c) {
      uint32_t v[16];
      tmem_ld_32x16(tS + c * 16, v);
      tmem_ld_wait();
      if (m < rows_valid) {
        uint32_t ow[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) ow[i >> 1] = pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
        *reinterpret_cast<uint4*>(orow + c * 16) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        *reinterpret_cast<uint4*>(orow + c * 16 + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ====================================================================================================================
// forward, schedule 7: P stays in TENSOR MEMORY (A operand of P V read from TMEM); one query tile per CTA, two CTAs per SM
// ====================================================================================================================
// ncu on a sixteen-softmax-warp version of schedule 1 (profiles/r02_ncu_attn_tc16.txt; since removed): per (image, head) pair
// the softmax warps spend 23 % of the CTA's life
// waiting for its 85 KB of TMA loads (every CTA of a wave loads at the same time, then computes while HBM idles), 9 % on
// P V, 5 % in the exit tail; and the 2 x 64 KB of swizzled shared memory that carry P to the second MMA are what keeps a
// second CTA (or a second K / V stage) off the SM.  Here the probabilities never leave tensor memory:
//   * S = Q K^T lands in TMEM columns [0, NKV); every row is shared by two threads (warps w / w + 4, the same TMEM lane
//     quarter) that read their half of the row ONCE, keep it in registers as packed bf16 pairs (keys outside the query's
//     image = -inf) and exchange the half-row maximum / sum through shared memory + a 64-thread named barrier;
//   * P = 2^((s - m) scale log2e) is written back with tcgen05.st as packed bf16 pairs into columns [0, NKV / 2) -- over S,
//     which is dead once both halves hold their rows -- and O = P V is issued with the A operand in TMEM (tcgen05.mma
//     [d], [a], b-desc), per half as soon as that half's four warps have stored their columns; O accumulates in columns
//     [128, 192) (also over dead S);
//   * shared memory is only Q (16 KB) + K + V: 69 KB at NKV = 208, TMEM 256 columns, so two CTAs share an SM and one's loads,
//     set-up and tail hide behind the other's softmax.
template <int NKV16>
struct Fwd7Cfg {
  static constexpr int NKV = NKV16 * 16;
  static constexpr int KV_BYTES = NKV * 128;
  static constexpr int OFF_K = TILE_BYTES;
  static constexpr int OFF_V = OFF_K + KV_BYTES;
  static constexpr int OFF_X = (OFF_V + KV_BYTES + 1023) / 1024 * 1024;  // floats: half-row max [2][128], half-row sum [2][128]
  static constexpr int OFF_BAR = OFF_X + 4 * 128 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr int THREADS = 9 * 32;
  static constexpr int HA = (NKV16 + 1) / 2;                // 16-key chunks of the first / second column half
  static constexpr int HB = NKV16 - HA;
  static constexpr int TM_O = 128;                          // O accumulator columns [128, 192)
  static constexpr int TMEM_COLS = 256;
  static_assert(NKV <= 256 && HB >= 1 && 2 * SMEM_BYTES <= 232448, "two CTAs per SM");
};

// one half of one query row: HALF 0 = key chunks [0, HA), HALF 1 = [HA, NKV16)
template <int NKV16, int HALF>
__device__ __forceinline__ void fwd7_worker(uint32_t tS, float* xmax, float* xsum, uint64_t* bar_s, uint64_t* bar_p, uint64_t* bar_o,
                                            int r, int q, int lane, int klo, int khi, bool all_inside, float sl2, float scale,
                                            float* lse_row, __nv_bfloat16* orow, uint32_t ph = 0, uint64_t* bar_tfree = nullptr) {
  using C = Fwd7Cfg<NKV16>;
  constexpr int C0 = HALF ? C::HA : 0, NC = HALF ? C::HB : C::HA;
  uint32_t srow[NC * 8];
  uint32_t buf[16];
  mbar_wait(bar_s, ph);
  tc_fence_after();
  // pass 1 (the only read of S): packed bf16 row half -> registers, row max on packed pairs
  uint32_t mx2 = 0xFF80FF80u;
  tmem_ld_32x16(tS + C0 * 16, buf);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) srow[c * 8 + i] = pack_bf16x2(__uint_as_float(buf[2 * i]), __uint_as_float(buf[2 * i + 1]));
    if (c + 1 < NC) tmem_ld_32x16(tS + (C0 + c + 1) * 16, buf);
    const int kc = (C0 + c) * 16;
    if (!(all_inside && kc + 16 <= khi)) {  // warp-uniform when one image per CTA: only the last chunk takes this path
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kc + 2 * i;
        if (!(k >= klo && k < khi)) srow[c * 8 + i] = (srow[c * 8 + i] & 0xFFFF0000u) | 0x0000FF80u;  // -inf
        if (!(k + 1 >= klo && k + 1 < khi)) srow[c * 8 + i] = (srow[c * 8 + i] & 0x0000FFFFu) | 0xFF800000u;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i += 2) mx2 = bf16x2_max(mx2, bf16x2_max(srow[c * 8 + i], srow[c * 8 + i + 1]));
  }
  float mx;
  {
    const float2 mm = unpack_bf16x2(mx2);
    mx = fmaxf(mm.x, mm.y);
  }
  xmax[HALF * 128 + r] = mx;
  tc_fence_before();              // my S reads are complete before the partner (after the barrier) overwrites those columns
  pair_bar_sync4(q);
  tc_fence_after();
  mx = fmaxf(mx, xmax[(HALF ^ 1) * 128 + r]);
  const float mb = -mx * sl2;
  // pass 2 from registers: p -> bf16 pairs -> TMEM columns [(C0 + c) * 8, + 8) (over the dead S)
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    uint32_t pw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 rr = unpack_bf16x2(srow[c * 8 + i]);
      const float p0 = ex2_ftz(fmaf(rr.x, sl2, mb)), p1 = ex2_ftz(fmaf(rr.y, sl2, mb));
      l0 += p0;
      l1 += p1;
      pw[i] = pack_bf16x2(p0, p1);
    }
    tmem_st_32x8(tS + (C0 + c) * 8, pw);
  }
  float l = l0 + l1;
  xsum[HALF * 128 + r] = l;
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_p + HALF);
  pair_bar_sync4(q);
  l += xsum[(HALF ^ 1) * 128 + r];
  if (HALF == 0 && lse_row) *lse_row = mx * scale + __logf(l);
  mbar_wait(bar_o, ph);
  tc_fence_after();
  const float inv = 1.f / l;
  uint32_t ov[32];
  tmem_ld_32x32(tS + C::TM_O + HALF * 32, ov);
  tmem_ld_wait();
  if (bar_tfree) {  // persistent schedule: the next tile's S may overwrite this TMEM region
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_tfree);
  }
  if (orow) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        ow[j] = pack_bf16x2(__uint_as_float(ov[i * 8 + 2 * j]) * inv, __uint_as_float(ov[i * 8 + 2 * j + 1]) * inv);
      *reinterpret_cast<uint4*>(orow + HALF * 32 + i * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
}

template <int NKV16>
__global__ void __launch_bounds__(Fwd7Cfg<NKV16>::THREADS, 2)
attn_fwd_tc7_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, int B, int N, int G, int h,
                    int tiles_per_group, float scale, __nv_bfloat16* __restrict__ out, long long ld_out, float* __restrict__ lse) {
  pdl_launch_dependents();
  using C = Fwd7Cfg<NKV16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* bar_load = bars + 0;
  uint64_t* bar_s = bars + 1;     // S in TMEM
  uint64_t* bar_p = bars + 2;     // [2 halves] P columns of that half stored to TMEM (4 warps each)
  uint64_t* bar_o = bars + 4;     // O complete in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* xmax = reinterpret_cast<float*>(smem + C::OFF_X);   // [half][128]
  float* xsum = xmax + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x % tiles_per_group;
  const int grp = blockIdx.x / tiles_per_group;
  const int bg = grp / h, head = grp % h;
  const int b0 = bg * G, n_img = min(G, B - b0);
  const int rows_valid = n_img * N;

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_p + 1, 4);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
    }
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ===================== control warp: TMA + MMA issue =====================
    const int row0 = b0 * N;
    if (lane == 0) {
      mbar_expect_tx(bar_load, TILE_BYTES + 2 * C::KV_BYTES);
      tma_load_2d(smem, &tmQ, bar_load, head * HD, row0 + t * BLOCK_Q);
      tma_load_2d(smem + C::OFF_K, &tmKV, bar_load, (h + head) * HD, row0);
      tma_load_2d(smem + C::OFF_V, &tmKV, bar_load, (2 * h + head) * HD, row0);
    }
    __syncwarp();
    mbar_wait(bar_load, 0);
    const uint32_t s0 = smem_u32(smem);
    const uint64_t dq = smem_desc(s0, 16, 1024);
    const uint64_t dk = smem_desc(s0 + C::OFF_K, 16, 1024);
    const uint64_t dv = smem_desc(s0 + C::OFF_V, 64 * 128, 1024);  // B of O: V rows, MN-major (N = head dim)
    const uint32_t idesc_s = instr_desc(BLOCK_Q, C::NKV, 0, 0);
    const uint32_t idesc_o = instr_desc(BLOCK_Q, HD, 0, 1);
    if (elect_one_sync()) {
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        umma_f16(tmem_base, dq + (uint64_t)(ks * 2), dk + (uint64_t)(ks * 2), idesc_s, ks > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncwarp();
    for (int half = 0; half < 2; ++half) {
      mbar_wait(bar_p + half, 0);
      tc_fence_after();
      if (elect_one_sync()) {
        const int k0 = half ? C::HA : 0, k1 = half ? NKV16 : C::HA;
#pragma unroll 1
        for (int ks = k0; ks < k1; ++ks)  // A: 16 keys = 8 packed columns of every lane; B: 16 V rows
          umma_f16_ts(tmem_base + C::TM_O, tmem_base + ks * 8, dv + (uint64_t)((ks * 16 * 128) >> 4), idesc_o, ks > 0 ? 1u : 0u);
        if (half == 1) umma_commit(bar_o);
      }
      __syncwarp();
    }
  } else {
    // ===================== worker warps: lane quarter q, column half =====================
    const int half = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const int m = t * BLOCK_Q + r;
    const int img = min(m / N, n_img - 1);
    const int klo = img * N, khi = klo + N;
    const uint32_t tS = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool valid = m < rows_valid;
    float* lse_row = (lse && valid) ? lse + ((size_t)(b0 + img) * h + head) * N + (m - klo) : nullptr;
    __nv_bfloat16* orow = valid ? out + ((size_t)b0 * N + m) * ld_out + head * HD : nullptr;
    if (half == 0)
      fwd7_worker<NKV16, 0>(tS, xmax, xsum, bar_s, bar_p, bar_o, r, q, lane, klo, khi, G == 1, scale * kLog2e, scale, lse_row, orow);
    else
      fwd7_worker<NKV16, 1>(tS, xmax, xsum, bar_s, bar_p, bar_o, r, q, lane, klo, khi, G == 1, scale * kLog2e, scale, lse_row, orow);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ====================================================================================================================
// forward, schedule 8: schedule 7 in persistent CTAs (two per SM) with the next tile's Q / K and V prefetched
// ====================================================================================================================
// ncu on schedule 7 (profiles/r02_ncu_attn_tc7.txt): 24 % of the softmax warps' time is still the wait for the tile's TMA
// loads + S (the two CTAs of an SM start together and stay in lock-step), 4 % the initial set-up barrier.  Q and K are dead
// as soon as the S MMAs have retired -- early in a tile -- so the NEXT tile's Q and K are loaded right then into the same
// buffers; V is dead once P V has retired, and the next V then has the whole softmax phase of the next tile to land.  Shared
// memory stays at schedule 7's 69 KB; barriers, TMEM and tensor maps are set up once per CTA.
template <int NKV16>
__global__ void __launch_bounds__(Fwd7Cfg<NKV16>::THREADS, 2)
attn_fwd_tc8_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, int B, int N, int G, int h,
                    int tiles_per_group, int n_tiles_total, float scale, __nv_bfloat16* __restrict__ out, long long ld_out,
                    float* __restrict__ lse) {
  pdl_launch_dependents();
  using C = Fwd7Cfg<NKV16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* bar_qk = bars + 0;     // Q tile + K landed
  uint64_t* bar_v = bars + 1;      // V landed
  uint64_t* bar_s = bars + 2;      // S in TMEM (= the S MMAs have retired: Q / K buffers free)
  uint64_t* bar_p = bars + 3;      // [2 halves] P columns stored (4 warps each)
  uint64_t* bar_o = bars + 5;      // O complete (= P V retired: V buffer free)
  uint64_t* bar_tfree = bars + 6;  // O drained by all 8 worker warps: the TMEM region may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* xmax = reinterpret_cast<float*>(smem + C::OFF_X);
  float* xsum = xmax + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_p + 1, 4);
    mbar_init(bar_o, 1);
    mbar_init(bar_tfree, 8);
    fence_barrier_init();
  }
  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
    }
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const int n_mine = (n_tiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles blockIdx.x, + gridDim.x, ...
  auto tile_coords = [&](int it, int& t, int& head, int& b0) {
    const int w = blockIdx.x + it * gridDim.x;
    t = w % tiles_per_group;
    const int grp = w / tiles_per_group;
    head = grp % h;
    b0 = (grp / h) * G;
  };

  if (warp == 8) {
    // ===================== control warp =====================
    const uint32_t s0 = smem_u32(smem);
    const uint64_t dq = smem_desc(s0, 16, 1024);
    const uint64_t dk = smem_desc(s0 + C::OFF_K, 16, 1024);
    const uint64_t dv = smem_desc(s0 + C::OFF_V, 64 * 128, 1024);
    const uint32_t idesc_s = instr_desc(BLOCK_Q, C::NKV, 0, 0);
    const uint32_t idesc_o = instr_desc(BLOCK_Q, HD, 0, 1);
    auto load_qk = [&](int it) {
      int t, head, b0;
      tile_coords(it, t, head, b0);
      mbar_expect_tx(bar_qk, TILE_BYTES + C::KV_BYTES);
      tma_load_2d(smem, &tmQ, bar_qk, head * HD, b0 * N + t * BLOCK_Q);
      tma_load_2d(smem + C::OFF_K, &tmKV, bar_qk, (h + head) * HD, b0 * N);
    };
    auto load_v = [&](int it) {
      int t, head, b0;
      tile_coords(it, t, head, b0);
      mbar_expect_tx(bar_v, C::KV_BYTES);
      tma_load_2d(smem + C::OFF_V, &tmKV, bar_v, (2 * h + head) * HD, b0 * N);
    };
    if (lane == 0 && n_mine > 0) {
      load_qk(0);
      load_v(0);
    }
    __syncwarp();
    for (int it = 0; it < n_mine; ++it) {
      const uint32_t ph = it & 1;
      mbar_wait(bar_qk, ph);
      if (it > 0) mbar_wait(bar_tfree, ph ^ 1);  // the previous tile's O has been read out of TMEM
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_f16(tmem_base, dq + (uint64_t)(ks * 2), dk + (uint64_t)(ks * 2), idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(bar_s);
      }
      __syncwarp();
      if (it + 1 < n_mine) {  // Q / K of the next tile as soon as this tile's S MMAs have read the buffers
        mbar_wait(bar_s, ph);
        if (lane == 0) load_qk(it + 1);
        __syncwarp();
      }
      mbar_wait(bar_v, ph);
      for (int half = 0; half < 2; ++half) {
        mbar_wait(bar_p + half, ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const int k0 = half ? C::HA : 0, k1 = half ? NKV16 : C::HA;
#pragma unroll 1
          for (int ks = k0; ks < k1; ++ks)
            umma_f16_ts(tmem_base + C::TM_O, tmem_base + ks * 8, dv + (uint64_t)((ks * 16 * 128) >> 4), idesc_o, ks > 0 ? 1u : 0u);
          if (half == 1) umma_commit(bar_o);
        }
        __syncwarp();
      }
      if (it + 1 < n_mine) {  // V of the next tile once P V has retired
        mbar_wait(bar_o, ph);
        if (lane == 0) load_v(it + 1);
        __syncwarp();
      }
    }
  } else {
    // ===================== worker warps: lane quarter q, column half =====================
    const int half = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t tS = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int it = 0; it < n_mine; ++it) {
      int t, head, b0;
      tile_coords(it, t, head, b0);
      const int n_img = min(G, B - b0);
      const int rows_valid = n_img * N;
      const int m = t * BLOCK_Q + r;
      const int img = min(m / N, n_img - 1);
      const int klo = img * N, khi = klo + N;
      const bool valid = m < rows_valid;
      float* lse_row = (lse && valid) ? lse + ((size_t)(b0 + img) * h + head) * N + (m - klo) : nullptr;
      __nv_bfloat16* orow = valid ? out + ((size_t)b0 * N + m) * ld_out + head * HD : nullptr;
      if (half == 0)
        fwd7_worker<NKV16, 0>(tS, xmax, xsum, bar_s, bar_p, bar_o, r, q, lane, klo, khi, G == 1, scale * kLog2e, scale, lse_row, orow,
                              it & 1, bar_tfree);
      else
        fwd7_worker<NKV16, 1>(tS, xmax, xsum, bar_s, bar_p, bar_o, r, q, lane, klo, khi, G == 1, scale * kLog2e, scale, lse_row, orow,
                              it & 1, bar_tfree);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ====================================================================================================================
// backward
// ====================================================================================================================
template <int NKV16, int NG = 2>
struct BwdCfg {
  static constexpr int NKV = NKV16 * 16;                   // padded key count, <= 208
  // NKV <= 128 (short sequences packed into ONE 128-row tile): one query tile, one key M-tile, half the P tile -- 99 KB of
  // shared memory and 256 TMEM columns, so two CTAs share an SM (with NG = 2: 9 warps each)
  static constexpr bool SMALL = NKV <= 128;
  static constexpr int QT = SMALL ? 1 : 2;                 // query tiles held in smem
  static constexpr int MT = SMALL ? 1 : 2;                 // 128-key M tiles of the dK / dV accumulators
  static constexpr int P_SLABS = SMALL ? 2 : 4;            // 64-key slabs of the P / dS tile
  static constexpr int KV_BYTES = NKV * 128;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_DO = QT * TILE_BYTES;
  static constexpr int OFF_K = 2 * QT * TILE_BYTES;
  static constexpr int OFF_V = OFF_K + KV_BYTES;
  static constexpr int OFF_P = (OFF_V + KV_BYTES + 1023) / 1024 * 1024;
  static constexpr int OFF_F = OFF_P + P_SLABS * TILE_BYTES;   // floats: L[256], D[256], colsum[192]
  static constexpr int OFF_BAR = OFF_F + (256 + 256 + 192) * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr int WORKERS = 4 * NG;                   // worker warps: NG column groups x 4 lane quarters
  static constexpr int THREADS = (WORKERS + 1) * 32;       // + the control warp (the last one)
  static constexpr int CTAS_PER_SM = (SMALL && NG == 2) ? 2 : 1;
  // TMEM columns: S / dP / dQ share [0, NKV); dK: M tile 0 | 1; dV likewise
  static constexpr int TM_SDP = 0, TM_DK = SMALL ? 128 : 256, TM_DV = SMALL ? 192 : 384;
  static constexpr int TMEM_COLS = SMALL ? 256 : 512;
  __host__ __device__ static constexpr int qb(int g) { return (NKV16 * g + NG - 1) / NG; }   // column group g owns the 16-key chunks [qb(g), qb(g+1))
  static_assert(NKV <= 256, "S / dP tile");
  static_assert(CTAS_PER_SM * SMEM_BYTES <= 232448, "shared memory");
};

// column sums over the 32 rows held by a warp (one row per lane, 32 consecutive columns in v): butterfly transpose-
// reduce, 31 shuffles instead of 160; on return lane j holds the sum of column j.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float mine = upper ? v[i + half] : v[i];
      const float give = upper ? v[i] : v[i + half];
      v[i] = mine + __shfl_xor_sync(0xffffffffu, give, half);
    }
  }
  return v[0];
}

// Drain one [128 x 64] fp32 accumulator tile (one row per thread, 4 warps): out = bf16(acc * mul) -> global row (when
// `valid`), and this warp's column sums of the bf16-rounded values -> smem atomics on cs[0..64).
// 32 columns [c*32, c*32+32) of the tile
__device__ __forceinline__ void drain_cols32(uint32_t taddr, int c, float mul, bool valid, __nv_bfloat16* grow, float* cs, int lane) {
  uint32_t v[32];
  tmem_ld_32x32(taddr + c * 32, v);
  tmem_ld_wait();
  float f[32];
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    w[i >> 1] = pack_bf16x2(__uint_as_float(v[i]) * mul, __uint_as_float(v[i + 1]) * mul);
    const float2 rr = unpack_bf16x2(w[i >> 1]);
    f[i] = valid ? rr.x : 0.f;
    f[i + 1] = valid ? rr.y : 0.f;
  }
  if (valid) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<uint4*>(grow + c * 32 + i * 8) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  }
  if (cs) {
    const float s = warp_colsum32(f, lane);
    atomicAdd(cs + c * 32 + lane, s);
  }
}
__device__ __forceinline__ void drain_tile(uint32_t taddr, float mul, bool valid, __nv_bfloat16* grow, float* cs, int lane) {
  drain_cols32(taddr, 0, mul, valid, grow, cs, lane);
  drain_cols32(taddr, 1, mul, valid, grow, cs, lane);
}

// Backward worker passes over the 16-key chunks [C0, C1) of one row.  The next chunk's TMEM load is issued as soon as the
// current one has been packed to bf16 pairs, so it is in flight during the arithmetic.
//   P pass : p = 2^(s*sl2 - L) (0 on padded rows / keys outside [klo, khi)) -> bf16 -> swizzled smem row.  Chunks below
//            `nfull` hold only valid keys for every row of the CTA and run a mask-free body (padded ROWS carry L = +inf there,
//            which makes their p exactly 0 without a test); the rest test every key.
//   dS pass: dS = p * (dP - D), bf16, written in place of p
template <int C0, int C1>
__device__ __forceinline__ void bwd_p_pass(uint32_t tcol, uint8_t* sP, int r, bool row_ok, int klo, int khi, int nfull, float sl2,
                                           float L) {
  uint32_t buf[16];
  const float Linf = row_ok ? L : INFINITY;
  tmem_ld_32x16(tcol + C0 * 16, buf);
  auto take = [&](int c, uint32_t (&pk)[8]) {
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(__uint_as_float(buf[2 * i]), __uint_as_float(buf[2 * i + 1]));
    if (c + 1 < C1) tmem_ld_32x16(tcol + (c + 1) * 16, buf);
  };
  auto store = [&](int c, const uint32_t (&pw)[8]) {
    uint8_t* slab = sP + (c >> 2) * TILE_BYTES;
    const int ch = (c & 3) * 2;
    *reinterpret_cast<uint4*>(slab + swz(r, ch)) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
    *reinterpret_cast<uint4*>(slab + swz(r, ch + 1)) = make_uint4(pw[4], pw[5], pw[6], pw[7]);
  };
  int c = C0;
  const int cf = min(C1, nfull);
#pragma unroll 1
  for (; c < cf; ++c) {
    uint32_t pk[8], pw[8];
    take(c, pk);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 rr = unpack_bf16x2(pk[i]);
      pw[i] = pack_bf16x2(ex2_ftz(fmaf(rr.x, sl2, -Linf)), ex2_ftz(fmaf(rr.y, sl2, -Linf)));
    }
    store(c, pw);
  }
#pragma unroll 1
  for (; c < C1; ++c) {
    uint32_t pk[8], pw[8];
    take(c, pk);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 rr = unpack_bf16x2(pk[i]);
      const int k = c * 16 + 2 * i;
      const float p0 = (row_ok && k >= klo && k < khi) ? ex2_ftz(fmaf(rr.x, sl2, -L)) : 0.f;
      const float p1 = (row_ok && k + 1 >= klo && k + 1 < khi) ? ex2_ftz(fmaf(rr.y, sl2, -L)) : 0.f;
      pw[i] = pack_bf16x2(p0, p1);
    }
    store(c, pw);
  }
}
template <int C0, int C1>
__device__ __forceinline__ void bwd_ds_pass(uint32_t tcol, uint8_t* sP, int r, float D) {
  uint32_t buf[16];
  tmem_ld_32x16(tcol + C0 * 16, buf);
#pragma unroll 1
  for (int c = C0; c < C1; ++c) {
    uint8_t* slab = sP + (c >> 2) * TILE_BYTES;
    const int ch = (c & 3) * 2;
    const uint4 pa = *reinterpret_cast<const uint4*>(slab + swz(r, ch));
    const uint4 pb = *reinterpret_cast<const uint4*>(slab + swz(r, ch + 1));
    tmem_ld_wait();
    uint32_t dpk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dpk[i] = pack_bf16x2(__uint_as_float(buf[2 * i]), __uint_as_float(buf[2 * i + 1]));  // bf16 dP
    if (c + 1 < C1) tmem_ld_32x16(tcol + (c + 1) * 16, buf);
    const uint32_t pin[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
    uint32_t dw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 dp = unpack_bf16x2(dpk[i]);
      const float2 pp = unpack_bf16x2(pin[i]);  // p = 0 on masked rows / keys: dS = 0 there whatever dP holds
      dw[i] = pack_bf16x2(pp.x * (dp.x - D), pp.y * (dp.y - D));
    }
    *reinterpret_cast<uint4*>(slab + swz(r, ch)) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
    *reinterpret_cast<uint4*>(slab + swz(r, ch + 1)) = make_uint4(dw[4], dw[5], dw[6], dw[7]);
  }
}

template <int NKV16, int NG>
__global__ void __launch_bounds__(BwdCfg<NKV16, NG>::THREADS, BwdCfg<NKV16, NG>::CTAS_PER_SM)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmDO, const __nv_bfloat16* __restrict__ outp,
                   const __nv_bfloat16* __restrict__ dout, long long ld_out, const float* __restrict__ lse, int B, int N, int G,
                   int h, float scale, __nv_bfloat16* __restrict__ dqkv, long long ld_dtok, float* __restrict__ colsum) {
  pdl_launch_dependents();  // the wait follows the barrier / TMEM set-up below
  using C = BwdCfg<NKV16, NG>;
  constexpr int CTRL = C::WORKERS;  // control warp index
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* sL = reinterpret_cast<float*>(smem + C::OFF_F);  // base-2 log-sum-exp per query row (256)
  float* sD = sL + 256;                                   // D_i = sum_d dO O
  float* sC = sD + 256;                                   // column sums of dq | dk | dv
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* bar_load = bars + 0;
  uint64_t* bar_s = bars + 1;     // S of the current tile in TMEM
  uint64_t* bar_p = bars + 2;     // P in smem (all worker warps)
  uint64_t* bar_dp = bars + 3;    // dP in TMEM (and dV MMAs of this tile retired: P may be overwritten)
  uint64_t* bar_ds = bars + 4;    // dS in smem (all worker warps)
  uint64_t* bar_dq = bars + 5;    // dQ in TMEM (and dK MMAs retired)
  uint64_t* bar_free = bars + 6;  // dQ drained (4 warps, or 8 with NG = 4): the shared TMEM columns may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bg = blockIdx.x / h, head = blockIdx.x % h;  // G consecutive images x one head (see the forward kernel)
  const int b0 = bg * G, n_img = min(G, B - b0);
  const int rows_valid = n_img * N;
  const int n_tiles = (rows_valid + BLOCK_Q - 1) / BLOCK_Q;

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, C::WORKERS);
    mbar_init(bar_dp, 1);
    mbar_init(bar_ds, C::WORKERS);
    mbar_init(bar_dq, 1);
    mbar_init(bar_free, NG == 2 ? 4 : 8);
    fence_barrier_init();
  }
  if (warp == CTRL) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      tma_prefetch_desc(&tmDO);
    }
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const size_t row0 = (size_t)b0 * N;

  if (warp == CTRL) {
    // ===================== control warp =====================
    if (lane == 0) {
      mbar_expect_tx(bar_load, 2 * C::QT * TILE_BYTES + 2 * C::KV_BYTES);
      tma_load_2d(smem + C::OFF_Q, &tmQ, bar_load, head * HD, (int)row0);
      tma_load_2d(smem + C::OFF_DO, &tmDO, bar_load, head * HD, (int)row0);
      if (C::QT == 2) {
        tma_load_2d(smem + C::OFF_Q + TILE_BYTES, &tmQ, bar_load, head * HD, (int)row0 + BLOCK_Q);
        tma_load_2d(smem + C::OFF_DO + TILE_BYTES, &tmDO, bar_load, head * HD, (int)row0 + BLOCK_Q);
      }
      tma_load_2d(smem + C::OFF_K, &tmKV, bar_load, (h + head) * HD, (int)row0);
      tma_load_2d(smem + C::OFF_V, &tmKV, bar_load, (2 * h + head) * HD, (int)row0);
    }
    __syncwarp();
    mbar_wait(bar_load, 0);
    const uint32_t s0 = smem_u32(smem);
    const uint64_t dK_kmaj = smem_desc(s0 + C::OFF_K, 16, 1024);         // B of S: [keys x d], K-major (N = keys, K = d)
    const uint64_t dV_kmaj = smem_desc(s0 + C::OFF_V, 16, 1024);         // B of dP
    const uint64_t dK_mn = smem_desc(s0 + C::OFF_K, 64 * 128, 1024);     // B of dQ: (K = keys, N = d), MN-major
    const uint64_t dP_kmaj = smem_desc(s0 + C::OFF_P, 16, 1024);         // A of dQ: dS [q x keys], K-major, 64-key slabs
    const uint64_t dP_mn = smem_desc(s0 + C::OFF_P, TILE_BYTES, 1024);   // A of dV / dK: P^T / dS^T, MN-major (M = keys), LBO = slab
    const uint32_t id_s = instr_desc(BLOCK_Q, C::NKV, 0, 0);             // S, dP
    const uint32_t id_kv = instr_desc(BLOCK_Q, HD, 1, 1);                // dV, dK: A MN-major, B MN-major
    const uint32_t id_dq = instr_desc(BLOCK_Q, HD, 0, 1);                // dQ
    for (int t = 0; t < n_tiles; ++t) {
      const uint32_t ph = t & 1;
      const uint64_t dQt_k = smem_desc(s0 + C::OFF_Q + t * TILE_BYTES, 16, 1024);        // A of S
      const uint64_t dQt_mn = smem_desc(s0 + C::OFF_Q + t * TILE_BYTES, 64 * 128, 1024);  // B of dK (K = q rows, N = d)
      const uint64_t dOt_k = smem_desc(s0 + C::OFF_DO + t * TILE_BYTES, 16, 1024);       // A of dP
      const uint64_t dOt_mn = smem_desc(s0 + C::OFF_DO + t * TILE_BYTES, 64 * 128, 1024);  // B of dV
      if (t > 0) {
        mbar_wait(bar_free, ph ^ 1);  // dQ of tile t-1 drained
        tc_fence_after();
      }
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem_base + C::TM_SDP, dQt_k + (uint64_t)(ks * 2), dK_kmaj + (uint64_t)(ks * 2), id_s, ks > 0 ? 1u : 0u);
        umma_commit(bar_s);
      }
      __syncwarp();
      mbar_wait(bar_p, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        // dV[mt] (+)= P^T dO_t : M = 128 keys of M tile mt (slabs 2mt, 2mt+1), K = 128 query rows in 8 steps of 16
#pragma unroll 1
        for (int mt = 0; mt < C::MT; ++mt)
#pragma unroll 1
          for (int ks = 0; ks < 8; ++ks)
            umma_f16(tmem_base + C::TM_DV + mt * 64, dP_mn + (uint64_t)((mt * 2 * TILE_BYTES + ks * 2048) >> 4),
                     dOt_mn + (uint64_t)((ks * 2048) >> 4), id_kv, (t > 0 || ks > 0) ? 1u : 0u);
        // dP = dO_t V^T into the columns S occupied (every worker has consumed S: bar_p)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem_base + C::TM_SDP, dOt_k + (uint64_t)(ks * 2), dV_kmaj + (uint64_t)(ks * 2), id_s, ks > 0 ? 1u : 0u);
        umma_commit(bar_dp);
      }
      __syncwarp();
      mbar_wait(bar_ds, ph);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll 1
        for (int mt = 0; mt < C::MT; ++mt)
#pragma unroll 1
          for (int ks = 0; ks < 8; ++ks)
            umma_f16(tmem_base + C::TM_DK + mt * 64, dP_mn + (uint64_t)((mt * 2 * TILE_BYTES + ks * 2048) >> 4),
                     dQt_mn + (uint64_t)((ks * 2048) >> 4), id_kv, (t > 0 || ks > 0) ? 1u : 0u);
        // dQ_t = dS K : K = keys in NKV16 steps of 16
#pragma unroll 1
        for (int ks = 0; ks < NKV16; ++ks)
          umma_f16(tmem_base + C::TM_SDP, dP_kmaj + (uint64_t)(((ks >> 2) * TILE_BYTES + (ks & 3) * 32) >> 4),
                   dK_mn + (uint64_t)((ks * 2048) >> 4), id_dq, ks > 0 ? 1u : 0u);
        umma_commit(bar_dq);
      }
      __syncwarp();
    }
  } else {
    // ===================== worker warps =====================
    const int g = warp >> 2, q = warp & 3;   // g: key-column group (and which accumulator / columns this group drains)
    const int r = q * 32 + lane;             // row inside a 128-row tile == TMEM lane
    const uint32_t tL = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sl2 = scale * kLog2e;
    uint8_t* sP = smem + C::OFF_P;
    // per-row constants: base-2 LSE and D_i = sum_d dO O straight from global (one 128-byte row per thread and tensor)
    if (g < C::QT) {
      const int m = g * BLOCK_Q + r;  // this thread prepares row m of the group (rows 0..255)
      float L = 0.f, D = 0.f;
      if (m < rows_valid) {
        const int im = m / N;
        L = lse[((size_t)(b0 + im) * h + head) * N + (m - im * N)] * kLog2e;
        const uint4* po = reinterpret_cast<const uint4*>(outp + (row0 + m) * ld_out + head * HD);
        const uint4* pd = reinterpret_cast<const uint4*>(dout + (row0 + m) * ld_out + head * HD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 a = __ldg(po + i), c = __ldg(pd + i);
          const uint32_t aa[4] = {a.x, a.y, a.z, a.w}, cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 x = unpack_bf16x2(aa[j]), y = unpack_bf16x2(cc[j]);
            D = fmaf(x.x, y.x, D);
            D = fmaf(x.y, y.y, D);
          }
        }
      }
      sL[m] = L;
      sD[m] = D;
    }
    if (threadIdx.x < 192) sC[threadIdx.x] = 0.f;
    named_bar_sync(1, C::WORKERS * 32);
    for (int t = 0; t < n_tiles; ++t) {
      const uint32_t ph = t & 1;
      const int m = t * BLOCK_Q + r;
      const bool row_ok = m < rows_valid;
      const int klo = min(m / N, n_img - 1) * N, khi = klo + N;  // this row's keys: its own image's tokens
      const float L = sL[m], D = sD[m];
      const int nfull = (G == 1) ? N / 16 : 0;  // chunks whose 16 keys are valid for every row (one image per CTA)
      // ---- P = 2^(s*sl2 - L) for my key columns (0 for padded rows / keys) -> smem
      mbar_wait(bar_s, ph);
      tc_fence_after();
      if (g == 0) bwd_p_pass<C::qb(0), C::qb(1)>(tL + C::TM_SDP, sP, r, row_ok, klo, khi, nfull, sl2, L);
      else if (g == 1) bwd_p_pass<C::qb(1), C::qb(2)>(tL + C::TM_SDP, sP, r, row_ok, klo, khi, nfull, sl2, L);
      else if (NG == 4 && g == 2) bwd_p_pass<C::qb(NG == 4 ? 2 : 0), C::qb(NG == 4 ? 3 : 1)>(tL + C::TM_SDP, sP, r, row_ok, klo, khi, nfull, sl2, L);
      else if (NG == 4) bwd_p_pass<C::qb(NG == 4 ? 3 : 0), C::qb(NG == 4 ? 4 : 1)>(tL + C::TM_SDP, sP, r, row_ok, klo, khi, nfull, sl2, L);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      // ---- dS = P (dP - D), bf16, in place of P
      mbar_wait(bar_dp, ph);
      tc_fence_after();
      if (g == 0) bwd_ds_pass<C::qb(0), C::qb(1)>(tL + C::TM_SDP, sP, r, D);
      else if (g == 1) bwd_ds_pass<C::qb(1), C::qb(2)>(tL + C::TM_SDP, sP, r, D);
      else if (NG == 4 && g == 2) bwd_ds_pass<C::qb(NG == 4 ? 2 : 0), C::qb(NG == 4 ? 3 : 1)>(tL + C::TM_SDP, sP, r, D);
      else if (NG == 4) bwd_ds_pass<C::qb(NG == 4 ? 3 : 0), C::qb(NG == 4 ? 4 : 1)>(tL + C::TM_SDP, sP, r, D);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ds);
      // ---- dQ_t: group 0 drains it (NG = 4: groups 0 / 1 take 32 columns each); the others run ahead to the next tile's waits
      if (g == 0 || (NG == 4 && g == 1)) {
        mbar_wait(bar_dq, ph);
        tc_fence_after();
        __nv_bfloat16* grow = dqkv + (row0 + m) * ld_dtok + head * HD;
        if (NG == 2) drain_tile(tL + C::TM_SDP, scale, row_ok, grow, colsum ? sC : nullptr, lane);
        else drain_cols32(tL + C::TM_SDP, g, scale, row_ok, grow, colsum ? sC : nullptr, lane);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free);
      }
    }
    // ---- dK / dV: keys on the lanes, two 128-key M tiles each; the last bar_dq completion covers every MMA issued
    mbar_wait(bar_dq, (n_tiles - 1) & 1);
    tc_fence_after();
    if (NG == 2 && C::MT == 1) {  // one key M-tile: group 0 drains dK, group 1 dV
      const bool ok = r < rows_valid;
      __nv_bfloat16* base = dqkv + (row0 + r) * ld_dtok + head * HD + (size_t)(1 + g) * h * HD;
      drain_tile(tL + (g ? C::TM_DV : C::TM_DK), g ? 1.f : scale, ok, base, colsum ? sC + 64 + 64 * g : nullptr, lane);
    } else if (NG == 2) {  // group g drains M tile g of both
      const int key = g * BLOCK_Q + r;
      const bool ok = key < rows_valid;
      __nv_bfloat16* base = dqkv + (row0 + key) * ld_dtok + head * HD;
      drain_tile(tL + C::TM_DK + g * 64, scale, ok, base + (size_t)h * HD, colsum ? sC + 64 : nullptr, lane);
      drain_tile(tL + C::TM_DV + g * 64, 1.f, ok, base + (size_t)2 * h * HD, colsum ? sC + 128 : nullptr, lane);
    } else {        // groups 0 / 1: dK tiles 0 / 1; groups 2 / 3: dV tiles 0 / 1
      const int mt = g & 1, is_v = g >> 1;
      const int key = mt * BLOCK_Q + r;
      const bool ok = key < rows_valid;
      __nv_bfloat16* base = dqkv + (row0 + key) * ld_dtok + head * HD + (size_t)(1 + is_v) * h * HD;
      drain_tile(tL + (is_v ? C::TM_DV : C::TM_DK) + mt * 64, is_v ? 1.f : scale, ok, base, colsum ? sC + 64 + 64 * is_v : nullptr, lane);
    }
    named_bar_sync(1, C::WORKERS * 32);
    if (colsum && threadIdx.x < 192) {
      const int part = threadIdx.x >> 6, d = threadIdx.x & 63;
      atomicAdd(colsum + (size_t)part * h * HD + head * HD + d, sC[threadIdx.x]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CTRL) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---- host side (driver entry point resolved at run time, as in gemm_tcgen05.cu) --------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && p) fn = (PFN_encodeTiled)p;
  }
  return fn;
}
static int tmap_rows(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return B200_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B200_OK : B200_ERR_DRIVER;
}

template <int NKV16>
static int launch_fwd(const void* qkv, long long ld_tok, int B, int N, int G, int h, float scale, void* out, long long ld_out,
                      float* lse, cudaStream_t s) {
  using C = FwdCfg<NKV16>;
  CUtensorMap tmQ, tmKV;
  int rc = tmap_rows(&tmQ, qkv, (long long)B * N, 3LL * h * HD, ld_tok, BLOCK_Q);
  if (rc) return rc;
  rc = tmap_rows(&tmKV, qkv, (long long)B * N, 3LL * h * HD, ld_tok, C::NKV);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_fwd_tc_kernel<NKV16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
      return B200_ERR_CUDA;
    attr = true;
  }
  launch_kernel(attn_fwd_tc_kernel<NKV16>, ((B + G - 1) / G) * h, C::THREADS, C::SMEM_BYTES, s, tmQ, tmKV, B, N, G, h, scale, (__nv_bfloat16*)out,
                                                                                  ld_out, lse);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

template <int NKV16>
static int launch_fwd7(const void* qkv, long long ld_tok, int B, int N, int G, int h, float scale, void* out, long long ld_out,
                       float* lse, cudaStream_t s) {
  using C = Fwd7Cfg<NKV16>;
  CUtensorMap tmQ, tmKV;
  int rc = tmap_rows(&tmQ, qkv, (long long)B * N, 3LL * h * HD, ld_tok, BLOCK_Q);
  if (rc) return rc;
  rc = tmap_rows(&tmKV, qkv, (long long)B * N, 3LL * h * HD, ld_tok, C::NKV);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_fwd_tc7_kernel<NKV16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
      return B200_ERR_CUDA;
    attr = true;
  }
  const int tiles = (G * N + BLOCK_Q - 1) / BLOCK_Q;  // query tiles per (image group, head): 1 or 2
  launch_kernel(attn_fwd_tc7_kernel<NKV16>, ((B + G - 1) / G) * h * tiles, C::THREADS, C::SMEM_BYTES, s, tmQ, tmKV, B, N, G, h, tiles,
                scale, (__nv_bfloat16*)out, ld_out, lse);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

template <int NKV16>
static int launch_fwd8(const void* qkv, long long ld_tok, int B, int N, int G, int h, float scale, void* out, long long ld_out,
                       float* lse, cudaStream_t s) {
  using C = Fwd7Cfg<NKV16>;
  CUtensorMap tmQ, tmKV;
  int rc = tmap_rows(&tmQ, qkv, (long long)B * N, 3LL * h * HD, ld_tok, BLOCK_Q);
  if (rc) return rc;
  rc = tmap_rows(&tmKV, qkv, (long long)B * N, 3LL * h * HD, ld_tok, C::NKV);
  if (rc) return rc;
  static bool attr = false;
  static int num_sms = 0;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_fwd_tc8_kernel<NKV16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
      return B200_ERR_CUDA;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  const int tiles = (G * N + BLOCK_Q - 1) / BLOCK_Q;
  const int total = ((B + G - 1) / G) * h * tiles;
  const int grid = total < 2 * num_sms ? total : 2 * num_sms;
  launch_kernel(attn_fwd_tc8_kernel<NKV16>, grid, C::THREADS, C::SMEM_BYTES, s, tmQ, tmKV, B, N, G, h, tiles, total, scale,
                (__nv_bfloat16*)out, ld_out, lse);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

template <int NKV16, int NG>
static int launch_bwd(const void* qkv, long long ld_tok, const void* out, const void* dout, long long ld_out, const float* lse, int B,
                      int N, int G, int h, float scale, void* dqkv, long long ld_dtok, float* colsum, cudaStream_t s) {
  using C = BwdCfg<NKV16, NG>;
  CUtensorMap tmQ, tmKV, tmDO;
  int rc = tmap_rows(&tmQ, qkv, (long long)B * N, 3LL * h * HD, ld_tok, BLOCK_Q);
  if (rc) return rc;
  rc = tmap_rows(&tmKV, qkv, (long long)B * N, 3LL * h * HD, ld_tok, C::NKV);
  if (rc) return rc;
  rc = tmap_rows(&tmDO, dout, (long long)B * N, (long long)h * HD, ld_out, BLOCK_Q);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(attn_bwd_tc_kernel<NKV16, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess)
      return B200_ERR_CUDA;
    attr = true;
  }
  launch_kernel(attn_bwd_tc_kernel<NKV16, NG>, ((B + G - 1) / G) * h, C::THREADS, C::SMEM_BYTES, s, tmQ, tmKV, tmDO, (const __nv_bfloat16*)out,
                                                                                  (const __nv_bfloat16*)dout, ld_out, lse, B, N, G, h,
                                                                                  scale, (__nv_bfloat16*)dqkv, ld_dtok, colsum);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace attn_tc
}  // namespace b200

// Forward on tcgen05.  128 < N <= 256 (shorter sequences: b200_attention_fwd routes them to the warp-level kernel).
extern "C" int b200_attention_fwd_tc(const void* qkv, long long ld_tok, int B, int N, int h, int head_dim, float scale, void* out,
                                     long long ld_out, float* lse, void* stream) {
  using namespace b200::attn_tc;
  if (!qkv || !out || B <= 0 || N <= 0 || h <= 0 || !(scale > 0.f)) return B200_ERR_INVALID_ARG;
  if (head_dim != HD || (ld_tok % 8) || (ld_out % 8) || ((uintptr_t)qkv & 15) || ((uintptr_t)out & 15)) return B200_ERR_UNSUPPORTED;
  if (N > 256) return B200_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  // B200_ATTN_FWD_SCHEDULE: 8 (default) P in tensor memory, persistent CTAs (two per SM) with the next tile's Q / K / V
  // prefetched: 41 us at 768 pairs x 197 tokens; 7 the same without persistence (45 us); 1 the first tcgen05 version: both tiles
  // per CTA, P through swizzled shared memory (62 us).  All numerically identical (tools/attn_check.py).  Three more schedules
  // that carried P through shared memory (one tile per CTA x two CTAs per SM: 68 us; persistent with TMA prefetch: 72 us;
  // sixteen softmax warps: 57 us) were measured and removed: profiles/r02_attn_*time*.log, DESIGN.md 6.3.
  static int sched = -1;
  if (sched < 0) {
    const char* e = std::getenv("B200_ATTN_FWD_SCHEDULE");
    sched = (e && (e[0] == '1' || e[0] == '7' || e[0] == '8')) ? e[0] - '0' : 8;
  }
  if (sched == 8) {
    if (N <= 128) return launch_fwd8<8>(qkv, ld_tok, B, N, 128 / N, h, scale, out, ld_out, lse, s);
    if (N <= 208) return launch_fwd8<13>(qkv, ld_tok, B, N, 1, h, scale, out, ld_out, lse, s);
    return launch_fwd8<16>(qkv, ld_tok, B, N, 1, h, scale, out, ld_out, lse, s);
  }
  if (sched == 7) {
    if (N <= 128) return launch_fwd7<8>(qkv, ld_tok, B, N, 128 / N, h, scale, out, ld_out, lse, s);
    if (N <= 208) return launch_fwd7<13>(qkv, ld_tok, B, N, 1, h, scale, out, ld_out, lse, s);
    return launch_fwd7<16>(qkv, ld_tok, B, N, 1, h, scale, out, ld_out, lse, s);
  }
  // N <= 128: as many whole images as fit one 128-row tile share a CTA (local crops: 3 x 37 tokens), keys padded to 128
  if (N <= 128) return launch_fwd<8>(qkv, ld_tok, B, N, 128 / N, h, scale, out, ld_out, lse, s);
  if (N <= 208) return launch_fwd<13>(qkv, ld_tok, B, N, 1, h, scale, out, ld_out, lse, s);
  return launch_fwd<16>(qkv, ld_tok, B, N, 1, h, scale, out, ld_out, lse, s);
}

// Backward on tcgen05.  N <= 208.
extern "C" int b200_attention_bwd_tc(const void* qkv, long long ld_tok, const void* out, const void* dout, long long ld_out,
                                     const float* lse, int B, int N, int h, int head_dim, float scale, void* dqkv,
                                     long long ld_dtok, float* dqkv_colsum, void* stream) {
  using namespace b200::attn_tc;
  if (!qkv || !out || !dout || !lse || !dqkv || B <= 0 || N <= 0 || h <= 0 || !(scale > 0.f)) return B200_ERR_INVALID_ARG;
  if (head_dim != HD || (ld_tok % 8) || (ld_out % 8) || (ld_dtok % 8)) return B200_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 15) || ((uintptr_t)dout & 15) || ((uintptr_t)dqkv & 15)) return B200_ERR_UNSUPPORTED;
  if (N > 208) return B200_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  static int groups = -1;  // B200_ATTN_BWD_GROUPS = 4 (default; 16 worker warps: each row's keys split four ways) | 2 (8 worker warps)
  if (groups < 0) {
    const char* e = std::getenv("B200_ATTN_BWD_GROUPS");
    groups = (e && e[0] == '2') ? 2 : 4;
  }
  // N <= 128 (packed short sequences): one tile per CTA, 9 warps, two CTAs per SM
  if (N <= 128) return launch_bwd<8, 2>(qkv, ld_tok, out, dout, ld_out, lse, B, N, 128 / N, h, scale, dqkv, ld_dtok, dqkv_colsum, s);
  if (groups == 4) {
    return launch_bwd<13, 4>(qkv, ld_tok, out, dout, ld_out, lse, B, N, 1, h, scale, dqkv, ld_dtok, dqkv_colsum, s);
  }
  return launch_bwd<13, 2>(qkv, ld_tok, out, dout, ld_out, lse, B, N, 1, h, scale, dqkv, ld_dtok, dqkv_colsum, s);
}
