// b200_gemm: C[M,N] = A[M,K] * B[N,K]^T on tcgen05 tensor cores (sm_100a).
//
// Replaces every nn.Linear / conv-as-GEMM on the DINOv2 hot path of the reference
// (qkv/proj: layers/attention.py:44-63, Mlp fc1/fc2: layers/mlp.py:31-42, PatchEmbed.proj:
// layers/patch_embed.py:77-110, DINOv2ProjectionHead: _methods/dinov2/dinov2_head.py:66-95)
// and their autograd backward (dgrad / wgrad), which the reference leaves to cuBLASLt.
//
// Design (one persistent CTA per SM, warp-specialised):
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      : tcgen05.mma issuer (single thread), accumulators in TMEM, double-buffered
//   warps 2..9  : epilogue (tcgen05.ld -> registers -> fused bias/GELU/LayerScale+residual/
//                 dGELU/atomic split-K) with vectorised global stores
// bf16 operands, fp32 accumulate. Either operand may be K-major (row-major [rows,K]) or
// MN-major (stored [K,rows]); the latter feeds wgrad (contraction over tokens) without
// materialising transposes.
#include <cuda.h>
#include "common.cuh"
#include <type_traits>
#include "../../include/b200dino.h"

namespace b200 {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
static constexpr int UMMA_K = 16;
// Epilogue warps: TMEM lane quarter q is only reachable from warps with (warp % 4) == q, so the epilogue has a
// multiple of 4 warps; each owns 32 rows x (BLOCK_N / (warps/4)) columns.  The fused epilogues are issue/latency
// bound, so the 192-wide tile (every D=384 GEMM) gets 12 warps x 64 columns; 128 -> 8 x 64, 256 -> 8 x 128
// (more warps would cap registers below what the aux-prefetching epilogues need).
template <int BLOCK_N>
struct EpiCfg {
  static constexpr int WARPS = (BLOCK_N == 192) ? 12 : 8;
  static constexpr int COLS_PER_WARP = BLOCK_N / (WARPS / 4);
  static constexpr int THREADS = (2 + WARPS) * 32;
  static constexpr int STAGING_BYTES = WARPS * 32 * 32 * 4;  // one 32x32 fp32 transpose tile per warp
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_STAGING_BYTES = EpiCfg<BLOCK_N>::STAGING_BYTES;
  static constexpr int STAGES_FIT = (232448 - 1024 - 256 - EPI_STAGING_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr int ACC_STRIDE = 256;  // TMEM columns between the two accumulator buffers
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmDev {
  int M, N, K;
  int a_mn, b_mn;
  int splits;
  int epi;
  float alpha;
  void* C;
  long long ldc;
  void* C2;
  long long ldc2;
  const void* aux;
  long long ldaux;
  const float* bias;
  const float* gamma;
  const float* rowscale;
  int rows_per_scale;
};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30),
  // SBO>>4 [32,46), version=1 [46,48), layout_type=SWIZZLE_128B(2) [61,64)
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn, int b_mn, int m = BLOCK_M) {
  // cute::UMMA::InstrDescriptor: c_format F32(1)@[4,6), a/b_format BF16(1)@[7,10)/[10,13),
  // a_major@15, b_major@16, N>>3@[17,23), M>>4@[24,29)
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (uint32_t)(a_mn ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn ? 1 : 0) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

// ---------------------------------------------------------------------------------------------
// Epilogue.  After an smem transpose a warp instruction covers 4 rows x 32 columns (full 128-byte lines of fp32,
// 64-byte runs of bf16): coalesced loads and stores.  EPI is a template parameter (branch-free per element) and
// all global addresses are strength-reduced to one pointer bump per row group: the epilogue is issue-bound,
// every instruction per element counts.
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

template <int EPI, bool HAS_C2>
__device__ __forceinline__ void epilogue_vec4(float alpha, float4 acc, char* c, char* c2_, float rs, const float4& bias4,
                                              const float4& gamma4, const float4& aux4) {
  char* const c2 = HAS_C2 ? c2_ : nullptr;  // compile-time null: no per-store pointer test
  const float v0 = fmaf(acc.x, alpha, bias4.x), v1 = fmaf(acc.y, alpha, bias4.y), v2 = fmaf(acc.z, alpha, bias4.z),
              v3 = fmaf(acc.w, alpha, bias4.w);
  if constexpr (EPI == B200_EPI_BF16) {
    *reinterpret_cast<uint2*>(c) = make_uint2(pack_bf16x2(v0, v1), pack_bf16x2(v2, v3));
  } else if constexpr (EPI == B200_EPI_F32) {
    *reinterpret_cast<float4*>(c) = make_float4(v0, v1, v2, v3);
  } else if constexpr (EPI == B200_EPI_F32_ATOMIC) {
    atomicAdd(reinterpret_cast<float4*>(c), make_float4(v0, v1, v2, v3));
  } else if constexpr (EPI == B200_EPI_BIAS_GELU) {
    // u = bf16(acc + bias) (what nn.Linear returns under bf16 autocast); h = bf16(gelu(u))
    const uint32_t p01 = pack_bf16x2(v0, v1), p23 = pack_bf16x2(v2, v3);
    if (c2) *reinterpret_cast<uint2*>(c2) = make_uint2(p01, p23);
    const float2 u01 = unpack_bf16x2(p01), u23 = unpack_bf16x2(p23);
    *reinterpret_cast<uint2*>(c) = make_uint2(pack_bf16x2(gelu_erf(u01.x), gelu_erf(u01.y)), pack_bf16x2(gelu_erf(u23.x), gelu_erf(u23.y)));
  } else if constexpr (EPI == B200_EPI_BIAS_GELU_DG) {
    // like BIAS_GELU, but the second output is gelu'(u) (bf16) instead of u: the backward then only multiplies
    const uint32_t p01 = pack_bf16x2(v0, v1), p23 = pack_bf16x2(v2, v3);
    const float2 u01 = unpack_bf16x2(p01), u23 = unpack_bf16x2(p23);
    const float h0 = gelu_tail(u01.x), h1 = gelu_tail(u01.y), h2 = gelu_tail(u23.x), h3 = gelu_tail(u23.y);
    *reinterpret_cast<uint2*>(c) = make_uint2(pack_bf16x2(gelu_from_tail(u01.x, h0), gelu_from_tail(u01.y, h1)),
                                              pack_bf16x2(gelu_from_tail(u23.x, h2), gelu_from_tail(u23.y, h3)));
    if (c2)
      *reinterpret_cast<uint2*>(c2) = make_uint2(
          pack_bf16x2(gelu_grad_from_phi(u01.x, phi_from_tail(u01.x, h0)), gelu_grad_from_phi(u01.y, phi_from_tail(u01.y, h1))),
          pack_bf16x2(gelu_grad_from_phi(u23.x, phi_from_tail(u23.x, h2)), gelu_grad_from_phi(u23.y, phi_from_tail(u23.y, h3))));
  } else if constexpr (EPI == B200_EPI_MUL_AUX) {
    // C = bf16( bf16(acc) * aux ), aux bf16 (e.g. the gelu'(u) saved by BIAS_GELU_DG)
    const float2 ga = unpack_bf16x2(__float_as_uint(aux4.x)), gc = unpack_bf16x2(__float_as_uint(aux4.y));
    // (rounding through the packed F2FP: the scalar F2F.BF16 conversion runs on the quarter-rate XU pipe)
    const float2 r01 = unpack_bf16x2(pack_bf16x2(v0, v1)), r23 = unpack_bf16x2(pack_bf16x2(v2, v3));
    *reinterpret_cast<uint2*>(c) = make_uint2(pack_bf16x2(r01.x * ga.x, r01.y * ga.y), pack_bf16x2(r23.x * gc.x, r23.y * gc.y));
  } else if constexpr (EPI == B200_EPI_RESIDUAL) {
    // o = bf16(acc + bias); x_out = x_in + gamma * o * rowscale   (fp32 residual stream)
    const uint32_t p01 = pack_bf16x2(v0, v1), p23 = pack_bf16x2(v2, v3);
    if (c2) *reinterpret_cast<uint2*>(c2) = make_uint2(p01, p23);
    const float2 o01 = unpack_bf16x2(p01), o23 = unpack_bf16x2(p23);
    float4 x = aux4;
    x.x += (o01.x * gamma4.x) * rs; x.y += (o01.y * gamma4.y) * rs; x.z += (o23.x * gamma4.z) * rs; x.w += (o23.y * gamma4.w) * rs;
    *reinterpret_cast<float4*>(c) = x;
  } else if constexpr (EPI == B200_EPI_DGELU) {
    // dU = bf16( bf16(acc) * gelu'(u) ), u = aux (bf16 pre-activation saved by the forward)
    const float2 ua = unpack_bf16x2(__float_as_uint(aux4.x)), uc = unpack_bf16x2(__float_as_uint(aux4.y));
    const float2 r01 = unpack_bf16x2(pack_bf16x2(v0, v1)), r23 = unpack_bf16x2(pack_bf16x2(v2, v3));
    *reinterpret_cast<uint2*>(c) = make_uint2(pack_bf16x2(r01.x * gelu_erf_grad(ua.x), r01.y * gelu_erf_grad(ua.y)),
                                              pack_bf16x2(r23.x * gelu_erf_grad(uc.x), r23.y * gelu_erf_grad(uc.y)));
  }
}

// global reads an epilogue needs for one 32-column chunk (8 row groups per lane): the fp32 residual stream or the
// saved bf16 pre-activation.  They never alias the outputs, so they are issued a whole chunk ahead.
template <int EPI>
__device__ __forceinline__ void load_aux_chunk(float4 (&aux4)[8], const char* aux, long long step, int rows_valid, int sub_row, bool col_ok) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const bool ok = col_ok && (it * 4 + sub_row < rows_valid);
    if constexpr (EPI == B200_EPI_RESIDUAL) {
      aux4[it] = ok ? __ldg(reinterpret_cast<const float4*>(aux + it * step)) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else if constexpr (EPI == B200_EPI_DGELU || EPI == B200_EPI_MUL_AUX) {
      const uint2 u = ok ? __ldg(reinterpret_cast<const uint2*>(aux + it * step)) : make_uint2(0u, 0u);
      aux4[it] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), 0.f, 0.f);
    }
  }
}

// The epilogue role of one warp, for the whole persistent loop over work items w = w_begin, w_begin+w_step, ...
// (tile = w / n_splits, m-tile = tile % tiles_m, n-tile = tile / tiles_m in both schedules).
// Per tile: TMEM -> registers -> smem transpose -> fused epilogue -> global.
//  * __noinline__ + by-value arguments: each EPI variant gets its own register allocation and reads no kernel
//    parameter from memory inside the loop;
//  * bias / gamma live in registers across tiles and are re-read only when the n-tile changes;
//  * software pipeline: the first chunk's aux reads are issued before the accumulator is ready; tcgen05.ld and the
//    aux reads of chunk c+1 are in flight while chunk c is stored.
struct EpiSched {
  int w_begin, w_end, w_step, n_splits, tiles_m, tiles_n;  // tiles_n > 0: n-fastest tile order (generic kernel)
  int m_stride = BLOCK_M, m_off = 0;                       // CTA-pair kernel: 256-row tiles, this CTA's half at m_off
  // LayerNorm-prologue kernel: the CTA walks whole units (row panels) of `unit_tiles` consecutive tiles; its k-th unit is
  // unit_first + k * unit_stride and w counts the CTA's own tiles 0, 1, 2, ...
  int unit_tiles = 0, unit_first = 0, unit_stride = 0;
  __device__ __forceinline__ int tile_of(int w) const {
    return unit_tiles > 0 ? (unit_first + (w / unit_tiles) * unit_stride) * unit_tiles + (w % unit_tiles) : w / n_splits;
  }
};

template <int BLOCK_N, int EPI, bool TWO_SM = false>
__device__ __forceinline__ void epilogue_role(const GemmDev& p, const EpiSched& sc, const uint32_t tmem_base, uint64_t* tmem_full,
                                              uint64_t* tmem_empty, const int quarter, const int half, const int lane, float* stg) {
  constexpr int COLS_PER_WARP = EpiCfg<BLOCK_N>::COLS_PER_WARP;
  constexpr int NC = COLS_PER_WARP / 32;
  constexpr bool HAS_AUX = (EPI == B200_EPI_RESIDUAL || EPI == B200_EPI_DGELU || EPI == B200_EPI_MUL_AUX);
  constexpr int C_ESIZE = (EPI == B200_EPI_F32 || EPI == B200_EPI_F32_ATOMIC || EPI == B200_EPI_RESIDUAL) ? 4 : 2;
  constexpr int AUX_ESIZE = (EPI == B200_EPI_RESIDUAL) ? 4 : 2;
  constexpr int ACC_STRIDE = 256;
  // kernel parameters -> registers once (this function is not inlined: `p` arrives through the stack)
  const int M = p.M, N = p.N, rows_per_scale = p.rows_per_scale;
  const float alpha = p.alpha;
  char* const C = reinterpret_cast<char*>(p.C);
  char* const C2 = reinterpret_cast<char*>(p.C2);
  const char* const AUX = reinterpret_cast<const char*>(p.aux);
  const float* const bias = p.bias;
  const float* const gamma = p.gamma;
  const float* const rowscale = p.rowscale;
  const long long ldc = p.ldc, ldc2 = p.ldc2, ldaux = p.ldaux;
  const int sub_row = lane >> 3;  // 0..3 : row within a 4-row group after the transpose
  const int g4 = lane & 7;        // 0..7 : which float4 (4 columns) of the 32-column chunk
  const long long c_step = 4 * ldc * C_ESIZE, c2_step = 8 * ldc2, aux_step = 4 * ldaux * AUX_ESIZE;
  const uint32_t stg_w = smem_u32(stg) + lane * 128;     // this lane's row (write side of the transpose)
  const uint32_t stg_r = smem_u32(stg) + sub_row * 128;  // first row of the read side
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int w = sc.w_begin; w < sc.w_end; w += sc.w_step) {
    const int tile = sc.tile_of(w);
    const int m0 = (sc.tiles_n > 0 ? tile / sc.tiles_n : tile % sc.tiles_m) * sc.m_stride + sc.m_off;
    const int n0 = (sc.tiles_n > 0 ? tile % sc.tiles_n : tile / sc.tiles_m) * BLOCK_N;
    const int row_first = m0 + quarter * 32 + sub_row;
    const int rows_valid = M - (m0 + quarter * 32);  // rows r (0..31) of this warp are valid iff r < rows_valid
    const int col0 = n0 + half * COLS_PER_WARP + g4 * 4;
    // bias / gamma of the first chunk (later chunks are fetched one chunk ahead inside the loop)
    float4 bias_nxt = (bias && col0 < N) ? __ldg(reinterpret_cast<const float4*>(bias + col0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 gamma_nxt = (EPI == B200_EPI_RESIDUAL && gamma && col0 < N) ? __ldg(reinterpret_cast<const float4*>(gamma + col0))
                                                                        : make_float4(1.f, 1.f, 1.f, 1.f);
    char* c_base = C + ((size_t)row_first * ldc + col0) * C_ESIZE;
    char* c2_base = C2 ? C2 + ((size_t)row_first * ldc2 + col0) * 2 : nullptr;
    const char* aux_base = HAS_AUX ? AUX + ((size_t)row_first * ldaux + col0) * AUX_ESIZE : nullptr;
    float rs[8];
    if constexpr (EPI == B200_EPI_RESIDUAL) {
#pragma unroll
      for (int it = 0; it < 8; ++it)
        rs[it] = (rowscale && it * 4 + sub_row < rows_valid) ? __ldg(rowscale + (row_first + it * 4) / rows_per_scale) : 1.0f;
    }
    float4 aux_cur[8], aux_nxt[8];
    if constexpr (HAS_AUX) load_aux_chunk<EPI>(aux_nxt, aux_base, aux_step, rows_valid, sub_row, col0 < N);
    if constexpr (HAS_AUX) {
      // the NEXT tile's slice of aux (32 rows x COLS_PER_WARP of this warp) -> L2: aux is read exactly once, and with
      // one chunk per warp in flight the fp32 residual stream was latency-bound (profiles/r01_gemm_stalls.md)
      const int wn = w + sc.w_step;
      if (wn < sc.w_end) {
        const int tn = sc.tile_of(wn);
        const int m0n = (sc.tiles_n > 0 ? tn / sc.tiles_n : tn % sc.tiles_m) * sc.m_stride + sc.m_off + quarter * 32;
        const int n0n = (sc.tiles_n > 0 ? tn % sc.tiles_n : tn / sc.tiles_m) * BLOCK_N + half * COLS_PER_WARP;
        constexpr int LINES = COLS_PER_WARP * AUX_ESIZE / 128;  // 128-byte lines per row slice (>= 1)
#pragma unroll
        for (int i = 0; i < LINES; ++i) {
          const int coln = n0n + i * (128 / AUX_ESIZE);
          if (m0n + lane < M && coln < N) prefetch_l2(AUX + ((size_t)(m0n + lane) * ldaux + coln) * AUX_ESIZE);
        }
      }
    }

    mbar_wait(&tmem_full[acc], acc_phase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(quarter * 32) << 16) + half * COLS_PER_WARP;
    uint32_t v[32];
    tmem_ld_32x32(taddr, v);
    // the chunk loop is deliberately NOT unrolled: the unrolled 8-row-group body below is already 1-2 K
    // instructions for the GELU epilogues, and unrolling it NC times thrashed the instruction cache
    // (stall_no_inst ~ 19 % of samples in profiles/r01_gemm_stalls.md)
#pragma unroll 1
    for (int c = 0; c < NC; ++c) {
      const float4 bias4 = bias_nxt, gamma4 = gamma_nxt;
      tmem_ld_wait();
      // transpose through smem: thread (= row `lane`) writes its 32 columns; XOR swizzle keeps both the
      // row-wise 16-byte writes and the column-group reads bank-conflict free
#pragma unroll
      for (int j = 0; j < 8; ++j) sts128(stg_w + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      const int col_next = col0 + (c + 1) * 32;
      if (c + 1 < NC) {
        tmem_ld_32x32(taddr + (c + 1) * 32, v);
        if (bias && col_next < N) bias_nxt = __ldg(reinterpret_cast<const float4*>(bias + col_next));
        if (EPI == B200_EPI_RESIDUAL && gamma && col_next < N) gamma_nxt = __ldg(reinterpret_cast<const float4*>(gamma + col_next));
      }
      __syncwarp();
      const bool col_ok = col0 + c * 32 < N;  // N % 8 == 0 and col % 4 == 0 -> the 4 columns are all valid
      if constexpr (HAS_AUX) {
#pragma unroll
        for (int it = 0; it < 8; ++it) aux_cur[it] = aux_nxt[it];
        if (c + 1 < NC) load_aux_chunk<EPI>(aux_nxt, aux_base + (c + 1) * 32 * AUX_ESIZE, aux_step, rows_valid, sub_row, col_next < N);
      }
      if (col_ok) {
        // the 8 row groups of this chunk, specialised at compile time on (second output present, all 32 rows valid):
        // per-store pointer tests and row predicates cost ~2 instructions per element in the issue-bound epilogues
        auto rows = [&](auto has_c2, auto full) {
          constexpr bool HC2 = decltype(has_c2)::value, FULL = decltype(full)::value;
          char* cp = c_base + c * 32 * C_ESIZE;
          char* c2p = HC2 ? c2_base + c * 64 : nullptr;
          // the transposed accumulators are read back in batches: the ld.shared are volatile (ordered against the
          // st.shared of the transpose), so reading one row group at a time exposed the full smem latency 8 times per
          // chunk (short-scoreboard stalls were 40 % of the epilogue's samples in profiles/r01_gemm_stalls.md)
          constexpr int BATCH = (EPI == B200_EPI_RESIDUAL) ? 4 : 8;
#pragma unroll
          for (int b0 = 0; b0 < 8; b0 += BATCH) {
            float4 a4[BATCH];
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
              const int r = (b0 + i) * 4 + sub_row;
              a4[i] = lds128(stg_r + (b0 + i) * 512 + ((g4 ^ (r & 7)) << 4));
            }
#pragma unroll
            for (int i = 0; i < BATCH; ++i) {
              const int it = b0 + i;
              const int r = it * 4 + sub_row;
              if (FULL || r < rows_valid)
                epilogue_vec4<EPI, HC2>(alpha, a4[i], cp, c2p, (EPI == B200_EPI_RESIDUAL) ? rs[it] : 1.0f, bias4, gamma4, aux_cur[it]);
              cp += c_step;
              if (HC2) c2p += c2_step;
            }
          }
        };
        if (rows_valid >= 32) {
          if (c2_base) rows(std::true_type{}, std::true_type{}); else rows(std::false_type{}, std::true_type{});
        } else {
          if (c2_base) rows(std::true_type{}, std::false_type{}); else rows(std::false_type{}, std::false_type{});
        }
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (TWO_SM) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));  // the leader CTA's barrier
      else mbar_arrive(&tmem_empty[acc]);
    }
    if ((acc ^= 1) == 0) acc_phase ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------
template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(EpiCfg<BLOCK_N>::THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmDev p) {
  pdl_launch_dependents();  // the wait follows the barrier / TMEM set-up below (no global memory touched before it)
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // swizzle-128B operand tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi_staging = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb_per_split = (kb_total + p.splits - 1) / p.splits;
  const int n_splits = (kb_total + kb_per_split - 1) / kb_per_split;  // effective, no empty splits
  const int total_work = tiles_m * tiles_n * n_splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], EpiCfg<BLOCK_N>::WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int split = w % n_splits;
      const int tile = w / n_splits;
      // n-fastest: the CTAs running concurrently share the same few A row-panels through L2 (A is the big operand;
      // with m-fastest order a K=1536 A matrix was re-read from DRAM once per n-tile: 188 MB instead of 116 MB)
      const int m0 = (tile / tiles_n) * BLOCK_M;
      const int n0 = (tile % tiles_n) * BLOCK_N;
      const int kb0 = split * kb_per_split;
      const int kb1 = min(kb_total, kb0 + kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          if (!p.a_mn) {
            tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int a = 0; a < BLOCK_M / 64; ++a)
              tma_load_2d(sa + a * (BLOCK_K * 128), &tmA, &full_bar[stage], m0 + a * 64, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int a = 0; a < BLOCK_N / 64; ++a)
              tma_load_2d(sb + a * (BLOCK_K * 128), &tmB, &full_bar[stage], n0 + a * 64, k0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // every operand below is warp-uniform; one elected lane issues (keeps descriptors in uniform registers)
    const uint32_t idesc = make_idesc(BLOCK_N, p.a_mn, p.b_mn);
    // K-major:  SBO = 8 rows * 128 B;                         K-step = 32 B inside the swizzle atom
    // MN-major: LBO = BLOCK_K*128 B between 64-wide MN atoms, SBO = 8 k-rows * 128 B, K-step = 16 k-rows * 128 B
    const uint32_t smem0 = smem_u32(smem);
    const uint64_t adesc_base = p.a_mn ? make_smem_desc(smem0, BLOCK_K * 128, 1024) : make_smem_desc(smem0, 16, 1024);
    const uint64_t bdesc_base = p.b_mn ? make_smem_desc(smem0 + Cfg::A_BYTES, BLOCK_K * 128, 1024)
                                       : make_smem_desc(smem0 + Cfg::A_BYTES, 16, 1024);
    const uint32_t a_step = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
    const uint32_t b_step = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int split = w % n_splits;
      const int kb0 = split * kb_per_split;
      const int kb1 = min(kb_total, kb0 + kb_per_split);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_STRIDE;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t ad = adesc_base + (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
        const uint64_t bd = bdesc_base + (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
        if (elect_one_sync()) {
          umma_f16(tmem_d, ad, bd, idesc, (kb > kb0) ? 1u : 0u);
          umma_f16(tmem_d, ad + a_step, bd + b_step, idesc, 1u);
          umma_f16(tmem_d, ad + 2 * a_step, bd + 2 * b_step, idesc, 1u);
          umma_f16(tmem_d, ad + 3 * a_step, bd + 3 * b_step, idesc, 1u);
          umma_commit(&empty_bar[stage]);  // frees the smem slot when the MMAs above retire
          if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if ((acc ^= 1) == 0) acc_phase ^= 1;
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    EpiSched sc{(int)blockIdx.x, total_work, (int)gridDim.x, n_splits, tiles_m, tiles_n};
    epilogue_role<BLOCK_N, EPI>(p, sc, tmem_base, tmem_full, tmem_empty, warp & 3, ew >> 2, lane, epi_staging + ew * (32 * 32));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cluster of 2, tcgen05 cta_group::2): one 256 x BLOCK_N output tile per pair.  Each CTA stages its
// own 128 rows of A and HALF of the B tile (BLOCK_N/2 rows); the leader CTA (cluster rank 0) issues M=256 MMAs that read
// both CTAs' shared memory and write each CTA's 128 accumulator rows into its own TMEM.  Operand bytes pulled through
// L2 per MAC drop by 1/3 (BLOCK_N = 256: 64 MAC/B against 42.7 for the single-CTA 128 x 256 tile) -- the K <= 1536 GEMMs
// of this model run at the L2 throughput cap (~6300 B/clk), not at the MMA rate (profiles/r01_gemm_sweep.log).
//   full[s]       : leader's barrier; the leader arms it for the bytes of BOTH CTAs, both CTAs' TMA complete on it
//   empty[s]      : one per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tmem_full[a]  : one per CTA (commit multicast); tmem_empty[a]: leader's, counts the epilogue warps of both CTAs
template <int BLOCK_N>
struct Gemm2Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;  // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_STAGING_BYTES = EpiCfg<BLOCK_N>::STAGING_BYTES;
  static constexpr int STAGES_FIT = (232448 - 1024 - 256 - EPI_STAGING_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int ACC_STRIDE = 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING_BYTES + 1024 + 256;
};

template <int BLOCK_N, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EpiCfg<BLOCK_N>::THREADS, 1)
gemm_tcgen05_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const GemmDev p) {
  pdl_launch_dependents();  // the wait follows the barrier / TMEM set-up below (no global memory touched before it)
  using Cfg = Gemm2Cfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi_staging = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  const int tiles_m = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb_per_split = (kb_total + p.splits - 1) / p.splits;
  const int n_splits = (kb_total + kb_per_split - 1) / kb_per_split;
  const int total_work = tiles_m * tiles_n * n_splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * EpiCfg<BLOCK_N>::WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_base_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits and TMEM allocations of both CTAs are visible before any remote signal
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    const int m_off = rank * BLOCK_M, n_off = rank * (BLOCK_N / 2);
    for (int w = pair; w < total_work; w += n_pairs) {
      const int split = w % n_splits;
      const int tile = w / n_splits;
      const int m0 = (tile / tiles_n) * (2 * BLOCK_M) + m_off;
      const int n0 = (tile % tiles_n) * BLOCK_N + n_off;
      const int kb0 = split * kb_per_split;
      const int kb1 = min(kb_total, kb0 + kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          const int k0 = kb * BLOCK_K;
          if (!p.a_mn) {
            tma_load_2d_2sm(sa, &tmA, bar, k0, m0);
          } else {
#pragma unroll
            for (int a = 0; a < BLOCK_M / 64; ++a) tma_load_2d_2sm(sa + a * (BLOCK_K * 128), &tmA, bar, m0 + a * 64, k0);
          }
          if (!p.b_mn) {
            tma_load_2d_2sm(sb, &tmB, bar, k0, n0);
          } else {
#pragma unroll
            for (int a = 0; a < BLOCK_N / 128; ++a) tma_load_2d_2sm(sb + a * (BLOCK_K * 128), &tmB, bar, n0 + a * 64, k0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      const uint32_t idesc = make_idesc(BLOCK_N, p.a_mn, p.b_mn, 2 * BLOCK_M);
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t adesc_base = p.a_mn ? make_smem_desc(smem0, BLOCK_K * 128, 1024) : make_smem_desc(smem0, 16, 1024);
      const uint64_t bdesc_base = p.b_mn ? make_smem_desc(smem0 + Cfg::A_BYTES, BLOCK_K * 128, 1024)
                                         : make_smem_desc(smem0 + Cfg::A_BYTES, 16, 1024);
      const uint32_t a_step = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      const uint32_t b_step = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = pair; w < total_work; w += n_pairs) {
        const int split = w % n_splits;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb_total, kb0 + kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t ad = adesc_base + (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
          const uint64_t bd = bdesc_base + (uint64_t)((stage * Cfg::STAGE_BYTES) >> 4);
          if (elect_one_sync()) {
            umma_f16_2sm(tmem_d, ad, bd, idesc, (kb > kb0) ? 1u : 0u);
            umma_f16_2sm(tmem_d, ad + a_step, bd + b_step, idesc, 1u);
            umma_f16_2sm(tmem_d, ad + 2 * a_step, bd + 2 * b_step, idesc, 1u);
            umma_f16_2sm(tmem_d, ad + 3 * a_step, bd + 3 * b_step, idesc, 1u);
            umma_commit_2sm(&empty_bar[stage], 3);  // frees the slot in both CTAs when the MMAs above retire
            if (kb == kb1 - 1) umma_commit_2sm(&tmem_full[acc], 3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int ew = warp - 2;
    EpiSched sc{pair, total_work, n_pairs, n_splits, tiles_m, tiles_n, 2 * BLOCK_M, rank * BLOCK_M};
    epilogue_role<BLOCK_N, EPI, true>(p, sc, tmem_base, tmem_full, tmem_empty, warp & 3, ew >> 2, lane, epi_staging + ew * (32 * 32));
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA of the pair leaves (or frees TMEM) while the other may still signal or read it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// Weight-stationary variant for small K (K <= WS_KB_MAX*64, e.g. every D=384 GEMM of the ViT):
// the [BLOCK_N x K] slab of B stays resident in shared memory while the CTA streams consecutive M tiles of A
// through a short ring.  The generic kernel re-fetches the B slab for every output tile, which makes the
// K=384 GEMMs L2->SM bandwidth bound (measured ~7 TB/s at 40 % tensor utilisation); here each CTA owns a
// contiguous run of the n-major tile list, so B is fetched once (at most twice) per CTA.
template <int BLOCK_N, int KB_MAX>
struct GemmWsCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;      // one k-block of the resident slab
  static constexpr int PANEL_BYTES = KB_MAX * B_BYTES;
  static constexpr int EPI_STAGING_BYTES = EpiCfg<BLOCK_N>::STAGING_BYTES;
  static constexpr int BUDGET = 232448 - 1024 - 256 - PANEL_BYTES - EPI_STAGING_BYTES;
  static constexpr int STAGES = BUDGET / A_BYTES > 6 ? 6 : BUDGET / A_BYTES;
  static constexpr int ACC_STRIDE = 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = PANEL_BYTES + STAGES * A_BYTES + EPI_STAGING_BYTES + 1024 + 256;
  static_assert(STAGES >= 2, "weight-stationary slab leaves no room for the A ring");
};

template <int BLOCK_N, int KB_MAX, int EPI>
__global__ void __launch_bounds__(EpiCfg<BLOCK_N>::THREADS, 1)
gemm_tcgen05_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const GemmDev p) {
  pdl_launch_dependents();  // the wait follows the barrier / TMEM set-up below (no global memory touched before it)
  using Cfg = GemmWsCfg<BLOCK_N, KB_MAX>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* panel = smem;
  uint8_t* ring = smem + Cfg::PANEL_BYTES;
  float* epi_staging = reinterpret_cast<float*>(ring + STAGES * Cfg::A_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + STAGES * Cfg::A_BYTES + Cfg::EPI_STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* b_full = tmem_empty + 2;
  uint64_t* b_empty = b_full + 1;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(b_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;  // <= KB_MAX (checked on the host)
  const int total_work = tiles_m * tiles_n;
  const int per_cta = (total_work + gridDim.x - 1) / gridDim.x;
  const int w0 = blockIdx.x * per_cta;
  const int w1 = min(total_work, w0 + per_cta);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EpiCfg<BLOCK_N>::WARPS); }
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) { tmem_alloc(tmem_base_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0, b_empty_phase = 0;
    int cur_nt = -1;
    for (int w = w0; w < w1; ++w) {
      const int nt = w / tiles_m, mt = w % tiles_m;
      const int m0 = mt * BLOCK_M, n0 = nt * BLOCK_N;
      if (nt != cur_nt) {
        if (cur_nt >= 0) { mbar_wait(b_empty, b_empty_phase); b_empty_phase ^= 1; }  // MMAs on the old slab retired
        if (lane == 0) {
          mbar_expect_tx(b_full, kb_total * Cfg::B_BYTES);
          for (int kb = 0; kb < kb_total; ++kb) {
            uint8_t* sb = panel + kb * Cfg::B_BYTES;
            if (!p.b_mn) {
              tma_load_2d(sb, &tmB, b_full, kb * BLOCK_K, n0);
            } else {
#pragma unroll
              for (int a = 0; a < BLOCK_N / 64; ++a) tma_load_2d(sb + a * (BLOCK_K * 128), &tmB, b_full, n0 + a * 64, kb * BLOCK_K);
            }
          }
        }
        __syncwarp();
        cur_nt = nt;
      }
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = ring + stage * Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::A_BYTES);
          if (!p.a_mn) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, m0);
          } else {
#pragma unroll
            for (int a = 0; a < BLOCK_M / 64; ++a) tma_load_2d(sa + a * (BLOCK_K * 128), &tmA, &full_bar[stage], m0 + a * 64, kb * BLOCK_K);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(BLOCK_N, p.a_mn, p.b_mn);
    const uint64_t adesc_base = p.a_mn ? make_smem_desc(smem_u32(ring), BLOCK_K * 128, 1024) : make_smem_desc(smem_u32(ring), 16, 1024);
    const uint64_t bdesc_base = p.b_mn ? make_smem_desc(smem_u32(panel), BLOCK_K * 128, 1024) : make_smem_desc(smem_u32(panel), 16, 1024);
    const uint32_t a_step = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
    const uint32_t b_step = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
    int stage = 0;
    uint32_t phase = 0, b_full_phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int cur_nt = -1;
    for (int w = w0; w < w1; ++w) {
      const int nt = w / tiles_m;
      if (nt != cur_nt) { mbar_wait(b_full, b_full_phase); b_full_phase ^= 1; cur_nt = nt; }
      const bool last_of_slab = (w + 1 == w1) || ((w + 1) / tiles_m != nt);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_STRIDE;
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t ad = adesc_base + (uint64_t)((stage * Cfg::A_BYTES) >> 4);
        const uint64_t bd = bdesc_base + (uint64_t)((kb * Cfg::B_BYTES) >> 4);
        if (elect_one_sync()) {
          umma_f16(tmem_d, ad, bd, idesc, (kb > 0) ? 1u : 0u);
          umma_f16(tmem_d, ad + a_step, bd + b_step, idesc, 1u);
          umma_f16(tmem_d, ad + 2 * a_step, bd + 2 * b_step, idesc, 1u);
          umma_f16(tmem_d, ad + 3 * a_step, bd + 3 * b_step, idesc, 1u);
          umma_commit(&empty_bar[stage]);
          if (kb == kb_total - 1) {
            umma_commit(&tmem_full[acc]);
            if (last_of_slab) umma_commit(b_empty);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if ((acc ^= 1) == 0) acc_phase ^= 1;
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    EpiSched sc{w0, w1, 1, 1, tiles_m, 0};
    epilogue_role<BLOCK_N, EPI>(p, sc, tmem_base, tmem_full, tmem_empty, warp & 3, ew >> 2, lane, epi_staging + ew * (32 * 32));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm as the GEMM's A-operand prologue (north_star: "LayerNorm fused into the adjacent GEMM").
//   y = epilogue( LN(x) @ B^T ),  x fp32 [M, K] (the residual stream), K = embed dim <= 384 (ViT-S / ViT-T)
// A-stationary: a CTA owns whole 128-row panels.  Four dedicated warps normalise the panel's fp32 rows (same arithmetic,
// in the same order, as ln_fwd_kernel: two-pass mean / variance, rsqrt, affine, bf16 rounding) and write them straight
// into the 128B-swizzled K-major tiles the MMA reads (one [128 x 64] tile per k-block, exactly the layout TMA would have
// produced) -- the normalised activations never make a round trip through HBM on their way to the GEMM.  Optionally the
// bf16 rows (student: the wgrad GEMM reads them later) and mean / rstd (LayerNorm backward) are also stored.  The B
// operand (weights, [N, K] K-major) streams through a TMA ring; the MMA warp walks the panel's n-tiles; accumulators are
// double-buffered in TMEM and drained by the shared epilogue role.  The panel is single-buffered (96 KB at K = 384), so the
// normalisation of a CTA's next panel waits for the MMAs of the current one.
template <int BLOCK_N>
struct GemmLnCfg {
  static constexpr int KB_MAX = 6;                             // K <= 384
  static constexpr int A_TILE = BLOCK_M * BLOCK_K * 2;         // one k-block of the panel: 16 KB
  static constexpr int PANEL_BYTES = KB_MAX * A_TILE;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int EPI_STAGING_BYTES = EpiCfg<BLOCK_N>::STAGING_BYTES;
  static constexpr int BUDGET = 232448 - 1024 - 256 - PANEL_BYTES - EPI_STAGING_BYTES;
  static constexpr int STAGES = BUDGET / B_BYTES > 6 ? 6 : BUDGET / B_BYTES;
  static constexpr int LN_WARPS = 4;
  static constexpr int THREADS = EpiCfg<BLOCK_N>::THREADS + LN_WARPS * 32;
  static constexpr int ACC_STRIDE = 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = PANEL_BYTES + STAGES * B_BYTES + EPI_STAGING_BYTES + 1024 + 256;
  static_assert(STAGES >= 2, "no room for the B ring next to the panel");
};

struct LnDev {
  const float* x;
  long long ldx;
  const float* w;
  const float* b;
  float eps;
  __nv_bfloat16* xn;   // optional bf16 copy of LN(x) [M, K]
  long long ldxn;
  float* mean;         // optional [M]
  float* rstd;
};

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GemmLnCfg<BLOCK_N>::THREADS, 1)
gemm_ln_tcgen05_kernel(const __grid_constant__ CUtensorMap tmB, const GemmDev p, const LnDev ln) {
  pdl_launch_dependents();
  using Cfg = GemmLnCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int EPW = EpiCfg<BLOCK_N>::WARPS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* panel = smem;
  uint8_t* ring = smem + Cfg::PANEL_BYTES;
  float* epi_staging = reinterpret_cast<float*>(ring + STAGES * Cfg::B_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + STAGES * Cfg::B_BYTES + Cfg::EPI_STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* a_full = tmem_empty + 2;    // the panel holds LN(x) of the current unit (LN_WARPS arrivals)
  uint64_t* a_empty = a_full + 1;       // every MMA that reads the panel has retired
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(a_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int kb_total = p.K / BLOCK_K;                       // K % 64 == 0, <= KB_MAX (checked on the host)
  const int n_units = (tiles_m - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // panels blockIdx.x, + gridDim.x, ...

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EPW); }
    mbar_init(a_full, Cfg::LN_WARPS);
    mbar_init(a_empty, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmB);
  if (warp == 1) { tmem_alloc(tmem_base_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer: B tiles of every (unit, n-tile, k-block) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int u = 0; u < n_units; ++u) {
      for (int nt = 0; nt < tiles_n; ++nt) {
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (lane == 0) {
            mbar_expect_tx(&full_bar[stage], Cfg::B_BYTES);
            tma_load_2d(ring + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, nt * BLOCK_N);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(BLOCK_N, 0, 0);
    const uint64_t adesc_base = make_smem_desc(smem_u32(panel), 16, 1024);
    const uint64_t bdesc_base = make_smem_desc(smem_u32(ring), 16, 1024);
    constexpr uint32_t kstep = (UMMA_K * 2) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = 0; u < n_units; ++u) {
      mbar_wait(a_full, u & 1);
      tc_fence_after();
      for (int nt = 0; nt < tiles_n; ++nt) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t ad = adesc_base + (uint64_t)((kb * Cfg::A_TILE) >> 4);
          const uint64_t bd = bdesc_base + (uint64_t)((stage * Cfg::B_BYTES) >> 4);
          if (elect_one_sync()) {
            umma_f16(tmem_d, ad, bd, idesc, (kb > 0) ? 1u : 0u);
            umma_f16(tmem_d, ad + kstep, bd + kstep, idesc, 1u);
            umma_f16(tmem_d, ad + 2 * kstep, bd + 2 * kstep, idesc, 1u);
            umma_f16(tmem_d, ad + 3 * kstep, bd + 3 * kstep, idesc, 1u);
            umma_commit(&empty_bar[stage]);
            if (kb == kb_total - 1) {
              umma_commit(&tmem_full[acc]);
              if (nt == tiles_n - 1) umma_commit(a_empty);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      }
    }
  } else if (warp < 2 + EPW) {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    EpiSched sc{0, n_units * tiles_n, 1, 1, tiles_m, tiles_n};
    sc.unit_tiles = tiles_n;
    sc.unit_first = (int)blockIdx.x;
    sc.unit_stride = (int)gridDim.x;
    epilogue_role<BLOCK_N, EPI>(p, sc, tmem_base, tmem_full, tmem_empty, warp & 3, ew >> 2, lane, epi_staging + ew * (32 * 32));
  } else {
    // ===================== LayerNorm warps: rows lw*32 .. lw*32+31 of every panel =====================
    // The panel's fp32 rows (196 KB at K = 384) are latency-bound if read row by row (first version: 40 us per panel).  Each
    // warp first asks L2 for ALL of its 32 rows (prefetch.global.L2: no registers held), then walks them four rows at a
    // time (12 float4 loads per lane in flight); the next panel's rows are requested as soon as this one is written.
    const int lw = warp - 2 - EPW;
    const int D = p.K, nv = D >> 2;
    const int row_lines = (D * 4 + 127) / 128;   // 128-byte lines per row
    auto prefetch_rows = [&](int m0) {
      for (int i = lane; i < 32 * row_lines; i += 32) {
        const int row = m0 + lw * 32 + i / row_lines;
        if (row < p.M) prefetch_l2(reinterpret_cast<const char*>(ln.x + (size_t)row * ln.ldx) + (i % row_lines) * 128);
      }
    };
    if (n_units > 0) prefetch_rows((int)blockIdx.x * BLOCK_M);
    const float4* w4 = reinterpret_cast<const float4*>(ln.w);
    const float4* b4 = reinterpret_cast<const float4*>(ln.b);
    for (int u = 0; u < n_units; ++u) {
      const int m0 = ((int)blockIdx.x + u * (int)gridDim.x) * BLOCK_M;
      if (u > 0) mbar_wait(a_empty, (u - 1) & 1);  // the previous panel's MMAs have retired
      if (nv == 96 && m0 + BLOCK_M <= p.M) {
        // fast path (K = 384, full panel): no bounds predicates, gamma / beta in registers, swizzle offsets hoisted.
        // (ncu on the generic path below: 325 SASS instructions per row and ~36 us per panel on these four warps)
        float4 gw[3], gb[3];
        uint32_t poff[3];                            // row-independent part of the panel offset of this lane's 8-byte pieces
        int chunk[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int c = lane + 32 * j;
          gw[j] = __ldg(w4 + c);
          gb[j] = __ldg(b4 + c);
          poff[j] = (uint32_t)((c >> 4) * Cfg::A_TILE + (c & 1) * 8);
          chunk[j] = (c & 15) >> 1;
        }
        const float inv_d = 1.0f / 384.0f;
        const float* xbase = ln.x + (size_t)(m0 + lw * 32) * ln.ldx;
        __nv_bfloat16* xnbase = ln.xn ? ln.xn + (size_t)(m0 + lw * 32) * ln.ldxn : nullptr;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 4) {
          float4 v[4][3];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float4* xr = reinterpret_cast<const float4*>(xbase + (size_t)(r0 + t) * ln.ldx);
#pragma unroll
            for (int j = 0; j < 3; ++j) v[t][j] = xr[lane + 32 * j];
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int r = lw * 32 + r0 + t;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) s += v[t][j].x + v[t][j].y + v[t][j].z + v[t][j].w;
            const float mu = warp_sum(s) / 384.0f;
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float a = v[t][j].x - mu, bb = v[t][j].y - mu, cc = v[t][j].z - mu, d = v[t][j].w - mu;
              q += a * a + bb * bb + cc * cc + d * d;
            }
            const float rs = rsqrtf(warp_sum(q) / 384.0f + ln.eps);
            if (lane == 0) {
              if (ln.mean) ln.mean[m0 + r] = mu;
              if (ln.rstd) ln.rstd[m0 + r] = rs;
            }
            uint8_t* prow = panel + r * 128;
            const int rx = r & 7;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              uint2 pk;
              pk.x = pack_bf16x2((v[t][j].x - mu) * rs * gw[j].x + gb[j].x, (v[t][j].y - mu) * rs * gw[j].y + gb[j].y);
              pk.y = pack_bf16x2((v[t][j].z - mu) * rs * gw[j].z + gb[j].z, (v[t][j].w - mu) * rs * gw[j].w + gb[j].w);
              if (xnbase) reinterpret_cast<uint2*>(xnbase + (size_t)(r0 + t) * ln.ldxn)[lane + 32 * j] = pk;
              *reinterpret_cast<uint2*>(prow + poff[j] + ((chunk[j] ^ rx) << 4)) = pk;
            }
          }
        }
        (void)inv_d;
      } else {
#pragma unroll 1
      for (int r0 = 0; r0 < 32; r0 += 4) {
        float4 v[4][3];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int row = m0 + lw * 32 + r0 + t;
          const float4* xr = reinterpret_cast<const float4*>(ln.x + (size_t)row * ln.ldx);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int c = lane + 32 * j;
            v[t][j] = (row < p.M && c < nv) ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int r = lw * 32 + r0 + t;            // row inside the panel
          const int row = m0 + r;
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (lane + 32 * j < nv) s += v[t][j].x + v[t][j].y + v[t][j].z + v[t][j].w;
          const float mu = warp_sum(s) / D;
          float q = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (lane + 32 * j < nv) {
              const float a = v[t][j].x - mu, bb = v[t][j].y - mu, cc = v[t][j].z - mu, d = v[t][j].w - mu;
              q += a * a + bb * bb + cc * cc + d * d;
            }
          }
          const float rs = rsqrtf(warp_sum(q) / D + ln.eps);
          if (lane == 0 && row < p.M) {
            if (ln.mean) ln.mean[row] = mu;
            if (ln.rstd) ln.rstd[row] = rs;
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int c = lane + 32 * j;             // float4 index: columns 4c .. 4c+3
            if (c < nv) {
              uint2 pk;
              if (row < p.M) {
                const float4 gw = __ldg(w4 + c), gb = __ldg(b4 + c);
                pk.x = pack_bf16x2((v[t][j].x - mu) * rs * gw.x + gb.x, (v[t][j].y - mu) * rs * gw.y + gb.y);
                pk.y = pack_bf16x2((v[t][j].z - mu) * rs * gw.z + gb.z, (v[t][j].w - mu) * rs * gw.w + gb.w);
                if (ln.xn) reinterpret_cast<uint2*>(ln.xn + (size_t)row * ln.ldxn)[c] = pk;
              } else {
                pk = make_uint2(0u, 0u);             // rows past M: zeros (their outputs are never stored)
              }
              // column 4c -> k-block (4c)/64 = c/16, 16-byte chunk ((4c)%64)/8 = (c%16)/2, 8-byte half c&1
              const int kb = c >> 4, chunk = (c & 15) >> 1;
              *reinterpret_cast<uint2*>(panel + kb * Cfg::A_TILE + r * 128 + ((chunk ^ (r & 7)) << 4) + (c & 1) * 8) = pk;
            }
          }
        }
      }
      }
      fence_proxy_async();  // the MMAs (async proxy) read what these threads wrote through the generic proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      if (u + 1 < n_units) prefetch_rows(((int)blockIdx.x + (u + 1) * (int)gridDim.x) * BLOCK_M);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess) return nullptr;
  if (qres != cudaDriverEntryPointSuccess) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

// rows x cols bf16 matrix, cols contiguous, leading dimension ld (elements); box = box_cols x box_rows
static int make_tmap(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_cols,
                     int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return B200_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B200_OK : B200_ERR_DRIVER;
}

static int g_num_sms = 0;

template <int BLOCK_N, int EPI>
static int launch_gemm_epi(const b200_gemm_args* a, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap tmA, tmB;
  int rc;
  if (!a->a_mn) rc = make_tmap(&tmA, a->A, a->M, a->K, a->lda, BLOCK_K, BLOCK_M);
  else rc = make_tmap(&tmA, a->A, a->K, a->M, a->lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!a->b_mn) rc = make_tmap(&tmB, a->B, a->N, a->K, a->ldb, BLOCK_K, BLOCK_N);
  else rc = make_tmap(&tmB, a->B, a->K, a->N, a->ldb, 64, BLOCK_K);
  if (rc) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_tcgen05_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg::SMEM_BYTES) != cudaSuccess)
      return B200_ERR_CUDA;
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  GemmDev p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_mn; p.b_mn = a->b_mn;
  p.splits = a->splits < 1 ? 1 : a->splits;
  p.epi = a->epi;
  p.alpha = a->alpha;
  p.C = a->C; p.ldc = a->ldc;
  p.C2 = a->C2; p.ldc2 = a->ldc2;
  p.aux = a->aux; p.ldaux = a->ldaux;
  p.bias = a->bias; p.gamma = a->gamma;
  p.rowscale = a->rowscale; p.rows_per_scale = a->rows_per_scale > 0 ? a->rows_per_scale : 1;

  const int tiles = ((a->M + BLOCK_M - 1) / BLOCK_M) * ((a->N + BLOCK_N - 1) / BLOCK_N);
  const int kb_total = (a->K + BLOCK_K - 1) / BLOCK_K;
  int splits = p.splits > kb_total ? kb_total : p.splits;
  p.splits = splits;
  long long work = (long long)tiles * splits;
  int grid = (int)(work < g_num_sms ? work : g_num_sms);
  if (grid < 1) grid = 1;
  launch_kernel(gemm_tcgen05_kernel<BLOCK_N, EPI>, grid, EpiCfg<BLOCK_N>::THREADS, Cfg::SMEM_BYTES, stream, tmA, tmB, p);
  B200_CHECK_LAUNCH();
  return B200_OK;
}


template <int BLOCK_N, int KB_MAX, int EPI>
static int launch_gemm_ws_epi(const b200_gemm_args* a, cudaStream_t stream) {
  using Cfg = GemmWsCfg<BLOCK_N, KB_MAX>;
  CUtensorMap tmA, tmB;
  int rc;
  if (!a->a_mn) rc = make_tmap(&tmA, a->A, a->M, a->K, a->lda, BLOCK_K, BLOCK_M);
  else rc = make_tmap(&tmA, a->A, a->K, a->M, a->lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!a->b_mn) rc = make_tmap(&tmB, a->B, a->N, a->K, a->ldb, BLOCK_K, BLOCK_N);
  else rc = make_tmap(&tmB, a->B, a->K, a->N, a->ldb, 64, BLOCK_K);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_tcgen05_ws_kernel<BLOCK_N, KB_MAX, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg::SMEM_BYTES) != cudaSuccess)
      return B200_ERR_CUDA;
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  GemmDev p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_mn; p.b_mn = a->b_mn;
  p.splits = 1;
  p.epi = a->epi;
  p.alpha = a->alpha;
  p.C = a->C; p.ldc = a->ldc;
  p.C2 = a->C2; p.ldc2 = a->ldc2;
  p.aux = a->aux; p.ldaux = a->ldaux;
  p.bias = a->bias; p.gamma = a->gamma;
  p.rowscale = a->rowscale; p.rows_per_scale = a->rows_per_scale > 0 ? a->rows_per_scale : 1;
  const long long tiles = (long long)((a->M + BLOCK_M - 1) / BLOCK_M) * ((a->N + BLOCK_N - 1) / BLOCK_N);
  int grid = (int)(tiles < g_num_sms ? tiles : g_num_sms);
  // even out the contiguous runs: ceil(tiles/grid) tiles per CTA, drop CTAs that would get nothing
  const long long per = (tiles + grid - 1) / grid;
  grid = (int)((tiles + per - 1) / per);
  launch_kernel(gemm_tcgen05_ws_kernel<BLOCK_N, KB_MAX, EPI>, grid, EpiCfg<BLOCK_N>::THREADS, Cfg::SMEM_BYTES, stream, tmA, tmB, p);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

template <int BLOCK_N, int EPI>
static int launch_gemm_2sm_epi(const b200_gemm_args* a, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BLOCK_N>;
  CUtensorMap tmA, tmB;
  int rc;
  if (!a->a_mn) rc = make_tmap(&tmA, a->A, a->M, a->K, a->lda, BLOCK_K, BLOCK_M);
  else rc = make_tmap(&tmA, a->A, a->K, a->M, a->lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!a->b_mn) rc = make_tmap(&tmB, a->B, a->N, a->K, a->ldb, BLOCK_K, BLOCK_N / 2);
  else rc = make_tmap(&tmB, a->B, a->K, a->N, a->ldb, 64, BLOCK_K);
  if (rc) return rc;
  if (a->b_mn && (BLOCK_N % 128) != 0) return B200_ERR_UNSUPPORTED;  // the B half must be whole 64-wide MN atoms
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_tcgen05_2sm_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg::SMEM_BYTES) != cudaSuccess)
      return B200_ERR_CUDA;
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  GemmDev p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_mn; p.b_mn = a->b_mn;
  p.splits = a->splits < 1 ? 1 : a->splits;
  p.epi = a->epi;
  p.alpha = a->alpha;
  p.C = a->C; p.ldc = a->ldc;
  p.C2 = a->C2; p.ldc2 = a->ldc2;
  p.aux = a->aux; p.ldaux = a->ldaux;
  p.bias = a->bias; p.gamma = a->gamma;
  p.rowscale = a->rowscale; p.rows_per_scale = a->rows_per_scale > 0 ? a->rows_per_scale : 1;
  const int tiles = ((a->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((a->N + BLOCK_N - 1) / BLOCK_N);
  const int kb_total = (a->K + BLOCK_K - 1) / BLOCK_K;
  if (p.splits > kb_total) p.splits = kb_total;
  const long long work = (long long)tiles * p.splits;
  const int max_pairs = g_num_sms / 2;
  int pairs = (int)(work < max_pairs ? work : max_pairs);
  if (pairs < 1) pairs = 1;
  launch_kernel(gemm_tcgen05_2sm_kernel<BLOCK_N, EPI>, 2 * pairs, EpiCfg<BLOCK_N>::THREADS, Cfg::SMEM_BYTES, stream, tmA, tmB, p);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// one kernel per (tile width, epilogue): the epilogue is inlined, its parameters stay in the constant bank and the
// register allocation is the epilogue's own
#define B200_EPI_SWITCH(CALL)                                              \
  switch (a->epi) {                                                        \
    case B200_EPI_BF16: return CALL(B200_EPI_BF16);                        \
    case B200_EPI_F32: return CALL(B200_EPI_F32);                          \
    case B200_EPI_F32_ATOMIC: return CALL(B200_EPI_F32_ATOMIC);            \
    case B200_EPI_BIAS_GELU: return CALL(B200_EPI_BIAS_GELU);              \
    case B200_EPI_RESIDUAL: return CALL(B200_EPI_RESIDUAL);                \
    case B200_EPI_DGELU: return CALL(B200_EPI_DGELU);                      \
    case B200_EPI_BIAS_GELU_DG: return CALL(B200_EPI_BIAS_GELU_DG);        \
    case B200_EPI_MUL_AUX: return CALL(B200_EPI_MUL_AUX);                  \
    default: return B200_ERR_INVALID_ARG;                                  \
  }
template <int BLOCK_N>
static int launch_gemm(const b200_gemm_args* a, cudaStream_t stream) {
#define B200_CALL(E) launch_gemm_epi<BLOCK_N, E>(a, stream)
  B200_EPI_SWITCH(B200_CALL)
#undef B200_CALL
}
template <int BLOCK_N, int KB_MAX>
static int launch_gemm_ws(const b200_gemm_args* a, cudaStream_t stream) {
#define B200_CALL(E) launch_gemm_ws_epi<BLOCK_N, KB_MAX, E>(a, stream)
  B200_EPI_SWITCH(B200_CALL)
#undef B200_CALL
}
template <int BLOCK_N>
static int launch_gemm_2sm(const b200_gemm_args* a, cudaStream_t stream) {
#define B200_CALL(E) launch_gemm_2sm_epi<BLOCK_N, E>(a, stream)
  B200_EPI_SWITCH(B200_CALL)
#undef B200_CALL
}

template <int BLOCK_N, int EPI>
static int launch_gemm_ln_epi(const b200_gemm_args* a, const b200_ln_args* l, cudaStream_t stream) {
  using Cfg = GemmLnCfg<BLOCK_N>;
  CUtensorMap tmB;
  int rc = make_tmap(&tmB, a->B, a->N, a->K, a->ldb, BLOCK_K, BLOCK_N);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_ln_tcgen05_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) !=
        cudaSuccess)
      return B200_ERR_CUDA;
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  GemmDev p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = 0; p.b_mn = 0;
  p.splits = 1;
  p.epi = a->epi;
  p.alpha = a->alpha;
  p.C = a->C; p.ldc = a->ldc;
  p.C2 = a->C2; p.ldc2 = a->ldc2;
  p.aux = nullptr; p.ldaux = 0;
  p.bias = a->bias; p.gamma = nullptr;
  p.rowscale = nullptr; p.rows_per_scale = 1;
  LnDev ln;
  ln.x = l->x; ln.ldx = l->ldx; ln.w = l->weight; ln.b = l->bias; ln.eps = l->eps;
  ln.xn = (__nv_bfloat16*)l->xn_out; ln.ldxn = l->ld_xn; ln.mean = l->mean; ln.rstd = l->rstd;
  const int tiles_m = (a->M + BLOCK_M - 1) / BLOCK_M;
  const int grid = tiles_m < g_num_sms ? tiles_m : g_num_sms;
  launch_kernel(gemm_ln_tcgen05_kernel<BLOCK_N, EPI>, grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, tmB, p, ln);
  B200_CHECK_LAUNCH();
  return B200_OK;
}


}  // namespace b200

// y = epilogue(LayerNorm(x) @ B^T): LayerNorm as the A-operand prologue of the GEMM (K = embed dim, K % 64 == 0, K <= 384).
extern "C" int b200_ln_gemm(const b200_gemm_args* a, const b200_ln_args* l, void* stream) {
  using namespace b200;
  if (!a || !l || !a->B || !a->C || !l->x || !l->weight || !l->bias) return B200_ERR_INVALID_ARG;
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return B200_ERR_INVALID_ARG;
  if ((a->K % BLOCK_K) || a->K > GemmLnCfg<128>::KB_MAX * BLOCK_K) return B200_ERR_UNSUPPORTED;
  if (a->a_mn || a->b_mn || a->splits > 1) return B200_ERR_UNSUPPORTED;
  if ((a->ldb % 8) || (a->N % 8) || (a->ldc % 8) || (l->ldx % 4) || (a->C2 && (a->ldc2 % 8)) || (l->xn_out && (l->ld_xn % 4)))
    return B200_ERR_UNSUPPORTED;
  if (((uintptr_t)a->B & 15) || ((uintptr_t)a->C & 15) || ((uintptr_t)l->x & 15)) return B200_ERR_UNSUPPORTED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int bn = a->block_n ? a->block_n : (a->epi == B200_EPI_BF16 ? 128 : 192);
  switch (a->epi) {
    case B200_EPI_BF16:
      return bn == 192 ? launch_gemm_ln_epi<192, B200_EPI_BF16>(a, l, s) : launch_gemm_ln_epi<128, B200_EPI_BF16>(a, l, s);
    case B200_EPI_BIAS_GELU:
      return bn == 192 ? launch_gemm_ln_epi<192, B200_EPI_BIAS_GELU>(a, l, s) : launch_gemm_ln_epi<128, B200_EPI_BIAS_GELU>(a, l, s);
    case B200_EPI_BIAS_GELU_DG:
      return bn == 192 ? launch_gemm_ln_epi<192, B200_EPI_BIAS_GELU_DG>(a, l, s)
                       : launch_gemm_ln_epi<128, B200_EPI_BIAS_GELU_DG>(a, l, s);
    default:
      return B200_ERR_UNSUPPORTED;
  }
}

extern "C" int b200_gemm(const b200_gemm_args* a, void* stream) {
  using namespace b200;
  if (!a || !a->A || !a->B || !a->C) return B200_ERR_INVALID_ARG;
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return B200_ERR_INVALID_ARG;
  // TMA: 16-byte aligned bases and row pitches; vectorised epilogue: N, ldc multiples of 8
  if ((a->lda % 8) || (a->ldb % 8) || (a->N % 8) || (a->ldc % 8)) return B200_ERR_UNSUPPORTED;
  if (((uintptr_t)a->A & 15) || ((uintptr_t)a->B & 15) || ((uintptr_t)a->C & 15)) return B200_ERR_UNSUPPORTED;
  if (a->epi < 0 || a->epi > B200_EPI_MUL_AUX) return B200_ERR_INVALID_ARG;
  if (a->splits > 1 && a->epi != B200_EPI_F32_ATOMIC) return B200_ERR_INVALID_ARG;
  b200_gemm_args tuned;
  if (a->splits == 0 && a->epi == B200_EPI_F32_ATOMIC) {
    // auto split-K (wgrad: few output tiles, very long K): pick (tile width, splits) so that tiles*splits fills whole
    // waves of the 148 SMs; cost ~ rounds * k-blocks per split * operand bytes per k-block
    if (g_num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int kbt = (a->K + BLOCK_K - 1) / BLOCK_K;
    const int tm = (a->M + BLOCK_M - 1) / BLOCK_M;
    const int cand[3] = {192, 256, 128};
    long long best = -1;
    int best_bn = 192, best_sp = 1;
    for (int i = 0; i < 3; ++i) {
      if (a->block_n != 0 && a->block_n != cand[i]) continue;
      const int tiles = tm * ((a->N + cand[i] - 1) / cand[i]);
      for (int r = 1; r <= 6; ++r) {
        int sp = (g_num_sms * r) / tiles;
        if (sp < 1) continue;
        if (sp > kbt / 4) sp = kbt / 4 > 0 ? kbt / 4 : 1;
        const int rounds = (tiles * sp + g_num_sms - 1) / g_num_sms;
        // per k-block a CTA pulls (BLOCK_M + BLOCK_N) x 128 B through L2: these long-K GEMMs run at the L2->SM
        // bandwidth (~11 TB/s measured), not at the MMA rate, so the cost is bytes, not BLOCK_N; +6: per-item overhead
        const long long cost = (long long)rounds * ((kbt + sp - 1) / sp + 6) * (BLOCK_M + cand[i]);
        if (best < 0 || cost < best) { best = cost; best_bn = cand[i]; best_sp = sp; }
      }
    }
    tuned = *a;
    tuned.block_n = best_bn;
    tuned.splits = best_sp;
    a = &tuned;
  }
  if ((a->epi == B200_EPI_RESIDUAL || a->epi == B200_EPI_DGELU || a->epi == B200_EPI_MUL_AUX) && (!a->aux || (a->ldaux % 8)))
    return B200_ERR_INVALID_ARG;
  if (a->C2 && (a->ldc2 % 8)) return B200_ERR_UNSUPPORTED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int kb_total = (a->K + BLOCK_K - 1) / BLOCK_K;
  const int tiles_m = (a->M + BLOCK_M - 1) / BLOCK_M;
  // weight-stationary schedule: small K (the B slab fits in smem next to the A ring) and enough M tiles to amortise it
  // measured on B200 (profiles/r01_gemm_*.log): the slab schedule wins only for K <= 256 (projection-head last layer);
  // for K = 384 the generic ring has more bytes in flight and is as fast or faster
  const bool want_ws = a->ws_mode == 1 || (a->ws_mode == 0 && a->splits <= 1 && tiles_m >= 8 && kb_total <= 4);
  int bn = a->block_n;
  bool use_pair = false;
  if (bn == 0) {
    // tile-N heuristic: minimise (waves over the SMs) x (operand bytes a CTA pulls through L2 per k-block, ~ BLOCK_M +
    // BLOCK_N).  These K<=1536 GEMMs run at the L2->SM bandwidth (~11 TB/s measured: tools/gemm_sweep.py,
    // profiles/r01_gemm_sweep.log), not at the MMA rate, so wider tiles win unless they add a wave or padding.
    if (g_num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int cand[3] = {192, 256, 128};
    const int splits = a->splits < 1 ? 1 : a->splits;
    long long best = -1;
    for (int i = 0; i < 3; ++i) {
      if (want_ws && kb_total > 4 && cand[i] == 256) continue;  // a 256 x 384 slab does not fit
      const long long tiles = (long long)tiles_m * ((a->N + cand[i] - 1) / cand[i]) * splits;
      long long cost = ((tiles + g_num_sms - 1) / g_num_sms) * (BLOCK_M + cand[i]) * 8;
      // the GELU / dGELU epilogues are issue-bound: the 12-warp (192-wide) configuration measured 10-15 % faster
      if ((a->epi == B200_EPI_BIAS_GELU || a->epi == B200_EPI_DGELU || a->epi == B200_EPI_BIAS_GELU_DG) && cand[i] != 192)
        cost += cost / 4;
      // the fp32 residual epilogue keeps two chunks of the stream in registers: 128-wide tiles (8 warps, 168 registers,
      // two chunks per warp) measured 29 / 43.5 us against 35 / 47-49 us for the wider ones (proj / fc2 shapes)
      if (a->epi == B200_EPI_RESIDUAL && cand[i] != 128) cost += cost / 4;
      if (best < 0 || cost < best) { best = cost; bn = cand[i]; }
    }
    // CTA-pair kernel (256-row tiles): measured 4 % faster than the single-CTA kernel at equal wave counts and
    // 10 % faster on large-K GEMMs (8192^3: 1.46 vs 1.33 PFLOP/s), but its waves are counted in pairs -- taken only
    // when that does not add a round, for the plain epilogues, and for K-major B or whole 64-wide MN atoms per half
    if (a->ws_mode == 0 && !want_ws && splits == 1 && (a->epi == B200_EPI_BF16 || a->epi == B200_EPI_F32)) {
      const int pairs = g_num_sms / 2;
      const int tiles_m2 = (a->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
      for (int i = 0; i < 3; ++i) {
        if (a->b_mn && (cand[i] % 128) != 0) continue;
        // no extra padding: an odd number of 128-row tiles or a partial last column tile costs the pair more than it
        // gains (M = 25216, N = 1152, 256-wide: 28.7 us against 25.8 us single-CTA 192-wide)
        if (2 * tiles_m2 != tiles_m || (a->N % cand[i]) != 0) continue;
        const long long tiles2 = (long long)tiles_m2 * ((a->N + cand[i] - 1) / cand[i]);
        const long long cost2 = ((tiles2 + pairs - 1) / pairs) * (BLOCK_M + cand[i]) * 8 * 19 / 20;
        if (cost2 < best) { best = cost2; bn = cand[i]; use_pair = true; }
      }
    }
  }
  if (a->ws_mode == 3 || use_pair) {
    switch (bn) {
      case 128: return launch_gemm_2sm<128>(a, s);
      case 192: return launch_gemm_2sm<192>(a, s);
      case 256: return launch_gemm_2sm<256>(a, s);
      default: return B200_ERR_INVALID_ARG;
    }
  }
  if (want_ws) {
    if (a->splits > 1) return B200_ERR_INVALID_ARG;
    if (kb_total <= 4 && bn == 256) return launch_gemm_ws<256, 4>(a, s);
    if (kb_total <= 6 && bn == 192) return launch_gemm_ws<192, 6>(a, s);
    if (kb_total <= 6 && bn == 128) return launch_gemm_ws<128, 6>(a, s);
    if (a->ws_mode == 1) return B200_ERR_UNSUPPORTED;
  }
  switch (bn) {
    case 128: return launch_gemm<128>(a, s);
    case 192: return launch_gemm<192>(a, s);
    case 256: return launch_gemm<256>(a, s);
    default: return B200_ERR_INVALID_ARG;
  }
}
