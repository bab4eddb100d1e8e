// Short-sequence attention FORWARD on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) -- BRING-UP, NOT WIRED.
//
// Status: compiles for sm_100a, exported as b200_attention_fwd_tc with the signature of b200_attention_fwd, but it is
// not called by lightly_train_b200/ops.py and has not run on hardware yet (the round's GPU budget was spent); the
// product path uses the mma.sync kernels of attention.cu.  tests/test_kernels_gpu.py::test_attention_fwd_tcgen05 is the
// parity test that must pass before it is switched on (opt-in through B200_TEST_TC_ATTENTION=1 until then).
// Design and expected gain: DESIGN.md section 7.
//
// Replaces (like attention.cu) Attention.forward of LT/_models/dinov2_vit/dinov2_vit_src/layers/attention.py:49-66.
//
// One CTA per (image, head).  qkv: bf16 [B*N, 3*h*64] (row pitch ld_tok); out: bf16 [B*N, h*64].
//   warp 4 (one elected lane): TMA loads of the pair's Q (two 128-row boxes), K and V rows (NKV = 16*ceil(N/16) rows;
//            rows past the image's N tokens belong to the next image or are zero-filled by TMA: they are MASKED, never
//            relied on), then per 128-row query tile  S = Q K^T  (4 x tcgen05.mma M128 x N=NKV x K16, fp32 in TMEM) and
//            O = P V  (NKV/16 x tcgen05.mma M128 x N64 x K16, A = P from smem K-major, B = V from smem MN-major).
//   warps 0-3: one query row per thread (TMEM lane = row): tcgen05.ld of the S row in 16-column chunks, row max and
//            sum thread-local (no shuffles), P = 2^((s - m) * scale * log2e) rounded to bf16 UNNORMALISED into a
//            128B-swizzled smem tile (flash-style: O is scaled by 1/l after the second MMA), then O row -> global.
// S is rounded to bf16 before the softmax (autocast reference; the power-of-two scale commutes with the rounding).
#include <cuda.h>
#include "common.cuh"
#include "../../include/b200dino.h"

namespace b200 {
namespace attn_tc {

static constexpr int HD = 64;
static constexpr int BLOCK_Q = 128;                 // query rows per MMA tile (= TMEM lanes)
static constexpr int Q_BYTES = 2 * BLOCK_Q * 128;   // two query tiles
static constexpr int P_SLAB = BLOCK_Q * 128;        // one 64-key slab of the P tile
static constexpr int THREADS = 5 * 32;

// same encodings as gemm_tcgen05.cu (kept local: that file is hardware-verified and stays untouched)
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t instr_desc(int m, int n, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                    // D = f32
  d |= 1u << 7;                    // A = bf16
  d |= 1u << 10;                   // B = bf16
  d |= (uint32_t)(b_mn ? 1 : 0) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}
// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

template <int NKV16>
struct Cfg {
  static constexpr int NKV = NKV16 * 16;                   // padded key count (MMA N of S, K of P V), <= 256
  static constexpr int KV_BYTES = NKV * 128;
  static constexpr int P_SLABS = (NKV + 63) / 64;
  static constexpr int OFF_K = Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_BYTES;
  static constexpr int OFF_P = (OFF_V + KV_BYTES + 1023) / 1024 * 1024;
  static constexpr int OFF_BAR = OFF_P + P_SLABS * P_SLAB;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;  // barriers + TMEM slot, alignment slack
  static constexpr int TM_S = 0, TM_O = NKV;               // TMEM columns: S [0, NKV), O [NKV, NKV + 64)
  static_assert(NKV <= 256 && NKV + HD <= 512, "one S tile and one O tile must fit TMEM");
};

template <int NKV16>
__global__ void __launch_bounds__(THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, int N, int h,
                   float scale, __nv_bfloat16* __restrict__ out, long long ld_out, float* __restrict__ lse) {
  using C = Cfg<NKV16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* bar_load = bars + 0;  // TMA bytes landed
  uint64_t* bar_s = bars + 1;     // S = Q K^T of the current tile is in TMEM
  uint64_t* bar_p = bars + 2;     // P tile written to smem (4 worker warps)
  uint64_t* bar_o = bars + 3;     // O = P V of the current tile is in TMEM
  uint64_t* bar_free = bars + 4;  // workers are done with this tile's TMEM (4 worker warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x, b = bh / h, head = bh % h;
  const int n_tiles = (N + BLOCK_Q - 1) / BLOCK_Q;  // 1 or 2 query tiles

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
    mbar_init(bar_free, 4);
    fence_barrier_init();
  }
  if (warp == 4) {
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
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================== control warp: TMA + MMA issue =====================
    const int row0 = b * N;  // first token row of this image in the [B*N, 3*h*64] qkv matrix
    if (lane == 0) {
      mbar_expect_tx(bar_load, Q_BYTES + 2 * C::KV_BYTES);
      tma_load_2d(smem, &tmQ, bar_load, head * HD, row0);
      tma_load_2d(smem + BLOCK_Q * 128, &tmQ, bar_load, head * HD, row0 + BLOCK_Q);
      tma_load_2d(smem + C::OFF_K, &tmKV, bar_load, (h + head) * HD, row0);
      tma_load_2d(smem + C::OFF_V, &tmKV, bar_load, (2 * h + head) * HD, row0);
    }
    __syncwarp();
    mbar_wait(bar_load, 0);
    const uint32_t s0 = smem_u32(smem);
    const uint64_t dq = smem_desc(s0, 16, 1024);                    // A of S: Q tile, K-major
    const uint64_t dk = smem_desc(s0 + C::OFF_K, 16, 1024);         // B of S: K rows, K-major (N = keys)
    const uint64_t dp = smem_desc(s0 + C::OFF_P, 16, 1024);         // A of O: P tile, K-major (K = keys)
    const uint64_t dv = smem_desc(s0 + C::OFF_V, 64 * 128, 1024);   // B of O: V rows, MN-major (N = head dim)
    const uint32_t idesc_s = instr_desc(BLOCK_Q, C::NKV, 0);
    const uint32_t idesc_o = instr_desc(BLOCK_Q, HD, 1);
    for (int t = 0; t < n_tiles; ++t) {
      const uint32_t ph = t & 1;
      if (t > 0) {
        mbar_wait(bar_free, ph ^ 1);  // workers drained tile t-1's S and O
        tc_fence_after();
      }
      if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_f16(tmem_base + C::TM_S, dq + (uint64_t)((t * BLOCK_Q * 128 + ks * 32) >> 4), dk + (uint64_t)((ks * 32) >> 4),
                   idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(bar_s);
      }
      __syncwarp();
      mbar_wait(bar_p, ph);  // P tile complete in smem (workers fenced the generic->async proxy before arriving)
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll 1
        for (int ks = 0; ks < NKV16; ++ks)
          umma_f16(tmem_base + C::TM_O, dp + (uint64_t)(((ks >> 2) * P_SLAB + (ks & 3) * 32) >> 4),
                   dv + (uint64_t)((ks * 16 * 128) >> 4), idesc_o, ks > 0 ? 1u : 0u);
        umma_commit(bar_o);
      }
      __syncwarp();
    }
  } else {
    // ===================== worker warps: one query row per thread =====================
    const int q = warp;                 // TMEM lane quarter == warp index (warps 0-3)
    const int r = q * 32 + lane;        // row inside the 128-row tile
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sl2 = scale * 1.4426950408889634f;
    uint8_t* sP = smem + C::OFF_P;
    for (int t = 0; t < n_tiles; ++t) {
      const uint32_t ph = t & 1;
      const int m = t * BLOCK_Q + r;    // query token index inside the image
      mbar_wait(bar_s, ph);
      tc_fence_after();
      // pass 1: row max of the bf16-rounded scores over the valid keys
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < NKV16; ++c) {
        uint32_t v[16];
        tmem_ld_32x16(lane_base + C::TM_S + c * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float a = __uint_as_float(v[i]), bq = __uint_as_float(v[i + 1]);
          const uint32_t pk = pack_bf16x2(a, bq);
          const float2 rr = unpack_bf16x2(pk);
          if (c * 16 + i < N) mx = fmaxf(mx, rr.x);
          if (c * 16 + i + 1 < N) mx = fmaxf(mx, rr.y);
        }
      }
      const float mb = -mx * sl2;
      // pass 2: p = 2^(s*sl2 - m*sl2), row sum in fp32, P (unnormalised, bf16) -> swizzled smem A tile
      float l = 0.f;
#pragma unroll 1
      for (int c = 0; c < NKV16; ++c) {
        uint32_t v[16];
        tmem_ld_32x16(lane_base + C::TM_S + c * 16, v);
        tmem_ld_wait();
        uint32_t pw[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float2 rr = unpack_bf16x2(pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          const float p0 = (c * 16 + i < N) ? ex2_ftz(fmaf(rr.x, sl2, mb)) : 0.f;
          const float p1 = (c * 16 + i + 1 < N) ? ex2_ftz(fmaf(rr.y, sl2, mb)) : 0.f;
          l += p0 + p1;
          pw[i >> 1] = pack_bf16x2(p0, p1);
        }
        // keys [c*16, c*16+16) = 16-byte chunks (c*2) and (c*2+1) of slab c/4; 128B swizzle: chunk ^ (row & 7)
        uint8_t* slab = sP + (c >> 2) * P_SLAB + r * 128;
        const int ch = (c & 3) * 2;
        *reinterpret_cast<uint4*>(slab + (((ch) ^ (r & 7)) << 4)) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
        *reinterpret_cast<uint4*>(slab + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pw[4], pw[5], pw[6], pw[7]);
      }
      fence_proxy_async();  // the MMA (async proxy) reads what this thread just wrote through the generic proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      if (lse && m < N) lse[(size_t)bh * N + m] = mx * scale + __logf(l);
      // O row: scale by 1/l, round to bf16, 128 contiguous bytes per row
      mbar_wait(bar_o, ph);
      tc_fence_after();
      const float inv = 1.f / l;
      __nv_bfloat16* orow = out + ((size_t)b * N + m) * ld_out + head * HD;
#pragma unroll
      for (int c = 0; c < HD / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x16(lane_base + C::TM_O + c * 16, v);
        tmem_ld_wait();
        if (m < N) {
          uint32_t ow[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) ow[i >> 1] = pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 16) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          *reinterpret_cast<uint4*>(orow + c * 16 + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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
static int launch(const void* qkv, long long ld_tok, int B, int N, int h, float scale, void* out, long long ld_out, float* lse,
                  cudaStream_t s) {
  using C = Cfg<NKV16>;
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
  attn_fwd_tc_kernel<NKV16><<<B * h, THREADS, C::SMEM_BYTES, s>>>(tmQ, tmKV, N, h, scale, (__nv_bfloat16*)out, ld_out, lse);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace attn_tc
}  // namespace b200

extern "C" int b200_attention_fwd_tc(const void* qkv, long long ld_tok, int B, int N, int h, int head_dim, float scale, void* out,
                                     long long ld_out, float* lse, void* stream) {
  using namespace b200::attn_tc;
  if (!qkv || !out || B <= 0 || N <= 0 || h <= 0 || !(scale > 0.f)) return B200_ERR_INVALID_ARG;
  if (head_dim != HD || (ld_tok % 8) || (ld_out % 8) || ((uintptr_t)qkv & 15) || ((uintptr_t)out & 15)) return B200_ERR_UNSUPPORTED;
  if (N > 256) return B200_ERR_UNSUPPORTED;  // two 128-row query tiles, one MMA N <= 256 of keys
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = (N + 15) / 16;
  if (nb <= 3) return launch<3>(qkv, ld_tok, B, N, h, scale, out, ld_out, lse, s);
  if (nb <= 13) return launch<13>(qkv, ld_tok, B, N, h, scale, out, ld_out, lse, s);
  return launch<16>(qkv, ld_tok, B, N, h, scale, out, ld_out, lse, s);
}
