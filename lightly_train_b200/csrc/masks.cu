// Device-side block-wise mask generation for the iBOT term (the input side of the step).
//
// Same ALGORITHM and distribution as MaskingGenerator / create_collated_masks of LT/_methods/dinov2/utils.py:41-152 (BEiT
// block masking: random rectangles with log-uniform aspect ratio, accepted when they add 1..budget new patches, up to 10
// proposals per block, until the per-crop target count is reached), with a counter-based device RNG instead of python's
// `random` stream: statistical parity, not bit parity (the host generator in _methods/dinov2/utils.py keeps bit parity
// and stays the default).  The per-crop TARGET counts are drawn on the host (a few dozen uniform draws per step) so that
// the padded masked-token capacity of the static-shape step is known before launch; everything else -- rectangles, the
// collated index list, the per-token weights 1/count, the padding masks -- is produced on the device, inside the captured
// step, with no host -> device staging of masks.
//
//   block_masks_kernel    one warp per crop; the crop's mask lives in shared memory; lane 0 draws, all lanes count / fill
//   collate_masks_kernel  one CTA: exclusive scan over the B*Np mask bytes -> mask_indices_list, masks_weight, 1/M row
//                         weights, padding masks, M
#include "common.cuh"
#include "../../include/b200dino.h"

namespace b200 {

struct Rng {  // splitmix64: counter-based, one independent stream per (seed, step, crop)
  unsigned long long s;
  __device__ __forceinline__ unsigned long long next() {
    unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  __device__ __forceinline__ float uniform() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }  // [0, 1)
  __device__ __forceinline__ int randint(int lo, int hi) {  // inclusive, like random.randint
    return lo + (int)(next() % (unsigned long long)(hi - lo + 1));
  }
};

static constexpr int MASK_WARPS = 4;

__global__ void __launch_bounds__(MASK_WARPS * 32)
block_masks_kernel(const int* __restrict__ targets, int B, int H, int W, int min_patches, int max_patches, float log_ar_lo,
                   float log_ar_hi, unsigned long long seed, const int* __restrict__ step_dev, unsigned char* __restrict__ masks) {
  B200_PDL_SYNC();
  extern __shared__ unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crop = blockIdx.x * MASK_WARPS + warp;
  if (crop >= B) return;
  const int Np = H * W;
  unsigned char* m = sm + (size_t)warp * ((Np + 15) / 16 * 16);
  for (int i = lane; i < Np; i += 32) m[i] = 0;
  __syncwarp();
  const int target = targets[crop];
  const unsigned long long step = step_dev ? (unsigned long long)(unsigned int)*step_dev : 0ull;
  Rng rng{seed * 0xD1B54A32D192ED03ull + step * 0x9E3779B97F4A7C15ull + (unsigned long long)crop * 0xBF58476D1CE4E5B9ull};
  int count = 0;
  while (count < target) {
    const int budget = min(target - count, max_patches);
    int added = 0;
    for (int attempt = 0; attempt < 10 && added == 0; ++attempt) {
      int bh = 0, bw = 0, top = 0, left = 0;
      if (lane == 0) {
        const float area = (float)min_patches + rng.uniform() * (float)(budget - min_patches);  // random.uniform(min, budget)
        const float ratio = __expf(log_ar_lo + rng.uniform() * (log_ar_hi - log_ar_lo));
        bh = (int)rintf(sqrtf(area * ratio));
        bw = (int)rintf(sqrtf(area / ratio));
        if (bw < W && bh < H) {
          top = rng.randint(0, H - bh);
          left = rng.randint(0, W - bw);
        } else {
          bh = -1;
        }
      }
      bh = __shfl_sync(0xffffffffu, bh, 0); bw = __shfl_sync(0xffffffffu, bw, 0);
      top = __shfl_sync(0xffffffffu, top, 0); left = __shfl_sync(0xffffffffu, left, 0);
      if (bh < 0) continue;
      const int cells = bh * bw;
      int covered = 0;
      for (int i = lane; i < cells; i += 32) covered += m[(top + i / bw) * W + left + i % bw];
      covered = (int)warp_sum((float)covered);
      const int fresh = cells - covered;
      if (fresh > 0 && fresh <= budget) {
        for (int i = lane; i < cells; i += 32) m[(top + i / bw) * W + left + i % bw] = 1;
        __syncwarp();
        added = fresh;
      }
    }
    if (added == 0) break;
    count += added;
  }
  __syncwarp();
  for (int i = lane; i < Np; i += 32) masks[(size_t)crop * Np + i] = m[i];
}

// masks u8 [B, Np] -> idx (flat positions of the masked cells, ascending), weight (1 / masked count of the cell's crop),
// row_w (1 / M), pad (0 for real rows, -1e30 for rows >= M), m_valid.  Rows >= M: idx 0, weight 0, row_w 0.
__global__ void __launch_bounds__(1024)
collate_masks_kernel(const unsigned char* __restrict__ masks, int B, int Np, int cap, long long* __restrict__ idx,
                     float* __restrict__ weight, float* __restrict__ row_w, float* __restrict__ pad, int* __restrict__ m_valid) {
  B200_PDL_SYNC();
  __shared__ int warp_tot[32];
  __shared__ int base;
  extern __shared__ int crop_count[];  // [B]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < B; c += 32) {
    int n = 0;
    for (int i = lane; i < Np; i += 32) n += masks[(size_t)c * Np + i];
    n = (int)warp_sum((float)n);
    if (lane == 0) crop_count[c] = n;
  }
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const long long total = (long long)B * Np;
  for (long long start = 0; start < total; start += 1024) {
    const long long i = start + threadIdx.x;
    const int f = (i < total) ? masks[i] : 0;
    // block exclusive scan of f
    int incl = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = warp_tot[lane], s = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += u;
      }
      warp_tot[lane] = s - v;  // exclusive
    }
    __syncthreads();
    const int pos = base + warp_tot[warp] + incl - f;
    if (f && pos < cap) {
      idx[pos] = i;
      weight[pos] = 1.0f / (float)max(crop_count[(int)(i / Np)], 1);
    }
    __syncthreads();
    if (threadIdx.x == 1023) base = pos + f;
    __syncthreads();
  }
  const int M = min(base, cap);
  for (int r = threadIdx.x; r < cap; r += 1024) {
    const bool real = r < M;
    if (!real) { idx[r] = 0; weight[r] = 0.f; }
    row_w[r] = real ? 1.0f / (float)max(M, 1) : 0.f;
    pad[r] = real ? 0.f : -1e30f;
  }
  if (threadIdx.x == 0) *m_valid = M;
}

// Uniformly random k-subset of {0..n-1} in random order: the k smallest of n counter-based pseudo-random keys, ranked by
// counting in one CTA (n <= 2048: at most 4 M comparisons).  Replaces `torch.randperm(b)[:k]` of
// drop_add_residual_stochastic_depth (LT/_models/dinov2_vit/dinov2_vit_src/layers/block.py:125-127) -- a radix sort plus
// three small launches (17 us) per residual branch -- with one ~2 us launch; *counter advances by one per launch, so CUDA-graph
// replays draw fresh subsets.
__global__ void __launch_bounds__(1024)
random_subset_kernel(int n, int k, unsigned long long seed, long long* __restrict__ counter, long long* __restrict__ idx_out) {
  B200_PDL_SYNC();
  extern __shared__ unsigned long long keys[];
  const unsigned long long ctr = (unsigned long long)*counter;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    Rng rng{seed * 0xD1B54A32D192ED03ull + ctr * 0x9E3779B97F4A7C15ull + (unsigned long long)i * 0xBF58476D1CE4E5B9ull};
    keys[i] = rng.next();
  }
  __syncthreads();  // every thread has read *counter
  if (threadIdx.x == 0) *counter = (long long)(ctr + 1);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long ki = keys[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const unsigned long long kj = keys[j];
      rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
    }
    if (rank < k) idx_out[rank] = i;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_random_subset(int n, int k, long long seed, long long* counter_dev, long long* idx_out, void* stream) {
  if (!counter_dev || !idx_out || n <= 0 || k <= 0 || k > n) return B200_ERR_INVALID_ARG;
  if (n > 2048) return B200_ERR_UNSUPPORTED;
  const int threads = n >= 1024 ? 1024 : ((n + 31) / 32) * 32;
  launch_kernel(random_subset_kernel, 1, threads, (size_t)n * sizeof(unsigned long long), (cudaStream_t)stream, n, k,
                (unsigned long long)seed, counter_dev, idx_out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_block_masks(const int* targets, int B, int H, int W, int min_patches, int max_patches, float min_aspect,
                                float max_aspect, long long seed, const int* step_dev, unsigned char* masks, void* stream) {
  if (!targets || !masks || B <= 0 || H <= 0 || W <= 0 || min_patches < 1 || max_patches < min_patches) return B200_ERR_INVALID_ARG;
  if (!(min_aspect > 0.f) || !(max_aspect >= min_aspect)) return B200_ERR_INVALID_ARG;
  const size_t per_warp = ((size_t)H * W + 15) / 16 * 16;
  if (per_warp * MASK_WARPS > 48 * 1024) return B200_ERR_UNSUPPORTED;
  launch_kernel(block_masks_kernel, (B + MASK_WARPS - 1) / MASK_WARPS, MASK_WARPS * 32, per_warp * MASK_WARPS, (cudaStream_t)stream, 
      targets, B, H, W, min_patches, max_patches, logf(min_aspect), logf(max_aspect), (unsigned long long)seed, step_dev, masks);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

extern "C" int b200_collate_masks(const unsigned char* masks, int B, int Np, int cap, long long* idx, float* weight, float* row_w,
                                  float* pad, int* m_valid, void* stream) {
  if (!masks || !idx || !weight || !row_w || !pad || !m_valid || B <= 0 || Np <= 0 || cap <= 0) return B200_ERR_INVALID_ARG;
  if ((size_t)B * sizeof(int) > 40 * 1024) return B200_ERR_UNSUPPORTED;
  launch_kernel(collate_masks_kernel, 1, 1024, (size_t)B * sizeof(int), (cudaStream_t)stream, masks, B, Np, cap, idx, weight, row_w, pad, m_valid);
  B200_CHECK_LAUNCH();
  return B200_OK;
}
