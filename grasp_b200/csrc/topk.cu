// Per-matrix top-k of singular-value importance scores (radix select + bitonic
// sort of the survivors), and the cumulative-threshold rank selection.
// Replaces torch.topk at reference modeling_grasp.py:404 and
// adaptive_rank_selection at tools/utils_func.py:45-57.
#include "common.cuh"

namespace grasp {

constexpr int TK_THREADS = 1024;
constexpr int TK_MAX_SORT = 16384;  // pow2 capacity of the smem sorter (8 B per slot)

// monotone map float -> uint32 (larger = better); NaN ranks highest like torch.topk
__device__ __forceinline__ uint32_t score_key(float x) {
  if (x != x) return 0xffffffffu;
  x += 0.0f;  // -0 -> +0
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct TopkBatch {
  const float* score[8];
  int64_t* idx[8];
  int r[8];
  int k[8];
};

// block-wide exclusive scan of a packed (lo16 = a, hi16 = b) count over one 1024-thread chunk
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_tot /*[33]*/, uint32_t& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t t = warp_tot[lane];
    uint32_t ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    warp_tot[lane] = ti - t;  // exclusive per-warp offset
    if (lane == 31) warp_tot[32] = ti;
  }
  __syncthreads();
  total = warp_tot[32];
  return warp_tot[w] + inc - v;
}

// sort slots [0,n) (n pow2) by (key desc, idx asc)
__device__ void bitonic_sort_desc(uint32_t* key, int32_t* idx, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);  // first half of each bitonic block sorted "better first"
        const uint32_t ka = key[lo], kb = key[hi];
        const int32_t ia = idx[lo], ib = idx[hi];
        const bool a_before_b = (ka > kb) || (ka == kb && ia < ib);
        if (a_before_b != desc) {
          key[lo] = kb; key[hi] = ka;
          idx[lo] = ib; idx[hi] = ia;
        }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TK_THREADS)
topk_kernel(TopkBatch bt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t hist[256];
  __shared__ uint32_t warp_tot[33];
  __shared__ uint32_t s_prefix, s_need;

  const int m = blockIdx.x;
  const float* __restrict__ score = bt.score[m];
  int64_t* __restrict__ out = bt.idx[m];
  const int r = bt.r[m], k = bt.k[m];
  if (k <= 0) return;
  int kp = 1;
  while (kp < k) kp <<= 1;
  uint32_t* skey = reinterpret_cast<uint32_t*>(smem_raw);
  int32_t* sidx = reinterpret_cast<int32_t*>(skey + kp);

  // ---- radix select: key of the k-th best element --------------------------
  uint32_t prefix = 0, mask = 0, need = (uint32_t)k;
  if (k < r) {
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < r; i += blockDim.x) {
        const uint32_t key = score_key(score[i]);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (threadIdx.x < 32) {
        // lane l owns digits [8l, 8l+8); suffix counts from the top digit down
        const int lane = threadIdx.x;
        uint32_t c[8], tot = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; tot += c[j]; }
        uint32_t above = tot;  // inclusive suffix over lanes >= lane
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t t = __shfl_down_sync(0xffffffffu, above, o);
          if (lane + o < 32) above += t;
        }
        above -= tot;  // count of keys in lanes strictly above
        uint32_t run = above;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
          if (run < need && need <= run + c[j]) {
            s_prefix = prefix | ((uint32_t)(lane * 8 + j) << shift);
            s_need = need - run;
          }
          run += c[j];
        }
      }
      __syncthreads();
      prefix = s_prefix;
      need = s_need;
      mask |= (0xffu << shift);
      __syncthreads();
    }
  }
  // now: take every key > prefix, plus the `need` lowest-index keys == prefix
  const uint32_t thr = prefix;
  const uint32_t n_gt = (uint32_t)k - need;

  // ---- ordered compaction into smem -----------------------------------------
  uint32_t base_gt = 0, base_eq = 0;
  for (int c0 = 0; c0 < r; c0 += TK_THREADS) {
    const int i = c0 + threadIdx.x;
    uint32_t key = 0;
    bool gt = false, eq = false;
    if (i < r) {
      key = score_key(score[i]);
      if (k >= r) gt = true;
      else { gt = key > thr; eq = key == thr; }
    }
    uint32_t total;
    const uint32_t pos = block_excl_scan((gt ? 1u : 0u) | (eq ? 0x10000u : 0u), warp_tot, total);
    if (gt) {
      const uint32_t p = base_gt + (pos & 0xffffu);
      skey[p] = key; sidx[p] = i;
    } else if (eq) {
      const uint32_t e = base_eq + (pos >> 16);
      if (e < need) { skey[n_gt + e] = key; sidx[n_gt + e] = i; }
    }
    base_gt += total & 0xffffu;
    base_eq += total >> 16;
  }
  for (int i = k + threadIdx.x; i < kp; i += blockDim.x) { skey[i] = 0u; sidx[i] = 0x7fffffff; }
  __syncthreads();

  bitonic_sort_desc(skey, sidx, kp);
  for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = (int64_t)sidx[i];
}

// full descending order + shortest prefix reaching target_ratio * sum
__global__ void __launch_bounds__(TK_THREADS)
adaptive_rank_kernel(const float* __restrict__ score, int r, float target_ratio, int64_t* __restrict__ idx,
                     int64_t* __restrict__ count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int rp = 1;
  while (rp < r) rp <<= 1;
  uint32_t* skey = reinterpret_cast<uint32_t*>(smem_raw);
  int32_t* sidx = reinterpret_cast<int32_t*>(skey + rp);
  for (int i = threadIdx.x; i < rp; i += blockDim.x) {
    skey[i] = (i < r) ? score_key(score[i]) : 0u;
    sidx[i] = (i < r) ? i : 0x7fffffff;
  }
  __syncthreads();
  bitonic_sort_desc(skey, sidx, rp);
  for (int i = threadIdx.x; i < r; i += blockDim.x) idx[i] = (int64_t)sidx[i];
  if (threadIdx.x == 0) {
    // reference sums sequentially in fp32 (python sum over a tensor), keep that order
    float total = 0.f;
    for (int i = 0; i < r; ++i) total += score[i];
    const float target = total * target_ratio;
    float run = 0.f;
    int n = 0;
    for (int i = 0; i < r; ++i) {
      run += score[sidx[i]];
      n = i + 1;
      if (run >= target) break;
    }
    *count = (int64_t)n;
  }
}

}  // namespace grasp

using namespace grasp;

extern "C" int grasp_topk_batched(int batch, const float* const* score, const int64_t* r, const int64_t* k,
                                  int64_t* const* idx, void* stream) {
  if (batch < 0) return bad_arg("topk: batch");
  if (batch == 0) return 0;
  if (!score || !r || !k || !idx) return bad_arg("topk: null");
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             TK_MAX_SORT * 8), "topk attr");
    if (rc) return rc;
    attr_set = true;
  }
  for (int b0 = 0; b0 < batch; b0 += 8) {
    TopkBatch bt;
    const int nb = (batch - b0 < 8) ? batch - b0 : 8;
    int kmax = 1;
    for (int j = 0; j < nb; ++j) {
      const int i = b0 + j;
      if (!score[i] || !idx[i]) return bad_arg("topk: null matrix pointer");
      if (r[i] <= 0 || r[i] > 65536) return bad_arg("topk: r must be in [1,65536]");
      if (k[i] < 0 || k[i] > r[i]) return bad_arg("topk: k must be in [0,r]");
      if (k[i] > TK_MAX_SORT) return bad_arg("topk: k > 16384 unsupported");
      bt.score[j] = score[i]; bt.idx[j] = idx[i]; bt.r[j] = (int)r[i]; bt.k[j] = (int)k[i];
      if (k[i] > kmax) kmax = (int)k[i];
    }
    for (int j = nb; j < 8; ++j) { bt.score[j] = nullptr; bt.idx[j] = nullptr; bt.r[j] = 0; bt.k[j] = 0; }
    int kp = 1;
    while (kp < kmax) kp <<= 1;
    GRASP_LAUNCH(topk_kernel, dim3(nb), dim3(TK_THREADS), (size_t)kp * 8, stream, bt);
    GRASP_CHECK_LAST("topk_kernel");
  }
  return 0;
}

extern "C" int grasp_adaptive_rank(const float* score, int64_t r, double target_ratio, int64_t* idx,
                                   int64_t* count, void* stream) {
  if (!score || !idx || !count) return bad_arg("adaptive_rank: null");
  if (r <= 0 || r > TK_MAX_SORT) return bad_arg("adaptive_rank: r must be in [1,16384]");
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(adaptive_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             TK_MAX_SORT * 8), "adaptive attr");
    if (rc) return rc;
    attr_set = true;
  }
  int rp = 1;
  while (rp < r) rp <<= 1;
  GRASP_LAUNCH(adaptive_rank_kernel, dim3(1), dim3(TK_THREADS), (size_t)rp * 8, stream, score, (int)r,
               (float)target_ratio, idx, count);
  GRASP_CHECK_LAST("adaptive_rank_kernel");
  return 0;
}
