// Tensor-core (tcgen05 + TMA) Gram and update steps of the block-Jacobi SVD.
//
// During the tensor-core phase the working matrix Z lives as three bf16 planes (hi, mid, lo:
// Z = p0 + p1 + p2 to ~24 bits) in a TILE-MAJOR layout  Zp[plane][column tile][row][128]: the
// 32-row x 128-column chunk a CTA loads or stores is one contiguous 8 KiB run (row-major Z would give
// 256-byte pieces 16-30 KB apart, which halves the achieved HBM bandwidth).  One CTA tile handles TWO block pairs = 4 chunks of 32 rows
// (I1, J1, I2, J2), so every MMA is the validated 128 x 128 x 16 shape:
//   gram  : G_tile = Yt Yt^T over a K range (A and B descriptors point at the same smem planes);
//           the two 64 x 64 diagonal blocks are the pair Grams (K split over CTAs, partials summed
//           by the eigen-solve kernel in a fixed order)
//   update: Znew_tile[128, 128 cols] = blockdiag(ET1, ET2) * Zt[128 rows, 128 cols], the rows of Z
//           enter as the MN-major B operand straight from the row-major planes; the epilogue writes
//           the new rows as fp32 and as planes, in place.
// Six plane products per K step (bf16x6), per-K-block TMEM accumulators added in fp32 registers
// (see gemm_tc.cu for why).  Included by svd_jacobi.cu only.
#pragma once
#include "tc_common.cuh"

namespace grasp {

using namespace tc;

constexpr int JT_THREADS = 320;            // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quadrant)
constexpr int JT_PLANE_TILE = 128 * 64 * 2;   // 16 KiB: 128 rows (or 64 k-rows x 2 halves) x 64 bf16
constexpr int JT_NACC = 4;

struct JtMaps {
  CUtensorMap z[J_MAXMAT];    // planes [3][ldz/128][rp][128] (4-D), box 64 cols x 32 rows
  CUtensorMap et[J_MAXMAT];   // planes [3][ntiles*128][128], box 64 k x 128 m
};

struct JtMat {
  __nv_bfloat16* Zp;          // planes [3][ldz/128][rp][128]
  float* Gpart;               // [npairs][nsplit][64*64]
  const int* pair_flag;
  const uint32_t* stats;
};

struct JtParams {
  JtMat mat[J_MAXMAT];
  int nmat, rp, Lp, ldz, p, npairs, ntiles, nsplit, round;
};

enum { JT_GRAM = 0, JT_UPDATE = 1, JT_GRAM3 = 2, JT_UPDATE2 = 3 };   // JT_GRAM3: all three planes / six products (clean-up
                                                                    // sweeps); JT_UPDATE2: two planes / three products (early sweeps)

template <int MODE>
struct JtCfg {
  // the Gram only steers the rotations of a phase that stops at 1e-4: two planes / three products
  // (bf16x3, 4e-6) are plenty there; the update keeps all three planes (six products)
  // JT_UPDATE2 (experiment, off by default -- see grasp_svd_batched): the early sweeps read and write two planes only,
  // 8 instead of 12 bytes per element and round of the HBM-bound update.
  static constexpr bool IS_GRAM = (MODE == JT_GRAM || MODE == JT_GRAM3);
  static constexpr bool IS_UPDATE = !IS_GRAM;
  static constexpr int GRAM_PLANES = (MODE == JT_GRAM) ? 2 : 3;
  static constexpr int UPD_PLANES = (MODE == JT_UPDATE2) ? 2 : 3;
  static constexpr int STAGE_BYTES = (IS_GRAM ? GRAM_PLANES : 2 * UPD_PLANES) * JT_PLANE_TILE;
  static constexpr int STAGES = (MODE == JT_GRAM) ? 6 : (MODE == JT_GRAM3 ? 4 : (MODE == JT_UPDATE2 ? 3 : 2));
  static constexpr int STORE_BYTES = IS_UPDATE ? 8 * 4096 : 0;   // one 32 x 64 bf16 staging tile per epilogue warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STORE_BYTES + 1024 + 256;
};

// chunk c (0..3) of tile t -> first row of the 32-row block; pairs beyond npairs alias pair 0 of the tile
__device__ __forceinline__ int jt_chunk_row(const JtParams& p, int tile, int c) {
  int pair = tile * 2 + (c >> 1);
  if (pair >= p.npairs) pair = tile * 2;
  int I, J;
  rr_pair(p.p, p.round, pair, I, J);
  return ((c & 1) ? J : I) * JB;
}

template <int MODE>
__global__ void __launch_bounds__(JT_THREADS, 1)
jacobi_tc_kernel(const __grid_constant__ JtMaps maps, const JtParams p) {
  using Cfg = JtCfg<MODE>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[JT_NACC], acc_empty[JT_NACC];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks_total = p.Lp / 64;
  const int per_split = (kblocks_total + p.nsplit - 1) / p.nsplit;
  const int ncol = p.ldz / 128;
  constexpr bool IS_GRAM = Cfg::IS_GRAM;
  const int per_mat = IS_GRAM ? p.ntiles * p.nsplit : p.ntiles * ncol;
  const int total = p.nmat * per_mat;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < JT_NACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 256); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  // work item decode shared by the three roles (must be identical in all of them)
  auto decode = [&](int w, int& m, int& tile, int& sub, int& kb0, int& kb1) -> bool {
    m = w / per_mat;
    const int r = w - m * per_mat;
    if (IS_GRAM) {
      tile = r / p.nsplit; sub = r - tile * p.nsplit;
      kb0 = sub * per_split; kb1 = min(kblocks_total, kb0 + per_split);
    } else {
      tile = r / ncol; sub = r - tile * ncol;
      kb0 = 0; kb1 = 2;
    }
    if (p.mat[m].stats[0]) return false;                      // matrix already converged
    if (Cfg::IS_UPDATE) {
      const int f0 = p.mat[m].pair_flag[tile * 2];
      const int f1 = (tile * 2 + 1 < p.npairs) ? p.mat[m].pair_flag[tile * 2 + 1] : 0;
      if (!f0 && !f1) return false;                           // both rotations are the identity
    }
    return kb1 > kb0 || IS_GRAM;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        int m, tile, sub, kb0, kb1;
        if (!decode(w, m, tile, sub, kb0, kb1)) continue;
        const CUtensorMap* zmap = &maps.z[m];
        int rows[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) rows[c] = jt_chunk_row(p, tile, c);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* st = smem + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (IS_GRAM) {
#pragma unroll
            for (int pl = 0; pl < Cfg::GRAM_PLANES; ++pl)
#pragma unroll
              for (int c = 0; c < 4; ++c)
                tma_load_4d(st + pl * JT_PLANE_TILE + c * 4096, zmap, &full_bar[stage], (kb & 1) * 64, rows[c], kb >> 1, pl);
          } else {
            unsigned char* sB = st + Cfg::UPD_PLANES * JT_PLANE_TILE;
#pragma unroll
            for (int pl = 0; pl < Cfg::UPD_PLANES; ++pl) {
              // A: ET planes, 128 m x 64 k of K block kb
              tma_load_3d(st + pl * JT_PLANE_TILE, &maps.et[m], &full_bar[stage], kb * 64, tile * 128, pl);
              // B: rows of Z as K (two 32-row chunks per K block), 128 columns as two 64-wide halves
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc)
                  tma_load_4d(sB + pl * JT_PLANE_TILE + h * 8192 + cc * 4096, zmap, &full_bar[stage], h * 64,
                              rows[kb * 2 + cc], sub, pl);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, Cfg::IS_UPDATE ? 1 : 0);
      // plane products, small terms first; two-plane modes use planes {0,1} only: (1,0) (0,1) (0,0)
      constexpr bool TWO = (MODE == JT_GRAM || MODE == JT_UPDATE2);
      constexpr int NPROD = TWO ? 3 : 6;
      constexpr int PA[6] = {TWO ? 1 : 2, 0, TWO ? 0 : 1, 1, 0, 0};
      constexpr int PB[6] = {0, TWO ? 1 : 2, TWO ? 0 : 1, 0, 1, 0};
      constexpr bool IS_GRAM_ = Cfg::IS_GRAM;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        int m, tile, sub, kb0, kb1;
        if (!decode(w, m, tile, sub, kb0, kb1)) continue;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&acc_empty[acc], acc_phase ^ 1);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
          const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sB = IS_GRAM_ ? sA : sA + Cfg::UPD_PLANES * JT_PLANE_TILE;
#pragma unroll
          for (int q = 0; q < NPROD; ++q) {
            const uint64_t da = umma_desc_kmajor_sw128(sA + PA[q] * JT_PLANE_TILE);
            const uint64_t db = IS_GRAM_ ? umma_desc_kmajor_sw128(sB + PB[q] * JT_PLANE_TILE)
                                                  : umma_desc_mnmajor_sw128(sB + PB[q] * JT_PLANE_TILE, 8192, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t bstep = IS_GRAM_ ? (uint64_t)(2 * k) : (uint64_t)(128 * k);
              umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + bstep, idesc, (q | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&acc_full[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++acc == JT_NACC) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    const int quad = warp & 3;                     // TMEM lane quadrant
    const int half = (warp - 2) >> 2;              // columns [64*half, 64*half + 64) of the tile
    const int mrow = quad * 32 + lane;             // row of the 128-row tile owned by this thread
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      int m, tile, sub, kb0, kb1;
      if (!decode(w, m, tile, sub, kb0, kb1)) continue;
      float racc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) racc[j] = 0.f;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after_sync();
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 128 + half * 64);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float t[32];
          tmem_ld_32x32(t_row + (uint32_t)(c * 32), t);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[c * 32 + j] += t[j];
        }
        tc_fence_before_sync();
        mbar_arrive(&acc_empty[acc]);
        if (++acc == JT_NACC) { acc = 0; acc_phase ^= 1; }
      }
      const JtMat& M = p.mat[m];
      const int pair = tile * 2 + (mrow >> 6);
      if (pair >= p.npairs) continue;               // second half of an odd last tile
      if (IS_GRAM) {
        // the 64 x 64 diagonal block of this row's pair lives in the column half equal to the row half
        if (half == (mrow >> 6)) {
          float* out = M.Gpart + ((int64_t)pair * p.nsplit + sub) * (JS * JS) + (mrow & 63) * JS;
#pragma unroll
          for (int j = 0; j < 64; j += 4)
            *reinterpret_cast<float4*>(out + j) = make_float4(racc[j], racc[j + 1], racc[j + 2], racc[j + 3]);
        }
      } else {
        // x = p0 + p1 + p2 with p0, p1 the TRUNCATED bf16 of the running remainder (remainders are exact
        // in fp32) and p2 the truncated rest: |x - sum| <= 2^-24 |x|; a plane word is just the two high
        // halves of consecutive fp32 bit patterns (one PRMT).  Each warp stages its 32 rows x 64 columns
        // of one plane in shared memory (128-byte swizzle of the tensor map) and one lane writes the tile
        // with a TMA store: full 128-byte lines instead of 32 scattered 16-byte pieces per instruction.
        unsigned char* stg = smem + STAGES * Cfg::STAGE_BYTES + (warp - 2) * 4096;
        const int row0 = jt_chunk_row(p, tile, quad);       // this warp's 32 rows are chunk `quad` of the tile
#pragma unroll
        for (int pl = 0; pl < Cfg::UPD_PLANES; ++pl) {
          if (lane == 0) tma_store_wait_read();              // the previous store has finished reading `stg`
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 64; j += 8) {
            uint32_t wv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x0 = racc[j + 2 * e], x1 = racc[j + 2 * e + 1];
              if (pl >= 1) {
                x0 -= __uint_as_float(__float_as_uint(x0) & 0xffff0000u);
                x1 -= __uint_as_float(__float_as_uint(x1) & 0xffff0000u);
              }
              if (pl == 2) {
                x0 -= __uint_as_float(__float_as_uint(x0) & 0xffff0000u);
                x1 -= __uint_as_float(__float_as_uint(x1) & 0xffff0000u);
              }
              wv[e] = __byte_perm(__float_as_uint(x0), __float_as_uint(x1), 0x7632);
            }
            // 16-byte chunk j/8 of row `lane`, XOR-swizzled with the row index (Swizzle<3,4,3>)
            *reinterpret_cast<uint4*>(stg + lane * 128 + (((j >> 3) ^ (lane & 7)) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&maps.z[m], stg, half * 64, row0, sub, pl);
            tma_store_commit();
          }
        }
      }
    }
  }

  if (Cfg::IS_UPDATE && warp >= 2 && lane == 0) tma_store_wait_all();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// fp32 Z [rp][ldz] -> three bf16 planes, tile-major [plane][ldz/128][rp][128] (once after init)
__global__ void jt_split_kernel(const float* __restrict__ Z, int rp, int ldz, __nv_bfloat16* __restrict__ Zp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 per thread
  const int64_t n4 = (int64_t)rp * ldz / 4;
  if (i >= n4) return;
  const int row = (int)(i / (ldz / 4)), col = (int)(i % (ldz / 4)) * 4;
  const float4 v = reinterpret_cast<const float4*>(Z)[i];
  float x[4] = {v.x, v.y, v.z, v.w};
  uint16_t h[3][4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      const __nv_bfloat16 b = __float2bfloat16_rn(x[e]);
      h[pl][e] = __bfloat16_as_ushort(b);
      x[e] -= __bfloat162float(b);
    }
  }
  const int64_t plane = (int64_t)rp * ldz;
  const int64_t dst = ((int64_t)(col >> 7) * rp + row) * 128 + (col & 127);
#pragma unroll
  for (int pl = 0; pl < 3; ++pl) {
    uint2 w;
    w.x = (uint32_t)h[pl][0] | ((uint32_t)h[pl][1] << 16);
    w.y = (uint32_t)h[pl][2] | ((uint32_t)h[pl][3] << 16);
    *reinterpret_cast<uint2*>(Zp + pl * plane + dst) = w;
  }
}

// planes -> fp32 Z (end of the tensor-core phase)
__global__ void jt_merge_kernel(const __nv_bfloat16* __restrict__ Zp, int rp, int ldz, float* __restrict__ Z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = (int64_t)rp * ldz / 4;
  if (i >= n4) return;
  const int row = (int)(i / (ldz / 4)), col = (int)(i % (ldz / 4)) * 4;
  const int64_t plane = (int64_t)rp * ldz;
  const int64_t src = ((int64_t)(col >> 7) * rp + row) * 128 + (col & 127);
  float x[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int pl = 2; pl >= 0; --pl) {     // small parts first
    const uint2 w = *reinterpret_cast<const uint2*>(Zp + pl * plane + src);
    x[0] += __uint_as_float(w.x << 16); x[1] += __uint_as_float(w.x & 0xffff0000u);
    x[2] += __uint_as_float(w.y << 16); x[3] += __uint_as_float(w.y & 0xffff0000u);
  }
  reinterpret_cast<float4*>(Z)[i] = make_float4(x[0], x[1], x[2], x[3]);
}

}  // namespace grasp
