// fp32-accurate GEMM on the sm_100a tensor cores.
//
// fp32 operands are split into NS bf16 planes (x = p0 + p1 (+ p2), each plane the bf16
// rounding of the running remainder) by a pre-pass that also puts both operands in
// "K-major" form  Ap[plane][m][k], Bp[plane][n][k].  The GEMM core is then
//     C[m,n] = sum_k sum_{(i,j) in PRODUCTS} Ap[i][m,k] * Bp[j][n,k]
// with PRODUCTS = {00,01,10} (NS=2, "BF16X3", rel. err ~4e-6) or
// {00,01,10,11,02,20} (NS=3, "BF16X6", ~1e-7), every product a tcgen05.mma (kind::f16,
// bf16 in, fp32 accumulate in TMEM) fed by TMA through a shared-memory ring.
//
// Kernel shape: persistent, one CTA per SM, 192 threads:
//   warp 0   TMA producer (one elected lane)
//   warp 1   TMEM allocator + MMA issuer (one elected lane)
//   warps 2-5 epilogue (TMEM -> registers -> global)
// The tensor core adds into its fp32 accumulator with truncation, so a long accumulation chain
// drifts (measured: 2.8e-5 relative at K=4096 with 6 products).  Each 64-wide K block is therefore
// accumulated into a fresh TMEM buffer (small correction products first, the hi*hi product last) and
// the epilogue warps add the block result into fp32 registers with round-to-nearest.  The TMEM
// buffers form a ring (512 / BN deep), so the MMA warp never waits for the epilogue.
#include "tc_common.cuh"
#include "split_f16.cuh"
#include <stdlib.h>

namespace grasp {

using namespace tc;

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;           // 64 bf16 = 128 bytes = one swizzle row
constexpr int TC_THREADS = 192;
constexpr int TC_A_TILE = TC_BM * TC_BK * 2;   // 16 KiB per plane

enum { EPI_STORE = 0, EPI_SIGMA = 1, EPI_PLANES = 2 };

struct TcParams {
  int M, N, K;
  int tiles_m, tiles_n;
  float alpha, beta;
  void* C;            // EPI_STORE: [M][ldc] fp32 or bf16
  int64_t ldc;
  int c_bf16;
  const float* Umul;  // EPI_SIGMA: [M][ldu], multiplied element-wise before the column reduction
  int64_t ldu;
  float* partial;     // EPI_SIGMA: [tiles_m][N]
  const float* inv_sa;  // F16 planes: per-row inverse scale of A (M) and B (N); the accumulator is
  const float* inv_sb;  //   multiplied by inv_sa[m] * inv_sb[n] before the epilogue
  int kgroup;           // K blocks accumulated in one TMEM buffer before the epilogue drains it (0 = 1)
  // EPI_PLANES: the result leaves as a prepared operand (fp16 hi / lo planes [2][M][out_pitch] + row scales) instead
  // of fp32: the rank-k intermediate of a factor pair (SVDLinear, reference modeling_grasp.py:57-59) feeds the next
  // GEMM directly.  The row scale comes from a bound, not from the row maximum (a tile sees only BN columns):
  // |acc| <= K * 2^30 in plane units, so acc * out_scale with out_scale = 2^-16 / pow2ceil(K) stays below 2^14;
  // out_inv[row] = inv_sa[row] * inv_sb[0] / out_scale (exact powers of two).  Needs a tensor-scaled B.
  uint16_t* out_planes;
  int64_t out_pitch, out_plane_stride;
  float* out_inv;
  float out_scale;
  int group_m;          // tile rasterisation: M blocks per super-row (see tile_coords)
  int staged;           // EPI_STORE fp32: transpose through shared memory, 128-byte row segments per store
};

// Tile order of the persistent loop.  Consecutive tile indices run at the same time on different SMs, so they
// should share operand rows: tiles are walked in super-rows of `group_m` M blocks, M fastest inside a super-row,
// then along N.  The A rows of a super-row (group_m x 128 x K planes) stay in L2 while the B tiles stream past
// once per super-row.  M-fastest over ALL M blocks (the round-1 order) re-read every A tile once per wave:
// ncu counted 3.17 GB of DRAM traffic for 0.68 GB of operands + output (profiles/r01_ncu_traffic.json).
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int group_m, int& m_blk, int& n_blk) {
  const int per_group = group_m * tiles_n;
  const int g = tile / per_group;
  const int first = g * group_m;
  const int rows = min(group_m, tiles_m - first);     // the last super-row may be shorter
  const int t = tile - g * per_group;
  m_blk = first + t % rows;
  n_blk = t / rows;
}

template <int NS, int BN>
struct TcCfg {
  static constexpr int B_TILE = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = NS * (TC_A_TILE + B_TILE);
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;
  static constexpr int NACC = 512 / BN;       // ring of per-K-block accumulators in TMEM
  static constexpr int TMEM_COLS = 512;
  // epilogue scratch: [4][BN] floats of the sigma reduction, or one 32 x 32 fp32 staging tile per epilogue warp
  static constexpr int EPI_SMEM = (BN / 32) * 4096;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + EPI_SMEM + 256;
  static constexpr int EPI_THREADS = BN;      // 4 epilogue warps per 128 output columns (128 fp32 accumulators per thread)
  static constexpr int THREADS = 64 + EPI_THREADS;
  static_assert(STAGES >= 2, "need at least a double-buffered ring");
  static_assert(BN == 128 || BN == 256, "supported tile widths");
};

// BMN = 1: the B planes are stored [plane][K][N] (N contiguous, i.e. op(B) as given when tb = 0) and are
// fed to the tensor core as an MN-major operand: no transposing pre-pass.
// F16 = 1: the planes hold fp16 (hi, lo) of row-scaled operands (2 planes, 3 products, ~3e-7).
template <int NS, int BN, int EPI, int BMN, int F16>
__global__ void __launch_bounds__(64 + BN, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcParams p) {
  using Cfg = TcCfg<NS, BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment for the 128-byte swizzle atoms
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  float* red = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);   // [4][BN]
  constexpr int NACC = Cfg::NACC;
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[NACC], acc_empty[NACC];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = (p.K + TC_BK - 1) / TC_BK;
  // The epilogue reads a whole 128 x BN fp32 accumulator per drain and TMEM reads run at 64 B/clk: at one drain
  // per 64-wide K block that is 2048 clk against 1536 clk of MMAs (BN = 256, three products), i.e. the kernel is
  // TMEM-read-bound at ~75 % tensor-pipe activity.  `kgroup` K blocks share one accumulator before it is drained.
  const int kgroup = p.kgroup > 0 ? p.kgroup : 1;
  const int kgroups = (kblocks + kgroup - 1) / kgroup;
  const int total_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < NACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], Cfg::EPI_THREADS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, p.tiles_m, p.tiles_n, p.group_m, m_blk, n_blk);
        const int m0 = m_blk * TC_BM, n0 = n_blk * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sA = smem + stage * Cfg::STAGE_BYTES;
          unsigned char* sB = sA + NS * TC_A_TILE;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
          for (int pl = 0; pl < NS; ++pl) {
            tma_load_3d(sA + pl * TC_A_TILE, &mapA, &full_bar[stage], kb * TC_BK, m0, pl);
            if constexpr (BMN) {
              // box = 64 n x 64 k; one box per 64-wide N chunk, chunks TC_BK*128 bytes apart
#pragma unroll
              for (int h = 0; h < BN / 64; ++h)
                tma_load_3d(sB + pl * Cfg::B_TILE + h * (TC_BK * 128), &mapB, &full_bar[stage], n0 + h * 64,
                            kb * TC_BK, pl);
            } else {
              tma_load_3d(sB + pl * Cfg::B_TILE, &mapB, &full_bar[stage], kb * TC_BK, n0, pl);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, BN, 0, BMN, F16);
      // plane products, small terms first
      constexpr int NPROD = (NS == 2) ? 3 : 6;
      constexpr int PA[6] = {NS == 2 ? 1 : 2, NS == 2 ? 0 : 0, NS == 2 ? 0 : 1, 1, 0, 0};
      constexpr int PB[6] = {NS == 2 ? 0 : 0, NS == 2 ? 1 : 2, NS == 2 ? 0 : 1, 0, 1, 0};
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int in_group = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          if (in_group == 0) mbar_wait(&acc_empty[acc], acc_phase ^ 1);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sB = sA + NS * TC_A_TILE;
#pragma unroll
          for (int q = 0; q < NPROD; ++q) {
            const uint64_t da = umma_desc_kmajor_sw128(sA + PA[q] * TC_A_TILE);
            const uint64_t db = BMN ? umma_desc_mnmajor_sw128(sB + PB[q] * Cfg::B_TILE, TC_BK * 128, 1024)
                                    : umma_desc_kmajor_sw128(sB + PB[q] * Cfg::B_TILE);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              // K-major: +32 bytes per 16-element K step inside the 128-byte swizzle row;
              // MN-major: 16 K rows = two 8-row groups = +2048 bytes
              const uint64_t kb_step = BMN ? (uint64_t)(128 * k) : (uint64_t)(2 * k);
              umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + kb_step, idesc, (in_group | q | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);           // frees the smem stage when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++in_group == kgroup || kb == kblocks - 1) {
            umma_commit(&acc_full[acc]);            // this group's partial product is complete
            in_group = 0;
            if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;               // which 128-column half of the tile this warp drains
    const int ep_tid = (int)threadIdx.x - 64;       // 0..EPI_THREADS-1
    int acc = 0; uint32_t acc_phase = 0;
    float* stg = red + (warp - 2) * 1024;          // this warp's 32 x 32 staging tile (EPI_STORE)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, p.tiles_m, p.tiles_n, p.group_m, m_blk, n_blk);
      const int row = m_blk * TC_BM + quad * 32 + lane;
      const int n0 = n_blk * BN;
      float racc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) racc[j] = 0.f;
      for (int grp = 0; grp < kgroups; ++grp) {
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after_sync();
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * 128);
        if constexpr (BN == 128) {
          // two TMEM loads in flight per wait (the second one's latency hides behind the first one's adds);
          // the 320-thread BN = 256 variant has no registers to spare for it (168 per thread)
#pragma unroll
          for (int c = 0; c < 4; c += 2) {
            float t0[32], t1[32];
            tmem_ld_32x32(t_row + (uint32_t)(c * 32), t0);
            tmem_ld_32x32(t_row + (uint32_t)(c * 32 + 32), t1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) racc[c * 32 + j] += t0[j];   // round-to-nearest fp32 add
#pragma unroll
            for (int j = 0; j < 32; ++j) racc[c * 32 + 32 + j] += t1[j];
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float t[32];
            tmem_ld_32x32(t_row + (uint32_t)(c * 32), t);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) racc[c * 32 + j] += t[j];   // round-to-nearest fp32 add
          }
        }
        tc_fence_before_sync();
        mbar_arrive(&acc_empty[acc]);               // hand the TMEM buffer back to the MMA warp
        if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
      }
      const bool staged = (EPI == EPI_STORE) && p.staged;
      float row_scale = p.alpha;                    // staged path: alpha and the row's inverse plane scale in one factor
      if constexpr (F16) {
        // undo the per-row power-of-two scaling of the fp16 planes (exact)
        const float ia = (row < p.M) ? p.inv_sa[row] : 0.f;
        if (staged || EPI == EPI_PLANES) {
          row_scale *= ia;                          // the column scales are applied after the transpose (8 vector loads
        } else {                                    // per lane and tile instead of 128 scalar ones)
#pragma unroll
          for (int j = 0; j < 128; ++j) {
            const int col = n0 + half * 128 + j;
            racc[j] *= ia * ((col < p.N) ? __ldg(p.inv_sb + col) : 0.f);
          }
        }
      }
      if constexpr (EPI == EPI_PLANES) {
        // (F16 only) planes of acc * out_scale: the plane scales of A and B stay folded into out_inv[row]
        if (row < p.M) {
          if (n_blk == 0 && half == 0) p.out_inv[row] = p.inv_sa[row] * p.inv_sb[0] / p.out_scale;
          uint16_t* hi_row = p.out_planes + (int64_t)row * p.out_pitch;
#pragma unroll
          for (int c8 = 0; c8 < 16; ++c8) {
            const int col = n0 + half * 128 + c8 * 8;
            if (col < p.out_pitch) {            // columns in [N, pitch) hold exact zeros (zero-filled B rows)
              uint32_t h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = racc[c8 * 8 + 2 * e] * p.out_scale, b2 = racc[c8 * 8 + 2 * e + 1] * p.out_scale;
                const __half2 h2 = __floats2half2_rn(a, b2);
                const float2 f = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(a - f.x, b2 - f.y);
                h[e] = *reinterpret_cast<const uint32_t*>(&h2);
                l[e] = *reinterpret_cast<const uint32_t*>(&l2);
              }
              *reinterpret_cast<uint4*>(hi_row + col) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4*>(hi_row + p.out_plane_stride + col) = make_uint4(l[0], l[1], l[2], l[3]);
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if constexpr (EPI == EPI_PLANES) break;
        float* v = &racc[c * 32];
        const int col0 = n0 + half * 128 + c * 32;
        if constexpr (EPI == EPI_STORE) {
          if (staged) {
            // Each lane holds 32 consecutive columns of ITS row: stored directly, one instruction writes 16 bytes
            // into 32 different rows (32 half-filled sectors).  Staged through shared memory (float4 chunks
            // XOR-swizzled with the row, conflict-free both ways) the warp writes 4 rows x 128 contiguous bytes
            // per instruction.  N % 4 == 0 here, so a float4 is inside or outside the matrix as a whole.
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) =
                  make_float4(row_scale * v[4 * q], row_scale * v[4 * q + 1], row_scale * v[4 * q + 2], row_scale * v[4 * q + 3]);
            __syncwarp();
            const int sub = lane >> 3, pos = lane & 7;
            const int row_base = m_blk * TC_BM + quad * 32;
            // row r = 4 i + sub of the staging tile holds column chunk pos ^ (r & 7) at position pos: a lane sees two
            // chunks only, (pos ^ sub) for even i and (pos ^ sub ^ 4) for odd i
            const int colA = col0 + ((pos ^ sub) << 2), colB = col0 + ((pos ^ sub ^ 4) << 2);
            float4 csA = make_float4(1.f, 1.f, 1.f, 1.f), csB = csA;
            if constexpr (F16) {
              csA = (colA < p.N) ? __ldg(reinterpret_cast<const float4*>(p.inv_sb + colA)) : make_float4(0.f, 0.f, 0.f, 0.f);
              csB = (colB < p.N) ? __ldg(reinterpret_cast<const float4*>(p.inv_sb + colB)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const bool full = (row_base + 32 <= p.M) && (col0 + 32 <= p.N);   // warp-uniform
            float* crow = static_cast<float*>(p.C) + (int64_t)(row_base + sub) * p.ldc;
            if (full && p.beta == 0.f) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float4 o = *reinterpret_cast<const float4*>(stg + (i * 4 + sub) * 32 + (pos << 2));
                const float4 cs = (i & 1) ? csB : csA;
                o.x *= cs.x; o.y *= cs.y; o.z *= cs.z; o.w *= cs.w;
                *reinterpret_cast<float4*>(crow + (int64_t)(i * 4) * p.ldc + ((i & 1) ? colB : colA)) = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + sub;
                const int col = (i & 1) ? colB : colA;
                float4 o = *reinterpret_cast<const float4*>(stg + r * 32 + (pos << 2));
                const float4 cs = (i & 1) ? csB : csA;
                o.x *= cs.x; o.y *= cs.y; o.z *= cs.z; o.w *= cs.w;
                if (row_base + r < p.M && col < p.N) {
                  float* dst = crow + (int64_t)(i * 4) * p.ldc + col;
                  if (p.beta != 0.f) {
                    const float4 old = *reinterpret_cast<const float4*>(dst);
                    o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
                  }
                  *reinterpret_cast<float4*>(dst) = o;
                }
              }
            }
            __syncwarp();
          } else if (row < p.M && col0 < p.N) {
            const int ncols = min(32, p.N - col0);
            if (p.c_bf16) {
              __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.C) + (int64_t)row * p.ldc + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j < ncols) {
                  float x = p.alpha * v[j];
                  if (p.beta != 0.f) x += p.beta * __bfloat162float(dst[j]);
                  dst[j] = __float2bfloat16_rn(x);
                }
              }
            } else {
              float* dst = static_cast<float*>(p.C) + (int64_t)row * p.ldc + col0;
              const bool vec = (ncols == 32) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
              if (vec) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 o = make_float4(p.alpha * v[j], p.alpha * v[j + 1], p.alpha * v[j + 2], p.alpha * v[j + 3]);
                  if (p.beta != 0.f) {
                    const float4 old = *reinterpret_cast<const float4*>(dst + j);
                    o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
                  }
                  *reinterpret_cast<float4*>(dst + j) = o;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (j < ncols) {
                    float x = p.alpha * v[j];
                    if (p.beta != 0.f) x += p.beta * dst[j];
                    dst[j] = x;
                  }
                }
              }
            }
          }
        } else {
          // multiply by U[row][col] and reduce over the 32 rows of this warp
          const float* u = p.Umul + (int64_t)row * p.ldu + col0;
          const bool row_ok = row < p.M;
          const bool vec = row_ok && (col0 + 32 <= p.N) && ((reinterpret_cast<uintptr_t>(u) & 15) == 0);
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 uu = *reinterpret_cast<const float4*>(u + j);
              v[j] *= uu.x; v[j + 1] *= uu.y; v[j + 2] *= uu.z; v[j + 3] *= uu.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (row_ok && col0 + j < p.N) ? v[j] * u[j] : 0.f;
          }
          // transpose-reduce: afterwards v[0] of lane l is the sum over lanes of column l
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
            for (int j = 0; j < o; ++j) {
              const bool up = (lane & o) != 0;
              const float send = up ? v[j] : v[j + o];
              const float keep = up ? v[j + o] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          red[quad * BN + half * 128 + c * 32 + lane] = v[0];
        }
      }
      if constexpr (EPI == EPI_SIGMA) {
        asm volatile("bar.sync 1, %0;" ::"n"(Cfg::EPI_THREADS) : "memory");
        for (int j = ep_tid; j < BN; j += Cfg::EPI_THREADS) {
          const int col = n0 + j;
          if (col < p.N)
            p.partial[(int64_t)m_blk * p.N + col] = (red[j] + red[BN + j]) + (red[2 * BN + j] + red[3 * BN + j]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(Cfg::EPI_THREADS) : "memory");
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) of the two-plane arithmetics: a cluster of two CTAs (two SMs of one
// TPC) computes a 256 x 256 output tile.  Each CTA loads only ITS 128 rows of A and ITS 128 columns of
// B (half the B bytes per SM of the single-CTA kernel), the leader CTA issues 256 x 256 x 16 MMAs that
// read both CTAs' shared memory and write both CTAs' TMEM; each CTA drains its own 128 accumulator rows.
// Barriers: TMA bytes of both CTAs are counted on the leader's `full` barrier; tcgen05.commit multicasts
// the `empty` / `acc_full` arrivals to both CTAs; the epilogue warps of both CTAs arrive on the leader's
// `acc_empty`.
// ---------------------------------------------------------------------------
struct Tc2Cfg {
  static constexpr int NS = 2;
  static constexpr int TILE = 128 * TC_BK * 2;             // 16 KiB: 128 rows x 64 x 16 bit
  static constexpr int STAGE_BYTES = NS * 2 * TILE;        // A half + B half per CTA: 64 KiB
  static constexpr int STAGES = 3;
  static constexpr int NACC = 2;                           // 2 x 256 TMEM columns
  static constexpr int THREADS = 64 + 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 4 * 256 * 4 + 256;
};

template <int EPI, int BMN, int F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc2Cfg::THREADS, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcParams p) {
  using Cfg = Tc2Cfg;
  constexpr int STAGES = Cfg::STAGES, NACC = Cfg::NACC, NS = Cfg::NS, BN = 256;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  float* red = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);   // [4][256]
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[NACC], acc_empty[NACC];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                   // 0 = leader
  const int kblocks = (p.K + TC_BK - 1) / TC_BK;
  const int tiles_m2 = (p.M + 255) / 256;
  const int total_tiles = tiles_m2 * p.tiles_n;               // p.tiles_n counts 256-wide column tiles
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < NACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 16); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(&tmem_base_smem);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        const int m0 = (tile % tiles_m2) * 256 + (int)rank * 128;
        const int n0 = (tile / tiles_m2) * BN + (int)rank * 128;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sA = smem + stage * Cfg::STAGE_BYTES;
          unsigned char* sB = sA + NS * Cfg::TILE;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          else mbar_arrive_cluster(mapa_u32(smem_u32(&full_bar[stage]), 0));
#pragma unroll
          for (int pl = 0; pl < NS; ++pl) {
            tma_load_3d_2sm(sA + pl * Cfg::TILE, &mapA, &full_bar[stage], kb * TC_BK, m0, pl);
            if constexpr (BMN) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
                tma_load_3d_2sm(sB + pl * Cfg::TILE + h * (TC_BK * 128), &mapB, &full_bar[stage], n0 + h * 64,
                                kb * TC_BK, pl);
            } else {
              tma_load_3d_2sm(sB + pl * Cfg::TILE, &mapB, &full_bar[stage], kb * TC_BK, n0, pl);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, BMN, F16);
      constexpr int PA[3] = {1, 0, 0};
      constexpr int PB[3] = {0, 1, 0};
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&acc_empty[acc], acc_phase ^ 1);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sB = sA + NS * Cfg::TILE;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const uint64_t da = umma_desc_kmajor_sw128(sA + PA[q] * Cfg::TILE);
            const uint64_t db = BMN ? umma_desc_mnmajor_sw128(sB + PB[q] * Cfg::TILE, TC_BK * 128, 1024)
                                    : umma_desc_kmajor_sw128(sB + PB[q] * Cfg::TILE);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint64_t kb_step = BMN ? (uint64_t)(128 * k) : (uint64_t)(2 * k);
              umma_2sm(d_tmem, da + (uint64_t)(2 * k), db + kb_step, idesc, (q | k) != 0);
            }
          }
          umma_commit_2sm(&empty_bar[stage], 3);   // both CTAs may refill this stage
          umma_commit_2sm(&acc_full[acc], 3);      // both CTAs' epilogues may drain this accumulator
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9 of both CTAs)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int ep_tid = (int)threadIdx.x - 64;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs) {
      const int m_blk2 = tile % tiles_m2, n_blk = tile / tiles_m2;
      const int row = m_blk2 * 256 + (int)rank * 128 + quad * 32 + lane;
      const int n0 = n_blk * BN;
      float racc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) racc[j] = 0.f;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after_sync();
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t[32];
          tmem_ld_32x32(t_row + (uint32_t)(c * 32), t);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[c * 32 + j] += t[j];
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[acc]), 0));   // leader's barrier
        if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
      }
      if constexpr (F16) {
        const float ia = (row < p.M) ? p.inv_sa[row] : 0.f;
#pragma unroll
        for (int j = 0; j < 128; ++j) {
          const int col = n0 + half * 128 + j;
          racc[j] *= ia * ((col < p.N) ? __ldg(p.inv_sb + col) : 0.f);
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float* v = &racc[c * 32];
        const int col0 = n0 + half * 128 + c * 32;
        if constexpr (EPI == EPI_STORE) {
          if (row < p.M && col0 < p.N) {
            const int ncols = min(32, p.N - col0);
            if (p.c_bf16) {
              __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.C) + (int64_t)row * p.ldc + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j < ncols) {
                  float x = p.alpha * v[j];
                  if (p.beta != 0.f) x += p.beta * __bfloat162float(dst[j]);
                  dst[j] = __float2bfloat16_rn(x);
                }
              }
            } else {
              float* dst = static_cast<float*>(p.C) + (int64_t)row * p.ldc + col0;
              const bool vec = (ncols == 32) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
              if (vec) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 o = make_float4(p.alpha * v[j], p.alpha * v[j + 1], p.alpha * v[j + 2], p.alpha * v[j + 3]);
                  if (p.beta != 0.f) {
                    const float4 old = *reinterpret_cast<const float4*>(dst + j);
                    o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
                  }
                  *reinterpret_cast<float4*>(dst + j) = o;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (j < ncols) {
                    float x = p.alpha * v[j];
                    if (p.beta != 0.f) x += p.beta * dst[j];
                    dst[j] = x;
                  }
                }
              }
            }
          }
        } else {
          const float* u = p.Umul + (int64_t)row * p.ldu + col0;
          const bool row_ok = row < p.M;
          const bool vec = row_ok && (col0 + 32 <= p.N) && ((reinterpret_cast<uintptr_t>(u) & 15) == 0);
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 uu = *reinterpret_cast<const float4*>(u + j);
              v[j] *= uu.x; v[j + 1] *= uu.y; v[j + 2] *= uu.z; v[j + 3] *= uu.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (row_ok && col0 + j < p.N) ? v[j] * u[j] : 0.f;
          }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
            for (int j = 0; j < o; ++j) {
              const bool up = (lane & o) != 0;
              const float send = up ? v[j] : v[j + o];
              const float keep = up ? v[j + o] : v[j];
              v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          red[quad * BN + half * 128 + c * 32 + lane] = v[0];
        }
      }
      if constexpr (EPI == EPI_SIGMA) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int j = ep_tid; j < BN; j += 256) {
          const int col = n0 + j;
          if (col < p.N)   // one partial row per 128-row half of the pair tile
            p.partial[(int64_t)(m_blk2 * 2 + (int)rank) * p.N + col] =
                (red[j] + red[BN + j]) + (red[2 * BN + j] + red[3 * BN + j]);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
  }

  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------
// split pre-pass: fp32 [rows x cols] (or its transpose) -> NS bf16 planes [NS][R][Kp]
// ---------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16* out) {
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    out[i] = h;
    x -= __bfloat162float(h);
  }
}

// planes[pl][r][k] = part_pl(src[r*ld + k]),  r < R, k < K   (no transpose; 4 elements per thread)
template <int NS>
__global__ void split_rows_kernel(const float* __restrict__ src, int64_t ld, int R, int K, int Kp,
                                  __nv_bfloat16* __restrict__ planes) {
  const int r = blockIdx.x;                                   // rows on grid.x (may exceed 65535)
  const int k0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (k0 >= Kp) return;
  float x[4];
  const float* s = src + (int64_t)r * ld + k0;
  if (k0 + 3 < K && ((reinterpret_cast<uintptr_t>(s) & 15) == 0)) {
    const float4 t = *reinterpret_cast<const float4*>(s);
    x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = (k0 + j < K) ? s[j] : 0.f;
  }
  __nv_bfloat16 parts[4][NS];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_bf16<NS>(x[j], parts[j]);
#pragma unroll
  for (int pl = 0; pl < NS; ++pl) {
    __nv_bfloat16* d = planes + ((int64_t)pl * R + r) * Kp + k0;   // Kp % 8 == 0 and k0 % 4 == 0 -> 8-byte aligned
    const __nv_bfloat162 lo = __halves2bfloat162(parts[0][pl], parts[1][pl]);
    const __nv_bfloat162 hi = __halves2bfloat162(parts[2][pl], parts[3][pl]);
    uint2 w;
    w.x = *reinterpret_cast<const uint32_t*>(&lo);
    w.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(d) = w;
  }
}

// planes[pl][r][k] = part_pl(src[k*ld + r])   (source holds the transpose: K rows of length R)
template <int NS>
__global__ void split_transposed_kernel(const float* __restrict__ src, int64_t ld, int R, int K, int Kp,
                                        __nv_bfloat16* __restrict__ planes) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int dy = ty; dy < 32; dy += 8) {
    const int k = k0 + dy, r = r0 + tx;
    tile[dy][tx] = (k < K && r < R) ? src[(int64_t)k * ld + r] : 0.f;
  }
  __syncthreads();
  for (int dy = ty; dy < 32; dy += 8) {
    const int r = r0 + dy, k = k0 + tx;
    if (r < R && k < Kp) {
      __nv_bfloat16 parts[NS];
      split_bf16<NS>(tile[tx][dy], parts);
#pragma unroll
      for (int pl = 0; pl < NS; ++pl) planes[((int64_t)pl * R + r) * Kp + k] = parts[pl];
    }
  }
}

// ---------------------------------------------------------------------------
// fp16 planes: x*s = hi + lo with s a per-row power of two that puts the row maximum in [2^14, 2^15)
// (fp16 keeps 11 bits, two planes 22; the row scale keeps every row inside fp16's range).
// ---------------------------------------------------------------------------
// scale[r], inv[r] from max_k |src[r*ld + k]|  (one warp per row)
__global__ void rowmax_scale_kernel(const float* __restrict__ src, int64_t ld, int R, int K, float* __restrict__ scale,
                                    float* __restrict__ inv) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* row = src + (int64_t)r * ld;
  float m = 0.f;
  for (int k = threadIdx.x & 31; k < K; k += 32) m = fmaxf(m, fabsf(row[k]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) scale_from_max(m, scale[r], inv[r]);
}

// inv[r] only (long rows: the split kernel recovers the scale as 1 / inv)
__global__ void rowmax_inv_kernel(const float* __restrict__ src, int64_t ld, int R, int K, float* __restrict__ inv) {
  const int r = blockIdx.x;
  if (r >= R) return;
  __shared__ float red[8];
  const float* row = src + (int64_t)r * ld;
  float m = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) m = fmaxf(m, fabsf(row[k]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    float s, i2;
    scale_from_max(m, s, i2);
    inv[r] = i2;
  }
}

// scale[c], inv[c] from max_r |src[r*ld + c]|  (32 columns per block, coalesced rows)
__global__ void colmax_scale_kernel(const float* __restrict__ src, int64_t ld, int R, int Ccols, float* __restrict__ scale,
                                    float* __restrict__ inv) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float m = 0.f;
  if (c < Ccols)
    for (int r = threadIdx.y; r < R; r += 8) m = fmaxf(m, fabsf(src[(int64_t)r * ld + c]));
  red[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0 && c < Ccols) {
#pragma unroll
    for (int y = 1; y < 8; ++y) m = fmaxf(m, red[y][threadIdx.x]);
    scale_from_max(m, scale[c], inv[c]);
  }
}

// planes[pl][r][k] = part_pl(src[r*ld + k] * rs[r] * cs[k]); rs / cs nullable (exactly one is used)
__global__ void split_rows_f16_kernel(const float* __restrict__ src, int64_t ld, int R, int K, int Kp,
                                      const float* __restrict__ rs, const float* __restrict__ cs,
                                      uint16_t* __restrict__ planes, int rs_inverted = 0) {
  const int r = blockIdx.x;                                   // rows on grid.x (may exceed 65535)
  const int k0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (k0 >= Kp) return;
  const float* srow = src + (int64_t)r * ld + k0;
  const float sr = rs ? (rs_inverted ? 1.f / rs[r] : rs[r]) : 1.f;   // powers of two: the reciprocal is exact
  uint16_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x = (k0 + j < K) ? srow[j] * sr : 0.f;
    if (cs && k0 + j < K) x *= cs[k0 + j];
    split_f16(x, hi[j], lo[j]);
  }
  uint16_t* d = planes + (int64_t)r * Kp + k0;
  uint2 w;
  w.x = (uint32_t)hi[0] | ((uint32_t)hi[1] << 16); w.y = (uint32_t)hi[2] | ((uint32_t)hi[3] << 16);
  *reinterpret_cast<uint2*>(d) = w;
  w.x = (uint32_t)lo[0] | ((uint32_t)lo[1] << 16); w.y = (uint32_t)lo[2] | ((uint32_t)lo[3] << 16);
  *reinterpret_cast<uint2*>(d + (int64_t)R * Kp) = w;
}

// planes[pl][r][k] = part_pl(src[k*ld + r] * rs[r])   (source holds the transpose)
__global__ void split_transposed_f16_kernel(const float* __restrict__ src, int64_t ld, int R, int K, int Kp,
                                            const float* __restrict__ rs, uint16_t* __restrict__ planes) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int dy = ty; dy < 32; dy += 8) {
    const int k = k0 + dy, r = r0 + tx;
    tile[dy][tx] = (k < K && r < R) ? src[(int64_t)k * ld + r] : 0.f;
  }
  __syncthreads();
  for (int dy = ty; dy < 32; dy += 8) {
    const int r = r0 + dy, k = k0 + tx;
    if (r < R && k < Kp) {
      uint16_t hi, lo;
      split_f16(tile[tx][dy] * rs[r], hi, lo);
      planes[(int64_t)r * Kp + k] = hi;
      planes[((int64_t)R + r) * Kp + k] = lo;
    }
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// planes [NS][R][Kp] bf16 -> 3-D map (k, r, plane), box 64 x box_rows x 1, 128-byte swizzle, zero OOB fill
static int make_plane_map(CUtensorMap* map, const void* planes, int NS, int R, int K, int Kp, int box_rows,
                          int box_inner = TC_BK) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -3; }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)R, (cuuint64_t)NS};
  cuuint64_t strides[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)R * Kp * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(planes), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return -3; }
  return 0;
}

// the same for other translation units (attention.cu)
int tc_make_plane_map(CUtensorMap* map, const void* planes, int NS, int R, int K, int Kp, int box_rows, int box_inner) {
  return make_plane_map(map, planes, NS, R, K, Kp, box_rows, box_inner);
}

// generic bf16 3-D map used by the SVD tensor-core kernels (declared in tc_common.cuh)
int tc_make_map_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                   uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -3; }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return -3; }
  return 0;
}

int tc_make_map_4d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                   uint64_t stride1_bytes, uint64_t stride2_bytes, uint64_t stride3_bytes, uint32_t box0, uint32_t box1) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return -3; }
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {stride1_bytes, stride2_bytes, stride3_bytes};
  cuuint32_t box[4] = {box0, box1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(4d) failed with CUresult %d", (int)r); return -3; }
  return 0;
}

static int kp_of(int64_t K) { return (int)plane_pitch(K); }
static size_t planes_bytes(int NS, int64_t R, int64_t K) { return round_up((int64_t)NS * R * kp_of(K) * 2, 1024); }
static int ns_of(int prec) { return prec == GRASP_PREC_BF16X6 ? 3 : 2; }
static bool is_f16(int prec) { return prec == GRASP_PREC_F16X3; }

template <int NS>
static int split_operand(const float* src, int64_t ld, int transposed, int R, int K, __nv_bfloat16* planes, void* stream) {
  const int Kp = kp_of(K);
  if (!transposed) {
    dim3 grid((unsigned)R, (unsigned)ceil_div(Kp, 4 * 256));
    GRASP_LAUNCH((split_rows_kernel<NS>), grid, dim3(256), 0, stream, src, ld, R, K, Kp, planes);
  } else {
    dim3 grid((unsigned)ceil_div(Kp, 32), (unsigned)ceil_div(R, 32));
    GRASP_LAUNCH((split_transposed_kernel<NS>), grid, dim3(32, 8), 0, stream, src, ld, R, K, Kp, planes);
  }
  GRASP_CHECK_LAST("split kernel");
  return 0;
}

// fp16 planes of an operand whose K-major form is [R][K]:
//   layout 0: src is [R][K] (scales per row)          -> planes [R][Kp]
//   layout 1: src is [K][R] (transposed split)        -> planes [R][Kp]
//   layout 2: src is [K][R], kept as it is (MN-major) -> planes [K][Rp], scales per column
static int split_operand_f16(const float* src, int64_t ld, int layout, int R, int K, __nv_bfloat16* planes,
                             float* scale, float* inv, void* stream) {
  uint16_t* pl = reinterpret_cast<uint16_t*>(planes);
  if (layout == 0) {
    const int Kp = kp_of(K);
    GRASP_LAUNCH(rowmax_scale_kernel, dim3((unsigned)ceil_div(R, 8)), dim3(256), 0, stream, src, ld, R, K, scale, inv);
    GRASP_LAUNCH(split_rows_f16_kernel, dim3((unsigned)R, (unsigned)ceil_div(Kp, 4 * 256)), dim3(256), 0, stream, src,
                 ld, R, K, Kp, (const float*)scale, (const float*)nullptr, pl);
  } else if (layout == 1) {
    const int Kp = kp_of(K);
    GRASP_LAUNCH(colmax_scale_kernel, dim3((unsigned)ceil_div(R, 32)), dim3(32, 8), 0, stream, src, ld, K, R, scale, inv);
    GRASP_LAUNCH(split_transposed_f16_kernel, dim3((unsigned)ceil_div(Kp, 32), (unsigned)ceil_div(R, 32)), dim3(32, 8),
                 0, stream, src, ld, R, K, Kp, (const float*)scale, pl);
  } else {
    const int Rp = kp_of(R);
    GRASP_LAUNCH(colmax_scale_kernel, dim3((unsigned)ceil_div(R, 32)), dim3(32, 8), 0, stream, src, ld, K, R, scale, inv);
    GRASP_LAUNCH(split_rows_f16_kernel, dim3((unsigned)K, (unsigned)ceil_div(Rp, 4 * 256)), dim3(256), 0, stream, src,
                 ld, K, R, Rp, (const float*)nullptr, (const float*)scale, pl);
  }
  GRASP_CHECK_LAST("fp16 split kernels");
  return 0;
}

static int gemm_kgroup(int ns, int K);
static int raster_group_m(int tiles_m, int tiles_n, int K, int ns);
static bool staged_epilogue();

template <int NS, int BN, int EPI, int BMN = 0, int F16 = 0>
static int launch_core(const __nv_bfloat16* Ap, const __nv_bfloat16* Bp, TcParams prm, void* stream) {
  using Cfg = TcCfg<NS, BN>;
  CUtensorMap mapA, mapB;
  int rc = make_plane_map(&mapA, Ap, NS, prm.M, prm.K, kp_of(prm.K), TC_BM);
  if (rc) return rc;
  if (BMN)   // planes [NS][K][Np]: inner dimension N, rows K
    rc = make_plane_map(&mapB, Bp, NS, prm.K, prm.N, kp_of(prm.N), TC_BK, 64);
  else
    rc = make_plane_map(&mapB, Bp, NS, prm.N, prm.K, kp_of(prm.K), BN);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(tc_gemm_kernel<NS, BN, EPI, BMN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES), "tc_gemm attr");
    if (rc) return rc;
    attr_set = true;
  }
  prm.tiles_m = (int)ceil_div(prm.M, TC_BM);
  prm.tiles_n = (int)ceil_div(prm.N, BN);
  prm.kgroup = gemm_kgroup(NS, prm.K);
  prm.group_m = raster_group_m(prm.tiles_m, prm.tiles_n, prm.K, NS);
  prm.staged = (EPI == EPI_STORE && !prm.c_bf16 && (prm.N & 3) == 0 && (prm.ldc & 3) == 0 &&
                (reinterpret_cast<uintptr_t>(prm.C) & 15) == 0 && (!F16 || (reinterpret_cast<uintptr_t>(prm.inv_sb) & 15) == 0) &&
                staged_epilogue()) ? 1 : 0;
  const int total = prm.tiles_m * prm.tiles_n;
  const int grid = total < sm_count() ? total : sm_count();
  GRASP_LAUNCH((tc_gemm_kernel<NS, BN, EPI, BMN, F16>), dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, mapA, mapB, prm);
  GRASP_CHECK_LAST("tc_gemm_kernel");
  return 0;
}

// Tile width of the two-plane arithmetics.  The kernel is persistent with static tile assignment, so
// the cost is (waves of tiles over the SMs) x (time of one tile); a 256-wide tile does twice the work
// of a 128-wide one in ~1.55x the time (fewer operand bytes per MMA).  GRASP_GEMM_BN forces a width.
static bool wide_tiles(int64_t M, int64_t N, int64_t K = 1 << 20) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("GRASP_GEMM_BN"); forced = e ? atoi(e) : 0; }
  if (forced == 128) return false;
  if (forced == 256) return N > 128;
  if (N <= 128) return false;
  // short K (the rank-k factors of compressed layers): with the staged epilogue and two K blocks per drain the
  // 256-wide tile is 8-17 % faster there as well (N >= 4096, K = 204 / 298: profiles/r02_gemm_shapes_tile_width.txt;
  // round 1's per-lane row stores made it the slower one), so the wave model below decides for every K.
  // GRASP_GEMM_SHORTK_NARROW=1 restores the 128-wide choice.
  static int narrow_short_k = -1;
  if (narrow_short_k < 0) { const char* e = getenv("GRASP_GEMM_SHORTK_NARROW"); narrow_short_k = e ? atoi(e) : 0; }
  if (narrow_short_k && K <= 512 && N > 256) return false;
  const int64_t sms = sm_count();
  const int64_t tm = ceil_div(M, TC_BM);
  const double c128 = (double)ceil_div(tm * ceil_div(N, 128), sms);
  const double c256 = (double)ceil_div(tm * ceil_div(N, 256), sms) * 1.55;
  return c256 < c128;
}

// K blocks per accumulator drain (see tc_gemm_kernel).  The tensor core adds into its accumulator with
// truncation, so longer chains drift towards zero.  Measured on B200 (tools/gemm_kgroup_check.py, K = 4096,
// operands with a non-zero mean): 1 block per drain 5.0e-7 max / -1.0e-7 mean signed error at 0.670 ms
// (8176 x 4096 x 4096), 2 blocks 6.1e-7 / -3.2e-7 at 0.652 ms, 4 blocks 9.8e-7 / -7.5e-7 at 0.638 ms,
// 8 blocks 1.9e-6 / -1.6e-6 at 0.634 ms (torch fp32 matmul: 3.5e-6, unbiased).  The drain is therefore NOT what
// limits the kernel (3-5 %): at 410-430 TFLOP/s fp32-equivalent = 1.25-1.3 PFLOP/s of fp16 MMA it runs at the
// power-capped sustained rate of the part.  Default 1 (accuracy first); GRASP_GEMM_KGROUP overrides.
// Short K (the rank-k factors of compressed layers, K <= 512): a tile lives for <= 8 K blocks and its time is the
// TMEM drain (128 x BN x 4 bytes at 64 B/clk per K block), ~2x the HBM-write bound of the output; two K blocks per
// drain bring the drain under the write bound at 24 accumulations per chain (error class of the row above).
static int gemm_kgroup(int ns, int K) {
  static int v = -1, vs = -1;
  if (v < 0) { const char* e = getenv("GRASP_GEMM_KGROUP"); v = e ? atoi(e) : 0; }
  if (vs < 0) { const char* e = getenv("GRASP_GEMM_KGROUP_SHORT"); vs = e ? atoi(e) : 2; }
  (void)ns;
  if (v > 0) return v;
  return (K <= 512 && vs > 0) ? vs : 1;
}

// M blocks per super-row of the tile order: as many as keep their A planes (128 x K x NS x 2 bytes each) within
// 32 MB of the 126 MB L2, at least 4.  Measured DRAM reads of x[8176,4096] W[11008,4096]^T (ncu, 0.31 GB of operands,
// profiles/r02_gemm_group_m_traffic.txt): 12 blocks 1.63 GB, 16 blocks 1.44 GB, 24 blocks 1.47 GB, 32 blocks 1.77 GB,
// all 64 (the round-1 M-fastest order) 2.80 GB.  GRASP_GEMM_GROUP_M overrides.
static int raster_group_m(int tiles_m, int tiles_n, int K, int ns) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("GRASP_GEMM_GROUP_M"); forced = e ? atoi(e) : 0; }
  if (forced > 0) return forced < tiles_m ? forced : tiles_m;
  (void)tiles_n;
  const double per_block = 128.0 * (double)kp_of(K) * ns * 2.0;
  int g = (int)(32.0 * 1024 * 1024 / per_block);
  if (g < 4) g = 4;
  if (g > tiles_m) g = tiles_m;
  return g;
}

static bool staged_epilogue() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GRASP_GEMM_STAGED"); v = e ? atoi(e) : 1; }
  return v != 0;
}

static bool use_bmn() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GRASP_GEMM_BMN"); v = e ? atoi(e) : 1; }
  return v != 0;
}

template <int EPI, int BMN, int F16>
static int launch_core2(const __nv_bfloat16* Ap, const __nv_bfloat16* Bp, TcParams prm, void* stream) {
  CUtensorMap mapA, mapB;
  int rc = make_plane_map(&mapA, Ap, 2, prm.M, prm.K, kp_of(prm.K), 128);
  if (rc) return rc;
  if (BMN) rc = make_plane_map(&mapB, Bp, 2, prm.K, prm.N, kp_of(prm.N), TC_BK, 64);
  else rc = make_plane_map(&mapB, Bp, 2, prm.N, prm.K, kp_of(prm.K), 128);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(tc_gemm2_kernel<EPI, BMN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Tc2Cfg::SMEM_BYTES), "tc_gemm2 attr");
    if (rc) return rc;
    attr_set = true;
  }
  prm.tiles_m = (int)ceil_div(prm.M, 256);
  prm.tiles_n = (int)ceil_div(prm.N, 256);
  const int pairs_total = prm.tiles_m * prm.tiles_n;
  const int max_pairs = sm_count() / 2;
  const int pairs = pairs_total < max_pairs ? pairs_total : max_pairs;
  GRASP_LAUNCH((tc_gemm2_kernel<EPI, BMN, F16>), dim3(2 * pairs), dim3(Tc2Cfg::THREADS), Tc2Cfg::SMEM_BYTES, stream,
               mapA, mapB, prm);
  GRASP_CHECK_LAST("tc_gemm2_kernel");
  return 0;
}

static bool use_2cta(int64_t M, int64_t N) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GRASP_GEMM_2CTA"); v = e ? atoi(e) : 0; }
  return v != 0 && M > 128 && N > 128;
}

size_t tc_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int prec) {
  const int NS = ns_of(prec);
  const size_t b = planes_bytes(NS, N, K) > planes_bytes(NS, K, N) ? planes_bytes(NS, N, K) : planes_bytes(NS, K, N);
  return planes_bytes(NS, M, K) + b + 2048 + (size_t)round_up(2 * (round_up(M, 4) + round_up(N, 4)) * 4, 1024);
}

size_t tc_sigma_workspace_bytes(int64_t out, int64_t in, int64_t r, int prec) {
  return (size_t)round_up(ceil_div(out, TC_BM) * r * 4, 1024) + tc_gemm_workspace_bytes(out, r, in, prec);
}

static bool dims_ok(int64_t M, int64_t N, int64_t K) {
  return M > 0 && N > 0 && K > 0 && M < (1 << 30) && N < (1 << 30) && K < (1 << 30);
}

int tc_gemm_f32(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                const float* B, int64_t ldb, float beta, void* C, int64_t ldc, int c_bf16, int prec, void* ws,
                size_t ws_bytes, void* stream) {
  if (!dims_ok(M, N, K)) return bad_arg("tc_gemm: M/N/K");
  if (!ws || ws_bytes < tc_gemm_workspace_bytes(M, N, K, prec)) return bad_arg("tc_gemm: workspace too small");
  const int NS = ns_of(prec);
  unsigned char* w = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* Ap = reinterpret_cast<__nv_bfloat16*>(w);
  __nv_bfloat16* Bp = reinterpret_cast<__nv_bfloat16*>(w + planes_bytes(NS, M, K));
  int rc;
  // A: op(A) is M x K; stored [M][K] (ta=0, already K-major) or [K][M] (ta=1)
  // B: op(B) is K x N; K-major form is [N][K]: stored [N][K] when tb=1, [K][N] when tb=0 (needs the transpose)
  // tb = 0: B is stored [K][N]; either transpose it in the split pre-pass (K-major operand) or keep it
  // as it is and feed it as an MN-major operand (planes [K][Np])
  const bool bmn = !tb && use_bmn();
  if (is_f16(prec)) {
    const size_t bbytes = planes_bytes(NS, N, K) > planes_bytes(NS, K, N) ? planes_bytes(NS, N, K) : planes_bytes(NS, K, N);
    // scale_a[M] inv_a[M] scale_b[N] inv_b[N], each on a 16-byte boundary (the epilogue reads inv_b as float4)
    float* sc = reinterpret_cast<float*>(w + planes_bytes(NS, M, K) + bbytes);
    float* inv_a = sc + round_up(M, 4);
    float* sc_b = sc + 2 * round_up(M, 4);
    float* inv_b = sc_b + round_up(N, 4);
    rc = split_operand_f16(A, lda, ta ? 1 : 0, (int)M, (int)K, Ap, sc, inv_a, stream); if (rc) return rc;
    rc = split_operand_f16(B, ldb, tb ? 0 : (bmn ? 2 : 1), (int)N, (int)K, Bp, sc_b, inv_b, stream); if (rc) return rc;
    TcParams prm{};
    prm.M = (int)M; prm.N = (int)N; prm.K = (int)K;
    prm.alpha = alpha; prm.beta = beta; prm.C = C; prm.ldc = ldc; prm.c_bf16 = c_bf16;
    prm.inv_sa = inv_a; prm.inv_sb = inv_b;
    if (use_2cta(M, N)) {
      if (bmn) return launch_core2<EPI_STORE, 1, 1>(Ap, Bp, prm, stream);
      return launch_core2<EPI_STORE, 0, 1>(Ap, Bp, prm, stream);
    }
    if (wide_tiles(M, N, K)) {
      if (bmn) return launch_core<2, 256, EPI_STORE, 1, 1>(Ap, Bp, prm, stream);
      return launch_core<2, 256, EPI_STORE, 0, 1>(Ap, Bp, prm, stream);
    }
    if (bmn) return launch_core<2, 128, EPI_STORE, 1, 1>(Ap, Bp, prm, stream);
    return launch_core<2, 128, EPI_STORE, 0, 1>(Ap, Bp, prm, stream);
  }
  if (NS == 2) {
    rc = split_operand<2>(A, lda, ta, (int)M, (int)K, Ap, stream); if (rc) return rc;
    rc = bmn ? split_operand<2>(B, ldb, 0, (int)K, (int)N, Bp, stream)
             : split_operand<2>(B, ldb, !tb, (int)N, (int)K, Bp, stream);
    if (rc) return rc;
  } else {
    rc = split_operand<3>(A, lda, ta, (int)M, (int)K, Ap, stream); if (rc) return rc;
    rc = bmn ? split_operand<3>(B, ldb, 0, (int)K, (int)N, Bp, stream)
             : split_operand<3>(B, ldb, !tb, (int)N, (int)K, Bp, stream);
    if (rc) return rc;
  }
  TcParams prm{};
  prm.M = (int)M; prm.N = (int)N; prm.K = (int)K;
  prm.alpha = alpha; prm.beta = beta; prm.C = C; prm.ldc = ldc; prm.c_bf16 = c_bf16;
  if (bmn) {
    if (NS == 2) return launch_core<2, 128, EPI_STORE, 1>(Ap, Bp, prm, stream);
    return launch_core<3, 128, EPI_STORE, 1>(Ap, Bp, prm, stream);
  }
  if (NS == 2) return launch_core<2, 128, EPI_STORE>(Ap, Bp, prm, stream);
  return launch_core<3, 128, EPI_STORE>(Ap, Bp, prm, stream);
}

// ---------------------------------------------------------------------------
// Prepared operands (F16X3 arithmetic).  The calibration engine multiplies the same weight by many
// micro-batches (forward as [N][K], backward as [K][N]) and the same activation by several weights, so the
// split pre-pass is exposed on its own: planes [2][rows][round8(cols)] fp16 in the STORED orientation plus
// inverse scales.  Scale per stored row (activations: one read of the row, staged in shared memory) or one
// scale for the whole tensor (weights: the same planes then serve as K-major and as MN-major operand).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rowsplit_f16_kernel(const float* __restrict__ src, int64_t ld, int R, int K, int Kp, uint16_t* __restrict__ planes,
                    float* __restrict__ inv) {
  extern __shared__ __align__(16) float rowbuf[];     // Kp floats
  __shared__ float red[9];
  const int r = blockIdx.x;
  const float* row = src + (int64_t)r * ld;
  float m = 0.f;
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec) {
    for (int k = threadIdx.x * 4; k < Kp; k += 256 * 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k + 3 < K) v = ldg_stream(reinterpret_cast<const float4*>(row + k));
      else { if (k < K) v.x = row[k]; if (k + 1 < K) v.y = row[k + 1]; if (k + 2 < K) v.z = row[k + 2]; }
      *reinterpret_cast<float4*>(rowbuf + k) = v;
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
  } else {
    for (int k = threadIdx.x; k < Kp; k += 256) {
      const float v = (k < K) ? row[k] : 0.f;
      rowbuf[k] = v;
      m = fmaxf(m, fabsf(v));
    }
  }
  block_row_split_256(rowbuf, m, Kp, planes + (int64_t)r * Kp, planes + ((int64_t)R + r) * Kp, inv + r, red);
}

// bits of max |src| over the whole tensor (non-negative floats order like their bit patterns)
__global__ void absmax_bits_kernel(const float* __restrict__ src, int64_t ld, int R, int K, uint32_t* __restrict__ out) {
  float m = 0.f;
  for (int r = blockIdx.x; r < R; r += gridDim.x)
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const float v = fabsf(src[(int64_t)r * ld + k]);
      if (v == v) m = fmaxf(m, v);
    }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

__global__ void tensorsplit_f16_kernel(const float* __restrict__ src, int64_t ld, int R, int K, int Kp,
                                       const uint32_t* __restrict__ maxbits, uint16_t* __restrict__ planes,
                                       float* __restrict__ inv, int n_inv) {
  float s, i;
  scale_from_max(__uint_as_float(*maxbits), s, i);
  const int r = blockIdx.x;                                   // rows on grid.x: vocabularies exceed 65535 rows
  const int k0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (r == 0)
    for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < n_inv; j += gridDim.y * blockDim.x) inv[j] = i;
  if (k0 >= Kp) return;
  const float* srow = src + (int64_t)r * ld + k0;
  uint16_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_f16((k0 + j < K) ? srow[j] * s : 0.f, hi[j], lo[j]);
  uint16_t* d = planes + (int64_t)r * Kp + k0;
  uint2 w;
  w.x = (uint32_t)hi[0] | ((uint32_t)hi[1] << 16); w.y = (uint32_t)hi[2] | ((uint32_t)hi[3] << 16);
  *reinterpret_cast<uint2*>(d) = w;
  w.x = (uint32_t)lo[0] | ((uint32_t)lo[1] << 16); w.y = (uint32_t)lo[2] | ((uint32_t)lo[3] << 16);
  *reinterpret_cast<uint2*>(d + (int64_t)R * Kp) = w;
}

size_t tc_planes_bytes(int64_t rows, int64_t cols) { return planes_bytes(2, rows, cols); }

int tc_split_f16(const float* src, int64_t ld, int64_t rows, int64_t cols, int scale_mode, void* planes, float* inv,
                 void* stream) {
  if (rows <= 0 || cols <= 0 || rows >= (1 << 30) || cols >= (1 << 30)) return bad_arg("split: rows/cols");
  if (reinterpret_cast<uintptr_t>(planes) & 1023) return bad_arg("split: planes must be 1024-byte aligned");
  const int R = (int)rows, K = (int)cols, Kp = kp_of(cols);
  uint16_t* pl = static_cast<uint16_t*>(planes);
  if (scale_mode == GRASP_SCALE_ROWS) {
    const size_t smem = (size_t)Kp * 4;
    if (smem > 200 * 1024) {
      // rows that do not fit in shared memory (logits of a 128k vocabulary): maximum and split as two passes
      GRASP_LAUNCH(rowmax_inv_kernel, dim3((unsigned)R), dim3(256), 0, stream, src, ld, R, K, inv);
      GRASP_LAUNCH(split_rows_f16_kernel, dim3((unsigned)R, (unsigned)ceil_div(Kp, 4 * 256)), dim3(256), 0, stream, src,
                   ld, R, K, Kp, (const float*)inv, (const float*)nullptr, pl, 1);
      GRASP_CHECK_LAST("long-row split kernels");
      return 0;
    }
    static size_t attr = 48 * 1024;
    if (smem > attr) {
      int rc = check_cuda(cudaFuncSetAttribute(rowsplit_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                          "rowsplit attr");
      if (rc) return rc;
      attr = 200 * 1024;
    }
    GRASP_LAUNCH(rowsplit_f16_kernel, dim3((unsigned)R), dim3(256), smem, stream, src, ld, R, K, Kp, pl, inv);
    GRASP_CHECK_LAST("rowsplit_f16_kernel");
    return 0;
  }
  if (scale_mode != GRASP_SCALE_TENSOR) return bad_arg("split: scale_mode");
  // inv holds max(rows, cols) entries followed by one scratch word for the maximum
  const int n_inv = (int)(rows > cols ? rows : cols);
  uint32_t* maxbits = reinterpret_cast<uint32_t*>(inv + n_inv);
  int rc = check_cuda(cudaMemsetAsync(maxbits, 0, 4, (cudaStream_t)stream), "split memset");
  if (rc) return rc;
  GRASP_LAUNCH(absmax_bits_kernel, dim3((unsigned)(R < 1184 ? R : 1184)), dim3(256), 0, stream, src, ld, R, K, maxbits);
  GRASP_LAUNCH(tensorsplit_f16_kernel, dim3((unsigned)R, (unsigned)ceil_div(Kp, 4 * 256)), dim3(256), 0, stream, src, ld,
               R, K, Kp, (const uint32_t*)maxbits, pl, inv, n_inv);
  GRASP_CHECK_LAST("tensor split kernels");
  return 0;
}

// C = alpha * A op(B) + beta * C on prepared operands.  A planes [2][M][Kp]; B planes [2][N][Kp] (b_kn = 0)
// or [2][K][Np] (b_kn = 1); inv_a [M], inv_b [N].
int tc_gemm_planes(int64_t M, int64_t N, int64_t K, float alpha, const void* Ap, const float* inv_a, const void* Bp,
                   int b_kn, const float* inv_b, float beta, float* C, int64_t ldc, void* stream) {
  if (!dims_ok(M, N, K)) return bad_arg("gemm_planes: M/N/K");
  if ((reinterpret_cast<uintptr_t>(Ap) & 1023) || (reinterpret_cast<uintptr_t>(Bp) & 1023))
    return bad_arg("gemm_planes: planes must be 1024-byte aligned");
  TcParams prm{};
  prm.M = (int)M; prm.N = (int)N; prm.K = (int)K;
  prm.alpha = alpha; prm.beta = beta; prm.C = C; prm.ldc = ldc; prm.c_bf16 = 0;
  prm.inv_sa = inv_a; prm.inv_sb = inv_b;
  const __nv_bfloat16* A16 = static_cast<const __nv_bfloat16*>(Ap);
  const __nv_bfloat16* B16 = static_cast<const __nv_bfloat16*>(Bp);
  if (use_2cta(M, N)) {
    if (b_kn) return launch_core2<EPI_STORE, 1, 1>(A16, B16, prm, stream);
    return launch_core2<EPI_STORE, 0, 1>(A16, B16, prm, stream);
  }
  if (wide_tiles(M, N, K)) {
    if (b_kn) return launch_core<2, 256, EPI_STORE, 1, 1>(A16, B16, prm, stream);
    return launch_core<2, 256, EPI_STORE, 0, 1>(A16, B16, prm, stream);
  }
  if (b_kn) return launch_core<2, 128, EPI_STORE, 1, 1>(A16, B16, prm, stream);
  return launch_core<2, 128, EPI_STORE, 0, 1>(A16, B16, prm, stream);
}

static float pow2_at_least(int64_t k) {
  float p = 1.f;
  while (p < (float)k) p *= 2.f;
  return p;
}

// out = A op(B) as a prepared operand (row-scaled planes [2][M][pitch(N)] + inv [M]); B must be tensor-scaled
int tc_gemm_planes_out(int64_t M, int64_t N, int64_t K, const void* Ap, const float* inv_a, const void* Bp, int b_kn,
                       const float* inv_b, void* out_planes, float* out_inv, void* stream) {
  if (!dims_ok(M, N, K)) return bad_arg("gemm_planes_out: M/N/K");
  if ((reinterpret_cast<uintptr_t>(Ap) & 1023) || (reinterpret_cast<uintptr_t>(Bp) & 1023) ||
      (reinterpret_cast<uintptr_t>(out_planes) & 1023))
    return bad_arg("gemm_planes_out: planes must be 1024-byte aligned");
  TcParams prm{};
  prm.M = (int)M; prm.N = (int)N; prm.K = (int)K;
  prm.alpha = 1.f; prm.beta = 0.f;
  prm.inv_sa = inv_a; prm.inv_sb = inv_b;
  prm.out_planes = static_cast<uint16_t*>(out_planes);
  prm.out_pitch = kp_of(N);
  prm.out_plane_stride = M * (int64_t)kp_of(N);
  prm.out_inv = out_inv;
  prm.out_scale = 1.52587890625e-05f /*2^-16*/ / pow2_at_least(K);
  const __nv_bfloat16* A16 = static_cast<const __nv_bfloat16*>(Ap);
  const __nv_bfloat16* B16 = static_cast<const __nv_bfloat16*>(Bp);
  if (wide_tiles(M, N, K)) {
    if (b_kn) return launch_core<2, 256, EPI_PLANES, 1, 1>(A16, B16, prm, stream);
    return launch_core<2, 256, EPI_PLANES, 0, 1>(A16, B16, prm, stream);
  }
  if (b_kn) return launch_core<2, 128, EPI_PLANES, 1, 1>(A16, B16, prm, stream);
  return launch_core<2, 128, EPI_PLANES, 0, 1>(A16, B16, prm, stream);
}

int tc_sigma_partials(const float* U, const float* G, const float* Vh, int64_t out, int64_t in, int64_t r, int prec,
                      float* partial, int64_t* n_partials, void* ws, size_t ws_bytes, void* stream) {
  if (!dims_ok(out, r, in)) return bad_arg("tc_sigma: out/in/r");
  if (ws_bytes < tc_sigma_workspace_bytes(out, in, r, prec)) return bad_arg("tc_sigma: workspace too small");
  const int NS = ns_of(prec);
  const int64_t tiles_m = ceil_div(out, TC_BM);
  unsigned char* w = static_cast<unsigned char*>(ws) + round_up(tiles_m * r * 4, 1024);
  w = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(w) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* Gp = reinterpret_cast<__nv_bfloat16*>(w);
  __nv_bfloat16* Vp = reinterpret_cast<__nv_bfloat16*>(w + planes_bytes(NS, out, in));
  int rc;
  if (is_f16(prec)) {
    float* sc = reinterpret_cast<float*>(w + planes_bytes(NS, out, in) +
                                         (planes_bytes(NS, r, in) > planes_bytes(NS, in, r) ? planes_bytes(NS, r, in)
                                                                                            : planes_bytes(NS, in, r)));
    float* inv_a = sc + out;
    float* sc_b = sc + 2 * out;
    float* inv_b = sc_b + r;
    rc = split_operand_f16(G, in, 0, (int)out, (int)in, Gp, sc, inv_a, stream); if (rc) return rc;
    rc = split_operand_f16(Vh, in, 0, (int)r, (int)in, Vp, sc_b, inv_b, stream); if (rc) return rc;
    TcParams prm{};
    prm.M = (int)out; prm.N = (int)r; prm.K = (int)in;
    prm.alpha = 1.f; prm.beta = 0.f;
    prm.Umul = U; prm.ldu = r; prm.partial = partial;
    prm.inv_sa = inv_a; prm.inv_sb = inv_b;
    *n_partials = tiles_m;
    if (wide_tiles(out, r)) return launch_core<2, 256, EPI_SIGMA, 0, 1>(Gp, Vp, prm, stream);
    return launch_core<2, 128, EPI_SIGMA, 0, 1>(Gp, Vp, prm, stream);
  }
  if (NS == 2) {
    rc = split_operand<2>(G, in, 0, (int)out, (int)in, Gp, stream); if (rc) return rc;
    rc = split_operand<2>(Vh, in, 0, (int)r, (int)in, Vp, stream); if (rc) return rc;
  } else {
    rc = split_operand<3>(G, in, 0, (int)out, (int)in, Gp, stream); if (rc) return rc;
    rc = split_operand<3>(Vh, in, 0, (int)r, (int)in, Vp, stream); if (rc) return rc;
  }
  TcParams prm{};
  prm.M = (int)out; prm.N = (int)r; prm.K = (int)in;
  prm.alpha = 1.f; prm.beta = 0.f;
  prm.Umul = U; prm.ldu = r; prm.partial = partial;
  *n_partials = tiles_m;
  if (NS == 2) return launch_core<2, 128, EPI_SIGMA>(Gp, Vp, prm, stream);
  return launch_core<3, 128, EPI_SIGMA>(Gp, Vp, prm, stream);
}

}  // namespace grasp
