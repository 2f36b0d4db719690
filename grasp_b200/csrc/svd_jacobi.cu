// Thin SVD by block one-sided Jacobi on row vectors (Hestenes), batched over
// same-shape matrices.  Replaces torch.linalg.svd(w, full_matrices=False) at
// reference modeling_grasp.py:231.
//
// Y0 = A (out <= in) or A^T (out > in) is r x L with r = min(out,in).  The working
// matrix Z = [Y | QT] (r x (L + r), QT = I at start) is rotated from the left,
// Z <- E^T Z on pairs of b-row blocks, until the rows of Y are mutually orthogonal:
//   Y = Q^T Y0  =>  Y0 = Q diag(sigma) W^T,  sigma_i = |Y_i|, W^T = rows of Y / sigma.
// One round = (gram) G = Yp Yp^T for every block pair, (evd) small symmetric
// Jacobi eigen-solve in shared memory producing E^T, (update) Z_pair <- E^T Z_pair.
// Rotation angles and the eigenvector accumulation are fp64 (fp32 rotations lose
// orthogonality ~1e-4 at n=4096); the Gram/updates are fp32.
#include "common.cuh"
#include <stdlib.h>

namespace grasp {

// from gemm_tc.cu
int tc_gemm_f32(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                const float* B, int64_t ldb, float beta, void* C, int64_t ldc, int c_bf16, int prec, void* ws,
                size_t ws_bytes, void* stream);
size_t tc_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int prec);

constexpr int JB = 32;        // rows per block
constexpr int JS = 2 * JB;    // rows per pair = order of the small eigenproblem
constexpr int J_THREADS = 256;
constexpr int EVD_THREADS = 1024;  // one 2x2 Gram block and two eigenvector items per thread and step
constexpr int J_MAXMAT = 8;   // matrices per launch group
constexpr int J_STATS = 64;   // uint32 slots of per-matrix status

struct SvdMat {
  float* Z;          // [rp][ldz]
  float* Gpart;      // [npairs][nsplit][JS*JS]
  float* ET;         // [npairs][JS*JS]
  __nv_bfloat16* ETp;  // tensor-core path: [3][ntiles*128][128] block-diagonal ET planes (null on the CUDA-core path)
  int* pair_flag;    // [npairs] 1 = ET is not the identity
  uint32_t* stats;   // [J_STATS]: [0]=converged flag, [1]=sweeps used, [2]=last maxoff bits, [8+s]=maxoff bits of sweep s
};

struct SvdGroup {
  SvdMat mat[J_MAXMAT];
  int nmat;
  int rp, Lp, ldz;   // padded rows, padded Y columns, row stride of Z
  int p;             // number of row blocks (even)
  int npairs;        // p/2
  int nsplit;        // K splits of the Gram
};

// round-robin tournament: pair k of round t among p players
__device__ __forceinline__ void rr_pair(int p, int t, int k, int& a, int& b) {
  int x, y;
  if (k == 0) { x = p - 1; y = t; }
  else {
    x = (t + k) % (p - 1);
    y = (t - k + (p - 1)) % (p - 1);
  }
  a = min(x, y);
  b = max(x, y);
}

// NaN -> lowest key, so that comparisons on it form a total order
__device__ __forceinline__ float finite_key(float d) { return (d == d) ? d : -3.0e38f; }

__device__ __forceinline__ const float* pair_row(const float* Z, int ldz, int I, int J, int rr) {
  const int row = (rr < JB) ? (I * JB + rr) : (J * JB + rr - JB);
  return Z + (int64_t)row * ldz;
}

}  // namespace grasp
#include "svd_tc.cuh"
namespace grasp {

// ---------------------------------------------------------------------------
// init: Z = [Y0 | I], zero padding.  trans: Y0[i][j] = A[j][i]
// ---------------------------------------------------------------------------
__global__ void svd_init_kernel(const float* __restrict__ A, int64_t lda, int r, int L, int trans,
                                float* __restrict__ Z, int rp, int Lp, int ldz) {
  __shared__ float tile[32][33];
  const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (j0 < Lp) {
    if (!trans) {
      for (int dy = ty; dy < 32; dy += 8) {
        const int i = i0 + dy, j = j0 + tx;
        if (i < rp && j < Lp) Z[(int64_t)i * ldz + j] = (i < r && j < L) ? A[(int64_t)i * lda + j] : 0.f;
      }
    } else {
      // A is L x r here (rows of A index j); read coalesced along A's rows, write transposed
      for (int dy = ty; dy < 32; dy += 8) {
        const int j = j0 + dy, i = i0 + tx;
        tile[dy][tx] = (i < r && j < L) ? A[(int64_t)j * lda + i] : 0.f;
      }
      __syncthreads();
      for (int dy = ty; dy < 32; dy += 8) {
        const int i = i0 + dy, j = j0 + tx;
        if (i < rp && j < Lp) Z[(int64_t)i * ldz + j] = tile[tx][dy];
      }
    }
  } else {
    const int jq0 = j0 - Lp;  // QT part
    for (int dy = ty; dy < 32; dy += 8) {
      const int i = i0 + dy, j = jq0 + tx;
      if (i < rp && Lp + j < ldz) Z[(int64_t)i * ldz + Lp + j] = (i == j) ? 1.f : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------
// gram: Gpart[pair][split] = Yp[:, kchunk] Yp[:, kchunk]^T   (64 x 64)
// grid (nsplit, npairs, nmat).  ACC = float: fp32 FMA.  ACC = double (clean-up sweeps): the products of fp32
// numbers are exact in fp64, so the Gram is exact for the rows as stored -- an fp32 sum over L ~ 4096..15000 terms
// carries ~1e-6 |y_i||y_j| of noise, which forced a rotation threshold of 1e-6 and left neighbouring singular
// vectors rotated by ~1e-6 sigma / (2 gap) ~ 2e-3 at n = 4096 (4x LAPACK's error in the rebuilt weights).
// ---------------------------------------------------------------------------
template <typename ACC>
__global__ void __launch_bounds__(J_THREADS)
svd_gram_kernel(SvdGroup g, int round) {
  const SvdMat& M = g.mat[blockIdx.z];
  if (M.stats[0]) return;
  constexpr int KC = 32;
  __shared__ float Ys[KC][JS + 1];
  int I, J;
  rr_pair(g.p, round, blockIdx.y, I, J);
  const int chunks = g.Lp / KC;
  const int per = (chunks + g.nsplit - 1) / g.nsplit;
  const int c_beg = blockIdx.x * per, c_end = min(chunks, c_beg + per);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // ACC = double: every 32-term chunk is summed with fp32 FMAs (exact products, 32 roundings) and the chunk sums are
  // added in fp64 -- the running sum an fp32 addition rounds against stays chunk-sized, so the result is as good as
  // a full fp64 accumulation for this purpose (diagonal: ~1e-8 relative) at the speed of the fp32 kernel
  // (a DFMA-bound version took 1.11 ms per round for eight 4096 x 8192 matrices, the fp32 one 0.62 ms).
  ACC acc[4][4] = {};
  for (int c = c_beg; c < c_end; ++c) {
    // 64 rows x 32 k: thread -> (row = e / 32, kk = e % 32): 128-byte coalesced row segments
#pragma unroll
    for (int it = 0; it < (JS * KC) / J_THREADS; ++it) {
      const int e = tid + it * J_THREADS;
      const int rr = e >> 5, kk = e & 31;
      Ys[kk][rr] = pair_row(M.Z, g.ldz, I, J, rr)[c * KC + kk];
    }
    __syncthreads();
    float part[4][4] = {};
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Ys[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ys[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += (ACC)part[i][j];
    __syncthreads();
  }
  float* out = M.Gpart + ((int64_t)blockIdx.y * g.nsplit + blockIdx.x) * (JS * JS);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(ty * 4 + i) * JS + tx * 4 + j] = (float)acc[i][j];
}

// ---------------------------------------------------------------------------
// evd: parallel-order two-sided Jacobi on the 64x64 Gram of one pair.
// G (fp32) and E (fp64) live in shared memory; rotation parameters in fp64.
// grid (npairs, nmat)
// ---------------------------------------------------------------------------
// ET = float: eigenvectors accumulated in fp32 (tensor-core phase: its drift is removed by the clean-up
// stage anyway); ET = double: fp64 accumulation with fp64-renormalised rotations (CUDA-core sweeps).
template <typename ET>
struct EvdSmem {
  float G[JS][JS + 1];
  ET E[JS][JS + 1];
  ET c[JS / 2], s[JS / 2];
  float cf[JS / 2], sf[JS / 2];
  float red[32];
  int rank[JS];
  int inv[JS];
  int rotated;
  int nonident;
};

// next pair of a fixed tournament slot k when the round advances by one (see rr_pair)
__device__ __forceinline__ void rr_advance(int& x, int& y, int k) {
  y = (y + 1 == JS - 1) ? 0 : y + 1;
  if (k != 0) x = (x + 1 == JS - 1) ? 0 : x + 1;
}

template <typename ET>
__global__ void __launch_bounds__(EVD_THREADS)
svd_evd_kernel(SvdGroup g, int round, int sweep, float tol, int inner_cap) {
  const SvdMat& M = g.mat[blockIdx.y];
  if (M.stats[0]) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EvdSmem<ET>& sm = *reinterpret_cast<EvdSmem<ET>*>(smem_raw);
  const int tid = threadIdx.x;
  const int pair = blockIdx.x;
  constexpr int H = JS / 2;

  // G = sum of K-split partials (fixed order => deterministic)
  const float* gp = M.Gpart + (int64_t)pair * g.nsplit * (JS * JS);
  for (int e = tid; e < JS * JS; e += EVD_THREADS) {
    float v = 0.f;
    for (int sp = 0; sp < g.nsplit; ++sp) v += gp[(int64_t)sp * (JS * JS) + e];
    sm.G[e / JS][e % JS] = v;
    sm.E[e / JS][e % JS] = (e / JS == e % JS) ? (ET)1 : (ET)0;
  }
  if (tid == 0) { sm.rotated = 0; sm.nonident = 0; }
  __syncthreads();
  // symmetrise and measure the largest relative off-diagonal
  float maxoff = 0.f;
  for (int e = tid; e < JS * JS; e += EVD_THREADS) {
    const int i = e / JS, j = e % JS;
    if (i < j) {
      const float v = 0.5f * (sm.G[i][j] + sm.G[j][i]);
      sm.G[i][j] = v;
      sm.G[j][i] = v;
      const float d = sm.G[i][i] * sm.G[j][j];
      if (d > 0.f) maxoff = fmaxf(maxoff, fabsf(v) * rsqrtf(d));
    }
  }
  maxoff = warp_max(maxoff);
  if ((tid & 31) == 0) sm.red[tid >> 5] = maxoff;
  __syncthreads();
  if (tid < 32) {
    float v = sm.red[tid];
    v = warp_max(v);
    if (tid == 0) {
      sm.red[0] = v;
      atomicMax(&M.stats[8 + sweep], __float_as_uint(v));
    }
  }
  __syncthreads();
  maxoff = sm.red[0];

  if (maxoff >= tol) {
    // static work assignment: thread -> one 2x2 block (k1,k2) of G and two (row, pair) items of E
    const int k1 = tid / H, k2 = tid % H;
    for (int isw = 0; isw < inner_cap; ++isw) {
      if (tid == 0) sm.rotated = 0;
      __syncthreads();
      // unordered pairs of slots k1 / k2 in round t, advanced incrementally: slot 0 is (JS-1, t),
      // slot k is ((t+k) mod (JS-1), (t-k) mod (JS-1))
      int x1 = (k1 == 0) ? JS - 1 : k1, y1 = (k1 == 0) ? 0 : (JS - 1 - k1);
      int x2 = (k2 == 0) ? JS - 1 : k2, y2 = (k2 == 0) ? 0 : (JS - 1 - k2);
      for (int t = 0; t < JS - 1; ++t) {
        const int p1 = min(x1, y1), q1 = max(x1, y1), p2 = min(x2, y2), q2 = max(x2, y2);
        if (tid < H) {   // k1 == 0, k2 == tid: this thread's (p2, q2) is the pair of slot tid
          const float gpp = sm.G[p2][p2], gqq = sm.G[q2][q2], gpq = sm.G[p2][q2];
          ET c = 1, s = 0;
          float c0 = 1.f, s0 = 0.f;
          if (gpq != 0.f && gpq * gpq > tol * tol * fabsf(gpp * gqq)) {
            // the angle only steers convergence: fp32 is enough for it
            const float d = gqq - gpp, x = 2.f * gpq;
            const float tt = x / (d + copysignf(sqrtf(fmaf(d, d, x * x)), d));
            c0 = rsqrtf(fmaf(tt, tt, 1.f));
            s0 = tt * c0;
            if constexpr (sizeof(ET) == 8) {
              // orthogonality needs c^2 + s^2 == 1 far below fp32 rounding: renormalise in fp64
              // (first order: 1/sqrt(1+e) = 1 - e/2 for e ~ 1e-7), no fp64 divide / sqrt
              const double cd = (double)c0, sd = (double)s0;
              const double corr = 1.0 - 0.5 * (cd * cd + sd * sd - 1.0);
              c = cd * corr;
              s = sd * corr;
            } else {
              // rsqrtf is approximate and its error has a preferred sign; a correction factor 1 + h with
              // h ~ 1e-7 is not representable in fp32, so add c0*h instead: h = (1 - c0^2 - s0^2) / 2 from two
              // fused multiply-adds (the residual is then unbiased fp32 rounding)
              const float h = -0.5f * fmaf(s0, s0, fmaf(c0, c0, -1.0f));
              c0 = fmaf(c0, h, c0);
              s0 = fmaf(s0, h, s0);
              c = c0; s = s0;
            }
            sm.rotated = 1;
            sm.nonident = 1;
          }
          sm.c[tid] = c; sm.s[tid] = s; sm.cf[tid] = c0; sm.sf[tid] = s0;
        }
        __syncthreads();
        // G <- J^T G J on disjoint 2x2 blocks (one block per thread)
        {
          const float s1 = sm.sf[k1], s2 = sm.sf[k2];
          if (s1 != 0.f || s2 != 0.f) {
            const float c1 = sm.cf[k1], c2 = sm.cf[k2];
            const float b00 = sm.G[p1][p2], b01 = sm.G[p1][q2], b10 = sm.G[q1][p2], b11 = sm.G[q1][q2];
            const float t00 = c1 * b00 - s1 * b10, t01 = c1 * b01 - s1 * b11;
            const float t10 = s1 * b00 + c1 * b10, t11 = s1 * b01 + c1 * b11;
            float n00 = c2 * t00 - s2 * t01, n01 = s2 * t00 + c2 * t01;
            float n10 = c2 * t10 - s2 * t11, n11 = s2 * t10 + c2 * t11;
            if (k1 == k2) { n01 = 0.f; n10 = 0.f; }
            sm.G[p1][p2] = n00; sm.G[p1][q2] = n01; sm.G[q1][p2] = n10; sm.G[q1][q2] = n11;
          }
          // E <- E J: rows i = k1 and k1 + 32, pair k2
          if (s2 != 0.f) {
            const ET c = sm.c[k2], s = sm.s[k2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int i = k1 + h * H;
              const ET ep = sm.E[i][p2], eq = sm.E[i][q2];
              sm.E[i][p2] = c * ep - s * eq;
              sm.E[i][q2] = s * ep + c * eq;
            }
          }
        }
        rr_advance(x1, y1, k1);
        rr_advance(x2, y2, k2);
        __syncthreads();
      }
      if (!sm.rotated) break;
      __syncthreads();
    }
  }

  // order the new rows by descending squared norm (diag of the rotated Gram)
  if (tid < JS) {
    // (NaN diagonals -- non-finite input -- sort last: the ranks must stay a permutation, they index shared memory)
    const float di = finite_key(sm.G[tid][tid]);
    int rk = 0;
    for (int j = 0; j < JS; ++j) {
      const float dj = finite_key(sm.G[j][j]);
      rk += (dj > di) || (dj == di && j < tid);
    }
    sm.rank[tid] = rk;
    if (rk != tid) sm.nonident = 1;
  }
  __syncthreads();
  if (tid < JS) sm.inv[sm.rank[tid]] = tid;
  __syncthreads();
  if (tid == 0) M.pair_flag[pair] = sm.nonident;
  if (M.ETp) {
    // block-diagonal ET of the two pairs of a tile as three bf16 planes (always written, identity included)
    const int ntiles = (g.npairs + 1) / 2;
    const int64_t plane = (int64_t)ntiles * 128 * 128;
    __nv_bfloat16* base = M.ETp + ((int64_t)(pair >> 1) * 128 + (pair & 1) * 64) * 128;
    for (int e = tid; e < JS * 128; e += EVD_THREADS) {
      const int rn = e >> 7, col = e & 127;
      const int i = col - (pair & 1) * 64;          // component index inside this pair's 64 columns
      float x = 0.f;
      if (i >= 0 && i < JS) x = (float)sm.E[i][sm.inv[rn]];   // row rn of ET = eigenvector of rank rn
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        const __nv_bfloat16 b = __float2bfloat16_rn(x);
        base[pl * plane + (int64_t)rn * 128 + col] = b;
        x -= __bfloat162float(b);
      }
    }
  }
  if (sm.nonident) {
    float* et = M.ET + (int64_t)pair * (JS * JS);
    for (int e = tid; e < JS * JS; e += EVD_THREADS) {
      const int c = e / JS, i = e % JS;       // eigenvector c, component i
      et[sm.rank[c] * JS + i] = (float)sm.E[i][c];
    }
  }
}

// ---------------------------------------------------------------------------
// evd, register-resident variant used by the tensor-core phase (fp32 eigenvectors, same rotations,
// thresholds and outputs as svd_evd_kernel<float>).
//
// svd_evd_kernel spends ~480 warp instructions per Jacobi step on each of 32 warps (index arithmetic,
// shared-memory round trips, two block barriers for one 2x2 update per thread): ncu shows it issue-bound
// (39 M warp instructions per launch at n = 4096).  Here the matrices stay in registers for the whole sweep:
//   lane k of every warp owns columns A = top[k], B = bot[k]
//   warps 0-3 (G): row slots T[8w..8w+7], B[8w..8w+7] of G           (32 registers)
//   warps 4-7 (E): rows 8w'..8w'+7 and 32+8w'..32+8w'+7 of E         (32 registers; rows of E never move)
// Rows live in fixed slots and pair j is always (T[j], B[j]), so every register index is a compile-time
// constant.  After each step rows and columns move one place round the Brent-Luk ring (T[0] fixed,
// T[1]..T[31] -> B[31]..B[0] -> T[1]): columns by one warp shuffle per value, rows by the choice of the
// destination register, the two slots that cross a warp boundary through shared memory.  One step is
// ~320 instructions on a G warp and two 256-thread barriers; a launch executes ~5x fewer instructions.
// grid (npairs, nmat), 256 threads.
// ---------------------------------------------------------------------------
constexpr int EVW_THREADS = 256;

struct EvdWarpSmem {
  float G[JS][JS + 1];
  float E[JS][JS + 1];
  float c[32], s[32];
  float xT[4][2][32];     // slot T[8w+7] of G warp w after the column move (-> T[8w+8])
  float xB[4][2][32];     // slot B[8w]                                     (-> B[8w-1])
  float red[EVW_THREADS / 32];
  int rank[JS];
  int inv[JS];
  int rotated;
  int nonident;
};

// one value pair (column A, column B of this lane) moves to its place for the next step
__device__ __forceinline__ void ring_cols(float vA, float vB, int lane, float& nA, float& nB) {
  const float x = (lane == 0) ? vB : vA;
  const float u = __shfl_up_sync(0xffffffffu, x, 1);
  const float d = __shfl_down_sync(0xffffffffu, vB, 1);
  nA = (lane == 0) ? vA : u;
  nB = (lane == 31) ? vA : d;
}

__global__ void __launch_bounds__(EVW_THREADS)
svd_evd_warp_kernel(SvdGroup g, int round, int sweep, float tol, int inner_cap) {
  const SvdMat& M = g.mat[blockIdx.y];
  if (M.stats[0]) return;
  __shared__ EvdWarpSmem sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pair = blockIdx.x;
  constexpr int H = JS / 2;
  constexpr int R = 8;                 // row slots of each kind per G warp
  static_assert(JS == 64, "the register layout assumes 64-row pairs");

  // G = sum of the K-split partials in a fixed order; four independent 16-byte streams per thread so that the
  // loads of all partials are in flight together (the kernel is a latency chain: 19 % of its samples sat here)
  const float4* gp4 = reinterpret_cast<const float4*>(M.Gpart + (int64_t)pair * g.nsplit * (JS * JS));
  {
    float4 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sp = 0; sp < g.nsplit; ++sp) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t = gp4[(int64_t)sp * (JS * JS / 4) + tid + i * EVW_THREADS];
        acc[i].x += t.x; acc[i].y += t.y; acc[i].z += t.z; acc[i].w += t.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = (tid + i * EVW_THREADS) * 4, r = e / JS, c = e % JS;
      sm.G[r][c] = acc[i].x; sm.G[r][c + 1] = acc[i].y; sm.G[r][c + 2] = acc[i].z; sm.G[r][c + 3] = acc[i].w;
#pragma unroll
      for (int j = 0; j < 4; ++j) sm.E[r][c + j] = (r == c + j) ? 1.f : 0.f;
    }
  }
  if (tid == 0) { sm.rotated = 0; sm.nonident = 0; }
  __syncthreads();
  float maxoff = 0.f;
  for (int e = tid; e < JS * JS; e += EVW_THREADS) {
    const int i = e / JS, j = e % JS;
    if (i < j) {
      const float v = 0.5f * (sm.G[i][j] + sm.G[j][i]);
      sm.G[i][j] = v;
      sm.G[j][i] = v;
      const float d = sm.G[i][i] * sm.G[j][j];
      if (d > 0.f) maxoff = fmaxf(maxoff, fabsf(v) * rsqrtf(d));
    }
  }
  maxoff = warp_max(maxoff);
  if (lane == 0) sm.red[warp] = maxoff;
  __syncthreads();
  maxoff = 0.f;
#pragma unroll
  for (int i = 0; i < EVW_THREADS / 32; ++i) maxoff = fmaxf(maxoff, sm.red[i]);
  if (tid == 0) atomicMax(&M.stats[8 + sweep], __float_as_uint(maxoff));

  if (maxoff >= tol) {
    const bool gw = warp < 4;            // G warp (else E warp)
    const int w = warp & 3;
    // G warps: [local slot] of column A / column B, t = slots T[8w + i], b = slots B[8w + i]
    // E warps: t = rows 8w + i, b = rows 32 + 8w + i
    float tA[R], tB[R], bA[R], bB[R];
    // starting arrangement of the ring: top = {63, 1, 2, .., 31}, bot = {0, 62, 61, .., 32}, i.e. the first step
    // pairs index k with 63 - k (rows arrive sorted by norm: large with small) and the following steps run the
    // same tournament as svd_evd_kernel; after 63 steps everything is back in these places
    const int colA = lane == 0 ? JS - 1 : lane, colB = lane == 0 ? 0 : JS - 1 - lane;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int j = 8 * w + i;
      const int rowT = j == 0 ? JS - 1 : j, rowB = j == 0 ? 0 : JS - 1 - j;
      if (gw) {
        tA[i] = sm.G[rowT][colA]; tB[i] = sm.G[rowT][colB];
        bA[i] = sm.G[rowB][colA]; bB[i] = sm.G[rowB][colB];
      } else {
        // E rows never move: this warp keeps rows j and 32 + j
        tA[i] = (j == colA) ? 1.f : 0.f; tB[i] = (j == colB) ? 1.f : 0.f;
        bA[i] = (H + j == colA) ? 1.f : 0.f; bB[i] = (H + j == colB) ? 1.f : 0.f;
      }
    }
    for (int isw = 0; isw < inner_cap; ++isw) {
      bool rotated = false;
      for (int t = 0; t < JS - 1; ++t) {
        // ---- A: the G warp that holds pair k's diagonal block (slot k, lane k) computes its rotation
        if (gw && (lane >> 3) == w) {
          float gpp = 0.f, gqq = 0.f, gpq = 0.f;
#pragma unroll
          for (int i = 0; i < R; ++i)
            if (i == (lane & 7)) { gpp = tA[i]; gpq = tB[i]; gqq = bB[i]; }
          float c = 1.f, s = 0.f;
          if (gpq != 0.f && gpq * gpq > tol * tol * fabsf(gpp * gqq)) {
            const float d = gqq - gpp, x = 2.f * gpq;
            const float tt = x / (d + copysignf(sqrtf(fmaf(d, d, x * x)), d));
            c = rsqrtf(fmaf(tt, tt, 1.f));
            s = tt * c;
            const float h = -0.5f * fmaf(s, s, fmaf(c, c, -1.0f));   // unbiased normalisation (see svd_evd_kernel)
            c = fmaf(c, h, c);
            s = fmaf(s, h, s);
            rotated = true;
          }
          sm.c[lane] = c;
          sm.s[lane] = s;
        }
        __syncthreads();
        // ---- B: rotate, move the columns, shift the row slots
        const float c = sm.c[lane], s = sm.s[lane];
        if (gw) {
          if (s != 0.f) {                                   // G <- G J: this lane's two columns
#pragma unroll
            for (int i = 0; i < R; ++i) {
              float a = tA[i], b = tB[i];
              tA[i] = c * a - s * b; tB[i] = s * a + c * b;
              a = bA[i]; b = bB[i];
              bA[i] = c * a - s * b; bB[i] = s * a + c * b;
            }
          }
#pragma unroll
          for (int i = 0; i < R; ++i) {                     // G <- J^T G: row pair 8w + i with the rotation of that lane
            const float cj = __shfl_sync(0xffffffffu, c, 8 * w + i), sj = __shfl_sync(0xffffffffu, s, 8 * w + i);
            float a = tA[i], b = bA[i];
            tA[i] = cj * a - sj * b; bA[i] = sj * a + cj * b;
            a = tB[i]; b = bB[i];
            tB[i] = cj * a - sj * b; bB[i] = sj * a + cj * b;
          }
          if (s != 0.f && (lane >> 3) == w) {
#pragma unroll
            for (int i = 0; i < R; ++i)
              if (i == (lane & 7)) { tB[i] = 0.f; bA[i] = 0.f; }     // the annihilated pivot, exactly
          }
          // column move of every slot, then the slot shift: T[j] <- T[j-1], B[j] <- B[j+1]
          float oA, oB;
          ring_cols(tA[R - 1], tB[R - 1], lane, oA, oB);            // T[8w+7] leaves for T[8w+8] (or B[31])
          sm.xT[w][0][lane] = oA; sm.xT[w][1][lane] = oB;
          float fA, fB;
          ring_cols(bA[0], bB[0], lane, fA, fB);                    // B[8w] leaves for B[8w-1] (or T[1])
          sm.xB[w][0][lane] = fA; sm.xB[w][1][lane] = fB;
#pragma unroll
          for (int i = R - 1; i >= 2; --i) ring_cols(tA[i - 1], tB[i - 1], lane, tA[i], tB[i]);
          {
            float zA, zB;
            ring_cols(tA[0], tB[0], lane, zA, zB);                  // old local T[0]
            if (w == 0) { tA[1] = fA; tB[1] = fB; tA[0] = zA; tB[0] = zB; }   // T[1] <- B[0]; T[0] stays
            else { tA[1] = zA; tB[1] = zB; }                        // T[8w+1] <- T[8w]; T[8w] is imported below
          }
#pragma unroll
          for (int i = 0; i < R - 1; ++i) ring_cols(bA[i + 1], bB[i + 1], lane, bA[i], bB[i]);
          if (w == 3) { bA[R - 1] = oA; bB[R - 1] = oB; }           // B[31] <- T[31]
        } else {
#pragma unroll
          for (int i = 0; i < R; ++i) {                     // E <- E J, then the same column move
            float a = tA[i], b = tB[i];
            if (s != 0.f) { const float na = c * a - s * b; b = s * a + c * b; a = na; }
            ring_cols(a, b, lane, tA[i], tB[i]);
            a = bA[i]; b = bB[i];
            if (s != 0.f) { const float na = c * a - s * b; b = s * a + c * b; a = na; }
            ring_cols(a, b, lane, bA[i], bB[i]);
          }
        }
        __syncthreads();
        // ---- C: the two slots that crossed a warp boundary
        if (gw) {
          if (w > 0) { tA[0] = sm.xT[w - 1][0][lane]; tB[0] = sm.xT[w - 1][1][lane]; }
          if (w < 3) { bA[R - 1] = sm.xB[w + 1][0][lane]; bB[R - 1] = sm.xB[w + 1][1][lane]; }
        }
      }
      // after JS - 1 steps the ring is back where it started: lane k holds columns k and 32 + k again
      const bool any = __any_sync(0xffffffffu, rotated);
      if (gw && lane == 0 && any) { sm.rotated = 1; sm.nonident = 1; }
      __syncthreads();
      const int again = sm.rotated;
      __syncthreads();
      if (tid == 0) sm.rotated = 0;
      if (!again) break;
    }
    if (gw) {
      if ((lane >> 3) == w) {
        float dA = 0.f, dB = 0.f;
#pragma unroll
        for (int i = 0; i < R; ++i)
          if (i == (lane & 7)) { dA = tA[i]; dB = bB[i]; }
        sm.G[colA][colA] = dA;
        sm.G[colB][colB] = dB;
      }
    } else {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int j = 8 * w + i;
        sm.E[j][colA] = tA[i]; sm.E[j][colB] = tB[i];
        sm.E[H + j][colA] = bA[i]; sm.E[H + j][colB] = bB[i];
      }
    }
    __syncthreads();
  }

  // order the new rows by descending squared norm (diag of the rotated Gram)
  if (tid < JS) {
    // (NaN diagonals -- non-finite input -- sort last: the ranks must stay a permutation, they index shared memory)
    const float di = finite_key(sm.G[tid][tid]);
    int rk = 0;
    for (int j = 0; j < JS; ++j) {
      const float dj = finite_key(sm.G[j][j]);
      rk += (dj > di) || (dj == di && j < tid);
    }
    sm.rank[tid] = rk;
    if (rk != tid) sm.nonident = 1;
  }
  __syncthreads();
  if (tid < JS) sm.inv[sm.rank[tid]] = tid;
  __syncthreads();
  if (tid == 0) M.pair_flag[pair] = sm.nonident;
  if (M.ETp) {
    const int ntiles = (g.npairs + 1) / 2;
    const int64_t plane = (int64_t)ntiles * 128 * 128;
    __nv_bfloat16* base = M.ETp + ((int64_t)(pair >> 1) * 128 + (pair & 1) * 64) * 128;
    // two adjacent columns per thread: 4-byte stores, 64 threads cover one 128-column row
    for (int e = tid; e < JS * 64; e += EVW_THREADS) {
      const int rn = e >> 6, col = (e & 63) * 2;
      const int i = col - (pair & 1) * 64;
      float x0 = 0.f, x1 = 0.f;
      if (i >= 0 && i < JS) { const int src = sm.inv[rn]; x0 = sm.E[i][src]; x1 = sm.E[i + 1][src]; }
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        const __nv_bfloat16 b0 = __float2bfloat16_rn(x0), b1 = __float2bfloat16_rn(x1);
        *reinterpret_cast<__nv_bfloat162*>(base + pl * plane + (int64_t)rn * 128 + col) = __halves2bfloat162(b0, b1);
        x0 -= __bfloat162float(b0);
        x1 -= __bfloat162float(b1);
      }
    }
  }
  if (sm.nonident) {
    float* et = M.ET + (int64_t)pair * (JS * JS);
    for (int e = tid; e < JS * JS; e += EVW_THREADS) {
      const int c = e / JS, i = e % JS;
      et[sm.rank[c] * JS + i] = sm.E[i][c];
    }
  }
}

// ---------------------------------------------------------------------------
// update: Z_pair[:, cols] <- ET (64x64) * Z_pair[:, cols], in place.
// grid (ldz/128, npairs, nmat); 256 threads, 4 rows x 8 cols per thread.
// ---------------------------------------------------------------------------
constexpr int UP_TN = 128;

__global__ void __launch_bounds__(J_THREADS)
svd_update_kernel(SvdGroup g, int round) {
  const SvdMat& M = g.mat[blockIdx.z];
  if (M.stats[0]) return;
  const int pair = blockIdx.y;
  if (!M.pair_flag[pair]) return;
  __shared__ __align__(16) float Est[JS][JS];   // Est[k][row] = ET[row][k]
  __shared__ __align__(16) float Zs[JS][UP_TN];
  int I, J;
  rr_pair(g.p, round, pair, I, J);
  const int tid = threadIdx.x;
  const int col0 = blockIdx.x * UP_TN;
  const float* et = M.ET + (int64_t)pair * (JS * JS);
  for (int e = tid; e < JS * JS; e += J_THREADS) Est[e % JS][e / JS] = et[e];
  // 64 rows x 128 cols as float4: 2048 vectors
#pragma unroll
  for (int it = 0; it < (JS * UP_TN / 4) / J_THREADS; ++it) {
    const int e = tid + it * J_THREADS;
    const int rr = e >> 5, v = e & 31;
    const float4 x = *reinterpret_cast<const float4*>(pair_row(M.Z, g.ldz, I, J, rr) + col0 + v * 4);
    *reinterpret_cast<float4*>(&Zs[rr][v * 4]) = x;
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;  // rows ty*4.., cols tx*4.. and 64+tx*4..
  float acc[4][8] = {};
#pragma unroll 8
  for (int k = 0; k < JS; ++k) {
    const float4 b0 = *reinterpret_cast<const float4*>(&Zs[k][tx * 4]);
    const float4 b1 = *reinterpret_cast<const float4*>(&Zs[k][64 + tx * 4]);
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float4 a4 = *reinterpret_cast<const float4*>(&Est[k][ty * 4]);
    const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* dst = const_cast<float*>(pair_row(M.Z, g.ldz, I, J, ty * 4 + i)) + col0;
    *reinterpret_cast<float4*>(dst + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(dst + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

// one thread per matrix: close the sweep
__global__ void svd_sweep_end_kernel(SvdGroup g, int sweep, float tol, int cleanup_idx) {
  const int m = threadIdx.x;
  if (m >= g.nmat) return;
  uint32_t* st = g.mat[m].stats;
  if (st[0]) return;
  const float off = __uint_as_float(st[8 + sweep]);
  st[1] = (cleanup_idx < 0) ? sweep + 1 : st[3] + cleanup_idx + 1;   // sweeps used so far
  st[2] = st[8 + sweep];
  if (off < tol) st[0] = 1;
}

// ---------------------------------------------------------------------------
// finalize: sigma_i = |Y_i| (fp64 accumulate), then emit sorted factors
// ---------------------------------------------------------------------------
__global__ void svd_norms_kernel(const float* __restrict__ Z, int ldz, int rp, int Lp, float* __restrict__ sigma) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rp) return;
  const float* y = Z + (int64_t)row * ldz;
  double acc = 0.0;
  for (int j = (threadIdx.x & 31) * 4; j < Lp; j += 128) {
    const float4 v = *reinterpret_cast<const float4*>(y + j);
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sigma[row] = (float)sqrt(acc);
}

// out[i][j] = src[perm[i]][col0 + j] * (normalize ? 1/sigma[perm[i]] : 1),  i < r, j < ncols
__global__ void svd_emit_rows_kernel(const float* __restrict__ Z, int ldz, int col0, const int64_t* __restrict__ perm,
                                     const float* __restrict__ sigma, int normalize, int r, int ncols,
                                     float* __restrict__ out, int64_t ldo, float* __restrict__ S_out) {
  const int i = blockIdx.x;
  const int src = (int)perm[i];
  const float sg = sigma[src];
  const float scale = normalize ? (sg > 0.f ? 1.f / sg : 0.f) : 1.f;
  if (S_out && threadIdx.x == 0) S_out[i] = sg;
  const float* z = Z + (int64_t)src * ldz + col0;
  for (int j = threadIdx.x; j < ncols; j += blockDim.x) out[(int64_t)i * ldo + j] = z[j] * scale;
}

// out[a][i] = src[perm[i]][col0 + a] * scale_i,  a < nrows_out, i < r   (tiled transpose)
__global__ void svd_emit_cols_kernel(const float* __restrict__ Z, int ldz, int col0, const int64_t* __restrict__ perm,
                                     const float* __restrict__ sigma, int normalize, int r, int nrows_out,
                                     float* __restrict__ out, int64_t ldo) {
  __shared__ float tile[32][33];
  const int i0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int dy = ty; dy < 32; dy += 8) {
    const int i = i0 + dy, a = a0 + tx;
    float v = 0.f;
    if (i < r && a < nrows_out) {
      const int src = (int)perm[i];
      const float sg = sigma[src];
      const float scale = normalize ? (sg > 0.f ? 1.f / sg : 0.f) : 1.f;
      v = Z[(int64_t)src * ldz + col0 + a] * scale;
    }
    tile[dy][tx] = v;
  }
  __syncthreads();
  for (int dy = ty; dy < 32; dy += 8) {
    const int a = a0 + dy, i = i0 + tx;
    if (a < nrows_out && i < r) out[(int64_t)a * ldo + i] = tile[tx][dy];
  }
}

// convergence check on the full Gram T = Y Y^T (one tensor-core GEMM) instead of a verification sweep:
// stats[60] = bits of max_{i != j} |t_ij| / sqrt(t_ii t_jj); grid = rows
__global__ void __launch_bounds__(256)
svd_offdiag_max_kernel(const float* __restrict__ T, int ld, int r, uint32_t* __restrict__ stats) {
  if (stats[0]) return;
  const int i = blockIdx.x;
  const float dii = T[(int64_t)i * ld + i];
  float m = 0.f;
  for (int j = threadIdx.x; j < r; j += blockDim.x) {
    if (j == i) continue;
    const float d = dii * T[(int64_t)j * ld + j];
    if (d > 0.f) m = fmaxf(m, fabsf(T[(int64_t)i * ld + j]) * rsqrtf(d));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(&stats[60], __float_as_uint(m));
}

__global__ void svd_check_end_kernel(SvdGroup g, float tol) {
  const int m = threadIdx.x;
  if (m >= g.nmat) return;
  uint32_t* st = g.mat[m].stats;
  if (st[0]) return;
  if (__uint_as_float(st[60]) < tol) { st[0] = 1; st[2] = st[60]; }
  st[60] = 0;
}

__global__ void svd_reopen_kernel(SvdGroup g) {
  if (threadIdx.x < g.nmat) {
    uint32_t* st = g.mat[threadIdx.x].stats;
    st[3] = st[1];   // sweeps of the tensor-core phase
    st[0] = 0;
  }
}

__global__ void svd_info_kernel(const uint32_t* __restrict__ stats, const uint32_t* __restrict__ pre_fail,
                                int32_t* __restrict__ info) {
  info[0] = (int32_t)stats[1];
  info[1] = (int32_t)(stats[0] && !(pre_fail && *pre_fail));   // converged, and the preconditioning (if any) was sound
  info[2] = (int32_t)stats[2];
  info[3] = (int32_t)stats[3];   // sweeps of the tensor-core phase (0 on the CUDA-core path)
}

// ---------------------------------------------------------------------------
// CholeskyQR2 preconditioning of wide working matrices (L >= 1.5 r).
// The Jacobi phase streams whole rows of Z = [Y | QT] through HBM once per round, so its time grows with the
// row length L + r.  For Y0 (r x L) = Lm * Q with Q (r x L) of orthonormal rows, the SVD of the SQUARE
// Lm = U S W^T gives Y0 = U S (W^T Q): the rounds run on rows of length 2 r instead of L + r and the long
// dimension is touched only by a handful of tensor-core GEMMs (tc_gemm_f32, fp32-class arithmetic):
//     G = Y0 Y0^T,  M1 = inv(chol(G)),  Q1 = M1 Y0,  G2 = Q1 Q1^T,  M2 = inv(chol(G2)),  Q = M2 Q1,  Lm = Y0 Q^T.
// The inverse Cholesky factor is built recursively: for G = [[G11, .], [G21, G22]],
//     M11 = invchol(G11), T = G21 M11^T, M22 = invchol(G22 - T T^T), M21 = -M22 T M11,
// with 128 x 128 blocks solved by one CTA in shared memory.  A non-positive pivot or a Q that is not orthonormal
// to 1e-4 (cond(Y0)^2 beyond fp32: does not happen for random-init weights, can for trained ones) sets a flag that
// ends up in info[1] = 0, and the caller repeats the matrix with GRASP_SVD_NO_PRECOND.
// ---------------------------------------------------------------------------
constexpr int IC_BASE = 128;
constexpr int IC_SMEM = 2 * IC_BASE * (IC_BASE + 1) * 4;

// M = inv(chol(G)) for one n x n block (n <= 128); only the lower triangle of G is read; M upper triangle = 0
__global__ void __launch_bounds__(IC_BASE)
svd_invchol_base_kernel(const float* __restrict__ G, int64_t ldg, int n, float* __restrict__ M, int64_t ldm,
                        uint32_t* __restrict__ fail) {
  extern __shared__ float ic_smem[];
  float (*Ls)[IC_BASE + 1] = reinterpret_cast<float (*)[IC_BASE + 1]>(ic_smem);
  float (*Xs)[IC_BASE + 1] = reinterpret_cast<float (*)[IC_BASE + 1]>(ic_smem + IC_BASE * (IC_BASE + 1));
  __shared__ float piv;
  const int t = threadIdx.x;
  for (int e = t; e < n * n; e += IC_BASE) {
    const int i = e / n, k = e - i * n;
    Ls[i][k] = (k <= i) ? G[(int64_t)i * ldg + k] : 0.f;
  }
  __syncthreads();
  // left-looking Cholesky, thread t owns row t
  for (int j = 0; j < n; ++j) {
    float s = 0.f;
    if (t >= j && t < n) {
      for (int k = 0; k < j; ++k) s = fmaf(Ls[t][k], Ls[j][k], s);
      s = Ls[t][j] - s;
    }
    if (t == j) {
      const float d0 = Ls[j][j];
      const float floor_ = (d0 > 0.f) ? 1e-7f * d0 : 1.f;
      if (!(s > floor_)) { atomicOr(fail, 1u); s = floor_; }
      piv = sqrtf(s);
    }
    __syncthreads();
    if (t >= j && t < n) Ls[t][j] = (t == j) ? piv : s / piv;
    __syncthreads();
  }
  // X = L^-1 by forward substitution, thread t owns column t
  if (t < n) {
    Xs[t][t] = 1.f / Ls[t][t];
    for (int i = t + 1; i < n; ++i) {
      float s = 0.f;
      for (int k = t; k < i; ++k) s = fmaf(Ls[i][k], Xs[k][t], s);
      Xs[i][t] = -s / Ls[i][i];
    }
  }
  __syncthreads();
  for (int e = t; e < n * n; e += IC_BASE) {
    const int i = e / n, k = e - i * n;
    M[(int64_t)i * ldm + k] = (k <= i) ? Xs[i][k] : 0.f;
  }
}

// flag |= (max_ij |T_ij - delta_ij| > tol);  grid = rows of T
__global__ void __launch_bounds__(256)
svd_orth_defect_kernel(const float* __restrict__ T, int64_t ld, int r, float tol, uint32_t* __restrict__ fail) {
  const int i = blockIdx.x;
  float m = 0.f;
  for (int j = threadIdx.x; j < r; j += blockDim.x) {
    const float v = T[(int64_t)i * ld + j] - (i == j ? 1.f : 0.f);
    m = fmaxf(m, (v == v) ? fabsf(v) : 1e30f);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > tol) atomicOr(fail, 2u);
}

// G[i][i] += rel * mean_i G[i][i]   (one CTA)
__global__ void __launch_bounds__(1024)
svd_shift_diag_kernel(float* __restrict__ G, int64_t ld, int r, float rel) {
  __shared__ float red[32];
  __shared__ float shift;
  float acc = 0.f;
  for (int i = threadIdx.x; i < r; i += blockDim.x) acc += G[(int64_t)i * ld + i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) shift = rel * v / (float)r;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < r; i += blockDim.x) G[(int64_t)i * ld + i] += shift;
}

struct IcCtx {
  void* stream;
  void* gws;
  size_t gws_bytes;
  uint32_t* fail;
};

// M (n x n, zero-initialised by the caller) = inv(chol(G)); G is overwritten; scratch holds (2/3) n^2 floats
static int invchol_rec(float* G, int64_t ldg, int n, float* M, int64_t ldm, float* scratch, const IcCtx& c) {
  if (n <= IC_BASE) {
    GRASP_LAUNCH(svd_invchol_base_kernel, dim3(1), dim3(IC_BASE), IC_SMEM, (cudaStream_t)c.stream, (const float*)G, ldg, n, M,
                 ldm, c.fail);
    return check_cuda(cudaGetLastError(), "svd_invchol_base_kernel");
  }
  const int h = ((n / 2 + 63) / 64) * 64, n2 = n - h;
  int rc = invchol_rec(G, ldg, h, M, ldm, scratch, c);
  if (rc) return rc;
  float* T = scratch;
  float* tmp = scratch + (size_t)n2 * h;
  float* next = scratch + (size_t)2 * n2 * h;
  float* G21 = G + (int64_t)h * ldg;
  float* G22 = G21 + h;
  float* M21 = M + (int64_t)h * ldm;
  float* M22 = M21 + h;
  const int P = GRASP_PREC_F16X3;
  rc = tc_gemm_f32(0, 1, n2, h, h, 1.f, G21, ldg, M, ldm, 0.f, T, h, 0, P, c.gws, c.gws_bytes, c.stream);       // T = G21 M11^T
  if (rc) return rc;
  rc = tc_gemm_f32(0, 1, n2, n2, h, -1.f, T, h, T, h, 1.f, G22, ldg, 0, P, c.gws, c.gws_bytes, c.stream);      // G22 -= T T^T
  if (rc) return rc;
  rc = invchol_rec(G22, ldg, n2, M22, ldm, next, c);
  if (rc) return rc;
  rc = tc_gemm_f32(0, 0, n2, h, n2, 1.f, M22, ldm, T, h, 0.f, tmp, h, 0, P, c.gws, c.gws_bytes, c.stream);      // tmp = M22 T
  if (rc) return rc;
  return tc_gemm_f32(0, 0, n2, h, h, -1.f, tmp, h, M, ldm, 0.f, M21, ldm, 0, P, c.gws, c.gws_bytes, c.stream); // M21 = -tmp M11
}

// shared scratch of the preconditioning (one per call; the matrices are preconditioned one after another)
struct PreShared {
  float *G, *M, *scratch, *Q1;
  void* gws;
  size_t gws_bytes;
};

static size_t pre_gws_bytes(int64_t r, int64_t L) {
  size_t a = tc_gemm_workspace_bytes(r, r, L, GRASP_PREC_F16X3);
  size_t b = tc_gemm_workspace_bytes(r, L, r, GRASP_PREC_F16X3);
  size_t c = tc_gemm_workspace_bytes(L, r, r, GRASP_PREC_F16X3);
  if (b > a) a = b;
  if (c > a) a = c;
  return (a + 1023) / 1024 * 1024;
}

static size_t pre_shared_bytes(int64_t r, int64_t L) {
  return (size_t)3 * r * r * 4 + (size_t)r * L * 4 + pre_gws_bytes(r, L) + 4 * 1024;
}

// Y0 = A (trans = 0, A is r x L) or A^T (trans = 1, A is L x r).  Writes Lm [r][r] = Y0 Q^T and Q [r][L].
// Three Cholesky passes: the first on the Gram shifted by 2e-5 of its mean diagonal, which keeps the fp32
// factorisation positive definite when cond(Y0)^2 > 1 / eps (square Gaussian matrices: cond ~ 3e4) and leaves a Q1
// of condition ~60; the second and third (unshifted) bring the rows of Q to orthonormality at rounding level.
static int svd_precondition(const float* A, int64_t lda, int trans, int r, int L, float* Lm, float* Q, uint32_t* fail,
                            const PreShared& sh, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr = false;
  if (!attr) {
    int rc = check_cuda(cudaFuncSetAttribute(svd_invchol_base_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IC_SMEM),
                        "svd_invchol attr");
    if (rc) return rc;
    attr = true;
  }
  const int P = GRASP_PREC_F16X3;
  IcCtx c{stream, sh.gws, sh.gws_bytes, fail};
  int rc = check_cuda(cudaMemsetAsync(fail, 0, 4, st), "svd pre memset");
  if (rc) return rc;
  const size_t rr = (size_t)r * r * 4;
  float shift_rel = 2e-5f;
  if (const char* e = getenv("GRASP_SVD_PRECOND_SHIFT")) shift_rel = (float)atof(e);
  // pass 1: G = Y0 Y0^T (+ shift), Q1 = inv(chol(G)) Y0.  Y0 as the left operand is (ta = trans, A); as op(B) = K x N
  // it is (tb = trans ? 1 : 0, A)
  rc = tc_gemm_f32(trans, trans ? 0 : 1, r, r, L, 1.f, A, lda, A, lda, 0.f, sh.G, r, 0, P, sh.gws, sh.gws_bytes, stream);
  if (rc) return rc;
  GRASP_LAUNCH(svd_shift_diag_kernel, dim3(1), dim3(1024), 0, st, sh.G, (int64_t)r, r, shift_rel);
  rc = check_cuda(cudaMemsetAsync(sh.M, 0, rr, st), "svd pre memset"); if (rc) return rc;
  rc = invchol_rec(sh.G, r, r, sh.M, r, sh.scratch, c); if (rc) return rc;
  rc = tc_gemm_f32(0, trans ? 1 : 0, r, L, r, 1.f, sh.M, r, A, lda, 0.f, Q, L, 0, P, sh.gws, sh.gws_bytes, stream);
  if (rc) return rc;
  // passes 2 and 3: Q -> Q1 -> Q
  for (int pass = 0; pass < 2; ++pass) {
    float* src = pass == 0 ? Q : sh.Q1;
    float* dst = pass == 0 ? sh.Q1 : Q;
    rc = tc_gemm_f32(0, 1, r, r, L, 1.f, src, L, src, L, 0.f, sh.G, r, 0, P, sh.gws, sh.gws_bytes, stream);
    if (rc) return rc;
    rc = check_cuda(cudaMemsetAsync(sh.M, 0, rr, st), "svd pre memset"); if (rc) return rc;
    rc = invchol_rec(sh.G, r, r, sh.M, r, sh.scratch, c); if (rc) return rc;
    rc = tc_gemm_f32(0, 0, r, L, r, 1.f, sh.M, r, src, L, 0.f, dst, L, 0, P, sh.gws, sh.gws_bytes, stream);
    if (rc) return rc;
  }
  // orthonormality of Q (also catches NaN / Inf from a failed factorisation)
  rc = tc_gemm_f32(0, 1, r, r, L, 1.f, Q, L, Q, L, 0.f, sh.G, r, 0, P, sh.gws, sh.gws_bytes, stream);
  if (rc) return rc;
  GRASP_LAUNCH(svd_orth_defect_kernel, dim3((unsigned)r), dim3(256), 0, st, (const float*)sh.G, (int64_t)r, r, 1e-4f, fail);
  // Lm = Y0 Q^T
  rc = tc_gemm_f32(trans, 1, r, r, L, 1.f, A, lda, Q, L, 0.f, Lm, r, 0, P, sh.gws, sh.gws_bytes, stream);
  if (rc) return rc;
  return check_cuda(cudaGetLastError(), "svd precondition");
}

struct SvdPlan {
  int64_t m, n;        // A is m x n
  int trans;           // 1 when m > n (work on A^T)
  int r, L, rp, Lp, ldz, p, npairs, nsplit, ntiles;
  size_t off_Z, off_G, off_ET, off_flag, off_stats, off_sigma, off_perm, off_Zp, off_ETp, off_T, off_gws, gws_bytes, bytes;
  // CholeskyQR2-preconditioned matrix: the plan above is the one of the square r x r factor Lm (trans = 0) and
  // these describe the original matrix
  int pre;             // 1 when preconditioned
  int trans0;          // original m > n
  int64_t m0, n0, L0;  // original shape and its long side
  size_t off_Lm, off_Q, off_prefail;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static SvdPlan make_plan_plain(int64_t m, int64_t n) {
  SvdPlan P{};
  P.m = m; P.n = n;
  P.trans = (m > n);
  P.r = (int)(P.trans ? n : m);
  P.L = (int)(P.trans ? m : n);
  P.rp = (int)round_up(P.r, JS);
  P.Lp = (int)round_up(P.L, UP_TN);
  P.ldz = P.Lp + (int)round_up(P.rp, UP_TN);
  P.p = P.rp / JB;
  P.npairs = P.p / 2;
  const int chunks = P.Lp / 32;
  int want = (2 * 148 + P.npairs - 1) / P.npairs;
  if (want < 1) want = 1;
  if (want > chunks / 4) want = chunks / 4 > 0 ? chunks / 4 : 1;
  if (want > 32) want = 32;
  P.nsplit = want;
  size_t o = 0;
  P.off_Z = o;     o = align_up(o + (size_t)P.rp * P.ldz * 4, 256);
  P.off_G = o;     o = align_up(o + (size_t)P.npairs * P.nsplit * JS * JS * 4, 256);
  P.off_ET = o;    o = align_up(o + (size_t)P.npairs * JS * JS * 4, 256);
  P.off_flag = o;  o = align_up(o + (size_t)P.npairs * 4, 256);
  P.off_stats = o; o = align_up(o + (size_t)J_STATS * 4, 256);
  P.off_sigma = o; o = align_up(o + (size_t)P.rp * 4, 256);
  P.off_perm = o;  o = align_up(o + (size_t)P.rp * 8, 256);
  // tensor-core path: Z and the block-diagonal ET as three bf16 planes each
  P.ntiles = (P.npairs + 1) / 2;
  o = align_up(o, 1024);
  P.off_Zp = o;    o = align_up(o + (size_t)3 * P.rp * P.ldz * 2, 1024);
  P.off_ETp = o;   o = align_up(o + (size_t)3 * P.ntiles * 128 * 128 * 2, 1024);
  // clean-up stage: T = QT QT^T and the workspace of its three GEMMs
  P.off_T = o;     o = align_up(o + (size_t)P.rp * P.rp * 4, 1024);
  size_t g1 = tc_gemm_workspace_bytes(P.rp, P.rp, P.rp, GRASP_PREC_BF16X6);
  size_t g2 = tc_gemm_workspace_bytes(P.rp, P.L, P.r, GRASP_PREC_BF16X6);
  const size_t g3 = tc_gemm_workspace_bytes(P.rp, P.rp, P.Lp, GRASP_PREC_F16X3);   // convergence check Y Y^T
  if (g3 > g2) g2 = g3;
  P.gws_bytes = align_up(g1 > g2 ? g1 : g2, 1024);
  P.off_gws = o;   o = align_up(o + P.gws_bytes, 1024);
  P.bytes = o;
  P.pre = 0; P.trans0 = P.trans; P.m0 = m; P.n0 = n; P.L0 = P.L;
  return P;
}

static bool pre_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GRASP_SVD_PRECOND"); v = e ? atoi(e) : 1; }
  return v != 0;
}

// Worth it when the rows shrink by a quarter or more (L + r -> 2 r) and the matrix is large enough for the tensor-core
// phase to dominate.  The Jacobi phase works on the rows of Lm^T, the triangular factor of the OTHER Gram (one step of
// the Cholesky-LR iteration towards the diagonal), which also saves sweeps: 17 -> 14 at 4096 x 4096, 16 -> 15 at
// 4096 x 11008 on the B200 (profiles/r02_svd_times_precond_v2.txt), where the rows of Lm itself (same Gram as Y0) need
// as many as the plain route.  For square matrices those three sweeps (~40 ms) just pay for the three Cholesky passes
// (~35 ms, launch-bound recursion), so they take the plain route unless GRASP_SVD_PRECOND_SQUARE=1.
// GRASP_SVD_PRECOND_T=0 factors Lm instead of Lm^T.
static bool pre_eligible(int64_t m, int64_t n) {
  static int sq = -1;
  if (sq < 0) { const char* e = getenv("GRASP_SVD_PRECOND_SQUARE"); sq = e ? atoi(e) : 0; }
  const int64_t r = m < n ? m : n, L = m < n ? n : m;
  return r >= 512 && (sq || 2 * L >= 3 * r);
}

static bool pre_transposed() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GRASP_SVD_PRECOND_T"); v = e ? atoi(e) : 1; }
  return v != 0;
}

static SvdPlan make_plan(int64_t m, int64_t n, bool allow_pre) {
  if (!(allow_pre && pre_eligible(m, n))) return make_plan_plain(m, n);
  const int64_t r = m < n ? m : n, L = m < n ? n : m;
  SvdPlan P = make_plan_plain(r, r);
  P.pre = 1; P.trans0 = (m > n); P.m0 = m; P.n0 = n; P.L0 = L;
  P.trans = pre_transposed() ? 1 : 0;        // the Jacobi phase factors Y0' = Lm^T (or Lm)
  size_t o = P.bytes;
  P.off_Lm = o;      o = align_up(o + (size_t)r * r * 4, 1024);
  P.off_Q = o;       o = align_up(o + (size_t)r * L * 4, 1024);
  P.off_prefail = o; o = align_up(o + 256, 1024);
  P.bytes = o;
  return P;
}

}  // namespace grasp

using namespace grasp;

extern "C" size_t grasp_svd_workspace_bytes(int batch, const int64_t* m, const int64_t* n) {
  if (batch <= 0 || !m || !n) return 0;
  // sized for either route (the caller may switch the preconditioning off per call with GRASP_SVD_NO_PRECOND)
  size_t total = 0, shared = 0;
  for (int i = 0; i < batch; ++i) {
    if (m[i] <= 0 || n[i] <= 0) return 0;
    const size_t a = make_plan(m[i], n[i], false).bytes, b = make_plan(m[i], n[i], true).bytes;
    total += a > b ? a : b;
    if (pre_eligible(m[i], n[i])) {
      const int64_t r = m[i] < n[i] ? m[i] : n[i], L = m[i] < n[i] ? n[i] : m[i];
      const size_t sb = pre_shared_bytes(r, L);
      if (sb > shared) shared = sb;
    }
  }
  return total + shared + 1024;
}

extern "C" int grasp_svd_batched(int batch, const float* const* A, const int64_t* m, const int64_t* n,
                                 const int64_t* lda, float* const* U, float* const* S, float* const* Vh,
                                 int32_t* info, int prec, int max_sweeps, void* ws, size_t ws_bytes,
                                 void* stream) {
  if (batch < 0) return bad_arg("svd: batch");
  if (batch == 0) return 0;
  if (!A || !m || !n || !lda || !U || !S || !Vh || !ws) return bad_arg("svd: null");
  const bool no_precond = (prec & GRASP_SVD_NO_PRECOND) != 0;
  prec &= ~GRASP_SVD_NO_PRECOND;
  if (prec != GRASP_PREC_SIMT && prec != GRASP_PREC_BF16X3 && prec != GRASP_PREC_BF16X6 && prec != GRASP_PREC_F16X3)
    return bad_arg("svd: prec");
  if (max_sweeps <= 0) max_sweeps = 32;
  if (max_sweeps > J_STATS - 16) max_sweeps = J_STATS - 16;   // 8 status words + 8 clean-up sweeps
  for (int i = 0; i < batch; ++i) {
    if (!A[i] || !U[i] || !S[i] || !Vh[i]) return bad_arg("svd: null matrix pointer");
    if (m[i] <= 0 || n[i] <= 0 || lda[i] < n[i]) return bad_arg("svd: m/n/lda");
    if (m[i] > 65536 || n[i] > 65536) return bad_arg("svd: dimension > 65536");
    if ((m[i] < n[i] ? m[i] : n[i]) > 16384) return bad_arg("svd: min(m,n) > 16384 unsupported");
  }
  if (ws_bytes < grasp_svd_workspace_bytes(batch, m, n)) return bad_arg("svd: workspace too small");
  if (reinterpret_cast<uintptr_t>(ws) & 255) return bad_arg("svd: workspace must be 256-byte aligned");

  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(svd_evd_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(EvdSmem<double>)), "svd_evd attr");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(svd_evd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(EvdSmem<float>)), "svd_evd attr");
    if (rc) return rc;
    attr_set = true;
  }
  // convergence: every pair's |g_pq| / sqrt(g_pp g_qq) below tol (the fp32 Gram noise floor is ~5e-7
  // at L = 4096, so tighter values never trigger).  Both knobs can be overridden for experiments.
  float tol = 1e-6f;
  int inner_cap = 2;
  if (const char* e = getenv("GRASP_SVD_TOL")) tol = (float)atof(e);
  if (const char* e = getenv("GRASP_SVD_INNER_CAP")) inner_cap = atoi(e);
  cudaStream_t st = (cudaStream_t)stream;
  bool use_tc = (prec != GRASP_PREC_SIMT);
  if (const char* e = getenv("GRASP_SVD_TC")) use_tc = atoi(e) != 0;
  const bool evd64_dbg = getenv("GRASP_SVD_EVD64") != nullptr;      // experiments only
  int tc_inner_cap = 1;   // one inner Jacobi sweep per pair visit in the tensor-core phase (measured fastest overall)
  if (const char* e = getenv("GRASP_SVD_TC_INNER_CAP")) tc_inner_cap = atoi(e);
  // clean-up sweeps on the tensor cores: 10% faster but the accumulator's truncation (also on the exact
  // identity part, p0+p1+p2 can span more than 24 bits) leaves ~2e-5 instead of ~2e-6 -> off by default
  bool tc_cleanup = false;
  if (const char* e = getenv("GRASP_SVD_TC_CLEANUP")) tc_cleanup = atoi(e) != 0;
  const bool no_cleanup_dbg = getenv("GRASP_SVD_NO_CLEANUP") != nullptr;
  bool gemm_check = true;   // confirm convergence of the clean-up with one Gram GEMM instead of a second sweep
  if (const char* e = getenv("GRASP_SVD_GEMM_CHECK")) gemm_check = atoi(e) != 0;
  bool evd_warp = true;   // register-resident 2-warp eigen-solve in the tensor-core phase (0: the 1024-thread kernel)
  if (const char* e = getenv("GRASP_SVD_EVD_WARP")) evd_warp = atoi(e) != 0;
  if (use_tc) {
    static bool tc_attr = false;
    if (!tc_attr) {
      int rc = check_cuda(cudaFuncSetAttribute(jacobi_tc_kernel<JT_GRAM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               JtCfg<JT_GRAM>::SMEM_BYTES), "jacobi_tc gram attr");
      if (rc) return rc;
      rc = check_cuda(cudaFuncSetAttribute(jacobi_tc_kernel<JT_UPDATE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           JtCfg<JT_UPDATE>::SMEM_BYTES), "jacobi_tc update attr");
      if (rc) return rc;
      rc = check_cuda(cudaFuncSetAttribute(jacobi_tc_kernel<JT_GRAM3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           JtCfg<JT_GRAM3>::SMEM_BYTES), "jacobi_tc gram3 attr");
      if (rc) return rc;
      rc = check_cuda(cudaFuncSetAttribute(jacobi_tc_kernel<JT_UPDATE2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           JtCfg<JT_UPDATE2>::SMEM_BYTES), "jacobi_tc update2 attr");
      if (rc) return rc;
      tc_attr = true;
    }
  }

  // plans + workspace carving
  SvdPlan* plans = new SvdPlan[batch];
  unsigned char** base = new unsigned char*[batch];
  bool* done = new bool[batch];
  const float** Aeff = new const float*[batch];     // what the Jacobi phase factors: A itself, or the square Lm
  int64_t* ldaeff = new int64_t[batch];
  const bool allow_pre = use_tc && pre_enabled() && !no_precond;
  PreShared sh{};
  int rc = 0;
  {
    unsigned char* cur = static_cast<unsigned char*>(ws);
    int64_t rmax = 0, rlmax = 0;
    size_t gmax = 0;
    for (int i = 0; i < batch; ++i) {
      plans[i] = make_plan(m[i], n[i], allow_pre);
      base[i] = cur;
      cur += plans[i].bytes;
      done[i] = false;
      Aeff[i] = A[i];
      ldaeff[i] = lda[i];
      if (plans[i].pre) {
        const int64_t r = plans[i].r, L = plans[i].L0;
        if (r > rmax) rmax = r;
        if (r * L > rlmax) rlmax = r * L;
        const size_t gb = pre_gws_bytes(r, L);
        if (gb > gmax) gmax = gb;
      }
    }
    if (rmax > 0) {
      cur = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cur) + 1023) & ~(uintptr_t)1023);
      sh.G = reinterpret_cast<float*>(cur);        cur += (size_t)rmax * rmax * 4;
      sh.M = reinterpret_cast<float*>(cur);        cur += (size_t)rmax * rmax * 4;
      sh.scratch = reinterpret_cast<float*>(cur);  cur += (size_t)rmax * rmax * 4;
      sh.Q1 = reinterpret_cast<float*>(cur);       cur += (size_t)rlmax * 4;
      cur = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cur) + 1023) & ~(uintptr_t)1023);
      sh.gws = cur; sh.gws_bytes = gmax;
      for (int i = 0; i < batch && !rc; ++i) {
        if (!plans[i].pre) continue;
        float* Lm = reinterpret_cast<float*>(base[i] + plans[i].off_Lm);
        rc = svd_precondition(A[i], lda[i], plans[i].trans0, plans[i].r, (int)plans[i].L0, Lm,
                              reinterpret_cast<float*>(base[i] + plans[i].off_Q),
                              reinterpret_cast<uint32_t*>(base[i] + plans[i].off_prefail), sh, stream);
        Aeff[i] = Lm;
        ldaeff[i] = plans[i].r;
      }
    }
  }
  for (int i0 = 0; i0 < batch && !rc; ++i0) {
    if (done[i0]) continue;
    // group up to J_MAXMAT matrices with the same working shape (a preconditioned wide matrix works on its
    // square factor and shares launches with square matrices of the same order)
    SvdGroup g{};
    int members[J_MAXMAT];
    const SvdPlan& P = plans[i0];
    for (int i = i0; i < batch && g.nmat < J_MAXMAT; ++i) {
      if (done[i]) continue;
      if (plans[i].rp == P.rp && plans[i].Lp == P.Lp) {
        members[g.nmat] = i;
        SvdMat& M = g.mat[g.nmat++];
        M.Z = reinterpret_cast<float*>(base[i] + plans[i].off_Z);
        M.Gpart = reinterpret_cast<float*>(base[i] + plans[i].off_G);
        M.ET = reinterpret_cast<float*>(base[i] + plans[i].off_ET);
        M.pair_flag = reinterpret_cast<int*>(base[i] + plans[i].off_flag);
        M.stats = reinterpret_cast<uint32_t*>(base[i] + plans[i].off_stats);
        M.ETp = use_tc ? reinterpret_cast<__nv_bfloat16*>(base[i] + plans[i].off_ETp) : nullptr;
        done[i] = true;
      }
    }
    g.rp = P.rp; g.Lp = P.Lp; g.ldz = P.ldz; g.p = P.p; g.npairs = P.npairs; g.nsplit = P.nsplit;

    for (int j = 0; j < g.nmat && !rc; ++j) {
      const int i = members[j];
      const SvdPlan& Q = plans[i];
      rc = check_cuda(cudaMemsetAsync(g.mat[j].stats, 0, J_STATS * 4, st), "svd memset");
      if (rc) break;
      dim3 grid((unsigned)(Q.ldz / 32), (unsigned)(Q.rp / 32));
      GRASP_LAUNCH(svd_init_kernel, grid, dim3(32, 8), 0, st, Aeff[i], ldaeff[i], Q.r, Q.L, Q.trans,
                   g.mat[j].Z, Q.rp, Q.Lp, Q.ldz);
    }
    if (rc) break;
    rc = check_cuda(cudaGetLastError(), "svd_init_kernel");
    if (rc) break;

    JtMaps* maps = nullptr;
    JtParams jp{};
    if (use_tc && g.p >= 2) {
      maps = new JtMaps;
      jp.nmat = g.nmat; jp.rp = g.rp; jp.Lp = g.Lp; jp.ldz = g.ldz; jp.p = g.p; jp.npairs = g.npairs;
      jp.ntiles = P.ntiles; jp.nsplit = g.nsplit;
      for (int j = 0; j < g.nmat && !rc; ++j) {
        const SvdPlan& Q = plans[members[j]];
        __nv_bfloat16* Zp = reinterpret_cast<__nv_bfloat16*>(base[members[j]] + Q.off_Zp);
        jp.mat[j].Zp = Zp; jp.mat[j].Gpart = g.mat[j].Gpart;
        jp.mat[j].pair_flag = g.mat[j].pair_flag; jp.mat[j].stats = g.mat[j].stats;
        const int64_t n4 = (int64_t)Q.rp * Q.ldz / 4;
        GRASP_LAUNCH(jt_split_kernel, dim3((unsigned)ceil_div(n4, 256)), dim3(256), 0, st, g.mat[j].Z, Q.rp, Q.ldz, Zp);
        rc = check_cuda(cudaMemsetAsync(g.mat[j].ETp, 0, (size_t)3 * Q.ntiles * 128 * 128 * 2, st), "svd ETp memset");
        if (rc) break;
        rc = tc_make_map_4d(&maps->z[j], Zp, 128, (uint64_t)Q.rp, (uint64_t)Q.ldz / 128, 3, 256, (uint64_t)Q.rp * 256,
                            (uint64_t)Q.rp * Q.ldz * 2, 64, 32);
        if (rc) break;
        rc = tc_make_map_3d(&maps->et[j], g.mat[j].ETp, 128, (uint64_t)Q.ntiles * 128, 3, 128 * 2,
                            (uint64_t)Q.ntiles * 128 * 128 * 2, 64, 128);
      }
      for (int j = g.nmat; j < J_MAXMAT && !rc; ++j) { maps->z[j] = maps->z[0]; maps->et[j] = maps->et[0]; }
      if (rc) { delete maps; break; }
    }
    if (g.p >= 2) {
      const int tc_grid_g = min(sm_count(), g.nmat * P.ntiles * g.nsplit);
      const int tc_grid_u = min(sm_count(), g.nmat * P.ntiles * (g.ldz / 128));
      // Two-plane updates (JT_UPDATE2: 8 instead of 12 bytes per element and round) in the first sweeps were
      // measured and rejected: six such sweeps at n = 4096 leave QT orthogonal to ~1e-2 only, beyond what the single
      // Newton-Schulz step of the clean-up repairs -- sigma error 4.4e-5 instead of 2e-7, reconstruction 8.8e-5
      // instead of 1.8e-6, and two more sweeps (profiles/r02_svd_times_two_plane_sweeps.txt).  Off unless
      // GRASP_SVD_2PLANE_SWEEPS=n asks for n of them.
      int two_plane_sweeps = 0;
      if (use_tc) {
        if (const char* e = getenv("GRASP_SVD_2PLANE_SWEEPS")) two_plane_sweeps = atoi(e);
      }
      for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        if (use_tc && sweep == two_plane_sweeps && two_plane_sweeps > 0) {
          // the third plane was not maintained: from here on Z = p0 + p1 exactly
          for (int j = 0; j < g.nmat && !rc; ++j) {
            const SvdPlan& Q = plans[members[j]];
            const size_t plane_bytes = (size_t)Q.rp * Q.ldz * 2;
            rc = check_cuda(cudaMemsetAsync(base[members[j]] + Q.off_Zp + 2 * plane_bytes, 0, plane_bytes, st), "svd plane memset");
          }
          if (rc) break;
        }
        for (int round = 0; round < g.p - 1; ++round) {
          if (use_tc) {
            jp.round = round;
            GRASP_LAUNCH(jacobi_tc_kernel<JT_GRAM>, dim3(tc_grid_g), dim3(JT_THREADS), JtCfg<JT_GRAM>::SMEM_BYTES, st,
                         *maps, jp);
          } else {
            GRASP_LAUNCH(svd_gram_kernel<float>, dim3(g.nsplit, g.npairs, g.nmat), dim3(J_THREADS), 0, st, g, round);
          }
          if (use_tc && !evd64_dbg && evd_warp)
            GRASP_LAUNCH(svd_evd_warp_kernel, dim3(g.npairs, g.nmat), dim3(EVW_THREADS), 0, st, g, round, sweep, tol,
                         tc_inner_cap);
          else if (use_tc && !evd64_dbg)
            GRASP_LAUNCH(svd_evd_kernel<float>, dim3(g.npairs, g.nmat), dim3(EVD_THREADS), sizeof(EvdSmem<float>), st, g,
                         round, sweep, tol, tc_inner_cap);
          else
            GRASP_LAUNCH(svd_evd_kernel<double>, dim3(g.npairs, g.nmat), dim3(EVD_THREADS), sizeof(EvdSmem<double>), st,
                         g, round, sweep, tol, inner_cap);
          if (use_tc && sweep < two_plane_sweeps) {
            GRASP_LAUNCH(jacobi_tc_kernel<JT_UPDATE2>, dim3(tc_grid_u), dim3(JT_THREADS), JtCfg<JT_UPDATE2>::SMEM_BYTES,
                         st, *maps, jp);
          } else if (use_tc) {
            GRASP_LAUNCH(jacobi_tc_kernel<JT_UPDATE>, dim3(tc_grid_u), dim3(JT_THREADS), JtCfg<JT_UPDATE>::SMEM_BYTES,
                         st, *maps, jp);
          } else {
            GRASP_LAUNCH(svd_update_kernel, dim3(g.ldz / UP_TN, g.npairs, g.nmat), dim3(J_THREADS), 0, st, g, round);
          }
        }
        // tensor-core phase: stop once the Gram is diagonal to 1e-4 (its own drift is of that order anyway)
        GRASP_LAUNCH(svd_sweep_end_kernel, dim3(1), dim3(32), 0, st, g, sweep, use_tc ? 1e-4f : tol, -1);
      }
      if (rc) { delete maps; break; }
      if (use_tc && no_cleanup_dbg) {
        for (int j = 0; j < g.nmat; ++j) {
          const SvdPlan& Q = plans[members[j]];
          const int64_t n4 = (int64_t)Q.rp * Q.ldz / 4;
          GRASP_LAUNCH(jt_merge_kernel, dim3((unsigned)ceil_div(n4, 256)), dim3(256), 0, st,
                       reinterpret_cast<const __nv_bfloat16*>(base[members[j]] + Q.off_Zp), Q.rp, Q.ldz, g.mat[j].Z);
        }
      } else if (use_tc) {
        // ---- clean-up.  The tensor-core accumulator truncates, so every update shrinks Z by ~1e-7:
        // after thousands of updates QT is only orthogonal to ~3e-4 and Y has drifted from QT*Y0 by as
        // much.  One Newton-Schulz step re-orthogonalises QT (error -> its square), Y is recomputed as
        // QT*Y0, and a few CUDA-core (unbiased fp32) sweeps finish from an almost diagonal Gram.
        for (int j = 0; j < g.nmat && !rc; ++j) {
          const int i = members[j];
          const SvdPlan& Q = plans[i];
          float* Zm = g.mat[j].Z;
          float* QT = Zm + Q.Lp;
          const int64_t n4 = (int64_t)Q.rp * Q.ldz / 4;
          GRASP_LAUNCH(jt_merge_kernel, dim3((unsigned)ceil_div(n4, 256)), dim3(256), 0, st,
                       reinterpret_cast<const __nv_bfloat16*>(base[i] + Q.off_Zp), Q.rp, Q.ldz, Zm);
          float* T = reinterpret_cast<float*>(base[i] + Q.off_T);
          void* gws = base[i] + Q.off_gws;
          rc = tc_gemm_f32(0, 1, Q.rp, Q.rp, Q.rp, 1.f, QT, Q.ldz, QT, Q.ldz, 0.f, T, Q.rp, 0, GRASP_PREC_BF16X6, gws,
                           Q.gws_bytes, stream);
          if (rc) break;
          rc = tc_gemm_f32(0, 0, Q.rp, Q.rp, Q.rp, -0.5f, T, Q.rp, QT, Q.ldz, 1.5f, QT, Q.ldz, 0, GRASP_PREC_BF16X6, gws,
                           Q.gws_bytes, stream);
          if (rc) break;
          rc = tc_gemm_f32(0, Q.trans ? 1 : 0, Q.rp, Q.L, Q.r, 1.f, QT, Q.ldz, Aeff[i], ldaeff[i], 0.f, Zm, Q.ldz, 0,
                           GRASP_PREC_BF16X6, gws, Q.gws_bytes, stream);
          if (!tc_cleanup) g.mat[j].ETp = nullptr;
          else GRASP_LAUNCH(jt_split_kernel, dim3((unsigned)ceil_div(n4, 256)), dim3(256), 0, st, Zm, Q.rp, Q.ldz,
                            reinterpret_cast<__nv_bfloat16*>(base[i] + Q.off_Zp));
        }
        if (rc) { delete maps; break; }
        GRASP_LAUNCH(svd_reopen_kernel, dim3(1), dim3(32), 0, st, g);
        // after the re-orthogonalisation the Gram is diagonal to ~3e-4, so Jacobi's quadratic convergence
        // needs two sweeps, by default on the CUDA cores (unbiased fp32 FMA; fp64 eigen-solve).
        // GRASP_SVD_TC_CLEANUP=1 runs them on the tensor cores (three-plane Gram) instead.
        const int extra = 3;
        const float cleanup_conv = 3e-6f;
        // rotation threshold of the clean-up: with the exact (fp64-accumulated) Gram every pair above 1e-7 is
        // rotated; with the fp32 Gram 1e-6 is its noise floor.  GRASP_SVD_CLEANUP_GRAM64=0 restores round 1.
        bool cleanup_gram64 = true;
        if (const char* e = getenv("GRASP_SVD_CLEANUP_GRAM64")) cleanup_gram64 = atoi(e) != 0;
        float cleanup_tol = cleanup_gram64 ? 1e-7f : tol;
        if (const char* e = getenv("GRASP_SVD_CLEANUP_TOL")) cleanup_tol = (float)atof(e);
        for (int s2 = 0; s2 < extra; ++s2) {
          const int sweep = max_sweeps + s2;
          for (int round = 0; round < g.p - 1; ++round) {
            if (tc_cleanup) {
              jp.round = round;
              GRASP_LAUNCH(jacobi_tc_kernel<JT_GRAM3>, dim3(tc_grid_g), dim3(JT_THREADS), JtCfg<JT_GRAM3>::SMEM_BYTES, st,
                           *maps, jp);
            } else if (cleanup_gram64) {
              GRASP_LAUNCH(svd_gram_kernel<double>, dim3(g.nsplit, g.npairs, g.nmat), dim3(J_THREADS), 0, st, g, round);
            } else {
              GRASP_LAUNCH(svd_gram_kernel<float>, dim3(g.nsplit, g.npairs, g.nmat), dim3(J_THREADS), 0, st, g, round);
            }
            GRASP_LAUNCH(svd_evd_kernel<double>, dim3(g.npairs, g.nmat), dim3(EVD_THREADS), sizeof(EvdSmem<double>), st,
                         g, round, sweep, cleanup_tol, inner_cap);
            if (tc_cleanup) {
              GRASP_LAUNCH(jacobi_tc_kernel<JT_UPDATE>, dim3(tc_grid_u), dim3(JT_THREADS), JtCfg<JT_UPDATE>::SMEM_BYTES,
                           st, *maps, jp);
            } else {
              GRASP_LAUNCH(svd_update_kernel, dim3(g.ldz / UP_TN, g.npairs, g.nmat), dim3(J_THREADS), 0, st, g, round);
            }
          }
          GRASP_LAUNCH(svd_sweep_end_kernel, dim3(1), dim3(32), 0, st, g, sweep, cleanup_conv, s2);
          if (s2 + 1 < extra && !tc_cleanup && gemm_check) {
            // A sweep only knows the off-diagonals it met BEFORE rotating them, so confirming convergence costs
            // another full sweep (Gram + eigen-solve of every pair, ~45% of a sweep).  The whole Gram as one
            // fp32-class tensor-core GEMM plus a max-reduction gives the same answer for ~1 ms per matrix.
            // Also after the second sweep: about a third of the 4096 x 4096 matrices leave the tensor-core phase just
            // under its 1e-4 bar and need two clean-up sweeps; without this check a third one would only confirm them
            // (profiles/r02_svd_batch_sizes.txt: 16 + 3 sweeps).
            for (int j = 0; j < g.nmat && !rc; ++j) {
              const SvdPlan& Q = plans[members[j]];
              float* T = reinterpret_cast<float*>(base[members[j]] + Q.off_T);
              rc = tc_gemm_f32(0, 1, Q.rp, Q.rp, Q.Lp, 1.f, g.mat[j].Z, Q.ldz, g.mat[j].Z, Q.ldz, 0.f, T, Q.rp, 0,
                               GRASP_PREC_F16X3, base[members[j]] + Q.off_gws, Q.gws_bytes, stream);
              if (rc) break;
              GRASP_LAUNCH(svd_offdiag_max_kernel, dim3((unsigned)Q.r), dim3(256), 0, st, T, Q.rp, Q.r, g.mat[j].stats);
            }
            if (rc) break;
            GRASP_LAUNCH(svd_check_end_kernel, dim3(1), dim3(32), 0, st, g, cleanup_conv);
          }
        }
        if (rc) { delete maps; break; }
        if (tc_cleanup) {
          for (int j = 0; j < g.nmat; ++j) {
            const SvdPlan& Q = plans[members[j]];
            const int64_t n4 = (int64_t)Q.rp * Q.ldz / 4;
            GRASP_LAUNCH(jt_merge_kernel, dim3((unsigned)ceil_div(n4, 256)), dim3(256), 0, st,
                         reinterpret_cast<const __nv_bfloat16*>(base[members[j]] + Q.off_Zp), Q.rp, Q.ldz, g.mat[j].Z);
          }
        }
      }
      delete maps;
      rc = check_cuda(cudaGetLastError(), "svd sweep kernels");
      if (rc) break;
    }

    // finalize each member
    for (int j = 0; j < g.nmat && !rc; ++j) {
      const int i = members[j];
      const SvdPlan& Q = plans[i];
      float* sigma = reinterpret_cast<float*>(base[i] + Q.off_sigma);
      int64_t* perm = reinterpret_cast<int64_t*>(base[i] + Q.off_perm);
      GRASP_LAUNCH(svd_norms_kernel, dim3((Q.rp + 7) / 8), dim3(256), 0, st, g.mat[j].Z, Q.ldz, Q.rp, Q.Lp, sigma);
      const float* sc = sigma;
      int64_t rr = Q.rp, kk = Q.rp;
      rc = grasp_topk_batched(1, &sc, &rr, &kk, &perm, stream);
      if (rc) break;
      if (Q.pre) {
        // The Jacobi phase factored Y0' = Usq S Wt (Wt = normalised rows of Y, Usq = QT^T) with Y0' = Lm or Lm^T,
        // and Y0 = Lm Qb.  Let (Ul, Vlt) be the factors of Lm = Ul S Vlt: (Usq, Wt) or (Wt^T, Usq^T).  Then
        //   A = Y0   (trans0 = 0):  U = Ul,           Vh = Vlt Qb
        //   A = Y0^T (trans0 = 1):  U = Qb^T Vlt^T,   Vh = Ul^T
        const int r = Q.r;
        const int64_t L0 = Q.L0;
        float* Wb = sh.G;                                             // [r][r] scratch, free since the preconditioning
        const float* Qb = reinterpret_cast<const float*>(base[i] + Q.off_Q);
        const int lt = Q.trans;                                       // 1: Y0' = Lm^T
        // rows of Vlt: normalised rows of Y (lt = 0) or rows of QT (lt = 1)
        const int vl_col0 = lt ? Q.Lp : 0, vl_norm = lt ? 0 : 1;
        // Ul: QT^T (lt = 0) or the normalised rows of Y transposed (lt = 1)
        const int ul_col0 = lt ? 0 : Q.Lp, ul_norm = lt ? 1 : 0;
        if (!Q.trans0) {
          GRASP_LAUNCH(svd_emit_rows_kernel, dim3(r), dim3(256), 0, st, g.mat[j].Z, Q.ldz, vl_col0, perm, sigma, vl_norm, r, r,
                       Wb, (int64_t)r, S[i]);
          GRASP_LAUNCH(svd_emit_cols_kernel, dim3((r + 31) / 32, (unsigned)((r + 31) / 32)), dim3(32, 8), 0, st,
                       g.mat[j].Z, Q.ldz, ul_col0, perm, sigma, ul_norm, r, r, U[i], (int64_t)r);
          rc = tc_gemm_f32(0, 0, r, L0, r, 1.f, Wb, r, Qb, L0, 0.f, Vh[i], Q.n0, 0, GRASP_PREC_F16X3, sh.gws, sh.gws_bytes,
                           stream);
        } else {
          GRASP_LAUNCH(svd_emit_rows_kernel, dim3(r), dim3(256), 0, st, g.mat[j].Z, Q.ldz, vl_col0, perm, sigma, vl_norm, r, r,
                       Wb, (int64_t)r, (float*)nullptr);
          // Vh = Ul^T: row i of Vh is column i of Ul = (sorted) row i of QT (lt = 0) / normalised row i of Y (lt = 1)
          GRASP_LAUNCH(svd_emit_rows_kernel, dim3(r), dim3(256), 0, st, g.mat[j].Z, Q.ldz, ul_col0, perm, sigma, ul_norm, r, r,
                       Vh[i], (int64_t)Q.n0, S[i]);
          rc = tc_gemm_f32(1, 1, L0, r, r, 1.f, Qb, L0, Wb, r, 0.f, U[i], r, 0, GRASP_PREC_F16X3, sh.gws, sh.gws_bytes,
                           stream);
        }
        if (rc) break;
      } else if (!Q.trans) {
        // Vh = normalised rows of Y, U = QT^T
        GRASP_LAUNCH(svd_emit_rows_kernel, dim3(Q.r), dim3(256), 0, st, g.mat[j].Z, Q.ldz, 0, perm, sigma, 1, Q.r,
                     (int)Q.n, Vh[i], (int64_t)Q.n, S[i]);
        GRASP_LAUNCH(svd_emit_cols_kernel, dim3((Q.r + 31) / 32, (unsigned)((Q.m + 31) / 32)), dim3(32, 8), 0, st,
                     g.mat[j].Z, Q.ldz, Q.Lp, perm, sigma, 0, Q.r, (int)Q.m, U[i], (int64_t)Q.r);
      } else {
        // U = (normalised rows of Y)^T, Vh = QT
        GRASP_LAUNCH(svd_emit_cols_kernel, dim3((Q.r + 31) / 32, (unsigned)((Q.m + 31) / 32)), dim3(32, 8), 0, st,
                     g.mat[j].Z, Q.ldz, 0, perm, sigma, 1, Q.r, (int)Q.m, U[i], (int64_t)Q.r);
        GRASP_LAUNCH(svd_emit_rows_kernel, dim3(Q.r), dim3(256), 0, st, g.mat[j].Z, Q.ldz, Q.Lp, perm, sigma, 0, Q.r,
                     (int)Q.n, Vh[i], (int64_t)Q.n, S[i]);
      }
      if (info)
        GRASP_LAUNCH(svd_info_kernel, dim3(1), dim3(1), 0, st, g.mat[j].stats,
                     Q.pre ? reinterpret_cast<const uint32_t*>(base[i] + Q.off_prefail) : (const uint32_t*)nullptr,
                     info + 4 * i);
      rc = check_cuda(cudaGetLastError(), "svd finalize");
    }
  }
  delete[] plans;
  delete[] base;
  delete[] done;
  delete[] Aeff;
  delete[] ldaeff;
  return rc;
}
