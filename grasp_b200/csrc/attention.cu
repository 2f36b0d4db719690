// Causal self-attention of the calibration passes on the sm_100a tensor cores, fp32-class arithmetic.
//
// The reference runs transformers' LlamaAttention inside `self.model(...)` + `loss.backward()`
// (reference modeling_grasp.py:347-354); in fp32 torch dispatches that to its sm_80 memory-efficient kernels
// (30-37 TFLOP/s on a B200, 28 % of a deep pass).  Here Q K^T, P V and the five products of the backward are
// tcgen05.mma on fp16 (hi, lo) planes -- x * s = hi + lo with one power-of-two scale per tensor, products lo*hi,
// hi*lo, hi*hi accumulated in fp32 in TMEM (the F16X3 arithmetic of gemm_tc.cu, ~3e-7) -- with the softmax in
// fp32 registers in between.  The planes of q, k, v (after RoPE) and of dO come from grasp_gemm_split_f16 with
// GRASP_SCALE_TENSOR on the [tokens, heads * head_dim] activations, so one tile in shared memory serves as a
// K-major operand (contracted over head_dim) and as an MN-major operand (contracted over tokens).
//
// Forward, one CTA per (batch, head, 128 queries), 320 threads:
//   warp 0    TMA producer: the Q tile once, then (K_j, V_j) tiles of 64 keys through a two-stage ring
//   warp 1    MMA issuer:   S_j = Q K_j^T (128 x 64, TMEM);  O_j = P_j V_j (128 x D, TMEM)
//   warps 2-9 softmax:      two threads per query row (= TMEM lane), 32 score columns each: S_j -> online max / sum
//                           (row maximum exchanged through shared memory) -> P_j as fp16 planes
//                           (x 2^14) in the K-major swizzled layout the MMA reads; O_j is drained into fp32
//                           registers, O = O * corr + O_j (round-to-nearest adds, no truncating TMEM chain)
// Keys beyond the causal diagonal are masked, so tiles right of the diagonal are never loaded; rows past the end
// of a sequence (S not a multiple of 128) are computed on whatever the tile holds and not stored.
// Outputs: O [tokens, H * D] fp32 and lse2 [B, H, S] = base-2 log-sum-exp of the scaled scores (for the backward).
#include "tc_common.cuh"
#include "split_f16.cuh"
#include <math.h>

namespace grasp {

using namespace tc;

constexpr int AT_BQ = 128;        // queries per CTA = TMEM lanes
constexpr int AT_BK = 64;         // keys per step
constexpr int AT_SM_WARPS = 8;    // softmax / epilogue warps: two per TMEM lane quadrant, each half of the columns
constexpr int AT_SM_THREADS = 32 * AT_SM_WARPS;
constexpr int AT_THREADS = 64 + AT_SM_THREADS;   // + TMA warp + MMA warp
constexpr int AT_HC = AT_BK / 2;   // score columns per softmax thread
constexpr float AT_P_SCALE = 16384.f;          // probabilities (<= 1) as fp16 planes: p * 2^14 = hi + lo
constexpr float AT_P_INV = 1.f / 16384.f;

// 2^x on the SFU (rel. error ~2^-22, -inf -> 0)
__device__ __forceinline__ float at_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 8 fp32 values -> 8 fp16 hi and 8 fp16 lo (x = hi + lo), packed in element order (two conversions per pair)
__device__ __forceinline__ void at_split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 h2 = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
    const float2 f = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(v[2 * e] - f.x, v[2 * e + 1] - f.y);
    h[e] = *reinterpret_cast<const uint32_t*>(&h2);
    l[e] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// 32 fp32 values of this thread's half row -> (hi, lo) fp16 planes of a [128][64] K-major tile with the 128-byte
// swizzle (16-byte chunk c8 of row r sits at c8 ^ (r & 7))
// (this thread writes the 32 columns [32 * half, 32 * half + 32) of its row: chunks 4 * half .. 4 * half + 3)
__device__ __forceinline__ void at_store_planes(unsigned char* T, int plane_bytes, int row, int half, const float* v) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int c8 = half * 4 + c;
    uint4 hi, lo;
    at_split8(v + c * 8, hi, lo);
    const int off = row * 128 + ((c8 ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(T + off) = hi;
    *reinterpret_cast<uint4*>(T + plane_bytes + off) = lo;
  }
}

struct AttnParams {
  int B, S, H, Hkv;
  float scale_log2;               // softmax scale * log2(e)
  const float* inv_q;             // inverse plane scales (tensor-scaled: element 0)
  const float* inv_k;
  const float* inv_v;
  float* out;                     // [B*S][ld_out]
  int64_t ld_out;
  float* lse2;                    // [B][H][S]
};

template <int D>
struct AtFwdCfg {
  static constexpr int CH = D / 64;                    // 64-wide chunks of the head dimension (one swizzle row each)
  static constexpr int Q_PLANE = CH * AT_BQ * 128;     // bytes of one plane of the Q tile
  static constexpr int KV_PLANE = CH * AT_BK * 128;    // bytes of one plane of a K or V tile
  static constexpr int STAGE_BYTES = 4 * KV_PLANE;     // K (2 planes) + V (2 planes)
  static constexpr int STAGES = 2;
  static constexpr int P_PLANE = AT_BQ * 128;          // [128 q][64 keys] fp16
  static constexpr int OFF_KV = 2 * Q_PLANE;
  static constexpr int OFF_P = OFF_KV + STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_P + 2 * P_PLANE + 1024;
  static constexpr int TMEM_COLS = (64 + D) <= 128 ? 128 : 256;
  static_assert(D == 64 || D == 128, "head dimensions 64 and 128");
};

template <int D>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                const __grid_constant__ CUtensorMap mapV, const AttnParams p) {
  using Cfg = AtFwdCfg<D>;
  constexpr int CH = Cfg::CH;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* Qs = smem;
  unsigned char* KVs = smem + Cfg::OFF_KV;
  unsigned char* Ps = smem + Cfg::OFF_P;
  __shared__ uint64_t q_full, kv_full[Cfg::STAGES], kv_empty[Cfg::STAGES], s_full, p_full, o_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float xch[2][AT_BQ];        // [column half][row]: partial row maxima, then partial sums (1 KB: 224 KB + barriers fill the CTA's shared memory)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (p.S + AT_BQ - 1) / AT_BQ;
  const int qt = nqt - 1 - (int)blockIdx.x;            // the longest tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qt * AT_BQ;
  const int kv_end = min(p.S, q0 + AT_BQ);             // keys [0, kv_end) can be attended by this tile
  const int nk = (kv_end + AT_BK - 1) / AT_BK;
  const int tok0 = b * p.S;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapQ); prefetch_tmap(&mapK); prefetch_tmap(&mapV);
    mbar_init(&q_full, 1);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(&s_full, 1); mbar_init(&o_full, 1); mbar_init(&p_full, AT_SM_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(&tmem_base_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(&q_full, 2 * Cfg::Q_PLANE);
#pragma unroll
      for (int pl = 0; pl < 2; ++pl)
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
          tma_load_3d(Qs + pl * Cfg::Q_PLANE + ch * (AT_BQ * 128), &mapQ, &q_full, h * D + ch * 64, tok0 + q0, pl);
      for (int j = 0; j < nk; ++j) {
        const int stage = j & 1;
        mbar_wait(&kv_empty[stage], ((j >> 1) & 1) ^ 1);
        unsigned char* Ks = KVs + stage * Cfg::STAGE_BYTES;
        unsigned char* Vs = Ks + 2 * Cfg::KV_PLANE;
        mbar_arrive_expect_tx(&kv_full[stage], Cfg::STAGE_BYTES);
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            tma_load_3d(Ks + pl * Cfg::KV_PLANE + ch * (AT_BK * 128), &mapK, &kv_full[stage], hk * D + ch * 64,
                        tok0 + j * AT_BK, pl);
            tma_load_3d(Vs + pl * Cfg::KV_PLANE + ch * (AT_BK * 128), &mapV, &kv_full[stage], hk * D + ch * 64,
                        tok0 + j * AT_BK, pl);
          }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(AT_BQ, AT_BK, 0, 0, 1);   // S = Q K^T: both operands K-major
      constexpr uint32_t idesc_o = umma_idesc_bf16(AT_BQ, D, 0, 1, 1);       // O = P V:   V is MN-major
      constexpr int PA[3] = {1, 0, 0};                                        // lo*hi, hi*lo, hi*hi (small terms first)
      constexpr int PB[3] = {0, 1, 0};
      const uint32_t t_s = tmem_base, t_o = tmem_base + 64;
      mbar_wait(&q_full, 0);
      for (int j = 0; j < nk; ++j) {
        const int stage = j & 1;
        mbar_wait(&kv_full[stage], (j >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t sQ = smem_u32(Qs);
        const uint32_t sK = smem_u32(KVs + stage * Cfg::STAGE_BYTES);
        const uint32_t sV = sK + 2 * Cfg::KV_PLANE;
        // S_j: the softmax warps finished reading S_{j-1} before they arrived on p_full_{j-1}
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            const uint64_t da = umma_desc_kmajor_sw128(sQ + PA[q] * Cfg::Q_PLANE + ch * (AT_BQ * 128));
            const uint64_t db = umma_desc_kmajor_sw128(sK + PB[q] * Cfg::KV_PLANE + ch * (AT_BK * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(t_s, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_s, (q | ch | k) != 0);
          }
        umma_commit(&s_full);
        // O_j = P_j V_j once the probabilities are in shared memory (O_{j-1} was drained before that arrival)
        mbar_wait(&p_full, j & 1);
        tc_fence_after_sync();
        const uint32_t sP = smem_u32(Ps);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint64_t da = umma_desc_kmajor_sw128(sP + PA[q] * Cfg::P_PLANE);
          const uint64_t db = umma_desc_mnmajor_sw128(sV + PB[q] * Cfg::KV_PLANE, AT_BK * 128, 1024);
#pragma unroll
          for (int k = 0; k < AT_BK / 16; ++k)
            umma_bf16(t_o, da + (uint64_t)(2 * k), db + (uint64_t)(128 * k), idesc_o, (q | k) != 0);
        }
        umma_commit(&o_full);
        umma_commit(&kv_empty[stage]);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + output (warps 2..9)
    constexpr int DH = D / 2;                          // output columns per thread
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                  // which 32 score columns / which half of the head dimension
    const int row = quad * 32 + lane;                  // row of the tile = TMEM lane
    const int qi = q0 + row;                           // query position in the sequence
    const uint32_t t_s = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * AT_HC);
    const uint32_t t_o = tmem_base + ((uint32_t)(quad * 32) << 16) + 64 + (uint32_t)(half * DH);
    const float c1 = p.inv_q[0] * p.inv_k[0] * p.scale_log2;
    const float c2 = p.inv_v[0] * AT_P_INV;
    float m = -INFINITY, l = 0.f;                      // l: this thread's share of the row sum
    float oacc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) oacc[i] = 0.f;
    for (int j = 0; j < nk; ++j) {
      mbar_wait(&s_full, j & 1);
      tc_fence_after_sync();
      float s[AT_HC];
      tmem_ld_32x32(t_s, s);
      tmem_ld_wait();
      const int key0 = j * AT_BK + half * AT_HC;
      float mx = -INFINITY;
      if (j * AT_BK + AT_BK - 1 <= q0) {                 // tile entirely left of the diagonal: nothing to mask
#pragma unroll
        for (int c = 0; c < AT_HC; ++c) { s[c] *= c1; mx = fmaxf(mx, s[c]); }
      } else {
#pragma unroll
        for (int c = 0; c < AT_HC; ++c) {
          s[c] = (key0 + c <= qi) ? s[c] * c1 : -INFINITY;
          mx = fmaxf(mx, s[c]);
        }
      }
      xch[half][row] = mx;
      asm volatile("bar.sync 1, %0;" ::"n"(AT_SM_THREADS) : "memory");
      const float m_new = fmaxf(m, fmaxf(mx, xch[half ^ 1][row]));   // finite from the first tile on
      asm volatile("bar.sync 1, %0;" ::"n"(AT_SM_THREADS) : "memory");   // (the slot is rewritten in the next step)
      const float corr = at_exp2(m - m_new);
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < AT_HC; ++c) {
        const float e = at_exp2(s[c] - m_new);
        lsum += e;
        s[c] = e * AT_P_SCALE;
      }
      l = l * corr + lsum;
      m = m_new;
      // P_j (x 2^14) as fp16 planes in the layout the MMA reads
      at_store_planes(Ps, Cfg::P_PLANE, row, half, s);
      fence_proxy_async();
      tc_fence_before_sync();
      mbar_arrive(&p_full);
      // O = O * corr + O_j
      mbar_wait(&o_full, j & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        float t[32];
        tmem_ld_32x32(t_o + (uint32_t)(c * 32), t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) oacc[c * 32 + i] = fmaf(oacc[c * 32 + i], corr, t[i] * c2);
      }
      tc_fence_before_sync();
    }
    // ---- row sum = the two halves' shares; normalise and store (P's shared memory is free now: one 32 x 32 float
    //      staging tile per warp)
    xch[half][row] = l;
    asm volatile("bar.sync 1, %0;" ::"n"(AT_SM_THREADS) : "memory");
    l += xch[half ^ 1][row];
    const float inv_l = 1.f / l;
    if (half == 0 && qi < p.S) p.lse2[((int64_t)b * p.H + h) * p.S + qi] = m + log2f(l);
    float* stg = reinterpret_cast<float*>(Ps) + (warp - 2) * 1024;
    const int sub = lane >> 3, pos = lane & 7;
    const int row_base = q0 + quad * 32;
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) =
            make_float4(oacc[c * 32 + 4 * q] * inv_l, oacc[c * 32 + 4 * q + 1] * inv_l, oacc[c * 32 + 4 * q + 2] * inv_l,
                        oacc[c * 32 + 4 * q + 3] * inv_l);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + sub;
        const int col = half * DH + c * 32 + ((pos ^ (r & 7)) << 2);
        const float4 o = *reinterpret_cast<const float4*>(stg + r * 32 + (pos << 2));
        if (row_base + r < p.S)
          *reinterpret_cast<float4*>(p.out + (int64_t)(tok0 + row_base + r) * p.ld_out + h * D + col) = o;
      }
      __syncwarp();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}


// =====================================================================================================
// Backward.  With P = softmax(scale * Q K^T) (causal), O = P V and delta_q = sum_d dO[q,d] O[q,d]:
//     dV = P^T dO,   dP = dO V^T,   dS = P o (dP - delta),   dQ = scale * dS K,   dK = scale * dS^T Q.
// Two kernels, no atomics and a fixed summation order:
//   attn_bwd_dq_kernel   one CTA per (batch, head, 128 queries), keys in steps of 64:
//                        S = Q K_j^T, dP = dO V_j^T -> dS planes [128 q x 64 keys] -> dQ += dS K_j   (K_j MN-major)
//   attn_bwd_dkv_kernel  one CTA per (batch, kv head, 128 keys), queries in steps of 64, all query heads of the
//                        group: S^T = K Q_i^T, dP^T = V dO_i^T -> P^T planes -> dV += P^T dO_i (dO_i MN-major);
//                        dS^T planes (same buffer) -> dK += dS^T Q_i (Q_i MN-major)
// P is recomputed from the forward's lse2 (p = exp2(s * c1 - lse2)).  dQ / dK / dV accumulate in TMEM over the
// steps of a CTA (chains of <= 96 MMAs for Hkv = H; the accumulator's truncation then costs <= ~2e-6 relative).
// Plane scales: P * 2^14; dS * kappa with kappa = 2^-17 / D on the raw (plane-unit) dP, which cannot overflow
// (|dP_raw| <= D * 2^30).
// =====================================================================================================
struct AttnBwdParams {
  int B, S, H, Hkv;
  float scale, scale_log2;
  const float* inv_q;
  const float* inv_k;
  const float* inv_v;
  const float* inv_do;
  const float* lse2;       // [B][H][S]
  const float* delta;      // [B][H][S]  sum_d dO O (true units)
  float* dq;               // [B*S][H*D]
  float* dk;               // [B*S][Hkv*D]
  float* dv;               // [B*S][Hkv*D]
};

template <int D>
struct AtBwdCfg {
  static constexpr int CH = D / 64;
  static constexpr int A_PLANE = CH * 128 * 128;       // one plane of a 128-row operand tile
  static constexpr int B_PLANE = CH * 64 * 128;        // one plane of a 64-row operand tile
  static constexpr int OFF_B = 4 * A_PLANE;            // two resident 128-row operands (2 planes each)
  static constexpr int STAGE_BYTES = 4 * B_PLANE;      // two streamed 64-row operands
  static constexpr int T_PLANE = 128 * 128;            // [128][64] fp16 plane of the P^T / dS tile
  static constexpr int OFF_T = OFF_B + STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_T + 2 * T_PLANE + 1024;
  static constexpr float KAPPA = (D == 128) ? 5.9604644775390625e-8f /*2^-24*/ : 1.1920928955078125e-7f /*2^-23*/;
};

// drain NC columns of a [128 lanes x .] fp32 accumulator from TMEM (taddr = first of them), scale and store them to
// rows [row_first, row_first + 128) x [col0, col0 + NC) of a row-major matrix; rows >= rows_valid are skipped.
// stg: this warp's 32 x 32 float staging tile.
template <int NC>
__device__ __forceinline__ void at_drain_store(uint32_t taddr, float factor, float* stg, float* dst, int64_t ld,
                                               int64_t row_first, int rows_valid, int col0, int quad, int lane) {
  const int sub = lane >> 3, pos = lane & 7;
#pragma unroll
  for (int c = 0; c < NC / 32; ++c) {
    float t[32];
    tmem_ld_32x32(taddr + (uint32_t)(c * 32), t);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) =
          make_float4(t[4 * q] * factor, t[4 * q + 1] * factor, t[4 * q + 2] * factor, t[4 * q + 3] * factor);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = quad * 32 + i * 4 + sub;
      const int col = col0 + c * 32 + ((pos ^ ((i * 4 + sub) & 7)) << 2);
      const float4 o = *reinterpret_cast<const float4*>(stg + (i * 4 + sub) * 32 + (pos << 2));
      if (r < rows_valid) *reinterpret_cast<float4*>(dst + (row_first + r) * ld + col) = o;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ dQ
template <int D>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapDO,
                   const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV,
                   const AttnBwdParams p) {
  using Cfg = AtBwdCfg<D>;
  constexpr int CH = Cfg::CH;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* Qs = smem;                            // A operands: Q_i then dO_i, 2 planes each
  unsigned char* DOs = smem + 2 * Cfg::A_PLANE;
  unsigned char* KVs = smem + Cfg::OFF_B;              // K_j (2 planes) then V_j (2 planes)
  unsigned char* Ts = smem + Cfg::OFF_T;               // dS planes
  __shared__ uint64_t a_full, kv_full, kv_empty, s_full, t_full;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (p.S + AT_BQ - 1) / AT_BQ;
  const int qt = nqt - 1 - (int)blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.H / p.Hkv);
  const int q0 = qt * AT_BQ;
  const int nk = (min(p.S, q0 + AT_BQ) + AT_BK - 1) / AT_BK;
  const int tok0 = b * p.S;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapQ); prefetch_tmap(&mapDO); prefetch_tmap(&mapK); prefetch_tmap(&mapV);
    mbar_init(&a_full, 1); mbar_init(&kv_full, 1); mbar_init(&kv_empty, 1);
    mbar_init(&s_full, 1); mbar_init(&t_full, AT_SM_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(&tmem_base_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&a_full, 4 * Cfg::A_PLANE);
#pragma unroll
      for (int pl = 0; pl < 2; ++pl)
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          tma_load_3d(Qs + pl * Cfg::A_PLANE + ch * (128 * 128), &mapQ, &a_full, h * D + ch * 64, tok0 + q0, pl);
          tma_load_3d(DOs + pl * Cfg::A_PLANE + ch * (128 * 128), &mapDO, &a_full, h * D + ch * 64, tok0 + q0, pl);
        }
      for (int j = 0; j < nk; ++j) {
        mbar_wait(&kv_empty, (j & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full, Cfg::STAGE_BYTES);
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            tma_load_3d(KVs + pl * Cfg::B_PLANE + ch * (64 * 128), &mapK, &kv_full, hk * D + ch * 64, tok0 + j * AT_BK, pl);
            tma_load_3d(KVs + (2 + pl) * Cfg::B_PLANE + ch * (64 * 128), &mapV, &kv_full, hk * D + ch * 64,
                        tok0 + j * AT_BK, pl);
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0, 1);
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, D, 0, 1, 1);
      constexpr int PA[3] = {1, 0, 0};
      constexpr int PB[3] = {0, 1, 0};
      const uint32_t t_s = tmem_base, t_dp = tmem_base + 64, t_dq = tmem_base + 128;
      const uint32_t sQ = smem_u32(Qs), sDO = smem_u32(DOs), sK = smem_u32(KVs), sV = sK + 2 * Cfg::B_PLANE, sT = smem_u32(Ts);
      mbar_wait(&a_full, 0);
      for (int j = 0; j < nk; ++j) {
        mbar_wait(&kv_full, j & 1);
        tc_fence_after_sync();
        // S = Q K_j^T and dP = dO V_j^T (the softmax warps are done with the previous pair: they arrived on t_full)
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          const uint32_t sa = which ? sDO : sQ, sb = which ? sV : sK, td = which ? t_dp : t_s;
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
              const uint64_t da = umma_desc_kmajor_sw128(sa + PA[q] * Cfg::A_PLANE + ch * (128 * 128));
              const uint64_t db = umma_desc_kmajor_sw128(sb + PB[q] * Cfg::B_PLANE + ch * (64 * 128));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_s, (q | ch | k) != 0);
            }
        }
        umma_commit(&s_full);
        mbar_wait(&t_full, j & 1);
        tc_fence_after_sync();
        // dQ += dS K_j
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint64_t da = umma_desc_kmajor_sw128(sT + PA[q] * Cfg::T_PLANE);
          const uint64_t db = umma_desc_mnmajor_sw128(sK + PB[q] * Cfg::B_PLANE, 64 * 128, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(t_dq, da + (uint64_t)(2 * k), db + (uint64_t)(128 * k), idesc_q, (j | q | k) != 0);
        }
        umma_commit(&kv_empty);                          // K_j / V_j and the dS tile are free once these retire
      }
      umma_commit(&s_full);                              // (nk-th phase) everything issued has completed
    }
  } else {
    constexpr int DH = D / 2;
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                  // 32 of the 64 keys of a step; half of the head dimension at the end
    const int row = quad * 32 + lane;
    const int qi = q0 + row;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float c1 = p.inv_q[0] * p.inv_k[0] * p.scale_log2;
    const float idv = p.inv_do[0] * p.inv_v[0];
    const bool live = qi < p.S;
    const float lse = live ? p.lse2[((int64_t)b * p.H + h) * p.S + qi] : 0.f;
    const float del_raw = live ? p.delta[((int64_t)b * p.H + h) * p.S + qi] / idv : 0.f;   // delta in plane units
    for (int j = 0; j < nk; ++j) {
      mbar_wait(&s_full, j & 1);
      tc_fence_after_sync();
      float s[AT_HC], dp[AT_HC];
      tmem_ld_32x32(t_lane + (uint32_t)(half * AT_HC), s);
      tmem_ld_32x32(t_lane + 64 + (uint32_t)(half * AT_HC), dp);
      tmem_ld_wait();
      const int key0 = j * AT_BK + half * AT_HC;
#pragma unroll
      for (int c = 0; c < AT_HC; ++c) {
        const float pr = (live && key0 + c <= qi) ? at_exp2(s[c] * c1 - lse) : 0.f;
        s[c] = pr * (dp[c] - del_raw) * Cfg::KAPPA;
      }
      // Ts is free here: the dQ product of the previous step completed before s_full of this step (same commit order)
      at_store_planes(Ts, Cfg::T_PLANE, row, half, s);
      fence_proxy_async();
      tc_fence_before_sync();
      mbar_arrive(&t_full);
    }
    // all MMAs done -> drain dQ
    mbar_wait(&s_full, nk & 1);
    tc_fence_after_sync();
    const float c_dq = p.scale * idv * p.inv_k[0] / Cfg::KAPPA;
    float* stg = reinterpret_cast<float*>(Ts) + (warp - 2) * 1024;
    at_drain_store<DH>(t_lane + 128 + (uint32_t)(half * DH), c_dq, stg, p.dq, (int64_t)p.H * D, (int64_t)tok0 + q0, p.S - q0,
                       h * D + half * DH, quad, lane);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<256>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ dK, dV
template <int D>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV,
                    const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapDO,
                    const AttnBwdParams p) {
  using Cfg = AtBwdCfg<D>;
  constexpr int CH = Cfg::CH;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* Ks = smem;                            // A operands: K_j then V_j (128 keys), 2 planes each
  unsigned char* Vs = smem + 2 * Cfg::A_PLANE;
  unsigned char* QDs = smem + Cfg::OFF_B;              // Q_i (2 planes) then dO_i (2 planes), 64 queries
  unsigned char* Ts = smem + Cfg::OFF_T;               // P^T planes, then dS^T planes
  __shared__ uint64_t a_full, qd_full, qd_empty, s_full, pt_full, dv_done, dst_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float col_lse[2][64], col_del[2][64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkt = (p.S + 127) / 128;
  const int kt = blockIdx.x;                           // key tiles: the first one has the most queries, it comes first
  const int hk = blockIdx.y, b = blockIdx.z;
  const int group = p.H / p.Hkv;
  const int k0 = kt * 128;
  const int tok0 = b * p.S;
  const int i_first = k0 / 64;                         // queries before the tile's first key see none of its keys
  const int nq = (p.S + 63) / 64;
  const int steps_per_head = nq - i_first;
  const int nsteps = group * steps_per_head;
  (void)nkt;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapK); prefetch_tmap(&mapV); prefetch_tmap(&mapQ); prefetch_tmap(&mapDO);
    mbar_init(&a_full, 1); mbar_init(&qd_full, 1); mbar_init(&qd_empty, 1);
    mbar_init(&s_full, 1); mbar_init(&dv_done, 1); mbar_init(&pt_full, AT_SM_THREADS); mbar_init(&dst_full, AT_SM_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&a_full, 4 * Cfg::A_PLANE);
#pragma unroll
      for (int pl = 0; pl < 2; ++pl)
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          tma_load_3d(Ks + pl * Cfg::A_PLANE + ch * (128 * 128), &mapK, &a_full, hk * D + ch * 64, tok0 + k0, pl);
          tma_load_3d(Vs + pl * Cfg::A_PLANE + ch * (128 * 128), &mapV, &a_full, hk * D + ch * 64, tok0 + k0, pl);
        }
      for (int t = 0; t < nsteps; ++t) {
        const int h = hk * group + t / steps_per_head;
        const int i = i_first + t % steps_per_head;
        mbar_wait(&qd_empty, (t & 1) ^ 1);
        mbar_arrive_expect_tx(&qd_full, Cfg::STAGE_BYTES);
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
          for (int ch = 0; ch < CH; ++ch) {
            tma_load_3d(QDs + pl * Cfg::B_PLANE + ch * (64 * 128), &mapQ, &qd_full, h * D + ch * 64, tok0 + i * 64, pl);
            tma_load_3d(QDs + (2 + pl) * Cfg::B_PLANE + ch * (64 * 128), &mapDO, &qd_full, h * D + ch * 64, tok0 + i * 64, pl);
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0, 1);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, D, 0, 1, 1);
      constexpr int PA[3] = {1, 0, 0};
      constexpr int PB[3] = {0, 1, 0};
      const uint32_t t_st = tmem_base, t_dpt = tmem_base + 64, t_dv = tmem_base + 128, t_dk = tmem_base + 128 + D;
      const uint32_t sK = smem_u32(Ks), sV = smem_u32(Vs), sQ = smem_u32(QDs), sDO = sQ + 2 * Cfg::B_PLANE, sT = smem_u32(Ts);
      mbar_wait(&a_full, 0);
      for (int t = 0; t < nsteps; ++t) {
        mbar_wait(&qd_full, t & 1);
        tc_fence_after_sync();
        // S^T = K Q_i^T, dP^T = V dO_i^T  (the softmax warps have consumed the previous pair: dst_full(t-1) was waited)
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          const uint32_t sa = which ? sV : sK, sb = which ? sDO : sQ, td = which ? t_dpt : t_st;
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
              const uint64_t da = umma_desc_kmajor_sw128(sa + PA[q] * Cfg::A_PLANE + ch * (128 * 128));
              const uint64_t db = umma_desc_kmajor_sw128(sb + PB[q] * Cfg::B_PLANE + ch * (64 * 128));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_s, (q | ch | k) != 0);
            }
        }
        umma_commit(&s_full);
        // dV += P^T dO_i
        mbar_wait(&pt_full, t & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint64_t da = umma_desc_kmajor_sw128(sT + PA[q] * Cfg::T_PLANE);
          const uint64_t db = umma_desc_mnmajor_sw128(sDO + PB[q] * Cfg::B_PLANE, 64 * 128, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(t_dv, da + (uint64_t)(2 * k), db + (uint64_t)(128 * k), idesc_g, (t | q | k) != 0);
        }
        umma_commit(&dv_done);
        // dK += dS^T Q_i
        mbar_wait(&dst_full, t & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const uint64_t da = umma_desc_kmajor_sw128(sT + PA[q] * Cfg::T_PLANE);
          const uint64_t db = umma_desc_mnmajor_sw128(sQ + PB[q] * Cfg::B_PLANE, 64 * 128, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(t_dk, da + (uint64_t)(2 * k), db + (uint64_t)(128 * k), idesc_g, (t | q | k) != 0);
        }
        umma_commit(&qd_empty);
      }
      umma_commit(&s_full);                              // phase nsteps: all accumulations have completed
    }
  } else {
    constexpr int DH = D / 2;
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                    // 32 of the 64 queries of a step; half of the head dimension at the end
    const int row = quad * 32 + lane;
    const int key = k0 + row;
    const int stid = threadIdx.x - 64;                   // 0..255
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const float c1 = p.inv_q[0] * p.inv_k[0] * p.scale_log2;
    const float idv = p.inv_do[0] * p.inv_v[0];
    for (int t = 0; t < nsteps; ++t) {
      const int h = hk * group + t / steps_per_head;
      const int i = i_first + t % steps_per_head;
      // lse2 / delta of the 64 queries of this step (columns of the transposed tiles)
      if (stid < 128) {
        const int c = stid & 63;
        const int qi = i * 64 + c;
        const float v = (qi < p.S) ? (stid < 64 ? p.lse2 : p.delta)[((int64_t)b * p.H + h) * p.S + qi] : 0.f;
        if (stid < 64) col_lse[t & 1][c] = v; else col_del[t & 1][c] = v / idv;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(AT_SM_THREADS) : "memory");
      mbar_wait(&s_full, t & 1);
      tc_fence_after_sync();
      float s[AT_HC], dp[AT_HC];
      tmem_ld_32x32(t_lane + (uint32_t)(half * AT_HC), s);
      tmem_ld_32x32(t_lane + 64 + (uint32_t)(half * AT_HC), dp);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < AT_HC; ++c) {
        const int cc = half * AT_HC + c;
        const int qi = i * 64 + cc;
        s[c] = (qi < p.S && key <= qi && key < p.S) ? at_exp2(s[c] * c1 - col_lse[t & 1][cc]) : 0.f;   // P^T
      }
      // the buffer's previous content (dS^T of step t-1) was consumed: qd_empty(t-1) preceded qd_full(t) <= s_full(t)
      {
        float ps[AT_HC];
#pragma unroll
        for (int c = 0; c < AT_HC; ++c) ps[c] = s[c] * AT_P_SCALE;
        at_store_planes(Ts, Cfg::T_PLANE, row, half, ps);
      }
      fence_proxy_async();
      tc_fence_before_sync();
      mbar_arrive(&pt_full);
      // dS^T = P^T o (dP^T - delta)
#pragma unroll
      for (int c = 0; c < AT_HC; ++c) s[c] = s[c] * (dp[c] - col_del[t & 1][half * AT_HC + c]) * Cfg::KAPPA;
      mbar_wait(&dv_done, t & 1);                        // the dV product has finished reading P^T
      at_store_planes(Ts, Cfg::T_PLANE, row, half, s);
      fence_proxy_async();
      tc_fence_before_sync();
      mbar_arrive(&dst_full);
    }
    mbar_wait(&s_full, nsteps & 1);
    tc_fence_after_sync();
    const float c_dv = p.inv_do[0] * AT_P_INV;
    const float c_dk = p.scale * idv * p.inv_q[0] / Cfg::KAPPA;
    float* stg = reinterpret_cast<float*>(Ts) + (warp - 2) * 1024;
    at_drain_store<DH>(t_lane + 128 + (uint32_t)(half * DH), c_dv, stg, p.dv, (int64_t)p.Hkv * D, (int64_t)tok0 + k0, p.S - k0,
                       hk * D + half * DH, quad, lane);
    at_drain_store<DH>(t_lane + 128 + D + (uint32_t)(half * DH), c_dk, stg, p.dk, (int64_t)p.Hkv * D, (int64_t)tok0 + k0,
                       p.S - k0, hk * D + half * DH, quad, lane);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// delta[b][h][q] = sum_d dO[q, h*D + d] * O[q, h*D + d]; one warp per (token, head)
__global__ void attn_delta_kernel(const float* __restrict__ dO, const float* __restrict__ O, int B, int S, int H, int D,
                                  float* __restrict__ delta) {
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)B * S * H) return;
  const int h = (int)(w % H);
  const int64_t tok = w / H;
  const float4* a = reinterpret_cast<const float4*>(dO + tok * (int64_t)H * D + (int64_t)h * D);
  const float4* o = reinterpret_cast<const float4*>(O + tok * (int64_t)H * D + (int64_t)h * D);
  float acc = 0.f;
  for (int i = threadIdx.x & 31; i < D / 4; i += 32) {
    const float4 x = a[i], y = o[i];
    acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) {
    const int b = (int)(tok / S), q = (int)(tok % S);
    delta[((int64_t)b * H + h) * S + q] = acc;
  }
}

// ------------------------------------------------------------------------------------------ operand preparation
// q, k, v leave their projections as fp32 [tokens, heads * D]; the attention kernels want RoPE applied to q and k
// and all three as tensor-scaled fp16 planes.  Done separately that is two RoPE passes (read + write each) and three
// split calls of two passes each (maximum, then planes).  Here: ONE pass for the three maxima (a rotation grows a
// pair by at most sqrt(2), so the scale of the rotated tensor follows from the maximum before the rotation with
// half a bit of headroom) and ONE pass that rotates, scales, splits and writes the planes; the rotated fp32 q / k
// are never written -- nothing else reads them.
__global__ void __launch_bounds__(256)
attn_absmax3_kernel(const float4* __restrict__ q, int64_t nq4, const float4* __restrict__ k, int64_t nk4,
                    const float4* __restrict__ v, int64_t nv4, uint32_t* __restrict__ ws) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float m[3] = {0.f, 0.f, 0.f};
  const float4* src[3] = {q, k, v};
  const int64_t n4[3] = {nq4, nk4, nv4};
#pragma unroll
  for (int a = 0; a < 3; ++a)
    for (int64_t i = tid; i < n4[a]; i += stride) {
      const float4 x = ldg_stream(src[a] + i);
      const float y = fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w)));
      if (y == y) m[a] = fmaxf(m[a], y);
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float w = warp_max(m[a]);
    if ((threadIdx.x & 31) == 0 && w > 0.f) atomicMax(ws + a, __float_as_uint(w));
  }
}

struct PrepTensor {
  const float* src;       // [tokens][heads * D]
  uint16_t* planes;       // [2][tokens][pitch]
  float* inv;             // [n_inv]
  int heads, pitch, n_inv, rope;
  int64_t items_per_token;   // rope: heads * D / 16 pair groups of 8; plain: heads * D / 8
};

__global__ void __launch_bounds__(256)
attn_rope_split_kernel(PrepTensor tq, PrepTensor tk, PrepTensor tv, int64_t tokens, int seq, int D,
                       const float* __restrict__ cosp, const float* __restrict__ sinp, int64_t cs_batch,
                       const uint32_t* __restrict__ ws) {
  const PrepTensor T[3] = {tq, tk, tv};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int half = D >> 1;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const PrepTensor& P = T[a];
    float sc, inv;
    scale_from_max(__uint_as_float(ws[a]) * (P.rope ? 1.5f : 1.f), sc, inv);
    for (int64_t i = tid; i < P.n_inv; i += stride) P.inv[i] = inv;
    const int64_t total = tokens * P.items_per_token;
    const int64_t plane = tokens * (int64_t)P.pitch;
    const int cols = P.heads * D;
    for (int64_t i = tid; i < total; i += stride) {
      const int64_t t = i / P.items_per_token;
      const int rem = (int)(i - t * P.items_per_token);
      if (!P.rope) {
        const int c = rem * 8;
        const float4 a0 = ldg_stream(reinterpret_cast<const float4*>(P.src + t * cols + c));
        const float4 a1 = ldg_stream(reinterpret_cast<const float4*>(P.src + t * cols + c + 4));
        const float x[8] = {a0.x * sc, a0.y * sc, a0.z * sc, a0.w * sc, a1.x * sc, a1.y * sc, a1.z * sc, a1.w * sc};
        uint4 hi, lo;
        at_split8(x, hi, lo);
        *reinterpret_cast<uint4*>(P.planes + t * P.pitch + c) = hi;
        *reinterpret_cast<uint4*>(P.planes + plane + t * P.pitch + c) = lo;
        continue;
      }
      const int groups = half >> 3;                      // groups of 8 pairs per head
      const int h = rem / groups, j = (rem - h * groups) * 8;
      const int64_t b = t / seq;
      const int s = (int)(t - b * seq);
      const float* cr = cosp + b * cs_batch + (int64_t)s * D;
      const float* sr = sinp + b * cs_batch + (int64_t)s * D;
      const float* p1 = P.src + t * cols + h * D + j;
      float x1[8], x2[8], c1[8], c2[8], s1[8], s2[8], o1[8], o2[8];
#pragma unroll
      for (int e = 0; e < 8; e += 4) {
        *reinterpret_cast<float4*>(x1 + e) = ldg_stream(reinterpret_cast<const float4*>(p1 + e));
        *reinterpret_cast<float4*>(x2 + e) = ldg_stream(reinterpret_cast<const float4*>(p1 + half + e));
        *reinterpret_cast<float4*>(c1 + e) = __ldg(reinterpret_cast<const float4*>(cr + j + e));
        *reinterpret_cast<float4*>(c2 + e) = __ldg(reinterpret_cast<const float4*>(cr + half + j + e));
        *reinterpret_cast<float4*>(s1 + e) = __ldg(reinterpret_cast<const float4*>(sr + j + e));
        *reinterpret_cast<float4*>(s2 + e) = __ldg(reinterpret_cast<const float4*>(sr + half + j + e));
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        o1[e] = (x1[e] * c1[e] - x2[e] * s1[e]) * sc;   // same expression as rope_kernel (layer_ops.cu), then the scale
        o2[e] = (x2[e] * c2[e] + x1[e] * s2[e]) * sc;
      }
      uint4 hi, lo;
      uint16_t* d = P.planes + t * P.pitch + h * D + j;
      at_split8(o1, hi, lo);
      *reinterpret_cast<uint4*>(d) = hi;
      *reinterpret_cast<uint4*>(d + plane) = lo;
      at_split8(o2, hi, lo);
      *reinterpret_cast<uint4*>(d + half) = hi;
      *reinterpret_cast<uint4*>(d + half + plane) = lo;
    }
  }
}

// ------------------------------------------------------------------------------------------ host
int tc_make_plane_map(CUtensorMap* map, const void* planes, int NS, int R, int K, int Kp, int box_rows, int box_inner);

static int attn_check(int B, int S, int H, int Hkv, int D) {
  if (B <= 0 || S <= 0 || H <= 0 || Hkv <= 0 || H % Hkv != 0) return bad_arg("attention: B/S/H/Hkv");
  if (D != 64 && D != 128) return bad_arg("attention: head_dim must be 64 or 128");
  if ((int64_t)B * S >= (1 << 30)) return bad_arg("attention: too many tokens");
  return 0;
}

template <int D>
static int attn_fwd_launch(const void* qp, const void* kp, const void* vp, AttnParams prm, void* stream) {
  using Cfg = AtFwdCfg<D>;
  const int tokens = prm.B * prm.S;
  CUtensorMap mq, mk, mv;
  int rc = tc_make_plane_map(&mq, qp, 2, tokens, prm.H * D, (int)plane_pitch(prm.H * D), AT_BQ, 64); if (rc) return rc;
  rc = tc_make_plane_map(&mk, kp, 2, tokens, prm.Hkv * D, (int)plane_pitch(prm.Hkv * D), AT_BK, 64); if (rc) return rc;
  rc = tc_make_plane_map(&mv, vp, 2, tokens, prm.Hkv * D, (int)plane_pitch(prm.Hkv * D), AT_BK, 64); if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    rc = check_cuda(cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES),
                    "attn_fwd attr");
    if (rc) return rc;
    attr = true;
  }
  dim3 grid((unsigned)ceil_div(prm.S, AT_BQ), (unsigned)prm.H, (unsigned)prm.B);
  GRASP_LAUNCH((attn_fwd_kernel<D>), grid, dim3(AT_THREADS), Cfg::SMEM_BYTES, stream, mq, mk, mv, prm);
  GRASP_CHECK_LAST("attn_fwd_kernel");
  return 0;
}

template <int D>
static int attn_bwd_launch(const void* qp, const void* kp, const void* vp, const void* dop, const float* dO, const float* O,
                           AttnBwdParams prm, float* delta, void* stream) {
  using Cfg = AtBwdCfg<D>;
  const int tokens = prm.B * prm.S;
  const int64_t nwarps = (int64_t)tokens * prm.H;
  GRASP_LAUNCH(attn_delta_kernel, dim3((unsigned)ceil_div(nwarps, 8)), dim3(256), 0, stream, dO, O, prm.B, prm.S, prm.H, D, delta);
  prm.delta = delta;
  CUtensorMap mq128, mdo128, mk64, mv64, mk128, mv128, mq64, mdo64;
  const int pq = (int)plane_pitch(prm.H * D), pk = (int)plane_pitch(prm.Hkv * D);
  int rc = tc_make_plane_map(&mq128, qp, 2, tokens, prm.H * D, pq, 128, 64); if (rc) return rc;
  rc = tc_make_plane_map(&mdo128, dop, 2, tokens, prm.H * D, pq, 128, 64); if (rc) return rc;
  rc = tc_make_plane_map(&mq64, qp, 2, tokens, prm.H * D, pq, 64, 64); if (rc) return rc;
  rc = tc_make_plane_map(&mdo64, dop, 2, tokens, prm.H * D, pq, 64, 64); if (rc) return rc;
  rc = tc_make_plane_map(&mk128, kp, 2, tokens, prm.Hkv * D, pk, 128, 64); if (rc) return rc;
  rc = tc_make_plane_map(&mv128, vp, 2, tokens, prm.Hkv * D, pk, 128, 64); if (rc) return rc;
  rc = tc_make_plane_map(&mk64, kp, 2, tokens, prm.Hkv * D, pk, 64, 64); if (rc) return rc;
  rc = tc_make_plane_map(&mv64, vp, 2, tokens, prm.Hkv * D, pk, 64, 64); if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    rc = check_cuda(cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES),
                    "attn_bwd_dq attr");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(attn_bwd_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES),
                    "attn_bwd_dkv attr");
    if (rc) return rc;
    attr = true;
  }
  dim3 gq((unsigned)ceil_div(prm.S, 128), (unsigned)prm.H, (unsigned)prm.B);
  GRASP_LAUNCH((attn_bwd_dq_kernel<D>), gq, dim3(AT_THREADS), Cfg::SMEM_BYTES, stream, mq128, mdo128, mk64, mv64, prm);
  dim3 gk((unsigned)ceil_div(prm.S, 128), (unsigned)prm.Hkv, (unsigned)prm.B);
  GRASP_LAUNCH((attn_bwd_dkv_kernel<D>), gk, dim3(AT_THREADS), Cfg::SMEM_BYTES, stream, mk128, mv128, mq64, mdo64, prm);
  GRASP_CHECK_LAST("attn_bwd kernels");
  return 0;
}

}  // namespace grasp

using namespace grasp;

extern "C" int grasp_attn_fwd(const void* q_planes, const float* inv_q, const void* k_planes, const float* inv_k,
                              const void* v_planes, const float* inv_v, int B, int S, int H, int Hkv, int D, float scale,
                              float* out, float* lse2, void* stream) {
  if (!q_planes || !k_planes || !v_planes || !inv_q || !inv_k || !inv_v || !out || !lse2) return bad_arg("attn_fwd: null");
  int rc = attn_check(B, S, H, Hkv, D);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(q_planes) | reinterpret_cast<uintptr_t>(k_planes) | reinterpret_cast<uintptr_t>(v_planes)) & 1023)
    return bad_arg("attn_fwd: planes must be 1024-byte aligned");
  if (reinterpret_cast<uintptr_t>(out) & 15) return bad_arg("attn_fwd: out must be 16-byte aligned");
  AttnParams prm{};
  prm.B = B; prm.S = S; prm.H = H; prm.Hkv = Hkv;
  prm.scale_log2 = scale * 1.4426950408889634f;
  prm.inv_q = inv_q; prm.inv_k = inv_k; prm.inv_v = inv_v;
  prm.out = out; prm.ld_out = (int64_t)H * D; prm.lse2 = lse2;
  return D == 128 ? attn_fwd_launch<128>(q_planes, k_planes, v_planes, prm, stream)
                  : attn_fwd_launch<64>(q_planes, k_planes, v_planes, prm, stream);
}

extern "C" int grasp_attn_bwd(const void* q_planes, const float* inv_q, const void* k_planes, const float* inv_k,
                              const void* v_planes, const float* inv_v, const void* do_planes, const float* inv_do,
                              const float* dO, const float* O, const float* lse2, int B, int S, int H, int Hkv, int D,
                              float scale, float* dq, float* dk, float* dv, float* delta_ws, void* stream) {
  if (!q_planes || !k_planes || !v_planes || !do_planes || !inv_q || !inv_k || !inv_v || !inv_do || !dO || !O || !lse2 ||
      !dq || !dk || !dv || !delta_ws)
    return bad_arg("attn_bwd: null");
  int rc = attn_check(B, S, H, Hkv, D);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(q_planes) | reinterpret_cast<uintptr_t>(k_planes) | reinterpret_cast<uintptr_t>(v_planes) |
       reinterpret_cast<uintptr_t>(do_planes)) & 1023)
    return bad_arg("attn_bwd: planes must be 1024-byte aligned");
  if ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv) |
       reinterpret_cast<uintptr_t>(dO) | reinterpret_cast<uintptr_t>(O)) & 15)
    return bad_arg("attn_bwd: tensors must be 16-byte aligned");
  AttnBwdParams prm{};
  prm.B = B; prm.S = S; prm.H = H; prm.Hkv = Hkv;
  prm.scale = scale; prm.scale_log2 = scale * 1.4426950408889634f;
  prm.inv_q = inv_q; prm.inv_k = inv_k; prm.inv_v = inv_v; prm.inv_do = inv_do;
  prm.lse2 = lse2; prm.dq = dq; prm.dk = dk; prm.dv = dv;
  return D == 128 ? attn_bwd_launch<128>(q_planes, k_planes, v_planes, do_planes, dO, O, prm, delta_ws, stream)
                  : attn_bwd_launch<64>(q_planes, k_planes, v_planes, do_planes, dO, O, prm, delta_ws, stream);
}

extern "C" int grasp_attn_prep_qkv(const float* q, const float* k, const float* v, int64_t tokens, int64_t seq, int H,
                                   int Hkv, int D, const float* cosp, const float* sinp, int64_t cs_batch, void* q_planes,
                                   float* q_inv, void* k_planes, float* k_inv, void* v_planes, float* v_inv, void* ws,
                                   void* stream) {
  if (!q || !k || !v || !cosp || !sinp || !q_planes || !k_planes || !v_planes || !q_inv || !k_inv || !v_inv || !ws)
    return bad_arg("attn_prep: null");
  if (tokens <= 0 || seq <= 0 || tokens % seq || H <= 0 || Hkv <= 0 || (D != 64 && D != 128))
    return bad_arg("attn_prep: tokens/seq/heads/head_dim");
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(cosp) | reinterpret_cast<uintptr_t>(sinp)) & 15 || (cs_batch & 3))
    return bad_arg("attn_prep: tensors must be 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(q_planes) | reinterpret_cast<uintptr_t>(k_planes) | reinterpret_cast<uintptr_t>(v_planes)) & 1023)
    return bad_arg("attn_prep: planes must be 1024-byte aligned");
  int rc = check_cuda(cudaMemsetAsync(ws, 0, 12, (cudaStream_t)stream), "attn_prep memset");
  if (rc) return rc;
  const int64_t nq4 = tokens * H * D / 4, nk4 = tokens * Hkv * D / 4;
  const int grid = sm_count() * 8;
  GRASP_LAUNCH(attn_absmax3_kernel, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const float4*>(q), nq4,
               reinterpret_cast<const float4*>(k), nk4, reinterpret_cast<const float4*>(v), nk4, static_cast<uint32_t*>(ws));
  auto make = [&](const float* src, void* planes, float* inv, int heads, int rope) {
    PrepTensor t{};
    t.src = src; t.planes = static_cast<uint16_t*>(planes); t.inv = inv; t.heads = heads;
    t.pitch = (int)plane_pitch((int64_t)heads * D);
    t.n_inv = (int)(tokens > (int64_t)heads * D ? tokens : (int64_t)heads * D);
    t.rope = rope;
    t.items_per_token = rope ? (int64_t)heads * D / 16 : (int64_t)heads * D / 8;
    return t;
  };
  GRASP_LAUNCH(attn_rope_split_kernel, dim3(grid), dim3(256), 0, stream, make(q, q_planes, q_inv, H, 1),
               make(k, k_planes, k_inv, Hkv, 1), make(v, v_planes, v_inv, Hkv, 0), tokens, (int)seq, D, cosp, sinp, cs_batch,
               static_cast<const uint32_t*>(ws));
  GRASP_CHECK_LAST("attn_prep kernels");
  return 0;
}
