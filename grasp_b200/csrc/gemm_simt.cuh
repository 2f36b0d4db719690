// Plain fp32 FMA GEMM on CUDA cores with arbitrary element strides.
// This is the validation arithmetic (GRASP_PREC_SIMT) every tensor-core path is
// checked against on the device; it is not the fast path.
#pragma once
#include "common.cuh"

namespace grasp {

struct GemmStrided {
  const float* A; int64_t sAm, sAk;   // A(m,k) = A[m*sAm + k*sAk]
  const float* B; int64_t sBk, sBn;   // B(k,n) = B[k*sBk + n*sBn]
  void* C; int64_t ldc;               // C row-major, fp32 or bf16 (c_bf16)
  int c_bf16;
  int64_t M, N, K;
  float alpha, beta;
  int64_t bsA, bsB, bsC;              // batch strides (blockIdx.z)
};

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

static __global__ void __launch_bounds__(SG_THREADS)
gemm_simt_kernel(GemmStrided g) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const float* __restrict__ A = g.A + (int64_t)blockIdx.z * g.bsA;
  const float* __restrict__ B = g.B + (int64_t)blockIdx.z * g.bsB;
  float* __restrict__ C = static_cast<float*>(g.C) + (int64_t)blockIdx.z * g.bsC;
  __nv_bfloat16* __restrict__ Cb = static_cast<__nv_bfloat16*>(g.C) + (int64_t)blockIdx.z * g.bsC;
  const int64_t m0 = (int64_t)blockIdx.y * SG_BM, n0 = (int64_t)blockIdx.x * SG_BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4x4 outputs each
  float acc[4][4] = {};

  // loader mapping: fast index along the unit-stride dimension of each operand
  const bool a_kfast = (g.sAk == 1);
  const bool b_nfast = (g.sBn == 1);
  for (int64_t k0 = 0; k0 < g.K; k0 += SG_BK) {
#pragma unroll
    for (int it = 0; it < (SG_BM * SG_BK) / SG_THREADS; ++it) {
      const int e = tid + it * SG_THREADS;
      const int kk = a_kfast ? (e % SG_BK) : (e / SG_BM);
      const int mm = a_kfast ? (e / SG_BK) : (e % SG_BM);
      const int64_t m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < g.M && k < g.K) ? A[m * g.sAm + k * g.sAk] : 0.f;
    }
#pragma unroll
    for (int it = 0; it < (SG_BN * SG_BK) / SG_THREADS; ++it) {
      const int e = tid + it * SG_THREADS;
      const int kk = b_nfast ? (e / SG_BN) : (e % SG_BK);
      const int nn = b_nfast ? (e % SG_BN) : (e / SG_BK);
      const int64_t n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < g.N && k < g.K) ? B[k * g.sBk + n * g.sBn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.c_bf16) {
        if (g.beta != 0.f) v += g.beta * __bfloat162float(Cb[m * g.ldc + n]);
        Cb[m * g.ldc + n] = __float2bfloat16_rn(v);
      } else {
        if (g.beta != 0.f) v += g.beta * C[m * g.ldc + n];
        C[m * g.ldc + n] = v;
      }
    }
  }
}

static inline int launch_gemm_simt(const GemmStrided& g, int batch, void* stream) {
  if (g.M <= 0 || g.N <= 0) return 0;
  dim3 grid((unsigned)ceil_div(g.N, SG_BN), (unsigned)ceil_div(g.M, SG_BM), (unsigned)batch);
  GRASP_LAUNCH(gemm_simt_kernel, grid, dim3(SG_THREADS), 0, stream, g);
  GRASP_CHECK_LAST("gemm_simt_kernel");
  return 0;
}

}  // namespace grasp
