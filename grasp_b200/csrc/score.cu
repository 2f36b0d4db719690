// Singular-value gradient and importance score:
//   dsigma[i] = u_i^T G v_i = sum_a U[a,i] * (G Vh^T)[a,i],   score = |dsigma| or |dsigma*S|
// Replaces the autograd path through GRASPLayer (reference modeling_grasp.py:75-79,
// :354-363) and the score of :392-395.  The product Z = G Vh^T is never written:
// each tile of Z is contracted with the matching tile of U in the epilogue.
// This file holds the CUDA-core (GRASP_PREC_SIMT) arithmetic and the dispatch;
// the tcgen05 version of the same contraction lives in gemm_tc.cu.
#include "common.cuh"

namespace grasp {

// from gemm_tc.cu
int tc_sigma_partials(const float* U, const float* G, const float* Vh, int64_t out, int64_t in, int64_t r,
                      int prec, float* partial, int64_t* n_partials, void* ws, size_t ws_bytes, void* stream);
size_t tc_sigma_workspace_bytes(int64_t out, int64_t in, int64_t r, int prec);

constexpr int SC_T = 64, SC_BK = 32, SC_THREADS = 256;

// grid (r/64, out/64): partial[atile][i] = sum_{a in tile} U[a,i] * sum_b G[a,b] Vh[i,b]
__global__ void __launch_bounds__(SC_THREADS)
sigma_partial_simt_kernel(const float* __restrict__ U, const float* __restrict__ G, const float* __restrict__ Vh,
                          int64_t out, int64_t in, int64_t r, float* __restrict__ partial) {
  __shared__ float As[SC_BK][SC_T + 1];   // G tile   [k][a]
  __shared__ float Bs[SC_BK][SC_T + 1];   // Vh tile  [k][i]
  __shared__ float red[16][SC_T];
  const int64_t i0 = (int64_t)blockIdx.x * SC_T, a0 = (int64_t)blockIdx.y * SC_T;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < in; k0 += SC_BK) {
#pragma unroll
    for (int it = 0; it < (SC_T * SC_BK) / SC_THREADS; ++it) {
      const int e = tid + it * SC_THREADS;
      const int rr = e >> 5, kk = e & 31;
      const int64_t k = k0 + kk;
      As[kk][rr] = (a0 + rr < out && k < in) ? G[(a0 + rr) * in + k] : 0.f;
      Bs[kk][rr] = (i0 + rr < r && k < in) ? Vh[(i0 + rr) * in + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SC_BK; ++kk) {
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
  // epilogue: multiply by U[a,i] and reduce over the tile's rows a
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t i = i0 + tx * 4 + j;
    float s = 0.f;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int64_t a = a0 + ty * 4 + ii;
      if (a < out && i < r) s = fmaf(U[a * r + i], acc[ii][j], s);
    }
    red[ty][tx * 4 + j] = s;
  }
  __syncthreads();
  if (tid < SC_T) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 16; ++y) s += red[y][tid];
    if (i0 + tid < r) partial[(int64_t)blockIdx.y * r + i0 + tid] = s;
  }
}

// dsigma[i] (+)= sum_t partial[t][i]; optional score
__global__ void sigma_reduce_kernel(const float* __restrict__ partial, int64_t n_partials, int64_t r,
                                    const float* __restrict__ S, int metric, int accumulate,
                                    float* __restrict__ dsigma, float* __restrict__ score) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r) return;
  float s = 0.f;
  for (int64_t t = 0; t < n_partials; ++t) s += partial[t * r + i];
  if (accumulate) s += dsigma[i];
  dsigma[i] = s;
  if (score) score[i] = (metric == GRASP_METRIC_TAYLOR) ? fabsf(s * S[i]) : fabsf(s);
}

__global__ void score_from_grad_kernel(const float* __restrict__ g, const float* __restrict__ S, int64_t r,
                                       int metric, float* __restrict__ score) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r) return;
  score[i] = (metric == GRASP_METRIC_TAYLOR) ? fabsf(g[i] * S[i]) : fabsf(g[i]);
}

}  // namespace grasp

using namespace grasp;

extern "C" size_t grasp_sigma_score_workspace_bytes(int64_t out, int64_t in, int64_t r, int prec) {
  if (out <= 0 || in <= 0 || r <= 0) return 0;
  if (prec == GRASP_PREC_SIMT) return (size_t)ceil_div(out, SC_T) * r * 4;
  return tc_sigma_workspace_bytes(out, in, r, prec);
}

extern "C" int grasp_sigma_score(const float* U, const float* G, const float* Vh, const float* S, int64_t out,
                                 int64_t in, int64_t r, int metric, int accumulate, float* dsigma, float* score,
                                 int prec, void* ws, size_t ws_bytes, void* stream) {
  if (!U || !G || !Vh || !dsigma || !ws) return bad_arg("sigma_score: null");
  if (out <= 0 || in <= 0 || r <= 0 || r > (out < in ? out : in)) return bad_arg("sigma_score: out/in/r");
  if (metric != GRASP_METRIC_GRADIENT && metric != GRASP_METRIC_TAYLOR) return bad_arg("sigma_score: metric");
  if (score && metric == GRASP_METRIC_TAYLOR && !S) return bad_arg("sigma_score: taylor needs S");
  if (ws_bytes < grasp_sigma_score_workspace_bytes(out, in, r, prec)) return bad_arg("sigma_score: workspace too small");
  float* partial = static_cast<float*>(ws);
  int64_t n_partials = 0;
  if (prec == GRASP_PREC_SIMT) {
    n_partials = ceil_div(out, SC_T);
    dim3 grid((unsigned)ceil_div(r, SC_T), (unsigned)n_partials);
    GRASP_LAUNCH(sigma_partial_simt_kernel, grid, dim3(SC_THREADS), 0, stream, U, G, Vh, out, in, r, partial);
    GRASP_CHECK_LAST("sigma_partial_simt_kernel");
  } else if (prec == GRASP_PREC_BF16X3 || prec == GRASP_PREC_BF16X6 || prec == GRASP_PREC_F16X3) {
    int rc = tc_sigma_partials(U, G, Vh, out, in, r, prec, partial, &n_partials, ws, ws_bytes, stream);
    if (rc) return rc;
  } else {
    return bad_arg("sigma_score: prec");
  }
  GRASP_LAUNCH(sigma_reduce_kernel, dim3((unsigned)ceil_div(r, 256)), dim3(256), 0, stream, partial, n_partials, r,
               S, metric, accumulate, dsigma, score);
  GRASP_CHECK_LAST("sigma_reduce_kernel");
  return 0;
}

extern "C" int grasp_score_from_grad(const float* dsigma, const float* S, int64_t r, int metric, float* score,
                                     void* stream) {
  if (!dsigma || !score) return bad_arg("score_from_grad: null");
  if (r <= 0) return bad_arg("score_from_grad: r");
  if (metric != GRASP_METRIC_GRADIENT && metric != GRASP_METRIC_TAYLOR) return bad_arg("score_from_grad: metric");
  if (metric == GRASP_METRIC_TAYLOR && !S) return bad_arg("score_from_grad: taylor needs S");
  GRASP_LAUNCH(score_from_grad_kernel, dim3((unsigned)ceil_div(r, 256)), dim3(256), 0, stream, dsigma, S, r, metric, score);
  GRASP_CHECK_LAST("score_from_grad_kernel");
  return 0;
}
