// compile step: gather the k retained triplets and emit either the merged dense
// weight W = U_k diag(S_k) Vh_k (reference modeling_grasp.py:440-442, :454) or the two
// sqrt(S)-scaled factors of SVDLinear (reference modeling_grasp.py:40-48).
#include "common.cuh"
#include "gemm_simt.cuh"

namespace grasp {

// from gemm_tc.cu
int tc_gemm_f32(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
                const float* B, int64_t ldb, float beta, void* C, int64_t ldc, int c_bf16, int prec, void* ws,
                size_t ws_bytes, void* stream);
size_t tc_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int prec);
size_t tc_planes_bytes(int64_t rows, int64_t cols);
int tc_split_f16(const float* src, int64_t ld, int64_t rows, int64_t cols, int scale_mode, void* planes, float* inv,
                 void* stream);
int tc_gemm_planes(int64_t M, int64_t N, int64_t K, float alpha, const void* Ap, const float* inv_a, const void* Bp,
                   int b_kn, const float* inv_b, float beta, float* C, int64_t ldc, void* stream);

int tc_gemm_planes_out(int64_t M, int64_t N, int64_t K, const void* Ap, const float* inv_a, const void* Bp, int b_kn,
                       const float* inv_b, void* out_planes, float* out_inv, void* stream);

// Uk[a][j] = U[a][idx[j]] * su(j),  j < kp (zero beyond k);  su = 1 (mode 0) or sqrt(S) (mode 1)
__global__ void gather_cols_kernel(const float* __restrict__ U, const float* __restrict__ S,
                                   const int64_t* __restrict__ idx, int64_t out, int64_t r, int64_t k, int64_t kp,
                                   int mode, float* __restrict__ Uk) {
  const int64_t a = blockIdx.x;
  for (int64_t j = threadIdx.x; j < kp; j += blockDim.x) {
    float v = 0.f;
    if (j < k) {
      const int64_t c = idx[j];
      v = U[a * r + c];
      if (mode == 1) v *= sqrtf(S[c]);
    }
    Uk[a * kp + j] = v;
  }
}

// Vk[j][b] = Vh[idx[j]][b] * sv(j);  sv = S (mode 0) or sqrt(S) (mode 1); rows j >= k are zero
__global__ void gather_rows_kernel(const float* __restrict__ Vh, const float* __restrict__ S,
                                   const int64_t* __restrict__ idx, int64_t in, int64_t k, int mode,
                                   float* __restrict__ Vk) {
  const int64_t j = blockIdx.x;
  float scale = 0.f;
  const float* src = Vh;
  if (j < k) {
    const int64_t c = idx[j];
    scale = (mode == 1) ? sqrtf(S[c]) : S[c];
    src = Vh + c * in;
  }
  for (int64_t b = threadIdx.x; b < in; b += blockDim.x) Vk[j * in + b] = (j < k) ? src[b] * scale : 0.f;
}

__global__ void validate_idx_kernel(const int64_t* __restrict__ idx, int64_t k, int64_t r, int* __restrict__ bad) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < k && (idx[j] < 0 || idx[j] >= r)) atomicExch(bad, 1);
}

}  // namespace grasp

using namespace grasp;

static int64_t rebuild_kp(int64_t k) { return round_up(k, 64); }

extern "C" size_t grasp_lowrank_rebuild_workspace_bytes(int64_t out, int64_t in, int64_t k, int prec) {
  if (out <= 0 || in <= 0 || k < 0) return 0;
  const int64_t kp = rebuild_kp(k > 0 ? k : 1);
  size_t b = (size_t)(out + in) * kp * 4 + 512;
  if (prec != GRASP_PREC_SIMT) b += tc_gemm_workspace_bytes(out, in, kp, prec);
  return b;
}

extern "C" int grasp_lowrank_rebuild(const float* U, const float* S, const float* Vh, const int64_t* idx, int64_t k,
                                     int64_t out, int64_t in, int64_t r, int out_dtype, void* W, int prec, void* ws,
                                     size_t ws_bytes, void* stream) {
  if (!U || !S || !Vh || !W || !ws) return bad_arg("rebuild: null");
  if (k > 0 && !idx) return bad_arg("rebuild: null idx");
  if (out <= 0 || in <= 0 || r <= 0 || k < 0 || k > r) return bad_arg("rebuild: out/in/r/k");
  if (out_dtype != GRASP_DTYPE_F32 && out_dtype != GRASP_DTYPE_BF16) return bad_arg("rebuild: out_dtype");
  if (ws_bytes < grasp_lowrank_rebuild_workspace_bytes(out, in, k, prec)) return bad_arg("rebuild: workspace too small");
  if (reinterpret_cast<uintptr_t>(ws) & 255) return bad_arg("rebuild: workspace must be 256-byte aligned");
  const int64_t kp = rebuild_kp(k > 0 ? k : 1);
  float* Uk = static_cast<float*>(ws);
  float* Vk = Uk + out * kp;
  GRASP_LAUNCH(gather_cols_kernel, dim3((unsigned)out), dim3(128), 0, stream, U, S, idx, out, r, k, kp, 0, Uk);
  GRASP_LAUNCH(gather_rows_kernel, dim3((unsigned)kp), dim3(256), 0, stream, Vh, S, idx, in, k, 0, Vk);
  GRASP_CHECK_LAST("rebuild gather");
  const int c_bf16 = (out_dtype == GRASP_DTYPE_BF16);
  if (prec == GRASP_PREC_SIMT) {
    GemmStrided g{};
    g.A = Uk; g.sAm = kp; g.sAk = 1;
    g.B = Vk; g.sBk = in; g.sBn = 1;
    g.C = W; g.ldc = in; g.c_bf16 = c_bf16;
    g.M = out; g.N = in; g.K = kp; g.alpha = 1.f; g.beta = 0.f;
    return launch_gemm_simt(g, 1, stream);
  }
  unsigned char* tws = reinterpret_cast<unsigned char*>(Vk + in * kp);
  tws = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tws) + 255) & ~(uintptr_t)255);
  const size_t used = (size_t)(tws - static_cast<unsigned char*>(ws));
  return tc_gemm_f32(0, 0, out, in, kp, 1.f, Uk, kp, Vk, in, 0.f, W, in, c_bf16, prec, tws, ws_bytes - used, stream);
}

extern "C" int grasp_factor_pack(const float* U, const float* S, const float* Vh, const int64_t* idx, int64_t k,
                                 int64_t out, int64_t in, int64_t r, float* in_w, float* out_w, void* stream) {
  if (!U || !S || !Vh || !in_w || !out_w) return bad_arg("factor_pack: null");
  if (out <= 0 || in <= 0 || r <= 0 || k < 0 || k > r) return bad_arg("factor_pack: out/in/r/k");
  if (k == 0) return 0;
  if (!idx) return bad_arg("factor_pack: null idx");
  GRASP_LAUNCH(gather_cols_kernel, dim3((unsigned)out), dim3(128), 0, stream, U, S, idx, out, r, k, k, 1, out_w);
  GRASP_LAUNCH(gather_rows_kernel, dim3((unsigned)k), dim3(256), 0, stream, Vh, S, idx, in, k, 1, in_w);
  GRASP_CHECK_LAST("factor_pack");
  return 0;
}

extern "C" size_t grasp_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int prec) {
  if (prec == GRASP_PREC_SIMT) return 0;
  return tc_gemm_workspace_bytes(M, N, K, prec);
}

extern "C" int grasp_gemm_f32(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
                              int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int prec,
                              void* ws, size_t ws_bytes, void* stream) {
  if (!A || !B || !C) return bad_arg("gemm: null");
  if (M < 0 || N < 0 || K < 0) return bad_arg("gemm: M/N/K");
  if (lda < (ta ? M : K) || ldb < (tb ? K : N) || ldc < N) return bad_arg("gemm: leading dimension");
  if (M == 0 || N == 0) return 0;
  if (prec == GRASP_PREC_SIMT) {
    GemmStrided g{};
    g.A = A; g.sAm = ta ? 1 : lda; g.sAk = ta ? lda : 1;
    g.B = B; g.sBk = tb ? 1 : ldb; g.sBn = tb ? ldb : 1;
    g.C = C; g.ldc = ldc; g.c_bf16 = 0;
    g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
    return launch_gemm_simt(g, 1, stream);
  }
  if (prec != GRASP_PREC_BF16X3 && prec != GRASP_PREC_BF16X6 && prec != GRASP_PREC_F16X3) return bad_arg("gemm: prec");
  return tc_gemm_f32(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, 0, prec, ws, ws_bytes, stream);
}

extern "C" size_t grasp_gemm_planes_bytes(int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  return tc_planes_bytes(rows, cols);
}

extern "C" int grasp_gemm_split_f16(const float* src, int64_t ld, int64_t rows, int64_t cols, int scale_mode,
                                    void* planes, float* inv, void* stream) {
  if (!src || !planes || !inv) return bad_arg("split: null");
  if (ld < cols) return bad_arg("split: leading dimension");
  return tc_split_f16(src, ld, rows, cols, scale_mode, planes, inv, stream);
}

extern "C" int grasp_gemm_f16x3_planes(int64_t M, int64_t N, int64_t K, float alpha, const void* A_planes,
                                       const float* inv_a, const void* B_planes, int b_kn, const float* inv_b,
                                       float beta, float* C, int64_t ldc, void* stream) {
  if (!A_planes || !B_planes || !inv_a || !inv_b || !C) return bad_arg("gemm_planes: null");
  if (ldc < N) return bad_arg("gemm_planes: leading dimension");
  return tc_gemm_planes(M, N, K, alpha, A_planes, inv_a, B_planes, b_kn, inv_b, beta, C, ldc, stream);
}

extern "C" int grasp_gemm_f16x3_planes_out(int64_t M, int64_t N, int64_t K, const void* A_planes, const float* inv_a,
                                           const void* B_planes, int b_kn, const float* inv_b, void* out_planes,
                                           float* out_inv, void* stream) {
  if (!A_planes || !B_planes || !inv_a || !inv_b || !out_planes || !out_inv) return bad_arg("gemm_planes_out: null");
  return tc_gemm_planes_out(M, N, K, A_planes, inv_a, B_planes, b_kn, inv_b, out_planes, out_inv, stream);
}
