/*
 * grasp_b200.h -- C ABI of the B200-native GRASP compression hot path.
 *
 * The reference (compressionOrg/GRASP) is pure Python/PyTorch and has no FFI of
 * its own; each entry point below replaces one torch library call site of the
 * reference hot path (file:line given per function, paths relative to the
 * reference root).  Contract shared by every function unless stated otherwise:
 *
 *   - all pointers are DEVICE pointers owned by the caller (inputs, outputs and
 *     workspace); matrices are row-major with explicit leading dimensions in
 *     ELEMENTS; nothing is allocated or freed inside the library;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); the call
 *     never synchronises the device and is re-entrant per stream;
 *   - return value: 0 = ok, <0 = bad argument (nothing was enqueued),
 *     >0 = CUDA error code; text via grasp_last_error() (thread-local);
 *   - there is no CPU fallback: without an sm_100 device the calls return a
 *     CUDA error.
 */
#ifndef GRASP_B200_H
#define GRASP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRASP_ABI_VERSION 1

/* element types of hidden states / outputs */
#define GRASP_DTYPE_F32  0
#define GRASP_DTYPE_BF16 1
#define GRASP_DTYPE_F16  2

/* importance metric, reference modeling_grasp.py:392-397 */
#define GRASP_METRIC_GRADIENT 0
#define GRASP_METRIC_TAYLOR   1

/* GEMM arithmetic used by the tensor-core paths: fp32 operands are split into
 * bf16 planes (hi, mid, lo) and multiplied on tcgen05 with fp32 accumulation.
 *   GRASP_PREC_SIMT  : plain fp32 FMA on CUDA cores (validation path)
 *   GRASP_PREC_BF16X3: 2 planes, 3 MMAs  (rel. err ~4e-6)
 *   GRASP_PREC_BF16X6: 3 planes, 6 MMAs  (rel. err ~1e-7, fp32-class)
 *   GRASP_PREC_F16X3 : 2 fp16 planes of row-scaled operands, 3 MMAs (rel. err ~3e-7, fp32-class
 *                      for operands whose rows span < 2^14 in magnitude around their maximum)   */
#define GRASP_PREC_SIMT   0
#define GRASP_PREC_BF16X3 3
#define GRASP_PREC_BF16X6 6
#define GRASP_PREC_F16X3  16
/* flag, OR-ed into the `prec` of grasp_svd_batched: factor the matrix as it is, without the CholeskyQR2
 * preconditioning of wide / tall matrices (the retry when info reports a failed preconditioning) */
#define GRASP_SVD_NO_PRECOND 0x100

int         grasp_abi_version(void);
const char* grasp_last_error(void);
/* number of kernels this library has launched in the calling process (all
 * streams); bench.py reports the difference over its timed region. */
uint64_t    grasp_launch_count(void);

/* ---------------------------------------------------------------------------
 * (a1) block influence of one layer pair.  Replaces tools/utils_func.py:3-25
 * (block_influence) as called from modeling_grasp.py:163-167: per token row t
 *   sim_t = <x_t,y_t>/(|x_t||y_t|), NaN -> 0.5, bi_t = 1-sim_t
 *   (angular != 0: bi_t = arccos(sim_t)/pi)
 * and *acc += scale * mean_t(bi_t)  (acc is a device double, so the per-batch sums
 * the reference keeps in Python floats stay on the device; scale = 1 reproduces the
 * reference, other values re-weight a launch that carries several batches).  per_row (nullable)
 * receives bi_t as fp32 [rows].  ld = row stride in elements.
 * ------------------------------------------------------------------------- */
int grasp_bi_accumulate(const void* h_in, const void* h_out, int64_t rows, int64_t d,
                        int64_t ld, int dtype, int angular, double scale,
                        double* acc, float* per_row, void* stream);

/* (a2) the whole chain of one forward pass: hiddens[0..n_states) are the
 * L+1 hidden states of modeling_grasp.py:180-183 (HOST array of device
 * pointers, each [rows, d] with row stride ld); acc[i] += scale * mean_t BI(h[i],h[i+1])
 * for i in [0, n_states-1).  Every hidden state is read from HBM exactly once.
 * n_states <= 130. */
int grasp_bi_chain(const void* const* hiddens, int n_states, int64_t rows, int64_t d,
                   int64_t ld, int dtype, double scale, double* acc, void* stream);

/* ---------------------------------------------------------------------------
 * (a3) thin SVD, replaces torch.linalg.svd(w, full_matrices=False) at
 * modeling_grasp.py:231.  A[i] is m[i] x n[i] fp32 (lda[i]); outputs
 * U[i] [m, r] (ld r), S[i] [r] descending, Vh[i] [r, n] (ld n), r = min(m,n).
 * Wide / tall matrices (long side >= 1.5 x short side, short side >= 512) are first reduced to the square factor
 * of a CholeskyQR2 (tensor-core GEMMs), unless GRASP_SVD_NO_PRECOND is set in `prec`.
 * info (device int32 [4*batch]): {sweeps used, converged(0/1; 0 also when the preconditioning was not sound),
 * float bits of the last sweep's max relative off-diagonal, sweeps of the
 * tensor-core phase}.
 * prec: GRASP_PREC_*; max_sweeps <= 0 selects the default (32).
 * All shape arrays are HOST arrays; A/U/S/Vh are HOST arrays of device ptrs.
 * ------------------------------------------------------------------------- */
size_t grasp_svd_workspace_bytes(int batch, const int64_t* m, const int64_t* n);
int    grasp_svd_batched(int batch, const float* const* A, const int64_t* m, const int64_t* n,
                         const int64_t* lda, float* const* U, float* const* S, float* const* Vh,
                         int32_t* info, int prec, int max_sweeps,
                         void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * (a5/a6) singular-value gradient and importance.  Replaces the autograd path
 * through GRASPLayer (modeling_grasp.py:75-79, :354-363) by the identity
 * dL/dS_i = u_i^T G v_i with G = dL/dW [out,in] accumulated over the
 * calibration samples, and the score of modeling_grasp.py:392-395:
 *   dsigma[i] (+)= sum_{a,b} U[a,i] G[a,b] Vh[i,b]
 *   score[i]   = |dsigma[i]| (gradient) or |dsigma[i]*S[i]| (taylor)
 * U [out,r], G [out,in], Vh [r,in], S [r], all fp32 contiguous.
 * accumulate != 0 adds into dsigma (multi-call / multi-rank partial sums);
 * score may be NULL (e.g. before the all-reduce of dsigma).
 * ------------------------------------------------------------------------- */
size_t grasp_sigma_score_workspace_bytes(int64_t out, int64_t in, int64_t r, int prec);
int    grasp_sigma_score(const float* U, const float* G, const float* Vh, const float* S,
                         int64_t out, int64_t in, int64_t r, int metric, int accumulate,
                         float* dsigma, float* score, int prec,
                         void* ws, size_t ws_bytes, void* stream);
/* score only (after an all-reduce of dsigma): score[i] = |g[i]| or |g[i]*S[i]| */
int    grasp_score_from_grad(const float* dsigma, const float* S, int64_t r, int metric,
                             float* score, void* stream);

/* (a6) per-matrix top-k, replaces torch.topk at modeling_grasp.py:404.
 * idx[i][0..k[i]) = indices of the k largest scores, sorted by score
 * descending (ties: lower index first; NaN ranks above +inf like torch).
 * score/idx are HOST arrays of device pointers, r/k HOST arrays. r <= 65536. */
int grasp_topk_batched(int batch, const float* const* score, const int64_t* r, const int64_t* k,
                       int64_t* const* idx, void* stream);

/* (a6, threshold mode) tools/utils_func.py:45-57: sort descending, keep the
 * shortest prefix whose running sum reaches target_ratio * sum(score).
 * idx [r] receives the full descending order, *count (device int64) the
 * prefix length. */
int grasp_adaptive_rank(const float* score, int64_t r, double target_ratio,
                        int64_t* idx, int64_t* count, void* stream);

/* ---------------------------------------------------------------------------
 * (a7) compile.  Replaces modeling_grasp.py:440-442 + :454 (merge) and
 * :47-48 (SVDLinear "UV" fuse).
 *   rebuild: W[out,in] = U[:,idx] diag(S[idx]) Vh[idx,:]   (fp32 or bf16 out)
 *   pack   : in_w [k,in] = Vh[idx,:]*sqrt(S[idx])[:,None],
 *            out_w[out,k] = U[:,idx]*sqrt(S[idx])[None,:]
 * ------------------------------------------------------------------------- */
size_t grasp_lowrank_rebuild_workspace_bytes(int64_t out, int64_t in, int64_t k, int prec);
int    grasp_lowrank_rebuild(const float* U, const float* S, const float* Vh, const int64_t* idx,
                             int64_t k, int64_t out, int64_t in, int64_t r,
                             int out_dtype, void* W, int prec,
                             void* ws, size_t ws_bytes, void* stream);
int    grasp_factor_pack(const float* U, const float* S, const float* Vh, const int64_t* idx,
                         int64_t k, int64_t out, int64_t in, int64_t r,
                         float* in_w, float* out_w, void* stream);

/* ---------------------------------------------------------------------------
 * (f1) fp32 GEMM on the same arithmetic the SVD uses (building block, also
 * used by the calibration engine for G += dY^T X):
 *   C[M,N] = alpha * op(A) op(B) + beta * C,  op = transpose when ta/tb != 0
 * A is [M,K] (ta=0) or [K,M] (ta=1); B is [K,N] (tb=0) or [N,K] (tb=1).
 * ------------------------------------------------------------------------- */
size_t grasp_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int prec);
int    grasp_gemm_f32(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha,
                      const float* A, int64_t lda, const float* B, int64_t ldb,
                      float beta, float* C, int64_t ldc, int prec,
                      void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * (f1) prepared GEMM operands of the F16X3 arithmetic.  The calibration passes
 * (modeling_grasp.py:340-354 run the whole model per sample) multiply one
 * weight by every micro-batch, forward as [N][K] and backward as [K][N], and one
 * activation by several weights; the split pre-pass is therefore callable on
 * its own and its result reusable:
 *   planes  [2][rows][pitch(cols)] fp16 (hi, lo) of src * scale, in the STORED
 *           orientation, grasp_gemm_planes_bytes() bytes, 1024-byte aligned
 *   inv     inverse scales: GRASP_SCALE_ROWS   -> [rows], one power of two per row
 *                           GRASP_SCALE_TENSOR -> [max(rows,cols) + 1] floats, one
 *                           power of two for the whole tensor (last word scratch);
 *                           valid as A, as B [N][K] and as B [K][N]
 * grasp_gemm_f16x3_planes: C[M,N] = alpha * A * op(B) + beta * C with
 *   A planes [2][M][Kp] (row or tensor scale), B planes [2][N][Kp] (b_kn = 0,
 *   row or tensor scale) or [2][K][Np] (b_kn = 1, tensor scale only).
 * ------------------------------------------------------------------------- */
#define GRASP_SCALE_ROWS   0
#define GRASP_SCALE_TENSOR 2
size_t grasp_gemm_planes_bytes(int64_t rows, int64_t cols);
int    grasp_gemm_split_f16(const float* src, int64_t ld, int64_t rows, int64_t cols, int scale_mode,
                            void* planes, float* inv, void* stream);
int    grasp_gemm_f16x3_planes(int64_t M, int64_t N, int64_t K, float alpha,
                               const void* A_planes, const float* inv_a,
                               const void* B_planes, int b_kn, const float* inv_b,
                               float beta, float* C, int64_t ldc, void* stream);

/* (f2) the first GEMM of a factor pair (SVDLinear.forward, modeling_grasp.py:57-59: OutLinear(InLinear(x)), and the
 * mirrored backward): out = A op(B) leaves as a prepared operand -- row-scaled planes [2][M][pitch(N)] + inv [M] --
 * that the second GEMM consumes, so the [tokens, k] intermediate is never written in fp32 nor split by a pre-pass.
 * B must be tensor-scaled (weights are). */
int    grasp_gemm_f16x3_planes_out(int64_t M, int64_t N, int64_t K,
                                   const void* A_planes, const float* inv_a,
                                   const void* B_planes, int b_kn, const float* inv_b,
                                   void* out_planes, float* out_inv, void* stream);

/* ---------------------------------------------------------------------------
 * (f1) row-wise pieces of the LLaMA decoder layer the calibration passes run
 * between the GEMMs (the reference leaves them to transformers' eager modules,
 * modeling_grasp.py:347 -> LlamaDecoderLayer.forward: ~25 elementwise launches
 * per layer).  fp32, row-major, contiguous rows of length d unless a stride is
 * given.  Each backward is the exact derivative of its forward.
 *   rmsnorm : y = x * rsqrt(mean(x^2) + eps) * w ; rstd[t] kept for backward
 *             bwd: dx = rstd * (g - x * rstd^2 * mean(g * x)) (+ add), g = dy * w
 *   rope    : in place on x [tokens][heads][hd] (token t has position t % seq):
 *             (x1, x2) -> (x1 c1 - x2 s1, x2 c2 + x1 s2), c/s rows of cos/sin [seq][hd]
 *             (batch stride cs_batch elements, 0 = shared); inverse != 0 applies the
 *             transpose (the backward)
 *   swiglu  : h = silu(g) * u ; bwd: dg = dh * u * s(g) (1 + g (1 - s(g))), du = dh * silu(g)
 *   ce_loss : per row t of logits [rows][V]: loss[t] = coef[t] * (lse_t - logit[t, label_t]),
 *             logits overwritten by dloss/dlogits = coef[t] * (softmax - onehot);
 *             label < 0 -> loss 0, gradient 0 (ignore_index)
 * ------------------------------------------------------------------------- */
/* y and/or the GEMM operand form of y (planes + inv as written by grasp_gemm_split_f16 with
 * GRASP_SCALE_ROWS) -- the consumer of a norm is always a linear; either may be NULL, not both */
int grasp_rmsnorm_fwd(const float* x, const float* w, int64_t rows, int64_t d, float eps,
                      float* y, float* rstd, void* planes, float* inv, void* stream);
int grasp_rmsnorm_bwd(const float* dy, const float* x, const float* w, const float* rstd,
                      const float* add, int64_t rows, int64_t d, float* dx, void* stream);
int grasp_rope_inplace(float* x, int64_t tokens, int64_t seq, int64_t heads, int64_t hd,
                       const float* cos, const float* sin, int64_t cs_batch, int inverse, void* stream);
/* g, u, h, dh, dg, du are [rows][cols] contiguous.  h / (dg, du) may be NULL when only their operand
 * forms are wanted (h feeds down_proj, dg / du the backward of gate_proj / up_proj); the operand forms
 * need cols <= 51200 (forward) / 25600 (backward).  dg / du may alias g / u. */
int grasp_swiglu_fwd(const float* g, const float* u, int64_t rows, int64_t cols, float* h,
                     void* planes, float* inv, void* stream);
int grasp_swiglu_bwd(const float* dh, const float* g, const float* u, int64_t rows, int64_t cols,
                     float* dg, float* du, void* dg_planes, float* dg_inv,
                     void* du_planes, float* du_inv, void* stream);
int grasp_ce_loss_bwd(float* logits, const int64_t* labels, const float* coef, int64_t rows, int64_t V,
                      float* loss, void* stream);

/* ---------------------------------------------------------------------------
 * (f1) causal self-attention of the calibration passes (the reference reaches it through
 * transformers' LlamaAttention inside self.model(...) / loss.backward(), modeling_grasp.py:347-354):
 *   O = softmax(scale * Q K^T + causal mask) V  per (batch, head), grouped-query heads (H % Hkv == 0),
 * in the F16X3 tensor-core arithmetic.  q / k / v / dO are given as prepared operands of the
 * [B*S, heads*D] activations (after RoPE), split with GRASP_SCALE_TENSOR; head_dim D is 64 or 128.
 *   fwd: out [B*S][H*D] fp32, lse2 [B][H][S] = log2 sum_k exp(scale * q.k) (kept for the backward)
 *   bwd: dq [B*S][H*D], dk / dv [B*S][Hkv*D] (summed over the query heads of a group), from dO (fp32 and its
 *        planes), the forward's out and lse2; delta_ws holds B*H*S floats.
 * ------------------------------------------------------------------------- */
/* RoPE (transformers' apply_rotary_pos_emb, cos / sin as for grasp_rope_inplace) on q and k, and the tensor-scaled
 * operand planes of the rotated q, k and of v, in two passes over the fp32 projections (which stay untouched).
 * inv arrays as for GRASP_SCALE_TENSOR (max(tokens, heads*D) + 1 floats); ws: 16 bytes of scratch. */
int grasp_attn_prep_qkv(const float* q, const float* k, const float* v, int64_t tokens, int64_t seq, int H, int Hkv,
                        int D, const float* cos, const float* sin, int64_t cs_batch,
                        void* q_planes, float* q_inv, void* k_planes, float* k_inv, void* v_planes, float* v_inv,
                        void* ws, void* stream);
int grasp_attn_fwd(const void* q_planes, const float* inv_q, const void* k_planes, const float* inv_k,
                   const void* v_planes, const float* inv_v, int B, int S, int H, int Hkv, int D, float scale,
                   float* out, float* lse2, void* stream);
int grasp_attn_bwd(const void* q_planes, const float* inv_q, const void* k_planes, const float* inv_k,
                   const void* v_planes, const float* inv_v, const void* do_planes, const float* inv_do,
                   const float* dO, const float* O, const float* lse2, int B, int S, int H, int Hkv, int D,
                   float scale, float* dq, float* dk, float* dv, float* delta_ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRASP_B200_H */
