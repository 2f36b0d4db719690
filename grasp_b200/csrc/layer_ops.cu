// Row-wise pieces of a LLaMA decoder layer between the GEMMs of a calibration pass:
// RMSNorm, rotary embedding, SwiGLU and the cross-entropy loss, each with its exact backward.
// The reference leaves these to transformers' eager modules (modeling_grasp.py:347 runs
// LlamaForCausalLM.forward; ~25 elementwise launches per layer and direction, every one a
// full read + write of the activations).  All of them are HBM-bound: one CTA per token row
// (row staged in shared memory, so every operand is read once and every result written once),
// 16-byte accesses, fp32 arithmetic.
#include "common.cuh"
#include "split_f16.cuh"

namespace grasp {

constexpr int ROW_THREADS = 256;

__device__ __forceinline__ float block_sum_256(float v, float* red /*[8]*/) {
  v = warp_sum(v);
  __syncthreads();                       // red may still be read from a previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_THREADS / 32; ++i) t += red[i];   // same order in every thread: deterministic
  return t;
}

// ------------------------------------------------------------------------------------ RMSNorm
// y = x * rstd * w, rstd = rsqrt(mean(x^2) + eps)   (transformers LlamaRMSNorm.forward)
// y (fp32) and/or its GEMM operand form (fp16 hi/lo planes + row scale) are written; the consumer of a norm is
// always a linear, so the passes usually want only the planes: x is read once and nothing is re-read.
__global__ void __launch_bounds__(ROW_THREADS)
rmsnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int d, int dp, float eps,
                   float* __restrict__ y, float* __restrict__ rstd, uint16_t* __restrict__ planes,
                   float* __restrict__ inv, int64_t rows) {
  extern __shared__ __align__(16) float rowbuf[];   // dp floats (d rounded up to 8)
  __shared__ float red[9];
  const int64_t r = blockIdx.x;
  const float* xr = x + r * d;
  float ss = 0.f;
  const bool vec = (d & 3) == 0;
  if (vec) {
    for (int k = threadIdx.x * 4; k < d; k += ROW_THREADS * 4) {
      const float4 v = ldg_stream(reinterpret_cast<const float4*>(xr + k));
      *reinterpret_cast<float4*>(rowbuf + k) = v;
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (int k = threadIdx.x; k < d; k += ROW_THREADS) {
      const float v = xr[k];
      rowbuf[k] = v;
      ss += v * v;
    }
  }
  ss = block_sum_256(ss, red);
  const float rs = rsqrtf(ss / (float)d + eps);
  if (threadIdx.x == 0) rstd[r] = rs;
  float* yr = y ? y + r * d : nullptr;
  float m = 0.f;
  if (vec) {
    for (int k = threadIdx.x * 4; k < d; k += ROW_THREADS * 4) {
      const float4 v = *reinterpret_cast<const float4*>(rowbuf + k);
      const float4 ww = *reinterpret_cast<const float4*>(w + k);
      float4 o;
      o.x = ww.x * (v.x * rs); o.y = ww.y * (v.y * rs); o.z = ww.z * (v.z * rs); o.w = ww.w * (v.w * rs);
      if (yr) *reinterpret_cast<float4*>(yr + k) = o;
      if (planes) {
        *reinterpret_cast<float4*>(rowbuf + k) = o;
        m = fmaxf(fmaxf(m, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
      }
    }
  } else {
    for (int k = threadIdx.x; k < d; k += ROW_THREADS) {
      const float o = w[k] * (rowbuf[k] * rs);
      if (yr) yr[k] = o;
      if (planes) { rowbuf[k] = o; m = fmaxf(m, fabsf(o)); }
    }
  }
  if (planes) {
    for (int k = d + threadIdx.x; k < dp; k += ROW_THREADS) rowbuf[k] = 0.f;
    block_row_split_256(rowbuf, m, dp, planes + r * dp, planes + (rows + r) * dp, inv + r, red);
  }
}

// dx = rstd * (g - x * rstd^2 * mean(g x)) + add,  g = dy * w
__global__ void __launch_bounds__(ROW_THREADS)
rmsnorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                   const float* __restrict__ rstd, const float* __restrict__ add, int d, float* __restrict__ dx) {
  extern __shared__ __align__(16) float rowbuf[];   // g [d] | x [d]
  __shared__ float red[8];
  float* gbuf = rowbuf;
  float* xbuf = rowbuf + d;
  const int64_t r = blockIdx.x;
  const float* dyr = dy + r * d;
  const float* xr = x + r * d;
  float dot = 0.f;
  const bool vec = (d & 3) == 0;
  if (vec) {
    for (int k = threadIdx.x * 4; k < d; k += ROW_THREADS * 4) {
      const float4 a = ldg_stream(reinterpret_cast<const float4*>(dyr + k));
      const float4 b = ldg_stream(reinterpret_cast<const float4*>(xr + k));
      const float4 ww = *reinterpret_cast<const float4*>(w + k);
      const float4 g = make_float4(a.x * ww.x, a.y * ww.y, a.z * ww.z, a.w * ww.w);
      *reinterpret_cast<float4*>(gbuf + k) = g;
      *reinterpret_cast<float4*>(xbuf + k) = b;
      dot += g.x * b.x + g.y * b.y + g.z * b.z + g.w * b.w;
    }
  } else {
    for (int k = threadIdx.x; k < d; k += ROW_THREADS) {
      const float g = dyr[k] * w[k], b = xr[k];
      gbuf[k] = g;
      xbuf[k] = b;
      dot += g * b;
    }
  }
  dot = block_sum_256(dot, red);
  const float rs = rstd[r];
  const float c = rs * rs * rs * (dot / (float)d);
  float* dxr = dx + r * d;
  const float* ar = add ? add + r * d : nullptr;
  if (vec) {
    for (int k = threadIdx.x * 4; k < d; k += ROW_THREADS * 4) {
      const float4 g = *reinterpret_cast<const float4*>(gbuf + k);
      const float4 b = *reinterpret_cast<const float4*>(xbuf + k);
      float4 o = make_float4(rs * g.x - c * b.x, rs * g.y - c * b.y, rs * g.z - c * b.z, rs * g.w - c * b.w);
      if (ar) {
        const float4 e = ldg_stream(reinterpret_cast<const float4*>(ar + k));
        o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
      }
      *reinterpret_cast<float4*>(dxr + k) = o;
    }
  } else {
    for (int k = threadIdx.x; k < d; k += ROW_THREADS) dxr[k] = rs * gbuf[k] - c * xbuf[k] + (ar ? ar[k] : 0.f);
  }
}

// --------------------------------------------------------------------------------------- RoPE
// transformers apply_rotary_pos_emb: out = x * cos + rotate_half(x) * sin with rotate_half(x) = (-x2, x1):
//   out1 = x1 c1 - x2 s1,  out2 = x2 c2 + x1 s2        (c1/s1 = first half of the cos/sin row, c2/s2 second)
// inverse (the transpose, used by the backward):  dx1 = dy1 c1 + dy2 s2,  dx2 = dy2 c2 - dy1 s1
__global__ void __launch_bounds__(256)
rope_kernel(float* __restrict__ x, int64_t tokens, int seq, int heads, int hd, const float* __restrict__ cosp,
            const float* __restrict__ sinp, int64_t cs_batch, int inverse) {
  const int half = hd >> 1;
  const int quads = half >> 2;                        // float4 groups per half head
  const int64_t per_token = (int64_t)heads * quads;
  const int64_t total = tokens * per_token;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / per_token;
    const int rem = (int)(i - t * per_token);
    const int h = rem / quads, j = (rem - h * quads) * 4;
    const int64_t b = t / seq;
    const int s = (int)(t - b * seq);
    const float* cr = cosp + b * cs_batch + (int64_t)s * hd;
    const float* sr = sinp + b * cs_batch + (int64_t)s * hd;
    float* p1 = x + (t * heads + h) * hd + j;
    float* p2 = p1 + half;
    const float4 x1 = *reinterpret_cast<const float4*>(p1), x2 = *reinterpret_cast<const float4*>(p2);
    const float4 c1 = *reinterpret_cast<const float4*>(cr + j), c2 = *reinterpret_cast<const float4*>(cr + half + j);
    const float4 s1 = *reinterpret_cast<const float4*>(sr + j), s2 = *reinterpret_cast<const float4*>(sr + half + j);
    float4 o1, o2;
    if (!inverse) {
      o1.x = x1.x * c1.x - x2.x * s1.x; o1.y = x1.y * c1.y - x2.y * s1.y;
      o1.z = x1.z * c1.z - x2.z * s1.z; o1.w = x1.w * c1.w - x2.w * s1.w;
      o2.x = x2.x * c2.x + x1.x * s2.x; o2.y = x2.y * c2.y + x1.y * s2.y;
      o2.z = x2.z * c2.z + x1.z * s2.z; o2.w = x2.w * c2.w + x1.w * s2.w;
    } else {
      o1.x = x1.x * c1.x + x2.x * s2.x; o1.y = x1.y * c1.y + x2.y * s2.y;
      o1.z = x1.z * c1.z + x2.z * s2.z; o1.w = x1.w * c1.w + x2.w * s2.w;
      o2.x = x2.x * c2.x - x1.x * s1.x; o2.y = x2.y * c2.y - x1.y * s1.y;
      o2.z = x2.z * c2.z - x1.z * s1.z; o2.w = x2.w * c2.w - x1.w * s1.w;
    }
    *reinterpret_cast<float4*>(p1) = o1;
    *reinterpret_cast<float4*>(p2) = o2;
  }
}

// ------------------------------------------------------------------------------------- SwiGLU
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const float* __restrict__ g, const float* __restrict__ u, int64_t n, float* __restrict__ h) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = ldg_stream(reinterpret_cast<const float4*>(g) + i);
    const float4 b = ldg_stream(reinterpret_cast<const float4*>(u) + i);
    float4 o;
    o.x = a.x * sigmoidf_(a.x) * b.x; o.y = a.y * sigmoidf_(a.y) * b.y;
    o.z = a.z * sigmoidf_(a.z) * b.z; o.w = a.w * sigmoidf_(a.w) * b.w;
    reinterpret_cast<float4*>(h)[i] = o;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    h[i] = g[i] * sigmoidf_(g[i]) * u[i];
}

__device__ __forceinline__ void swiglu_bwd_1(float dh, float g, float u, float& dg, float& du) {
  const float s = sigmoidf_(g);
  du = dh * (g * s);
  dg = dh * u * (s * (1.f + g * (1.f - s)));
}

__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ g, const float* __restrict__ u, int64_t n,
                  float* __restrict__ dg, float* __restrict__ du) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 d = ldg_stream(reinterpret_cast<const float4*>(dh) + i);
    const float4 a = reinterpret_cast<const float4*>(g)[i];      // dg / du may alias g / u: plain loads
    const float4 b = reinterpret_cast<const float4*>(u)[i];
    float4 og, ou;
    swiglu_bwd_1(d.x, a.x, b.x, og.x, ou.x); swiglu_bwd_1(d.y, a.y, b.y, og.y, ou.y);
    swiglu_bwd_1(d.z, a.z, b.z, og.z, ou.z); swiglu_bwd_1(d.w, a.w, b.w, og.w, ou.w);
    reinterpret_cast<float4*>(dg)[i] = og;
    reinterpret_cast<float4*>(du)[i] = ou;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float og, ou;
    swiglu_bwd_1(dh[i], g[i], u[i], og, ou);
    dg[i] = og;
    du[i] = ou;
  }
}

// row forms that also (or only) emit the GEMM operand planes of their results: h feeds down_proj, dg / du feed
// the backward of gate_proj / up_proj, so the fp32 tensors are optional
__global__ void __launch_bounds__(ROW_THREADS)
swiglu_fwd_rows_kernel(const float* __restrict__ g, const float* __restrict__ u, int cols, int cp, float* __restrict__ h,
                       uint16_t* __restrict__ planes, float* __restrict__ inv, int64_t rows) {
  extern __shared__ __align__(16) float rowbuf[];   // cp floats
  __shared__ float red[9];
  const int64_t r = blockIdx.x;
  const float* gr = g + r * cols;
  const float* ur = u + r * cols;
  float* hr = h ? h + r * cols : nullptr;
  float m = 0.f;
  if ((cols & 3) == 0) {
    for (int k = threadIdx.x * 4; k < cols; k += ROW_THREADS * 4) {
      const float4 a = ldg_stream(reinterpret_cast<const float4*>(gr + k));
      const float4 b = ldg_stream(reinterpret_cast<const float4*>(ur + k));
      float4 o;
      o.x = a.x * sigmoidf_(a.x) * b.x; o.y = a.y * sigmoidf_(a.y) * b.y;
      o.z = a.z * sigmoidf_(a.z) * b.z; o.w = a.w * sigmoidf_(a.w) * b.w;
      if (hr) *reinterpret_cast<float4*>(hr + k) = o;
      *reinterpret_cast<float4*>(rowbuf + k) = o;
      m = fmaxf(fmaxf(m, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
    }
  } else {
    for (int k = threadIdx.x; k < cols; k += ROW_THREADS) {
      const float o = gr[k] * sigmoidf_(gr[k]) * ur[k];
      if (hr) hr[k] = o;
      rowbuf[k] = o;
      m = fmaxf(m, fabsf(o));
    }
  }
  for (int k = cols + threadIdx.x; k < cp; k += ROW_THREADS) rowbuf[k] = 0.f;
  block_row_split_256(rowbuf, m, cp, planes + r * cp, planes + (rows + r) * cp, inv + r, red);
}

__global__ void __launch_bounds__(ROW_THREADS)
swiglu_bwd_rows_kernel(const float* __restrict__ dh, const float* __restrict__ g, const float* __restrict__ u, int cols,
                       int cp, float* __restrict__ dg, float* __restrict__ du, uint16_t* __restrict__ dg_planes,
                       float* __restrict__ dg_inv, uint16_t* __restrict__ du_planes, float* __restrict__ du_inv,
                       int64_t rows) {
  extern __shared__ __align__(16) float rowbuf[];   // dg [cp] | du [cp]
  __shared__ float red[9];
  float* bg = rowbuf;
  float* bu = rowbuf + cp;
  const int64_t r = blockIdx.x;
  const float* dr = dh + r * cols;
  const float* gr = g + r * cols;
  const float* ur = u + r * cols;
  float* dgr = dg ? dg + r * cols : nullptr;
  float* dur = du ? du + r * cols : nullptr;
  float mg = 0.f, mu = 0.f;
  if ((cols & 3) == 0) {
    for (int k = threadIdx.x * 4; k < cols; k += ROW_THREADS * 4) {
      const float4 d = ldg_stream(reinterpret_cast<const float4*>(dr + k));
      const float4 a = *reinterpret_cast<const float4*>(gr + k);       // dg / du may alias g / u: plain loads
      const float4 b = *reinterpret_cast<const float4*>(ur + k);
      float4 og, ou;
      swiglu_bwd_1(d.x, a.x, b.x, og.x, ou.x); swiglu_bwd_1(d.y, a.y, b.y, og.y, ou.y);
      swiglu_bwd_1(d.z, a.z, b.z, og.z, ou.z); swiglu_bwd_1(d.w, a.w, b.w, og.w, ou.w);
      if (dgr) *reinterpret_cast<float4*>(dgr + k) = og;
      if (dur) *reinterpret_cast<float4*>(dur + k) = ou;
      *reinterpret_cast<float4*>(bg + k) = og;
      *reinterpret_cast<float4*>(bu + k) = ou;
      mg = fmaxf(fmaxf(mg, fmaxf(fabsf(og.x), fabsf(og.y))), fmaxf(fabsf(og.z), fabsf(og.w)));
      mu = fmaxf(fmaxf(mu, fmaxf(fabsf(ou.x), fabsf(ou.y))), fmaxf(fabsf(ou.z), fabsf(ou.w)));
    }
  } else {
    for (int k = threadIdx.x; k < cols; k += ROW_THREADS) {
      float og, ou;
      swiglu_bwd_1(dr[k], gr[k], ur[k], og, ou);
      if (dgr) dgr[k] = og;
      if (dur) dur[k] = ou;
      bg[k] = og; bu[k] = ou;
      mg = fmaxf(mg, fabsf(og)); mu = fmaxf(mu, fabsf(ou));
    }
  }
  for (int k = cols + threadIdx.x; k < cp; k += ROW_THREADS) { bg[k] = 0.f; bu[k] = 0.f; }
  block_row_split_256(bg, mg, cp, dg_planes + r * cp, dg_planes + (rows + r) * cp, dg_inv + r, red);
  block_row_split_256(bu, mu, cp, du_planes + r * cp, du_planes + (rows + r) * cp, du_inv + r, red);
}

// ------------------------------------------------------------------------------ cross entropy
// One CTA per logits row.  Pass 1 (online max / sum of exponentials) reads the row from HBM, pass 2
// re-reads it (L2-resident: a row is V*4 bytes and only a few hundred rows are in flight) and
// overwrites it with coef * (softmax - onehot).
constexpr int CE_THREADS = 512;

__global__ void __launch_bounds__(CE_THREADS)
ce_loss_bwd_kernel(float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ coef,
                   int V, float* __restrict__ loss) {
  __shared__ float red_m[CE_THREADS / 32], red_s[CE_THREADS / 32];
  const int64_t r = blockIdx.x;
  float* row = logits + r * V;
  const int64_t label = labels[r];
  const float cf = coef[r];
  const bool vec = (V & 3) == 0;
  if (label < 0 || label >= V || cf == 0.f) {          // ignored position: no loss, no gradient
    for (int k = threadIdx.x; k < V; k += CE_THREADS) row[k] = 0.f;
    if (threadIdx.x == 0) loss[r] = 0.f;
    return;
  }
  float m = -INFINITY, s = 0.f;
  auto push = [&](float v) {
    if (v > m) { s = s * expf(m - v) + 1.f; m = v; }
    else s += expf(v - m);
  };
  if (vec) {
    for (int k = threadIdx.x * 4; k < V; k += CE_THREADS * 4) {
      const float4 v = *reinterpret_cast<const float4*>(row + k);
      push(v.x); push(v.y); push(v.z); push(v.w);
    }
  } else {
    for (int k = threadIdx.x; k < V; k += CE_THREADS) push(row[k]);
  }
  // combine (m, s) over the block: global max first, then the rescaled sums in a fixed order
  const float wm = warp_max(m);
  float ws = (m == -INFINITY) ? 0.f : s * expf(m - wm);
  ws = warp_sum(ws);
  if ((threadIdx.x & 31) == 0) { red_m[threadIdx.x >> 5] = wm; red_s[threadIdx.x >> 5] = ws; }
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int i = 0; i < CE_THREADS / 32; ++i) M = fmaxf(M, red_m[i]);
  float S = 0.f;
#pragma unroll
  for (int i = 0; i < CE_THREADS / 32; ++i) S += (red_m[i] == -INFINITY) ? 0.f : red_s[i] * expf(red_m[i] - M);
  const float lse = M + logf(S);
  if (threadIdx.x == 0) loss[r] = cf * (lse - row[label]);
  __syncthreads();                                      // row[label] is read before anyone overwrites it
  if (vec) {
    for (int k = threadIdx.x * 4; k < V; k += CE_THREADS * 4) {
      float4 v = *reinterpret_cast<const float4*>(row + k);
      v.x = cf * expf(v.x - lse); v.y = cf * expf(v.y - lse); v.z = cf * expf(v.z - lse); v.w = cf * expf(v.w - lse);
      const int64_t dl = label - k;
      if (dl == 0) v.x -= cf; else if (dl == 1) v.y -= cf; else if (dl == 2) v.z -= cf; else if (dl == 3) v.w -= cf;
      *reinterpret_cast<float4*>(row + k) = v;
    }
  } else {
    for (int k = threadIdx.x; k < V; k += CE_THREADS) row[k] = cf * expf(row[k] - lse) - (k == label ? cf : 0.f);
  }
}

static int row_smem_attr(const void* fn, size_t bytes, size_t& granted, const char* what) {
  if (bytes <= granted) return 0;
  int rc = check_cuda(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), what);
  if (!rc) granted = 200 * 1024;
  return rc;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace grasp

using namespace grasp;

extern "C" int grasp_rmsnorm_fwd(const float* x, const float* w, int64_t rows, int64_t d, float eps, float* y,
                                 float* rstd, void* planes, float* inv, void* stream) {
  if (!x || !w || !rstd) return bad_arg("rmsnorm_fwd: null");
  if (!y && !planes) return bad_arg("rmsnorm_fwd: neither y nor planes requested");
  if (planes && !inv) return bad_arg("rmsnorm_fwd: planes need inv");
  if (rows < 0 || d <= 0 || d > 50 * 1024) return bad_arg("rmsnorm_fwd: rows/d");
  if (!aligned16(x) || !aligned16(w) || (y && !aligned16(y))) return bad_arg("rmsnorm_fwd: pointers must be 16-byte aligned");
  if (planes && (reinterpret_cast<uintptr_t>(planes) & 1023)) return bad_arg("rmsnorm_fwd: planes must be 1024-byte aligned");
  if (rows == 0) return 0;
  const int64_t dp = plane_pitch(d);
  static size_t granted = 48 * 1024;
  int rc = row_smem_attr((const void*)rmsnorm_fwd_kernel, (size_t)dp * 4, granted, "rmsnorm_fwd attr");
  if (rc) return rc;
  GRASP_LAUNCH(rmsnorm_fwd_kernel, dim3((unsigned)rows), dim3(ROW_THREADS), (size_t)dp * 4, stream, x, w, (int)d, (int)dp,
               eps, y, rstd, static_cast<uint16_t*>(planes), inv, rows);
  GRASP_CHECK_LAST("rmsnorm_fwd_kernel");
  return 0;
}

extern "C" int grasp_rmsnorm_bwd(const float* dy, const float* x, const float* w, const float* rstd, const float* add,
                                 int64_t rows, int64_t d, float* dx, void* stream) {
  if (!dy || !x || !w || !rstd || !dx) return bad_arg("rmsnorm_bwd: null");
  if (rows < 0 || d <= 0 || d > 25 * 1024) return bad_arg("rmsnorm_bwd: rows/d");
  if (!aligned16(dy) || !aligned16(x) || !aligned16(w) || !aligned16(dx) || (add && !aligned16(add)))
    return bad_arg("rmsnorm_bwd: pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  static size_t granted = 48 * 1024;
  int rc = row_smem_attr((const void*)rmsnorm_bwd_kernel, (size_t)d * 8, granted, "rmsnorm_bwd attr");
  if (rc) return rc;
  GRASP_LAUNCH(rmsnorm_bwd_kernel, dim3((unsigned)rows), dim3(ROW_THREADS), (size_t)d * 8, stream, dy, x, w, rstd, add,
               (int)d, dx);
  GRASP_CHECK_LAST("rmsnorm_bwd_kernel");
  return 0;
}

extern "C" int grasp_rope_inplace(float* x, int64_t tokens, int64_t seq, int64_t heads, int64_t hd, const float* cosp,
                                  const float* sinp, int64_t cs_batch, int inverse, void* stream) {
  if (!x || !cosp || !sinp) return bad_arg("rope: null");
  if (tokens < 0 || seq <= 0 || heads <= 0 || hd <= 0 || (hd & 7)) return bad_arg("rope: head_dim must be a multiple of 8");
  if (tokens % seq) return bad_arg("rope: tokens must be a multiple of seq");
  if (!aligned16(x) || !aligned16(cosp) || !aligned16(sinp) || (cs_batch & 3)) return bad_arg("rope: alignment");
  if (tokens == 0) return 0;
  const int64_t total = tokens * heads * (hd / 8);
  int64_t blocks = ceil_div(total, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  GRASP_LAUNCH(rope_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, x, tokens, (int)seq, (int)heads, (int)hd, cosp,
               sinp, cs_batch, inverse);
  GRASP_CHECK_LAST("rope_kernel");
  return 0;
}

static unsigned elementwise_blocks(int64_t n) {
  int64_t blocks = ceil_div(ceil_div(n, 4), 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

extern "C" int grasp_swiglu_fwd(const float* g, const float* u, int64_t rows, int64_t cols, float* h, void* planes,
                                float* inv, void* stream) {
  if (!g || !u) return bad_arg("swiglu_fwd: null");
  if (!h && !planes) return bad_arg("swiglu_fwd: neither h nor planes requested");
  if (planes && !inv) return bad_arg("swiglu_fwd: planes need inv");
  if (rows < 0 || cols <= 0) return bad_arg("swiglu_fwd: rows/cols");
  if (!aligned16(g) || !aligned16(u) || (h && !aligned16(h))) return bad_arg("swiglu_fwd: pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  const int64_t n = rows * cols;
  if (!planes) {
    GRASP_LAUNCH(swiglu_fwd_kernel, dim3(elementwise_blocks(n)), dim3(256), 0, stream, g, u, n, h);
    GRASP_CHECK_LAST("swiglu_fwd_kernel");
    return 0;
  }
  if (reinterpret_cast<uintptr_t>(planes) & 1023) return bad_arg("swiglu_fwd: planes must be 1024-byte aligned");
  const int64_t cp = plane_pitch(cols);
  if (cp * 4 > 200 * 1024) return bad_arg("swiglu_fwd: row too long for the fused operand form");
  static size_t granted = 48 * 1024;
  int rc = row_smem_attr((const void*)swiglu_fwd_rows_kernel, (size_t)cp * 4, granted, "swiglu_fwd attr");
  if (rc) return rc;
  GRASP_LAUNCH(swiglu_fwd_rows_kernel, dim3((unsigned)rows), dim3(ROW_THREADS), (size_t)cp * 4, stream, g, u, (int)cols,
               (int)cp, h, static_cast<uint16_t*>(planes), inv, rows);
  GRASP_CHECK_LAST("swiglu_fwd_rows_kernel");
  return 0;
}

extern "C" int grasp_swiglu_bwd(const float* dh, const float* g, const float* u, int64_t rows, int64_t cols, float* dg,
                                float* du, void* dg_planes, float* dg_inv, void* du_planes, float* du_inv, void* stream) {
  if (!dh || !g || !u) return bad_arg("swiglu_bwd: null");
  const bool want_planes = dg_planes || du_planes;
  if (want_planes && !(dg_planes && du_planes && dg_inv && du_inv)) return bad_arg("swiglu_bwd: both operand forms or none");
  if (!want_planes && !(dg && du)) return bad_arg("swiglu_bwd: nothing requested");
  if ((dg == nullptr) != (du == nullptr)) return bad_arg("swiglu_bwd: dg and du come together");
  if (rows < 0 || cols <= 0) return bad_arg("swiglu_bwd: rows/cols");
  if (!aligned16(dh) || !aligned16(g) || !aligned16(u) || (dg && (!aligned16(dg) || !aligned16(du))))
    return bad_arg("swiglu_bwd: pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  const int64_t n = rows * cols;
  if (!want_planes) {
    GRASP_LAUNCH(swiglu_bwd_kernel, dim3(elementwise_blocks(n)), dim3(256), 0, stream, dh, g, u, n, dg, du);
    GRASP_CHECK_LAST("swiglu_bwd_kernel");
    return 0;
  }
  if ((reinterpret_cast<uintptr_t>(dg_planes) & 1023) || (reinterpret_cast<uintptr_t>(du_planes) & 1023))
    return bad_arg("swiglu_bwd: planes must be 1024-byte aligned");
  const int64_t cp = plane_pitch(cols);
  if (cp * 8 > 200 * 1024) return bad_arg("swiglu_bwd: row too long for the fused operand form");
  static size_t granted = 48 * 1024;
  int rc = row_smem_attr((const void*)swiglu_bwd_rows_kernel, (size_t)cp * 8, granted, "swiglu_bwd attr");
  if (rc) return rc;
  GRASP_LAUNCH(swiglu_bwd_rows_kernel, dim3((unsigned)rows), dim3(ROW_THREADS), (size_t)cp * 8, stream, dh, g, u, (int)cols,
               (int)cp, dg, du, static_cast<uint16_t*>(dg_planes), dg_inv, static_cast<uint16_t*>(du_planes), du_inv, rows);
  GRASP_CHECK_LAST("swiglu_bwd_rows_kernel");
  return 0;
}

extern "C" int grasp_ce_loss_bwd(float* logits, const int64_t* labels, const float* coef, int64_t rows, int64_t V,
                                 float* loss, void* stream) {
  if (!logits || !labels || !coef || !loss) return bad_arg("ce_loss: null");
  if (rows < 0 || V <= 0 || V >= (1 << 30)) return bad_arg("ce_loss: rows/V");
  if (!aligned16(logits)) return bad_arg("ce_loss: logits must be 16-byte aligned");
  if (rows == 0) return 0;
  GRASP_LAUNCH(ce_loss_bwd_kernel, dim3((unsigned)rows), dim3(CE_THREADS), 0, stream, logits, labels, coef, (int)V, loss);
  GRASP_CHECK_LAST("ce_loss_bwd_kernel");
  return 0;
}
