// Block-influence (layer redundancy) scoring: 1 - cos(h_in, h_out) per token,
// mean over tokens, accumulated per layer pair.
// Behaviour follows reference tools/utils_func.py:3-25 (block_influence) as used by
// modeling_grasp.py:135-193 (compute_bi); the reference forms a full [N,N] Gram to read
// its diagonal -- here each token row is one CTA-wide dot product and every hidden
// state is streamed from HBM exactly once per forward pass (chain kernel).
#include "common.cuh"

namespace grasp {

constexpr int BI_THREADS = 256;
constexpr int BI_MAXV = 8;  // 16-byte vectors cached per thread per row

template <typename T> struct Vec16;  // 16 bytes of T -> floats
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static void unpack(const uint4& r, float* f) {
    f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y);
    f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void unpack(const uint4& r, float* f) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <> struct Vec16<__half> {
  static constexpr int N = 8;
  __device__ static void unpack(const uint4& r, float* f) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      float2 p = __half22float2(h);
      f[2 * i] = p.x; f[2 * i + 1] = p.y;
    }
  }
};

// block-wide sum of two floats; result valid in every thread
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red /*[2*32]*/) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // protect red from the previous use
  if (l == 0) { red[w] = a; red[32 + w] = b; }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float x = (l < nw) ? red[l] : 0.f;
  float y = (l < nw) ? red[32 + l] : 0.f;
  a = warp_sum(x);
  b = warp_sum(y);
}

__device__ __forceinline__ float bi_from(float dot, float nx2, float ny2, int angular) {
  // reference: sim = dot / (|x| * |y|); nan_to_num(nan=0.5); 1 - sim (or arccos(sim)/pi)
  float sim = dot / (sqrtf(nx2) * sqrtf(ny2));
  if (sim != sim) sim = 0.5f;
  if (angular) return acosf(sim) * 0.31830988618379067154f;
  return 1.f - sim;
}

struct ChainPtrs {
  const void* h[130];
};

// one CTA per token row; walks the L+1 hidden states keeping the previous row in registers
template <typename T>
__global__ void __launch_bounds__(BI_THREADS)
bi_chain_kernel(ChainPtrs ptrs, int n_states, int64_t rows, int64_t d, int64_t ld, double scale, double* acc) {
  __shared__ float red[64];
  using V = Vec16<T>;
  const int64_t row = blockIdx.x;
  const int nvec = (int)(d / V::N);  // host guarantees d % V::N == 0 and nvec <= BI_THREADS*BI_MAXV
  const int tid = threadIdx.x;

  uint4 prev[BI_MAXV], cur[BI_MAXV];
  auto load_row = [&](int s, uint4* dst) {
    const uint4* p = reinterpret_cast<const uint4*>(static_cast<const T*>(ptrs.h[s]) + row * ld);
#pragma unroll
    for (int j = 0; j < BI_MAXV; ++j) {
      const int v = tid + j * BI_THREADS;
      dst[j] = (v < nvec) ? ldg_stream(p + v) : make_uint4(0, 0, 0, 0);
    }
  };

  load_row(0, prev);
  float dummy = 0.f, nprev = 0.f;
#pragma unroll
  for (int j = 0; j < BI_MAXV; ++j) {
    float f[V::N];
    V::unpack(prev[j], f);
#pragma unroll
    for (int e = 0; e < V::N; ++e) nprev = fmaf(f[e], f[e], nprev);
  }
  if (n_states > 1) load_row(1, cur);
  block_sum2(nprev, dummy, red);

  const double inv_rows = scale / (double)rows;
  for (int s = 1; s < n_states; ++s) {
    float dot = 0.f, ncur = 0.f;
#pragma unroll
    for (int j = 0; j < BI_MAXV; ++j) {
      float a[V::N], b[V::N];
      V::unpack(prev[j], a);
      V::unpack(cur[j], b);
#pragma unroll
      for (int e = 0; e < V::N; ++e) {
        dot = fmaf(a[e], b[e], dot);
        ncur = fmaf(b[e], b[e], ncur);
      }
      prev[j] = cur[j];
    }
    // issue the next state's loads before the reduction so HBM latency overlaps it
    if (s + 1 < n_states) load_row(s + 1, cur);
    block_sum2(dot, ncur, red);
    if (tid == 0) atomicAdd(&acc[s - 1], (double)bi_from(dot, nprev, ncur, 0) * inv_rows);
    nprev = ncur;
  }
}

// generic pair kernel (any d, optional per-row output, angular variant)
template <typename T>
__global__ void __launch_bounds__(BI_THREADS)
bi_pair_kernel(const T* __restrict__ x, const T* __restrict__ y, int64_t rows, int64_t d, int64_t ld,
               int angular, double scale, double* acc, float* per_row) {
  __shared__ float red[64];
  using V = Vec16<T>;
  const int64_t row = blockIdx.x;
  const T* xr = x + row * ld;
  const T* yr = y + row * ld;
  float dot = 0.f, nx = 0.f, ny = 0.f;
  const bool vec_ok = (d % V::N == 0) && (ld % V::N == 0) &&
                      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (vec_ok) {
    const int nvec = (int)(d / V::N);
    const uint4* xp = reinterpret_cast<const uint4*>(xr);
    const uint4* yp = reinterpret_cast<const uint4*>(yr);
    for (int v = threadIdx.x; v < nvec; v += BI_THREADS) {
      float a[V::N], b[V::N];
      V::unpack(ldg_stream(xp + v), a);
      V::unpack(ldg_stream(yp + v), b);
#pragma unroll
      for (int e = 0; e < V::N; ++e) {
        dot = fmaf(a[e], b[e], dot);
        nx = fmaf(a[e], a[e], nx);
        ny = fmaf(b[e], b[e], ny);
      }
    }
  } else {
    for (int64_t i = threadIdx.x; i < d; i += BI_THREADS) {
      const float a = (float)xr[i], b = (float)yr[i];
      dot = fmaf(a, b, dot);
      nx = fmaf(a, a, nx);
      ny = fmaf(b, b, ny);
    }
  }
  float z = 0.f;
  block_sum2(dot, nx, red);
  block_sum2(ny, z, red);
  if (threadIdx.x == 0) {
    const float bi = bi_from(dot, nx, ny, angular);
    if (per_row) per_row[row] = bi;
    if (acc) atomicAdd(acc, (double)bi * scale / (double)rows);
  }
}

template <typename T>
static int launch_pair(const void* x, const void* y, int64_t rows, int64_t d, int64_t ld, int angular,
                       double scale, double* acc, float* per_row, void* stream) {
  GRASP_LAUNCH((bi_pair_kernel<T>), dim3((unsigned)rows), dim3(BI_THREADS), 0, stream,
               static_cast<const T*>(x), static_cast<const T*>(y), rows, d, ld, angular, scale, acc, per_row);
  GRASP_CHECK_LAST("bi_pair_kernel");
  return 0;
}

}  // namespace grasp

using namespace grasp;

extern "C" int grasp_bi_accumulate(const void* h_in, const void* h_out, int64_t rows, int64_t d,
                                   int64_t ld, int dtype, int angular, double scale, double* acc,
                                   float* per_row, void* stream) {
  if (rows < 0 || d <= 0 || ld < d) return bad_arg("bi: rows/d/ld");
  if (!acc && !per_row) return bad_arg("bi: no output");
  if (rows == 0) return 0;
  if (!h_in || !h_out) return bad_arg("bi: null hidden state");
  if (rows > 0x7fffffffLL) return bad_arg("bi: rows too large");
  switch (dtype) {
    case GRASP_DTYPE_F32: return launch_pair<float>(h_in, h_out, rows, d, ld, angular, scale, acc, per_row, stream);
    case GRASP_DTYPE_BF16: return launch_pair<__nv_bfloat16>(h_in, h_out, rows, d, ld, angular, scale, acc, per_row, stream);
    case GRASP_DTYPE_F16: return launch_pair<__half>(h_in, h_out, rows, d, ld, angular, scale, acc, per_row, stream);
  }
  return bad_arg("bi: dtype");
}

extern "C" int grasp_bi_chain(const void* const* hiddens, int n_states, int64_t rows, int64_t d,
                              int64_t ld, int dtype, double scale, double* acc, void* stream) {
  if (!hiddens || !acc) return bad_arg("bi_chain: null");
  if (n_states < 2 || n_states > 130) return bad_arg("bi_chain: n_states must be in [2,130]");
  if (rows < 0 || d <= 0 || ld < d) return bad_arg("bi_chain: rows/d/ld");
  if (rows == 0) return 0;
  if (rows > 0x7fffffffLL) return bad_arg("bi_chain: rows too large");
  const int vecn = (dtype == GRASP_DTYPE_F32) ? 4 : 8;
  bool fast = (d % vecn == 0) && (ld % vecn == 0) && (d / vecn <= (int64_t)BI_THREADS * BI_MAXV);
  for (int i = 0; i < n_states; ++i) {
    if (!hiddens[i]) return bad_arg("bi_chain: null hidden state");
    if (reinterpret_cast<uintptr_t>(hiddens[i]) & 15) fast = false;
  }
  if (!fast) {  // rows too wide for the register cache or unaligned: one pair launch per layer
    for (int i = 0; i + 1 < n_states; ++i) {
      int rc = grasp_bi_accumulate(hiddens[i], hiddens[i + 1], rows, d, ld, dtype, 0, scale, acc + i, nullptr, stream);
      if (rc) return rc;
    }
    return 0;
  }
  ChainPtrs p;
  for (int i = 0; i < n_states; ++i) p.h[i] = hiddens[i];
  switch (dtype) {
    case GRASP_DTYPE_F32:
      GRASP_LAUNCH((bi_chain_kernel<float>), dim3((unsigned)rows), dim3(BI_THREADS), 0, stream, p, n_states, rows, d, ld, scale, acc);
      break;
    case GRASP_DTYPE_BF16:
      GRASP_LAUNCH((bi_chain_kernel<__nv_bfloat16>), dim3((unsigned)rows), dim3(BI_THREADS), 0, stream, p, n_states, rows, d, ld, scale, acc);
      break;
    case GRASP_DTYPE_F16:
      GRASP_LAUNCH((bi_chain_kernel<__half>), dim3((unsigned)rows), dim3(BI_THREADS), 0, stream, p, n_states, rows, d, ld, scale, acc);
      break;
    default:
      return bad_arg("bi_chain: dtype");
  }
  GRASP_CHECK_LAST("bi_chain_kernel");
  return 0;
}
