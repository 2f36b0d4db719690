// fp16 (hi, lo) planes of the F16X3 GEMM arithmetic: helpers shared by the operand pre-pass (gemm_tc.cu) and
// by the row kernels that emit their result directly as a GEMM operand (layer_ops.cu).
#pragma once
#include "common.cuh"

namespace grasp {

// x*s = hi + lo with s a power of two that puts the maximum in [2^14, 2^15)
// (fp16 keeps 11 bits, two planes 22; the scale keeps every row inside fp16's range).
__device__ __forceinline__ void scale_from_max(float m, float& s, float& inv) {
  const int ef = (int)((__float_as_uint(m) >> 23) & 0xffu);     // biased exponent of the row maximum
  if (ef == 0 || ef == 0xff) { s = 1.f; inv = 1.f; return; }    // zero / denormal / inf / nan row: leave as is
  int e = 14 - (ef - 127);                                      // s = 2^e
  e = e > 100 ? 100 : (e < -100 ? -100 : e);
  s = __uint_as_float((uint32_t)(127 + e) << 23);
  inv = __uint_as_float((uint32_t)(127 - e) << 23);
}


__device__ __forceinline__ void split_f16(float x, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(x);
  const __half l = __float2half_rn(x - __half2float(h));
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(l);
}


// A 256-thread CTA holds one row of Kp floats (Kp % 8 == 0, zero padded) in shared memory and every thread its
// partial maximum `m`: find the row scale, write inv[0] and the two planes (16-byte stores).  red: 9 floats.
__device__ __forceinline__ void block_row_split_256(const float* rowbuf, float m, int Kp, uint16_t* __restrict__ d0,
                                                    uint16_t* __restrict__ d1, float* __restrict__ inv, float* red) {
  m = warp_max(m);
  __syncthreads();                                  // rowbuf complete; red free
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
    v = warp_max(v);
    if (threadIdx.x == 0) {
      float s, i;
      scale_from_max(v, s, i);
      red[8] = s;
      *inv = i;
    }
  }
  __syncthreads();
  const float s = red[8];
  for (int k = threadIdx.x * 8; k < Kp; k += 256 * 8) {
    uint16_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_f16(rowbuf[k + j] * s, hi[j], lo[j]);
    uint4 w;
    w.x = hi[0] | ((uint32_t)hi[1] << 16); w.y = hi[2] | ((uint32_t)hi[3] << 16);
    w.z = hi[4] | ((uint32_t)hi[5] << 16); w.w = hi[6] | ((uint32_t)hi[7] << 16);
    *reinterpret_cast<uint4*>(d0 + k) = w;
    w.x = lo[0] | ((uint32_t)lo[1] << 16); w.y = lo[2] | ((uint32_t)lo[3] << 16);
    w.z = lo[4] | ((uint32_t)lo[5] << 16); w.w = lo[6] | ((uint32_t)lo[7] << 16);
    *reinterpret_cast<uint4*>(d1 + k) = w;
  }
}

}  // namespace grasp
