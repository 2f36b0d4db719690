// Shared helpers for the grasp_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/grasp_b200.h"

namespace grasp {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int bad_arg(const char* what) {
  set_error("bad argument: %s", what);
  return -1;
}

inline int check_cuda(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", where, cudaGetErrorString(e));
  return (int)e;
}

// every kernel launch in the library goes through this macro so that
// grasp_launch_count() is an honest count.
#define GRASP_LAUNCH(kernel, grid, block, smem, stream, ...)                          \
  do {                                                                                 \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);          \
    ::grasp::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
  } while (0)

#define GRASP_CHECK_LAST(where)                                                        \
  do {                                                                                 \
    int _rc = ::grasp::check_cuda(cudaGetLastError(), where);                          \
    if (_rc) return _rc;                                                               \
  } while (0)

// ---- device helpers ---------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 16-byte load that does not pollute L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }
// Row pitch (elements) of a 16-bit operand plane: whole 128-byte lines, so that the 64-element row segments the TMA
// boxes fetch never straddle a line (at the rank-k widths 204 / 298 a pitch rounded to 8 made every segment touch two)
inline __host__ __device__ int64_t plane_pitch(int64_t cols) { return (cols + 63) / 64 * 64; }

// number of SMs of the current device (cached)
int sm_count();

}  // namespace grasp
