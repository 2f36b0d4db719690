// tcgen05 / TMA split-bf16 GEMM paths (GRASP_PREC_BF16X3 / BF16X6).
// Placeholder until the tensor-core kernels are validated on hardware: every
// entry point reports "not available" so callers fail loudly instead of
// silently computing on another path.
#include "common.cuh"

namespace grasp {

size_t tc_gemm_workspace_bytes(int64_t, int64_t, int64_t, int) { return 0; }
size_t tc_sigma_workspace_bytes(int64_t, int64_t, int64_t, int) { return 0; }

int tc_gemm_f32(int, int, int64_t, int64_t, int64_t, float, const float*, int64_t, const float*, int64_t, float,
                void*, int64_t, int, int, void*, size_t, void*) {
  set_error("tensor-core GEMM path not built in this revision");
  return -2;
}

int tc_sigma_partials(const float*, const float*, const float*, int64_t, int64_t, int64_t, int, float*, int64_t*,
                      void*, size_t, void*) {
  set_error("tensor-core sigma-score path not built in this revision");
  return -2;
}

}  // namespace grasp
