"""ctypes binding of libgrasp_b200.so -- mirrors include/grasp_b200.h one to one."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB = None
_LIB_PATH = Path(__file__).resolve().parent / "libgrasp_b200.so"

PREC_SIMT, PREC_BF16X3, PREC_BF16X6, PREC_F16X3 = 0, 3, 6, 16
SVD_NO_PRECOND = 0x100          # flag for grasp_svd_batched's prec (include/grasp_b200.h)
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
METRIC_GRADIENT, METRIC_TAYLOR = 0, 1
SCALE_ROWS, SCALE_TENSOR = 0, 2

_i64p = C.POINTER(C.c_int64)
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); keep in the order of include/grasp_b200.h
SIGNATURES = {
    "grasp_abi_version": (C.c_int, []),
    "grasp_last_error": (C.c_char_p, []),
    "grasp_launch_count": (C.c_uint64, []),
    "grasp_bi_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                      C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grasp_bi_chain": (C.c_int, [_vpp, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_void_p,
                                 C.c_void_p]),
    "grasp_svd_workspace_bytes": (C.c_size_t, [C.c_int, _i64p, _i64p]),
    "grasp_svd_batched": (C.c_int, [C.c_int, _vpp, _i64p, _i64p, _i64p, _vpp, _vpp, _vpp, C.c_void_p, C.c_int,
                                    C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "grasp_sigma_score_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "grasp_sigma_score": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                    C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                    C.c_void_p]),
    "grasp_score_from_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "grasp_topk_batched": (C.c_int, [C.c_int, _vpp, _i64p, _i64p, _vpp, C.c_void_p]),
    "grasp_adaptive_rank": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grasp_lowrank_rebuild_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "grasp_lowrank_rebuild": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                        C.c_void_p]),
    "grasp_factor_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                    C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grasp_gemm_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "grasp_gemm_f32": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                 C.c_size_t, C.c_void_p]),
    "grasp_gemm_planes_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "grasp_gemm_split_f16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "grasp_gemm_f16x3_planes": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]),
    "grasp_gemm_f16x3_planes_out": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grasp_rmsnorm_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "grasp_rmsnorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_void_p]),
    "grasp_rope_inplace": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_int, C.c_void_p]),
    "grasp_swiglu_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "grasp_swiglu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grasp_ce_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "grasp_attn_prep_qkv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grasp_attn_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "grasp_attn_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}


class GraspLibraryError(RuntimeError):
    pass


def lib_path() -> Path:
    return Path(os.environ.get("GRASP_B200_LIB", str(_LIB_PATH)))


def load():
    """Load the shared library (built by grasp_b200.build / __graft_entry__.build). Fails loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not path.exists():
        raise GraspLibraryError(
            f"{path} is missing: run `python -m grasp_b200.build` (or __graft_entry__.build()). "
            "grasp_b200 has no CPU/library fallback."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and this table drifted
        fn.restype = res
        fn.argtypes = args
    if lib.grasp_abi_version() != 1:
        raise GraspLibraryError(f"ABI version mismatch: {lib.grasp_abi_version()}")
    _LIB = lib
    return lib


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = load().grasp_last_error().decode(errors="replace")
    kind = "bad argument" if rc < 0 else "CUDA error"
    raise GraspLibraryError(f"{what}: {kind} ({rc}): {msg}")


def i64_array(vals):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


def ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])
