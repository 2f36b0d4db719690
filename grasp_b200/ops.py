"""torch-tensor wrappers over the C ABI (include/grasp_b200.h).

PyTorch is only the memory/stream provider here: every function hands raw device
pointers and the current CUDA stream to libgrasp_b200.so.  Nothing falls back to
torch.linalg / torch.topk / CPU -- a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (DTYPE_BF16, DTYPE_F16, DTYPE_F32, METRIC_GRADIENT, METRIC_TAYLOR, PREC_BF16X3, PREC_BF16X6,
                   PREC_F16X3, PREC_SIMT, GraspLibraryError, check)

# arithmetic of the GEMM-shaped stages; see include/grasp_b200.h
_DEFAULT_PREC = PREC_F16X3


def set_default_precision(prec: int) -> None:
    global _DEFAULT_PREC
    if prec not in (PREC_SIMT, PREC_BF16X3, PREC_BF16X6, PREC_F16X3):
        raise ValueError(f"unknown precision {prec}")
    _DEFAULT_PREC = prec


def default_precision() -> int:
    return _DEFAULT_PREC


def _prec(prec: Optional[int]) -> int:
    return _DEFAULT_PREC if prec is None else prec


_DTYPES = {torch.float32: DTYPE_F32, torch.bfloat16: DTYPE_BF16, torch.float16: DTYPE_F16}


def _need_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise GraspLibraryError("grasp_b200 ops need CUDA tensors (there is no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise GraspLibraryError("tensors live on different devices")
    return dev


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _workspace(nbytes: int, device) -> torch.Tensor:
    # torch's caching allocator hands out 512-byte aligned blocks
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def launch_count() -> int:
    return int(_lib.load().grasp_launch_count())


class _Timers:
    """Optional CUDA-event timing of the C-ABI calls (bench.py's roofline numbers).  Off by default."""

    def __init__(self):
        self.enabled = False
        self.spans = {}

    def reset(self, enabled: bool = False):
        self.enabled = enabled
        self.spans = {}

    def start(self):
        if not self.enabled:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def stop(self, tag: str, e0, flops: float = 0.0, bytes_: float = 0.0, mma_per_flop: float = 1.0):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        rec = self.spans.setdefault(tag, {"events": [], "flops": 0.0, "bytes": 0.0, "mma_per_flop": mma_per_flop})
        rec["events"].append((e0, e1))
        rec["flops"] += flops
        rec["bytes"] += bytes_

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for tag, rec in self.spans.items():
            ms = sum(a.elapsed_time(b) for a, b in rec["events"])
            out[tag] = {"ms": ms, "n": len(rec["events"]), "flops": rec["flops"], "bytes": rec["bytes"],
                        "mma_per_flop": rec["mma_per_flop"]}
        return out


timers = _Timers()


def _mma_per_flop(prec: int) -> float:
    return {PREC_SIMT: 0.0, PREC_BF16X3: 3.0, PREC_BF16X6: 6.0, PREC_F16X3: 3.0}[prec]


# --------------------------------------------------------------------------- BI
def bi_accumulate(h_in: torch.Tensor, h_out: torch.Tensor, acc: Optional[torch.Tensor] = None,
                  angular: bool = False, per_row: bool = False, scale: float = 1.0) -> Optional[torch.Tensor]:
    """acc[0] += mean_t BI(h_in[t], h_out[t]); optionally returns the per-token BI (fp32 [rows])."""
    lib = _lib.load()
    dev = _need_cuda(h_in, h_out, acc)
    if h_in.shape != h_out.shape or h_in.dtype != h_out.dtype:
        raise ValueError("hidden states must have identical shape and dtype")
    if h_in.dtype not in _DTYPES:
        raise TypeError(f"unsupported hidden-state dtype {h_in.dtype}")
    d = h_in.shape[-1]
    x = h_in.reshape(-1, d).contiguous()
    y = h_out.reshape(-1, d).contiguous()
    rows = x.shape[0]
    out = torch.empty(rows, dtype=torch.float32, device=dev) if per_row else None
    if acc is not None and (acc.dtype != torch.float64 or acc.numel() < 1):
        raise TypeError("acc must be a float64 device tensor")
    if acc is None and out is None:
        raise ValueError("nothing to compute: pass acc and/or per_row=True")
    if rows == 0:
        return out
    with torch.cuda.device(dev):
        check(lib.grasp_bi_accumulate(x.data_ptr(), y.data_ptr(), rows, d, d, _DTYPES[x.dtype], int(bool(angular)),
                                      float(scale), acc.data_ptr() if acc is not None else None,
                                      out.data_ptr() if out is not None else None, _stream()), "grasp_bi_accumulate")
    return out


def bi_chain(hiddens: Sequence[torch.Tensor], acc: torch.Tensor, scale: float = 1.0) -> None:
    """acc[i] += mean_t BI(hiddens[i], hiddens[i+1]) for the L+1 hidden states of one forward pass."""
    lib = _lib.load()
    dev = _need_cuda(*hiddens, acc)
    n = len(hiddens)
    if n < 2:
        raise ValueError("need at least two hidden states")
    d = hiddens[0].shape[-1]
    hs = []
    for h in hiddens:
        if h.shape != hiddens[0].shape or h.dtype != hiddens[0].dtype:
            raise ValueError("hidden states must share shape and dtype")
        hs.append(h.reshape(-1, d).contiguous())
    if hs[0].dtype not in _DTYPES:
        raise TypeError(f"unsupported hidden-state dtype {hs[0].dtype}")
    if acc.dtype != torch.float64 or acc.numel() < n - 1 or not acc.is_contiguous():
        raise TypeError("acc must be a contiguous float64 device tensor with >= len(hiddens)-1 entries")
    rows = hs[0].shape[0]
    ptrs = _lib.ptr_array([h.data_ptr() for h in hs])
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_bi_chain(ptrs, n, rows, d, d, _DTYPES[hs[0].dtype], float(scale), acc.data_ptr(), _stream()),
              "grasp_bi_chain")
        timers.stop("grasp_bi_chain", t0, bytes_=2.0 * (n - 1) * rows * d * hs[0].element_size())


# -------------------------------------------------------------------------- SVD
def svd_batched(mats: Sequence[torch.Tensor], prec: Optional[int] = None, max_sweeps: int = 0,
                return_info: bool = False, precondition: bool = True):
    """Thin SVD of each fp32 matrix: returns [(U [m,r], S [r] descending, Vh [r,n])].
    precondition=False factors wide / tall matrices as they are (no CholeskyQR2 reduction to a square factor)."""
    lib = _lib.load()
    dev = _need_cuda(*mats)
    if len(mats) == 0:
        return ([], None) if return_info else []
    As = []
    for w in mats:
        if w.dim() != 2:
            raise ValueError("svd expects 2-D matrices")
        As.append(_f32c(w, "matrix"))
    m = [a.shape[0] for a in As]
    n = [a.shape[1] for a in As]
    r = [min(a.shape) for a in As]
    Us = [torch.empty(mi, ri, dtype=torch.float32, device=dev) for mi, ri in zip(m, r)]
    Ss = [torch.empty(ri, dtype=torch.float32, device=dev) for ri in r]
    Vs = [torch.empty(ri, ni, dtype=torch.float32, device=dev) for ri, ni in zip(r, n)]
    info = torch.zeros(4 * len(As), dtype=torch.int32, device=dev)
    m_a, n_a = _lib.i64_array(m), _lib.i64_array(n)
    nbytes = lib.grasp_svd_workspace_bytes(len(As), m_a, n_a)
    if nbytes == 0:
        raise GraspLibraryError("grasp_svd_workspace_bytes rejected the shapes")
    ws = _workspace(nbytes, dev)
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_svd_batched(len(As), _lib.ptr_array([a.data_ptr() for a in As]), m_a, n_a, _lib.i64_array(n),
                                    _lib.ptr_array([u.data_ptr() for u in Us]),
                                    _lib.ptr_array([s.data_ptr() for s in Ss]),
                                    _lib.ptr_array([v.data_ptr() for v in Vs]), info.data_ptr(),
                                    _prec(prec) | (0 if precondition else _lib.SVD_NO_PRECOND), int(max_sweeps), ws.data_ptr(), ws.numel(), _stream()), "grasp_svd_batched")
        timers.stop("grasp_svd_batched", t0,
                    flops=sum(8.0 * max(a, b) * min(a, b) ** 2 + 4.0 / 3.0 * min(a, b) ** 3 for a, b in zip(m, n)),
                    mma_per_flop=0.0)
    # keep the workspace alive until the stream has consumed it
    ws.record_stream(torch.cuda.current_stream())
    out = list(zip(Us, Ss, Vs))
    return (out, info.view(-1, 4)) if return_info else out


def svd(w: torch.Tensor, prec: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Drop-in for torch.linalg.svd(w, full_matrices=False) on a CUDA fp32 matrix."""
    return svd_batched([w], prec=prec)[0]


# ---------------------------------------------------------------------- scoring
def _metric(metric) -> int:
    if metric in (METRIC_GRADIENT, "gradient"):
        return METRIC_GRADIENT
    if metric in (METRIC_TAYLOR, "taylor"):
        return METRIC_TAYLOR
    raise RuntimeError(f"{metric} not support")


def sigma_score(U: torch.Tensor, G: torch.Tensor, Vh: torch.Tensor, S: Optional[torch.Tensor], metric="taylor",
                dsigma: Optional[torch.Tensor] = None, want_score: bool = True, prec: Optional[int] = None):
    """dsigma[i] (+)= u_i^T G v_i ; score = |dsigma| (gradient) or |dsigma*S| (taylor).

    Pass an existing `dsigma` to accumulate into it.  Returns (dsigma, score or None)."""
    lib = _lib.load()
    dev = _need_cuda(U, G, Vh, S, dsigma)
    U, G, Vh = _f32c(U, "U"), _f32c(G, "G"), _f32c(Vh, "Vh")
    out, r = U.shape
    in_ = Vh.shape[1]
    if G.shape != (out, in_) or Vh.shape[0] != r:
        raise ValueError(f"shape mismatch: U {tuple(U.shape)} G {tuple(G.shape)} Vh {tuple(Vh.shape)}")
    if S is not None:
        S = _f32c(S, "S")
    accumulate = dsigma is not None
    if dsigma is None:
        dsigma = torch.empty(r, dtype=torch.float32, device=dev)
    elif dsigma.dtype != torch.float32 or not dsigma.is_contiguous() or dsigma.numel() != r:
        raise TypeError("dsigma must be a contiguous float32 tensor of length r")
    score = torch.empty(r, dtype=torch.float32, device=dev) if want_score else None
    p = _prec(prec)
    ws = _workspace(lib.grasp_sigma_score_workspace_bytes(out, in_, r, p), dev)
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_sigma_score(U.data_ptr(), G.data_ptr(), Vh.data_ptr(), S.data_ptr() if S is not None else None,
                                    out, in_, r, _metric(metric), int(accumulate), dsigma.data_ptr(),
                                    score.data_ptr() if score is not None else None, p, ws.data_ptr(), ws.numel(),
                                    _stream()), "grasp_sigma_score")
        timers.stop("grasp_sigma_score", t0, flops=2.0 * r * out * in_ + 2.0 * r * in_, mma_per_flop=_mma_per_flop(p))
    ws.record_stream(torch.cuda.current_stream())
    return dsigma, score


def score_from_grad(dsigma: torch.Tensor, S: Optional[torch.Tensor], metric="taylor") -> torch.Tensor:
    lib = _lib.load()
    dev = _need_cuda(dsigma, S)
    g = _f32c(dsigma, "dsigma")
    s = _f32c(S, "S") if S is not None else None
    score = torch.empty_like(g)
    with torch.cuda.device(dev):
        check(lib.grasp_score_from_grad(g.data_ptr(), s.data_ptr() if s is not None else None, g.numel(),
                                        _metric(metric), score.data_ptr(), _stream()), "grasp_score_from_grad")
    return score


def topk_batched(scores: Sequence[torch.Tensor], ks: Sequence[int]) -> List[torch.Tensor]:
    """Indices of the k largest scores of each vector, sorted by score descending (int64)."""
    lib = _lib.load()
    dev = _need_cuda(*scores)
    if len(scores) != len(ks):
        raise ValueError("scores and ks differ in length")
    if not scores:
        return []
    sc = [_f32c(s.reshape(-1), "score") for s in scores]
    for s, k in zip(sc, ks):
        if not 0 <= int(k) <= s.numel():
            raise ValueError(f"k={k} out of range for {s.numel()} scores")
    outs = [torch.empty(int(k), dtype=torch.int64, device=dev) for k in ks]
    live = [i for i, k in enumerate(ks) if int(k) > 0]
    if live:
        with torch.cuda.device(dev):
            check(lib.grasp_topk_batched(len(live), _lib.ptr_array([sc[i].data_ptr() for i in live]),
                                         _lib.i64_array([sc[i].numel() for i in live]),
                                         _lib.i64_array([ks[i] for i in live]),
                                         _lib.ptr_array([outs[i].data_ptr() for i in live]), _stream()),
                  "grasp_topk_batched")
    return outs


def topk(score: torch.Tensor, k: int) -> torch.Tensor:
    return topk_batched([score], [k])[0]


def adaptive_rank(score: torch.Tensor, target_ratio: float) -> torch.Tensor:
    """tools/utils_func.py:45-57 on the device; returns the kept indices (descending score)."""
    lib = _lib.load()
    dev = _need_cuda(score)
    s = _f32c(score.reshape(-1), "score")
    idx = torch.empty(s.numel(), dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.grasp_adaptive_rank(s.data_ptr(), s.numel(), float(target_ratio), idx.data_ptr(), count.data_ptr(),
                                      _stream()), "grasp_adaptive_rank")
    return idx[: int(count.item())]


# ---------------------------------------------------------------------- compile
def lowrank_rebuild(U: torch.Tensor, S: torch.Tensor, Vh: torch.Tensor, idx: torch.Tensor,
                    out_dtype: torch.dtype = torch.float32, prec: Optional[int] = None) -> torch.Tensor:
    """W = U[:, idx] diag(S[idx]) Vh[idx, :]."""
    lib = _lib.load()
    dev = _need_cuda(U, S, Vh, idx)
    U, S, Vh = _f32c(U, "U"), _f32c(S, "S"), _f32c(Vh, "Vh")
    if idx.dtype != torch.int64:
        raise TypeError("idx must be int64")
    idx = idx.contiguous()
    out, r = U.shape
    in_ = Vh.shape[1]
    k = idx.numel()
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("out_dtype must be float32 or bfloat16")
    W = torch.empty(out, in_, dtype=out_dtype, device=dev)
    p = _prec(prec)
    ws = _workspace(lib.grasp_lowrank_rebuild_workspace_bytes(out, in_, k, p), dev)
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_lowrank_rebuild(U.data_ptr(), S.data_ptr(), Vh.data_ptr(), idx.data_ptr(), k, out, in_, r,
                                        _DTYPES[out_dtype], W.data_ptr(), p, ws.data_ptr(), ws.numel(), _stream()),
              "grasp_lowrank_rebuild")
        timers.stop("grasp_lowrank_rebuild", t0, flops=2.0 * out * in_ * k,
                    bytes_=4.0 * k * (out + in_) + W.element_size() * out * in_, mma_per_flop=_mma_per_flop(p))
    ws.record_stream(torch.cuda.current_stream())
    return W


def factor_pack(U: torch.Tensor, S: torch.Tensor, Vh: torch.Tensor, idx: torch.Tensor):
    """(in_w [k,in] = Vh[idx]*sqrt(S[idx])[:,None], out_w [out,k] = U[:,idx]*sqrt(S[idx]))."""
    lib = _lib.load()
    dev = _need_cuda(U, S, Vh, idx)
    U, S, Vh = _f32c(U, "U"), _f32c(S, "S"), _f32c(Vh, "Vh")
    if idx.dtype != torch.int64:
        raise TypeError("idx must be int64")
    idx = idx.contiguous()
    out, r = U.shape
    in_ = Vh.shape[1]
    k = idx.numel()
    in_w = torch.empty(k, in_, dtype=torch.float32, device=dev)
    out_w = torch.empty(out, k, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.grasp_factor_pack(U.data_ptr(), S.data_ptr(), Vh.data_ptr(), idx.data_ptr(), k, out, in_, r,
                                    in_w.data_ptr(), out_w.data_ptr(), _stream()), "grasp_factor_pack")
    return in_w, out_w


# ------------------------------------------------------------------------- GEMM
def gemm(A: torch.Tensor, B: torch.Tensor, ta: bool = False, tb: bool = False, alpha: float = 1.0,
         beta: float = 0.0, C_out: Optional[torch.Tensor] = None, prec: Optional[int] = None) -> torch.Tensor:
    """C = alpha * op(A) op(B) + beta * C (fp32, row-major)."""
    lib = _lib.load()
    dev = _need_cuda(A, B, C_out)
    A, B = _f32c(A, "A"), _f32c(B, "B")
    M, K = (A.shape[1], A.shape[0]) if ta else (A.shape[0], A.shape[1])
    K2, N = (B.shape[1], B.shape[0]) if tb else (B.shape[0], B.shape[1])
    if K != K2:
        raise ValueError("inner dimensions differ")
    if C_out is None:
        if beta != 0.0:
            raise ValueError("beta != 0 needs C_out")
        C_out = torch.empty(M, N, dtype=torch.float32, device=dev)
    elif C_out.dtype != torch.float32 or not C_out.is_contiguous() or C_out.shape != (M, N):
        raise TypeError("C_out must be a contiguous float32 [M,N] tensor")
    p = _prec(prec)
    ws = _workspace(lib.grasp_gemm_workspace_bytes(M, N, K, p), dev)
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_gemm_f32(int(ta), int(tb), M, N, K, float(alpha), A.data_ptr(), A.shape[1], B.data_ptr(),
                                 B.shape[1], float(beta), C_out.data_ptr(), N, p, ws.data_ptr(), ws.numel(),
                                 _stream()), "grasp_gemm_f32")
        timers.stop("grasp_gemm_f32", t0, flops=2.0 * M * N * K, mma_per_flop=_mma_per_flop(p))
    ws.record_stream(torch.cuda.current_stream())
    return C_out


# ------------------------------------------------------- prepared GEMM operands
class Operand:
    """fp16 (hi, lo) planes + inverse scales of one fp32 matrix in its stored orientation
    (include/grasp_b200.h, grasp_gemm_split_f16).  `src` keeps the source storage alive."""
    __slots__ = ("planes", "inv", "rows", "cols", "mode", "src", "version", "_base")

    def nbytes(self) -> int:
        return self._base.numel() + self.inv.numel() * 4


def _new_operand(rows: int, cols: int, mode: int, dev) -> Operand:
    lib = _lib.load()
    nbytes = lib.grasp_gemm_planes_bytes(rows, cols)
    base = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    op = Operand()
    op._base = base
    op.planes = base.data_ptr() + (-base.data_ptr()) % 1024
    op.inv = torch.empty(rows if mode == _lib.SCALE_ROWS else max(rows, cols) + 1, dtype=torch.float32, device=dev)
    op.rows, op.cols, op.mode, op.src, op.version = rows, cols, mode, None, 0
    return op


def split_f16(x: torch.Tensor, mode: int = _lib.SCALE_ROWS, keep_src: bool = False) -> Operand:
    """Planes of a contiguous fp32 [rows, cols] matrix; SCALE_ROWS for activations, SCALE_TENSOR for weights."""
    lib = _lib.load()
    dev = _need_cuda(x)
    if x.dim() != 2:
        raise ValueError("split_f16 expects a 2-D matrix")
    x = _f32c(x, "x")
    rows, cols = x.shape
    if rows == 0 or cols == 0:
        raise ValueError("split_f16: empty matrix")
    op = _new_operand(rows, cols, mode, dev)
    op.src = x if keep_src else None
    op.version = x._version
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_gemm_split_f16(x.data_ptr(), cols, rows, cols, int(mode), op.planes, op.inv.data_ptr(),
                                       _stream()), "grasp_gemm_split_f16")
        timers.stop("grasp_gemm_split_f16", t0, bytes_=8.0 * rows * cols)
    return op


def gemm_planes(a: Operand, b: Operand, b_kn: bool = False, alpha: float = 1.0, beta: float = 0.0,
                C_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = alpha * A op(B) + beta * C on prepared operands: b_kn=False -> B stored [N, K] (x W^T);
    b_kn=True -> B stored [K, N] (dy W), which needs a tensor-scaled B."""
    lib = _lib.load()
    M, K = a.rows, a.cols
    if b_kn:
        K2, N = b.rows, b.cols
        if b.mode != _lib.SCALE_TENSOR:
            raise ValueError("a [K, N] operand must be tensor-scaled")
    else:
        N, K2 = b.rows, b.cols
    if K != K2:
        raise ValueError(f"inner dimensions differ: {K} vs {K2}")
    dev = a.inv.device
    if C_out is None:
        if beta != 0.0:
            raise ValueError("beta != 0 needs C_out")
        C_out = torch.empty(M, N, dtype=torch.float32, device=dev)
    elif C_out.dtype != torch.float32 or not C_out.is_contiguous() or C_out.shape != (M, N) or not C_out.is_cuda:
        raise TypeError("C_out must be a contiguous float32 CUDA [M,N] tensor")
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_gemm_f16x3_planes(M, N, K, float(alpha), a.planes, a.inv.data_ptr(), b.planes, int(bool(b_kn)),
                                          b.inv.data_ptr(), float(beta), C_out.data_ptr(), N, _stream()),
              "grasp_gemm_f16x3_planes")
        timers.stop("grasp_gemm_f16x3_planes", t0, flops=2.0 * M * N * K, mma_per_flop=3.0)
    return C_out


def gemm_planes_to_operand(a: Operand, b: Operand, b_kn: bool = False) -> Operand:
    """A op(B) handed over as a prepared (row-scaled) operand: the first GEMM of a factor pair.  B tensor-scaled."""
    lib = _lib.load()
    M, K = a.rows, a.cols
    K2, N = (b.rows, b.cols) if b_kn else (b.cols, b.rows)
    if K != K2:
        raise ValueError(f"inner dimensions differ: {K} vs {K2}")
    if b.mode != _lib.SCALE_TENSOR:
        raise ValueError("gemm_planes_to_operand needs a tensor-scaled B")
    dev = a.inv.device
    out = _new_operand(M, N, _lib.SCALE_ROWS, dev)
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_gemm_f16x3_planes_out(M, N, K, a.planes, a.inv.data_ptr(), b.planes, int(bool(b_kn)),
                                              b.inv.data_ptr(), out.planes, out.inv.data_ptr(), _stream()),
              "grasp_gemm_f16x3_planes_out")
        timers.stop("grasp_gemm_f16x3_planes", t0, flops=2.0 * M * N * K, mma_per_flop=3.0)
    return out


# ------------------------------------------------------- decoder-layer row kernels
def _rows2d(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D")
    return _f32c(t, name)


def rmsnorm_fwd(x: torch.Tensor, w: torch.Tensor, eps: float, want_y: bool = True, want_operand: bool = False):
    """(y, rstd) -- or (y or None, rstd, Operand) with want_operand: y as a prepared GEMM operand straight from the kernel."""
    lib = _lib.load()
    dev = _need_cuda(x, w)
    x, w = _rows2d(x, "x"), _f32c(w, "w")
    rows, d = x.shape
    if not (want_y or want_operand):
        raise ValueError("rmsnorm_fwd: nothing requested")
    y = torch.empty_like(x) if want_y else None
    rstd = torch.empty(rows, dtype=torch.float32, device=dev)
    op = _new_operand(rows, d, _lib.SCALE_ROWS, dev) if (want_operand and rows > 0) else None
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_rmsnorm_fwd(x.data_ptr(), w.data_ptr(), rows, d, float(eps), y.data_ptr() if y is not None else None,
                                    rstd.data_ptr(), op.planes if op is not None else None,
                                    op.inv.data_ptr() if op is not None else None, _stream()), "grasp_rmsnorm_fwd")
        timers.stop("grasp_rowops", t0, bytes_=4.0 * rows * d * (1 + int(want_y) + int(op is not None)))
    return (y, rstd, op) if want_operand else (y, rstd)


def rmsnorm_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, rstd: torch.Tensor,
                add: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    dev = _need_cuda(dy, x, w, rstd, add)
    dy, x, w, rstd = _rows2d(dy, "dy"), _rows2d(x, "x"), _f32c(w, "w"), _f32c(rstd, "rstd")
    if add is not None:
        add = _rows2d(add, "add")
    rows, d = x.shape
    dx = torch.empty_like(x)
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_rmsnorm_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), rstd.data_ptr(),
                                    add.data_ptr() if add is not None else None, rows, d, dx.data_ptr(), _stream()),
              "grasp_rmsnorm_bwd")
        timers.stop("grasp_rowops", t0, bytes_=(12.0 + (4.0 if add is not None else 0.0)) * rows * d)
    return dx


def rope_(x: torch.Tensor, seq: int, heads: int, hd: int, cos: torch.Tensor, sin: torch.Tensor,
          inverse: bool = False) -> torch.Tensor:
    """In place on x [tokens, heads*hd]; cos/sin [1 or B, seq, hd]."""
    lib = _lib.load()
    dev = _need_cuda(x, cos, sin)
    if x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 2 or x.shape[1] != heads * hd:
        raise TypeError("rope_: x must be a contiguous float32 [tokens, heads*hd] tensor")
    cos, sin = _f32c(cos, "cos"), _f32c(sin, "sin")
    if cos.shape[-2:] != (seq, hd) or sin.shape != cos.shape:
        raise ValueError("rope_: cos/sin must be [*, seq, head_dim]")
    batch = cos.shape[0] if cos.dim() == 3 else 1
    cs_batch = 0 if batch == 1 else seq * hd
    tokens = x.shape[0]
    if batch != 1 and batch * seq != tokens:
        raise ValueError("rope_: cos/sin batch does not match the token count")
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_rope_inplace(x.data_ptr(), tokens, seq, heads, hd, cos.data_ptr(), sin.data_ptr(), cs_batch,
                                     int(bool(inverse)), _stream()), "grasp_rope_inplace")
        timers.stop("grasp_rowops", t0, bytes_=8.0 * x.numel())
    return x


def swiglu_fwd(g: torch.Tensor, u: torch.Tensor, want_h: bool = True, want_operand: bool = False):
    """h = silu(g) * u -- or (h or None, Operand of h) with want_operand."""
    lib = _lib.load()
    dev = _need_cuda(g, u)
    g, u = _rows2d(g, "g"), _rows2d(u, "u")
    if g.shape != u.shape:
        raise ValueError("swiglu: shapes differ")
    rows, cols = g.shape
    if not (want_h or want_operand):
        raise ValueError("swiglu_fwd: nothing requested")
    h = torch.empty_like(g) if want_h else None
    op = _new_operand(rows, cols, _lib.SCALE_ROWS, dev) if (want_operand and rows > 0) else None
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_swiglu_fwd(g.data_ptr(), u.data_ptr(), rows, cols, h.data_ptr() if h is not None else None,
                                   op.planes if op is not None else None, op.inv.data_ptr() if op is not None else None,
                                   _stream()), "grasp_swiglu_fwd")
        timers.stop("grasp_rowops", t0, bytes_=4.0 * g.numel() * (2 + int(want_h) + int(op is not None)))
    return (h, op) if want_operand else h


def swiglu_bwd(dh: torch.Tensor, g: torch.Tensor, u: torch.Tensor, inplace: bool = False, want_grads: bool = True,
               want_operands: bool = False):
    """(dg, du); inplace=True overwrites g and u with their gradients.  With want_operands the result is
    (dg or None, du or None, Operand of dg, Operand of du): the gradients as prepared GEMM operands."""
    lib = _lib.load()
    dev = _need_cuda(dh, g, u)
    dh = _rows2d(dh, "dh")
    if g.dtype != torch.float32 or u.dtype != torch.float32 or not g.is_contiguous() or not u.is_contiguous():
        raise TypeError("swiglu_bwd: g and u must be contiguous float32")
    if g.dim() != 2 or g.shape != u.shape or dh.shape != g.shape:
        raise ValueError("swiglu: shapes differ")
    if not (want_grads or want_operands):
        raise ValueError("swiglu_bwd: nothing requested")
    rows, cols = g.shape
    dg = (g if inplace else torch.empty_like(g)) if want_grads else None
    du = (u if inplace else torch.empty_like(u)) if want_grads else None
    og = _new_operand(rows, cols, _lib.SCALE_ROWS, dev) if (want_operands and rows > 0) else None
    ou = _new_operand(rows, cols, _lib.SCALE_ROWS, dev) if (want_operands and rows > 0) else None
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_swiglu_bwd(dh.data_ptr(), g.data_ptr(), u.data_ptr(), rows, cols,
                                   dg.data_ptr() if dg is not None else None, du.data_ptr() if du is not None else None,
                                   og.planes if og is not None else None, og.inv.data_ptr() if og is not None else None,
                                   ou.planes if ou is not None else None, ou.inv.data_ptr() if ou is not None else None,
                                   _stream()), "grasp_swiglu_bwd")
        timers.stop("grasp_rowops", t0, bytes_=4.0 * g.numel() * (3 + 2 * int(want_grads) + 2 * int(og is not None)))
    return (dg, du, og, ou) if want_operands else (dg, du)


def ce_loss_bwd_(logits: torch.Tensor, labels: torch.Tensor, coef: torch.Tensor) -> torch.Tensor:
    """loss[t] = coef[t] * CE(logits[t], labels[t]); logits is overwritten by dloss/dlogits.  labels < 0 are ignored."""
    lib = _lib.load()
    dev = _need_cuda(logits, labels, coef)
    if logits.dtype != torch.float32 or not logits.is_contiguous() or logits.dim() != 2:
        raise TypeError("ce_loss_bwd_: logits must be a contiguous float32 [rows, V] tensor")
    if labels.dtype != torch.int64:
        raise TypeError("ce_loss_bwd_: labels must be int64")
    labels, coef = labels.contiguous(), _f32c(coef, "coef")
    rows, V = logits.shape
    if labels.numel() != rows or coef.numel() != rows:
        raise ValueError("ce_loss_bwd_: one label and one coefficient per row")
    loss = torch.empty(rows, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_ce_loss_bwd(logits.data_ptr(), labels.data_ptr(), coef.data_ptr(), rows, V, loss.data_ptr(),
                                    _stream()), "grasp_ce_loss_bwd")
        timers.stop("grasp_rowops", t0, bytes_=8.0 * rows * V)
    return loss


# ----------------------------------------------------------------------- attention
ATTN_HEAD_DIMS = (64, 128)


def attn_supported(head_dim: int) -> bool:
    return int(head_dim) in ATTN_HEAD_DIMS


def attn_prep_qkv(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, S: int, H: int, Hkv: int, D: int, cos: torch.Tensor,
                  sin: torch.Tensor):
    """Operands of attn_fwd_prepared from the projections as they leave their GEMMs: RoPE on q and k (cos / sin
    [1 or B, S, D]), then tensor-scaled planes of the rotated q, k and of v.  q, k, v are left untouched."""
    lib = _lib.load()
    dev = _need_cuda(q, k, v, cos, sin)
    q, k, v = _rows2d(q, "q"), _rows2d(k, "k"), _rows2d(v, "v")
    tokens = q.shape[0]
    if q.shape[1] != H * D or k.shape != (tokens, Hkv * D) or v.shape != k.shape or tokens % S:
        raise ValueError("attn_prep_qkv: q [tokens, H*D], k / v [tokens, Hkv*D], tokens a multiple of S")
    cos, sin = _f32c(cos, "cos"), _f32c(sin, "sin")
    if cos.shape[-2:] != (S, D) or sin.shape != cos.shape:
        raise ValueError("attn_prep_qkv: cos/sin must be [*, S, head_dim]")
    batch = cos.shape[0] if cos.dim() == 3 else 1
    if batch != 1 and batch * S != tokens:
        raise ValueError("attn_prep_qkv: cos/sin batch does not match the token count")
    cs_batch = 0 if batch == 1 else S * D
    qo = _new_operand(tokens, H * D, _lib.SCALE_TENSOR, dev)
    ko = _new_operand(tokens, Hkv * D, _lib.SCALE_TENSOR, dev)
    vo = _new_operand(tokens, Hkv * D, _lib.SCALE_TENSOR, dev)
    ws = torch.empty(4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_attn_prep_qkv(q.data_ptr(), k.data_ptr(), v.data_ptr(), tokens, S, H, Hkv, D, cos.data_ptr(),
                                      sin.data_ptr(), cs_batch, qo.planes, qo.inv.data_ptr(), ko.planes, ko.inv.data_ptr(),
                                      vo.planes, vo.inv.data_ptr(), ws.data_ptr(), _stream()), "grasp_attn_prep_qkv")
        timers.stop("grasp_rowops", t0, bytes_=(8.0 + 4.0) * (q.numel() + k.numel() + v.numel()))
    return qo, ko, vo


def attn_fwd_prepared(qo: Operand, ko: Operand, vo: Operand, B: int, S: int, H: int, Hkv: int, D: int, scale: float):
    """Causal attention on prepared (RoPE'd, tensor-scaled) operands.  Returns (out [B*S, H*D], ctx for attn_bwd)."""
    lib = _lib.load()
    dev = qo.inv.device
    if (qo.rows, qo.cols) != (B * S, H * D) or (ko.rows, ko.cols) != (B * S, Hkv * D) or (vo.rows, vo.cols) != (ko.rows, ko.cols):
        raise ValueError("attn_fwd: q [B*S, H*D], k / v [B*S, Hkv*D] expected")
    if not (qo.mode == ko.mode == vo.mode == _lib.SCALE_TENSOR):
        raise ValueError("attn_fwd: operands must be tensor-scaled")
    out = torch.empty(B * S, H * D, dtype=torch.float32, device=dev)
    lse2 = torch.empty(B * H * S, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_attn_fwd(qo.planes, qo.inv.data_ptr(), ko.planes, ko.inv.data_ptr(), vo.planes, vo.inv.data_ptr(),
                                 B, S, H, Hkv, D, float(scale), out.data_ptr(), lse2.data_ptr(), _stream()), "grasp_attn_fwd")
        timers.stop("grasp_attn", t0, flops=2.0 * B * H * S * S * D, mma_per_flop=3.0)      # causal: half of 4 S^2 D
    return out, (qo, ko, vo, out, lse2, (B, S, H, Hkv, D, float(scale)))


def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, S: int, H: int, Hkv: int, D: int, scale: float):
    """Causal attention of [B*S, H*D] / [B*S, Hkv*D] activations (RoPE already applied)."""
    _need_cuda(q, k, v)
    if q.shape != (B * S, H * D) or k.shape != (B * S, Hkv * D) or v.shape != k.shape:
        raise ValueError("attn_fwd: q [B*S, H*D], k / v [B*S, Hkv*D] expected")
    qo, ko, vo = (split_f16(t, _lib.SCALE_TENSOR) for t in (q, k, v))
    return attn_fwd_prepared(qo, ko, vo, B, S, H, Hkv, D, scale)


def attn_bwd(ctx, d_out: torch.Tensor):
    """(dq [B*S, H*D], dk, dv [B*S, Hkv*D]) for the ctx of attn_fwd and dL/d(out)."""
    lib = _lib.load()
    qo, ko, vo, out, lse2, (B, S, H, Hkv, D, scale) = ctx
    dev = _need_cuda(d_out)
    d_out = _f32c(d_out, "d_out")
    if d_out.shape != (B * S, H * D):
        raise ValueError("attn_bwd: d_out must be [B*S, H*D]")
    doo = split_f16(d_out, _lib.SCALE_TENSOR)
    dq = torch.empty(B * S, H * D, dtype=torch.float32, device=dev)
    dk = torch.empty(B * S, Hkv * D, dtype=torch.float32, device=dev)
    dv = torch.empty(B * S, Hkv * D, dtype=torch.float32, device=dev)
    delta = torch.empty(B * H * S, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        t0 = timers.start()
        check(lib.grasp_attn_bwd(qo.planes, qo.inv.data_ptr(), ko.planes, ko.inv.data_ptr(), vo.planes, vo.inv.data_ptr(),
                                 doo.planes, doo.inv.data_ptr(), d_out.data_ptr(), out.data_ptr(), lse2.data_ptr(), B, S, H, Hkv,
                                 D, scale, dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), delta.data_ptr(), _stream()),
              "grasp_attn_bwd")
        timers.stop("grasp_attn", t0, flops=7.0 * B * H * S * S * D, mma_per_flop=3.0)      # 7 products, causal half
    return dq, dk, dv
