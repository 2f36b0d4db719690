"""Stage orchestration behind modeling_grasp.GRASPModel.

Holds the pieces that have no counterpart in the reference because the reference leaves
them to autograd / per-call library dispatch:

  BlockInfluence        device-side accumulator for compute_bi (one launch per forward pass,
                        one D2H copy per calibration run instead of one per layer per batch)
  batched_svd           groups same-shape weights into batched SVD launches
  SigmaLinearFn         forward/backward of a GRASPLayer against the dense weight; harvests
                        G = dY^T X instead of differentiating through U diag(S) Vh
  deferred_sigma_grads  accumulate G over a calibration pass, contract diag(U^T G V) once
"""
from __future__ import annotations

import contextlib
from collections import OrderedDict
from typing import Iterable, List, Sequence

import torch

from . import ops

# "torch": the three linear GEMMs of a GRASPLayer go through torch.matmul (cuBLAS fp32);
# "grasp": they go through grasp_gemm_f32 (split-bf16 tcgen05 path of this library).
_LINEAR_BACKEND = "torch"


def set_linear_backend(name: str) -> None:
    global _LINEAR_BACKEND
    if name not in ("torch", "grasp"):
        raise ValueError(name)
    _LINEAR_BACKEND = name


def linear_backend() -> str:
    return _LINEAR_BACKEND


def _mm(a: torch.Tensor, b: torch.Tensor, ta=False, tb=False, out=None, accumulate=False) -> torch.Tensor:
    if _LINEAR_BACKEND == "grasp":
        return ops.gemm(a, b, ta=ta, tb=tb, beta=1.0 if accumulate else 0.0, C_out=out)
    A = a.t() if ta else a
    B = b.t() if tb else b
    if out is None:
        return A @ B
    if accumulate:
        return out.addmm_(A, B)
    return torch.mm(A, B, out=out)


class BlockInfluence:
    """Accumulates sum over batches of mean-over-tokens block influence per layer
    (reference modeling_grasp.py:148-167) in a device-resident float64 vector."""

    def __init__(self, n_layers: int, angular: bool = False, stride: int = 1):
        self.n_layers = n_layers
        self.angular = bool(angular)
        self.stride = max(int(stride or 1), 1) if angular else 1
        self.acc = None

    def add(self, hiddens: Sequence[torch.Tensor]) -> None:
        hiddens = list(hiddens)
        if self.acc is None:
            n = max(self.n_layers, len(hiddens) - 1)
            self.acc = torch.zeros(n, dtype=torch.float64, device=hiddens[0].device)
        if not self.angular:
            ops.bi_chain(hiddens, self.acc)
            return
        # angular variant: last token only, layer i against layer i+stride
        for i in range(len(hiddens) - self.stride):
            ops.bi_accumulate(hiddens[i][:, -1:], hiddens[i + self.stride][:, -1:], self.acc[i:i + 1], angular=True)

    def result(self) -> List[float]:
        if self.acc is None:
            return [0.0] * self.n_layers
        return self.acc[: self.n_layers].cpu().tolist()


def batched_svd(weights: Sequence[torch.Tensor], max_group: int = 8):
    """SVD of every weight; same-shape matrices share launches (latency-bound eigen-solves overlap)."""
    groups: "OrderedDict[tuple, list]" = OrderedDict()
    for i, w in enumerate(weights):
        groups.setdefault(tuple(w.shape), []).append(i)
    out = [None] * len(weights)
    for idxs in groups.values():
        for s in range(0, len(idxs), max_group):
            chunk = idxs[s:s + max_group]
            for i, usv in zip(chunk, ops.svd_batched([weights[i] for i in chunk])):
                out[i] = usv
    return out


class SigmaLinearFn(torch.autograd.Function):
    """y = x W^T with W the dense weight of a GRASPLayer; the gradient w.r.t. the singular
    values is dS_i = u_i^T (dY^T X) v_i  (identical to autograd through U diag(S) Vh)."""

    @staticmethod
    def forward(ctx, x, S, layer):
        ctx.layer = layer
        ctx.save_for_backward(x)
        return _mm(x, layer.dense_weight(), tb=True)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        layer = ctx.layer
        dy = dy.contiguous()
        dx = _mm(dy, layer.dense_weight()) if ctx.needs_input_grad[0] else None
        dS = None
        if ctx.needs_input_grad[1]:
            if layer._defer:
                if layer._G is None:
                    layer._G = _mm(dy, x, ta=True)
                else:
                    _mm(dy, x, ta=True, out=layer._G, accumulate=True)
            else:
                G = _mm(dy, x, ta=True)
                dS, _ = ops.sigma_score(layer.U.data, G, layer.Vh.data, None, metric="gradient", want_score=False)
        return dx, dS, None


@contextlib.contextmanager
def deferred_sigma_grads(layers: Iterable):
    layers = list(layers)
    for layer in layers:
        layer._defer = True
        layer._G = None
    try:
        yield
    finally:
        for layer in layers:
            layer._defer = False
            layer._G = None


def contract_sigma_grad(layer, dsigma: torch.Tensor = None) -> torch.Tensor:
    """dL/dS from the accumulated G of one calibration pass (zeros if the layer saw no batch)."""
    if layer._G is None:
        g = torch.zeros_like(layer.S.data)
        return g if dsigma is None else dsigma
    g, _ = ops.sigma_score(layer.U.data, layer._G, layer.Vh.data, None, metric="gradient", dsigma=dsigma,
                           want_score=False)
    return g
