"""Multi-GPU plan of the hot path (one process per GPU, torch.distributed over NCCL/NVLink).

The reference is single-process; this partitioning is new:

  SVD + rebuild     independent weight matrices -> ranks (longest-processing-time on the nominal
                    SVD flop count), then each matrix's factors are broadcast from its owner
  BI / sigma-grads  calibration samples -> ranks (contiguous shards); the per-layer BI sums and the
                    per-block sigma-gradient vectors are summed with ONE all-reduce each
  selection/compile replicated: every rank holds the same scores, so the replicas stay identical

Everything is a no-op when torch.distributed is not initialised (single GPU).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as td


def initialized() -> bool:
    return td.is_available() and td.is_initialized()


def rank_world() -> Tuple[int, int]:
    if initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def svd_cost(m: int, n: int) -> float:
    """Nominal thin-SVD work 8 m n^2 + 4/3 n^3 with n = min side: the figure the
    roofline arithmetic uses."""
    big, small = max(m, n), min(m, n)
    return 8.0 * big * small * small + 4.0 / 3.0 * small ** 3


def svd_working_shape(m: int, n: int) -> Tuple[int, int]:
    """(rows, row length) the Jacobi phase of grasp_svd_batched works on: wide / tall matrices with a short side
    >= 512 and a long side >= 1.5 x that are first reduced to their square CholeskyQR factor (csrc/svd_jacobi.cu,
    pre_eligible).  Matrices of equal working shape share launches."""
    r, L = min(m, n), max(m, n)
    if r >= 512 and 2 * L >= 3 * r:
        return r, r
    return r, L


def svd_time_cost(m: int, n: int) -> float:
    """What the block-Jacobi SVD actually costs, up to a constant: rounds ~ r, bytes per round ~ r (L' + r), plus the
    GEMMs of the preconditioning over the long side (measured: 4096 x 11008 takes 1.08x a 4096 x 4096).  The
    nominal flop count would rate the MLP matrices 2.45x an attention matrix and leave the ranks that own attention
    matrices with twice the work."""
    r, L = min(m, n), max(m, n)
    rw, Lw = svd_working_shape(m, n)
    cost = float(rw) * rw * (Lw + rw)
    if (rw, Lw) != (r, L):
        cost += 0.04 * float(r) * r * L
    return cost


def partition_lpt(costs: Sequence[float], n_parts: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of items to parts; deterministic on every rank."""
    parts: List[List[int]] = [[] for _ in range(n_parts)]
    load = [0.0] * n_parts
    for i in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        p = min(range(n_parts), key=lambda p: (load[p], p))
        parts[p].append(i)
        load[p] += costs[i]
    for p in parts:
        p.sort()
    return parts


def owners_of(shapes: Sequence[Tuple[int, int]], world: int) -> List[int]:
    """Owner rank of every matrix."""
    owner = [0] * len(shapes)
    for r, items in enumerate(partition_lpt([svd_time_cost(m, n) for m, n in shapes], world)):
        for i in items:
            owner[i] = r
    return owner


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) shard of n samples for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if initialized() and td.get_world_size() > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def all_min_int(value: int, device) -> int:
    """The smallest `value` over the ranks (plans that decide how many collectives follow must agree)."""
    if not (initialized() and td.get_world_size() > 1):
        return int(value)
    dev = torch.device(device)
    t = torch.tensor([int(value)], dtype=torch.int64, device=dev if dev.type == "cuda" else "cpu")
    td.all_reduce(t, op=td.ReduceOp.MIN)
    return int(t.item())


def all_reduce_sum_many_(tensors: Sequence[torch.Tensor]) -> None:
    """One collective for a list of same-dtype vectors (the sigma-gradients of a block: 48-64 KB)."""
    if not (initialized() and td.get_world_size() > 1) or not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    td.all_reduce(flat, op=td.ReduceOp.SUM)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()


def exchange_factors(local: Dict[int, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                     shapes: Sequence[Tuple[int, int]], owner: Sequence[int], device, dtype=torch.float32):
    """Every rank ends up with (U, S, Vh) of every matrix: owners send, the others receive."""
    rank, world = rank_world()
    out = []
    pending = []
    for i, (m, n) in enumerate(shapes):
        r = min(m, n)
        if owner[i] == rank:
            U, S, Vh = local[i]
        else:
            U = torch.empty(m, r, dtype=dtype, device=device)
            S = torch.empty(r, dtype=dtype, device=device)
            Vh = torch.empty(r, n, dtype=dtype, device=device)
        if world > 1:
            for t in (U, S, Vh):
                pending.append(td.broadcast(t, src=owner[i], async_op=True))
        out.append((U, S, Vh))
    for w in pending:
        w.wait()
    return out
