"""CPU, world_size 2 over gloo: the multi-GPU plan (matrix partition + factor exchange, sample shards +
summed sigma-gradients) without any kernel."""
import os
import socket

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from grasp_b200 import dist


def test_partition_is_balanced_and_deterministic():
    shapes = [(4096, 4096)] * 4 + [(11008, 4096), (11008, 4096), (4096, 11008)]
    shapes = shapes * 8                                   # LLaMA-2-7B, 8 layers: 56 matrices
    for world in (1, 2, 4, 8):
        owner = dist.owners_of(shapes, world)
        assert owner == dist.owners_of(shapes, world)
        load = [0.0] * world
        for (m, n), o in zip(shapes, owner):
            load[o] += dist.svd_time_cost(m, n)          # what the Jacobi SVD costs (the nominal flops mis-rate MLP matrices)
        assert max(load) / (sum(load) / world) < 1.05, (world, load)
        assert max(owner.count(r) for r in range(world)) - min(owner.count(r) for r in range(world)) <= 1
    assert abs(dist.svd_cost(4096, 4096) - 641e9) / 641e9 < 0.01
    assert abs(dist.svd_cost(11008, 4096) - 1569e9) / 1569e9 < 0.01
    # working shapes: wide / tall matrices run their Jacobi phase on the square CholeskyQR factor
    assert dist.svd_working_shape(11008, 4096) == (4096, 4096) and dist.svd_working_shape(4096, 4096) == (4096, 4096)
    assert dist.svd_working_shape(256, 2048) == (256, 2048) and dist.svd_working_shape(1024, 4096) == (1024, 1024)


def test_sample_shards_cover_everything_once():
    for n in (512, 7, 1, 0):
        for world in (1, 2, 3, 8):
            spans = [dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # sigma-gradient exchange: every rank holds the partial sums of its sample shard
        g = torch.Generator().manual_seed(0)
        per_sample = torch.randn(10, 3, 16, generator=g)             # 10 samples, 3 matrices, r = 16
        lo, hi = dist.shard_range(10, rank, world)
        parts = [per_sample[lo:hi, j].sum(0) for j in range(3)]
        dist.all_reduce_sum_many_(parts)
        want = per_sample.sum(0)
        assert all(torch.allclose(parts[j], want[j], atol=1e-5) for j in range(3))
        acc = torch.full((4,), float(rank + 1), dtype=torch.float64)
        dist.all_reduce_sum_(acc)
        assert torch.all(acc == sum(range(1, world + 1)))
        # SVD factor exchange: owners compute (here: fill with their matrix id), everyone receives
        shapes = [(8, 8), (12, 8), (8, 12), (8, 8), (12, 8)]
        owner = dist.owners_of(shapes, world)
        local = {}
        for i, (m, n) in enumerate(shapes):
            if owner[i] == rank:
                r = min(m, n)
                local[i] = (torch.full((m, r), float(i)), torch.full((r,), float(i)), torch.full((r, n), float(i)))
        got = dist.exchange_factors(local, shapes, owner, "cpu")
        for i, (U, S, Vh) in enumerate(got):
            assert U.shape == (shapes[i][0], min(shapes[i])) and Vh.shape == (min(shapes[i]), shapes[i][1])
            assert torch.all(U == i) and torch.all(S == i) and torch.all(Vh == i)
        # plans that decide how many collectives follow (SVD hoisting window) must agree on every rank
        assert dist.all_min_int(5 + 3 * rank, "cpu") == 5
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        td.destroy_process_group()


def test_two_rank_exchange_over_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
