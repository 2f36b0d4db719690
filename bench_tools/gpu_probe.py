"""Ad-hoc GPU probe: timings of the individual kernels at LLaMA shapes (CUDA events)."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from grasp_b200 import ops

def timed(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

dev = "cuda"
res = {}
sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048]
for n in sizes:
    A = torch.randn(n, n, device=dev) * 0.02
    t = timed(lambda: ops.svd_batched([A]), n=1, warm=1)
    outs, info = ops.svd_batched([A], return_info=True)
    U, S, Vh = outs[0]
    Sref = torch.linalg.svdvals(A.double())
    rec = (torch.linalg.norm((U * S) @ Vh - A) / torch.linalg.norm(A)).item()
    orth = (U.T @ U - torch.eye(n, device=dev)).abs().max().item()
    tt = timed(lambda: torch.linalg.svd(A, full_matrices=False), n=1, warm=1)
    res[f"svd_{n}"] = dict(ms=t, torch_ms=tt, info=info.cpu().tolist(), sigma_err=((S - Sref).abs().max() / Sref[0]).item(), recon=rec, orthU=orth)
    print(res[f"svd_{n}"], flush=True)
    G = torch.randn(n, n, device=dev)
    res[f"score_{n}"] = timed(lambda: ops.sigma_score(U, G, Vh, S))
    sc = ops.sigma_score(U, G, Vh, S)[1]
    k = int(n * n * 0.1 / (2 * n))
    res[f"topk_{n}"] = timed(lambda: ops.topk(sc, k), n=10)
    idx = ops.topk(sc, k)
    res[f"rebuild_{n}"] = timed(lambda: ops.lowrank_rebuild(U, S, Vh, idx))
    print({k_: v for k_, v in res.items() if not k_.startswith("svd")}, flush=True)
hs = [torch.randn(1, 511, 4096, device=dev) for _ in range(33)]
acc = torch.zeros(32, dtype=torch.float64, device=dev)
t = timed(lambda: ops.bi_chain(hs, acc), n=20, warm=3)
res["bi_chain_33x511x4096_ms"] = t
res["bi_alg_GBs"] = 2 * 32 * 511 * 4096 * 4 / t / 1e6
print(res["bi_chain_33x511x4096_ms"], res["bi_alg_GBs"])
json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)
