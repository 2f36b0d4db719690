"""BASELINE.json configs[4]: standalone SVD + importance + top-k + rebuild sweep over LLaMA-shaped matrices,
fp32 and bf16 inputs, against torch's own CUDA library path on the same B200 (cuSOLVER gesvd via
torch.linalg.svd, cuBLAS, torch.topk) and, optionally, the reference's CPU path."""
import json, os, sys, time
import torch
sys.path.insert(0, ".")
from grasp_b200 import ops

dev = "cuda"
shapes = [(4096, 4096), (4096, 11008), (11008, 4096)]
if "--big" in sys.argv:
    shapes.append((8192, 28672))
cpu = "--cpu" in sys.argv
out = {}

def ev(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)

for (m, n) in shapes:
    for in_dtype in (torch.float32, torch.bfloat16):
        torch.manual_seed(0)
        W = (torch.randn(m, n, device=dev) * 0.02).to(in_dtype)
        Wf = W.float()                      # bf16 inputs are up-cast (the reference CPU svd rejects bf16)
        G = torch.randn(m, n, device=dev)
        k = int(m * n * 0.1 / (m + n))
        key = f"{m}x{n}_{str(in_dtype).split('.')[-1]}"
        ops.svd(Wf[:256, :256].contiguous())       # warm the kernels
        (U, S, Vh), t_svd = ev(lambda: ops.svd(Wf))
        (g, sc), t_score = ev(lambda: ops.sigma_score(U, G, Vh, S))
        idx, t_topk = ev(lambda: ops.topk(sc, k))
        Wk, t_reb = ev(lambda: ops.lowrank_rebuild(U, S, Vh, idx, out_dtype=torch.bfloat16 if in_dtype == torch.bfloat16 else torch.float32))
        rec = {"k": k, "ours_ms": {"svd": t_svd, "score": t_score, "topk": t_topk, "rebuild": t_reb,
                                    "total": t_svd + t_score + t_topk + t_reb}}
        # fp64 truth on the GPU (accuracy reference for both paths)
        if max(m, n) <= 16384:
            U64, S64, V64 = torch.linalg.svd(Wf.double(), full_matrices=False)
            g64 = ((U64.T @ G.double()) * V64).sum(-1)
            s64 = (g64 * S64).abs()
            i64 = torch.topk(s64, k).indices
            kth = s64[i64[-1]].item()
            a, b = set(idx.tolist()), set(i64.tolist())
            rec["vs_fp64"] = {"sigma_err": ((S.double() - S64).abs().max() / S64[0]).item(),
                              "jaccard": len(a & b) / len(a | b),
                              "mismatches_within_5pct_of_kth": all(abs(s64[i].item() - kth) <= 0.05 * kth for i in a ^ b),
                              "score_err_over_max": ((sc.double() - s64).abs().max() / s64.max()).item()}
            del U64, V64
        # torch CUDA library path on the same GPU
        if max(m, n) <= 16384:
            torch.linalg.svd(Wf[:256, :256], full_matrices=False)
            _w = ((U.T[:64] @ G) * Vh[:64]).sum(-1); torch.topk(_w.abs(), 8); torch.cuda.synchronize()
            (Ut, St, Vt), tt_svd = ev(lambda: torch.linalg.svd(Wf, full_matrices=False))
            gt, tt_score = ev(lambda: ((Ut.T @ G) * Vt).sum(-1))
            it, tt_topk = ev(lambda: torch.topk((gt * St).abs(), k).indices)
            Wt, tt_reb = ev(lambda: Ut[:, it] @ (torch.diag(St[it]) @ Vt[it, :]))
            rec["torch_cuda_ms"] = {"svd": tt_svd, "score": tt_score, "topk": tt_topk, "rebuild": tt_reb,
                                    "total": tt_svd + tt_score + tt_topk + tt_reb}
            b2 = set(it.tolist())
            rec["torch_cuda_vs_fp64"] = {"sigma_err": ((St.double() - S64).abs().max() / S64[0]).item(),
                                         "jaccard": len(b2 & b) / len(b2 | b),
                                         "recon": (torch.linalg.norm((Ut * St) @ Vt - Wf) / torch.linalg.norm(Wf)).item()}
        rec["recon"] = (torch.linalg.norm((U * S) @ Vh - Wf) / torch.linalg.norm(Wf)).item()
        r = min(m, n)
        rec["orthU"] = (U.T @ U - torch.eye(r, device=dev)).abs().max().item()
        if cpu and in_dtype == torch.float32 and max(m, n) <= 11008:
            Wc, Gc = Wf.cpu(), G.cpu()
            torch.set_num_threads(os.cpu_count())
            t0 = time.perf_counter(); Uc, Sc, Vc = torch.linalg.svd(Wc, full_matrices=False); t1 = time.perf_counter()
            gc = ((Uc.T @ Gc) * Vc).sum(-1); ic = torch.topk((gc * Sc).abs(), k).indices
            Wc2 = Uc[:, ic] @ (torch.diag(Sc[ic]) @ Vc[ic, :]); t2 = time.perf_counter()
            rec["cpu_reference_s"] = {"svd": t1 - t0, "rest": t2 - t1, "cores": os.cpu_count()}
        out[key] = rec
        print(key, json.dumps(rec), flush=True)
        del U, S, Vh, W, Wf, G
        torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)
