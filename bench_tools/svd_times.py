"""SVD stage timings at the 7B shapes: batches as the job forms them, with / without the CholeskyQR2 preconditioning,
plus accuracy against torch fp64 on one matrix per shape (sigma, reconstruction, orthogonality, rebuilt rank-k weight)."""
import sys, time, torch
sys.path.insert(0, ".")
from grasp_b200 import ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
def ev(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)
cases = [("4096x4096 x8", [(4096, 4096)] * 8), ("4096x4096 x4", [(4096, 4096)] * 4), ("4096x11008 x4", [(4096, 11008)] * 4),
         ("11008x4096 x8", [(11008, 4096)] * 8), ("4096x4096 x1", [(4096, 4096)]), ("4096x11008 x1", [(4096, 11008)])]
if len(sys.argv) > 1:
    cases = [c for c in cases if any(a in c[0] for a in sys.argv[1:])]
for name, shapes in cases:
    mats = [torch.randn(m, n, device=dev, generator=g) * 0.02 for m, n in shapes]
    for pre in (True, False):
        ops.svd_batched(mats[:1], precondition=pre, max_sweeps=1)          # warm (attributes, tensor maps)
        (outs, info), ms = ev(lambda: ops.svd_batched(mats, precondition=pre, return_info=True))
        info = info.cpu()
        A = mats[0].double(); U, S, Vh = (t.double() for t in outs[0])
        S64 = torch.linalg.svdvals(A)
        rec = (torch.linalg.norm((U * S) @ Vh - A) / torch.linalg.norm(A)).item()
        r = S.numel(); eye = torch.eye(r, device=dev, dtype=torch.float64)
        orth = max((U.T @ U - eye).abs().max().item(), (Vh @ Vh.T - eye).abs().max().item())
        print(f"{name} precond={int(pre)}: {ms:8.1f} ms total, {ms / len(mats):7.1f} ms per matrix; sweeps {info[:, 0].tolist()} "
              f"(tensor-core {info[:, 3].tolist()}) converged {info[:, 1].tolist()}; sigma {((S - S64).abs().max() / S64[0]).item():.1e} "
              f"recon {rec:.1e} orth {orth:.1e}", flush=True)
        del outs
    del mats
    torch.cuda.empty_cache()
