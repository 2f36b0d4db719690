"""How the per-matrix SVD time depends on how many same-working-shape matrices share one grasp_svd_batched call
(engine._batched_svd_local's max_group).  Shapes of one LLaMA-2-7B layer: 4 square + 3 MLP matrices."""
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
def ev(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)
layer = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
sizes = [int(a) for a in sys.argv[1:]] or [7, 8, 14, 28]
ops.svd_batched([torch.randn(4096, 4096, device=dev, generator=g) * 0.02], max_sweeps=1)
for b in sizes:
    shapes = (layer * ((b + 6) // 7))[:b]
    mats = [torch.randn(m, n, device=dev, generator=g) * 0.02 for m, n in shapes]
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    (outs, info), ms = ev(lambda: ops.svd_batched(mats, return_info=True))
    info = info.cpu()
    print(f"batch {b:3d}: {ms:8.1f} ms total, {ms / b:7.1f} ms per matrix; sweeps {sorted(set(info[:, 0].tolist()))} "
          f"(per matrix {[(shapes[i][0] // 1000, int(info[i, 3]), int(info[i, 0])) for i in range(b)] if b <= 16 else ''}) "
          f"converged {int(info[:, 1].min())}; peak extra memory {(torch.cuda.max_memory_allocated() - base) / 2**30:.1f} GiB", flush=True)
    del outs, mats
    torch.cuda.empty_cache()
