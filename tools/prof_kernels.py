"""Short driver for ncu --set full captures of the individual kernels (one GPU)."""
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops
what = sys.argv[1]
dev = "cuda"
torch.manual_seed(0)
if what == "gemm":
    A = torch.randn(4096, 4096, device=dev); B = torch.randn(4096, 4096, device=dev)
    for _ in range(3): ops.gemm(A, B, tb=True, prec=6)
    for _ in range(3): ops.gemm(A, B, tb=True, prec=3)
elif what == "sigma":
    U = torch.randn(4096, 4096, device=dev); Vh = torch.randn(4096, 4096, device=dev)
    G = torch.randn(4096, 4096, device=dev); S = torch.rand(4096, device=dev)
    for _ in range(3): ops.sigma_score(U, G, Vh, S, prec=6)
elif what == "bi":
    hs = [torch.randn(8, 511, 4096, device=dev) for _ in range(33)]
    acc = torch.zeros(32, dtype=torch.float64, device=dev)
    for _ in range(4): ops.bi_chain(hs, acc)
elif what == "svd":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    A = torch.randn(n, n, device=dev) * 0.02
    ops.svd_batched([A], max_sweeps=2)
torch.cuda.synchronize()
print("done", what)
