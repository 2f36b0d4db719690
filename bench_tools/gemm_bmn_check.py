import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops
dev = "cuda"; torch.manual_seed(0)
ok = True
for prec, tol in ((3, 3e-5), (6, 2e-6)):
    for (M, N, K) in [(128, 128, 64), (128, 256, 128), (200, 300, 100), (1000, 777, 333), (512, 4096, 4096)]:
        for ta in (False, True):
            A = torch.randn((K, M) if ta else (M, K), device=dev); B = torch.randn(K, N, device=dev)
            want = (A.T if ta else A).double() @ B.double()
            C = ops.gemm(A, B, ta=ta, tb=False, prec=prec); torch.cuda.synchronize()
            e = (torch.linalg.norm(C.double() - want) / torch.linalg.norm(want)).item()
            print(f"BMN prec {prec} {M}x{N}x{K} ta={int(ta)}: {e:.3e} {'ok' if e < tol else 'FAIL'}", flush=True)
            if e >= tol:
                ok = False
                err = (C.double() - want).abs()
                bad_cols = (err.max(dim=0).values > 1e-3 * want.abs().max()).nonzero().flatten().tolist()
                bad_rows = (err.max(dim=1).values > 1e-3 * want.abs().max()).nonzero().flatten().tolist()
                print("   bad cols", len(bad_cols), bad_cols[:16], "bad rows", len(bad_rows), bad_rows[:8])
                print("   got", [round(v, 3) for v in C[0, :8].tolist()], "want", [round(v, 3) for v in want[0, :8].tolist()])
        if not ok: break
    if not ok: break
def timed(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
if ok:
    A = torch.randn(4096, 4096, device=dev); B = torch.randn(4096, 4096, device=dev)
    print("4096^3 tb=0 prec6 ms", timed(lambda: ops.gemm(A, B, prec=6)))
print("BMN OK" if ok else "BMN FAILED")
