"""GPU check of the tcgen05 split-bf16 GEMM against fp64 (run under `timeout`)."""
import sys, json, time
import torch
sys.path.insert(0, ".")
from grasp_b200 import ops

dev = "cuda"
torch.manual_seed(0)
res = {}

def relf(a, b):
    return (torch.linalg.norm(a.double() - b) / torch.linalg.norm(b)).item()

def timed(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def diag(C, want, tag):
    err = (C.double() - want).abs()
    i = int(err.argmax())
    r, c = divmod(i, C.shape[1])
    print(f"  [{tag}] worst at ({r},{c}): got {C[r,c].item():.6g} want {want[r,c].item():.6g}")
    print("  got  row0[:8]", [round(v, 5) for v in C[0, :8].tolist()])
    print("  want row0[:8]", [round(v, 5) for v in want[0, :8].tolist()])
    bad_rows = (err.max(dim=1).values > 1e-3 * want.abs().max()).nonzero().flatten().tolist()
    bad_cols = (err.max(dim=0).values > 1e-3 * want.abs().max()).nonzero().flatten().tolist()
    print(f"  bad rows {len(bad_rows)} (first {bad_rows[:12]}), bad cols {len(bad_cols)} (first {bad_cols[:12]})")

ok = True
shapes = [(128, 256, 64), (128, 128, 64), (128, 256, 128), (256, 512, 256), (200, 300, 100), (1000, 777, 333),
          (511, 4096, 4096)]
for prec, tol in ((16, 1.5e-6), (3, 3e-5), (6, 2e-6)):
    for (M, N, K) in shapes:
        for ta, tb in ((False, True), (False, False), (True, False), (True, True)):
            A = torch.randn((K, M) if ta else (M, K), device=dev)
            B = torch.randn((N, K) if tb else (K, N), device=dev)
            want = (A.T if ta else A).double() @ (B.T if tb else B).double()
            try:
                C = ops.gemm(A, B, ta=ta, tb=tb, prec=prec)
                torch.cuda.synchronize()
            except Exception as e:
                print(f"prec {prec} {M}x{N}x{K} ta={ta} tb={tb}: EXCEPTION {e}")
                sys.exit(2)
            e = relf(C, want)
            flag = "ok" if e < tol else "FAIL"
            if e >= tol:
                ok = False
            print(f"prec {prec} {M}x{N}x{K} ta={int(ta)} tb={int(tb)}: rel_fro {e:.3e} {flag}", flush=True)
            if e >= tol:
                diag(C, want, f"{M}x{N}x{K}")
            if (M, N, K) == shapes[0] and not (not ta and tb):
                pass
        # alpha/beta path
        A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C0 = torch.randn(M, N, device=dev)
        C = ops.gemm(A, B, tb=True, alpha=0.5, beta=2.0, C_out=C0.clone(), prec=prec)
        want = 0.5 * A.double() @ B.double().T + 2.0 * C0.double()
        e = relf(C, want)
        print(f"prec {prec} {M}x{N}x{K} alpha/beta: rel_fro {e:.3e} {'ok' if e < tol else 'FAIL'}", flush=True)
        ok = ok and e < tol
    if not ok:
        break

if ok:
    # rows with wildly different magnitudes (gradients ~1e-7, activations ~1e2): the row scaling of F16X3 must cope
    for prec, tol in ((16, 2e-6), (6, 2e-6)):
        M, N, K = 300, 500, 700
        A = torch.randn(M, K, device=dev) * torch.logspace(-9, 3, M, device=dev)[:, None]
        B = torch.randn(N, K, device=dev) * torch.logspace(-6, 2, N, device=dev)[:, None]
        want = A.double() @ B.double().T
        C = ops.gemm(A, B, tb=True, prec=prec)
        rowcol = (A.double().norm(dim=1)[:, None] * B.double().norm(dim=1)[None, :])
        e = ((C.double() - want).abs() / rowcol).max().item()
        print(f"prec {prec} graded rows: max elementwise err / (|a_m||b_n|) {e:.3e} {'ok' if e < tol else 'FAIL'}", flush=True)
        ok = ok and e < tol
        Bt = B.T.contiguous()
        C2 = ops.gemm(A, Bt, tb=False, prec=prec)
        e2 = ((C2.double() - want).abs() / rowcol).max().item()
        print(f"prec {prec} graded rows (B given [K,N]): {e2:.3e} {'ok' if e2 < tol else 'FAIL'}", flush=True)
        ok = ok and e2 < tol
if ok:
    # sigma score on the tensor cores vs fp64
    for (o, i) in [(512, 512), (1376, 512), (512, 1376), (4096, 4096)]:
        r = min(o, i)
        U = torch.linalg.qr(torch.randn(o, r, device=dev))[0].contiguous()
        Vh = torch.linalg.qr(torch.randn(i, r, device=dev))[0].T.contiguous()
        G = torch.randn(o, i, device=dev); S = torch.rand(r, device=dev)
        want = ((U.double().T @ G.double()) * Vh.double()).sum(-1)
        for prec in (16, 3, 6):
            g, sc = ops.sigma_score(U, G, Vh, S, prec=prec)
            e = ((g.double() - want).abs().max() / want.abs().max()).item()
            print(f"sigma_score prec {prec} {o}x{i}: max rel {e:.3e}", flush=True)
        g0, _ = ops.sigma_score(U, G, Vh, S, prec=0)
        e = ((g0.double() - want).abs().max() / want.abs().max()).item()
        print(f"sigma_score simt {o}x{i}: max rel {e:.3e}", flush=True)
    # timings
    for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 8192), (511, 11008, 4096)]:
        A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
        for prec in (16, 3, 6):
            t = timed(lambda: ops.gemm(A, B, tb=True, prec=prec))
            res[f"gemm_{M}x{N}x{K}_p{prec}_ms"] = t
            print(f"gemm {M}x{N}x{K} prec {prec}: {t:.3f} ms  {2*M*N*K/t/1e9:.1f} TF/s fp32-equivalent", flush=True)
        t = timed(lambda: A @ B.T)
        print(f"torch fp32 matmul {M}x{N}x{K}: {t:.3f} ms {2*M*N*K/t/1e9:.1f} TF/s", flush=True)
    o = i = 4096
    U = torch.randn(o, o, device=dev); Vh = torch.randn(o, o, device=dev); G = torch.randn(o, o, device=dev); S = torch.rand(o, device=dev)
    for prec in (0, 16, 3, 6):
        t = timed(lambda: ops.sigma_score(U, G, Vh, S, prec=prec))
        print(f"sigma_score 4096^2 prec {prec}: {t:.3f} ms  {2*o*o*o/t/1e9:.1f} TF/s", flush=True)
        res[f"sigma_4096_p{prec}_ms"] = t
json.dump(res, open("gpurun_out/tc_check.json", "w"), indent=1)
print("ALL OK" if ok else "FAILED")
