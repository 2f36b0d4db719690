"""K blocks per TMEM drain (GRASP_GEMM_KGROUP): accuracy against fp64 and time at the in-situ shapes."""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops, _lib
dev = "cuda"
torch.manual_seed(0)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K) in [(1024, 1024, 4096), (1024, 1024, 11008), (512, 2048, 32000)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.02
    x[:, ::7] = x[:, ::7].abs()          # a non-zero mean makes the partial sums grow (worst case for truncation)
    w[:, ::7] = w[:, ::7].abs()
    y = ops.gemm_planes(ops.split_f16(x), ops.split_f16(w, _lib.SCALE_TENSOR))
    ref = x.double() @ w.double().t()
    y32 = x @ w.t()
    e = ((y.double() - ref).abs().max() / ref.abs().max()).item()
    bias = ((y.double() - ref).mean() / ref.abs().mean()).item()
    e32 = ((y32.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"acc {M}x{N}x{K}: max rel err {e:.2e} (mean signed {bias:+.1e}); torch fp32 matmul {e32:.2e}", flush=True)
for (M, N, K) in [(8176, 4096, 4096), (8176, 11008, 4096), (8176, 4096, 11008), (8160, 32000, 4096), (8176, 4096, 204), (8176, 11008, 298)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.02; dy = torch.randn(M, N, device=dev)
    xo, wo, dyo = ops.split_f16(x), ops.split_f16(w, _lib.SCALE_TENSOR), ops.split_f16(dy)
    ms1 = t(lambda: ops.gemm_planes(xo, wo)); ms2 = t(lambda: ops.gemm_planes(dyo, wo, b_kn=True))
    fl = 2.0 * M * N * K
    print(f"{M}x{N}x{K}: xWt {ms1:.3f} ms {fl/ms1/1e9:.0f} TF/s | dyW {ms2:.3f} ms {fl/ms2/1e9:.0f} TF/s", flush=True)
'''
for kg in ("1", "2", "4", "8"):
    print("GRASP_GEMM_KGROUP=" + kg, flush=True)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GRASP_GEMM_KGROUP=kg), capture_output=True, text=True, timeout=200)
    print(r.stdout, r.stderr[-400:], flush=True)
