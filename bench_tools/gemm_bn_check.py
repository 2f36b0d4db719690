"""Tile width (GRASP_GEMM_BN = 128 / 256 / model) at the short-K shapes of compressed layers."""
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
for (M, N, K) in [(8176, 4096, 204), (8176, 204, 4096), (8176, 11008, 298), (8176, 298, 4096), (8176, 4096, 298), (8176, 298, 11008), (8176, 4096, 4096)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.02; dy = torch.randn(M, N, device=dev)
    xo, wo, dyo = ops.split_f16(x), ops.split_f16(w, _lib.SCALE_TENSOR), ops.split_f16(dy)
    ms1 = t(lambda: ops.gemm_planes(xo, wo)); ms2 = t(lambda: ops.gemm_planes(dyo, wo, b_kn=True))
    print(f"{M}x{N}x{K}: xWt {ms1*1e3:.0f} us | dyW (K={N}) {ms2*1e3:.0f} us", flush=True)
'''
for bn in ("0", "128", "256"):
    print("GRASP_GEMM_BN=" + bn, flush=True)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GRASP_GEMM_BN=bn), capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr[-400:], flush=True)
