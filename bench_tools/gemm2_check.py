"""CTA-pair GEMM (GRASP_GEMM_2CTA=1) against the single-CTA kernel: results and time at the in-situ shapes."""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops, _lib
dev = "cuda"
torch.manual_seed(0)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K) in [(300, 520, 1000), (1024, 1024, 512), (8176, 4096, 204), (8176, 208, 4096), (8176, 11008, 298), (8176, 4096, 4096), (8176, 11008, 4096), (8176, 4096, 11008), (8160, 32000, 4096)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.02; dy = torch.randn(M, N, device=dev)
    xo, wo, dyo = ops.split_f16(x), ops.split_f16(w, _lib.SCALE_TENSOR), ops.split_f16(dy)
    y = ops.gemm_planes(xo, wo); dx = ops.gemm_planes(dyo, wo, b_kn=True)
    if M * N * K < 3e10:
        e1 = ((y.double() - x.double() @ w.double().t()).abs().max() / y.abs().max()).item()
        e2 = ((dx.double() - dy.double() @ w.double()).abs().max() / dx.abs().max()).item()
    else:
        e1 = e2 = float("nan")
    ms1 = t(lambda: ops.gemm_planes(xo, wo)); ms2 = t(lambda: ops.gemm_planes(dyo, wo, b_kn=True))
    fl = 2.0 * M * N * K
    print(f"{M}x{N}x{K}: xWt {ms1:.3f} ms {fl/ms1/1e9:.0f} TF/s err {e1:.1e} | dyW {ms2:.3f} ms {fl/ms2/1e9:.0f} TF/s err {e2:.1e}", flush=True)
'''
for flag in ("0", "1"):
    print("GRASP_GEMM_2CTA=" + flag, flush=True)
    try:
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GRASP_GEMM_2CTA=flag), capture_output=True, text=True, timeout=90)
        print(r.stdout, r.stderr[-600:], flush=True)
    except subprocess.TimeoutExpired as e:
        print("TIMEOUT", (e.stdout or b"")[-500:], flush=True)
