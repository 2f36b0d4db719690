"""Prepared-operand GEMM at the shapes of the calibration passes (16 samples x 511 tokens), under the epilogue /
tile-order / drain-group switches of gemm_tc.cu.  Prints time, fp32-equivalent TFLOP/s and, for the short-K
shapes, the achieved fraction of the HBM-write bound of the output."""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops, _lib
dev = "cuda"
torch.manual_seed(0)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K) in [(8176, 4096, 204), (8176, 204, 4096), (8176, 11008, 298), (8176, 298, 4096), (8176, 4096, 298),
                  (8176, 298, 11008), (8176, 4096, 4096), (8176, 11008, 4096), (8176, 4096, 11008), (8160, 32000, 4096)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.02; dy = torch.randn(M, N, device=dev)
    xo, wo, dyo = ops.split_f16(x), ops.split_f16(w, _lib.SCALE_TENSOR), ops.split_f16(dy)
    y = ops.gemm_planes(xo, wo); dx = ops.gemm_planes(dyo, wo, b_kn=True)
    if M * N * K < 6e11:
        e1 = ((y.double() - x.double() @ w.double().t()).abs().max() / y.abs().max()).item()
        e2 = ((dx.double() - dy.double() @ w.double()).abs().max() / dx.abs().max()).item()
    else:
        e1 = e2 = float("nan")
    ms1 = t(lambda: ops.gemm_planes(xo, wo)); ms2 = t(lambda: ops.gemm_planes(dyo, wo, b_kn=True))
    fl = 2.0 * M * N * K
    wb1, wb2 = 4.0 * M * N / 6.5252e12 * 1e3, 4.0 * M * K / 6.5252e12 * 1e3          # ms to write the output at the copy peak
    print(f"{M}x{N}x{K}: xWt {ms1*1e3:7.1f} us {fl/ms1/1e9:5.0f} TF/s (write bound {wb1*1e3:5.1f} us) err {e1:.1e} | "
          f"dyW {ms2*1e3:7.1f} us {fl/ms2/1e9:5.0f} TF/s (write bound {wb2*1e3:5.1f} us) err {e2:.1e}", flush=True)
'''
variants = [("round-1 behaviour", dict(GRASP_GEMM_STAGED="0", GRASP_GEMM_GROUP_M="100000", GRASP_GEMM_KGROUP_SHORT="1")),
            ("staged epilogue only", dict(GRASP_GEMM_STAGED="1", GRASP_GEMM_GROUP_M="100000", GRASP_GEMM_KGROUP_SHORT="1")),
            ("staged + super-row tile order", dict(GRASP_GEMM_STAGED="1", GRASP_GEMM_KGROUP_SHORT="1")),
            ("default (staged, super-rows, 2 K blocks per drain at K <= 512)", dict()),
            ("default, 256-wide tiles forced", dict(GRASP_GEMM_BN="256")),
            ("default, 128-wide tiles forced", dict(GRASP_GEMM_BN="128"))]
if len(sys.argv) > 1:
    variants = [v for v in variants if any(a in v[0] for a in sys.argv[1:])]
for name, env in variants:
    print("== " + name + "  " + " ".join(f"{k}={v}" for k, v in env.items()), flush=True)
    try:
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=240)
        print(r.stdout, r.stderr[-800:], flush=True)
    except subprocess.TimeoutExpired as e:
        print("TIMEOUT", (e.stdout or b"")[-500:], flush=True)
