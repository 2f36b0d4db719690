"""Register-resident eigen-solve (GRASP_SVD_EVD_WARP=1, default) vs the 1024-thread kernel: accuracy and time."""
import os, sys, subprocess
code = r'''
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops
m = int(sys.argv[1]); n = int(sys.argv[2]); batch = int(sys.argv[3])
torch.manual_seed(0)
As = [torch.randn(m, n, device="cuda") * 0.02 for _ in range(batch)]
ops.svd_batched(As[:1]); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); outs, info = ops.svd_batched(As, return_info=True); e1.record(); torch.cuda.synchronize()
U, S, Vh = outs[-1]; A = As[-1]
Sref = torch.linalg.svdvals(A.double())
r = min(m, n)
print(f"{m}x{n} batch={batch} per_matrix_ms={e0.elapsed_time(e1)/batch:.1f} sweeps={info[:,0].tolist()} tc={info[:,3].tolist()} conv={info[:,1].tolist()} sigma={((S-Sref).abs().max()/Sref[0]).item():.2e} recon={(torch.linalg.norm((U*S)@Vh-A)/torch.linalg.norm(A)).item():.2e} orthU={(U.T@U-torch.eye(r,device='cuda')).abs().max().item():.2e} orthV={(Vh@Vh.T-torch.eye(r,device='cuda')).abs().max().item():.2e}")
'''
shapes = [(1024, 1024, 2), (4096, 4096, 4), (4096, 11008, 3)]
VAR = sys.argv[1] if len(sys.argv) > 1 else "GRASP_SVD_EVD_WARP"
for m, n, b in shapes:
    for flag in ("0", "1"):
        try:
            r = subprocess.run([sys.executable, "-c", code, str(m), str(n), str(b)], env=dict(os.environ, **{VAR: flag}),
                               capture_output=True, text=True, timeout=300)
            print(VAR + "=" + flag, r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
        except subprocess.TimeoutExpired:
            print("EVD_WARP=" + flag, "TIMEOUT", flush=True)
