import os, sys, subprocess
code = r'''
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops
n = int(sys.argv[1])
torch.manual_seed(0)
A = torch.randn(n, n, device="cuda") * 0.02
outs, info = ops.svd_batched([A], return_info=True); torch.cuda.synchronize()
U, S, Vh = outs[0]
Sref = torch.linalg.svdvals(A.double())
print(f"n={n} sweeps={info[:,0].tolist()} sigma={((S-Sref).abs().max()/Sref[0]).item():.2e} recon={(torch.linalg.norm((U*S)@Vh-A)/torch.linalg.norm(A)).item():.2e} orthU={(U.T@U-torch.eye(n,device='cuda')).abs().max().item():.2e} orthV={(Vh@Vh.T-torch.eye(n,device='cuda')).abs().max().item():.2e}")
'''
for label, env in (("default", {}), ("no_cleanup", {"GRASP_SVD_NO_CLEANUP": "1"}), ("no_cleanup+evd64", {"GRASP_SVD_NO_CLEANUP": "1", "GRASP_SVD_EVD64": "1"}), ("evd64", {"GRASP_SVD_EVD64": "1"})):
    r = subprocess.run([sys.executable, "-c", code, sys.argv[1] if len(sys.argv) > 1 else "2048"], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
    print(label, r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
