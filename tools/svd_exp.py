import os, sys, subprocess
code = r'''
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops
n = int(sys.argv[1]); batch = int(sys.argv[2])
torch.manual_seed(0)
As = [torch.randn(n, n, device="cuda") * 0.02 for _ in range(batch)]
ops.svd_batched(As[:1]); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); outs, info = ops.svd_batched(As, return_info=True); e1.record(); torch.cuda.synchronize()
U, S, Vh = outs[-1]; A = As[-1]
Sref = torch.linalg.svdvals(A.double())
print(f"n={n} batch={batch} per_matrix_ms={e0.elapsed_time(e1)/batch:.1f} sweeps={info[:,0].tolist()} tc={info[:,3].tolist()} conv={info[:,1].tolist()} sigma={((S-Sref).abs().max()/Sref[0]).item():.2e} recon={(torch.linalg.norm((U*S)@Vh-A)/torch.linalg.norm(A)).item():.2e} orthU={(U.T@U-torch.eye(n,device='cuda')).abs().max().item():.2e} orthV={(Vh@Vh.T-torch.eye(n,device='cuda')).abs().max().item():.2e}")
'''
exps = [("default", {}), ("simt_cleanup", {"GRASP_SVD_TC_CLEANUP": "0"}), ("tc_cap1", {"GRASP_SVD_TC_INNER_CAP": "1"}), ("tc_cap3", {"GRASP_SVD_TC_INNER_CAP": "3"})]
for n, b in ((2048, 1), (4096, 4)):
    for label, env in exps:
        r = subprocess.run([sys.executable, "-c", code, str(n), str(b)], env=dict(os.environ, **env), capture_output=True, text=True, timeout=400)
        print(label, r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
