import os, sys, subprocess, json
code = r'''
import sys, torch, time
sys.path.insert(0, ".")
from grasp_b200 import ops
n = int(sys.argv[1]); batch = int(sys.argv[2])
As = [torch.randn(n, n, device="cuda") * 0.02 for _ in range(batch)]
ops.svd_batched(As[:1]); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); outs, info = ops.svd_batched(As, return_info=True); e1.record(); torch.cuda.synchronize()
U, S, Vh = outs[0]; A = As[0]
rec = (torch.linalg.norm((U * S) @ Vh - A) / torch.linalg.norm(A)).item()
orth = (U.T @ U - torch.eye(n, device="cuda")).abs().max().item()
print(f"n={n} batch={batch} ms={e0.elapsed_time(e1):.1f} per_matrix={e0.elapsed_time(e1)/batch:.1f} sweeps={info[:,0].tolist()} conv={info[:,1].tolist()} recon={rec:.2e} orthU={orth:.2e}")
'''
for n, batch in ((1024, 1), (2048, 1), (2048, 4)):
    for cap in (2, 4, 8):
        env = dict(os.environ, GRASP_SVD_INNER_CAP=str(cap), GRASP_SVD_TOL="1e-6")
        r = subprocess.run([sys.executable, "-c", code, str(n), str(batch)], env=env, capture_output=True, text=True, timeout=300)
        print(f"cap={cap}", r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
env = dict(os.environ, GRASP_SVD_INNER_CAP="4", GRASP_SVD_TOL="1e-6")
for n, batch in ((4096, 1), (4096, 4)):
    r = subprocess.run([sys.executable, "-c", code, str(n), str(batch)], env=env, capture_output=True, text=True, timeout=300)
    print("cap=4", r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
