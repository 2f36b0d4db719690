import os, sys, subprocess
code = r'''
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops
n = int(sys.argv[1]); m = int(sys.argv[2]); batch = int(sys.argv[3])
torch.manual_seed(0)
As = [torch.randn(m, n, device="cuda") * 0.02 for _ in range(batch)]
ops.svd_batched(As[:1]); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); outs, info = ops.svd_batched(As, return_info=True); e1.record(); torch.cuda.synchronize()
U, S, Vh = outs[-1]; A = As[-1]
r = min(m, n)
Sref = torch.linalg.svdvals(A.double())
rec = (torch.linalg.norm((U * S) @ Vh - A) / torch.linalg.norm(A)).item()
orth = (U.T @ U - torch.eye(r, device="cuda")).abs().max().item()
orthv = (Vh @ Vh.T - torch.eye(r, device="cuda")).abs().max().item()
serr = ((S - Sref).abs().max() / Sref[0]).item()
print(f"{m}x{n} batch={batch} ms={e0.elapsed_time(e1):.1f} per_matrix={e0.elapsed_time(e1)/batch:.1f} sweeps={info[:,0].tolist()} conv={info[:,1].tolist()} sigma={serr:.2e} recon={rec:.2e} orthU={orth:.2e} orthV={orthv:.2e}")
'''
cases = [(64, 64, 1), (256, 256, 1), (704, 256, 1), (256, 704, 2), (192, 192, 1), (1024, 1024, 1), (2048, 2048, 1)]
for tc in ("0", "1"):
    for (n, m, b) in cases:
        env = dict(os.environ, GRASP_SVD_TC=tc)
        r = subprocess.run([sys.executable, "-c", code, str(n), str(m), str(b)], env=env, capture_output=True, text=True, timeout=300)
        print(f"TC={tc}", r.stdout.strip(), r.stderr.strip()[-400:], flush=True)
for (n, m, b) in [(4096, 4096, 1), (4096, 4096, 4), (4096, 11008, 3)]:
    env = dict(os.environ, GRASP_SVD_TC="1")
    r = subprocess.run([sys.executable, "-c", code, str(n), str(m), str(b)], env=env, capture_output=True, text=True, timeout=400)
    print("TC=1", r.stdout.strip(), r.stderr.strip()[-400:], flush=True)
