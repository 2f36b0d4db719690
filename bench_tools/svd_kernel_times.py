"""Per-kernel device time of two sweeps of a batched SVD (torch profiler / CUPTI, not serialised like ncu)."""
import os, sys, subprocess
code = r'''
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops
from torch.profiler import profile, ProfilerActivity
m = int(sys.argv[1]); n = int(sys.argv[2]); batch = int(sys.argv[3])
torch.manual_seed(0)
As = [torch.randn(m, n, device="cuda") * 0.02 for _ in range(batch)]
ops.svd_batched(As, max_sweeps=1); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ops.svd_batched(As, max_sweeps=2); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
'''
for m, n, b in [(4096, 4096, 4), (4096, 11008, 3)]:
    for flag in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", code, str(m), str(n), str(b)], env=dict(os.environ, GRASP_SVD_EVD_WARP=flag),
                           capture_output=True, text=True, timeout=300)
        print(f"=== {m}x{n} batch {b} EVD_WARP={flag}")
        for line in r.stdout.splitlines():
            if "grasp::" in line or "Self CUDA time" in line:
                print(line[:62], line[-75:])
        print(r.stderr[-300:])
