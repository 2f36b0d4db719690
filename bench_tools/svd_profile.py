import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = torch.randn(n, n, device="cuda") * 0.02
ops.svd_batched([A]); torch.cuda.synchronize()
