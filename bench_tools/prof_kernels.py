"""Short driver for ncu --set full captures of the individual kernels (one GPU)."""
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import ops
what = sys.argv[1]
dev = "cuda"
torch.manual_seed(0)
if what == "gemm":
    A = torch.randn(4096, 4096, device=dev); B = torch.randn(4096, 4096, device=dev)
    for _ in range(3): ops.gemm(A, B, tb=True, prec=6)
    for _ in range(3): ops.gemm(A, B, tb=True, prec=3)
elif what == "sigma":
    U = torch.randn(4096, 4096, device=dev); Vh = torch.randn(4096, 4096, device=dev)
    G = torch.randn(4096, 4096, device=dev); S = torch.rand(4096, device=dev)
    for _ in range(3): ops.sigma_score(U, G, Vh, S, prec=6)
elif what == "bi":
    hs = [torch.randn(8, 511, 4096, device=dev) for _ in range(33)]
    acc = torch.zeros(32, dtype=torch.float64, device=dev)
    for _ in range(4): ops.bi_chain(hs, acc)
elif what == "planes":
    # the GEMM of the calibration passes at its in-situ shapes (16 samples x 511 tokens), prepared operands
    from grasp_b200 import _lib
    T = 8176
    x = torch.randn(T, 4096, device=dev); w = torch.randn(11008, 4096, device=dev) * 0.02
    dy = torch.randn(T, 11008, device=dev)
    for _ in range(2):
        xo, wo, dyo = ops.split_f16(x), ops.split_f16(w, _lib.SCALE_TENSOR), ops.split_f16(dy)
        ops.gemm_planes(xo, wo)                 # x W^T   (K-major B)
        ops.gemm_planes(dyo, wo, b_kn=True)     # dy W    (MN-major B)
elif what == "shortk":
    # the rank-k GEMMs of a compressed layer: t[T,k] * OutW[11008,k]^T (short K) and x[T,4096] * InW[k,4096]^T (skinny N)
    from grasp_b200 import _lib
    T = 8176
    t_ = torch.randn(T, 298, device=dev); w = torch.randn(11008, 298, device=dev) * 0.02
    x = torch.randn(T, 4096, device=dev); wi = torch.randn(204, 4096, device=dev) * 0.02
    to, wo, xo, wio = ops.split_f16(t_), ops.split_f16(w, _lib.SCALE_TENSOR), ops.split_f16(x), ops.split_f16(wi, _lib.SCALE_TENSOR)
    for _ in range(3):
        ops.gemm_planes(to, wo)
        ops.gemm_planes(xo, wio)
elif what == "rowops":
    T = 8176
    x = torch.randn(T, 4096, device=dev); w = torch.ones(4096, device=dev)
    g = torch.randn(T, 11008, device=dev); u = torch.randn(T, 11008, device=dev)
    lg = torch.randn(T, 32000, device=dev); lab = torch.randint(0, 32000, (T,), device=dev)
    cos = torch.randn(1, 511, 128, device=dev); sin = torch.randn(1, 511, 128, device=dev)
    for _ in range(2):
        y, r = ops.rmsnorm_fwd(x, w, 1e-5); ops.rmsnorm_bwd(y, x, w, r, add=x)
        ops.rope_(y, 511, 32, 128, cos, sin)
        h = ops.swiglu_fwd(g, u); ops.swiglu_bwd(h, g.clone(), u.clone(), inplace=True)
        ops.ce_loss_bwd_(lg.clone(), lab, torch.ones(T, device=dev))
elif what == "svd4":
    # one batch of four 4096 x 4096 matrices, two sweeps (what the 7B job's attention projections look like)
    As = [torch.randn(4096, 4096, device=dev) * 0.02 for _ in range(4)]
    ops.svd_batched(As, max_sweeps=1)
elif what == "svd8":
    # the job's batches: eight 4096 x 4096 matrices (attention projections of two layers), one sweep
    As = [torch.randn(4096, 4096, device=dev) * 0.02 for _ in range(8)]
    ops.svd_batched(As, max_sweeps=1)
elif what == "attn":
    import math
    from grasp_b200 import _lib
    B, S, H, D = 16, 511, 32, 128
    q = torch.randn(B * S, H * D, device=dev); k = torch.randn(B * S, H * D, device=dev); v = torch.randn(B * S, H * D, device=dev)
    do = torch.randn(B * S, H * D, device=dev) * 1e-3
    cos = torch.randn(1, S, D, device=dev); sin = torch.randn(1, S, D, device=dev)
    for _ in range(2):
        qo, ko, vo = ops.attn_prep_qkv(q, k, v, S, H, H, D, cos, sin)
        out, ctx = ops.attn_fwd_prepared(qo, ko, vo, B, S, H, H, D, 1 / math.sqrt(D))
        ops.attn_bwd(ctx, do)
elif what == "svd":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    A = torch.randn(n, n, device=dev) * 0.02
    ops.svd_batched([A], max_sweeps=2)
torch.cuda.synchronize()
print("done", what)
