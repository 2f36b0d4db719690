"""Attention of one calibration micro-batch at the 7B shape (16 samples x 511 tokens, 32 heads x 128) and at the
Llama-3-8B shape (32 / 8 heads): grasp_attn_fwd / bwd (operand splits included and separately) against torch's fp32
scaled_dot_product_attention + autograd on the same B200."""
import math, sys, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from grasp_b200 import ops, _lib
dev = "cuda"
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (B, S, H, Hkv, D) in [(16, 511, 32, 32, 128), (16, 511, 32, 8, 128), (16, 511, 32, 4, 64)]:
    q = torch.randn(B * S, H * D, device=dev); k = torch.randn(B * S, Hkv * D, device=dev); v = torch.randn(B * S, Hkv * D, device=dev)
    do = torch.randn(B * S, H * D, device=dev) * 1e-3
    scale = 1 / math.sqrt(D)
    out, ctx = ops.attn_fwd(q, k, v, B, S, H, Hkv, D, scale)
    us_f = t(lambda: ops.attn_fwd(q, k, v, B, S, H, Hkv, D, scale))
    us_b = t(lambda: ops.attn_bwd(ctx, do))
    us_split = t(lambda: ops.split_f16(q, _lib.SCALE_TENSOR))
    def torch_fwd(keep=False):
        q4 = q.view(B, S, H, D).transpose(1, 2).detach().requires_grad_(keep)
        k4 = k.view(B, S, Hkv, D).transpose(1, 2)[:, :, None].expand(B, Hkv, H // Hkv, S, D).reshape(B, H, S, D).detach().requires_grad_(keep)
        v4 = v.view(B, S, Hkv, D).transpose(1, 2)[:, :, None].expand(B, Hkv, H // Hkv, S, D).reshape(B, H, S, D).detach().requires_grad_(keep)
        o = F.scaled_dot_product_attention(q4, k4, v4, is_causal=True, scale=scale)
        return o, (q4, k4, v4)
    with torch.no_grad():
        us_tf = t(lambda: torch_fwd())
    o_t, (q4, k4, v4) = torch_fwd(True)
    do4 = do.view(B, S, H, D).transpose(1, 2)
    us_tb = t(lambda: torch.autograd.grad(o_t, (q4, k4, v4), do4, retain_graph=True))
    err = ((out.view(B, S, H, D).transpose(1, 2) - o_t).abs().max() / o_t.abs().max()).item()
    fl = 2.0 * B * H * S * S * D
    print(f"B={B} S={S} H={H}/{Hkv} D={D}: fwd {us_f:7.1f} us ({fl / us_f / 1e6:5.0f} TF/s causal-useful), bwd {us_b:7.1f} us; one operand split "
          f"{us_split:5.1f} us | torch sdpa fp32 fwd {us_tf:7.1f} us, bwd {us_tb:7.1f} us | out vs torch {err:.1e}", flush=True)
