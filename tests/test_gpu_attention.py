"""GPU: the tensor-core causal attention (grasp_attn_fwd / grasp_attn_bwd through the C ABI) against fp64 torch
autograd -- the attention the reference reaches through transformers' LlamaAttention inside model(...) and
loss.backward() (modeling_grasp.py:347-354): grouped-query heads, ragged sequence lengths, both head dimensions."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def reference(q, k, v, d_out, B, S, H, Hkv, D, scale):
    q4 = q.double().view(B, S, H, D).transpose(1, 2).detach().requires_grad_(True)
    k4 = k.double().view(B, S, Hkv, D).transpose(1, 2).detach().requires_grad_(True)
    v4 = v.double().view(B, S, Hkv, D).transpose(1, 2).detach().requires_grad_(True)
    rep = H // Hkv
    kk = k4.repeat_interleave(rep, dim=1)
    vv = v4.repeat_interleave(rep, dim=1)
    s = (q4 @ kk.transpose(-1, -2)) * scale
    mask = torch.ones(S, S, dtype=torch.bool, device=q.device).tril()
    s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ vv).transpose(1, 2).reshape(B * S, H * D)
    o.backward(d_out.double())
    back = lambda g, heads: g.transpose(1, 2).reshape(B * S, heads * D)
    lse2 = torch.logsumexp(s, dim=-1) / math.log(2.0)
    return o.detach(), back(q4.grad, H), back(k4.grad, Hkv), back(v4.grad, Hkv), lse2.reshape(-1)


@pytest.mark.parametrize("B,S,H,Hkv,D", [(2, 511, 4, 4, 128), (1, 130, 4, 2, 64), (2, 64, 2, 1, 128), (1, 700, 2, 2, 128),
                                         (3, 1, 2, 2, 64), (1, 257, 8, 2, 128)])
def test_attention_forward_and_backward_match_fp64(cuda, parity_log, B, S, H, Hkv, D):
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + S + H + D)
    # activations with the spread of real ones: a few loud heads / tokens, small gradients
    q = (torch.randn(B * S, H * D, generator=g) * torch.logspace(-1, 0.5, H * D)[None, :]).to(cuda)
    k = torch.randn(B * S, Hkv * D, generator=g).to(cuda)
    v = (torch.randn(B * S, Hkv * D, generator=g) * 0.3).to(cuda)
    d_out = (torch.randn(B * S, H * D, generator=g) * 1e-4).to(cuda)
    scale = 1.0 / math.sqrt(D)
    out, ctx = ops.attn_fwd(q, k, v, B, S, H, Hkv, D, scale)
    dq, dk, dv = ops.attn_bwd(ctx, d_out)
    o_ref, dq_ref, dk_ref, dv_ref, lse_ref = reference(q, k, v, d_out, B, S, H, Hkv, D, scale)
    errs = {"out": rel(out, o_ref), "dq": rel(dq, dq_ref), "dk": rel(dk, dk_ref), "dv": rel(dv, dv_ref)}
    lse_err = (ctx[4].double() - lse_ref).abs().max().item()
    parity_log(f"attention B={B} S={S} H={H}/{Hkv} D={D}: " + ", ".join(f"{k_} {e:.1e}" for k_, e in errs.items())
               + f", lse2 abs {lse_err:.1e}")
    assert torch.isfinite(out).all() and torch.isfinite(dq).all() and torch.isfinite(dk).all() and torch.isfinite(dv).all()
    assert errs["out"] < 2e-6 and lse_err < 1e-5
    if S == 1:          # a single key: dq = dk = 0 exactly, nothing to be relative to
        assert dq.abs().max().item() < 1e-9 and dk.abs().max().item() < 1e-9 and errs["dv"] < 1e-5
    else:
        assert errs["dq"] < 1e-5 and errs["dk"] < 1e-5 and errs["dv"] < 1e-5


def test_attention_with_massive_activations(cuda, parity_log):
    """Trained LLaMAs carry a few hidden dimensions / tokens 100-1000x larger than the rest.  The operands carry one
    power-of-two scale per tensor, so the small entries keep fewer bits; the outputs must still be fp32-class
    relative to their own scale (error measured against the largest entry of each result)."""
    from grasp_b200 import ops
    B, S, H, Hkv, D = 1, 300, 4, 2, 128
    g = torch.Generator().manual_seed(77)
    q = torch.randn(B * S, H * D, generator=g)
    k = torch.randn(B * S, Hkv * D, generator=g)
    v = torch.randn(B * S, Hkv * D, generator=g)
    q[:, 5] *= 10.0; k[:, 5] *= 6.0                     # an outlier dimension shared by q and k (logits up to ~+-50)
    v[0] *= 1000.0                                      # the attention-sink token
    v[:, 200] *= 300.0
    d_out = torch.randn(B * S, H * D, generator=g) * 1e-3
    d_out[17] *= 500.0
    q, k, v, d_out = (t.to(cuda) for t in (q, k, v, d_out))
    scale = 1.0 / math.sqrt(D)
    out, ctx = ops.attn_fwd(q, k, v, B, S, H, Hkv, D, scale)
    dq, dk, dv = ops.attn_bwd(ctx, d_out)
    o_ref, dq_ref, dk_ref, dv_ref, _ = reference(q, k, v, d_out, B, S, H, Hkv, D, scale)
    errs = {"out": rel(out, o_ref), "dq": rel(dq, dq_ref), "dk": rel(dk, dk_ref), "dv": rel(dv, dv_ref)}
    parity_log("attention with 100-1000x outliers: " + ", ".join(f"{n} {e:.1e}" for n, e in errs.items()))
    assert all(torch.isfinite(t).all() for t in (out, dq, dk, dv))
    # logits of magnitude ~50 carry an absolute fp32-class error of ~50 * 3e-7, i.e. ~1.5e-5 relative in the probabilities
    assert errs["out"] < 1e-4 and max(errs["dq"], errs["dk"], errs["dv"]) < 5e-4


@pytest.mark.parametrize("B,S,H,Hkv,D", [(2, 70, 4, 2, 64), (1, 511, 2, 2, 128)])
def test_rope_and_operand_planes_in_one_preparation(cuda, B, S, H, Hkv, D):
    """grasp_attn_prep_qkv (RoPE + tensor-scaled planes in two passes, fp32 q / k untouched) feeds the same attention
    as grasp_rope_inplace followed by three operand splits."""
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(S + D)
    q = torch.randn(B * S, H * D, generator=g).to(cuda)
    k = torch.randn(B * S, Hkv * D, generator=g).to(cuda)
    v = torch.randn(B * S, Hkv * D, generator=g).to(cuda)
    pos = torch.arange(S, dtype=torch.float32)
    freq = 1.0 / (10000.0 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    ang = torch.cat([pos[:, None] * freq[None, :]] * 2, dim=-1)
    cos, sin = ang.cos()[None].to(cuda), ang.sin()[None].to(cuda)
    q0, k0 = q.clone(), k.clone()
    qo, ko, vo = ops.attn_prep_qkv(q, k, v, S, H, Hkv, D, cos, sin)
    assert torch.equal(q, q0) and torch.equal(k, k0)
    out, _ = ops.attn_fwd_prepared(qo, ko, vo, B, S, H, Hkv, D, D ** -0.5)
    qr, kr = ops.rope_(q.clone(), S, H, D, cos, sin), ops.rope_(k.clone(), S, Hkv, D, cos, sin)
    ref, _ = ops.attn_fwd(qr, kr, v, B, S, H, Hkv, D, D ** -0.5)
    assert rel(out, ref) < 1e-6
    # per-batch cos / sin rows
    cosb, sinb = cos.expand(B, S, D).contiguous(), sin.expand(B, S, D).contiguous()
    qo2, ko2, vo2 = ops.attn_prep_qkv(q, k, v, S, H, Hkv, D, cosb, sinb)
    out2, _ = ops.attn_fwd_prepared(qo2, ko2, vo2, B, S, H, Hkv, D, D ** -0.5)
    assert torch.equal(out, out2)


def test_attention_rejects_unsupported_shapes(cuda):
    from grasp_b200 import _lib, ops
    q = torch.zeros(8, 2 * 32, device=cuda)
    with pytest.raises(_lib.GraspLibraryError):
        ops.attn_fwd(q, q, q, 1, 8, 2, 2, 32, 1.0)             # head_dim 32: the caller keeps torch's kernel for it
    with pytest.raises(ValueError):
        ops.attn_fwd(q, q, q, 1, 8, 2, 2, 64, 1.0)             # shapes do not match the declared layout
    assert ops.attn_supported(128) and not ops.attn_supported(32)


def test_fused_layer_uses_the_attention_kernels_and_matches_autograd(cuda):
    """A decoder layer with 64-wide heads through fused.FusedLlama (own attention) against transformers + autograd."""
    import copy
    from grasp_b200 import synth
    from modeling_grasp import GRASPModel
    model = synth.random_llama("small", seed=21, num_attention_heads=4, num_key_value_heads=2).to(cuda)   # 256 / 4 = 64
    assert model.config.hidden_size // model.config.num_attention_heads == 64
    tokens = synth.random_tokens(5, 70, model.config.vocab_size, seed=4)
    grads = {}
    for fused in (True, False):
        gm = GRASPModel(copy.deepcopy(model))
        gm.micro_batch = 3
        gm._engine_runner().use_fused = fused
        dl = synth.calibration_dataloader(0, 0, 0, tokens=tokens)
        gm.compress_block(2, "attention", ["q_proj", "k_proj", "v_proj", "o_proj"], device=cuda)
        grads[fused] = gm.get_svdlayer_gradients(dl, cuda)
    for name in grads[True]:
        assert rel(grads[True][name], grads[False][name]) < 2e-4, name
