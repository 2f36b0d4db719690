"""GPU: the prepared-operand GEMM, the decoder-layer row kernels (through the C ABI) and the explicit
forward/backward built on them, against fp64 torch / autograd and against the transformers + autograd route
the reference takes (modeling_grasp.py:347-354)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------- prepared operands
@pytest.mark.parametrize("M,N,K", [(300, 204, 1000), (1024, 4096, 4096), (511, 11008, 4096), (129, 72, 20), (64, 8, 8),
                                   (2300, 520, 298),      # short K (two K blocks per TMEM drain), ragged M and N tiles
                                   (2300, 298, 1028),     # N % 4 != 0: the direct (unstaged) epilogue
                                   (2049, 4096, 4096)])   # more M blocks than one super-row of the tile order
def test_gemm_planes_forward_and_backward_forms(cuda, M, N, K):
    from grasp_b200 import _lib, ops
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) * 0.02).to(cuda)
    dy = torch.randn(M, N, generator=g).to(cuda)
    xo, dyo = ops.split_f16(x), ops.split_f16(dy)
    wo = ops.split_f16(w, _lib.SCALE_TENSOR)
    wr = ops.split_f16(w, _lib.SCALE_ROWS)
    y_ref = x.double() @ w.double().t()
    dx_ref = dy.double() @ w.double()
    tol = 2e-6                                   # fp32-class (the arithmetic itself measures 3e-7 at K = 4096)
    assert rel(ops.gemm_planes(xo, wo), y_ref) < tol
    assert rel(ops.gemm_planes(xo, wr), y_ref) < tol
    assert rel(ops.gemm_planes(dyo, wo, b_kn=True), dx_ref) < tol
    # beta = 1 accumulates into C (the q/k/v and gate/up gradient sums of the backward)
    c = torch.randn(M, K, generator=g).to(cuda)
    got = ops.gemm_planes(dyo, wo, b_kn=True, beta=1.0, C_out=c.clone())
    assert rel(got, dx_ref + c.double()) < tol
    with pytest.raises(ValueError):
        ops.gemm_planes(dyo, wr, b_kn=True)      # a [K, N] operand needs one scale for the whole tensor


@pytest.mark.parametrize("M,K,k,N", [(300, 1000, 204, 520), (1030, 4096, 298, 4096), (129, 72, 22, 64), (2300, 11008, 298, 4096)])
def test_factor_pair_without_an_fp32_intermediate(cuda, M, K, k, N):
    """SVDLinear forward OutLinear(InLinear(x)) (reference modeling_grasp.py:57-59) and its backward
    dy OutW InW: the first GEMM hands its [tokens, k] result over as operand planes (grasp_gemm_f16x3_planes_out),
    the second consumes them -- against fp64, and against the route through an fp32 intermediate + split."""
    from grasp_b200 import _lib, ops
    g = torch.Generator().manual_seed(M + K + k + N)
    x = torch.randn(M, K, generator=g).to(cuda)
    in_w = (torch.randn(k, K, generator=g) * 0.02).to(cuda)        # InLinear.weight  [k, in]
    out_w = (torch.randn(N, k, generator=g) * 0.05).to(cuda)       # OutLinear.weight [out, k]
    dy = (torch.randn(M, N, generator=g) * 1e-3).to(cuda)
    xo, dyo = ops.split_f16(x), ops.split_f16(dy)
    wi, wo = ops.split_f16(in_w, _lib.SCALE_TENSOR), ops.split_f16(out_w, _lib.SCALE_TENSOR)
    # forward
    to = ops.gemm_planes_to_operand(xo, wi)
    assert (to.rows, to.cols, to.mode) == (M, k, _lib.SCALE_ROWS)
    y = ops.gemm_planes(to, wo)
    y_ref = (x.double() @ in_w.double().t()) @ out_w.double().t()
    y_two = ops.gemm_planes(ops.split_f16(ops.gemm_planes(xo, wi)), wo)
    assert rel(y, y_ref) < 2e-6 and rel(y, y_two) < 2e-6
    # backward: dx = (dy OutW) InW, both weights as [K, N] operands
    dto = ops.gemm_planes_to_operand(dyo, wo, b_kn=True)
    dx = ops.gemm_planes(dto, wi, b_kn=True)
    dx_ref = (dy.double() @ out_w.double()) @ in_w.double()
    assert rel(dx, dx_ref) < 2e-6
    with pytest.raises(ValueError):
        ops.gemm_planes_to_operand(xo, ops.split_f16(in_w, _lib.SCALE_ROWS))     # the bound needs one scale for B


def test_operands_with_more_than_65535_rows(cuda):
    """lm_head of a 128k-token vocabulary: the row index must not sit on a 16-bit grid dimension."""
    from grasp_b200 import _lib, ops
    g = torch.Generator().manual_seed(9)
    x = torch.randn(96, 72, generator=g).to(cuda)
    w = torch.randn(70001, 72, generator=g).to(cuda)
    ref = x.double() @ w.double().t()
    xo = ops.split_f16(x)
    assert rel(ops.gemm_planes(xo, ops.split_f16(w, _lib.SCALE_TENSOR)), ref) < 2e-6
    assert rel(ops.gemm_planes(xo, ops.split_f16(w, _lib.SCALE_ROWS)), ref) < 2e-6
    assert rel(ops.gemm(x, w, tb=True), ref) < 2e-6                       # split inside grasp_gemm_f32
    dy = torch.randn(96, 70001, generator=g).to(cuda)
    assert rel(ops.gemm_planes(ops.split_f16(dy), ops.split_f16(w, _lib.SCALE_TENSOR), b_kn=True),
               dy.double() @ w.double()) < 2e-6
    assert rel(ops.gemm(dy, w), dy.double() @ w.double()) < 2e-6


def test_split_scales_badly_scaled_rows_and_zero_rows(cuda):
    from grasp_b200 import _lib, ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(96, 520, generator=g)
    x[3] *= 1e-20
    x[7] *= 1e18
    x[11] = 0
    x = x.to(cuda)
    w = torch.randn(40, 520, generator=g).to(cuda)
    y = ops.gemm_planes(ops.split_f16(x), ops.split_f16(w, _lib.SCALE_TENSOR))
    ref = x.double() @ w.double().t()
    for r in (0, 3, 7):                          # every row keeps its own relative accuracy
        assert rel(y[r], ref[r]) < 2e-6
    assert y[11].abs().max().item() == 0.0
    assert torch.isfinite(y).all()


# ------------------------------------------------------------------------------- row kernels
@pytest.mark.parametrize("rows,d", [(37, 4096), (5, 1000), (3, 11008), (2, 66)])
def test_rmsnorm_forward_backward(cuda, rows, d):
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(d)
    x = torch.randn(rows, d, generator=g).to(cuda)
    w = (1 + 0.1 * torch.randn(d, generator=g)).to(cuda)
    dy = torch.randn(rows, d, generator=g).to(cuda)
    add = torch.randn(rows, d, generator=g).to(cuda)
    eps = 1e-5
    xd = x.double().requires_grad_(True)
    yd = w.double() * (xd * torch.rsqrt(xd.pow(2).mean(-1, keepdim=True) + eps))
    (dxd,) = torch.autograd.grad(yd, xd, dy.double())
    y, rstd = ops.rmsnorm_fwd(x, w, eps)
    assert rel(y, yd.detach()) < 1e-6
    assert rel(rstd, torch.rsqrt(x.double().pow(2).mean(-1) + eps)) < 1e-6
    assert rel(ops.rmsnorm_bwd(dy, x, w, rstd), dxd) < 2e-6
    assert rel(ops.rmsnorm_bwd(dy, x, w, rstd, add=add), dxd + add.double()) < 2e-6


@pytest.mark.parametrize("B,S,H,D,per_batch", [(2, 33, 4, 128, False), (3, 7, 2, 16, True), (1, 511, 32, 128, False)])
def test_rope_matches_transformers_and_its_transpose(cuda, B, S, H, D, per_batch):
    from grasp_b200 import ops
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    g = torch.Generator().manual_seed(S)
    q = torch.randn(B, S, H, D, generator=g).to(cuda)
    nb = B if per_batch else 1
    ang = torch.rand(nb, S, D // 2, generator=g) * 6.28
    # independent halves (not the duplicated layout) so that the c1/c2, s1/s2 indexing is pinned down
    cos = torch.cat((ang.cos(), (ang * 1.3).cos()), -1).to(cuda)
    sin = torch.cat((ang.sin(), (ang * 0.7).sin()), -1).to(cuda)
    ql = q.clone().requires_grad_(True)
    ref, _ = apply_rotary_pos_emb(ql.transpose(1, 2), ql.transpose(1, 2), cos, sin)     # [B, H, S, D]
    ref = ref.transpose(1, 2)
    got = ops.rope_(q.clone().view(B * S, H * D), S, H, D, cos, sin).view(B, S, H, D)
    assert rel(got, ref.detach()) < 1e-6
    dy = torch.randn(B, S, H, D, generator=g).to(cuda)
    (dq,) = torch.autograd.grad(ref, ql, dy)
    got = ops.rope_(dy.clone().view(B * S, H * D), S, H, D, cos, sin, inverse=True).view(B, S, H, D)
    assert rel(got, dq) < 1e-6


@pytest.mark.parametrize("shape", [(13, 11008), (3, 177), (1, 5)])
def test_swiglu_forward_backward(cuda, shape):
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(shape[1])
    a = (torch.randn(*shape, generator=g) * 3).to(cuda)
    b = torch.randn(*shape, generator=g).to(cuda)
    dh = torch.randn(*shape, generator=g).to(cuda)
    ad, bd = a.double().requires_grad_(True), b.double().requires_grad_(True)
    hd = torch.nn.functional.silu(ad) * bd
    dga, dgb = torch.autograd.grad(hd, (ad, bd), dh.double())
    assert rel(ops.swiglu_fwd(a, b), hd.detach()) < 1e-6
    dg, du = ops.swiglu_bwd(dh, a, b)
    assert rel(dg, dga) < 1e-6 and rel(du, dgb) < 1e-6
    a2, b2 = a.clone(), b.clone()
    dg2, du2 = ops.swiglu_bwd(dh, a2, b2, inplace=True)
    assert dg2.data_ptr() == a2.data_ptr() and torch.equal(dg2, dg) and torch.equal(du2, du)


@pytest.mark.parametrize("rows,V", [(64, 32000), (9, 1001), (4, 128256)])
def test_cross_entropy_loss_and_gradient(cuda, rows, V):
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(V)
    logits = (torch.randn(rows, V, generator=g) * 4).to(cuda)
    labels = torch.randint(0, V, (rows,), generator=g).to(cuda)
    labels[1] = -100
    coef = torch.rand(rows, generator=g).to(cuda)
    ld = logits.double().requires_grad_(True)
    per = torch.nn.functional.cross_entropy(ld, labels, reduction="none", ignore_index=-100) * coef.double()
    (dl,) = torch.autograd.grad(per.sum(), ld)
    work = logits.clone()
    loss = ops.ce_loss_bwd_(work, labels, coef)
    assert rel(loss, per.detach()) < 1e-6
    assert rel(work, dl) < 2e-6
    assert loss[1].item() == 0.0 and work[1].abs().max().item() == 0.0


# ------------------------------------------------------------- the explicit forward / backward
def _prepare(cuda, name="small", seed=0, layer_id=3, block="attention"):
    from grasp_b200 import synth
    from modeling_grasp import GRASPModel
    model = synth.random_llama(name, seed=seed, device=cuda)
    gm = GRASPModel(model)
    tokens = synth.random_tokens(6, 40, model.config.vocab_size, seed=1)
    dl = synth.calibration_dataloader(0, 0, 0, tokens=tokens)
    return gm, dl


@pytest.mark.parametrize("block", ["mlp", "attention"])
def test_fused_route_equals_transformers_autograd_route(cuda, block):
    """Same model, same calibration set: dL/dS through the explicit forward/backward on the row kernels and
    prepared-operand GEMMs vs through transformers' modules + autograd (both on grasp_gemm_f32 arithmetic)."""
    from grasp_b200 import engine
    types = ["gate_proj", "up_proj", "down_proj"] if block == "mlp" else ["q_proj", "k_proj", "v_proj", "o_proj"]
    grads = {}
    states = {}
    for fused in (True, False):
        gm, dl = _prepare(cuda)
        gm.micro_batch = 4
        runner = gm._engine_runner()
        runner.use_fused = fused
        # a deeper layer already compressed to factor pairs, as in the real loop (deepest first)
        gm.compress_block(5, "mlp", ["down_proj", "up_proj", "gate_proj"], device=cuda)
        g5 = gm.get_svdlayer_gradients(dl, cuda)
        gm.compile_grasp_model(gm.dynamic_svd_selection(g5, compression_ratio=0.8), merge=False, device=cuda)
        gm.compress_block(3, block, types, device=cuda)
        assert (runner.fused(torch.zeros(1, device=cuda)) is not None) == fused
        grads[fused] = gm.get_svdlayer_gradients(dl, cuda)
        calib = gm._calibration_set(dl, cuda)
        with torch.no_grad(), engine.grasp_linear(True):
            states[fused] = runner.hidden_states(calib.input_ids[:3], runner.fused(torch.zeros(1, device=cuda)))
    assert set(grads[True]) == set(grads[False]) and len(grads[True]) == len(types)
    for name in grads[True]:
        assert rel(grads[True][name], grads[False][name]) < 2e-4, name
    for a, b in zip(states[True], states[False]):
        assert rel(a, b) < 2e-5


def test_weight_plane_cache_follows_module_replacement(cuda):
    from grasp_b200 import fused as fz
    gm, dl = _prepare(cuda, name="tiny")
    runner = gm._engine_runner()
    f = runner.fused(torch.zeros(1, device=cuda))
    assert f is not None
    calib = gm._calibration_set(dl, cuda)
    with torch.no_grad():
        runner.hidden_states(calib.input_ids[:2], f)
    n0 = len(f.be._w)
    assert n0 == 7 * runner.n_layers                      # the head is only used by the loss
    with torch.no_grad():
        runner.hidden_states(calib.input_ids[:2], f)
    assert len(f.be._w) == n0                             # second pass: every weight served from the cache
    runner.invalidate_above(1)
    assert len(f.be._w) == n0 - 7
    w = runner.layers[0].mlp.up_proj.weight
    op = f.be.wprep(w, 0)
    with torch.no_grad():
        w.mul_(2.0)                                       # an in-place change bumps the version: planes are redone
    assert f.be.wprep(w, 0) is not op
    assert isinstance(f.be, fz.CudaBackend)


def test_perplexity_evaluator_on_the_fused_route(cuda):
    """evaluate_grasp.py:99-127 through the fused forward + grasp_ce_loss_bwd, against the CPU oracle; also on
    a model whose layers are partly compressed to factor pairs."""
    import copy
    from grasp_b200 import evaluate, synth
    from oracle import restate
    from modeling_grasp import GRASPModel
    cpu_model = synth.random_llama("small", seed=2)
    tok = synth.random_tokens(6, 33, cpu_model.config.vocab_size, seed=5)
    ref = restate.perplexity(cpu_model, tok)
    model = copy.deepcopy(cpu_model)
    got = evaluate.evaluate_perplexity(model, tok, None, cuda, micro_batch=4)
    assert abs(got - ref) / ref < 1e-4
    gm = GRASPModel(model)
    dl = synth.calibration_dataloader(0, 0, 0, tokens=tok)
    gm.compress_block(4, "mlp", ["down_proj", "up_proj", "gate_proj"], device=cuda)
    gm.compile_grasp_model(gm.dynamic_svd_selection(gm.get_svdlayer_gradients(dl, cuda), compression_ratio=0.5),
                           merge=False, device=cuda)
    got_c = evaluate.evaluate_perplexity(gm, tok, None, cuda)
    ref_c = restate.perplexity(copy.deepcopy(gm.model).to("cpu"), tok)
    assert abs(got_c - ref_c) / ref_c < 1e-4


@pytest.mark.parametrize("rows,d,f", [(37, 4096, 11008), (5, 1000, 2744), (3, 66, 177)])
def test_row_kernels_emit_gemm_operands(cuda, rows, d, f):
    """RMSNorm / SwiGLU forward / SwiGLU backward hand their results over as prepared GEMM operands (planes +
    row scales) with or without the fp32 tensor: the product of that operand with a weight must equal the
    product of the separately split fp32 result, bit for bit, and the fp32 outputs must not change."""
    from grasp_b200 import _lib, ops
    g = torch.Generator().manual_seed(rows + d)
    x = torch.randn(rows, d, generator=g).to(cuda)
    w = (1 + 0.1 * torch.randn(d, generator=g)).to(cuda)
    W = (torch.randn(24, d, generator=g) * 0.05).to(cuda)
    Wo = ops.split_f16(W, _lib.SCALE_TENSOR)
    y, rstd = ops.rmsnorm_fwd(x, w, 1e-5)
    y2, rstd2, op = ops.rmsnorm_fwd(x, w, 1e-5, want_y=True, want_operand=True)
    none, rstd3, op3 = ops.rmsnorm_fwd(x, w, 1e-5, want_y=False, want_operand=True)
    assert none is None and torch.equal(y, y2) and torch.equal(rstd, rstd2) and torch.equal(rstd, rstd3)
    ref = ops.gemm_planes(ops.split_f16(y), Wo)
    assert torch.equal(ops.gemm_planes(op, Wo), ref) and torch.equal(ops.gemm_planes(op3, Wo), ref)

    a = (torch.randn(rows, f, generator=g) * 3).to(cuda)
    b = torch.randn(rows, f, generator=g).to(cuda)
    dh = torch.randn(rows, f, generator=g).to(cuda)
    Wf = (torch.randn(16, f, generator=g) * 0.05).to(cuda)
    Wfo = ops.split_f16(Wf, _lib.SCALE_TENSOR)
    h = ops.swiglu_fwd(a, b)
    h2, hop = ops.swiglu_fwd(a, b, want_h=True, want_operand=True)
    h3, hop3 = ops.swiglu_fwd(a, b, want_h=False, want_operand=True)
    assert h3 is None and torch.equal(h, h2)
    ref = ops.gemm_planes(ops.split_f16(h), Wfo)
    assert torch.equal(ops.gemm_planes(hop, Wfo), ref) and torch.equal(ops.gemm_planes(hop3, Wfo), ref)

    dg, du = ops.swiglu_bwd(dh, a, b)
    dg2, du2, gop, uop = ops.swiglu_bwd(dh, a, b, want_grads=True, want_operands=True)
    n1, n2, gop3, uop3 = ops.swiglu_bwd(dh, a, b, want_grads=False, want_operands=True)
    assert n1 is None and n2 is None and torch.equal(dg, dg2) and torch.equal(du, du2)
    for o, o3, t in ((gop, gop3, dg), (uop, uop3, du)):
        ref = ops.gemm_planes(ops.split_f16(t), Wfo)
        assert torch.equal(ops.gemm_planes(o, Wfo), ref) and torch.equal(ops.gemm_planes(o3, Wfo), ref)
    # in place: g and u are overwritten by their gradients while the operands are produced
    a2, b2 = a.clone(), b.clone()
    dg4, du4, gop4, _ = ops.swiglu_bwd(dh, a2, b2, inplace=True, want_operands=True)
    assert dg4.data_ptr() == a2.data_ptr() and torch.equal(dg4, dg) and torch.equal(du4, du)
    assert torch.equal(ops.gemm_planes(gop4, Wfo), ops.gemm_planes(gop, Wfo))
