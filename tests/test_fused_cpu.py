"""CPU: the explicit forward/backward of grasp_b200.fused (orchestration only, torch arithmetic from
tests/torch_backend.py) against transformers' modules + autograd, i.e. against what the reference runs at
modeling_grasp.py:347-354.  The CUDA kernels behind the same interface are checked in test_gpu_fused.py."""
import pytest
import torch

from grasp_b200 import engine, synth
from grasp_b200.fused import FusedLlama, layer_supported, linear_kind
from modeling_grasp import GRASPLayer, SVDLinear
from torch_backend import TorchBackend


def _get(model, name):
    return model.get_submodule(name)


def _set(model, name, mod):
    *parents, leaf = name.split(".")
    owner = model
    for p in parents:
        owner = getattr(owner, p)
    setattr(owner, leaf, mod)


def _to_grasp(model, name):
    lin = _get(model, name)
    U, S, Vh = torch.linalg.svd(lin.weight.data, full_matrices=False)
    layer = GRASPLayer(U, S, Vh, lin.bias, None, weight=lin.weight.data)
    _set(model, name, layer)
    return layer


def _to_svdlinear(model, name, k):
    lin = _get(model, name)
    U, S, Vh = torch.linalg.svd(lin.weight.data, full_matrices=False)
    root = S[:k].sqrt()
    new = SVDLinear.from_packed((Vh[:k] * root[:, None]).contiguous(), (U[:, :k] * root).contiguous(), lin.bias)
    new.requires_grad_(False)
    _set(model, name, new)


def _model(seed=0, **kw):
    model = synth.random_llama("tiny", seed=seed, **kw)
    for p in model.parameters():
        p.requires_grad = False
    return model


def test_fused_forward_matches_hf_hidden_states():
    model = _model(1)
    runner = engine.LlamaRunner(model, micro_batch=2, use_grasp_gemm=False)
    f = FusedLlama(runner, TorchBackend())
    assert f.supported() and all(layer_supported(l) for l in runner.layers)
    ids = synth.random_tokens(3, 17, 256, seed=2)
    with torch.no_grad():
        ref = model(input_ids=ids, output_hidden_states=True, use_cache=False, return_dict=True).hidden_states
        got = runner.hidden_states(ids, f)
        mid = f.run_layers(ref[1], 1, 3)
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert torch.allclose(a, b, atol=2e-5, rtol=1e-5)
    assert torch.allclose(mid, ref[3], atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("block,layer_id,upper_svd", [("mlp", 1, False), ("attention", 2, True), ("attention", 0, True),
                                                      ("mlp", 3, False)])
def test_fused_pass_harvests_the_same_G_as_autograd(block, layer_id, upper_svd):
    torch.manual_seed(0)
    model = _model(5)
    n_layers = model.config.num_hidden_layers
    if upper_svd and layer_id + 1 < n_layers:
        # a deeper layer already compiled to factor pairs (compile_grasp_model, merge=False)
        _to_svdlinear(model, f"model.layers.{layer_id + 1}.mlp.down_proj", 20)
        _to_svdlinear(model, f"model.layers.{layer_id + 1}.self_attn.k_proj", 9)
    types = ("gate_proj", "up_proj", "down_proj") if block == "mlp" else ("q_proj", "k_proj", "v_proj", "o_proj")
    owner = "mlp" if block == "mlp" else "self_attn"
    layers = {f"model.layers.{layer_id}.{owner}.{t}": None for t in types}
    for name in layers:
        layers[name] = _to_grasp(model, name)
    assert all(linear_kind(m) == "grasp" for m in layers.values())

    runner = engine.LlamaRunner(model, micro_batch=2, use_grasp_gemm=False)
    tokens = synth.random_tokens(2, 14, 256, seed=3)
    ids, labels = tokens[:, :-1], tokens[:, 1:]
    weights = torch.tensor([1.0, 0.5])
    with torch.no_grad():
        src = runner.hidden_states(ids)[layer_id]

    # reference route: transformers modules + autograd, G harvested by engine.SigmaLinearFn
    with engine.deferred_sigma_grads(layers.values()):
        hidden = runner.run_layers(src, layer_id, n_layers)
        loss_ref = runner.loss_sum(hidden, labels, weights)
        loss_ref.backward()
        G_ref = {n: l._G.clone() for n, l in layers.items()}

    f = FusedLlama(runner, TorchBackend())
    assert f.supported()
    with engine.deferred_sigma_grads(layers.values()), torch.no_grad():
        loss = f.forward_backward(src, labels, weights, layer_id, layer_id)
        G = {n: l._G.clone() for n, l in layers.items()}
        # a second micro-batch accumulates
        f.forward_backward(src, labels, weights, layer_id, layer_id)
        G2 = {n: l._G.clone() for n, l in layers.items()}
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    for n in layers:
        scale = G_ref[n].abs().max().item()
        assert scale > 0
        assert (G[n] - G_ref[n]).abs().max().item() <= 2e-5 * scale + 1e-9, n
        assert (G2[n] - 2 * G_ref[n]).abs().max().item() <= 4e-5 * scale + 1e-9, n


def test_fused_pass_starting_below_the_grasp_block():
    # start_layer < lowest GRASP layer: the layers in between run without saving anything
    model = _model(7)
    name = "model.layers.2.mlp.up_proj"
    layer = _to_grasp(model, name)
    runner = engine.LlamaRunner(model, micro_batch=2, use_grasp_gemm=False)
    tokens = synth.random_tokens(2, 10, 256, seed=4)
    ids, labels = tokens[:, :-1], tokens[:, 1:]
    weights = torch.ones(2)
    with torch.no_grad():
        src = runner.hidden_states(ids)[1]
    with engine.deferred_sigma_grads([layer]):
        runner.loss_sum(runner.run_layers(src, 1, 4), labels, weights).backward()
        G_ref = layer._G.clone()
    f = FusedLlama(runner, TorchBackend())
    with engine.deferred_sigma_grads([layer]), torch.no_grad():
        f.forward_backward(src, labels, weights, 1, 2)
        G = layer._G.clone()
    assert (G - G_ref).abs().max().item() <= 2e-5 * G_ref.abs().max().item()


def test_unsupported_modules_fall_back():
    model = _model(2)
    runner = engine.LlamaRunner(model, micro_batch=2, use_grasp_gemm=False)
    f = FusedLlama(runner, TorchBackend())
    assert f.supported()
    model.model.layers[1].mlp.act_fn = torch.nn.GELU()
    assert not f.supported()
    # and the runner never picks the fused route for CPU tensors (the product backend is CUDA-only)
    assert runner.fused(torch.zeros(1, 4, 64)) is None


def test_fused_pass_with_biased_linears():
    """LlamaConfig(attention_bias=True, mlp_bias=True): dense linears above the block add their bias, a GRASPLayer
    ignores its own (reference modeling_grasp.py:77-79), a compiled factor pair carries it on OutLinear (:41-44)."""
    model = _model(11, attention_bias=True, mlp_bias=True)
    g = torch.Generator().manual_seed(1)
    for n, p in model.named_parameters():
        if n.endswith(".bias"):
            p.data = torch.randn(p.shape, generator=g) * 0.05      # HF initialises them to zero
    _to_svdlinear(model, "model.layers.3.self_attn.v_proj", 7)
    _to_svdlinear(model, "model.layers.3.mlp.gate_proj", 12)
    names = [f"model.layers.1.self_attn.{t}" for t in ("q_proj", "k_proj", "v_proj", "o_proj")]
    layers = {n: _to_grasp(model, n) for n in names}
    assert all(l.bias is not None for l in layers.values())
    runner = engine.LlamaRunner(model, micro_batch=2, use_grasp_gemm=False)
    tokens = synth.random_tokens(2, 12, 256, seed=6)
    ids, labels = tokens[:, :-1], tokens[:, 1:]
    weights = torch.ones(2)
    with torch.no_grad():
        src = runner.hidden_states(ids)[1]
        # the runner's layer-wise forward is HF's own forward, biases included
        ref = model(input_ids=ids, output_hidden_states=True, use_cache=False, return_dict=True).hidden_states
        assert torch.allclose(runner.hidden_states(ids)[-1], ref[-1], atol=2e-5, rtol=1e-5)
    with engine.deferred_sigma_grads(layers.values()):
        loss_ref = runner.loss_sum(runner.run_layers(src, 1, 4), labels, weights)
        loss_ref.backward()
        G_ref = {n: l._G.clone() for n, l in layers.items()}
    f = FusedLlama(runner, TorchBackend())
    assert f.supported()
    with engine.deferred_sigma_grads(layers.values()), torch.no_grad():
        loss = f.forward_backward(src, labels, weights, 1, 1)
        G = {n: l._G.clone() for n, l in layers.items()}
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    for n in layers:
        assert (G[n] - G_ref[n]).abs().max().item() <= 2e-5 * G_ref[n].abs().max().item() + 1e-9, n
