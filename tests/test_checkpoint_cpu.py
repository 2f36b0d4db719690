"""CPU: the state-dict checkpoint (reference grasp.py:129-136 / evaluate.py:42) round-trips a compressed
model -- SVDLinear pairs and merged dense layers -- with identical logits."""
import torch

from grasp_b200 import checkpoint, synth
from modeling_grasp import GRASPModel, SVDLinear
from oracle import restate


def _compress_on_cpu(model):
    """A compressed tiny model without the GPU: the oracle's factors packed into the product's modules."""
    gm = GRASPModel(model)
    for name, merge in (("model.layers.3.mlp.down_proj", False), ("model.layers.3.self_attn.q_proj", False),
                        ("model.layers.2.mlp.up_proj", True)):
        lin = model.get_submodule(name)
        U, S, Vh = restate.svd(lin.weight.data)
        k = restate.preserve_rank(lin.in_features, lin.out_features, 0.5)
        idx = torch.arange(k)
        if merge:
            lin.weight.data = restate.merged_weight(U, S, Vh, idx)
        else:
            in_w, out_w = restate.packed_factors(U, S, Vh, idx)
            gm._set_module(model, name, SVDLinear.from_packed(in_w.contiguous(), out_w.contiguous(), None))
    gm.redundant_layers = [3, 2]
    return gm


def test_checkpoint_round_trip(tmp_path):
    model = synth.random_llama("tiny", seed=7)
    gm = _compress_on_cpu(model)
    tokens = synth.random_tokens(2, 16, 256, seed=1)
    with torch.no_grad():
        ref = gm.model(input_ids=tokens).logits
    path = str(tmp_path / "tiny.pth")
    checkpoint.save(gm, path)
    loaded = checkpoint.load(path, device="cpu")
    assert isinstance(loaded, GRASPModel) and loaded.redundant_layers == [3, 2]
    assert isinstance(loaded.model.get_submodule("model.layers.3.mlp.down_proj"), SVDLinear)
    assert not any(b.is_meta for b in loaded.model.buffers()) and not any(p.is_meta for p in loaded.model.parameters())
    with torch.no_grad():
        out = loaded.model(input_ids=tokens).logits
    assert torch.equal(out, ref)
    # the fallback of grasp._save: whole-module pickling fails -> this format, loadable by evaluate.py-style callers
    assert abs(restate.perplexity(loaded.model, tokens) - restate.perplexity(gm.model, tokens)) < 1e-6


def test_checkpoint_refuses_a_mismatching_file(tmp_path):
    import pytest
    model = synth.random_llama("tiny", seed=7)
    gm = GRASPModel(model)
    path = str(tmp_path / "bad.pth")
    checkpoint.save(gm, path)
    blob = torch.load(path, weights_only=False)
    del blob["state_dict"]["model.layers.0.mlp.up_proj.weight"]
    torch.save(blob, path)
    with pytest.raises(RuntimeError):
        checkpoint.load(path)
