"""GPU: the engine's memory plan.  With an artificially small activation store / process budget the job takes the
bounded paths (a subset of layer inputs resident, the rest recomputed per pass; smaller pass micro-batches) and
must retain the same singular triplets as the unbounded run (reference grasp.py:79-126 needs O(1) extra memory
per layer; any NUM_PRUNE_LAYERS has to work)."""
import pytest
import torch

from grasp_b200 import engine, synth

pytestmark = pytest.mark.gpu


def _run(cuda, store_mb=None, budget_extra_mb=None, layers=9):
    import grasp
    from modeling_grasp import GRASPModel
    model = synth.random_llama("small", seed=11, num_hidden_layers=12).to(cuda)
    tokens = synth.random_tokens(6, 24, model.config.vocab_size, seed=3)
    gm = GRASPModel(model)
    gm.micro_batch = 4
    runner = gm._engine_runner()
    if store_mb is not None:
        runner.store_cap_bytes = int(store_mb * 2**20)
    if budget_extra_mb is not None:
        runner.budget_bytes = torch.cuda.memory_allocated(cuda) + int(budget_extra_mb * 2**20)
    seen = {"max_entries": 0, "builds": 0, "pass_mb": []}
    build = runner.build_cache

    def spy(calib, layer_ids, keep_only=False):
        build(calib, layer_ids, keep_only=keep_only)
        seen["builds"] += 1
        seen["max_entries"] = max(seen["max_entries"], len(runner.cache))
    runner.build_cache = spy
    grads_fn = runner.sigma_gradients

    def spy_grads(calib, layers_, start):
        out = grads_fn(calib, layers_, start)
        seen["pass_mb"].append(runner.last_pass_micro_batch)
        return out
    runner.sigma_gradients = spy_grads
    picked = []
    select = gm.dynamic_svd_selection

    def spy_sel(*a, **kw):
        out = select(*a, **kw)
        picked.append({k: v.clone() for k, v in out.items()})
        return out
    gm.dynamic_svd_selection = spy_sel
    dl = synth.calibration_dataloader(0, 0, 0, tokens=tokens)
    grasp.compress(gm, dl, num_prune_layers=layers, compression_ratio=0.8, device=cuda)
    torch.cuda.synchronize()
    state = {k: v.detach().cpu() for k, v in gm.model.state_dict().items()}
    return gm, picked, state, seen


def test_bounded_activation_store_gives_the_same_compression(cuda):
    per_mb = 6 * 23 * 256 * 4 / 2**20                      # one store entry
    gm0, picked0, state0, seen0 = _run(cuda)
    assert seen0["max_entries"] >= 9                        # everything resident when memory is plentiful
    gm1, picked1, state1, seen1 = _run(cuda, store_mb=3.2 * per_mb)
    assert seen1["max_entries"] <= 3                        # the store never exceeded its budget
    assert gm1.redundant_layers == gm0.redundant_layers
    assert len(picked0) == len(picked1) == 18
    for a, b in zip(picked0, picked1):
        assert a.keys() == b.keys()
        for k in a:
            assert torch.equal(a[k], b[k]), k               # recomputed activations are bit-identical
    for k in state0:
        assert torch.equal(state0[k], state1[k]), k


def test_process_budget_shrinks_the_pass_micro_batch(cuda):
    gm0, picked0, state0, seen0 = _run(cuda, layers=3)
    assert set(seen0["pass_mb"]) == {4}
    gm1, picked1, state1, seen1 = _run(cuda, budget_extra_mb=12.0, layers=3)
    assert max(seen1["pass_mb"]) < 4                        # deep passes no longer fit four samples
    assert gm1.redundant_layers == gm0.redundant_layers
    for a, b in zip(picked0, picked1):
        for k in a:
            inter = len(set(a[k].tolist()) & set(b[k].tolist()))
            assert inter >= 0.97 * a[k].numel(), k          # G sums in a different order: ties may flip
