"""CPU: host-side logic of the calibration engine that does not touch a kernel -- the layer-wise runner
reproduces HF's own forward (hidden states, loss with the reference's double label shift)."""
import torch

from grasp_b200 import engine, synth


def test_runner_matches_hf_forward_and_loss():
    model = synth.random_llama("tiny", seed=3)
    assert engine.LlamaRunner.supports(model)
    runner = engine.LlamaRunner(model, micro_batch=2, use_grasp_gemm=False)
    tokens = synth.random_tokens(3, 20, 256, seed=1)
    ids, labels = tokens[:, :-1], tokens[:, 1:]
    with torch.no_grad():
        out = model(input_ids=ids, labels=labels, output_hidden_states=True, use_cache=False, return_dict=True)
        states = runner.hidden_states(ids)
    assert len(states) == len(out.hidden_states) == model.config.num_hidden_layers + 1
    for a, b in zip(states, out.hidden_states):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-5)
    # batch loss of HF == mean over samples of the per-sample means (equal lengths)
    with torch.no_grad():
        hidden = runner.run_layers(runner.embed(ids), 0, runner.n_layers)
        w = torch.full((3,), 1.0 / 3)
        loss = runner.loss_sum(hidden, labels, w)
    assert abs(loss.item() - out.loss.item()) < 1e-5
    # per-sample losses (reference batch size 1) add up
    with torch.no_grad():
        singles = sum(model(input_ids=ids[i:i + 1], labels=labels[i:i + 1], use_cache=False)[0].item() for i in range(3))
        loss1 = runner.loss_sum(hidden, labels, torch.ones(3))
    assert abs(loss1.item() - singles) < 1e-4


def test_prefix_cache_equals_full_forward_and_invalidation():
    model = synth.random_llama("tiny", seed=4)
    runner = engine.LlamaRunner(model, micro_batch=2, use_grasp_gemm=False)
    dl = synth.calibration_dataloader(5, 12, 256, batch_size=1, seed=2)
    calib = engine.CalibrationSet(dl, "cpu")
    assert calib.supported and len(calib) == 5 and calib.n_batches == 5
    assert torch.all(calib.weights == 1.0)
    runner.build_cache(calib, [1, 3])
    with torch.no_grad():
        states = runner.hidden_states(calib.input_ids)
    assert torch.allclose(runner.cache[1], states[1], atol=1e-5)
    assert torch.allclose(runner.cache[3], states[3], atol=1e-5)
    runner.invalidate_above(2)
    assert 1 in runner.cache and 3 not in runner.cache
    dl2 = synth.calibration_dataloader(4, 12, 256, batch_size=2, seed=2)
    calib2 = engine.CalibrationSet(dl2, "cpu")
    assert torch.all(calib2.weights == 0.5)


def test_layer_index_parsing():
    assert engine.layer_index("model.layers.17.mlp.down_proj") == 17
    assert engine.layer_index("layers.3.self_attn.q_proj") == 3
    assert engine.layer_index("lm_head") is None


def test_cache_from_checkpoint_matches_cache_from_scratch():
    model = synth.random_llama("tiny", seed=5)
    runner = engine.LlamaRunner(model, micro_batch=2, use_grasp_gemm=False)
    calib = engine.CalibrationSet(synth.calibration_dataloader(5, 12, 256, seed=3), "cpu")
    with torch.no_grad():
        states = runner.hidden_states(calib.input_ids)
    runner.cache, runner.cache_key = {1: states[1].clone(), 2: states[2].clone()}, id(calib)   # as left by the scoring pass
    runner.build_cache(calib, [2, 3], keep_only=True)
    assert torch.allclose(runner.cache[2], states[2], atol=1e-6)
    assert torch.allclose(runner.cache[3], states[3], atol=1e-5)
    assert set(runner.cache) == {2, 3}                                   # the unselected checkpoint was released
    runner.cache = {}
    runner.build_cache(calib, [1])                                       # no usable checkpoint: from the embeddings
    assert torch.allclose(runner.cache[1], states[1], atol=1e-5)


def test_perplexity_evaluator_matches_reference_formula():
    """grasp_b200.evaluate.evaluate_perplexity == reference evaluate_grasp.py:99-127 (oracle restatement)."""
    from grasp_b200 import evaluate
    from oracle import restate
    model = synth.random_llama("tiny", seed=3)
    tok = synth.random_tokens(5, 20, 256, seed=2)
    ref = restate.perplexity(model, tok)
    assert abs(evaluate.evaluate_perplexity(model, tok, None, "cpu", micro_batch=2) - ref) / ref < 1e-5
    assert abs(evaluate.evaluate_perplexity(model, tok, 2, "cpu") - restate.perplexity(model, tok[:2])) / ref < 1e-5


def test_bounded_store_keeps_a_subset_and_recomputes_the_rest():
    """With a store budget of three entries, build_cache keeps the lowest + evenly spaced layers and the
    missing ones are recomputed from the nearest resident entry below when their pass asks for them."""
    model = synth.random_llama("small", seed=6)                          # 6 layers
    runner = engine.LlamaRunner(model, micro_batch=2, use_grasp_gemm=False)
    calib = engine.CalibrationSet(synth.calibration_dataloader(4, 10, 1024, seed=4), "cpu")
    with torch.no_grad():
        states = runner.hidden_states(calib.input_ids)
    per = states[0].numel() * 4
    runner.store_key, runner.store_budget, runner.store_per = id(calib), 3 * per, per
    runner.build_cache(calib, [1, 2, 3, 4, 5], keep_only=True)
    assert set(runner.cache) == {1, 5}                                   # 3 slots: two kept, one free for a pass
    runner.invalidate_above(4)                                           # layer 5 done, layer 4 is next
    runner.build_cache(calib, [4])
    assert set(runner.cache) == {1, 4}
    assert torch.allclose(runner.cache[4], states[4], atol=1e-5)
    runner.invalidate_above(3)
    runner.build_cache(calib, [3])
    assert torch.allclose(runner.cache[3], states[3], atol=1e-5)


def test_runner_refuses_other_model_families_and_masked_labels_average_like_hf():
    from transformers import MistralConfig, MistralForCausalLM
    cfg = MistralConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                        num_key_value_heads=2, vocab_size=64, sliding_window=8)
    assert not engine.LlamaRunner.supports(MistralForCausalLM(cfg))      # sliding window: generic path
    model = synth.random_llama("tiny", seed=3)
    runner = engine.LlamaRunner(model, micro_batch=2, use_grasp_gemm=False)
    tokens = synth.random_tokens(2, 16, 256, seed=5)
    ids, labels = tokens[:, :-1], tokens[:, 1:].clone()
    labels[0, 5:] = -100                                                 # padded tail, ignored by HF's loss
    with torch.no_grad():
        hidden = runner.run_layers(runner.embed(ids), 0, runner.n_layers)
        ours = runner.loss_sum(hidden, labels, torch.ones(2))
        ref = sum(model(input_ids=ids[i:i + 1], labels=labels[i:i + 1], use_cache=False)[0].item() for i in range(2))
    assert abs(ours.item() - ref) < 1e-4


def test_svd_host_logic_groups_by_working_shape_and_climbs_the_retry_ladder(monkeypatch):
    """engine._batched_svd_local (the host side of reference modeling_grasp.py:231 for many matrices): one library call
    per group of <= 8 matrices with the same WORKING shape, same actual shapes side by side, results in input order,
    and the recovery of a matrix the library flags (info[:, 1] == 0): as it is, then on the fp32 path, then refused.
    The library call is replaced by a stand-in here (LAPACK on CPU) -- this tests the host logic, not a kernel."""
    import pytest
    from grasp_b200 import ops
    calls = []
    flagged = {"mode": None}

    def fake_svd_batched(mats, prec=None, max_sweeps=0, return_info=False, precondition=True):
        calls.append({"shapes": [tuple(m.shape) for m in mats], "prec": prec, "precondition": precondition,
                      "max_sweeps": max_sweeps})
        out = [tuple(torch.linalg.svd(m.double(), full_matrices=False)) for m in mats]
        out = [tuple(t.float() for t in usv) for usv in out]
        info = torch.zeros(len(mats), 4, dtype=torch.int32)
        info[:, 0] = 15
        for j, m in enumerate(mats):
            bad = flagged["mode"] is not None and m.shape == (16, 8) and m[0, 0].item() == 7.0
            ok = not bad or (flagged["mode"] == "precond" and not precondition) or \
                (flagged["mode"] == "fp32" and prec == ops.PREC_SIMT)
            info[j, 1] = int(ok)
        return out, info

    monkeypatch.setattr(ops, "svd_batched", fake_svd_batched)
    g = torch.Generator().manual_seed(0)
    # working shape (dist.svd_working_shape) = (short side, long side): tall and wide matrices share a group
    shapes = [(16, 8)] * 5 + [(8, 8)] * 9 + [(8, 16)] * 2 + [(16, 8)] * 4
    mats = [torch.randn(s, generator=g) for s in shapes]
    mats[2][0, 0] = 7.0                                   # the matrix the stand-in will flag
    out = engine._batched_svd_local(mats)
    assert [c["shapes"] for c in calls] == [[(8, 16)] * 2 + [(16, 8)] * 6, [(16, 8)] * 3, [(8, 8)] * 8, [(8, 8)]]
    for w, (U, S, Vh) in zip(mats, out):                  # input order survives the grouping
        assert U.shape == (w.shape[0], min(w.shape)) and torch.allclose((U * S) @ Vh, w, atol=1e-5)

    for mode, expect in (("precond", [False]), ("fp32", [False, True])):
        calls.clear(); flagged["mode"] = mode
        out = engine._batched_svd_local(mats)
        retries = [c for c in calls if c["shapes"] == [(16, 8)] and len(c["shapes"]) == 1 and
                   (not c["precondition"] or c["prec"] == ops.PREC_SIMT)]
        assert [c["precondition"] for c in retries] == expect
        if mode == "fp32":
            assert retries[-1]["prec"] == ops.PREC_SIMT and retries[-1]["max_sweeps"] == 48
        assert torch.allclose((out[2][0] * out[2][1]) @ out[2][2], mats[2], atol=1e-5)
    flagged["mode"] = "never"
    with pytest.raises(engine.SvdNotConverged):
        engine._batched_svd_local(mats)
