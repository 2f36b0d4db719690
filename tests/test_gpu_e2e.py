"""GPU: the whole hot path through the drop-in front (modeling_grasp / grasp.compress) against the
outputs of the real reference on the same seeded model and tokens (tests/golden/e2e_*.pt)."""
import copy

import pytest
import torch

from grasp_b200 import synth
from oracle import restate
from oracle.make_golden import state_checksum

pytestmark = pytest.mark.gpu


from parity_util import compare_blocks, compare_final_weights, rel


def run_ours(model, tokens, num_prune_layers, ratio, merge, device, use_engine=True, **options):
    """grasp.compress with per-block artefacts captured from the public GRASPModel methods."""
    import grasp
    from modeling_grasp import GRASPModel
    gm = GRASPModel(model)
    gm.use_engine = use_engine      # False: the generic whole-model forward/backward of the reference loop
    gm.micro_batch = 3
    gm.model.to(device)
    rec = {"blocks": []}
    orig_sel = gm.dynamic_svd_selection

    def spy(grads, **kw):
        names = list(grads.keys())
        S = {n: gm.model.get_submodule(n).S.data.clone() for n in names}
        idx = orig_sel(grads, **kw)
        rec["blocks"].append({"names": names, "S": S, "grads": {n: grads[n].clone() for n in names},
                              "indices": {n: torch.as_tensor(idx[n]).clone() for n in names}})
        return idx

    gm.dynamic_svd_selection = spy
    dl = synth.calibration_dataloader(0, 0, 0, tokens=tokens)
    grasp.compress(gm, dl, num_prune_layers=num_prune_layers, compression_ratio=ratio, merge=merge, device=device,
                   **options)
    rec["layer_importances"] = gm.layer_importances
    rec["layers_id"] = gm.redundant_layers
    return gm, rec


@pytest.mark.parametrize("fname,merge,use_engine", [("e2e_tiny.pt", False, True), ("e2e_tiny.pt", True, True),
                                                    ("e2e_small.pt", False, True), ("e2e_tiny.pt", False, False),
                                                    ("e2e_small.pt", False, False)])
def test_end_to_end_parity_with_reference(cuda, golden, parity_log, fname, merge, use_engine):
    fx = golden(fname)
    ref = fx["merge" if merge else "factored"]
    model = synth.random_llama(fx["model"], seed=fx["seed"])
    assert state_checksum(model) == fx["model_sha256"]
    dense = copy.deepcopy(model)
    gm, rec = run_ours(model, fx["tokens"], fx["num_prune_layers"], fx["ratio"], merge, "cuda", use_engine)
    tag = f"e2e {fname} merge={merge} engine={use_engine}"

    # stage 1: identical layer choice, BI within 1e-4 relative
    assert rec["layers_id"] == ref["layers_id"]
    assert rel(rec["layer_importances"], ref["layer_importances"]) < 1e-4

    # stages 2-3b: singular values, scores, retained index sets per block
    compare_blocks(rec, ref, parity_log, tag)

    # stage 3c: rebuilt weights of every compressed matrix
    if "final_state" in ref:
        ours_sd = {k: v.detach().cpu() for k, v in gm.model.state_dict().items()}
        compare_final_weights(rec, ref, ours_sd, ref["final_state"], dense.state_dict(), parity_log, tag)

    # downstream perplexity on the calibration tokens within 0.5 %
    gm.model.to("cpu")
    ppl = restate.perplexity(gm.model, fx["tokens"])
    assert abs(ppl - ref["ppl_compressed"]) / ref["ppl_compressed"] < 5e-3, (ppl, ref["ppl_compressed"])
    assert abs(restate.perplexity(dense, fx["tokens"]) - ref["ppl_dense"]) / ref["ppl_dense"] < 1e-4


@pytest.mark.parametrize("variant", ["gradient_metric", "threshold_merge"])
def test_option_variants_match_the_reference(cuda, golden, parity_log, variant):
    """The option branches of the path, against runs of the reference with the same options
    (tests/golden/e2e_tiny_variants.pt, oracle/make_golden.py:VARIANTS): the |gradient| score (modeling_grasp.py:393-394)
    and ranks chosen by the cumulative-score threshold (:408-410, tools/utils_func.py:45-57) with the merged rebuild."""
    fx = golden("e2e_tiny_variants.pt")
    ref = fx["variants"][variant]
    opt = ref["options"]
    model = synth.random_llama(fx["model"], seed=fx["seed"])
    assert state_checksum(model) == fx["model_sha256"]
    dense = copy.deepcopy(model)
    gm, rec = run_ours(model, fx["tokens"], opt["num_prune_layers"], opt["ratio"], opt["merge"], "cuda",
                       metric=opt["metric"], threshold_ratio=opt["threshold_ratio"])
    tag = f"e2e variant {variant}"
    assert rec["layers_id"] == ref["layers_id"]
    assert rel(rec["layer_importances"], ref["layer_importances"]) < 1e-4
    compare_blocks(rec, ref, parity_log, tag, metric=opt["metric"], adaptive=opt["threshold_ratio"] is not None)
    ours_sd = {k: v.detach().cpu() for k, v in gm.model.state_dict().items()}
    compare_final_weights(rec, ref, ours_sd, ref["final_state"], dense.state_dict(), parity_log, tag)
    gm.model.to("cpu")
    ppl = restate.perplexity(gm.model, fx["tokens"])
    assert abs(ppl - ref["ppl_compressed"]) / ref["ppl_compressed"] < 5e-3, (ppl, ref["ppl_compressed"])


def test_grasp_layer_standalone_backward_matches_oracle(cuda):
    """A GRASPLayer used outside the engine still yields S.grad == autograd through U diag(S) Vh."""
    from modeling_grasp import GRASPLayer
    g = torch.Generator().manual_seed(4)
    W = torch.randn(96, 160, generator=g) * 0.02
    x = torch.randn(2, 7, 160, generator=g)
    U, S, Vh = restate.svd(W)
    ref = restate.OracleGRASPLayer(U, S, Vh)
    (ref(x) ** 2).sum().backward()
    layer = GRASPLayer(U.to(cuda), S.to(cuda), Vh.to(cuda), None, None, weight=W.to(cuda))
    xg = x.to(cuda).requires_grad_(True)
    y = layer(xg)
    (y ** 2).sum().backward()
    assert rel(y, ref(x)) < 1e-4
    assert rel(layer.S.grad, ref.S.grad) < 1e-3
    assert xg.grad is not None


def test_sample_sharded_gradients_select_the_same_triplets_up_to_ties(cuda, parity_log, monkeypatch):
    """Multi-GPU plan: every rank contracts the G of its own sample shard and the sigma-gradients are
    summed.  Emulated here on one GPU (two CalibrationSets built as rank 0 / rank 1 of a world of 2, their
    gradients added): the sum equals the single-rank gradients to fp32 summation order, and the retained index sets
    differ only where scores tie."""
    from grasp_b200 import dist, engine, ops
    from modeling_grasp import GRASPModel
    from parity_util import TIE_TOL
    model = synth.random_llama("small", seed=9).to(cuda)
    tokens = synth.random_tokens(10, 32, model.config.vocab_size, seed=2)
    gm = GRASPModel(model)
    gm.micro_batch = 3
    runner = gm._engine_runner()
    gm.compress_block(4, "mlp", ["down_proj", "up_proj", "gate_proj"], device=cuda)
    names = gm.check_exists_grasp_layer()
    layers = {n: gm.model.get_submodule(n) for n in names}
    dl = synth.calibration_dataloader(0, 0, 0, tokens=tokens)

    def grads_as(rank, world):
        monkeypatch.setattr(dist, "rank_world", lambda: (rank, world))
        calib = engine.CalibrationSet(dl, cuda)
        monkeypatch.undo()
        return calib, runner.sigma_gradients(calib, layers, 4)

    c_all, g_all = grads_as(0, 1)
    c0, g0 = grads_as(0, 2)
    c1, g1 = grads_as(1, 2)
    assert len(c_all) == 10 and len(c0) == 5 and len(c1) == 5
    worst_g, worst_tie, swapped = 0.0, 0.0, 0
    for n in names:
        summed = g0[n] + g1[n]
        worst_g = max(worst_g, rel(summed, g_all[n]))
        S = layers[n].S.data
        k = gm.compute_preserve_rank(layers[n], 0.8)
        sc_a, sc_b = ops.score_from_grad(g_all[n], S, "taylor"), ops.score_from_grad(summed, S, "taylor")
        ia, ib = ops.topk(sc_a, k).tolist(), ops.topk(sc_b, k).tolist()
        kth = sc_a[ia[-1]].item()
        for i in set(ia) ^ set(ib):
            worst_tie = max(worst_tie, abs(sc_a[i].item() - kth) / kth)
            swapped += 1
    parity_log(f"2-rank emulation: summed shard gradients vs single rank {worst_g:.1e}; {swapped // 2} swapped indices, "
               f"worst tie margin {100 * worst_tie:.4f} %")
    assert worst_g < 1e-5
    assert worst_tie <= TIE_TOL
