"""CPU: the oracle restatement (oracle/restate.py) against the committed outputs of the real
reference (tests/golden/, written by oracle/make_golden.py)."""
import copy

import pytest
import torch

from grasp_b200 import synth
from oracle import restate
from oracle.make_golden import state_checksum


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def test_block_influence_matches_reference(golden):
    fx = golden("bi_small.pt")
    hs = fx["hiddens"]
    for i in range(len(hs) - 1):
        got = restate.block_influence(hs[i], hs[i + 1])
        assert rel(got, fx["per_pair"][i]) < 1e-5
        ang = restate.block_influence(hs[i][:, -1:], hs[i + 1][:, -1:], angular=True)
        assert rel(ang, fx["angular_last_token"][i]) < 1e-5
    # edge cases planted by the generator: zero row -> 0.5, identical -> 0, opposite -> 2
    assert abs(fx["per_pair"][1][3].item() - 0.5) < 1e-6
    assert abs(fx["per_pair"][2][17 + 5].item()) < 1e-6
    assert abs(fx["per_pair"][3][0].item() - 2.0) < 1e-5
    imp = [0.0] * 5
    restate.compute_bi_hiddens(hs, imp)
    assert rel(imp, fx["means"]) < 1e-6


def test_svd_cases_are_consistent(golden):
    for case in golden("svd_small.pt"):
        A, U, S, Vh = case["A"], case["U"], case["S"], case["Vh"]
        U2, S2, Vh2 = restate.svd(A)
        assert rel(S2, S) < 1e-5
        assert (torch.linalg.norm((U * S) @ Vh - A) / torch.linalg.norm(A)).item() < 1e-5
        assert torch.all(S[:-1] >= S[1:])


@pytest.mark.parametrize("metric", ["taylor", "gradient"])
def test_selection_and_compile_match_reference(golden, metric):
    for case in golden("select_small.pt"):
        U, S, Vh, G = case["U"], case["S"], case["Vh"], case["G"]
        g = restate.sigma_grad_from_G(U, G, Vh)
        assert rel(g, case["grad"]) < 2e-5
        for ratio in (0.9, 0.5):
            ref = case[f"{metric}_{ratio}"]
            o, i = U.shape[0], Vh.shape[1]
            assert restate.preserve_rank(i, o, ratio) == ref["k"]
            sc = restate.importance(case["grad"], S, metric)
            assert torch.equal(sc, ref["score"])
            idx = torch.topk(sc, k=ref["k"]).indices
            assert torch.equal(idx, ref["idx"])
            assert rel(restate.merged_weight(U, S, Vh, idx), ref["merged"]) < 1e-6
            iw, ow = restate.packed_factors(U, S, Vh, idx)
            assert torch.equal(iw, ref["in_w"]) and torch.equal(ow, ref["out_w"])
        thr = case["threshold_0.6"]
        assert restate.adaptive_rank_selection(thr["score"], 0.6) == thr["idx"].tolist()


def test_rank_table():
    # SURVEY.md appendix C, ratio 0.9
    assert restate.preserve_rank(4096, 4096, 0.9) == 204
    assert restate.preserve_rank(4096, 11008, 0.9) == 298
    assert restate.preserve_rank(2048, 2048, 0.9) == 102
    assert restate.preserve_rank(2048, 256, 0.9) == 22
    assert restate.preserve_rank(2048, 5632, 0.9) == 150
    assert restate.preserve_rank(4096, 14336, 0.9) == 318
    assert restate.preserve_rank(8192, 28672, 0.9) == 637


@pytest.mark.parametrize("fname,merge", [("e2e_tiny.pt", False), ("e2e_tiny.pt", True), ("e2e_small.pt", False)])
def test_end_to_end_restatement_matches_reference(golden, fname, merge):
    fx = golden(fname)
    model = synth.random_llama(fx["model"], seed=fx["seed"])
    assert state_checksum(model) == fx["model_sha256"], "random init is not reproducible on this box"
    ref = fx["merge" if merge else "factored"]
    rec = restate.run_grasp(copy.deepcopy(model), fx["tokens"], num_prune_layers=fx["num_prune_layers"],
                            compression_ratio=fx["ratio"], merge=merge)
    assert rec["layers_id"] == ref["layers_id"]
    assert rel(rec["layer_importances"], ref["layer_importances"]) < 1e-5
    for b, br in zip(rec["blocks"], ref["blocks"]):
        assert b["names"] == br["names"]
        for n in b["names"]:
            assert rel(b["S"][n], br["S"][n]) < 1e-6
            assert rel(b["grads"][n], br["grads"][n]) < 1e-4
            assert set(b["indices"][n].tolist()) == set(br["indices"][n].tolist())


@pytest.mark.parametrize("variant", ["gradient_metric", "threshold_merge"])
def test_option_variants_of_the_restatement_match_reference(golden, variant):
    """|gradient| score and threshold-chosen ranks (modeling_grasp.py:393-394, 408-410) against the reference's runs."""
    fx = golden("e2e_tiny_variants.pt")
    ref = fx["variants"][variant]
    opt = ref["options"]
    model = synth.random_llama(fx["model"], seed=fx["seed"])
    assert state_checksum(model) == fx["model_sha256"]
    rec = restate.run_grasp(copy.deepcopy(model), fx["tokens"], num_prune_layers=opt["num_prune_layers"],
                            compression_ratio=opt["ratio"], metric=opt["metric"], merge=opt["merge"],
                            threshold_ratio=opt["threshold_ratio"])
    assert rec["layers_id"] == ref["layers_id"]
    for b, br in zip(rec["blocks"], ref["blocks"]):
        assert b["names"] == br["names"]
        for n in b["names"]:
            assert rel(b["S"][n], br["S"][n]) < 1e-6
            assert rel(b["grads"][n], br["grads"][n]) < 1e-4
            assert rel(b["scores"][n], br["scores"][n]) < 1e-4
            assert b["indices"][n].tolist() == br["indices"][n].tolist()
