"""TEST INFRASTRUCTURE -- generate tests/golden/*.pt from the REAL reference.

Run in the build container (where /root/reference is mounted):

    python -m oracle.make_golden

Every fixture stores the seeded inputs and the outputs of the unmodified reference functions
(imported by oracle/ref_import.py, CPU, fp32).  While writing them, the script also asserts that
oracle/restate.py reproduces the reference on the same inputs, which is what pins the oracle.
"""
from __future__ import annotations

import copy
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from grasp_b200 import synth  # noqa: E402  (synthetic configs/tokens only; no kernels involved)
from oracle import ref_import, restate  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def state_checksum(model) -> str:
    h = hashlib.sha256()
    for k, v in sorted(model.state_dict().items()):
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def close(a, b, tol, what):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
    assert err <= tol, f"restate != reference for {what}: rel err {err:.3e} > {tol}"


def golden_bi(ref):
    g = torch.Generator().manual_seed(1)
    hs = [torch.randn(2, 17, 96, generator=g) for _ in range(6)]
    hs[2][0, 3] = 0.0                      # zero-norm token -> NaN -> 0.5
    hs[3][1, 5] = hs[2][1, 5]              # identical rows -> BI 0
    hs[4][0, 0] = -hs[3][0, 0]             # opposite rows -> BI 2
    per_pair = [ref.utils_func.block_influence(hs[i], hs[i + 1]) for i in range(5)]
    angular = [ref.utils_func.block_influence(hs[i][:, -1:], hs[i + 1][:, -1:], angular=True) for i in range(5)]
    for i in range(5):
        close(restate.block_influence(hs[i], hs[i + 1]), per_pair[i], 1e-5, "block_influence")
        close(restate.block_influence(hs[i][:, -1:], hs[i + 1][:, -1:], True), angular[i], 1e-5, "angular")
    means = [p.mean().item() for p in per_pair]
    torch.save({"hiddens": hs, "per_pair": per_pair, "angular_last_token": angular, "means": means},
               os.path.join(GOLDEN, "bi_small.pt"))


def golden_svd():
    g = torch.Generator().manual_seed(2)
    cases = []
    for (m, n) in [(48, 32), (64, 64), (40, 100), (128, 96), (7, 5), (1, 9), (130, 70)]:
        A = torch.randn(m, n, generator=g) * 0.02
        U, S, Vh = torch.linalg.svd(A, full_matrices=False)  # the reference call (modeling_grasp.py:231)
        cases.append({"A": A, "U": U, "S": S, "Vh": Vh})
    # rank-deficient and graded cases
    B = torch.randn(64, 8, generator=g) @ torch.randn(8, 80, generator=g)
    U, S, Vh = torch.linalg.svd(B, full_matrices=False)
    cases.append({"A": B, "U": U, "S": S, "Vh": Vh, "note": "rank 8"})
    Q1, _ = torch.linalg.qr(torch.randn(96, 96, generator=g))
    Q2, _ = torch.linalg.qr(torch.randn(96, 96, generator=g))
    Cm = Q1 @ torch.diag(torch.logspace(0, -5, 96)) @ Q2
    U, S, Vh = torch.linalg.svd(Cm, full_matrices=False)
    cases.append({"A": Cm, "U": U, "S": S, "Vh": Vh, "note": "graded 1e0..1e-5"})
    torch.save(cases, os.path.join(GOLDEN, "svd_small.pt"))


def golden_select(ref):
    g = torch.Generator().manual_seed(3)
    out = []
    for (o, i) in [(64, 64), (176, 64), (64, 176), (32, 64)]:
        W = torch.randn(o, i, generator=g) * 0.02
        G = torch.randn(o, i, generator=g)
        U, S, Vh = torch.linalg.svd(W, full_matrices=False)
        layer = ref.modeling.GRASPLayer(U, S, Vh, None, None)
        # reference autograd path for dL/dS with L = <G, W_reconstructed>
        x = torch.eye(i).unsqueeze(0)
        y = layer(x)                                   # [1, i, o] = W^T
        (y[0].t() * G).sum().backward()
        grad = layer.S.grad.clone()
        close(restate.sigma_grad_from_G(U, G, Vh), grad, 2e-5, "sigma_grad identity")
        case = {"W": W, "G": G, "U": U, "S": S, "Vh": Vh, "grad": grad}
        for ratio in (0.9, 0.5):
            k = int(i * o * (1 - ratio) / (i + o))
            assert restate.preserve_rank(i, o, ratio) == k
            for metric in ("taylor", "gradient"):
                sc = torch.abs(grad * S) if metric == "taylor" else torch.abs(grad)
                idx = torch.topk(sc, k=k).indices        # modeling_grasp.py:404
                Wm = torch.mm(U[:, idx], torch.mm(torch.diag(S[idx]), Vh[idx, :]))
                sv = ref.modeling.SVDLinear(U[:, idx], S[idx], Vh[idx, :], None, "UV")
                close(restate.merged_weight(U, S, Vh, idx), Wm, 1e-6, "merged")
                iw, ow = restate.packed_factors(U, S, Vh, idx)
                close(iw, sv.InLinear.weight.data, 1e-7, "in_w")
                close(ow, sv.OutLinear.weight.data, 1e-7, "out_w")
                case[f"{metric}_{ratio}"] = {"k": k, "score": sc, "idx": idx, "merged": Wm,
                                             "in_w": sv.InLinear.weight.data.clone(),
                                             "out_w": sv.OutLinear.weight.data.clone()}
        sc = torch.abs(grad * S)
        keep = ref.utils_func.adaptive_rank_selection(sc, 0.6)
        assert restate.adaptive_rank_selection(sc, 0.6) == keep
        case["threshold_0.6"] = {"score": sc, "idx": torch.tensor(keep)}
        out.append(case)
    torch.save(out, os.path.join(GOLDEN, "select_small.pt"))


def run_reference(ref, model, tokens, num_prune_layers, ratio, merge, metric="taylor", threshold_ratio=None,
                  angular=False):
    """The reference's own GRASPModel methods in grasp.main order (grasp.py:61-126), CPU."""
    batches = restate.batches_from_tokens(tokens)
    gm = ref.modeling.GRASPModel(model)
    imp, layers = gm.compute_bi(num_prune_layers=num_prune_layers, calibration_dataloader=batches, angular=angular,
                                device="cpu")
    rec = {"layer_importances": list(imp), "layers_id": sorted(layers, reverse=True), "blocks": []}
    for lid in rec["layers_id"]:
        for block_type, types in (("mlp", ["down_proj", "up_proj", "gate_proj"]),
                                  ("attention", ["q_proj", "k_proj", "v_proj", "o_proj"])):
            gm.compress_block(layer_id=lid, block_type=block_type, target_layer_types=types, device="cpu")
            names = gm.check_exists_grasp_layer()
            S = {n: gm.model.get_submodule(n).S.data.clone() for n in names}
            grads = gm.get_svdlayer_gradients(batches, "cpu")
            idx = gm.dynamic_svd_selection(grads, metric=metric, compression_ratio=ratio, threshold_ratio=threshold_ratio)
            rec["blocks"].append({"layer": lid, "block": block_type, "names": names, "S": S,
                                  "grads": {n: grads[n].clone() for n in names},
                                  "scores": {n: torch.abs(grads[n] * S[n]) if metric == "taylor" else torch.abs(grads[n])
                                             for n in names},
                                  "indices": {n: torch.as_tensor(idx[n]).clone() for n in names}})
            gm.compile_grasp_model(idx, merge=merge, device="cpu")
    return gm, rec


def golden_e2e(ref, name, n_samples, seq_len, num_prune_layers, ratio, fname):
    cfg_vocab = synth.MODEL_CONFIGS[name]["vocab_size"]
    tokens = synth.random_tokens(n_samples, seq_len, cfg_vocab, seed=0)
    fixture = {"model": name, "seed": 0, "tokens": tokens, "ratio": ratio, "num_prune_layers": num_prune_layers}
    for merge in (False, True):
        model = synth.random_llama(name, seed=0)
        fixture["model_sha256"] = state_checksum(model)
        ppl0 = restate.perplexity(model, tokens)
        gm, rec = run_reference(ref, copy.deepcopy(model), tokens, num_prune_layers, ratio, merge)
        rec["ppl_dense"] = ppl0
        rec["ppl_compressed"] = restate.perplexity(gm.model, tokens)
        if name == "tiny" or not merge:  # keep the committed fixtures small
            rec["final_state"] = {k: v.clone() for k, v in gm.model.state_dict().items()
                                  if any(f"layers.{l}." in k for l in rec["layers_id"])}
        # pin the restatement against the reference on the same model
        rec2 = restate.run_grasp(copy.deepcopy(model), tokens, num_prune_layers=num_prune_layers,
                                 compression_ratio=ratio, merge=merge)
        assert rec2["layers_id"] == rec["layers_id"], (rec2["layers_id"], rec["layers_id"])
        close(rec2["layer_importances"], rec["layer_importances"], 1e-5, "BI")
        for b_ref, b_re in zip(rec["blocks"], rec2["blocks"]):
            assert b_ref["names"] == b_re["names"]
            for n in b_ref["names"]:
                close(b_re["S"][n], b_ref["S"][n], 1e-6, f"S {n}")
                close(b_re["grads"][n], b_ref["grads"][n], 1e-4, f"grad {n}")
                assert set(b_re["indices"][n].tolist()) == set(b_ref["indices"][n].tolist()), n
        fixture["merge" if merge else "factored"] = rec
        print(f"{fname} merge={merge}: layers {rec['layers_id']} ppl {ppl0:.3f} -> {rec['ppl_compressed']:.3f}")
    torch.save(fixture, os.path.join(GOLDEN, fname))


VARIANTS = {
    # the option branches of the same path: the |gradient| score (modeling_grasp.py:393-394); rank chosen by the
    # cumulative-score threshold (:408-410) with merged rebuild.  (compute_bi(angular=True) cannot be pinned: the
    # reference's own branch raises UnboundLocalError at :154; block_influence(angular=True) itself is in bi_small.pt.)
    "gradient_metric": dict(num_prune_layers=2, ratio=0.5, merge=False, metric="gradient", threshold_ratio=None,
                            angular=False),
    "threshold_merge": dict(num_prune_layers=1, ratio=None, merge=True, metric="taylor", threshold_ratio=0.7,
                            angular=False),
}


def golden_e2e_variants(ref, fname="e2e_tiny_variants.pt"):
    name = "tiny"
    tokens = synth.random_tokens(6, 33, synth.MODEL_CONFIGS[name]["vocab_size"], seed=1)
    model = synth.random_llama(name, seed=0)
    fixture = {"model": name, "seed": 0, "tokens": tokens, "model_sha256": state_checksum(model), "variants": {}}
    for vname, kw in VARIANTS.items():
        gm, rec = run_reference(ref, copy.deepcopy(model), tokens, kw["num_prune_layers"], kw["ratio"], kw["merge"],
                                metric=kw["metric"], threshold_ratio=kw["threshold_ratio"], angular=kw["angular"])
        rec["options"] = dict(kw)
        rec["ppl_compressed"] = restate.perplexity(gm.model, tokens)
        rec["final_state"] = {k: v.clone() for k, v in gm.model.state_dict().items()
                              if any(f"layers.{l}." in k for l in rec["layers_id"])}
        rec2 = restate.run_grasp(copy.deepcopy(model), tokens, num_prune_layers=kw["num_prune_layers"],
                                 compression_ratio=kw["ratio"], metric=kw["metric"], merge=kw["merge"],
                                 threshold_ratio=kw["threshold_ratio"])
        assert rec2["layers_id"] == rec["layers_id"], (rec2["layers_id"], rec["layers_id"])
        close(rec2["layer_importances"], rec["layer_importances"], 1e-5, "BI")
        for b_ref, b_re in zip(rec["blocks"], rec2["blocks"]):
            for n in b_ref["names"]:
                close(b_re["S"][n], b_ref["S"][n], 1e-6, f"S {n}")
                close(b_re["grads"][n], b_ref["grads"][n], 1e-4, f"grad {n}")
                assert set(b_re["indices"][n].tolist()) == set(b_ref["indices"][n].tolist()), n
        fixture["variants"][vname] = rec
        print(f"{fname} {vname}: layers {rec['layers_id']} ranks "
              f"{[len(b['indices'][n]) for b in rec['blocks'] for n in b['names']]} ppl {rec['ppl_compressed']:.3f}")
    # SVDLinear's other sigma placements (modeling_grasp.py:49-54)
    g = torch.Generator().manual_seed(5)
    U, S, Vh = torch.linalg.svd(torch.randn(48, 80, generator=g), full_matrices=False)
    idx = torch.tensor([0, 3, 4, 9, 17])
    fuse = {}
    for mode in ("U", "V"):
        m = ref.modeling.SVDLinear(U[:, idx], S[idx], Vh[idx, :], None, mode)
        fuse[mode] = {"in_w": m.InLinear.weight.data.clone()}
        if mode == "U":
            fuse[mode]["out_w"] = m.OutLinear.weight.data.clone()
    fixture["sigma_fuse"] = {"U": U[:, idx].clone(), "S": S[idx].clone(), "Vh": Vh[idx, :].clone(), "modes": fuse}
    torch.save(fixture, os.path.join(GOLDEN, fname))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    ref = ref_import.load()
    if "variants" in sys.argv[1:]:      # only the option-variant fixture (the others stay byte-identical)
        golden_e2e_variants(ref)
        return
    golden_bi(ref)
    golden_svd()
    golden_select(ref)
    golden_e2e(ref, "tiny", n_samples=6, seq_len=33, num_prune_layers=2, ratio=0.5, fname="e2e_tiny.pt")
    golden_e2e(ref, "small", n_samples=4, seq_len=65, num_prune_layers=2, ratio=0.8, fname="e2e_small.pt")
    golden_e2e_variants(ref)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
