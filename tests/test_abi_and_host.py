"""CPU: the C-ABI library loads and exports every symbol of include/grasp_b200.h; host-side logic
of the drop-in front (flags, names, rank formula, loader format).  No compute call needs a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "grasp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(grasp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from grasp_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/grasp_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.grasp_abi_version() == 1


def test_bad_arguments_are_rejected_before_any_launch():
    from grasp_b200 import _lib
    lib = _lib.load()
    before = lib.grasp_launch_count()
    assert lib.grasp_topk_batched(-1, None, None, None, None, None) < 0
    assert b"batch" in lib.grasp_last_error()
    assert lib.grasp_bi_accumulate(None, None, 4, 8, 8, 0, 0, 1.0, None, None, None) < 0
    assert lib.grasp_svd_batched(1, None, None, None, None, None, None, None, None, 0, 0, None, 0, None) < 0
    assert lib.grasp_sigma_score(None, None, None, None, 4, 4, 4, 1, 0, None, None, 0, None, 0, None) < 0
    assert lib.grasp_lowrank_rebuild(None, None, None, None, 1, 4, 4, 4, 0, None, 0, None, 0, None) < 0
    assert lib.grasp_gemm_f32(0, 0, 4, 4, 4, 1.0, None, 4, None, 4, 0.0, None, 4, 0, None, 0, None) < 0
    assert lib.grasp_gemm_split_f16(None, 8, 4, 8, 0, None, None, None) < 0
    assert lib.grasp_gemm_f16x3_planes(4, 4, 4, 1.0, None, None, None, 0, None, 0.0, None, 4, None) < 0
    assert lib.grasp_rmsnorm_fwd(None, None, 4, 8, 1e-5, None, None, None, None, None) < 0
    assert lib.grasp_rmsnorm_bwd(None, None, None, None, None, 4, 8, None, None) < 0
    assert lib.grasp_rope_inplace(None, 4, 4, 1, 8, None, None, 0, 0, None) < 0
    assert lib.grasp_swiglu_fwd(None, None, 4, 8, None, None, None, None) < 0
    assert lib.grasp_swiglu_bwd(None, None, None, 4, 8, None, None, None, None, None, None, None) < 0
    assert lib.grasp_ce_loss_bwd(None, None, None, 4, 8, None, None) < 0
    assert lib.grasp_gemm_planes_bytes(8176, 4096) == 2 * 8176 * 4096 * 2      # two fp16 planes, already 1 KiB-aligned
    assert lib.grasp_gemm_planes_bytes(3, 5) == 1024                            # cols padded to 8, size to 1 KiB
    assert lib.grasp_gemm_planes_bytes(0, 5) == 0
    m = (ctypes.c_int64 * 1)(4096)
    assert lib.grasp_svd_workspace_bytes(1, m, m) > 4096 * 8192 * 4
    assert lib.grasp_launch_count() == before


def test_ops_refuse_cpu_tensors():
    from grasp_b200 import ops
    from grasp_b200._lib import GraspLibraryError
    with pytest.raises(GraspLibraryError):
        ops.svd(torch.randn(8, 8))
    with pytest.raises(GraspLibraryError):
        ops.topk(torch.randn(8), 2)
    with pytest.raises(GraspLibraryError):
        ops.bi_accumulate(torch.randn(1, 2, 8), torch.randn(1, 2, 8), per_row=True)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from grasp_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setenv("GRASP_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GraspLibraryError):
        _lib.load()


def test_cli_flags_match_reference_surface():
    import grasp
    a = grasp.parse_args(["--model_name_or_path", "x", "--num_prune_layers", "8", "--compression_ratio", "0.9",
                          "--num_samples", "512", "--mlp_target_layer_types", "down_proj,up_proj", "--merge",
                          "--layers_id", "3", "4"])
    assert a.num_prune_layers == 8 and a.compression_ratio == 0.9 and a.num_samples == 512
    assert a.mlp_target_layer_types == ["down_proj", "up_proj"]
    assert a.attn_target_layer_types == ["q_proj", "k_proj", "v_proj", "o_proj"]
    assert a.merge and not a.recovery and a.layers_id == [3, 4]
    assert a.metric == "taylor" and a.batch_size == 1 and a.seq_len == 512 and a.dataset_name == "wikitext2"
    for flag in ("threshold_ratio", "angular", "allocation_aware", "verbose", "padding", "log_file", "data_path",
                 "train_batch_size", "micro_batch_size", "num_epochs", "learning_rate", "max_length",
                 "val_set_size", "train_on_inputs", "add_eos_token", "resume_from_checkpoint",
                 "prompt_template_name", "train_device", "evaluate", "eval_ppl", "eval_tasks", "num_fewshot",
                 "limit", "save_path", "device"):
        assert hasattr(a, flag), flag


def test_params_script_keeps_the_reference_knobs():
    text = open(os.path.join(ROOT, "scripts", "params_script.sh")).read()
    for knob in ("NUM_PRUNE_LAYERS=7", "COMPRESSION_RATIO=0.9", "NUM_SAMPLES=512", "SEQ_LEN=512", "BATCH_SIZE=1",
                 'METRIC="taylor"', "MERGE=false"):
        assert knob in text, knob


def test_model_front_names_and_errors():
    from grasp_b200 import synth
    from modeling_grasp import GRASPLayer, GRASPModel, SVDLinear  # noqa: F401  (pickle-visible names)
    gm = GRASPModel(synth.random_llama("tiny"))
    assert all(not p.requires_grad for p in gm.model.parameters())
    assert gm.block_target_names(3, "mlp", ["down_proj"]) == ["model.layers.3.mlp.down_proj"]
    assert gm.block_target_names(1, "attention", None)[0] == "model.layers.1.self_attn.q_proj"
    with pytest.raises(ValueError):
        gm.compress_block(0, "mlp", ["q_proj"])
    with pytest.raises(NotImplementedError):
        gm.compress_block(0, "conv", ["q_proj"])
    with pytest.raises(ValueError):
        gm.compress_block(None, "mlp", ["down_proj"])
    assert gm.compress_block(0, "mlp", None) is True
    with pytest.raises(TypeError):
        gm.replace_with_GRASPLayer("model.layers.0.mlp")

    class L:  # compute_preserve_rank only reads the two feature counts
        in_features, out_features = 4096, 4096
    assert gm.compute_preserve_rank(L, 0.9) == 204
    with pytest.raises(ValueError):
        gm.compute_preserve_rank(L, None)
    with pytest.raises(ValueError):
        gm.dynamic_svd_selection({})
    assert gm._extract_layer_index("model.layers.23.mlp") == 23


def test_synthetic_loader_has_the_reference_batch_format():
    from grasp_b200 import synth
    dl = synth.calibration_dataloader(5, 16, 100, batch_size=2)
    batches = list(dl)
    assert len(batches) == 3 and len(batches[0]) == 2          # 2 keys => attention_mask=None in the model code
    b = batches[0]
    assert b["input_ids"].shape == (2, 15) and b["labels"].shape == (2, 15)
    toks = synth.random_tokens(5, 16, 100)
    assert torch.equal(b["input_ids"], toks[:2, :-1]) and torch.equal(b["labels"], toks[:2, 1:])


def test_sigma_fuse_placements_match_the_reference(golden):
    """SVDLinear with the singular values on one side only (modeling_grasp.py:49-54), against the reference's module.
    "V" leaves OutLinear at its default initialisation there and here, so only InLinear is compared."""
    from modeling_grasp import SVDLinear
    fx = golden("e2e_tiny_variants.pt")["sigma_fuse"]
    for mode, want in fx["modes"].items():
        m = SVDLinear(fx["U"], fx["S"], fx["Vh"], None, mode)
        assert torch.equal(m.InLinear.weight.data, want["in_w"])
        if "out_w" in want:
            assert torch.equal(m.OutLinear.weight.data, want["out_w"])
