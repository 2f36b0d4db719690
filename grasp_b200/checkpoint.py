"""State-dict checkpoint of a compressed model (replaces the fragile whole-module pickle of
reference grasp.py:129-136 when transformers' hooks make the module unpicklable).

File = torch.save({"config": HF config dict, "structure": {module name: kind/shape}, "state_dict": ...}).
`load` rebuilds the HF model from the config, swaps SVDLinear modules back in by name (merged layers are
plain nn.Linear of the original shape and need no entry), and returns a modeling_grasp.GRASPModel, so
`evaluate.py`-style callers (reference evaluate.py:42: `torch.load(path).model`) get `.model`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

FORMAT = "grasp_b200.checkpoint.v2"


def save(grasp_model, path: str) -> None:
    from modeling_grasp import GRASPLayer, SVDLinear
    structure = {}
    for name, module in grasp_model.model.named_modules():
        if isinstance(module, SVDLinear):
            structure[name] = {"kind": "svd", "rank": module.InLinear.out_features,
                               "in": module.InLinear.in_features, "out": module.OutLinear.out_features,
                               "bias": module.OutLinear.bias is not None}
        elif isinstance(module, GRASPLayer):
            raise ValueError(f"{name} is still a GRASPLayer: run compile_grasp_model before saving")
    cfg = grasp_model.model.config
    state = {k: v.detach().cpu() for k, v in grasp_model.model.state_dict().items()}
    torch.save({"format": FORMAT, "config_class": type(cfg).__name__, "config": cfg.to_dict(),
                "structure": structure, "state_dict": state,
                "redundant_layers": getattr(grasp_model, "redundant_layers", None)}, path)


def _set(model, name, module):
    *parents, leaf = name.split(".")
    owner = model
    for p in parents:
        owner = getattr(owner, p)
    setattr(owner, leaf, module)


def load(path: str, device="cpu"):
    import transformers
    from modeling_grasp import GRASPModel, SVDLinear
    blob = torch.load(path, map_location="cpu", weights_only=False)
    cfg = getattr(transformers, blob["config_class"])(**blob["config"])
    # build without allocating, give every parameter real (uninitialised) storage, then fill from the file
    with torch.device("meta"):
        model = transformers.AutoModelForCausalLM.from_config(cfg)
    for name, s in blob["structure"].items():
        new = SVDLinear.__new__(SVDLinear)
        nn.Module.__init__(new)
        new.InLinear = nn.Linear(s["in"], s["rank"], bias=False, device="meta")
        new.OutLinear = nn.Linear(s["rank"], s["out"], bias=s["bias"], device="meta")
        _set(model, name, new)
    model.to_empty(device=device)
    # non-persistent buffers (rotary inv_freq) are not in a state dict: re-create their modules from the config
    for name, module in list(model.named_modules()):
        if type(module).__name__.endswith("RotaryEmbedding"):
            _set(model, name, type(module)(config=cfg, device=device))
    missing, unexpected = model.load_state_dict(blob["state_dict"], strict=False)
    tied = getattr(cfg, "tie_word_embeddings", False)
    missing = [k for k in missing if not (tied and k == "lm_head.weight")]
    if missing or unexpected:
        raise RuntimeError(f"checkpoint does not match the rebuilt model: missing {missing[:5]}, "
                           f"unexpected {unexpected[:5]}")
    if tied:
        model.tie_weights()
    model.requires_grad_(False)
    model.eval()
    gm = GRASPModel(model)
    if blob.get("redundant_layers") is not None:
        gm.redundant_layers = blob["redundant_layers"]
    return gm
