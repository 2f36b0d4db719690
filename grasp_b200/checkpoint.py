"""State-dict checkpoint of a compressed model (replaces the fragile whole-module pickle of
reference grasp.py:129-136 when transformers' hooks make the module unpicklable).

File = torch.save({"config": HF config dict, "structure": {module name: kind/shape}, "state_dict": ...}).
`load` rebuilds the HF model from the config, swaps SVDLinear / merged Linear modules back in by
name, and returns a modeling_grasp.GRASPModel, so `evaluate.py`-style callers get `.model`.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def save(grasp_model, path: str) -> None:
    from modeling_grasp import SVDLinear
    structure = {}
    for name, module in grasp_model.model.named_modules():
        if isinstance(module, SVDLinear):
            structure[name] = {"kind": "svd", "rank": module.InLinear.out_features,
                               "in": module.InLinear.in_features, "out": module.OutLinear.out_features,
                               "bias": module.OutLinear.bias is not None}
    cfg = grasp_model.model.config
    torch.save({"config_class": type(cfg).__name__, "config": cfg.to_dict(), "structure": structure,
                "state_dict": grasp_model.model.state_dict()}, path)


def load(path: str, device="cpu"):
    import transformers
    from modeling_grasp import GRASPModel, SVDLinear
    blob = torch.load(path, map_location="cpu", weights_only=False)
    cfg = getattr(transformers, blob["config_class"])(**blob["config"])
    with torch.device("meta"):
        model = transformers.AutoModelForCausalLM.from_config(cfg)
    for name, s in blob["structure"].items():
        new = SVDLinear.__new__(SVDLinear)
        nn.Module.__init__(new)
        new.InLinear = nn.Linear(s["in"], s["rank"], bias=False, device="meta")
        new.OutLinear = nn.Linear(s["rank"], s["out"], bias=s["bias"], device="meta")
        *parents, leaf = name.split(".")
        owner = model
        for p in parents:
            owner = getattr(owner, p)
        setattr(owner, leaf, new)
    model.load_state_dict(blob["state_dict"], assign=True, strict=False)
    model.to(device)
    return GRASPModel(model)
