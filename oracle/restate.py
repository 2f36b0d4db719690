"""TEST INFRASTRUCTURE -- CPU restatement of the reference GRASP hot path (torch CPU, fp32).

Each function follows the cited lines of the reference (paths relative to the reference root,
/root/reference in the build container).  The arithmetic of the path lives in third-party code
the reference calls -- torch (`torch.linalg.svd` -> LAPACK sgesdd, `torch.mm`, `torch.topk`,
autograd; README pins torch==2.3.1, this image has 2.11.0) and transformers' LLaMA forward/loss
(requirements.txt pins 4.45.2, this image has 5.5.0) -- so the restatement calls the same
library entry points on the CPU.  Pinned by tests/golden/ (see oracle/__init__.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn


# ---------------------------------------------------------------------------- stage 1
def block_influence(x: torch.Tensor, y: torch.Tensor, angular: bool = False) -> torch.Tensor:
    """tools/utils_func.py:3-25.  x, y: [B, S, D] -> [B*S]."""
    d = x.shape[-1]
    x = x.reshape(-1, d).float()
    y = y.reshape(-1, d).float()
    nx = x.norm(dim=-1)
    ny = y.norm(dim=-1)
    # the reference takes the diagonal of the full Gram (:19-20); the diagonal is the row-wise dot
    sim = (x * y).sum(-1) / (nx * ny)
    sim = sim.nan_to_num(nan=0.5)
    if angular:
        return torch.arccos(sim) / torch.pi
    return 1 - sim


def compute_bi_hiddens(hiddens: Sequence[torch.Tensor], importances: List[float]) -> None:
    """modeling_grasp.py:148-167, non-angular branch: stride 1, += mean over tokens (python float)."""
    for i in range(len(hiddens) - 1):
        importances[i] += block_influence(hiddens[i], hiddens[i + 1]).mean().cpu().item()


def compute_bi(model, batches, num_prune_layers: int):
    """modeling_grasp.py:135-193 (non-angular).  Returns (importances, layers_to_remove)."""
    importances = [0 for _ in model.model.layers]
    for batch in batches:
        with torch.no_grad():
            out = model(input_ids=batch["input_ids"], attention_mask=None, use_cache=False,
                        output_hidden_states=True, return_dict=True)
        compute_bi_hiddens(out.hidden_states, importances)
    layers = np.argsort(np.array(importances))[:num_prune_layers].tolist()
    return importances, layers


# ---------------------------------------------------------------------------- stage 2
def svd(w: torch.Tensor):
    """modeling_grasp.py:230-231: thin SVD, fp32, S descending."""
    return torch.linalg.svd(w.float(), full_matrices=False)


class OracleGRASPLayer(nn.Module):
    """modeling_grasp.py:62-79: y = x (U diag(S) Vh)^T, only S trainable, bias never applied."""

    def __init__(self, U, S, Vh, bias=None, compression_ratio=None):
        super().__init__()
        self.U = nn.Parameter(U.clone().detach().requires_grad_(False))
        self.S = nn.Parameter(S.clone().detach().requires_grad_(True))
        self.Vh = nn.Parameter(Vh.clone().detach().requires_grad_(False))
        self.in_features = self.Vh.shape[1]
        self.out_features = self.U.shape[0]
        self.bias = bias
        self.compression_ratio = compression_ratio

    def forward(self, x):
        b, s, _ = x.shape
        w = torch.mm(self.U, torch.mm(torch.diag(self.S), self.Vh))
        return torch.mm(x.view(b * s, -1), w.t()).view(b, s, -1)


def _set_module(model, key, module):
    """modeling_grasp.py:218-223."""
    *parents, leaf = key.split(".")
    owner = model
    for p in parents:
        owner = getattr(owner, p)
    setattr(owner, leaf, module)


def block_names(layer_id: int, block_type: str, types: Sequence[str]) -> List[str]:
    """modeling_grasp.py:264-286."""
    prefix = {"attention": "self_attn.", "mlp": "mlp."}[block_type]
    return [f"model.layers.{layer_id}.{prefix}{t}" for t in types]


def compress_block(model, layer_id, block_type, types) -> None:
    """modeling_grasp.py:244-309 + :225-242."""
    for name in block_names(layer_id, block_type, types):
        lin = model.get_submodule(name)
        assert isinstance(lin, nn.Linear)
        U, S, Vh = svd(lin.weight.data)
        _set_module(model, name, OracleGRASPLayer(U, S, Vh, lin.bias, getattr(lin, "compression_ratio", None)))


def grasp_layer_names(model) -> List[str]:
    """modeling_grasp.py:319-329: named_modules order (gate, up, down / q, k, v, o)."""
    return [n for n, m in model.named_modules() if isinstance(m, OracleGRASPLayer)]


# ---------------------------------------------------------------------------- stage 3a
def svdlayer_gradients(model, batches) -> Dict[str, torch.Tensor]:
    """modeling_grasp.py:331-370: sum over batches of dLoss/dS; loss = HF causal-LM loss on the
    loader's already shifted labels (double shift)."""
    names = grasp_layer_names(model)
    grads: Dict[str, torch.Tensor] = {}
    for batch in batches:
        out = model(input_ids=batch["input_ids"], attention_mask=None, labels=batch["labels"], use_cache=False)
        loss = out[0]
        model.zero_grad()
        loss.backward()
        for n in names:
            g = model.get_submodule(n).S.grad
            grads[n] = g if n not in grads else grads[n] + g
    return grads


def sigma_grad_from_G(U: torch.Tensor, G: torch.Tensor, Vh: torch.Tensor) -> torch.Tensor:
    """The identity the CUDA path uses: dL/dS_i = u_i^T G v_i (autograd of modeling_grasp.py:77-79)."""
    return ((U.t() @ G) * Vh).sum(-1)


# ---------------------------------------------------------------------------- stage 3b
def preserve_rank(in_features: int, out_features: int, ratio: float) -> int:
    """modeling_grasp.py:311-317 (python float64 arithmetic, truncation)."""
    return int(in_features * out_features * (1 - ratio) / (in_features + out_features))


def importance(grad: torch.Tensor, S: torch.Tensor, metric: str) -> torch.Tensor:
    """modeling_grasp.py:392-397."""
    if metric == "gradient":
        return torch.abs(grad)
    if metric == "taylor":
        return torch.abs(grad * S)
    raise RuntimeError(f"{metric} not support")


def adaptive_rank_selection(scores, target_ratio: float) -> List[int]:
    """tools/utils_func.py:45-57 (sequential fp32 sums, stable descending sort)."""
    total = sum(scores)
    target = total * target_ratio
    order = sorted(enumerate(scores), key=lambda t: -t[1])
    run, keep = 0, []
    for i, v in order:
        run += v
        keep.append(i)
        if run >= target:
            break
    return keep


def select(grads: Dict[str, torch.Tensor], layers: Dict[str, OracleGRASPLayer], metric="taylor",
           compression_ratio: Optional[float] = None, threshold_ratio: Optional[float] = None):
    """modeling_grasp.py:372-421.  Returns ({name: index tensor/list}, {name: score})."""
    out, scores = {}, {}
    for name, g in grads.items():
        layer = layers[name]
        sc = importance(g, layer.S.data, metric)
        if layer.compression_ratio is not None:
            compression_ratio = layer.compression_ratio
        if compression_ratio is not None:
            k = preserve_rank(layer.in_features, layer.out_features, compression_ratio)
            out[name] = torch.topk(sc, k=k).indices
        else:
            assert threshold_ratio
            out[name] = adaptive_rank_selection(sc, threshold_ratio)
        scores[name] = sc
    return out, scores


# ---------------------------------------------------------------------------- stage 3c
def merged_weight(U, S, Vh, idx) -> torch.Tensor:
    """modeling_grasp.py:440-442 + :454."""
    idx = torch.as_tensor(idx)
    return torch.mm(U[:, idx], torch.mm(torch.diag(S[idx]), Vh[idx, :]))


def packed_factors(U, S, Vh, idx):
    """modeling_grasp.py:440-442 + :47-48 -> (InLinear.weight [k,in], OutLinear.weight [out,k])."""
    idx = torch.as_tensor(idx)
    s = S[idx]
    return Vh[idx, :].mul(s.sqrt().view(-1, 1)).contiguous(), U[:, idx].mul(s.sqrt()).contiguous()


class OracleSVDLinear(nn.Module):
    """modeling_grasp.py:25-59 with sigma_fuse="UV"."""

    def __init__(self, in_w, out_w, bias=None):
        super().__init__()
        k, i = in_w.shape
        o = out_w.shape[0]
        self.InLinear = nn.Linear(i, k, bias=False)
        self.OutLinear = nn.Linear(k, o, bias=bias is not None)
        self.InLinear.weight.data = in_w
        self.OutLinear.weight.data = out_w
        if bias is not None:
            self.OutLinear.bias.data = bias

    def forward(self, x):
        return self.OutLinear(self.InLinear(x))


def compile_model(model, indices: Dict[str, torch.Tensor], merge: bool) -> None:
    """modeling_grasp.py:423-469."""
    for name, idx in indices.items():
        layer: OracleGRASPLayer = model.get_submodule(name)
        U, S, Vh = layer.U.data, layer.S.data, layer.Vh.data
        if merge:
            lin = nn.Linear(layer.in_features, layer.out_features, bias=layer.bias is not None)
            lin.weight.data = merged_weight(U, S, Vh, idx)
            if layer.bias is not None:
                lin.bias = layer.bias
            lin.requires_grad_(False)
            _set_module(model, name, lin)
        else:
            new = OracleSVDLinear(*packed_factors(U, S, Vh, idx), layer.bias)
            new.requires_grad_(False)
            _set_module(model, name, new)


# ---------------------------------------------------------------------------- driver
def perplexity(model, tokens: torch.Tensor) -> float:
    """evaluate_grasp.py:99-127 on a [n, seqlen] token tensor (single shift, CE mean per row)."""
    n, seqlen = tokens.shape
    nlls = []
    with torch.no_grad():
        for i in range(n):
            logits = model(input_ids=tokens[i:i + 1, :-1])[0]
            loss = nn.CrossEntropyLoss()(logits.view(-1, logits.size(-1)), tokens[i:i + 1, 1:].reshape(-1))
            nlls.append(loss.float() * seqlen)
    return torch.exp(torch.stack(nlls).sum() / (len(nlls) * seqlen)).item()


def batches_from_tokens(tokens: torch.Tensor, batch_size: int = 1):
    """dataset/loader.py:24-36 batch format, deterministic order."""
    out = []
    for s in range(0, tokens.shape[0], batch_size):
        rows = tokens[s:s + batch_size]
        out.append({"input_ids": rows[:, :-1], "labels": rows[:, 1:]})
    return out


def run_grasp(model, tokens: torch.Tensor, layers_id=None, num_prune_layers=None, compression_ratio=0.9,
              metric="taylor", merge=False, threshold_ratio=None,
              mlp_types=("down_proj", "up_proj", "gate_proj"), attn_types=("q_proj", "k_proj", "v_proj", "o_proj"),
              record=None):
    """grasp.py:61-126: BI -> descending layers -> per layer MLP then attention.  `record` (dict)
    receives per-stage artefacts for parity tests."""
    for p in model.parameters():
        p.requires_grad = False  # modeling_grasp.py:86-87
    batches = batches_from_tokens(tokens)
    rec = record if record is not None else {}
    if layers_id is None:
        imp, layers_id = compute_bi(model, batches, num_prune_layers)
        rec["layer_importances"] = imp
    layers_id = sorted(layers_id, reverse=True)
    rec["layers_id"] = list(layers_id)
    rec["blocks"] = []
    for lid in layers_id:
        for block_type, types in (("mlp", mlp_types), ("attention", attn_types)):
            compress_block(model, lid, block_type, types)
            names = grasp_layer_names(model)
            layers = {n: model.get_submodule(n) for n in names}
            grads = svdlayer_gradients(model, batches)
            idx, scores = select(grads, layers, metric, compression_ratio, threshold_ratio)
            rec["blocks"].append({
                "layer": lid, "block": block_type, "names": names,
                "S": {n: layers[n].S.data.clone() for n in names},
                "grads": {n: grads[n].clone() for n in names},
                "scores": {n: scores[n].clone() for n in names},
                "indices": {n: torch.as_tensor(idx[n]).clone() for n in names},
            })
            compile_model(model, idx, merge)
    return rec
