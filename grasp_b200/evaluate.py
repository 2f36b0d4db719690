"""Perplexity of a (compressed) model on a token matrix -- the evaluator of reference
evaluate_grasp.py:99-127 (`evaluate_perplexity`), run through the same layer executor as the
calibration passes: for every row, logits of tokens[:-1] against tokens[1:] (single shift), the mean
cross-entropy per row, and ppl = exp(mean over rows).  The reference feeds one row at a time through
`model(input_ids)`; here several rows share a micro-batch and, on a CUDA device, the forward is
grasp_b200.fused (row kernels + prepared-operand GEMMs) with the loss taken by grasp_ce_loss_bwd.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import engine


@torch.no_grad()
def evaluate_perplexity(model, dataset: torch.Tensor, limit: Optional[int] = None, device="cuda",
                        micro_batch: int = 8) -> float:
    """model: a HF causal LM (or a GRASPModel wrapping one); dataset: [n, seqlen] int64 token ids."""
    hf = getattr(model, "model", model) if not hasattr(model, "lm_head") else model
    n, seqlen = dataset.shape
    if limit is not None and limit >= 0:
        n = min(n, int(limit))
    if n == 0:
        return float("nan")
    dev = torch.device(device)
    hf.to(dev)
    if not engine.LlamaRunner.supports(hf):
        # not a LLaMA-shaped model: plain module call, one row at a time like the reference
        total = 0.0
        for i in range(n):
            row = dataset[i:i + 1].to(dev)
            logits = hf(input_ids=row[:, :-1])[0]
            total += torch.nn.functional.cross_entropy(logits.view(-1, logits.size(-1)).float(), row[:, 1:].reshape(-1)).item()
        return math.exp(total / n)
    runner = engine.LlamaRunner(hf, micro_batch=micro_batch)
    total = torch.zeros((), dtype=torch.float64, device=dev)
    fused = None
    for s in range(0, n, micro_batch):
        rows = dataset[s:min(n, s + micro_batch)].to(dev)
        ids, labels = rows[:, :-1], rows[:, 1:]
        hidden = runner.embed(ids)
        if s == 0:
            fused = runner.fused(hidden)
        B, S, d = hidden.shape
        if fused is not None:
            x = fused.run_layers(hidden, 0, runner.n_layers).reshape(B * S, d)
            logits = fused.lin_fwd(runner.head, fused.be.prep(fused.final_norm(x)), tag="head")
            coef = torch.full((B * S,), 1.0 / S, dtype=torch.float32, device=dev)
            loss_rows = fused.be.ce_loss_bwd_(logits, labels.reshape(-1).contiguous(), coef)
            total += loss_rows.double().sum()
        else:
            with engine.grasp_linear(runner.use_grasp_gemm and hidden.is_cuda):
                x = runner.run_layers(hidden, 0, runner.n_layers)
                logits = runner.head(runner.norm(x))
            per_tok = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.size(-1)).float(), labels.reshape(-1),
                                                        reduction="none")
            total += per_tok.view(B, S).mean(dim=1).double().sum()
    return math.exp(total.item() / n)
