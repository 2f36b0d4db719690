"""Synthetic inputs for the five BASELINE.json configurations: random-init LLaMA-family
models (no checkpoints exist offline) and uniform random calibration tokens in the batch
format of the reference loader (dataset/loader.py:24-36: input_ids = row[:-1], labels = row[1:]).
"""
from __future__ import annotations

import torch
from torch.utils.data import DataLoader, Dataset

# name -> LlamaConfig kwargs
MODEL_CONFIGS = {
    "tiny": dict(hidden_size=64, intermediate_size=176, num_hidden_layers=4, num_attention_heads=4,
                 num_key_value_heads=2, vocab_size=256, max_position_embeddings=128),
    "small": dict(hidden_size=256, intermediate_size=704, num_hidden_layers=6, num_attention_heads=8,
                  num_key_value_heads=4, vocab_size=1024, max_position_embeddings=256),
    "tinyllama-1.1b": dict(hidden_size=2048, intermediate_size=5632, num_hidden_layers=22, num_attention_heads=32,
                           num_key_value_heads=4, vocab_size=32000, max_position_embeddings=2048),
    "llama2-7b": dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
                      num_key_value_heads=32, vocab_size=32000, max_position_embeddings=4096),
    "llama3-8b": dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                      num_key_value_heads=8, vocab_size=128256, max_position_embeddings=8192),
}


def llama_config(name: str, **overrides):
    from transformers import LlamaConfig
    kw = dict(MODEL_CONFIGS[name])
    kw.update(overrides)
    kw.setdefault("tie_word_embeddings", False)
    cfg = LlamaConfig(**kw)
    cfg._name_or_path = f"synthetic/{name}"
    return cfg


def random_llama(name: str, seed: int = 0, device="cpu", dtype=torch.float32, **overrides):
    """Random-init LlamaForCausalLM (HF default init, normal std 0.02) built directly on `device`."""
    from transformers import LlamaForCausalLM
    cfg = llama_config(name, **overrides)
    torch.manual_seed(seed)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device(device):
            model = LlamaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(prev)
    model.eval()
    return model


def random_tokens(n_samples: int, seq_len: int, vocab_size: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, vocab_size, (n_samples, seq_len), generator=g)


class TokenRows(Dataset):
    def __init__(self, tokens: torch.Tensor):
        self.tokens = tokens

    def __len__(self):
        return self.tokens.shape[0]

    def __getitem__(self, i):
        row = self.tokens[i]
        return {"input_ids": row[:-1], "labels": row[1:]}


def calibration_dataloader(n_samples: int, seq_len: int, vocab_size: int, batch_size: int = 1, seed: int = 0,
                           tokens: torch.Tensor = None, pin_memory: bool = False) -> DataLoader:
    """Deterministic stand-in for get_calibration_dataloader (shuffle off so runs are repeatable)."""
    if tokens is None:
        tokens = random_tokens(n_samples, seq_len, vocab_size, seed)
    return DataLoader(TokenRows(tokens), batch_size=batch_size, shuffle=False, pin_memory=pin_memory)
