"""Calibration batches in the format of the reference's loader (dataset/loader.py:10-107):
raw text rows are sampled with python's `random` under a fixed seed, joined with blank lines, tokenised as one stream
and cut into rows of `seq_len` tokens; an item is {"input_ids": row[:-1], "labels": row[1:]} (the labels are
pre-shifted, and HF's loss shifts again -- the reference's double shift, which the engine reproduces); batches are
dicts of exactly two tensors, which is what makes `compute_bi` / `get_svdlayer_gradients` pass attention_mask=None.

The text corpora themselves are not part of this repository (no network): `wikitext2` and `c4` are read with
`datasets.load_from_disk` from `<data_root>/wikitext/train` and `<data_root>/c4/train` exactly where the reference
expects them, `texts=` takes the rows directly, and `synthetic` draws uniform random tokens.
"""
from __future__ import annotations

import random
from typing import List, Optional, Sequence

import torch
from torch.utils.data import DataLoader, Dataset

_DISK = {"wikitext2": ("wikitext", "train", "text"), "c4": ("c4", "train", "text")}


class ShiftedRows(Dataset):
    """[n, seq_len] token rows -> {"input_ids": row[:-1], "labels": row[1:]} (reference loader.py:24-36)."""

    def __init__(self, rows: torch.Tensor):
        self.rows = rows

    def __len__(self):
        return self.rows.shape[0]

    def __getitem__(self, i):
        row = self.rows[i]
        return {"input_ids": row[:-1], "labels": row[1:]}


def rows_from_texts(texts: Sequence[str], tokenizer, seq_len: int) -> torch.Tensor:
    """One token stream of the rows joined by blank lines, cut into total // seq_len rows (loader.py:59-68);
    the tail that does not fill a row is dropped."""
    ids = tokenizer("\n\n".join(texts), return_tensors="pt").input_ids[0]
    n = ids.numel() // seq_len
    if n == 0:
        raise ValueError(f"the sampled text holds {ids.numel()} tokens, fewer than one row of {seq_len}")
    return ids[: n * seq_len].view(n, seq_len).clone()


def sample_rows(n_total: int, num_samples: int, seed: int) -> List[int]:
    """The reference's `random.seed(seed); random.sample(range(len(data)), num_samples)` (loader.py:20,82)."""
    rng = random.Random(seed)
    return rng.sample(range(n_total), num_samples)


def get_calibration_dataloader(dataset_name: str, tokenizer, num_samples: int = 128, seq_len: int = 2048,
                               padding="max_length", batch_size: int = 1, seed: int = 42, mix: bool = False,
                               data_root: str = "datasets", texts: Optional[Sequence[str]] = None, shuffle: bool = True):
    """Same arguments and return value as the reference's function; `padding` is accepted and, as there, unused by the
    pretraining-text branch.  NUM_SAMPLES counts raw text rows, so the number of sequences is total_tokens // seq_len."""
    if dataset_name == "synthetic":
        from . import synth
        vocab = getattr(tokenizer, "vocab_size", None) or 32000
        rows = synth.random_tokens(num_samples, seq_len, vocab, seed=seed)
    else:
        if texts is None:
            key = next((k for k in _DISK if k in dataset_name), None)
            if key is None:
                raise NotImplementedError(dataset_name)
            try:
                from datasets import load_from_disk
            except ImportError as exc:
                raise NotImplementedError("reading calibration corpora needs the `datasets` package") from exc
            folder, split, field = _DISK[key]
            data = load_from_disk(f"{data_root}/{folder}/{split}")
            data = data.select(sample_rows(len(data), num_samples, seed))
            texts = data[field]
        else:
            texts = list(texts)
            if num_samples is not None and num_samples < len(texts):
                texts = [texts[i] for i in sample_rows(len(texts), num_samples, seed)]
        rows = rows_from_texts(texts, tokenizer, seq_len)
    dataset = ShiftedRows(rows)
    if mix:
        return dataset
    return DataLoader(dataset, batch_size=batch_size, shuffle=shuffle)
