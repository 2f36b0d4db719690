"""CPU: calibration batches in the reference loader's format (dataset/loader.py:24-36, 59-68, 75-107)."""
import random

import pytest
import torch

from grasp_b200 import loader


class WordTokenizer:
    """Whitespace tokenizer with the call signature the loader uses."""
    vocab_size = 1000

    def __call__(self, text, return_tensors=None):
        ids = [hash(w) % self.vocab_size for w in text.split()]

        class Out:
            input_ids = torch.tensor([ids])
        return Out()


def test_text_rows_follow_the_reference_layout():
    texts = [" ".join(f"w{i}_{j}" for j in range(37)) for i in range(50)]
    tok = WordTokenizer()
    dl = loader.get_calibration_dataloader("wikitext2", tok, num_samples=20, seq_len=64, batch_size=2, seed=42, texts=texts,
                                           shuffle=False)
    picked = loader.sample_rows(50, 20, 42)
    random.seed(42)
    assert picked == random.sample(range(50), 20)                       # the reference's sampling
    stream = tok("\n\n".join(texts[i] for i in picked)).input_ids[0]
    n = stream.numel() // 64
    assert len(dl.dataset) == n == (20 * 37) // 64                     # NUM_SAMPLES counts text rows, not sequences
    batch = next(iter(dl))
    assert set(batch.keys()) == {"input_ids", "labels"} and len(batch) == 2   # -> attention_mask=None downstream
    assert batch["input_ids"].shape == (2, 63) and batch["labels"].shape == (2, 63)
    assert torch.equal(batch["input_ids"][0], stream[:63]) and torch.equal(batch["labels"][0], stream[1:64])
    assert torch.equal(batch["labels"][0][:-1], batch["input_ids"][0][1:])      # labels are the inputs shifted by one
    ds = loader.get_calibration_dataloader("wikitext2", tok, num_samples=20, seq_len=64, texts=texts, mix=True)
    assert len(ds) == n and not isinstance(ds, torch.utils.data.DataLoader)


def test_synthetic_and_errors():
    tok = WordTokenizer()
    dl = loader.get_calibration_dataloader("synthetic", tok, num_samples=6, seq_len=32, batch_size=3, shuffle=False)
    b = next(iter(dl))
    assert b["input_ids"].shape == (3, 31) and int(b["input_ids"].max()) < tok.vocab_size
    with pytest.raises(NotImplementedError):
        loader.get_calibration_dataloader("boolq", tok, texts=None)
    with pytest.raises(ValueError):
        loader.get_calibration_dataloader("c4", tok, num_samples=1, seq_len=4096, texts=["too short"])
