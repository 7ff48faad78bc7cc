"""Causal-LM datasets for the benchmarks (counterpart of reference benchmarks/dataset.py).

* ``SyntheticLM``  : uniform random tokens (throughput runs; no files needed).
* ``MarkovLM``     : tokens from a fixed random bigram chain -- learnable, so loss curves are meaningful and two
                     implementations can be compared step by step (accuracy benchmark) without downloading a corpus.
* ``TextFileLM``   : a local UTF-8 text file, byte-level tokens (vocab 256+), chunked into fixed-length samples
                     (drop a wikitext-2 ``train.txt`` next to the script to reproduce the reference's data set-up).
"""
from __future__ import annotations

import os
from typing import Dict, Iterator

import torch
from torch.utils.data import Dataset


class SyntheticLM(Dataset):

    def __init__(self, vocab_size: int, seq_len: int, num_samples: int = 4096, seed: int = 0):
        g = torch.Generator().manual_seed(seed)
        self.data = torch.randint(0, vocab_size, (num_samples, seq_len), generator=g)

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i) -> Dict[str, torch.Tensor]:
        x = self.data[i]
        return {"input_ids": x, "labels": x}


class MarkovLM(Dataset):
    """First-order Markov chain over ``vocab_size`` tokens with ``branching`` successors per token: the optimal
    loss is ``log(branching)``, a model that learns bigrams approaches it within a few hundred steps."""

    def __init__(self, vocab_size: int, seq_len: int, num_samples: int = 4096, branching: int = 4, seed: int = 0):
        g = torch.Generator().manual_seed(seed)
        succ = torch.randint(0, vocab_size, (vocab_size, branching), generator=g)
        tok = torch.randint(0, vocab_size, (num_samples,), generator=g)
        out = torch.empty(num_samples, seq_len, dtype=torch.long)
        for t in range(seq_len):
            out[:, t] = tok
            pick = torch.randint(0, branching, (num_samples,), generator=g)
            tok = succ[tok, pick]
        self.data = out
        self.optimal_loss = float(torch.log(torch.tensor(float(branching))))

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        x = self.data[i]
        return {"input_ids": x, "labels": x}


class TextFileLM(Dataset):

    def __init__(self, path: str, seq_len: int):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: no local corpus (this sandbox has no network; use MarkovLM instead)")
        raw = torch.frombuffer(bytearray(open(path, "rb").read()), dtype=torch.uint8).long()
        n = raw.numel() // seq_len
        self.data = raw[:n * seq_len].view(n, seq_len)
        self.vocab_size = 256

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        x = self.data[i]
        return {"input_ids": x, "labels": x}


def batches(ds: Dataset, batch_size: int, rank: int = 0, world: int = 1, seed: int = 0) -> Iterator[Dict[str, torch.Tensor]]:
    """Infinite, rank-sharded, deterministic batch stream."""
    g = torch.Generator().manual_seed(seed)
    n = len(ds)
    while True:
        perm = torch.randperm(n, generator=g)
        per = batch_size * world
        for s in range(0, n - per + 1, per):
            idx = perm[s + rank * batch_size: s + (rank + 1) * batch_size]
            items = [ds[int(i)] for i in idx]
            yield {k: torch.stack([it[k] for it in items]) for k in items[0]}
