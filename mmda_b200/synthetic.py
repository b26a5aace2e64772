"""Seeded synthetic MOSI/MOSEI-shaped batches.

The datasets and GloVe are not available offline, so every measurement and parity run uses
tensors of the layout the reference's collate function emits (reference
``src/data_loader.py:59-122``): time-major padded ``sentences (T,B) int64`` (PAD id 1,
``create_dataset.py:25-27``), ``visual (T,B,d_v)`` / ``acoustic (T,B,d_a)`` float32 zero past
each length, ``labels (B,6)`` float multi-hot, ``lengths (B,) int64`` on the CPU sorted
descending, and batch-first BERT id / type / mask tensors of width T+2.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

PAD_ID = 1


@dataclass
class Batch:
    sentences: torch.Tensor      # (T,B) int64
    visual: torch.Tensor         # (T,B,d_v) f32
    acoustic: torch.Tensor       # (T,B,d_a) f32
    labels: torch.Tensor         # (B,6) f32 multi-hot (emo_label in the reference)
    lengths: torch.Tensor        # (B,) int64, CPU
    bert_sent: torch.Tensor      # (B,T+2) int64
    bert_sent_type: torch.Tensor
    bert_sent_mask: torch.Tensor

    def model_args(self):
        return (self.sentences, self.visual, self.acoustic, self.lengths,
                self.bert_sent, self.bert_sent_type, self.bert_sent_mask)

    def slice(self, lo: int, hi: int) -> "Batch":
        """Batch shard [lo, hi) along the sample axis (data-parallel partition)."""
        return Batch(self.sentences[:, lo:hi].contiguous(), self.visual[:, lo:hi].contiguous(),
                     self.acoustic[:, lo:hi].contiguous(), self.labels[lo:hi].contiguous(),
                     self.lengths[lo:hi].contiguous(), self.bert_sent[lo:hi].contiguous(),
                     self.bert_sent_type[lo:hi].contiguous(), self.bert_sent_mask[lo:hi].contiguous())


def make_lengths(batch: int, seq_len: int, mode: str, gen: torch.Generator) -> torch.Tensor:
    if mode == "full":
        return torch.full((batch,), seq_len, dtype=torch.int64)
    if mode == "ragged":
        ln = torch.randint(1, seq_len + 1, (batch,), generator=gen, dtype=torch.int64)
        ln, _ = torch.sort(ln, descending=True)
        ln[0] = seq_len
        return ln
    if mode == "shuffled":   # not length-sorted: exercises enforce_sorted=False
        ln = torch.randint(1, seq_len + 1, (batch,), generator=gen, dtype=torch.int64)
        ln[int(torch.randint(0, batch, (1,), generator=gen))] = seq_len
        return ln
    raise ValueError(mode)


def make_batch(batch: int, seq_len: int, d_text_vocab: int, d_visual: int, d_acoustic: int,
               seed: int = 1234, lengths: str = "full", num_classes: int = 6) -> Batch:
    g = torch.Generator().manual_seed(seed)
    ln = make_lengths(batch, seq_len, lengths, g)
    T = int(ln.max())
    t_idx = torch.arange(T).unsqueeze(1)                       # (T,1)
    valid = t_idx < ln.unsqueeze(0)                            # (T,B)
    sent = torch.randint(2, d_text_vocab, (T, batch), generator=g, dtype=torch.int64)
    sent = torch.where(valid, sent, torch.full_like(sent, PAD_ID))
    vis = torch.randn(T, batch, d_visual, generator=g) * valid.unsqueeze(-1)
    aco = torch.randn(T, batch, d_acoustic, generator=g) * valid.unsqueeze(-1)
    while True:
        y = (torch.rand(batch, num_classes, generator=g) < 0.3).float()
        if bool((y.sum(0) > 0).all()):
            break
    bert = torch.randint(1000, 30522, (batch, T + 2), generator=g, dtype=torch.int64)
    bmask = (torch.arange(T + 2).unsqueeze(0) < (ln + 2).unsqueeze(1)).long()
    bert = bert * bmask
    btype = torch.zeros_like(bert)
    return Batch(sent, vis, aco, y, ln, bert, btype, bmask)


def batch_for(cfg, seed: int = 1234, lengths: str = "full", batch: int | None = None,
              seq_len: int = 50) -> Batch:
    return make_batch(batch or cfg.batch_size, seq_len, len(cfg.word2id), cfg.visual_size,
                      cfg.acoustic_size, seed=seed, lengths=lengths, num_classes=cfg.num_classes)
