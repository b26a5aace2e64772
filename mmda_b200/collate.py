"""Device-resident dataset + collate.

Host-side mirror of ``MSADataset`` / ``get_loader`` / ``collate_fn`` at reference
src/data_loader.py:18-131 for the caller side of the hot path.  The wire format is the
reference's pickled split: a list of ``((words, visual, acoustic, actual_words), label,
segment)`` tuples (src/create_dataset.py:388).  ``DeviceDataset`` uploads the whole split to HBM
once as ragged flat arrays; ``collate(indices)`` then assembles one batch on the device
(``mmda_collate_batch`` / ``mmda_collate_bert``), returning the same 10-tuple ``collate_fn``
returns -- already on the GPU, so the reference's ``to_gpu`` calls (src/utils/convert.py:4-11)
become no-ops -- with ``lengths`` on the CPU as ``Solver.train`` expects (src/solver.py:149).

There is no CPU fallback: without libmmda_b200.so / a GPU the calls raise ``MmdaError``.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Sequence

import numpy as np
import torch

from ._lib import LIB

PAD = 1                              # create_dataset.py:26-27
CLS, SEP, BERT_PAD = 101, 102, 0     # bert-base-uncased specials


def flatten_split(samples: Sequence, wordpiece_ids: Optional[Callable] = None):
    """Wire-format samples -> ragged flat numpy arrays (host side, once per split)."""
    if len(samples) == 0:
        raise ValueError("empty split")
    lens = np.array([np.asarray(s[0][0]).shape[0] for s in samples], dtype=np.int64)
    if (lens <= 0).any():
        raise ValueError("a sample has no words (pack_padded_sequence would reject it too)")
    offsets = np.zeros(len(samples) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    words = np.concatenate([np.asarray(s[0][0], dtype=np.int64) for s in samples])
    visual = np.concatenate([np.asarray(s[0][1], dtype=np.float32) for s in samples])
    acoustic = np.concatenate([np.asarray(s[0][2], dtype=np.float32) for s in samples])
    labels = np.stack([np.asarray(s[1], dtype=np.float32).reshape(-1) for s in samples])
    out = {"words": words, "visual": visual, "acoustic": acoustic, "labels": labels,
           "offsets": offsets, "lengths": lens, "segments": [s[2] for s in samples]}
    if wordpiece_ids is not None:
        wp = [np.asarray(list(wordpiece_ids(s)), dtype=np.int64) for s in samples]
        wpo = np.zeros(len(samples) + 1, dtype=np.int64)
        np.cumsum([len(w) for w in wp], out=wpo[1:])
        out["wp_ids"] = np.concatenate(wp) if wpo[-1] else np.zeros(1, dtype=np.int64)
        out["wp_offsets"] = wpo
    return out


class DeviceDataset:
    """One split resident in HBM.  ``len()``, ``visual_size`` / ``acoustic_size`` as
    ``MSADataset`` sets them on the config (data_loader.py:35-36)."""

    def __init__(self, samples: Sequence, device="cuda:0", wordpiece_ids: Optional[Callable] = None):
        flat = flatten_split(samples, wordpiece_ids)
        self.device = torch.device(device)
        self.lengths = flat["lengths"]                       # host copy: the sort key
        self.segments = flat["segments"]
        self.n = len(self.lengths)
        self.visual_size = int(flat["visual"].shape[1])
        self.acoustic_size = int(flat["acoustic"].shape[1])
        self.n_label = int(flat["labels"].shape[1])
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.words, self.visual, self.acoustic = up(flat["words"]), up(flat["visual"]), up(flat["acoustic"])
        self.labels, self.offsets = up(flat["labels"]), up(flat["offsets"])
        self.has_bert = "wp_ids" in flat
        if self.has_bert:
            self.wp_ids, self.wp_offsets = up(flat["wp_ids"]), up(flat["wp_offsets"])
        self.resident_bytes = sum(t.numel() * t.element_size() for t in
                                  (self.words, self.visual, self.acoustic, self.labels, self.offsets))

    def __len__(self):
        return self.n

    def sort_batch(self, indices: Iterable[int]) -> np.ndarray:
        """data_loader.py:64 -- stable descending sort by length (``sorted(..., reverse=True)``
        keeps the incoming order among equal lengths)."""
        idx = np.asarray(list(indices), dtype=np.int64)
        if idx.size == 0:
            raise ValueError("empty batch")
        if idx.min() < 0 or idx.max() >= self.n:
            raise IndexError("sample index out of range")
        return idx[np.argsort(-self.lengths[idx], kind="stable")]

    def collate(self, indices: Iterable[int]):
        """-> (sentences, visual, acoustic, labels, emo_labels, lengths[CPU], bert_sentences,
        bert_sentence_types, bert_sentence_att_mask, ids) exactly as ``collate_fn`` orders them."""
        if self.n_label != 7:
            raise TypeError("collate: emo_labels is None for label width != 7 (the reference's "
                            "torch.from_numpy(np.array(None)) raises, data_loader.py:109-118)")
        order = self.sort_batch(indices)
        B, T = int(order.size), int(self.lengths[order[0]])
        dev = self.device
        order_dev = torch.from_numpy(order).to(dev, non_blocking=True)
        sentences = torch.empty(T, B, dtype=torch.int64, device=dev)
        visual = torch.empty(T, B, self.visual_size, device=dev)
        acoustic = torch.empty(T, B, self.acoustic_size, device=dev)
        labels = torch.empty(B, device=dev)
        emo = torch.empty(B, 6, device=dev)
        lengths_dev = torch.empty(B, dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        LIB.call("mmda_collate_batch", self.words.data_ptr(), self.visual.data_ptr(),
                 self.acoustic.data_ptr(), self.labels.data_ptr(), self.offsets.data_ptr(),
                 order_dev.data_ptr(), B, T, self.visual_size, self.acoustic_size, self.n_label, PAD,
                 sentences.data_ptr(), visual.data_ptr(), acoustic.data_ptr(), labels.data_ptr(),
                 emo.data_ptr(), lengths_dev.data_ptr(), stream)
        W = T + 2
        if self.has_bert:
            ids = torch.empty(B, W, dtype=torch.int64, device=dev)
            types, mask = torch.empty_like(ids), torch.empty_like(ids)
            LIB.call("mmda_collate_bert", self.wp_ids.data_ptr(), self.wp_offsets.data_ptr(),
                     order_dev.data_ptr(), B, T, CLS, SEP, BERT_PAD, ids.data_ptr(),
                     types.data_ptr(), mask.data_ptr(), stream)
        else:
            ids = torch.zeros(B, W, dtype=torch.int64, device=dev)
            types, mask = torch.zeros_like(ids), torch.zeros_like(ids)
        lengths = torch.from_numpy(self.lengths[order].copy())        # CPU, solver.py:149
        return (sentences, visual, acoustic, labels, emo, lengths, ids, types, mask,
                [self.segments[i] for i in order])


class DeviceLoader:
    """``get_loader`` (data_loader.py:50-131): batches of ``batch_size`` indices, shuffled per
    epoch with a seeded generator when ``shuffle``; the last batch may be short."""

    def __init__(self, dataset: DeviceDataset, batch_size: int, shuffle: bool = True, seed: int = 0):
        self.dataset, self.batch_size, self.shuffle = dataset, int(batch_size), shuffle
        self.gen = torch.Generator().manual_seed(seed)

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        perm = torch.randperm(n, generator=self.gen).numpy() if self.shuffle else np.arange(n)
        for i in range(0, n, self.batch_size):
            yield self.dataset.collate(perm[i:i + self.batch_size])
