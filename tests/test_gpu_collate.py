"""Device-resident collate through the C ABI vs the reference's collate_fn
(golden fixture) and vs the oracle restatement at MOSEI widths.  Bit-exact: the collate only moves
bytes, compares labels with 0 and replaces NaN."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu

KEYS = ("sentences", "visual", "acoustic", "labels", "emo_labels", "lengths", "bert_sentences",
        "bert_sentence_types", "bert_sentence_att_mask")


def _same(got, ref, tag):
    got = got.cpu().numpy()
    assert got.shape == ref.shape and got.dtype == ref.dtype, (tag, got.shape, ref.shape, got.dtype, ref.dtype)
    assert np.array_equal(got, ref, equal_nan=True), tag


def test_collate_matches_reference_golden():
    from mmda_b200.collate import DeviceDataset
    from oracle.collate_oracle import make_samples, wordpieces
    z = np.load(os.path.join(GOLDEN, "collate_small.npz"), allow_pickle=False)
    meta = json.loads(bytes(z["meta"]).decode())
    samples = make_samples(meta["n"], meta["dv"], meta["da"], seed=meta["seed"])
    ds = DeviceDataset(samples, "cuda:0", wordpiece_ids=lambda s: wordpieces(s[0][3]))
    for bi in range(3):
        out = ds.collate(z[f"b{bi}/index"])
        for k, t in zip(KEYS, out[:9]):
            _same(t, z[f"b{bi}/{k}"], (bi, k))
        assert out[5].device.type == "cpu"                    # lengths stay on the host
        assert out[9] == list(z[f"b{bi}/ids"])


def test_collate_mosei_widths_vs_oracle_and_feeds_the_step():
    from mmda_b200 import MISA, FusedTrainer
    from mmda_b200.collate import DeviceDataset, DeviceLoader
    from mmda_b200.config import mosei_config
    from oracle.collate_oracle import collate, make_samples, wordpieces
    samples = make_samples(300, 35, 74, seed=9, max_len=50, vocab=2000)
    ds = DeviceDataset(samples, "cuda:0", wordpiece_ids=lambda s: wordpieces(s[0][3]))
    g = np.random.RandomState(3)
    for B in (1, 64, 256):
        idx = g.permutation(300)[:B]
        out = ds.collate(idx)
        ref = collate([samples[i] for i in idx], wp_ids=lambda s: wordpieces(s[0][3]))
        for k, t in zip(KEYS, out[:9]):
            _same(t, ref[k], (B, k))
        assert out[9] == ref["ids"]
    # the collated batch drives the fused step directly (no host copies of the tensors)
    clean = make_samples(128, 35, 74, seed=10, max_len=20, vocab=2000, with_nan=False)
    for s in clean:
        s[1][0, 1:] = np.abs(s[1][0, 1:]) + 0.1 * (np.arange(6) % 2 == 0)   # every class has positives
    dl = DeviceLoader(DeviceDataset(clean, "cuda:0"), batch_size=64, shuffle=True, seed=1)
    cfg = mosei_config(vocab_size=2000, batch_size=64)
    torch.manual_seed(0)
    tr = FusedTrainer(MISA(cfg).to("cuda:0").train())
    n = 0
    for t, v, a, y, emo, l, *_ in dl:
        L = tr.step(t, v, a, l, emo)
        assert torch.isfinite(L[:6]).all()
        n += 1
    assert n == len(dl) == 2
