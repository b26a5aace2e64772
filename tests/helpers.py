"""Shared helpers for the parity tests (CPU and GPU)."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_small(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def small_cfg(meta, **kw):
    from mmda_b200.config import MisaConfig
    c = dict(meta["cfg"])
    c.pop("seq_len")
    return MisaConfig(use_confidNet=meta["use_confidNet"], use_cmd_sim=meta.get("use_cmd_sim", True),
                      rnncell=meta.get("rnncell", "lstm"), **c, **kw)


def small_batch(z):
    from mmda_b200.synthetic import Batch
    T, B = z["in/sentences"].shape
    ln = torch.from_numpy(z["in/lengths"])
    bert = torch.zeros(B, T + 2, dtype=torch.int64)
    return Batch(torch.from_numpy(z["in/sentences"]), torch.from_numpy(z["in/visual"]),
                 torch.from_numpy(z["in/acoustic"]), torch.from_numpy(z["in/labels"]), ln,
                 bert, bert.clone(), bert.clone())


def state_from_npz(z, prefix="param/"):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).flatten()
    b = torch.as_tensor(b, dtype=torch.float64).flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_rel(a, b):
    """max |a-b| / max|b|  (scale-relative, the tolerance BASELINE.json states)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
