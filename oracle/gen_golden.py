"""Generate tests/golden/* by running the UNMODIFIED reference (/root/reference/src) on CPU.
TEST INFRASTRUCTURE.  Runs only in the build container (the GPU box has no /root/reference);
the fixtures it writes are committed and are what the tests read.

    python oracle/gen_golden.py            # rewrites tests/golden/
    python oracle/gen_golden.py --collate-only   # only tests/golden/collate_small.npz

Recipe: put the reference on sys.path, sanitise argv, stub the packages
that are absent offline (gensim, hypertune, mmsdk, wandb), neutralise the import-time
``BertTokenizer.from_pretrained`` network call (solver.py:39), build through ``Solver.build``
with ``is_train=False`` (torch 2.11 rejects ``ReduceLROnPlateau(verbose=...)``, solver.py:100),
then create the criteria and the optimizer exactly as solver.py:97-99,108-118 do and call the
reference's own ``get_*_loss`` / ``backward`` / ``clip_grad_value_`` / ``Adam.step``.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
sys.path.insert(0, ROOT)


def import_reference():
    sys.argv = ["gen_golden"]
    sys.path.insert(0, REF)
    for name in ("gensim", "hypertune", "mmsdk", "mmsdk.mmdatasdk", "wandb"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mmsdk"].mmdatasdk = sys.modules["mmsdk.mmdatasdk"]
    sys.modules["hypertune"].HyperTune = lambda: None
    import transformers
    transformers.BertTokenizer.from_pretrained = classmethod(lambda cls, *a, **k: None)
    import warnings
    warnings.simplefilter("ignore")
    import models as ref_models          # noqa: F401
    import solver as ref_solver
    import config as ref_config
    return ref_solver, ref_config


def build_reference(ref_solver, ref_config, cfg, seed):
    import torch
    import torch.nn as nn
    rc = ref_config.get_config(parse=False, use_bert=False, data=cfg.data,
                               use_confidNet=cfg.use_confidNet, batch_size=cfg.batch_size,
                               embedding_size=cfg.embedding_size, hidden_size=cfg.hidden_size,
                               dropout=cfg.dropout, learning_rate=cfg.learning_rate,
                               use_cmd_sim=cfg.use_cmd_sim, rnncell=cfg.rnncell)
    rc.visual_size, rc.acoustic_size = cfg.visual_size, cfg.acoustic_size
    rc.word2id = {i: i for i in range(cfg.vocab_size)}
    rc.pretrained_emb = None
    torch.manual_seed(seed)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        s = ref_solver.Solver(rc, None, None, None, None, None, is_train=False)
        s.build()
    s.optimizer = rc.optimizer(filter(lambda p: p.requires_grad, s.model.parameters()),
                               lr=rc.learning_rate)
    s.criterion = nn.BCELoss(reduction="mean")
    s.domain_loss_criterion = nn.CrossEntropyLoss(reduction="mean")
    s.loss_diff = ref_solver.DiffLoss()
    s.loss_recon = nn.MSELoss(reduction="mean")
    s.loss_cmd = ref_solver.CMD()
    s.loss_mcp = nn.CrossEntropyLoss(reduction="mean")
    s.loss_tcp = nn.MSELoss(reduction="mean")
    return s, rc


def reference_step(s, rc, batch, do_step=True):
    """solver.py:139-186 with the reference's own functions."""
    import torch
    m = s.model
    m.zero_grad()
    scores, labels = m(*batch.model_args())
    y = batch.labels.type(torch.float)
    L = {"cls": s.get_cls_loss(scores, y), "diff": s.get_diff_loss(),
         "recon": s.get_recon_loss(),
         "sim": s.get_cmd_loss() if rc.use_cmd_sim else s.get_domain_loss(),   # solver.py:170-173
         "conf": s.get_conf_loss(scores, y)}
    loss = L["cls"] + rc.diff_weight * L["diff"] + rc.sim_weight * L["sim"] + \
        rc.recon_weight * L["recon"]
    if rc.use_confidNet:
        loss = loss + rc.conf_weight * L["conf"]
    L["total"] = loss
    loss.backward()
    grads = {n: (None if p.grad is None else p.grad.detach().clone())
             for n, p in m.named_parameters()}
    if do_step:
        torch.nn.utils.clip_grad_value_([p for p in m.parameters() if p.requires_grad], rc.clip)
        s.optimizer.step()
    return scores, labels, L, grads


ATTRS = ["utt_t_orig", "utt_v_orig", "utt_a_orig", "utt_private_t", "utt_private_v",
         "utt_private_a", "utt_shared_t", "utt_shared_v", "utt_shared_a", "utt_t_recon",
         "utt_v_recon", "utt_a_recon", "tcp", "shared_or_private_p_t", "shared_or_private_p_v",
         "shared_or_private_p_a", "shared_or_private_s"]


def gen_small(ref_solver, ref_config, name, seed, confid, lengths_mode, use_cmd_sim=True,
              rnncell="lstm"):
    import torch
    from mmda_b200.config import MisaConfig
    from mmda_b200.synthetic import batch_for
    cfg = MisaConfig(embedding_size=12, visual_size=5, acoustic_size=7, hidden_size=16,
                     vocab_size=50, batch_size=6, use_confidNet=confid, dropout=0.1,
                     use_cmd_sim=use_cmd_sim, rnncell=rnncell)
    s, rc = build_reference(ref_solver, ref_config, cfg, seed)
    s.model.eval()                                   # deterministic parity mode
    batch = batch_for(cfg, seed=seed + 1, lengths=lengths_mode, seq_len=7)
    params0 = {n: p.detach().clone() for n, p in s.model.named_parameters()}
    scores, labels, L, grads = reference_step(s, rc, batch, do_step=True)
    packed = torch.nn.utils.rnn.pack_padded_sequence(batch.visual, batch.lengths,
                                                     enforce_sorted=False)
    arrs = {}
    for n, p in params0.items():
        arrs["param/" + n] = p.numpy()
    for n, g in grads.items():
        if g is not None:
            arrs["grad/" + n] = g.numpy()
    for n, p in s.model.named_parameters():
        arrs["after/" + n] = p.detach().numpy()
    for a in ATTRS + ([] if use_cmd_sim else ["domain_label_t", "domain_label_v", "domain_label_a"]):
        arrs["out/" + a] = getattr(s.model, a).detach().numpy()
    arrs["out/scores"] = scores.detach().numpy()
    arrs["out/labels"] = labels.detach().numpy()
    for k, v in L.items():
        arrs["loss/" + k] = np.asarray(float(v), dtype=np.float64)
    for f in ("sentences", "visual", "acoustic", "labels", "lengths"):
        arrs["in/" + f] = getattr(batch, f).numpy()
    arrs["pack/batch_sizes"] = packed.batch_sizes.numpy()
    arrs["pack/sorted_indices"] = packed.sorted_indices.numpy()
    arrs["pack/unsorted_indices"] = packed.unsorted_indices.numpy()
    arrs["pack/data"] = packed.data.numpy()
    meta = {"seed": seed, "use_confidNet": confid, "lengths": lengths_mode, "use_cmd_sim": use_cmd_sim,
            "rnncell": rnncell,
            "none_grads": sorted(n for n, g in grads.items() if g is None),
            "cfg": {"embedding_size": 12, "visual_size": 5, "acoustic_size": 7,
                    "hidden_size": 16, "vocab_size": 50, "batch_size": 6, "seq_len": 7}}
    arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"),
                        **{k: (v.astype(np.float32) if v.dtype == np.float64 and not k.startswith("loss/") else v)
                           for k, v in arrs.items()})
    print(name, {k: float(v) for k, v in L.items()})


def stat(t):
    t = t.detach().double()
    return [float(t.sum()), float(t.abs().sum()), float(t.norm())]


def gen_summary(ref_solver, ref_config, name, cfg, seed, lengths_mode, steps=2):
    """Full-size configs: keep seeds + scalar summaries only (weights regenerate from the seed)."""
    from mmda_b200.synthetic import batch_for
    s, rc = build_reference(ref_solver, ref_config, cfg, seed)
    s.model.eval()
    rec = {"seed": seed, "lengths": lengths_mode, "batch_seed": seed + 1, "steps": []}
    rec["param0"] = {n: stat(p) for n, p in s.model.named_parameters()}
    for it in range(steps):
        batch = batch_for(cfg, seed=seed + 1 + it, lengths=lengths_mode)
        scores, labels, L, grads = reference_step(s, rc, batch, do_step=True)
        rec["steps"].append({
            "losses": {k: float(v) for k, v in L.items()},
            "scores": scores.detach().double().numpy().round(9).tolist(),
            "labels_sum": float(labels.sum()),
            "grads": {n: (None if g is None else stat(g)) for n, g in grads.items()},
            "params_after": {n: stat(p) for n, p in s.model.named_parameters()},
        })
        print(name, it, rec["steps"][-1]["losses"])
    with open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w") as f:
        json.dump(rec, f)


class _FakeTokenizer:
    """Stand-in for bert-base-uncased's tokenizer (no network / vocab file offline): word pieces
    from oracle.collate_oracle.wordpieces, then the encode_plus contract the reference relies on
    (data_loader.py:84-85): specials added, truncated to max_length, padded to max_length."""

    def encode_plus(self, text, max_length=None, add_special_tokens=True, pad_to_max_length=False):
        from oracle.collate_oracle import BERT_PAD, CLS, SEP, wordpieces
        wp = wordpieces(text.split(" "))[:max_length - 2]
        ids = [CLS] + wp + [SEP]
        n = len(ids)
        ids = ids + [BERT_PAD] * (max_length - n)
        return {"input_ids": ids, "token_type_ids": [0] * max_length,
                "attention_mask": [1] * n + [0] * (max_length - n)}


def gen_collate():
    """Run the reference's own collate_fn (data_loader.py:59-122) on seeded ragged samples."""
    import types as _types
    import data_loader as ref_dl
    from oracle.collate_oracle import make_samples
    samples = make_samples(11, 5, 7, seed=77)

    class FakeDataset:
        def __init__(self, config):
            self.data = samples
        def __len__(self):
            return len(self.data)

    ref_dl.MSADataset = FakeDataset
    ref_dl.DataLoader = lambda dataset, batch_size, shuffle, collate_fn: _types.SimpleNamespace(
        collate_fn=collate_fn)
    ref_dl.bert_tokenizer = _FakeTokenizer()
    cfg = _types.SimpleNamespace(mode="train", batch_size=4)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        loader = ref_dl.get_loader(cfg, shuffle=False)
    arrs = {}
    for bi, idx in enumerate([[0, 1, 2, 3, 4, 5, 6], [10, 3, 8, 7, 9, 2], [4]]):
        out = loader.collate_fn([samples[i] for i in idx])
        names = ["sentences", "visual", "acoustic", "labels", "emo_labels", "lengths",
                 "bert_sentences", "bert_sentence_types", "bert_sentence_att_mask"]
        arrs[f"b{bi}/index"] = np.array(idx, dtype=np.int64)
        for n, t in zip(names, out[:9]):
            arrs[f"b{bi}/{n}"] = t.numpy()
        arrs[f"b{bi}/ids"] = np.array(out[9])
    arrs["meta"] = np.frombuffer(json.dumps({"n": 11, "dv": 5, "da": 7, "seed": 77}).encode(),
                                 dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "collate_small.npz"), **arrs)
    print("collate_small", {k: v.shape for k, v in arrs.items() if k.startswith("b1/")})


def main():
    adv_only = "--adversarial-only" in sys.argv      # import_reference() sanitises argv
    collate_only = "--collate-only" in sys.argv
    gru_only = "--gru-only" in sys.argv
    ref_solver, ref_config = import_reference()
    if collate_only:
        gen_collate()
        return
    if gru_only:
        gen_small(ref_solver, ref_config, "small_gru", 41, True, "shuffled", rnncell="gru")
        return
    from mmda_b200.config import mosi_config, mosei_config
    if adv_only:
        gen_small(ref_solver, ref_config, "small_adversarial", 31, False, "shuffled", use_cmd_sim=False)
        return
    gen_small(ref_solver, ref_config, "small_adversarial", 31, False, "shuffled", use_cmd_sim=False)
    gen_small(ref_solver, ref_config, "small_gru", 41, True, "shuffled", rnncell="gru")
    gen_small(ref_solver, ref_config, "small_ragged", 11, False, "ragged")
    gen_small(ref_solver, ref_config, "small_shuffled_confid", 23, True, "shuffled")
    gen_summary(ref_solver, ref_config, "c1_mosi_b64", mosi_config(vocab_size=2000), 1234, "ragged")
    gen_summary(ref_solver, ref_config, "c2_mosei_b256", mosei_config(vocab_size=2000), 1234,
                "full", steps=1)
    gen_summary(ref_solver, ref_config, "c3_mosei_confid_b256",
                mosei_config(vocab_size=2000, use_confidNet=True), 4321, "ragged", steps=1)
    gen_collate()


if __name__ == "__main__":
    main()
