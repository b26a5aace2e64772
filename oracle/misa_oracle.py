"""CPU oracle for the MISA hot path.  TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` -- its ``cpu_baseline`` / ``--impl
reference`` legs (the timed CPU baseline) and, at N > 1, the ``dp_check`` parity evidence outside any
timed region (the checker of the data-parallel step) -- may import this module.  The
product path (``mmda_b200``) never routes through it and fails loudly without its CUDA library.

What it restates (all citations are into /root/reference):

* the model forward, ``src/models.py:163-285`` (``extract_features``, ``alignment``,
  ``shared_private``, ``reconstruct``) -- as plain PyTorch CPU modules, which is the arithmetic
  the reference itself executes (it has no kernels of its own; SURVEY.md section 8c row O2: torch
  2.11.0 and transformers 5.5.0 in this image, unpinned upstream);
* the loss functions, ``src/solver.py:373-462`` and ``src/utils/functions.py:49-115``;
* the optimisation step, ``src/solver.py:175-186`` (weighted sum, backward, clip_grad_value_, Adam).

Pinning: the reference has no tests or golden vectors, so this oracle is pinned
against outputs of the reference itself, run in the build container by
``oracle/gen_golden.py`` (which imports /root/reference/src) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks this file against those fixtures
on every CPU test run.  ``oracle/explicit.py`` separately restates the third-party pieces
(LSTM cell, packing, LayerNorm, attention) from their published definitions.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

_ACT = {"leakyrelu": nn.LeakyReLU, "relu": nn.ReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}
_MODS = ("t", "v", "a")


def _act_module(act):
    if isinstance(act, str):
        return _ACT[act.lower()]()
    return act() if isinstance(act, type) else act


def _seq(**named):
    s = nn.Sequential()
    for k, v in named.items():
        s.add_module(k, v)
    return s


class OracleMISA(nn.Module):
    """Same parameter names, registration order and init RNG consumption as
    ``models.MISA.__init__`` (models.py:17-161), so that the same ``torch.manual_seed`` yields the
    same weights and reference ``state_dict``s load unchanged."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        d = cfg.hidden_size
        sizes = {"t": cfg.embedding_size, "v": cfg.visual_size, "a": cfg.acoustic_size}
        self.sizes = sizes
        act = _act_module(cfg.activation)
        if cfg.extractor != "lstm":
            raise NotImplementedError("extractor='transformer' exits in the reference (models.py:33-36)")
        rnn = nn.LSTM if cfg.rnncell == "lstm" else nn.GRU          # models.py:39
        if cfg.use_bert:
            from transformers import BertConfig, BertModel
            # bert-base-uncased geometry == BertConfig() defaults; no network here.
            self.bertmodel = BertModel(BertConfig(output_hidden_states=True))
        else:
            self.embed = nn.Embedding(len(cfg.word2id), sizes["t"])
            self.trnn1 = rnn(sizes["t"], sizes["t"], bidirectional=True)
            self.trnn2 = rnn(2 * sizes["t"], sizes["t"], bidirectional=True)
        for m in ("v", "a"):
            setattr(self, f"{m}rnn1", rnn(sizes[m], sizes[m], bidirectional=True))
            setattr(self, f"{m}rnn2", rnn(2 * sizes[m], sizes[m], bidirectional=True))
        for m in _MODS:
            fan_in = 768 if (m == "t" and cfg.use_bert) else 4 * sizes[m]
            setattr(self, f"project_{m}", _seq(**{
                f"project_{m}": nn.Linear(fan_in, d),
                f"project_{m}_activation": act,
                f"project_{m}_layer_norm": nn.LayerNorm(d)}))
        for m, tag in zip(_MODS, ("1", "1", "3")):       # private_a uses suffix _3 (models.py:95)
            setattr(self, f"private_{m}", _seq(**{
                f"private_{m}_{tag}": nn.Linear(d, d),
                f"private_{m}_activation_{tag}": nn.Sigmoid()}))
        self.shared = _seq(shared_1=nn.Linear(d, d), shared_activation_1=nn.Sigmoid())
        for m in _MODS:
            setattr(self, f"recon_{m}", _seq(**{f"recon_{m}_1": nn.Linear(d, d)}))
        if not cfg.use_cmd_sim:
            self.discriminator = _seq(
                discriminator_layer_1=nn.Linear(d, d),
                discriminator_layer_1_activation=act,
                discriminator_layer_1_dropout=nn.Dropout(cfg.dropout),
                discriminator_layer_2=nn.Linear(d, 3))
        self.sp_discriminator = _seq(sp_discriminator_layer_1=nn.Linear(d, 4))
        self.confidence = _seq(confidence_layer_1=nn.Linear(6 * d, 6),
                               confidence_layer_activation=nn.Sigmoid())
        self.classifier = _seq(classifier_layer=nn.Linear(6 * d, cfg.num_classes),
                               classifier_layer_dropout=nn.Dropout(cfg.dropout),
                               classifier_layer_activation=nn.Sigmoid())
        self.tlayer_norm = nn.LayerNorm((2 * sizes["t"],))
        self.vlayer_norm = nn.LayerNorm((2 * sizes["v"],))
        self.alayer_norm = nn.LayerNorm((2 * sizes["a"],))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.transformer_encoder = nn.TransformerEncoder(
                nn.TransformerEncoderLayer(d_model=d, nhead=2), num_layers=1)

    # -- models.py:163-180 -------------------------------------------------------------------
    @staticmethod
    def encode(seq, lengths, rnn1, rnn2, norm):
        B = lengths.size(0)
        out1, h1 = rnn1(pack_padded_sequence(seq, lengths, enforce_sorted=False))
        h1 = h1[0] if isinstance(h1, tuple) else h1             # LSTM: (h, c); GRU: h  (models.py:166-169)
        padded, _ = pad_packed_sequence(out1)
        _, h2 = rnn2(pack_padded_sequence(norm(padded), lengths, enforce_sorted=False))
        h2 = h2[0] if isinstance(h2, tuple) else h2
        # models.py:203 -> per sample [h1_fwd | h2_fwd | h1_bwd | h2_bwd]
        return torch.cat((h1, h2), dim=2).permute(1, 0, 2).contiguous().view(B, -1)

    # -- models.py:182-285 -------------------------------------------------------------------
    def forward(self, sentences, visual, acoustic, lengths, bert_sent=None, bert_sent_type=None,
                bert_sent_mask=None) -> Dict[str, torch.Tensor]:
        cfg, out = self.cfg, {}
        if cfg.use_bert:
            hid = self.bertmodel(input_ids=bert_sent, attention_mask=bert_sent_mask,
                                 token_type_ids=bert_sent_type)[0]
            m = bert_sent_mask.unsqueeze(2)
            utt = {"t": (m * hid).sum(1) / bert_sent_mask.sum(1, keepdim=True)}
        else:
            utt = {"t": self.encode(self.embed(sentences), lengths, self.trnn1, self.trnn2,
                                    self.tlayer_norm)}
        utt["v"] = self.encode(visual, lengths, self.vrnn1, self.vrnn2, self.vlayer_norm)
        utt["a"] = self.encode(acoustic, lengths, self.arnn1, self.arnn2, self.alayer_norm)
        for m in _MODS:
            out[f"utterance_{m}"] = utt[m]
            o = getattr(self, f"project_{m}")(utt[m])
            out[f"utt_{m}_orig"] = o
            out[f"utt_private_{m}"] = getattr(self, f"private_{m}")(o)
            out[f"utt_shared_{m}"] = self.shared(o)
        if not cfg.use_cmd_sim:
            for m in _MODS:
                out[f"domain_label_{m}"] = self.discriminator(
                    _GradReverse.apply(out[f"utt_shared_{m}"], cfg.reverse_grad_weight))
        for m in _MODS:
            out[f"shared_or_private_p_{m}"] = self.sp_discriminator(out[f"utt_private_{m}"])
        out["shared_or_private_s"] = self.sp_discriminator(
            (out["utt_shared_t"] + out["utt_shared_v"] + out["utt_shared_a"]) / 3.0)
        for m in _MODS:
            out[f"utt_{m}"] = out[f"utt_private_{m}"] + out[f"utt_shared_{m}"]
            out[f"utt_{m}_recon"] = getattr(self, f"recon_{m}")(out[f"utt_{m}"])
        tokens = torch.stack([out[f"utt_private_{m}"] for m in _MODS] +
                             [out[f"utt_shared_{m}"] for m in _MODS], dim=0)      # (6,B,d)
        fused = self.transformer_encoder(tokens)
        h = torch.cat([fused[i] for i in range(6)], dim=1)                        # (B,6d)
        out["fused"] = h
        out["tcp"] = self.confidence(h)
        out["scores"] = self.classifier(h)
        out["labels"] = (out["scores"] > cfg.threshold).to(out["scores"].dtype)   # functions.py:112-115
        return out


class _GradReverse(torch.autograd.Function):          # functions.py:9-21
    @staticmethod
    def forward(ctx, x, p):
        ctx.p = p
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.neg() * ctx.p, None


# ------------------------------------------------------------------------------------------
# losses: solver.py:373-462, functions.py:49-109
# ------------------------------------------------------------------------------------------
def diff_loss(a, b):                                   # functions.py:54-78
    a = torch.nan_to_num(a.reshape(a.size(0), -1))
    b = torch.nan_to_num(b.reshape(b.size(0), -1))
    a = a - a.mean(0, keepdim=True)
    b = b - b.mean(0, keepdim=True)
    a = a / (a.norm(p=2, dim=1, keepdim=True).detach() + 1e-6)
    b = b / (b.norm(p=2, dim=1, keepdim=True).detach() + 1e-6)
    return (a.t() @ b).pow(2).mean()


def cmd_loss(x1, x2, n_moments=5):                     # functions.py:88-109
    m1, m2 = x1.mean(0), x2.mean(0)
    c1, c2 = x1 - m1, x2 - m2
    total = ((m1 - m2) ** 2).sum() ** 0.5
    for k in range(2, n_moments + 1):
        total = total + ((c1.pow(k).mean(0) - c2.pow(k).mean(0)) ** 2).sum() ** 0.5
    return total


def oracle_losses(out, y, cfg) -> Dict[str, torch.Tensor]:
    s, P, S = out["scores"], [out[f"utt_private_{m}"] for m in _MODS], \
        [out[f"utt_shared_{m}"] for m in _MODS]
    y = y.to(s.dtype)
    L = {}
    # solver.py:373-385  (sum over classes of per-class mean BCE)
    L["cls"] = sum(F.binary_cross_entropy(s[:, c], y[:, c]) for c in range(y.size(1)))
    # solver.py:422-441  pairs: (p_t,s_t) (p_v,s_v) (p_a,s_a) (p_a,p_t) (p_a,p_v) (p_t,p_v)
    L["diff"] = (diff_loss(P[0], S[0]) + diff_loss(P[1], S[1]) + diff_loss(P[2], S[2]) +
                 diff_loss(P[2], P[0]) + diff_loss(P[2], P[1]) + diff_loss(P[0], P[1]))
    # solver.py:443-449
    L["recon"] = sum(F.mse_loss(out[f"utt_{m}_recon"], out[f"utt_{m}_orig"]) for m in _MODS) / 3.0
    # solver.py:409-420  pairs: (s_t,s_v) (s_t,s_a) (s_a,s_v)
    if cfg.use_cmd_sim:
        L["sim"] = (cmd_loss(S[0], S[1]) + cmd_loss(S[0], S[2]) + cmd_loss(S[2], S[1])) / 3.0
    else:                                               # solver.py:388-407
        pred = torch.cat([out[f"domain_label_{m}"] for m in _MODS], 0)
        n = s.size(0)
        true = torch.cat([torch.full((n,), i, dtype=torch.long, device=pred.device) for i in range(3)], 0)
        L["sim"] = F.cross_entropy(pred, true)
    # solver.py:451-462 ; CrossEntropyLoss on 1-D float input+target == soft-label CE over batch axis
    tcp = out["tcp"]
    conf = 0.0
    for c in range(y.size(1)):
        nnz = torch.count_nonzero(y[:, c])
        conf = conf + F.mse_loss(tcp[:, c], y[:, c] * s[:, c]) / nnz
        conf = conf + (-(y[:, c] * F.log_softmax(s[:, c], dim=0)).sum()) / nnz
    L["conf"] = conf
    total = L["cls"] + cfg.diff_weight * L["diff"] + cfg.sim_weight * L["sim"] + \
        cfg.recon_weight * L["recon"]
    if cfg.use_confidNet:
        total = total + cfg.conf_weight * L["conf"]
    L["total"] = total
    return L


def oracle_build(cfg, seed: int, dtype=torch.float32) -> OracleMISA:
    """``Solver.build`` (solver.py:60-94): construct, orthogonal-init every ``weight_hh``."""
    torch.manual_seed(seed)
    model = OracleMISA(cfg)
    for name, p in model.named_parameters():
        if cfg.data == "mosei" and "bertmodel.encoder.layer" in name:
            if int(name.split("encoder.layer.")[-1].split(".")[0]) <= 8:
                p.requires_grad = False
        if "weight_hh" in name:
            nn.init.orthogonal_(p)
    return model.to(dtype)


def oracle_optimizer(model, cfg):
    return torch.optim.Adam([p for p in model.parameters() if p.requires_grad],
                            lr=cfg.learning_rate)          # solver.py:97-99 (no weight decay)


def oracle_step(model, batch, cfg, opt: Optional[torch.optim.Optimizer] = None):
    """One pass of solver.py:139-186.  Returns (outputs, losses, grads-before-clipping)."""
    model.zero_grad()
    out = model(*batch.model_args())
    L = oracle_losses(out, batch.labels, cfg)
    L["total"].backward()
    grads = {n: (None if p.grad is None else p.grad.detach().clone())
             for n, p in model.named_parameters()}
    if opt is not None:
        torch.nn.utils.clip_grad_value_([p for p in model.parameters() if p.requires_grad],
                                        cfg.clip)
        opt.step()
    return out, L, grads
