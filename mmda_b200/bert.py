"""BERT-base text encoder on the hand-written kernels.

Replaces ``self.bertmodel(input_ids, attention_mask, token_type_ids)[0]`` + the masked mean at
reference src/models.py:41-45,186-198.  ``model.bertmodel`` stays an HF ``BertModel`` *parameter
container* (same ``state_dict`` keys, so bert-base checkpoints load); none of its ``forward``
methods runs.  Dense layers go through the tcgen05 GEMM (3xTF32 in fp32 mode, bf16 operands in
bf16 mode), everything else through ``csrc/bert.cu`` / ``csrc/elementwise.cu``.

Freeze contract (src/solver.py:66-73): parameters with ``requires_grad=False`` get no weight
gradient; the data gradient still flows through every layer down to the lowest trainable tensor
(the embeddings stay trainable in the reference, so all 12 layers run their dgrad).  The pooler
is never used: its gradients stay ``None``.
"""
from __future__ import annotations

from typing import Dict, Optional

import os

import torch

from ._lib import MmdaError

PFX = "bertmodel."


def _ptr(t):
    return None if t is None else t.data_ptr()


class BertEngine:
    def __init__(self, eng):
        self.eng, self.k = eng, eng.k
        bc = eng.model.bertmodel.config
        self.H, self.nh, self.I, self.L = (bc.hidden_size, bc.num_attention_heads,
                                           bc.intermediate_size, bc.num_hidden_layers)
        self.eps = float(bc.layer_norm_eps)
        self.p_h, self.p_a = float(bc.hidden_dropout_prob), float(bc.attention_probs_dropout_prob)
        self.V, self.max_pos = bc.vocab_size, bc.max_position_embeddings
        if bc.hidden_act != "gelu":
            raise MmdaError(f"bert: hidden_act={bc.hidden_act!r} (erf GELU only)")
        if getattr(bc, "position_embedding_type", "absolute") != "absolute":
            raise MmdaError("bert: absolute position embeddings only")
        if self.H != self.nh * 64:
            raise MmdaError("bert: head_dim must be 64 (bert-base geometry)")
        self._pver = None
        self.saved = None

    # ------------------------------------------------------------------ parameters ---------
    def params(self) -> Dict[str, torch.Tensor]:
        ps = list(self.eng.model.bertmodel.named_parameters())
        ver = tuple(p.data_ptr() for _, p in ps)
        if ver != self._pver:
            from . import engine as _engine
            for n, p in ps:
                if _engine._DRYRUN:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise MmdaError(f"bert parameter {n} must be a contiguous CUDA float32 tensor "
                                    "(no CPU fallback)")
            self._P = {PFX + n: p.data for n, p in ps}
            self._pver = ver
        return self._P

    def trainable(self):
        return {PFX + n for n, p in self.eng.model.bertmodel.named_parameters()
                if p.requires_grad and not n.startswith("pooler.")}

    # ------------------------------------------------------------------ GEMM helpers -------
    def _op(self, name, x):
        """tensor-core operand copy (hi/lo tf32 split or bf16) of the 2-D fp32 view x"""
        return self.eng._prep(name, x)

    def _wop(self, P, name):
        ops = self._wops
        if name not in ops:
            ops[name] = self.eng._prep("bertW_" + name, P[name], split=True)
        return ops[name]

    def _mm(self, a_mn, b_mn, M, N, K, A, B, C, bias=None, acc=False):
        self.k.gemm_tc(self.eng.tc_kind, a_mn, b_mn, M, N, K, A, B, C, bias=bias,
                       mode=1 if acc else 0, split_k=0 if acc else 1)

    def _gelu(self, l, pre, M):
        """GELU of layer l's intermediate pre-activation -> tensor-core operand.  In bf16 mode the
        activation only ever feeds tcgen05 GEMMs, so the kernel writes the bf16 copy directly."""
        I = self.I
        if self.eng.tc_kind == 1:
            act_bf = self.eng.buf(f"bert_actbf_{l}", M, I, dtype=torch.bfloat16)
            self.k._c("mmda_gelu_forward", _ptr(pre), None, _ptr(act_bf), M * I)
            return act_bf, None
        act = self.eng.buf(f"bert_act_{l}", M, I)
        self.k._c("mmda_gelu_forward", _ptr(pre), _ptr(act), None, M * I)
        return self._op("bert_opI", act)

    def _att_sfx(self, S):
        """bf16 mode runs the attention core on the tensor pipe (sequences up to 64 tokens)"""
        return "_mma" if self.eng.tc_kind == 1 and S <= 64 and \
            os.environ.get("MMDA_BERT_ATT", "mma") != "simt" else ""

    def _ln(self, x, res, g, b, y, mean, rstd):
        self.k._c("mmda_layernorm_forward", _ptr(x), x.stride(0), _ptr(res),
                  0 if res is None else res.stride(0), _ptr(g), _ptr(b), _ptr(y), y.stride(0),
                  _ptr(mean), _ptr(rstd), x.shape[0], x.shape[1], self.eps)

    # ------------------------------------------------------------------ forward ------------
    def forward(self, ids, types, mask, train: bool, drop: bool, seed: int,
                seed_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B,S) int64 ids / token types / attention mask -> utterance_text (B, hidden)."""
        eng, k, H, I, nh = self.eng, self.k, self.H, self.I, self.nh
        eng.params()                 # binds the workspace to the model's device
        P = self.params()
        k.bind_stream()
        if ids.dim() != 2 or ids.shape != mask.shape or ids.shape != types.shape:
            raise MmdaError("bert: bert_sent / bert_sent_type / bert_sent_mask must be (B, S)")
        from . import engine as _engine
        for t in (ids, types, mask):
            if t.dtype != torch.int64 or not (t.is_cuda or _engine._DRYRUN):
                raise MmdaError("bert: inputs must be CUDA int64 tensors")
        ids, types, mask = ids.contiguous(), types.contiguous(), mask.contiguous()
        B, S = ids.shape
        M = B * S
        p_h = self.p_h if drop else 0.0
        p_a = self.p_a if drop else 0.0
        self._wops = {}
        buf = eng.buf
        E = PFX + "embeddings."
        emb = buf("bert_emb", M, H)
        k._c("mmda_bert_embed_forward", _ptr(P[E + "word_embeddings.weight"]),
             _ptr(P[E + "position_embeddings.weight"]), _ptr(P[E + "token_type_embeddings.weight"]),
             _ptr(ids), _ptr(types), B, S, H, self.V, self.max_pos, _ptr(emb))
        stats = buf("bert_ln_stats", 2 * self.L + 1, 2, M)
        x = buf("bert_x_0", M, H)
        self._ln(emb, None, P[E + "LayerNorm.weight"], P[E + "LayerNorm.bias"], x, stats[0, 0],
                 stats[0, 1])
        if p_h > 0:
            k.dropout(x, x, p_h, seed, 200, seed_dev)
        for l in range(self.L):
            Lp = f"{PFX}encoder.layer.{l}."
            QKV = buf(f"bert_qkv_{l}", M, 3 * H)
            xop = self._op("bert_opH", x)
            for j, nm in enumerate(("query", "key", "value")):
                self._mm(0, 0, M, H, H, xop, self._wop(P, Lp + f"attention.self.{nm}.weight"),
                         QKV[:, j * H:(j + 1) * H], bias=P[Lp + f"attention.self.{nm}.bias"])
            ctx = buf(f"bert_ctx_{l}", M, H)
            probs = buf(f"bert_probs_{l}", B, nh, S, S) if train else None
            k._c("mmda_bert_attention_forward" + self._att_sfx(S), _ptr(QKV), _ptr(mask), _ptr(ctx), _ptr(probs), B, S,
                 nh, 64, p_a, seed, _ptr(seed_dev), 201 + 4 * l)
            ao = buf(f"bert_ao_{l}", M, H)
            self._mm(0, 0, M, H, H, self._op("bert_opH", ctx),
                     self._wop(P, Lp + "attention.output.dense.weight"), ao,
                     bias=P[Lp + "attention.output.dense.bias"])
            if p_h > 0:
                k.dropout(ao, ao, p_h, seed, 202 + 4 * l, seed_dev)
            h1 = buf(f"bert_h1_{l}", M, H)
            self._ln(ao, x, P[Lp + "attention.output.LayerNorm.weight"],
                     P[Lp + "attention.output.LayerNorm.bias"], h1, stats[1 + 2 * l, 0],
                     stats[1 + 2 * l, 1])
            pre = buf(f"bert_pre_{l}", M, I)
            self._mm(0, 0, M, I, H, self._op("bert_opH", h1),
                     self._wop(P, Lp + "intermediate.dense.weight"), pre,
                     bias=P[Lp + "intermediate.dense.bias"])
            fo = buf(f"bert_fo_{l}", M, H)
            self._mm(0, 0, M, H, I, self._gelu(l, pre, M),
                     self._wop(P, Lp + "output.dense.weight"), fo, bias=P[Lp + "output.dense.bias"])
            if p_h > 0:
                k.dropout(fo, fo, p_h, seed, 203 + 4 * l, seed_dev)
            xn = buf(f"bert_x_{l + 1}", M, H)
            self._ln(fo, h1, P[Lp + "output.LayerNorm.weight"], P[Lp + "output.LayerNorm.bias"], xn,
                     stats[2 + 2 * l, 0], stats[2 + 2 * l, 1])
            x = xn
        utt = buf("bert_utt", B, H)
        k._c("mmda_masked_mean_forward", _ptr(x), _ptr(mask), B, S, H, _ptr(utt))
        self.saved = dict(B=B, S=S, ids=ids, types=types, mask=mask, p_h=p_h, p_a=p_a, seed=seed,
                          seed_dev=seed_dev, train=train)
        return utt

    # ------------------------------------------------------------------ backward -----------
    def backward(self, G: Dict[str, torch.Tensor], d_utt: torch.Tensor):
        """Accumulates (+=) the gradients of the trainable BERT tensors into ``G[name]``."""
        sv = self.saved
        if sv is None or not sv["train"]:
            raise MmdaError("bert.backward needs a preceding forward(train=True)")
        eng, k, H, I, nh = self.eng, self.k, self.H, self.I, self.nh
        P = self.params()
        k.bind_stream()
        B, S, mask = sv["B"], sv["S"], sv["mask"]
        M = B * S
        p_h, p_a, seed, seed_dev = sv["p_h"], sv["p_a"], sv["seed"], sv["seed_dev"]
        buf = eng.buf
        train = {n for n in self.trainable() if n in G}
        E = PFX + "embeddings."
        emb_train = any(n.startswith(E) for n in train)
        lowest = self.L
        for l in range(self.L):
            if any(n.startswith(f"{PFX}encoder.layer.{l}.") for n in train):
                lowest = l
                break
        if emb_train:
            lowest = -1
        if lowest == self.L:
            return
        stats = buf("bert_ln_stats", 2 * self.L + 1, 2, M)
        scr_g, scr_b = buf("bert_scr_g", H), buf("bert_scr_b", H)

        def gw(name):        # gradient destination of a weight, None if frozen
            return G[name] if name in train else None

        def gw_or(name, scratch):
            return G[name] if name in train else scratch

        def wgrad(dy_op, dy, x_op, wname, bname, N, K):
            """dW[N][K] += dy^T x ; db += colsum(dy)   (dy: [M][N], x: [M][K])"""
            if gw(wname) is not None:
                self._mm(1, 1, N, K, M, dy_op, x_op, G[wname], acc=True)
            if gw(bname) is not None:
                k.colsum(dy, G[bname])

        dx = buf("bert_dxA", M, H)
        k._c("mmda_masked_mean_backward", _ptr(d_utt.contiguous()), _ptr(mask), B, S, H, _ptr(dx))
        other = buf("bert_dxB", M, H)
        for l in range(self.L - 1, max(lowest, 0) - 1, -1):
            Lp = f"{PFX}encoder.layer.{l}."
            x = buf(f"bert_x_{l}", M, H)
            QKV, ctx = buf(f"bert_qkv_{l}", M, 3 * H), buf(f"bert_ctx_{l}", M, H)
            probs = buf(f"bert_probs_{l}", B, nh, S, S)
            ao, h1 = buf(f"bert_ao_{l}", M, H), buf(f"bert_h1_{l}", M, H)
            pre = buf(f"bert_pre_{l}", M, I)
            bf = eng.tc_kind == 1
            fo = buf(f"bert_fo_{l}", M, H)
            # ---- output LayerNorm(fo + h1) ----
            dsum = other
            k.layernorm_bwd(dx, fo, h1, P[Lp + "output.LayerNorm.weight"], stats[2 + 2 * l, 0],
                            stats[2 + 2 * l, 1], dsum, gw_or(Lp + "output.LayerNorm.weight", scr_g),
                            gw_or(Lp + "output.LayerNorm.bias", scr_b))
            dfo = buf("bert_dH", M, H)
            if p_h > 0:
                k.dropout(dsum, dfo, p_h, seed, 203 + 4 * l, seed_dev)
            else:
                dfo = dsum
            dfo_op = self._op("bert_opH_d", dfo)
            # ---- output.dense: act [M][I] -> fo [M][H] ----
            dact = buf("bert_dI", M, I)
            self._mm(0, 1, M, I, H, dfo_op, self._wop(P, Lp + "output.dense.weight"), dact)
            if gw(Lp + "output.dense.weight") is not None:
                act_op = (buf(f"bert_actbf_{l}", M, I, dtype=torch.bfloat16), None) if bf else \
                    self._op("bert_opI", buf(f"bert_act_{l}", M, I))
                wgrad(dfo_op, dfo, act_op, Lp + "output.dense.weight", Lp + "output.dense.bias", H, I)
            if bf:     # d(pre) only feeds GEMMs (+ the bias column sum when that bias is trainable)
                dpre_bf = buf("bert_dIbf", M, I, dtype=torch.bfloat16)
                need32 = gw(Lp + "intermediate.dense.bias") is not None
                k._c("mmda_gelu_backward", _ptr(dact), _ptr(pre), _ptr(dact) if need32 else None,
                     _ptr(dpre_bf), M * I)
                dpre_op = (dpre_bf, None)
            else:
                k._c("mmda_gelu_backward", _ptr(dact), _ptr(pre), _ptr(dact), None, M * I)
                dpre_op = self._op("bert_opI_d", dact)
            # ---- intermediate.dense: h1 [M][H] -> pre [M][I]; dh1 = dsum + dpre W_i ----
            self._mm(0, 1, M, H, I, dpre_op, self._wop(P, Lp + "intermediate.dense.weight"), dsum,
                     acc=True)
            if gw(Lp + "intermediate.dense.weight") is not None:
                wgrad(dpre_op, dact, self._op("bert_opH", h1), Lp + "intermediate.dense.weight",
                      Lp + "intermediate.dense.bias", I, H)
            # ---- attention.output LayerNorm(ao + x) ----
            dsum1 = dx
            k.layernorm_bwd(dsum, ao, x, P[Lp + "attention.output.LayerNorm.weight"],
                            stats[1 + 2 * l, 0], stats[1 + 2 * l, 1], dsum1,
                            gw_or(Lp + "attention.output.LayerNorm.weight", scr_g),
                            gw_or(Lp + "attention.output.LayerNorm.bias", scr_b))
            dao = buf("bert_dH", M, H)
            if p_h > 0:
                k.dropout(dsum1, dao, p_h, seed, 202 + 4 * l, seed_dev)
            else:
                dao = dsum1
            dao_op = self._op("bert_opH_d", dao)
            dctx = buf("bert_dctx", M, H)
            self._mm(0, 1, M, H, H, dao_op, self._wop(P, Lp + "attention.output.dense.weight"), dctx)
            if gw(Lp + "attention.output.dense.weight") is not None:
                wgrad(dao_op, dao, self._op("bert_opH", ctx), Lp + "attention.output.dense.weight",
                      Lp + "attention.output.dense.bias", H, H)
            dQKV = buf("bert_dqkv", M, 3 * H)
            k._c("mmda_bert_attention_backward" + self._att_sfx(S), _ptr(QKV), _ptr(probs), _ptr(dctx), _ptr(dQKV), B, S,
                 nh, 64, p_a, seed, _ptr(seed_dev), 201 + 4 * l)
            if l > lowest or any(Lp + f"attention.self.{nm}.{wb}" in train
                                 for nm in ("query", "key", "value") for wb in ("weight", "bias")):
                dq_op = self._op("bert_op3H_d", dQKV)
                x_op = None
                for j, nm in enumerate(("query", "key", "value")):
                    blk = eng._cols(dq_op, j * H, (j + 1) * H)
                    if l > lowest:      # the layer below (or the embeddings) needs dx
                        self._mm(0, 1, M, H, H, blk, self._wop(P, Lp + f"attention.self.{nm}.weight"),
                                 dsum1, acc=True)
                    wn, bn = Lp + f"attention.self.{nm}.weight", Lp + f"attention.self.{nm}.bias"
                    if gw(wn) is not None:
                        if x_op is None:
                            x_op = self._op("bert_opH", x)
                        wgrad(blk, dQKV[:, j * H:(j + 1) * H], x_op, wn, bn, H, H)
                    elif gw(bn) is not None:
                        k.colsum(dQKV[:, j * H:(j + 1) * H], G[bn])
            dx, other = dsum1, dsum
        if emb_train:
            if p_h > 0:
                k.dropout(dx, dx, p_h, seed, 200, seed_dev)
            emb = buf("bert_emb", M, H)
            demb = other
            k.layernorm_bwd(dx, emb, None, P[E + "LayerNorm.weight"], stats[0, 0], stats[0, 1], demb,
                            gw_or(E + "LayerNorm.weight", scr_g), gw_or(E + "LayerNorm.bias", scr_b))
            k._c("mmda_bert_embed_backward", _ptr(demb), _ptr(sv["ids"]), _ptr(sv["types"]), B, S, H,
                 self.V, _ptr(gw(E + "word_embeddings.weight")),
                 _ptr(gw(E + "position_embeddings.weight")),
                 _ptr(gw(E + "token_type_embeddings.weight")))
