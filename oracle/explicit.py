"""First-principles restatement of the third-party arithmetic on the hot path.
TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT (see oracle/misa_oracle.py header).

The reference delegates all arithmetic to torch (unpinned upstream; 2.11.0 in this image):
``nn.LSTM``, ``pack_padded_sequence`` / ``pad_packed_sequence``, ``nn.LayerNorm``,
``nn.TransformerEncoderLayer``, ``nn.BCELoss``, ``clip_grad_value_`` and ``Adam`` at the call
sites /root/reference/src/models.py:47-55,155-173 and src/solver.py:108-118,185-186.  This file
restates their published definitions with explicit loops (numpy / small torch ops), so the
semantics the CUDA kernels must reproduce are written down rather than inherited:

* LSTM gate order i,f,g,o; h0=c0=0; reverse direction runs t=L_b-1..0 per sample;
* packing: ``batch_sizes[t] = #{b: L_b > t}``, rows time-major inside the length-sorted order;
* LayerNorm biased variance, eps inside the sqrt;
* post-norm encoder layer, 2 heads, scale 1/sqrt(head_dim);
* BCE with log clamped at -100; Adam without weight decay, bias-corrected.

``tests/test_oracle_explicit.py`` checks each against torch on CPU.
"""
from __future__ import annotations

import numpy as np


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def pack_indices(lengths: np.ndarray, sorted_idx: np.ndarray):
    """Packing indices for ``pack_padded_sequence(enforce_sorted=False)``.

    ``sorted_idx`` is the permutation returned by the host's descending sort of ``lengths``
    (torch's CPU sort is not stable, SURVEY.md hard part 4, so it is an *input*).  Returns
    ``batch_sizes (Tmax,)``, ``offsets (Tmax+1,)`` (exclusive prefix sum) and ``unsorted_idx``.
    Packed row of token (t, sorted position j) is ``offsets[t] + j``.
    """
    lengths = np.asarray(lengths, dtype=np.int64)
    sorted_idx = np.asarray(sorted_idx, dtype=np.int64)
    ls = lengths[sorted_idx]
    assert np.all(ls[:-1] >= ls[1:]), "sorted_idx does not sort lengths descending"
    tmax = int(ls[0]) if ls.size else 0
    batch_sizes = np.array([(ls > t).sum() for t in range(tmax)], dtype=np.int64)
    offsets = np.zeros(tmax + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(batch_sizes)
    unsorted = np.empty_like(sorted_idx)
    unsorted[sorted_idx] = np.arange(sorted_idx.size)
    return batch_sizes, offsets, unsorted


def lstm_direction(x, lengths, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """One LSTM direction over a padded time-major batch.

    x (T,B,I) float64; returns y (T,B,H) (zero past each length) and final h (B,H), c (B,H).
    """
    T, B, _ = x.shape
    H = w_hh.shape[1]
    y = np.zeros((T, B, H), dtype=x.dtype)
    hn = np.zeros((B, H), dtype=x.dtype)
    cn = np.zeros((B, H), dtype=x.dtype)
    for b in range(B):
        h = np.zeros(H, dtype=x.dtype)
        c = np.zeros(H, dtype=x.dtype)
        L = int(lengths[b])
        steps = range(L - 1, -1, -1) if reverse else range(L)
        for t in steps:
            g = w_ih @ x[t, b] + b_ih + w_hh @ h + b_hh
            i, f, gg, o = sigmoid(g[:H]), sigmoid(g[H:2 * H]), np.tanh(g[2 * H:3 * H]), \
                sigmoid(g[3 * H:])
            c = f * c + i * gg
            h = o * np.tanh(c)
            y[t, b] = h
        hn[b], cn[b] = h, c
    return y, hn, cn


def gru_direction(x, lengths, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """One nn.GRU direction (gate rows r,z,n; reference models.py:39 when rnncell != 'lstm'):
    r = s(W_ir x + b_ir + W_hr h + b_hr), z likewise, n = tanh(W_in x + b_in + r*(W_hn h + b_hn)),
    h' = (1-z) n + z h.  Returns y (T,B,H) and the final h (B,H)."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    y = np.zeros((T, B, H), dtype=x.dtype)
    hn = np.zeros((B, H), dtype=x.dtype)
    for b in range(B):
        h = np.zeros(H, dtype=x.dtype)
        L = int(lengths[b])
        for t in (range(L - 1, -1, -1) if reverse else range(L)):
            gx = w_ih @ x[t, b] + b_ih
            gh = w_hh @ h + b_hh
            r = sigmoid(gx[:H] + gh[:H])
            z = sigmoid(gx[H:2 * H] + gh[H:2 * H])
            n = np.tanh(gx[2 * H:] + r * gh[2 * H:])
            h = (1.0 - z) * n + z * h
            y[t, b] = h
        hn[b] = h
    return y, hn


def bilstm(x, lengths, p, prefix=""):
    """Bidirectional layer.  ``p`` maps torch's names (weight_ih_l0, ..._reverse) to arrays.
    Returns y (T,B,2H) = [fwd | bwd] and final h (2,B,H)."""
    yf, hf, _ = lstm_direction(x, lengths, p[prefix + "weight_ih_l0"], p[prefix + "weight_hh_l0"],
                               p[prefix + "bias_ih_l0"], p[prefix + "bias_hh_l0"], False)
    yb, hb, _ = lstm_direction(x, lengths, p[prefix + "weight_ih_l0_reverse"],
                               p[prefix + "weight_hh_l0_reverse"], p[prefix + "bias_ih_l0_reverse"],
                               p[prefix + "bias_hh_l0_reverse"], True)
    return np.concatenate([yf, yb], axis=2), np.stack([hf, hb], axis=0)


def layer_norm(x, gamma, beta, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * gamma + beta


def encoder_features(x, lengths, p1, p2, gamma, beta):
    """models.py:163-180 + :203 -> (B,4H) = [h1_fwd | h2_fwd | h1_bwd | h2_bwd]."""
    T = int(np.max(lengths))
    x = x[:T]
    y1, h1 = bilstm(x, lengths, p1)
    n1 = layer_norm(y1, gamma, beta)            # padded rows become beta, then dropped by re-pack
    _, h2 = bilstm(n1, lengths, p2)
    return np.concatenate([h1[0], h2[0], h1[1], h2[1]], axis=1)


def encoder_layer(x, p, nhead=2, eps=1e-5):
    """Post-norm TransformerEncoderLayer in eval mode.  x (S,B,d).  ``p`` uses torch's key names
    relative to ``transformer_encoder.layers.0.``."""
    S, B, d = x.shape
    hd = d // nhead
    qkv = x @ p["self_attn.in_proj_weight"].T + p["self_attn.in_proj_bias"]
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    ctx = np.zeros_like(x)
    for b in range(B):
        for h in range(nhead):
            sl = slice(h * hd, (h + 1) * hd)
            s = (q[:, b, sl] @ k[:, b, sl].T) / np.sqrt(hd)
            s = np.exp(s - s.max(-1, keepdims=True))
            a = s / s.sum(-1, keepdims=True)
            ctx[:, b, sl] = a @ v[:, b, sl]
    attn = ctx @ p["self_attn.out_proj.weight"].T + p["self_attn.out_proj.bias"]
    x1 = layer_norm(x + attn, p["norm1.weight"], p["norm1.bias"], eps)
    ff = np.maximum(x1 @ p["linear1.weight"].T + p["linear1.bias"], 0.0)
    ff = ff @ p["linear2.weight"].T + p["linear2.bias"]
    return layer_norm(x1 + ff, p["norm2.weight"], p["norm2.bias"], eps)


def bce_mean(s, y):
    """nn.BCELoss(reduction='mean') with the log clamp at -100."""
    ls = np.maximum(np.log(s), -100.0)
    l1s = np.maximum(np.log(1.0 - s), -100.0)
    return float(np.mean(-(y * ls + (1.0 - y) * l1s)))


def adam_clip_step(p, g, m, v, step, lr, clip=1.0, b1=0.9, b2=0.999, eps=1e-8):
    """clip_grad_value_(clip) then torch.optim.Adam (no weight decay, no amsgrad); ``step`` is
    the 1-based step count after the increment."""
    g = np.clip(g, -clip, clip)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v
