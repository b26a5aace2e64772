"""Check the first-principles restatement (oracle/explicit.py) against torch on CPU: this pins
the third-party semantics (gate order, packing, LayerNorm, encoder layer, BCE, Adam) that the
CUDA kernels reproduce."""
import numpy as np
import torch
import torch.nn as nn

from oracle import explicit as E


def _np(d):
    return {k: v.detach().double().numpy() for k, v in d.items()}


def test_bilstm_and_encoder_features():
    torch.manual_seed(0)
    T, B, I = 6, 5, 4
    lengths = torch.tensor([3, 6, 1, 6, 2])
    x = torch.randn(T, B, I, dtype=torch.float64)
    r1 = nn.LSTM(I, I, bidirectional=True).double()
    r2 = nn.LSTM(2 * I, I, bidirectional=True).double()
    ln = nn.LayerNorm(2 * I).double()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5); ln.bias.uniform_(-0.5, 0.5)
    from oracle.misa_oracle import OracleMISA
    ref = OracleMISA.encode(x, lengths, r1, r2, ln).detach().numpy()
    got = E.encoder_features(x.numpy(), lengths.numpy(), _np(dict(r1.named_parameters())),
                             _np(dict(r2.named_parameters())), ln.weight.detach().numpy(),
                             ln.bias.detach().numpy())
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)


def test_gru_direction_matches_torch():
    """nn.GRU semantics (gate rows r,z,n; b_hn inside the r product) -- models.py:39 variant."""
    from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
    torch.manual_seed(1)
    T, B, I, H = 6, 5, 4, 3
    lengths = torch.tensor([3, 6, 1, 6, 2])
    x = torch.randn(T, B, I, dtype=torch.float64)
    gru = nn.GRU(I, H, bidirectional=True).double()
    out, hn = gru(pack_padded_sequence(x, lengths, enforce_sorted=False))
    y_ref, _ = pad_packed_sequence(out, total_length=T)
    p = _np(dict(gru.named_parameters()))
    for d, suf in enumerate(("", "_reverse")):
        y, h = E.gru_direction(x.numpy(), lengths.numpy(), p["weight_ih_l0" + suf],
                               p["weight_hh_l0" + suf], p["bias_ih_l0" + suf],
                               p["bias_hh_l0" + suf], reverse=bool(d))
        np.testing.assert_allclose(y, y_ref[:, :, d * H:(d + 1) * H].detach().numpy(), atol=1e-12)
        np.testing.assert_allclose(h, hn[d].detach().numpy(), atol=1e-12)


def test_pack_indices_match_torch():
    g = torch.Generator().manual_seed(3)
    for B in (1, 7, 40):
        lengths = torch.randint(1, 9, (B,), generator=g)
        x = torch.randn(int(lengths.max()), B, 3, generator=g)
        pk = nn.utils.rnn.pack_padded_sequence(x, lengths, enforce_sorted=False)
        bs, off, uns = E.pack_indices(lengths.numpy(), pk.sorted_indices.numpy())
        np.testing.assert_array_equal(bs, pk.batch_sizes.numpy())
        np.testing.assert_array_equal(uns, pk.unsorted_indices.numpy())
        # packed row (t, j) holds x[t, sorted_idx[j]]
        data = pk.data.numpy()
        si = pk.sorted_indices.numpy()
        for t in range(len(bs)):
            for j in range(bs[t]):
                np.testing.assert_array_equal(data[off[t] + j], x[t, si[j]].numpy())


def test_encoder_layer():
    torch.manual_seed(1)
    layer = nn.TransformerEncoderLayer(d_model=16, nhead=2).double().eval()
    x = torch.randn(6, 3, 16, dtype=torch.float64)
    ref = layer(x).detach().numpy()
    got = E.encoder_layer(x.numpy(), _np(dict(layer.named_parameters())))
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)


def test_bce_clamp_and_adam():
    s = np.array([0.0, 1.0, 0.3, 0.9]); y = np.array([1.0, 0.0, 1.0, 0.0])
    ref = nn.BCELoss()(torch.tensor(s), torch.tensor(y)).item()
    assert abs(E.bce_mean(s, y) - ref) < 1e-12
    torch.manual_seed(2)
    p = torch.randn(50, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([p], lr=1e-2)
    pn, m, v = p.detach().numpy().copy(), np.zeros(50), np.zeros(50)
    for step in range(1, 4):
        g = torch.randn(50, dtype=torch.float64) * 3
        p.grad = g.clone()
        torch.nn.utils.clip_grad_value_([p], 1.0)
        opt.step()
        pn, m, v = E.adam_clip_step(pn, g.numpy(), m, v, step, 1e-2)
        np.testing.assert_allclose(pn, p.detach().numpy(), rtol=0, atol=1e-12)
