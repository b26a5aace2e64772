"""GPU dev driver for the tensor-core recurrence (csrc/lstm_tc.cu): parity against the SIMT
recurrence (csrc/lstm.cu) and an fp64 torch restatement on the same packed inputs, then timing.

    python tools/dev_lstm_tc.py [B H T ragged]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmda_b200._lib import LIB  # noqa: E402


def P(t):
    return None if t is None else t.data_ptr()


def make_case(B, H, T, ragged, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    if ragged:
        lens = torch.randint(1, T + 1, (B,), generator=g)
        lens[0] = T
    else:
        lens = torch.full((B,), T)
    ls, si = torch.sort(lens, descending=True)
    Tmax = int(ls[0])
    bs = torch.tensor([(ls > t).sum() for t in range(Tmax)])
    off = torch.zeros(Tmax + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(bs, 0)
    N = int(off[-1])
    gates = torch.randn(N, 8 * H, generator=g) * 0.7
    whh = [torch.nn.init.orthogonal_(torch.empty(4 * H, H), generator=g) for _ in range(2)]
    dy = torch.randn(N, 2 * H, generator=g) * 0.1
    dutt = torch.randn(B, 4 * H, generator=g) * 0.1
    d = dict(B=B, H=H, Tmax=Tmax, N=N, lens=ls.int().to(dev), sidx=si.int().to(dev),
             off=off.int().to(dev), gates=gates.to(dev), whh=[w.to(dev) for w in whh],
             dy=dy.to(dev), dutt=dutt.to(dev), ls_cpu=ls, si_cpu=si, off_cpu=off)
    return d


def ref64(c):
    """fp64 restatement (oracle/explicit.py::lstm_direction semantics) on the packed layout."""
    B, H, Tmax = c["B"], c["H"], c["Tmax"]
    dev = c["gates"].device
    G = c["gates"].double().view(-1, 2, H, 4)
    ls, off = c["ls_cpu"], c["off_cpu"]
    y = torch.zeros(c["N"], 2, H, dtype=torch.float64, device=dev)
    cc = torch.zeros_like(y)
    act = torch.zeros(c["N"], 2, H, 4, dtype=torch.float64, device=dev)
    utt = torch.zeros(B, 4 * H, dtype=torch.float64, device=dev)
    for d in range(2):
        W = c["whh"][d].double().view(4, H, H)        # [g][u][k]
        h = torch.zeros(B, H, dtype=torch.float64, device=dev)
        cs = torch.zeros_like(h)
        ts = range(Tmax) if d == 0 else range(Tmax - 1, -1, -1)
        for t in ts:
            n = int((ls > t).sum())
            rows = slice(int(off[t]), int(off[t]) + n)
            pre = G[rows, d] + torch.einsum("guk,bk->bug", W, h[:n])
            i, f, g_, o = (torch.sigmoid(pre[..., 0]), torch.sigmoid(pre[..., 1]),
                           torch.tanh(pre[..., 2]), torch.sigmoid(pre[..., 3]))
            cn = f * cs[:n] + i * g_
            hn = o * torch.tanh(cn)
            h = h.clone(); cs = cs.clone()
            h[:n] = hn; cs[:n] = cn
            y[rows, d] = hn; cc[rows, d] = cn
            act[rows, d] = torch.stack([i, f, g_, o], -1)
        # final states in original order: fwd = h at t = L-1, bwd = h at t = 0
    si = c["si_cpu"].to(dev)
    for j in range(B):
        L = int(ls[j])
        utt[si[j], 0:H] = y[int(off[L - 1]) + j, 0]
        utt[si[j], 2 * H:3 * H] = y[int(off[0]) + j, 1]
    return y.view(-1, 2 * H), cc.view(-1, 2 * H), act.view(-1, 8 * H), utt


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def run(c, which, dbg=None):
    B, H, Tmax, N = c["B"], c["H"], c["Tmax"], c["N"]
    dev = c["gates"].device
    st = torch.cuda.current_stream().cuda_stream
    G = c["gates"].clone()
    y = torch.zeros(N, 2 * H, device=dev)
    cc = torch.zeros(N, 2 * H, device=dev)
    utt = torch.zeros(B, 4 * H, device=dev)
    out = {}
    if which == "tc":
        nb = LIB.raw("mmda_lstm_tc_workspace_bytes")(B, H, Tmax)
        assert nb > 0, nb
        ws = torch.zeros(nb // 4 + 1, dtype=torch.int32, device=dev)
        LIB.call("mmda_lstm_tc_forward", P(G), P(c["whh"][0]), P(c["whh"][1]), P(y), P(cc),
                 P(c["lens"]), P(c["sidx"]), P(c["off"]), P(utt), 4 * H, 0, 2 * H, B, H, Tmax, 1,
                 P(ws), st)
        torch.cuda.synchronize()
        out["err_fwd"] = int(ws[0])
    else:
        LIB.call("mmda_lstm_forward", P(G), P(c["whh"][0]), P(c["whh"][1]), P(y), P(cc),
                 P(c["lens"]), P(c["sidx"]), P(c["off"]), P(utt), 4 * H, 0, 2 * H, B, H, Tmax, 1, st)
        torch.cuda.synchronize()
    out.update(y=y.clone(), c=cc.clone(), act=G.clone(), utt=utt.clone())
    # backward on the forward's own saved state
    if which == "tc":
        LIB.call("mmda_lstm_tc_backward", P(G), P(c["whh"][0]), P(c["whh"][1]), P(cc), P(c["dy"]),
                 P(c["dutt"]), 4 * H, 0, 2 * H, P(c["lens"]), P(c["sidx"]), P(c["off"]), B, H, Tmax,
                 P(ws), st)
        torch.cuda.synchronize()
        out["err_bwd"] = int(ws[0])
    else:
        nb = LIB.raw("mmda_lstm_scratch_bytes")(B, H)
        scr = torch.zeros(max(1, nb // 4), device=dev)
        LIB.call("mmda_lstm_backward", P(G), P(c["whh"][0]), P(c["whh"][1]), P(cc), P(c["dy"]),
                 P(c["dutt"]), 4 * H, 0, 2 * H, P(c["lens"]), P(c["sidx"]), P(c["off"]), P(scr), B, H,
                 Tmax, st)
        torch.cuda.synchronize()
    out["dG"] = G.clone()
    return out


def ref64_bwd(c, act, cc):
    """fp64 BPTT from fp64 saved activations: d(pre-activation gates) [N][8H]."""
    B, H, Tmax = c["B"], c["H"], c["Tmax"]
    dev = act.device
    A = act.view(-1, 2, H, 4)
    C = cc.view(-1, 2, H)
    dy = c["dy"].double().view(-1, 2, H)
    dutt = c["dutt"].double()
    ls, off, si = c["ls_cpu"], c["off_cpu"], c["si_cpu"].to(dev)
    dG = torch.zeros_like(A)
    for d in range(2):
        W = c["whh"][d].double().view(4, H, H)
        dh_rec = torch.zeros(B, H, dtype=torch.float64, device=dev)
        dc = torch.zeros_like(dh_rec)
        ts = range(Tmax - 1, -1, -1) if d == 0 else range(Tmax)
        for t in ts:
            n = int((ls > t).sum())
            rows = slice(int(off[t]), int(off[t]) + n)
            dh = dy[rows, d].clone()
            lsn = ls[:n].to(dev)
            fin = (lsn - 1 == t) if d == 0 else torch.full((n,), t == 0, device=dev)
            uo = 0 if d == 0 else 2 * H
            dh = dh + fin[:, None] * dutt[si[:n]][:, uo:uo + H]
            dh = dh + dh_rec[:n]      # rows that just became active still carry zeros
            i, f, g_, o = A[rows, d, :, 0], A[rows, d, :, 1], A[rows, d, :, 2], A[rows, d, :, 3]
            ct = C[rows, d]
            tp = t - 1 if d == 0 else t + 1
            cp = torch.zeros_like(ct)
            if 0 <= tp < Tmax:
                npv = int((ls > tp).sum())
                m = min(n, npv)
                cp[:m] = C[int(off[tp]):int(off[tp]) + m, d]
            tc = torch.tanh(ct)
            do = dh * tc * o * (1 - o)
            dcc = dc[:n]
            dcc = dcc + dh * o * (1 - tc * tc)
            di = dcc * g_ * i * (1 - i)
            df = dcc * cp * f * (1 - f)
            dg = dcc * i * (1 - g_ * g_)
            dG[rows, d] = torch.stack([di, df, dg, do], -1)
            dc = dc.clone(); dh_rec = dh_rec.clone()
            dc[:n] = dcc * f
            pre = torch.stack([di, df, dg, do], 1)      # [n][4][H]
            dh_rec[:n] = torch.einsum("bgu,guk->bk", pre, W)
    return dG.view(-1, 8 * H)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    B, H, T, ragged = 256, 300, 50, 0
    if len(sys.argv) > 4:
        B, H, T, ragged = (int(x) for x in sys.argv[1:5])
    dev = torch.device("cuda:0")
    c = make_case(B, H, T, ragged, dev)
    plan = (torch.zeros(8, dtype=torch.int32))
    import ctypes
    arr = (ctypes.c_int * 8)()
    LIB.call("mmda_lstm_tc_plan", B, H, c["Tmax"], arr)
    print("plan S,G,BT,NT,Kp,smem_f,smem_b,ctas:", list(arr), flush=True)
    y64, c64, a64, u64 = ref64(c)
    dG64 = ref64_bwd(c, a64, c64)
    res = {}
    for which in ("simt", "tc"):
        o = run(c, which)
        res[which] = o
        row = {k: rel(o[k], r) for k, r in (("y", y64), ("c", c64), ("act", a64), ("utt", u64), ("dG", dG64))}
        row.update({k: v for k, v in o.items() if k.startswith("err")})
        print(which, "vs fp64:", json.dumps(row), flush=True)
    print("tc vs simt:", json.dumps({k: rel(res["tc"][k], res["simt"][k]) for k in ("y", "c", "act", "utt", "dG")}),
          flush=True)

    # ---- timing (forward, backward) ----
    st = torch.cuda.current_stream().cuda_stream
    N = c["N"]
    G = c["gates"].clone()
    y = torch.zeros(N, 2 * H, device=dev); cc = torch.zeros(N, 2 * H, device=dev)
    utt = torch.zeros(B, 4 * H, device=dev)
    nb = LIB.raw("mmda_lstm_tc_workspace_bytes")(B, H, c["Tmax"])
    ws = torch.zeros(nb // 4 + 1, dtype=torch.int32, device=dev)
    scr = torch.zeros(max(1, LIB.raw("mmda_lstm_scratch_bytes")(B, H) // 4), device=dev)
    Gs = res["tc"]["act"]

    def f_tc():
        G.copy_(c["gates"])
        LIB.call("mmda_lstm_tc_forward", P(G), P(c["whh"][0]), P(c["whh"][1]), P(y), P(cc), P(c["lens"]),
                 P(c["sidx"]), P(c["off"]), P(utt), 4 * H, 0, 2 * H, B, H, c["Tmax"], 1, P(ws), st)

    def f_simt():
        G.copy_(c["gates"])
        LIB.call("mmda_lstm_forward", P(G), P(c["whh"][0]), P(c["whh"][1]), P(y), P(cc), P(c["lens"]),
                 P(c["sidx"]), P(c["off"]), P(utt), 4 * H, 0, 2 * H, B, H, c["Tmax"], 1, st)

    def b_tc():
        G.copy_(Gs)
        LIB.call("mmda_lstm_tc_backward", P(G), P(c["whh"][0]), P(c["whh"][1]), P(res["tc"]["c"]), P(c["dy"]),
                 P(c["dutt"]), 4 * H, 0, 2 * H, P(c["lens"]), P(c["sidx"]), P(c["off"]), B, H, c["Tmax"],
                 P(ws), st)

    def b_simt():
        G.copy_(Gs)
        LIB.call("mmda_lstm_backward", P(G), P(c["whh"][0]), P(c["whh"][1]), P(res["tc"]["c"]), P(c["dy"]),
                 P(c["dutt"]), 4 * H, 0, 2 * H, P(c["lens"]), P(c["sidx"]), P(c["off"]), P(scr), B, H,
                 c["Tmax"], st)

    def copy_only():
        G.copy_(Gs)

    t_copy = timeit(copy_only)
    print("ms (copy subtracted): " + json.dumps({
        "fwd_tc": timeit(f_tc) - t_copy, "fwd_simt": timeit(f_simt) - t_copy,
        "bwd_tc": timeit(b_tc) - t_copy, "bwd_simt": timeit(b_simt) - t_copy, "copy": t_copy}), flush=True)

    # ---- per-step phase stamps of CTA 0 (16 slots per step; see TCL_TS in lstm_tc.cu) ----
    dbg = torch.zeros(16 * (c["Tmax"] + 1), dtype=torch.int64, device=dev)
    LIB.call("mmda_lstm_tc_set_debug_buffer", P(dbg))
    for name, fn in (("fwd", f_tc), ("bwd", b_tc)):
        dbg.zero_()
        fn(); torch.cuda.synchronize()
        d = dbg.view(-1, 16)[: c["Tmax"]].cpu().double()
        lo, hi = 5, c["Tmax"] - 2
        if hi - lo < 2:
            continue
        rel0 = (d[lo:hi] - d[lo:hi, 0:1])
        rel0[d[lo:hi] == 0] = float("nan")
        means = rel0.nanmean(0).tolist()
        step = (d[lo + 1:hi + 1, 0] - d[lo:hi, 0]).mean().item()
        print(name, "stamps (clk after step start):",
              {i: round(v) for i, v in enumerate(means) if v == v and i > 0}, "step", round(step), flush=True)
    LIB.call("mmda_lstm_tc_set_debug_buffer", None)


if __name__ == "__main__":
    main()
