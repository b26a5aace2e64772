"""Time the long-K / small-output forward linears: plain 32x32 SIMT tiles vs the in-CTA split-K
kernel (split_k=-1) vs atomic split-K."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmda_b200.engine import Kernels

dev = torch.device("cuda:0")
k = Kernels(); k.bind_stream()
for (M, N, K) in [(256, 128, 1200), (1536, 128, 2048), (256, 128, 128)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev)
    for sk in (1, -1, 4):
        kw = dict(tb=True, split_k=sk)
        if sk != 4:
            kw.update(bias=b, act=2)
        for _ in range(5):
            k.gemm(x, w, out, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            k.gemm(x, w, out, **kw)
        e1.record(); torch.cuda.synchronize()
        print(f"{M}x{N}x{K} split_k={sk}: {e0.elapsed_time(e1) * 10:.1f} us/launch")
