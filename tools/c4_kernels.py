"""Per-kernel time of the C4 step (BERT-base text + LSTM v/a, B=512), aggregated by kernel name with
CUPTI through torch.profiler.  Informational (profiler overhead inflates launch gaps, not kernels).

    python tools/c4_kernels.py [--precision bf16] [--batch 512]
"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    from mmda_b200 import MISA, FusedTrainer, mosei_config
    from mmda_b200.synthetic import batch_for
    dev = torch.device("cuda:0")
    cfg = mosei_config(vocab_size=20000, batch_size=a.batch, use_bert=True, precision=a.precision)
    torch.manual_seed(1234)
    model = MISA(cfg)
    for n, p in model.named_parameters():
        if "bertmodel.encoder.layer" in n and int(n.split("encoder.layer.")[-1].split(".")[0]) <= 8:
            p.requires_grad = False
    model = model.to(dev).train()
    tr = FusedTrainer(model)
    b = batch_for(cfg, seed=1, lengths="full", seq_len=50)
    args = [t.to(dev) for t in (b.sentences, b.visual, b.acoustic)] + [b.lengths, b.labels.to(dev)] + \
           [t.to(dev) for t in (b.bert_sent, b.bert_sent_type, b.bert_sent_mask)]
    for _ in range(3):
        tr.step(*args)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            tr.step(*args)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            e = agg[ev.name[:90]]
            e[0] += 1
            e[1] += ev.device_time
    tot = sum(v[1] for v in agg.values())
    print(f"C4 {a.precision} B={a.batch}: kernel time {tot / a.steps / 1e3:.2f} ms/step over {a.steps} steps")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
        print(f"{t / a.steps / 1e3:8.3f} ms {100 * t / tot:5.1f}%  x{n // a.steps:4d}  {name}")


if __name__ == "__main__":
    main()
