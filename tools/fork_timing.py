"""Per-stream timeline of the four fork/join regions of one eager C2 step (encoders forward, head
chains forward, head chains backward, encoders backward): when does each modality's stream finish
relative to the fork?  Shows which stream the join waits for.

    python tools/fork_timing.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from mmda_b200 import MISA, FusedTrainer, mosei_config
    from mmda_b200.synthetic import batch_for
    dev = torch.device("cuda:0")
    cfg = mosei_config(vocab_size=20000, batch_size=256)
    torch.manual_seed(0)
    tr = FusedTrainer(MISA(cfg).to(dev).train(), use_graph=False)
    b = batch_for(cfg, seed=1, lengths="full")
    args = [b.sentences.to(dev), b.visual.to(dev), b.acoustic.to(dev), b.lengths, b.labels.to(dev)]
    for _ in range(3):
        tr.step(*args)
    torch.cuda.synchronize()
    tr.eng.fork_log = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    tr.step(*args)
    t1.record()
    torch.cuda.synchronize()
    log, tr.eng.fork_log = tr.eng.fork_log, None
    print(f"eager step {t0.elapsed_time(t1):.3f} ms")
    names = ["encoders fwd", "heads fwd", "heads bwd", "encoders bwd"]
    region, start = -1, None
    for tag, ev in log:
        if tag == "start":
            region += 1
            start = ev
            print(f"{names[region % 4]}: fork at {t0.elapsed_time(ev):.3f} ms")
        elif tag.startswith("  "):
            print(f"      {tag.strip()}: +{start.elapsed_time(ev):.3f} ms")
        else:
            print(f"    {tag}: done +{start.elapsed_time(ev):.3f} ms")


if __name__ == "__main__":
    main()
