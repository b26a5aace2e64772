"""C4 (BASELINE configs[3]): BERT-base text encoder + LSTM visual/acoustic encoders, fused level-2
step, batch 512, seq 50 (+2 specials), layers 0-8 frozen as Solver.build does (solver.py:66-73).
Informational timing of the hand-written path; not the bench.py headline (that is configs[1]).

    python tools/bench_c4.py [--batch 512] [--precision fp32|bf16] [--steps 5]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--seq", type=int, default=50)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=2)
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
    b = batch_for(cfg, seed=1, lengths="full", seq_len=a.seq)
    args = [t.to(dev) for t in (b.sentences, b.visual, b.acoustic)] + [b.lengths, b.labels.to(dev)] + \
           [t.to(dev) for t in (b.bert_sent, b.bert_sent_type, b.bert_sent_mask)]
    for _ in range(max(3, a.warmup)):
        L = tr.step(*args)
    torch.cuda.synchronize()
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import ClockSampler        # nvidia-smi clocks / throttle reasons during the timed region
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    clocks = ClockSampler(0)
    clocks.start()
    l0 = tr.eng.k.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        flush.zero_()                      # L2 flush between steps, inside the timed region
        L = tr.step(*args)
    e1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1) / a.steps
    flops = 3 * 2 * a.batch * (a.seq + 2) * 12 * (4 * 768 * 768 + 2 * 768 * 3072) * (1 - 0.25 * 9 / 12)
    print(json.dumps({"workload": f"C4 BERT-base + LSTM v/a, B={a.batch}, seq {a.seq}+2, layers 0-8 frozen, "
                                  "train mode, fused step", "precision": a.precision,
                      "ms_per_step": ms, "samples_per_s": a.batch / ms * 1e3, "steps": a.steps,
                      "warmup": max(3, a.warmup), "clocks": clk,
                      "launches_per_step": (tr.eng.k.launches - l0) / a.steps,
                      "bert_dense_tflops_algorithmic": flops / ms / 1e9,
                      "losses": [round(x, 5) for x in L[:6].tolist()],
                      "mem_gb": torch.cuda.max_memory_allocated() / 2**30}))


if __name__ == "__main__":
    main()
