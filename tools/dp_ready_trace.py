"""When does each gradient bucket of a data-parallel step become ready?  Eager step, timed events
at the notifications (rank 0 prints ms after the start of the step).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_ready_trace.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from mmda_b200 import MISA, FusedTrainer, mosei_config
    from mmda_b200.synthetic import batch_for
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    cfg = mosei_config(vocab_size=20000, batch_size=256)
    torch.manual_seed(0)
    tr = FusedTrainer(MISA(cfg).to(dev).train(), process_group=dist.group.WORLD, use_graph=False)
    b = batch_for(cfg, seed=1 + rank, lengths=os.environ.get("LENGTHS", "full"))
    args = [b.sentences.to(dev), b.visual.to(dev), b.acoustic.to(dev), b.lengths, b.labels.to(dev)]
    for _ in range(5):
        tr.step(*args)
    torch.cuda.synchronize()
    for rep in range(3):
        tr.trace_ready = {}
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        tr.step(*args)
        t1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print({k: round(t0.elapsed_time(v), 3) for k, v in tr.trace_ready.items()},
                  "step", round(t0.elapsed_time(t1), 3))
    tr.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
