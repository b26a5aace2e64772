"""Kernel timeline of one C2 training step under CUDA-graph replay (torch.profiler / CUPTI):
which stream runs what, when, and where the critical path has gaps.

    python tools/timeline.py [out.txt]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/timeline.py [out.txt]
        (data-parallel step: rank 0 prints its own timeline, NCCL kernels included)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from torch.profiler import ProfilerActivity, profile
    from mmda_b200 import MISA, FusedTrainer, mosei_config
    from mmda_b200.synthetic import batch_for
    world, rank, pg = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), None
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    cfg = mosei_config(vocab_size=20000, batch_size=256)
    torch.manual_seed(0)
    tr = FusedTrainer(MISA(cfg).to(dev).train(), process_group=pg)
    b = batch_for(cfg, seed=1 + rank, lengths=os.environ.get("LENGTHS", "full"))
    args = [b.sentences.to(dev), b.visual.to(dev), b.acoustic.to(dev), b.lengths, b.labels.to(dev)]
    for _ in range(6):
        tr.step(*args)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr.step(*args)
        torch.cuda.synchronize()
    tr.close()
    if world > 1:
        dist.barrier(device_ids=[dev.index])
        dist.destroy_process_group()
        if rank != 0:
            return
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = []
    for e in evs:
        tr_ = e.time_range
        ks.append((tr_.start, tr_.end, e.name, getattr(e, "device_resource_id", -1)))
    ks.sort()
    if not ks:
        print("no CUDA events captured")
        return
    t0 = ks[0][0]
    streams = sorted({k[3] for k in ks})
    sid = {s: i for i, s in enumerate(streams)}
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
    print(f"kernels: {len(ks)}  streams: {len(streams)}  span: {(max(k[1] for k in ks) - t0) / 1e3:.3f} ms", file=out)
    for s, e, n, st in ks:
        short = n.replace("void ", "").replace("(anonymous namespace)::", "")[:60]
        print(f"{(s - t0) / 1e3:8.3f} {(e - s) / 1e3:7.3f}  s{sid[st]:<2d} {short}", file=out)


if __name__ == "__main__":
    main()
