"""world_size-2 gloo tests (CPU) of the data-parallel host logic: (1) the batch-coupled losses
computed from sharded batches + all-reduced statistic segments equal the global-batch loss and
gradients; (2) the trainer's bucketed gradient all-reduce covers exactly the
active range of the arena, bucket by bucket, in the backward's ready order."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _worker_loss(rank, world, port, q):
    _init(rank, world, port)
    from loss_phases_ref import phased_loss
    from test_loss_phases_cpu import _rand_outputs
    X0, O, R, s, t, y = _rand_outputs(32, 16, 6, 5)
    w = dict(diff=0.3, sim=0.7, recon=0.7, conf=0.3)
    Lg, gg = phased_loss(X0, O, R, s, t, y, 32.0, w)        # global batch, no communication
    sl = slice(rank * 16, (rank + 1) * 16)

    def reduce(tn):
        tn = tn.clone()
        dist.all_reduce(tn, op=dist.ReduceOp.SUM)
        return tn
    L, g = phased_loss(X0[sl], O[:, sl], R[:, sl], s[sl], t[sl], y[sl], 32.0, w, reduce=reduce)
    ok = all(abs(float(L[k]) - float(Lg[k])) < 1e-10 for k in Lg)
    ok = ok and float((g["d_tokens"] - gg["d_tokens"][sl]).abs().max()) < 1e-12
    ok = ok and float((g["d_scores"] - gg["d_scores"][sl]).abs().max()) < 1e-12
    ok = ok and float((g["d_tcp"] - gg["d_tcp"][sl]).abs().max()) < 1e-12
    q.put((rank, ok))
    dist.destroy_process_group()


def _worker_buckets(rank, world, port, q):
    _init(rank, world, port)
    import mmda_b200.engine as E
    from mmda_b200._lib import LIB
    E._DRYRUN = True
    LIB.call = lambda name, *a: 0
    LIB.raw = lambda name: (lambda *a: 4096)
    from mmda_b200 import MISA, mosei_config
    from mmda_b200.trainer import FusedTrainer
    torch.manual_seed(0)
    model = MISA(mosei_config(vocab_size=40))
    tr = FusedTrainer(model, process_group=dist.group.WORLD)
    assert tr.world == 2
    tr.g_arena.fill_(float(rank + 1))
    order = []
    orig = tr._allreduce

    def spy(t, async_op=False):
        order.append((t.data_ptr() - tr.g_arena.data_ptr()) // 4)
        return orig(t, async_op=async_op)
    tr._allreduce = spy
    for tag in ("fusion", "heads", "enc_v", "enc_a", "enc_t"):
        tr._on_ready(tag)
    for wk in tr._pending:
        wk.wait()
    ok = bool((tr.g_arena[:tr.n_active] == 3.0).all()) and bool((tr.g_arena[tr.n_active:] == rank + 1).all())
    ok = ok and order == [lo for lo, hi in tr.ranges if hi > lo] and order == sorted(order)
    # a full (stubbed) step with the stats all-reduces must not deadlock
    from mmda_b200.synthetic import batch_for
    b = batch_for(model.config, seed=rank, lengths="ragged", batch=8, seq_len=5)
    tr._allreduce = orig
    tr.step(b.sentences, b.visual, b.acoustic, b.lengths, b.labels)
    q.put((rank, ok))
    dist.destroy_process_group()


def _worker_unequal(rank, world, port, q):
    """ranks with different local batch sizes: every rank raises (no deadlock, no skewed means)"""
    _init(rank, world, port)
    import mmda_b200.engine as E
    from mmda_b200._lib import LIB, MmdaError
    E._DRYRUN = True
    LIB.call = lambda name, *a: 0
    LIB.raw = lambda name: (lambda *a: 4096)
    from mmda_b200 import MISA, mosei_config
    from mmda_b200.trainer import FusedTrainer
    from mmda_b200.synthetic import batch_for
    torch.manual_seed(0)
    model = MISA(mosei_config(vocab_size=40))
    tr = FusedTrainer(model, process_group=dist.group.WORLD)
    b = batch_for(model.config, seed=rank, lengths="ragged", batch=8 - 2 * rank, seq_len=5)
    try:
        tr.step(b.sentences, b.visual, b.acoustic, b.lengths, b.labels)
        ok = False
    except MmdaError as e:
        ok = "unequal local batches (6..8" in str(e)
    # equal batches afterwards: the step goes through and the size is remembered
    b = batch_for(model.config, seed=rank, lengths="ragged", batch=8, seq_len=5)
    tr.step(b.sentences, b.visual, b.acoustic, b.lengths, b.labels)
    ok = ok and tr._equal_B == 8
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_worker_loss, _worker_buckets, _worker_unequal])
def test_world2_gloo(worker):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=180) for _ in range(2))
    [p.join(60) for p in procs]
    assert res == [(0, True), (1, True)], res
    assert all(p.exitcode == 0 for p in procs)
