"""Informational measurement: the oracle (= the reference's own
torch modules) moved to ``cuda:0`` -- eager PyTorch + cuDNN LSTM, the only pre-existing
Blackwell code path for this workload -- timed next to the fused step on the same C2 batch.
Nothing here is a parity gate beyond "both arms see the same loss"; the numbers are written to
``gpurun_out/info_oracle_cuda.json`` and copied to ``profiles/`` by hand.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_info_oracle_on_cuda_vs_fused_step():
    from mmda_b200 import MISA, FusedTrainer
    from mmda_b200.config import mosei_config
    from mmda_b200.synthetic import batch_for
    from oracle.misa_oracle import oracle_build, oracle_optimizer, oracle_step

    dev = torch.device("cuda:0")
    cfg = mosei_config(vocab_size=20000)     # train mode: cuDNN's RNN backward requires it
    batch = batch_for(cfg, seed=1, lengths="full")
    ref = oracle_build(cfg, seed=1234)
    state = {k: v.clone() for k, v in ref.state_dict().items()}
    ref = ref.to(dev).train()
    opt = oracle_optimizer(ref, cfg)
    dbatch = type(batch)(batch.sentences.to(dev), batch.visual.to(dev), batch.acoustic.to(dev),
                         batch.labels.to(dev), batch.lengths, batch.bert_sent.to(dev),
                         batch.bert_sent_type.to(dev), batch.bert_sent_mask.to(dev))

    def time_arm(fn, warm=3, steps=10):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    first = {}

    def ref_fn():
        _, L, _ = oracle_step(ref, dbatch, cfg, opt)
        first.setdefault("ref", float(L["total"]))

    ref_ms = time_arm(ref_fn)

    torch.manual_seed(1234)
    model = MISA(cfg)
    model.load_state_dict(state)
    model = model.to(dev).train()
    tr = FusedTrainer(model)
    args = (dbatch.sentences, dbatch.visual, dbatch.acoustic, batch.lengths, dbatch.labels)

    def our_fn():
        L = tr.step(*args)
        first.setdefault("ours", float(L[5]))

    our_ms = time_arm(our_fn)
    B = cfg.batch_size
    info = {"workload": "C2 MOSEI-shape, B=256, T=50, train mode (dropout on), resident inputs",
            "oracle_on_cuda_ms": ref_ms, "oracle_on_cuda_samples_per_s": B / ref_ms * 1e3,
            "fused_step_ms": our_ms, "fused_step_samples_per_s": B / our_ms * 1e3,
            "speedup": ref_ms / our_ms, "first_step_total_loss": first,
            "note": "oracle arm = eager PyTorch modules + cuDNN LSTM + torch autograd + "
                    "torch.optim.Adam; CUDA_LAUNCH_BLOCKING unset"}
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "info_oracle_cuda.json"), "w") as f:
        json.dump(info, f, indent=1)
    assert abs(first["ref"] - first["ours"]) <= 0.05 * abs(first["ref"])    # dropout masks differ


def test_info_hf_bert_on_cuda_c4():
    """Informational: the oracle's HF BertModel + torch modules on the same B200 at the C4 shape
    (B=512, seq 50+2, layers 0-8 frozen, fp32, eager PyTorch) -- the library baseline the
    hand-written BERT path (tools/bench_c4.py) is compared with."""
    from mmda_b200.config import mosei_config
    from mmda_b200.synthetic import batch_for
    from oracle.misa_oracle import oracle_build, oracle_optimizer, oracle_step
    dev = torch.device("cuda:0")
    cfg = mosei_config(vocab_size=20000, batch_size=512, use_bert=True)
    ref = oracle_build(cfg, seed=1234)
    for n, p in ref.named_parameters():
        if "bertmodel.encoder.layer" in n and int(n.split("encoder.layer.")[-1].split(".")[0]) <= 8:
            p.requires_grad = False
    ref = ref.to(dev).train()
    opt = oracle_optimizer(ref, cfg)
    b = batch_for(cfg, seed=1, lengths="full")
    db = type(b)(b.sentences.to(dev), b.visual.to(dev), b.acoustic.to(dev), b.labels.to(dev),
                 b.lengths, b.bert_sent.to(dev), b.bert_sent_type.to(dev), b.bert_sent_mask.to(dev))
    res = {}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for _ in range(2):
            oracle_step(ref, db, cfg, opt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            oracle_step(ref, db, cfg, opt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        res["tf32_matmul" if tf32 else "fp32_matmul"] = {"ms_per_step": ms,
                                                         "samples_per_s": 512 / ms * 1e3}
    torch.backends.cuda.matmul.allow_tf32 = False
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "info_hf_bert_cuda_c4.json"), "w") as f:
        json.dump({"workload": "C4: oracle (HF BertModel + cuDNN LSTM, eager PyTorch) on cuda:0, "
                               "B=512, seq 50+2, layers 0-8 frozen, train mode", **res}, f, indent=1)
