"""Where does the step go?  CUDA-event timing of the main-stream phases of the fused step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmda_b200 import MISA, mosei_config
from mmda_b200.synthetic import batch_for
from mmda_b200.trainer import FusedTrainer

dev = torch.device("cuda:0")
cfg = mosei_config(vocab_size=20000)
torch.manual_seed(0)
m = MISA(cfg).to(dev).train()
tr = FusedTrainer(m)
eng = m.engine
b = batch_for(cfg, seed=1, lengths=sys.argv[1] if len(sys.argv) > 1 else "full")
args = (b.sentences.to(dev), b.visual.to(dev), b.acoustic.to(dev), b.lengths)
lab = b.labels.to(dev)
for _ in range(3):
    tr.step(*args, lab)
marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
# monkeypatch phase boundaries
orig_fork = eng._fork
calls = [0]
def fork(fns):
    calls[0] += 1
    mark(f"fork{calls[0]}_begin"); orig_fork(fns); mark(f"fork{calls[0]}_end")
eng._fork = fork
orig_lag = tr.loss_and_grads
def lag(*a, **k):
    mark("loss_begin"); r = orig_lag(*a, **k); mark("loss_end"); return r
tr.loss_and_grads = lag
N = 10
acc = {}
for it in range(N):
    marks.clear(); calls[0] = 0
    mark("step_begin"); tr.step(*args, lab); mark("step_end")
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(marks, marks[1:]):
        acc.setdefault(f"{n0} -> {n1}", 0.0)
        acc[f"{n0} -> {n1}"] += e0.elapsed_time(e1)
tot = 0
for k, v in acc.items():
    print(f"{v / N:8.3f} ms  {k}"); tot += v / N
print(f"{tot:8.3f} ms total")
