"""Per-step phase timing of the text-size LSTM forward kernel (clock64 stamps of CTA 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmda_b200 import MISA, mosei_config
from mmda_b200.synthetic import batch_for
from mmda_b200._lib import LIB
from mmda_b200.engine import _ptr

dev = torch.device("cuda:0")
cfg = mosei_config(vocab_size=2000)
torch.manual_seed(0)
m = MISA(cfg).to(dev).eval()
b = batch_for(cfg, seed=1, lengths="full")
eng = m.engine
args = (b.sentences.to(dev), b.visual.to(dev), b.acoustic.to(dev), b.lengths)
os.environ["MMDA_STREAMS"] = "0"
eng.multi_stream = False
for _ in range(2):
    eng.forward(*args, train=True)
dbg = torch.zeros(8 * 64, dtype=torch.int64, device=dev)
LIB.call("mmda_lstm_set_debug_buffer", _ptr(dbg))
eng.forward(*args, train=True)
torch.cuda.synchronize()
LIB.call("mmda_lstm_set_debug_buffer", None)
# last forward launch that wrote = arnn2 (smallest); we want the text one: re-run only text via H filter
d = dbg.cpu().view(64, 8)
print("NOTE: stamps are from the last lstm_forward launch of the step (acoustic rnn2)")
# text only: call the engine's text encoder directly
pk = eng._pack(b.lengths)
P = eng.params(); eng.k.bind_stream()
dbg.zero_()
LIB.call("mmda_lstm_set_debug_buffer", _ptr(dbg))
eng._encode("t", eng.saved["X"]["t"], pk, True, P)
torch.cuda.synchronize()
LIB.call("mmda_lstm_set_debug_buffer", None)
d = dbg.cpu().view(64, 8)[:50].double()
names = ["matvec+reduce", "gate math", "wait A", "dsmem stores+arrive B", "global stores", "wait B"]
ph = d[:, 1:7] - d[:, 0:6]
tot = d[1:, 0] - d[:-1, 0]
print("text rnn2 (last text launch), clocks per step: mean total %.0f" % tot[5:].mean())
for i, n in enumerate(names):
    print(f"  {n:24s} mean {ph[5:, i].mean():8.0f}  min {ph[5:, i].min():8.0f}  max {ph[5:, i].max():8.0f}")

# ---- backward (text rnn1 = last text backward launch) ----
from mmda_b200.trainer import FusedTrainer
tr = FusedTrainer(m)
lab = b.labels.to(dev)
tr.forward_backward(*args, lab)
dbg.zero_()
G = tr.G
du = eng.buf("dutt_t", 256, 1200)
eng.k.bind_stream()
LIB.call("mmda_lstm_set_debug_buffer", _ptr(dbg))
eng.multi_stream = False
eng._encode_backward("t", du, G, pk, eng.params())
torch.cuda.synchronize()
LIB.call("mmda_lstm_set_debug_buffer", None)
d = dbg.cpu().view(64, 8)[:50].double()
tot = d[1:, 0] - d[:-1, 0]
print("text BACKWARD (rnn1), clocks per step: mean total %.0f" % tot[5:].mean())
phases = {"matvec+scratch write": d[:, 2] - d[:, 0], "arrive + fetch issue": d[:, 1] - d[:, 2],
          "cluster wait": d[:, 3] - d[:, 1], "reduce+gates+stores": d[:, 4] - d[:, 3],
          "syncthreads": d[:, 5] - d[:, 4]}
for n, v in phases.items():
    print(f"  {n:24s} mean {v[5:].mean():8.0f}  min {v[5:].min():8.0f}  max {v[5:].max():8.0f}")
