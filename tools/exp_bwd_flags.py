import sys, torch
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import dev_lstm_tc as D
from mmda_b200._lib import LIB
dev = torch.device("cuda:0")
c = D.make_case(256, 300, 50, 0, dev)
o = D.run(c, "tc")
st = torch.cuda.current_stream().cuda_stream
B, H, N = c["B"], c["H"], c["N"]
nb = LIB.raw("mmda_lstm_tc_workspace_bytes")(B, H, c["Tmax"])
ws = torch.zeros(nb // 4 + 1, dtype=torch.int32, device=dev)
G = o["act"].clone()
def b_tc():
    LIB.call("mmda_lstm_tc_backward", D.P(G), D.P(c["whh"][0]), D.P(c["whh"][1]), D.P(o["c"]), D.P(c["dy"]),
             D.P(c["dutt"]), 4 * H, 0, 2 * H, D.P(c["lens"]), D.P(c["sidx"]), D.P(c["off"]), B, H, c["Tmax"], D.P(ws), st)
for fl in (0, 8, 16, 24, 31):
    LIB.call("mmda_lstm_tc_set_debug_flags", fl)
    print("flags", fl, "bwd ms", round(D.timeit(b_tc), 4), flush=True)
