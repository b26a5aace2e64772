import torch, sys
sys.path.insert(0, "/root/repo")
from mmda_b200._lib import LIB
dev = torch.device("cuda:0")
B, S, H, V = 6, 11, 768, 30522
word = torch.randn(V, H, device=dev); pos = torch.randn(512, H, device=dev); typ = torch.randn(2, H, device=dev)
ids = torch.randint(0, V, (B, S), device=dev); types = torch.zeros(B, S, dtype=torch.int64, device=dev)
out = torch.empty(B * S, H, device=dev)
st = torch.cuda.current_stream().cuda_stream
LIB.call("mmda_bert_embed_forward", word.data_ptr(), pos.data_ptr(), typ.data_ptr(), ids.data_ptr(), types.data_ptr(), B, S, H, V, 512, out.data_ptr(), st)
torch.cuda.synchronize()
ref = word[ids.view(-1)] + pos[:S].repeat(B, 1) + typ[0]
print("embed err", float((out - ref).abs().max()))
