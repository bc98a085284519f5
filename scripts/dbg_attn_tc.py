import os, sys, torch
os.environ.setdefault("ESVIT_ATTN_TC", "2")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvit_b200 import ops
d = torch.device("cuda:0")
B, H, C, nH, shift = [int(v) for v in os.environ.get("CASE", "2,14,96,3,0").split(",")]
qkv = torch.randn(B, H * H, 3 * C, device=d).to(torch.bfloat16).requires_grad_(True)
bias = torch.randn(3 * C, device=d) * 0.1
table = torch.randn(169, nH, device=d) * 0.2
go = torch.randn(B, H * H, C, device=d).to(torch.bfloat16)
out = ops.WindowAttentionFn.apply(qkv, bias, table, H, H, nH, 7, shift, 32 ** -0.5, None)
torch.cuda.synchronize()
print("fwd ok", flush=True)
out.backward(go)
torch.cuda.synchronize()
print("bwd ok", qkv.grad.float().abs().mean().item())
