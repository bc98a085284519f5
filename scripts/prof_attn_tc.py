"""Per-role cycle accounting of the tcgen05 attention backward (ESVIT_ATTN_TC=2 ESVIT_ATTN_PROF=1): one launch at the
stage-0 global-crop shape; CTA (0, 0) prints where each warp role spent its cycles."""
import os
import sys

import torch

os.environ.setdefault("ESVIT_ATTN_TC", "2")
os.environ.setdefault("ESVIT_ATTN_PROF", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvit_b200 import ops  # noqa: E402

d = torch.device("cuda:0")
for (B, H, C, nH, shift) in [(128, 56, 96, 3, 3), (512, 24, 96, 3, 0)]:
    qkv = torch.randn(B, H * H, 3 * C, device=d).to(torch.bfloat16).requires_grad_(True)
    bias = torch.randn(3 * C, device=d) * 0.1
    table = torch.randn(169, nH, device=d) * 0.2
    go = torch.randn(B, H * H, C, device=d).to(torch.bfloat16)
    for rep in range(2):
        out = ops.WindowAttentionFn.apply(qkv, bias, table, H, H, nH, 7, shift, 32 ** -0.5, None)
        torch.cuda.synchronize()
        print(f"--- B={B} H={H} shift={shift} rep {rep}", flush=True)
        out.backward(go)
        torch.cuda.synchronize()
