"""Micro-benchmark of the ws=14 window-attention kernels at the Swin-B W14 stage shapes (B=32: 64 global + 256 local
images).  (profiles/r01_v9_attn14_microbench.txt holds the A/B against the first, generic ws-templated kernels.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvit_b200 import ops  # noqa: E402

d = torch.device("cuda:0")
# (images, H, C, heads, shift, blocks of this kind per forward)
shapes = [(64, 56, 128, 4, 0, 1), (64, 56, 128, 4, 7, 1), (64, 28, 256, 8, 0, 1), (64, 28, 256, 8, 7, 1),
          (64, 14, 512, 16, 0, 18), (256, 24, 128, 4, 0, 1), (256, 24, 128, 4, 7, 1), (256, 12, 256, 8, 0, 1),
          (256, 12, 256, 8, 7, 1), (256, 6, 512, 16, 0, 18)]
for generic in ("0",):
    tot_f = tot_b = 0.0
    for (B, H, C, nH, shift, nblk) in shapes:
        qkv = torch.randn(B, H * H, 3 * C, device=d).to(torch.bfloat16).requires_grad_(True)
        bias = torch.randn(3 * C, device=d) * 0.1
        table = torch.randn(27 * 27, nH, device=d) * 0.2
        go = torch.randn(B, H * H, C, device=d).to(torch.bfloat16)
        f = lambda: ops.WindowAttentionFn.apply(qkv, bias, table, H, H, nH, 14, shift, 32 ** -0.5, None)
        out = f()
        out.backward(go)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for _ in range(3):
            out = f()
        e[1].record()
        for _ in range(3):
            out.backward(go, retain_graph=True)
        e[2].record()
        torch.cuda.synchronize()
        tf, tb = e[0].elapsed_time(e[1]) / 3, e[1].elapsed_time(e[2]) / 3
        tot_f += tf * nblk
        tot_b += tb * nblk
        print(f"B={B} H={H} C={C} nH={nH} shift={shift}: fwd {tf*1e3:.0f} us  bwd {tb*1e3:.0f} us", flush=True)
    print(f"weighted sum fwd {tot_f:.3f} ms  bwd {tot_b:.3f} ms (student pass, Swin-B W14, B=32)", flush=True)
