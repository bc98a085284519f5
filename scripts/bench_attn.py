"""Micro-benchmark of the window-attention kernels at the Swin-T stage shapes (B=64: 128 global + 512 local images).
ESVIT_ATTN_DBG=1: gathers disabled after the first window per CTA (math only); =2: math disabled (gathers only)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esvit_b200 import ops  # noqa: E402

d = torch.device("cuda:0")
shapes = [(128, 56, 96, 3), (512, 24, 96, 3), (128, 28, 192, 6), (512, 12, 192, 6), (128, 14, 384, 12), (512, 6, 384, 12),
          (128, 7, 768, 24), (512, 3, 768, 24)]
for dbg in (("0",) if os.environ.get("ESVIT_ATTN_TC") or os.environ.get("ESVIT_ATTN_ONLY0") else ("0", "1", "2")):
    os.environ["ESVIT_ATTN_DBG"] = dbg
    tot_f = tot_b = 0.0
    for (B, H, C, nH) in shapes:
        for shift in (0, 3 if H > 7 or H in (6, 12, 24) else 0):
            qkv = torch.randn(B, H * H, 3 * C, device=d).to(torch.bfloat16).requires_grad_(True)
            bias = torch.randn(3 * C, device=d) * 0.1
            table = torch.randn(169, nH, device=d) * 0.2
            go = torch.randn(B, H * H, C, device=d).to(torch.bfloat16)
            f = lambda: ops.WindowAttentionFn.apply(qkv, bias, table, H, H, nH, 7, shift, 32 ** -0.5, None)
            out = f()
            out.backward(go)
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            for _ in range(5):
                out = f()
            e[1].record()
            for _ in range(5):
                out.backward(go, retain_graph=True)
            e[2].record()
            torch.cuda.synchronize()
            tf, tb = e[0].elapsed_time(e[1]) / 5, e[1].elapsed_time(e[2]) / 5
            tot_f += tf
            tot_b += tb
            if True:
                print(f"dbg={dbg} " +f"B={B} H={H} C={C} nH={nH} shift={shift}: fwd {tf*1e3:.0f} us  bwd {tb*1e3:.0f} us  "
                      f"fwd {B*H*H*C*8/tf/1e6:.0f} GB/s  bwd {B*H*H*C*16/tb/1e6:.0f} GB/s")
    print(f"ESVIT_ATTN_DBG={dbg}: sum fwd {tot_f:.3f} ms  sum bwd {tot_b:.3f} ms")
