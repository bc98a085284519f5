"""tcgen05 / TMA GEMM with fused bias (+GELU) epilogue against torch (bf16 inputs, fp32 reference math)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import assert_close

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


@pytest.mark.parametrize("M,K,N", [(4096, 96, 384), (1000, 192, 768), (300, 384, 1536), (256, 768, 3072),
                                   (512, 3072, 768), (777, 64, 96), (128, 128, 288), (33, 96, 288), (20000, 96, 96)])
@pytest.mark.parametrize("act", [0, 1])
def test_gemm_bias_act(M, K, N, act):
    from esvit_b200 import ops
    torch.manual_seed(M + K + N)
    d = torch.device("cuda:0")
    a = (torch.randn(M, K, device=d) * 0.5).to(BF16)
    w = (torch.randn(N, K, device=d) / K ** 0.5).to(BF16)
    b = torch.randn(N, device=d) * 0.2
    ref_pre = a.float() @ w.float().t() + b
    ref = F.gelu(ref_pre) if act else ref_pre
    if act:
        out, gp = ops.gemm_bias_act(a, w, b, act=1, want_pre=True)
        xr = ref_pre.clone().requires_grad_(True)
        F.gelu(xr).sum().backward()
        assert_close(gp, xr.grad, 5e-3, "gelu'(pre-activation)")
    else:
        out = ops.gemm_bias_act(a, w, b, act=0)
    torch.cuda.synchronize()
    assert_close(out, ref, 5e-3, "out")
    out2 = ops.gemm_bias_act(a, w, None, act=0)
    assert_close(out2, ref_pre - b, 5e-3, "no bias")


def test_gemm_speed_vs_library(capsys):
    """prints achieved throughput next to the cuBLASLt GEMM + separate GELU kernel it replaces (informational)."""
    from esvit_b200 import ops
    d = torch.device("cuda:0")
    rows = []
    for (M, K, N) in [(696320, 96, 384), (174080, 192, 768), (43520, 384, 1536), (10880, 768, 3072), (43520, 1536, 384)]:
        a = (torch.randn(M, K, device=d) * 0.5).to(BF16)
        w = (torch.randn(N, K, device=d) / K ** 0.5).to(BF16)
        b = torch.randn(N, device=d) * 0.2
        bb = b.to(BF16)

        def mine():
            return ops.gemm_bias_act(a, w, b, act=1, want_pre=True)

        def lib():
            h = F.linear(a, w, bb)
            return ops.GeluFn.apply(h)

        res = []
        for f in (mine, lib):
            for _ in range(3):
                f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                f()
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 10)
        rows.append((M, K, N, res[0], res[1], 2.0 * M * K * N / res[0] / 1e9))
    with capsys.disabled():
        for r in rows:
            print("tcgen05 gemm+bias+gelu M=%d K=%d N=%d: %.3f ms (library gemm + gelu kernel %.3f ms) %.0f TFLOP/s" % r)
