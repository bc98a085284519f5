"""One launch of every hot kernel at a KNOWN shape of the Swin-T B=64 step, for `ncu --set full` captures whose DRAM
traffic can be set against the algorithmic bytes (bench.py `roofline.traffic`, profiles/ncu_traffic.json).

    python scripts/ncu_kernels.py <case>            # runs the case's kernel(s) 3 times (ncu: -k regex:<kernel> -s 2 -c 1)
    python scripts/ncu_kernels.py --list            # case name, kernel regex, algorithmic bytes / flops (JSON)

Cases use the stage-0 / stage-1 student shapes (the launches that carry the bytes): T0 = 696 320 tokens at C = 96."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

T0, C0 = 696320, 96
B_G, H_G = 128, 56     # stage 0 global-crop group: 128 maps of 56 x 56 tokens
K_OUT, ROWS_S, ROWS_T = 65536, 64 * (2 * 49 + 8 * 9), 64 * 2 * 49

CASES = {
    # name: (kernel regex, algorithmic bytes, algorithmic flops, note)
    "gemm_fwd_qkv0": ("gemm_kernel", 2 * (T0 * C0 + 3 * C0 * C0 + T0 * 3 * C0), 2 * T0 * C0 * 3 * C0, "qkv Linear stage 0, bias epilogue"),
    "gemm_gelu_fc1_0": ("gemm_kernel", 2 * (T0 * C0 + 4 * C0 * C0 + 2 * T0 * 4 * C0), 2 * T0 * C0 * 4 * C0, "fc1 + GELU (+gelu') stage 0"),
    "gemm_mul_fc2dgrad_0": ("gemm_kernel", 2 * (T0 * C0 + 4 * C0 * C0 + 2 * T0 * 4 * C0), 2 * T0 * C0 * 4 * C0, "fc2 dgrad * gelu' + colsum stage 0"),
    "gemm_wgrad_qkv0": ("gemm_kernel", 2 * (T0 * 3 * C0 + T0 * C0) + 4 * 3 * C0 * C0, 2 * T0 * C0 * 3 * C0, "qkv weight gradient stage 0"),
    "gemm_fwd_fc2_2": ("gemm_kernel", 2 * (43520 * 1536 + 384 * 1536 + 43520 * 384), 2 * 43520 * 1536 * 384, "fc2 stage 2 (compute-bound shape)"),
    "gemm_fwd_lastlayer": ("gemm_kernel", 2 * (ROWS_S * 256 + K_OUT * 256 + ROWS_S * K_OUT), 2 * ROWS_S * 256 * K_OUT, "DINOHead last layer, student region rows"),
    "attn_fwd7_s0": ("window_attn_fwd7", 8 * B_G * H_G * H_G * C0, 2 * 2 * 49 * 49 * 32 * 3 * B_G * 64, "ws 7 attention fwd, stage 0 global crops, shifted"),
    "attn_bwd7_s0": ("window_attn_bwd7", 16 * B_G * H_G * H_G * C0, 5 * 2 * 49 * 49 * 32 * 3 * B_G * 64, "ws 7 attention bwd, stage 0 global crops, shifted"),
    "add_ln_fwd_96": ("add_ln_fwd", 12 * T0 * C0, 0, "residual add + LN fwd, C = 96"),
    "add_ln_bwd_96": ("add_ln_bwd", 16 * T0 * C0, 0, "residual add + LN bwd, C = 96"),
    "dino_ce_fwd": ("dino_ce_q_fwd", 2 * K_OUT * (ROWS_S + ROWS_T), 0, "CE fwd, region rows"),
    "dino_ce_bwd": ("dino_ce_q_bwd", 2 * K_OUT * (2 * ROWS_S + ROWS_T), 0, "CE bwd, region rows"),
    "patch_embed_fwd": ("patch_embed_fwd", 128 * 3 * 224 * 224 * 4 + 128 * 3136 * 96 * 4, 2 * 48 * 96 * 128 * 3136, "PatchEmbed fwd, global crops"),
    "patch_embed_bwd": ("patch_embed_bwd", 128 * 3 * 224 * 224 * 4 + 128 * 3136 * 96 * 4, 4 * 48 * 96 * 128 * 3136, "PatchEmbed bwd, global crops"),
    "region_match": ("region_match", 4 * 768 * (ROWS_S + ROWS_T), 2 * 768 * 49 * 64 * 2 * (49 + 8 * 9 + 49), "cosine arg-max"),
}


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--list":
        print(json.dumps({k: {"kernel": v[0], "algorithmic_bytes": v[1], "algorithmic_flops": v[2], "note": v[3]} for k, v in CASES.items()}))
        return
    case = sys.argv[1]
    import torch

    from esvit_b200 import ops
    from esvit_b200.losses import DDINOLoss
    d = torch.device("cuda:0")
    BF = torch.bfloat16
    torch.manual_seed(0)
    reps = 3

    def rnd(*shape, s=0.5, dt=BF):
        return (torch.randn(*shape, device=d) * s).to(dt)

    if case == "gemm_fwd_qkv0":
        a, w, b = rnd(T0, C0), rnd(3 * C0, C0, s=0.1), torch.randn(3 * C0, device=d)
        f = lambda: ops.gemm(a, w, b)
    elif case == "gemm_gelu_fc1_0":
        a, w, b = rnd(T0, C0), rnd(4 * C0, C0, s=0.1), torch.randn(4 * C0, device=d)
        f = lambda: ops.gemm(a, w, b, act=1, want_pre=True)
    elif case == "gemm_mul_fc2dgrad_0":
        a, w, m, cs = rnd(T0, C0), rnd(C0, 4 * C0, s=0.1), rnd(T0, 4 * C0), torch.zeros(4 * C0, device=d)
        f = lambda: ops.gemm_mul_colsum(a, w, m, cs, b_mn=True)
    elif case == "gemm_wgrad_qkv0":
        dy, x = rnd(T0, 3 * C0), rnd(T0, C0)
        f = lambda: ops.gemm_wgrad(dy, x)
    elif case == "gemm_fwd_fc2_2":
        a, w, b = rnd(43520, 1536), rnd(384, 1536, s=0.05), torch.randn(384, device=d)
        f = lambda: ops.gemm(a, w, b)
    elif case == "gemm_fwd_lastlayer":
        a, w = rnd(ROWS_S, 256, s=0.06), rnd(K_OUT, 256, s=0.06)
        f = lambda: ops.gemm(a, w, None)
    elif case in ("attn_fwd7_s0", "attn_bwd7_s0"):
        qkv = rnd(B_G, H_G * H_G, 3 * C0).requires_grad_(True)
        bias, table, go = torch.randn(3 * C0, device=d) * 0.1, torch.randn(169, 3, device=d) * 0.2, rnd(B_G, H_G * H_G, C0)
        fw = lambda: ops.WindowAttentionFn.apply(qkv, bias, table, H_G, H_G, 3, 7, 3, 32 ** -0.5, None)
        if case == "attn_fwd7_s0":
            f = fw
        else:
            out = fw()
            f = lambda: out.backward(go, retain_graph=True)
    elif case in ("add_ln_fwd_96", "add_ln_bwd_96"):
        x = torch.randn(64, T0 // 64, C0, device=d).requires_grad_(True)
        dl = rnd(64, T0 // 64, C0).requires_grad_(True)
        g, b = torch.ones(C0, device=d, requires_grad=True), torch.zeros(C0, device=d, requires_grad=True)
        fw = lambda: ops.add_layer_norm(x, dl, None, g, b, 1e-6)
        if case == "add_ln_fwd_96":
            f = fw
        else:
            xo, y = fw()
            gx, gy = torch.randn_like(xo), rnd(*y.shape)
            f = lambda: torch.autograd.backward([xo, y], [gx, gy], retain_graph=True)
    elif case in ("dino_ce_fwd", "dino_ce_bwd"):
        B, ncrops = 64, 10
        mod = DDINOLoss(K_OUT, ncrops, 0.04, 0.04, 0, 10).to(d)
        s_cls, s_reg = rnd(ncrops * B, K_OUT).requires_grad_(True), rnd(ROWS_S, K_OUT).requires_grad_(True)
        t_cls, t_reg = rnd(2 * B, K_OUT), rnd(ROWS_T, K_OUT)
        s_fea, t_fea = torch.randn(ROWS_S, 768, device=d), torch.randn(ROWS_T, 768, device=d)
        fw = lambda: mod((s_cls, s_reg, s_fea, [49, 9]), (t_cls, t_reg, t_fea, [49]), 0, None)
        if case == "dino_ce_fwd":
            f = fw
        else:
            l = fw()
            f = lambda: l.backward(retain_graph=True)
    elif case in ("patch_embed_fwd", "patch_embed_bwd"):
        img = torch.randn(128, 3, 224, 224, device=d)
        w = (torch.randn(96, 3, 4, 4, device=d) * 0.1).requires_grad_(True)
        pb, g, b = [torch.randn(96, device=d).requires_grad_(True) for _ in range(3)]
        fw = lambda: ops.PatchEmbedFn.apply(img, w, pb, g, b, 1e-6)
        if case == "patch_embed_fwd":
            f = fw
        else:
            y = fw()
            gy = torch.randn_like(y)
            f = lambda: y.backward(gy, retain_graph=True)
    elif case == "region_match":
        s_fea, t_fea = torch.randn(ROWS_S, 768, device=d), torch.randn(ROWS_T, 768, device=d)
        f = lambda: ops.region_match(s_fea, t_fea, 64, 10, 49, 9)
    else:
        raise SystemExit("unknown case " + case)
    for _ in range(reps):
        f()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
