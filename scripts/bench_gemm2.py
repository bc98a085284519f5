"""GPU micro-benchmark of the second-generation tcgen05 GEMM family on the REAL shapes of the Swin-T B=64 step, every
tile configuration next to the library GEMM (torch / cuBLASLt) it replaces.  Prints one line per (shape, kind)."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from esvit_b200 import ops

BF16 = torch.bfloat16
d = torch.device("cuda:0")
TILES = [1128, 1256, 2128, 2256]


def timeit(f, n=10, warm=3):
    for _ in range(warm):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def line(kind, M, N, K, lib_ms, mine):
    fl = 2.0 * M * N * K
    best = min(mine.items(), key=lambda kv: kv[1])
    s = " ".join(f"{t}:{ms:.3f}" for t, ms in mine.items())
    print(f"{kind:10s} M={M:7d} N={N:6d} K={K:6d} lib {lib_ms:.3f} ms ({fl / lib_ms / 1e9:6.0f} TF/s) | {s} | best {best[0]} "
          f"{best[1]:.3f} ms ({fl / best[1] / 1e9:6.0f} TF/s) x{lib_ms / best[1]:.2f}", flush=True)


def main():
    T0 = 696320
    shapes = []
    for st, C in enumerate((96, 192, 384, 768)):
        T = T0 // 4 ** st
        shapes += [("fwd", T, 3 * C, C), ("fwd", T, C, C), ("gelu", T, 4 * C, C), ("fwd", T, C, 4 * C),
                   ("dgrad", T, C, 3 * C), ("dgrad", T, C, 4 * C), ("mul", T, 4 * C, C),
                   ("wgrad", T, 3 * C, C), ("wgrad", T, 4 * C, C), ("wgrad", T, C, 4 * C)]
        if st < 3:
            shapes += [("fwd", T // 4, 2 * C, 4 * C)]
    shapes += [("gelu", 11520, 2048, 768), ("gelu", 11520, 2048, 2048), ("fwd", 11520, 256, 2048), ("fwd", 11520, 65536, 256),
               ("fwd", 6400, 65536, 256), ("dgrad", 11520, 256, 65536), ("wgrad", 11520, 65536, 256)]
    only = sys.argv[1:] or None
    for kind, M, N, K in shapes:
        if only and kind not in only:
            continue
        torch.manual_seed(0)
        if kind in ("fwd", "gelu"):
            a = (torch.randn(M, K, device=d) * 0.5).to(BF16)
            w = (torch.randn(N, K, device=d) / K ** 0.5).to(BF16)
            b = torch.randn(N, device=d) * 0.2
            bb = b.to(BF16)
            if kind == "fwd":
                lib = timeit(lambda: F.linear(a, w, bb))
                mine = {t: timeit(lambda: ops.gemm(a, w, b, tile=t)) for t in TILES}
            else:
                lib = timeit(lambda: ops.GeluFn.apply(F.linear(a, w, bb)))
                mine = {t: timeit(lambda: ops.gemm(a, w, b, act=1, want_pre=True, tile=t)) for t in TILES}
        elif kind == "dgrad":   # dx[M,N] = dy[M,K] @ W[K,N]
            dy = (torch.randn(M, K, device=d) * 0.5).to(BF16)
            w = (torch.randn(K, N, device=d) / K ** 0.5).to(BF16)
            lib = timeit(lambda: dy @ w)
            mine = {t: timeit(lambda: ops.gemm(dy, w, None, b_mn=True, tile=t)) for t in TILES}
        elif kind == "mul":     # d(pre)[M,N] = (dy[M,K] @ W2[K,N]) * gelu'
            dy = (torch.randn(M, K, device=d) * 0.5).to(BF16)
            w = (torch.randn(K, N, device=d) / K ** 0.5).to(BF16)
            mult = torch.rand(M, N, device=d).to(BF16)
            cs = torch.zeros(N, device=d)
            lib = timeit(lambda: (dy @ w) * mult)
            mine = {t: timeit(lambda: ops.gemm_mul_colsum(dy, w, mult, cs, b_mn=True, tile=t)) for t in TILES}
        else:                   # wgrad dw[N,K] = dy[M,N]^T @ x[M,K]   (here M = tokens)
            dy = (torch.randn(M, N, device=d) * 0.5).to(BF16)
            x = (torch.randn(M, K, device=d) * 0.5).to(BF16)
            lib = timeit(lambda: (dy.t() @ x).float())
            mine = {t: timeit(lambda: ops.gemm_wgrad(dy, x, tile=t)) for t in TILES}
            if M <= 50000:  # forced split counts (tile + 10000 * splits): where does the time go for the short-K shapes?
                for sp in (1, 2, 4, 8):
                    mine[f"2256/s{sp}"] = timeit(lambda: ops.gemm_wgrad(dy, x, tile=2256 + 10000 * sp))
        line(kind, M, N, K, lib, mine)
        del mine
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
