"""world_size-2 `gloo` tests (CPU) of the N>1 host logic: sharding by rank, the packed center all-reduce and its
rows*world scaling, gradient averaging, and max-over-ranks timing - checked against the single-process oracle on the
concatenated batch (world-size invariance)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import losses as L
    K, rows = 64, 6
    g = torch.Generator().manual_seed(5)
    t_cls_full = torch.randn(world * rows, K, generator=g)
    t_reg_full = torch.randn(world * rows * 3, K, generator=g)
    t_cls = t_cls_full[rank * rows:(rank + 1) * rows]
    t_reg = t_reg_full[rank * rows * 3:(rank + 1) * rows * 3]

    # (1) the product's packed center reduction, with the CUDA EMA kernel stubbed by its arithmetic definition
    from esvit_b200 import losses as PL
    from esvit_b200 import ops

    def center_ema_cpu(center, colsum_total, rows_total, momentum, out=None):
        res = center * momentum + (colsum_total / rows_total) * (1 - momentum)
        if out is not None:
            out.copy_(res)
            return out
        return res

    ops.center_ema = center_ema_cpu
    mod = PL.DDINOLoss(K, 4, 0.04, 0.04, 0, 10)
    mod.center.normal_(generator=torch.Generator().manual_seed(1))
    mod.center_grid.normal_(generator=torch.Generator().manual_seed(2))
    c0, cg0 = mod.center.clone(), mod.center_grid.clone()
    sums = torch.stack([t_cls.sum(0), t_reg.sum(0)])
    mod._reduce_and_update(sums, [t_cls.shape[0], t_reg.shape[0]], ["center", "center_grid"])
    ref_c = L.center_update(c0, t_cls_full, 0.9)       # single process, concatenated batch
    ref_g = L.center_update(cg0, t_reg_full, 0.9)
    assert torch.allclose(mod.center, ref_c, atol=1e-6), "center not world-size invariant"
    assert torch.allclose(mod.center_grid, ref_g, atol=1e-6)

    # (2) oracle-level invariance through the all_reduce hook (what DDP AVG + dist.all_reduce do)
    c_rank = L.center_update(c0, t_cls, 0.9, world, lambda x: dist.all_reduce(x))
    assert torch.allclose(c_rank, ref_c, atol=1e-6)

    # (3) gradient averaging == gradient of the mean loss over the global batch
    w = torch.randn(K, generator=torch.Generator().manual_seed(3), requires_grad=True)
    (t_cls @ w).pow(2).mean().backward()
    gr = w.grad.clone()
    dist.all_reduce(gr)
    gr /= world
    w2 = w.detach().clone().requires_grad_(True)
    (t_cls_full @ w2).pow(2).mean().backward()
    assert torch.allclose(gr, w2.grad, atol=1e-5)

    # (4) bench.py helpers: per-rank crops differ, timing is the max over ranks
    import bench
    a = bench.synthetic_crops(1, 1, rank)[0]
    other = [torch.empty_like(a) for _ in range(world)]
    dist.all_gather(other, a)
    assert not torch.equal(other[0], other[1])
    ms = bench.max_over_ranks(10.0 + rank, torch.device("cpu"))
    assert ms == 10.0 + world - 1
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


def test_world_size_two_host_logic(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
