"""Host-side autograd plumbing of the module mirror, exercised WITHOUT a GPU: every C-ABI call is replaced by a no-op
(outputs stay uninitialised), so only the wiring is checked - fused pending residuals, the x-less add+LN after
PatchMerging, the per-pass DropPath plan, the shared gradient accumulators of the two crop groups, MlpFn.  The numerics
of the same paths are the -m gpu tests."""
import torch

from esvit_b200 import _lib, engine, ops


def _patch(monkeypatch):
    monkeypatch.setattr(_lib, "call", lambda name, *a: None)
    monkeypatch.setattr(ops, "_stream", lambda: 0)

    def chk(t, dtype, name):
        if t is None:
            return None
        assert t.dtype == dtype, (name, t.dtype, dtype)
        return t if t.is_contiguous() else t.contiguous()

    def gemm_bias_act(a, w, bias, act=0, want_pre=False):
        out = torch.zeros(*a.shape[:-1], w.shape[0], dtype=torch.bfloat16)
        return (out, torch.zeros_like(out)) if (act and want_pre) else out

    monkeypatch.setattr(ops, "_chk", chk)
    monkeypatch.setattr(ops, "gemm_bias_act", gemm_bias_act)


import pytest


@pytest.mark.parametrize("fuse_groups", [False, True])
def test_every_parameter_gets_one_gradient(monkeypatch, fuse_groups):
    from esvit_b200 import swin_transformer
    _patch(monkeypatch)
    monkeypatch.setattr(swin_transformer, "USE_FUSED_GROUPS", fuse_groups)
    spec = dict(engine.SWIN_SPECS["swin_tiny_w7"])
    spec["depths"] = [1, 1, 2, 1]
    torch.manual_seed(0)
    net = engine.build_network(spec, 256, True, False, True, 224, None)
    net.train()
    B = 2
    crops = [torch.randn(B, 3, 224, 224) for _ in range(2)] + [torch.randn(B, 3, 96, 96) for _ in range(3)]
    cls, region, fea, npatch = net(crops)
    assert cls.shape == (5 * B, 256) and npatch == [49, 9]
    assert region.shape == (B * (2 * 49 + 3 * 9), 256) and fea.shape == (B * (2 * 49 + 3 * 9), 768)
    loss = (cls.float() ** 2).sum() + (region.float() ** 2).sum()
    ops.begin_step("cpu", 1 << 22)
    try:
        loss.backward()
        n_shared = len(ops._Arena.accs)
    finally:
        ops.end_step()
    assert ops._Arena.accs is None
    assert n_shared > 20  # LN / bias / rel-pos-table accumulators (shared by the two crop groups when run per group)
    missing = [n for n, p in net.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing
    for n, p in net.named_parameters():
        if p.grad is not None:
            assert p.grad.shape == p.shape and p.grad.dtype == torch.float32, n


def test_accumulators_are_private_outside_a_step(monkeypatch):
    _patch(monkeypatch)
    a, first_a = ops._acc(("k", 1), (4,), "cpu")
    b, first_b = ops._acc(("k", 1), (4,), "cpu")
    assert first_a and first_b and a.data_ptr() != b.data_ptr()
    ops.begin_step("cpu", 1 << 10)
    try:
        a, first_a = ops._acc(("k", 1), (4,), "cpu")
        b, first_b = ops._acc(("k", 1), (4,), "cpu")
        assert first_a and not first_b and a.data_ptr() == b.data_ptr()
    finally:
        ops.end_step()


def test_group_geometry_of_the_concatenated_layout():
    """rows -> samples map and the per-stage geometry the group-aware Functions receive (pure host logic)."""
    spec = dict(engine.SWIN_SPECS["swin_tiny_w7"])
    spec["depths"] = [1, 1, 1, 1]
    net = engine.build_network(spec, 64, True, False, True, 224, None)
    grp = [(2, 56, 56, 0), (3, 24, 24, 2 * 56 * 56)]
    rs = net._row_samples(grp, torch.device("cpu"))
    assert rs.numel() == 2 * 56 * 56 + 3 * 24 * 24
    assert rs[0] == 0 and rs[56 * 56 - 1] == 0 and rs[56 * 56] == 1 and rs[2 * 56 * 56] == 2 and rs[-1] == 4
    assert net._row_samples(grp, torch.device("cpu")) is rs  # cached per geometry
    merged = []
    row0 = 0
    for B, H, W, _ in grp:  # what PatchMerging.fused_groups hands to the next stage
        merged.append((B, (H + 1) // 2, (W + 1) // 2, row0))
        row0 += B * ((H + 1) // 2) * ((W + 1) // 2)
    assert merged == [(2, 28, 28, 0), (3, 12, 12, 2 * 28 * 28)]


def test_loss_row_order_is_an_image_major_permutation():
    """losses._order: the CTA -> student-row map of the CE kernels is a permutation of the (crop, image, token) storage
    order that visits all rows of image 0, then image 1, ... (so the CTAs resident together share teacher rows)."""
    from esvit_b200.losses import DDINOLoss
    B, ncrops, Tg, Tl = 3, 5, 4, 2
    m = DDINOLoss(64, ncrops, 0.04, 0.04, 0, 10)
    o = m._order(B, [(2, Tg), (ncrops - 2, Tl)], "cpu").tolist()
    R = B * (2 * Tg + (ncrops - 2) * Tl)
    assert sorted(o) == list(range(R))

    def image_of(r):
        if r < 2 * B * Tg:
            return (r // Tg) % B
        return ((r - 2 * B * Tg) // Tl) % B
    imgs = [image_of(r) for r in o]
    assert imgs == sorted(imgs)                      # image-major
    per = 2 * Tg + (ncrops - 2) * Tl
    assert all(imgs[i * per] == i for i in range(B))
    oc = m._order(B, [(ncrops, 1)], "cpu").tolist()  # cls rows: (crop, image) -> (image, crop)
    assert oc == [v * B + b for b in range(B) for v in range(ncrops)]


def test_cat_adjacent_views_back_to_back_crops():
    """ops.cat_adjacent == torch.cat; a view (no copy) exactly when the tensors lie back to back in one storage."""
    import torch
    from esvit_b200 import ops
    buf = torch.randn(3, 4, 3, 8, 8)
    ts = [buf[k] for k in range(3)]
    v = ops.cat_adjacent(ts)
    assert torch.equal(v, torch.cat(ts)) and v.data_ptr() == buf.data_ptr()
    gap = ops.cat_adjacent([buf[0], buf[2]])                       # not adjacent: a real concatenation
    assert torch.equal(gap, torch.cat([buf[0], buf[2]])) and gap.data_ptr() != buf.data_ptr()
    sep = [torch.randn(4, 3, 8, 8) for _ in range(2)]               # separate allocations
    assert torch.equal(ops.cat_adjacent(sep), torch.cat(sep))
    assert ops.cat_adjacent([buf[1]]) is not None and ops.cat_adjacent([buf[1]]).shape == buf[1].shape
