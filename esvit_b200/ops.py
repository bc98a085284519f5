"""torch.autograd.Function wrappers over the esvit_b200 C ABI (one entry point per kernel).

PyTorch is used here for device memory (caching allocator), the current CUDA stream and autograd bookkeeping;
all arithmetic of these ops happens in the hand-written sm_100a kernels.  There is no fallback path: every op
requires CUDA tensors and the built library.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib, shadow

Tensor = torch.Tensor
BF16 = torch.bfloat16
F32 = torch.float32


def _p(t: Optional[Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _po(t: Tensor, off_elems: int):
    """pointer to element `off_elems` of a contiguous tensor (one resolution group inside a concatenated buffer)"""
    return ctypes.c_void_p(t.data_ptr() + off_elems * t.element_size())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Arena:
    """One zero-filled fp32 buffer per training step from which the backward kernels' small ACCUMULATED outputs
    (dgamma / dbeta / bias gradients / rel-pos-table gradients) are carved: one memset per step instead of ~340 tiny
    fill kernels.  Regions are never handed out twice, so a stale arena is still correct (just smaller)."""
    buf: Optional[Tensor] = None
    off = 0
    accs: Optional[dict] = None  # live only between begin_step() and end_step(): parameter -> its accumulator
    warned = False


def begin_step(device, nfloats: int = 1 << 22) -> None:
    """Call right before ONE backward() (engine.SelfDistillStep does), end_step() right after; optional - without it
    ops fall back to torch.zeros per accumulator."""
    _Arena.buf = torch.zeros(nfloats, dtype=F32, device=device)
    _Arena.off = 0
    _Arena.accs = {}


def end_step() -> None:
    _Arena.accs = None


def _acc(key, shape, device) -> Tuple[Tensor, bool]:
    """Accumulated-gradient buffer of one parameter for THIS backward pass -> (buffer, first).  The backbone runs once per
    resolution group (global / local crops) through the same weights; both backward kernels += into the same
    zero-filled buffer and only the first call hands it to autograd (later calls return None for that input), so no
    separate gradient-accumulation kernels run.  Valid because autograd holds the first tensor by reference until every
    contribution to the leaf has been produced, and the kernels are ordered on one stream.  Outside
    begin_step()/end_step() every call gets a fresh buffer."""
    d = _Arena.accs
    if d is None:
        return _zeros(shape, device), True
    t = d.get(key)
    if t is not None and tuple(t.shape) == tuple(shape):
        return t, False
    t = d[key] = _zeros(shape, device)
    return t.view(t.shape), True  # a fresh alias: autograd adopts it as .grad instead of cloning (sole reference)


def _zeros(shape, device) -> Tensor:
    n = 1
    for d in shape:
        n *= int(d)
    buf = _Arena.buf
    n16 = (n + 3) & ~3  # keep every region 16-byte aligned (vectorised optimiser reads)
    if buf is not None and buf.device == torch.device(device) and _Arena.off + n16 <= buf.numel():
        v = buf[_Arena.off:_Arena.off + n].view(*shape)
        _Arena.off += n16
        return v
    if buf is not None and not _Arena.warned:
        _Arena.warned = True
        import warnings
        warnings.warn("esvit_b200: gradient-accumulator arena exhausted (%d floats); falling back to per-tensor torch.zeros - "
                      "pass a larger nfloats to ops.begin_step" % buf.numel())
    return torch.zeros(*shape, dtype=F32, device=device)


def _chk(t: Optional[Tensor], dtype, name: str) -> Optional[Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"esvit_b200 op input `{name}` must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise TypeError(f"`{name}` must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------------------
# residual add + LayerNorm


def _add_ln_fwd(x, delta, keep, tps, gamma, beta, eps, want_y, y_bf16):
    ref = x if x is not None else delta  # x = None: xout = fp32(delta), the start of a residual stream
    T, C = ref.numel() // ref.shape[-1], ref.shape[-1]
    xout = torch.empty(ref.shape, dtype=F32, device=ref.device) if delta is not None else None
    y = mean = rstd = None
    if want_y:
        y = torch.empty(ref.shape, dtype=BF16 if y_bf16 else F32, device=ref.device)
        mean = torch.empty(T, dtype=F32, device=ref.device)
        rstd = torch.empty(T, dtype=F32, device=ref.device)
    _lib.call("esvit_add_ln_fwd", _p(x), _p(delta), _p(keep), tps, _p(gamma), _p(beta), eps, _p(xout), _p(y),
              1 if y_bf16 else 0, _p(mean), _p(rstd), T, C, _stream())
    return (xout if delta is not None else x), y, mean, rstd


class AddLayerNormFn(Function):
    """(x, delta, delta_bias, keep) -> (xout = x + keep*delta, y = LN(xout)).
    keep: per-sample DropPath scale or None.  delta_bias (fp32 [C] parameter or None) is the bias the producing GEMM
    already added to delta in its epilogue: it is only routed here so that ITS GRADIENT (column sums of ddelta) comes
    out of this backward kernel instead of a separate reduction."""

    @staticmethod
    def forward(ctx, x, delta, dbias, keep, gamma, beta, eps: float, y_bf16: bool):
        x = _chk(x, F32, "x")
        delta = _chk(delta, BF16, "delta")
        dbias = _chk(dbias, F32, "delta_bias")
        keep = _chk(keep, F32, "keep")
        gamma, beta = _chk(gamma, F32, "gamma"), _chk(beta, F32, "beta")
        ref = x if x is not None else delta  # x = None: the stream starts here as fp32(delta) (after PatchMerging)
        tps = ref.numel() // ref.shape[-1] // ref.shape[0]
        xout, y, mean, rstd = _add_ln_fwd(x, delta, keep, tps, gamma, beta, eps, True, y_bf16)
        ctx.save_for_backward(xout, mean, rstd, gamma, keep)
        ctx.tps, ctx.y_bf16, ctx.has_dbias, ctx.has_x = tps, y_bf16, dbias is not None, x is not None
        return xout, y

    @staticmethod
    @once_differentiable
    def backward(ctx, g_xout, g_y):
        xout, mean, rstd, gamma, keep = ctx.saved_tensors
        T, C = xout.numel() // xout.shape[-1], xout.shape[-1]
        g_xout = _chk(g_xout, F32, "g_xout") if g_xout is not None else None
        if g_y is not None:
            g_y = _chk(g_y, BF16 if ctx.y_bf16 else F32, "g_y")
        dx = torch.empty_like(xout) if ctx.has_x else None
        ddelta = torch.empty(xout.shape, dtype=BF16, device=xout.device)
        acc, first = _acc(("add_ln", gamma.data_ptr()), (3, C), xout.device)  # dgamma | dbeta | ddelta_bias
        _lib.call("esvit_add_ln_bwd", _p(g_y), 1 if ctx.y_bf16 else 0, _p(g_xout), _p(xout), _p(mean), _p(rstd),
                  _p(gamma), _p(keep), ctx.tps, _p(dx), _p(ddelta), _p(acc[0]), _p(acc[1]),
                  _p(acc[2]) if ctx.has_dbias else None, T, C, _stream())
        if not first:
            return dx, ddelta, None, None, None, None, None, None
        return dx, ddelta, (acc[2] if ctx.has_dbias else None), None, acc[0], acc[1], None, None


class LayerNormFn(Function):
    """x fp32 -> y = LN(x) (bf16 or fp32)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float, y_bf16: bool):
        x = _chk(x, F32, "x")
        gamma, beta = _chk(gamma, F32, "gamma"), _chk(beta, F32, "beta")
        _, y, mean, rstd = _add_ln_fwd(x, None, None, 1, gamma, beta, eps, True, y_bf16)
        ctx.save_for_backward(x, mean, rstd, gamma)
        ctx.y_bf16 = y_bf16
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g_y):
        x, mean, rstd, gamma = ctx.saved_tensors
        T, C = x.numel() // x.shape[-1], x.shape[-1]
        g_y = _chk(g_y, BF16 if ctx.y_bf16 else F32, "g_y")
        dx = torch.empty_like(x)
        acc, first = _acc(("ln", gamma.data_ptr()), (2, C), x.device)
        _lib.call("esvit_add_ln_bwd", _p(g_y), 1 if ctx.y_bf16 else 0, None, _p(x), _p(mean), _p(rstd), _p(gamma),
                  None, 1, _p(dx), None, _p(acc[0]), _p(acc[1]), None, T, C, _stream())
        return (dx, acc[0], acc[1], None, None) if first else (dx, None, None, None, None)


class LayerNormResidFn(Function):
    """x fp32 -> (x, y = LN(x)): the start of a residual stream whose input is also the shortcut (the first block after
    PatchEmbed).  Handing x out as an OUTPUT makes autograd deliver the shortcut's gradient to this backward, where the
    add_ln backward kernel adds it to the LN input gradient (dxo) - otherwise x has two consumers and autograd sums the
    two full-size fp32 gradients with a separate add kernel (112 us for Swin-T at B = 64)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float, y_bf16: bool):
        x = _chk(x, F32, "x")
        gamma, beta = _chk(gamma, F32, "gamma"), _chk(beta, F32, "beta")
        _, y, mean, rstd = _add_ln_fwd(x, None, None, 1, gamma, beta, eps, True, y_bf16)
        ctx.save_for_backward(x, mean, rstd, gamma)
        ctx.y_bf16 = y_bf16
        return x.view(x.shape), y

    @staticmethod
    @once_differentiable
    def backward(ctx, g_x, g_y):
        x, mean, rstd, gamma = ctx.saved_tensors
        T, C = x.numel() // x.shape[-1], x.shape[-1]
        g_x = _chk(g_x, F32, "g_x") if g_x is not None else None
        if g_y is None:
            return g_x, None, None, None, None
        g_y = _chk(g_y, BF16 if ctx.y_bf16 else F32, "g_y")
        dx = torch.empty_like(x)
        acc, first = _acc(("ln", gamma.data_ptr()), (2, C), x.device)
        _lib.call("esvit_add_ln_bwd", _p(g_y), 1 if ctx.y_bf16 else 0, _p(g_x), _p(x), _p(mean), _p(rstd), _p(gamma),
                  None, 1, _p(dx), None, _p(acc[0]), _p(acc[1]), None, T, C, _stream())
        return (dx, acc[0], acc[1], None, None) if first else (dx, None, None, None, None)


class ResidualAddFn(Function):
    """xout = x + keep * delta (fp32 + bf16), no norm; delta_bias only receives its gradient (see AddLayerNormFn)."""

    @staticmethod
    def forward(ctx, x, delta, dbias, keep):
        x, delta, keep = _chk(x, F32, "x"), _chk(delta, BF16, "delta"), _chk(keep, F32, "keep")
        dbias = _chk(dbias, F32, "delta_bias")
        tps = x.numel() // x.shape[-1] // x.shape[0]
        xout, _, _, _ = _add_ln_fwd(x, delta, keep, tps, None, None, 0.0, False, False)
        ctx.save_for_backward(keep)
        ctx.tps, ctx.has_dbias = tps, dbias is not None
        ctx.dbias_ptr = dbias.data_ptr() if dbias is not None else 0
        return xout

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (keep,) = ctx.saved_tensors
        g = _chk(g, F32, "g")
        T, C = g.numel() // g.shape[-1], g.shape[-1]
        ddelta = torch.empty(g.shape, dtype=BF16, device=g.device)
        db, first = _acc(("res", ctx.dbias_ptr), (C,), g.device) if ctx.has_dbias else (None, True)
        _lib.call("esvit_add_ln_bwd", None, 0, _p(g), None, None, None, None, _p(keep), ctx.tps, None, _p(ddelta),
                  None, None, _p(db), T, C, _stream())
        return g, ddelta, (db if first else None), None


def add_layer_norm(x: Tensor, delta: Optional[Tensor], keep: Optional[Tensor], gamma: Tensor, beta: Tensor,
                   eps: float, y_bf16: bool = True, delta_bias: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    if delta is None:
        if x.requires_grad and torch.is_grad_enabled():
            return LayerNormResidFn.apply(x, gamma, beta, eps, y_bf16)
        return x, LayerNormFn.apply(x, gamma, beta, eps, y_bf16)
    return AddLayerNormFn.apply(x, delta, delta_bias, keep, gamma, beta, eps, y_bf16)  # x may be None: xout = fp32(delta)


def residual_add(x: Tensor, delta: Optional[Tensor], keep: Optional[Tensor],
                 delta_bias: Optional[Tensor] = None) -> Tensor:
    return x if delta is None else ResidualAddFn.apply(x, delta, delta_bias, keep)


# ------------------------------------------------------------------------------------------------------------
class PatchMergeLNFn(Function):
    """x fp32 [B, H*W, C] -> LN(2x2 gather) bf16 [B, ceil(H/2)*ceil(W/2), 4C]."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float, H: int, W: int):
        x, gamma, beta = _chk(x, F32, "x"), _chk(gamma, F32, "gamma"), _chk(beta, F32, "beta")
        B, L, C = x.shape
        assert L == H * W
        Lo = ((H + 1) // 2) * ((W + 1) // 2)
        y = torch.empty(B, Lo, 4 * C, dtype=BF16, device=x.device)
        mean = torch.empty(B * Lo, dtype=F32, device=x.device)
        rstd = torch.empty(B * Lo, dtype=F32, device=x.device)
        _lib.call("esvit_patch_merge_ln_fwd", _p(x), _p(gamma), _p(beta), eps, _p(y), _p(mean), _p(rstd), B, H, W, C,
                  _stream())
        ctx.save_for_backward(x, mean, rstd, gamma)
        ctx.hw = (H, W)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, mean, rstd, gamma = ctx.saved_tensors
        H, W = ctx.hw
        B, L, C = x.shape
        g = _chk(g, BF16, "g")
        dx = torch.empty_like(x)
        acc, first = _acc(("merge", gamma.data_ptr()), (2, gamma.numel()), x.device)
        dgamma, dbeta = acc[0], acc[1]
        _lib.call("esvit_patch_merge_ln_bwd", _p(g), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dx), _p(dgamma),
                  _p(dbeta), B, H, W, C, _stream())
        return (dx, dgamma, dbeta, None, None, None) if first else (dx, None, None, None, None, None)


class TokenMeanFn(Function):
    """region fp32 [B, N, C] -> pooled [B, C]."""

    @staticmethod
    def forward(ctx, region):
        region = _chk(region, F32, "region")
        B, N, C = region.shape
        pooled = torch.empty(B, C, dtype=F32, device=region.device)
        _lib.call("esvit_token_mean_fwd", _p(region), _p(pooled), B, N, C, _stream())
        ctx.shape = (B, N, C)
        return pooled

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        B, N, C = ctx.shape
        g = _chk(g, F32, "g")
        d = torch.empty(B, N, C, dtype=F32, device=g.device)
        _lib.call("esvit_token_mean_bwd", _p(g), None, _p(d), B, N, C, _stream())
        return d


class PatchEmbedFn(Function):
    """img fp32 [B,3,H,W] -> LN(conv4x4/4) fp32 [B, (H/4)(W/4), E]."""

    @staticmethod
    def forward(ctx, img, w, bias, gamma, beta, eps: float):
        img, w, bias = _chk(img, F32, "img"), _chk(w, F32, "w"), _chk(bias, F32, "bias")
        gamma, beta = _chk(gamma, F32, "gamma"), _chk(beta, F32, "beta")
        B, Cin, H, W = img.shape
        E = w.shape[0]
        if Cin != 3 or tuple(w.shape[1:]) != (3, 4, 4):
            raise ValueError("PatchEmbed kernel supports in_chans=3, patch_size=4")
        T = B * (H // 4) * (W // 4)
        out = torch.empty(B, (H // 4) * (W // 4), E, dtype=F32, device=img.device)
        mean = torch.empty(T, dtype=F32, device=img.device)
        rstd = torch.empty(T, dtype=F32, device=img.device)
        _lib.call("esvit_patch_embed_fwd", _p(img), _p(w), _p(bias), _p(gamma), _p(beta), eps, _p(out), _p(mean),
                  _p(rstd), B, H, W, E, _stream())
        ctx.save_for_backward(img, w, bias, gamma, mean, rstd)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        img, w, bias, gamma, mean, rstd = ctx.saved_tensors
        B, _, H, W = img.shape
        E = w.shape[0]
        g = _chk(g, F32, "g")
        dw, first = _acc(("pe_w", w.data_ptr()), tuple(w.shape), w.device)
        acc, _ = _acc(("pe_b", w.data_ptr()), (3, E), w.device)
        db, dgamma, dbeta = acc[0], acc[1], acc[2]
        _lib.call("esvit_patch_embed_bwd", _p(img), _p(w), _p(bias), _p(gamma), _p(mean), _p(rstd), _p(g), _p(dw),
                  _p(db), _p(dgamma), _p(dbeta), B, H, W, E, _stream())
        return (None, dw, db, dgamma, dbeta, None) if first else (None, None, None, None, None, None)


class PatchEmbedGroupsFn(Function):
    """PatchEmbedFn over several resolution groups written into ONE token buffer fp32 [sum_g B_g L_g, E] (the
    concatenated residual stream of the fused multi-crop forward): no torch.cat of the groups' outputs (267 MB copied
    per step for Swin-T at B = 64) and no split of the gradient."""

    @staticmethod
    def forward(ctx, w, bias, gamma, beta, eps: float, *imgs):
        w, bias = _chk(w, F32, "w"), _chk(bias, F32, "bias")
        gamma, beta = _chk(gamma, F32, "gamma"), _chk(beta, F32, "beta")
        imgs = [_chk(im, F32, "img") for im in imgs]
        E = w.shape[0]
        if tuple(w.shape[1:]) != (3, 4, 4) or any(im.shape[1] != 3 for im in imgs):
            raise ValueError("PatchEmbed kernel supports in_chans=3, patch_size=4")
        rows = [im.shape[0] * (im.shape[2] // 4) * (im.shape[3] // 4) for im in imgs]
        T = sum(rows)
        dev = imgs[0].device
        out = torch.empty(T, E, dtype=F32, device=dev)
        mean = torch.empty(T, dtype=F32, device=dev)
        rstd = torch.empty(T, dtype=F32, device=dev)
        r0 = 0
        for im, n in zip(imgs, rows):
            B, _, H, W = im.shape
            _lib.call("esvit_patch_embed_fwd", _p(im), _p(w), _p(bias), _p(gamma), _p(beta), eps, _p(out[r0:]),
                      _p(mean[r0:]), _p(rstd[r0:]), B, H, W, E, _stream())
            r0 += n
        ctx.save_for_backward(w, bias, gamma, mean, rstd, *imgs)
        ctx.rows = rows
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        w, bias, gamma, mean, rstd = ctx.saved_tensors[:5]
        imgs = ctx.saved_tensors[5:]
        E = w.shape[0]
        g = _chk(g, F32, "g")
        dw, first = _acc(("pe_w", w.data_ptr()), tuple(w.shape), w.device)
        acc, _ = _acc(("pe_b", w.data_ptr()), (3, E), w.device)
        r0 = 0
        for im, n in zip(imgs, ctx.rows):
            B, _, H, W = im.shape
            _lib.call("esvit_patch_embed_bwd", _p(im), _p(w), _p(bias), _p(gamma), _p(mean[r0:]), _p(rstd[r0:]),
                      _p(g[r0:]), _p(dw), _p(acc[0]), _p(acc[1]), _p(acc[2]), B, H, W, E, _stream())
            r0 += n
        none = (None,) * len(imgs)
        return ((dw, acc[0], acc[1], acc[2], None) if first else (None, None, None, None, None)) + none


def cat_adjacent(ts):
    """torch.cat(ts) along dim 0 - as a VIEW when the tensors already lie back to back in one storage (the engine's
    static crop buffers do: the 2 global / 8 local crops of a step), so the multi-crop forward copies nothing."""
    ts = list(ts)
    if len(ts) == 1:
        return ts[0]
    t0 = ts[0]
    ok = all(t.shape == t0.shape and t.dtype == t0.dtype and t.device == t0.device and t.is_contiguous()
             and not t.requires_grad for t in ts)
    if ok:
        st, n = t0.untyped_storage(), t0.numel()
        ok = all(t.untyped_storage().data_ptr() == st.data_ptr() and t.storage_offset() == t0.storage_offset() + i * n
                 for i, t in enumerate(ts))
    if not ok:
        return torch.cat(ts)
    shape = (t0.shape[0] * len(ts),) + tuple(t0.shape[1:])
    return torch.as_strided(t0, shape, t0.stride(), t0.storage_offset())


# ------------------------------------------------------------------------------------------------------------
ATTN_WS_FLOATS = 8192  # per head; include/esvit_b200.h (expanded bias for ws 7, bias-gradient accumulator for ws 14)


class WindowAttentionFn(Function):
    """qkv bf16 [B, H*W, 3C] (qkv GEMM output incl. bias) -> attention output bf16 [B, H*W, C] in token order
    (pad / roll / partition / reverse folded in).  qkv_bias (fp32 [3C] parameter) supplies the value of padded slots
    and receives the COMPLETE qkv-bias gradient from the backward kernel (column sums of dq/dk/dv)."""

    @staticmethod
    def forward(ctx, qkv, qkv_bias, bias_table, H: int, W: int, num_heads: int, ws: int, shift: int, scale: float,
                bias_exp: Optional[Tensor]):
        qkv = _chk(qkv, BF16, "qkv")
        qkv_bias = _chk(qkv_bias, F32, "qkv_bias")
        bias_table = _chk(bias_table, F32, "relative_position_bias_table")
        B, L, C3 = qkv.shape
        C = C3 // 3
        assert L == H * W
        qb = shadow.lookup(qkv_bias)  # optimiser-maintained bf16 copy (no cast kernel per call) when registered
        if qb is None:
            qb = qkv_bias.detach().to(BF16)
        nwin = B * (-(-H // ws)) * (-(-W // ws))
        out = torch.empty(B, L, C, dtype=BF16, device=qkv.device)
        lse = torch.empty(nwin * num_heads * ws * ws, dtype=F32, device=qkv.device)
        # bias_exp: the table already expanded for this step by expand_rel_pos_bias (shared by every call and by the
        # backward); without it each call expands into its own scratch
        ready = 1 if (bias_exp is not None and ws == 7) else 0
        bws = bias_exp if ready else torch.empty(num_heads * ATTN_WS_FLOATS, dtype=F32, device=qkv.device)
        _lib.call("esvit_window_attn_fwd", _p(qkv), _p(qb), _p(bias_table), _p(bws), ready, _p(out), _p(lse), B, H, W, C,
                  num_heads, ws, shift, scale, _stream())
        ctx.save_for_backward(qkv, qb, bias_table, out, lse, bws if ready else None)
        ctx.geo = (B, H, W, C, num_heads, ws, shift, scale)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        qkv, qb, bias_table, out, lse, bias_exp = ctx.saved_tensors
        B, H, W, C, nH, ws, shift, scale = ctx.geo
        ready = 1 if bias_exp is not None else 0
        g = _chk(g, BF16, "g")
        dqkv = torch.empty_like(qkv)
        dtable, first = _acc(("attn_t", bias_table.data_ptr()), tuple(bias_table.shape), qkv.device)
        dqb, _ = _acc(("attn_b", bias_table.data_ptr()), (3 * C,), qkv.device)
        bws = bias_exp if ready else torch.empty(nH * ATTN_WS_FLOATS, dtype=F32, device=qkv.device)
        _lib.call("esvit_window_attn_bwd", _p(qkv), _p(qb), _p(bias_table), _p(bws), ready, _p(out), _p(g), _p(lse), _p(dqkv),
                  _p(dtable), _p(dqb), B, H, W, C, nH, ws, shift, scale, _stream())
        return dqkv, (dqb if first else None), (dtable if first else None), None, None, None, None, None, None, None


class WindowAttentionGroupsFn(Function):
    """WindowAttentionFn over several resolution groups stored back to back in ONE token-major tensor: qkv bf16
    [T, 3C], group g = (B, H, W, row0) = rows [row0, row0 + B*H*W) holding B maps of H x W tokens.  One kernel launch
    per group on pointer offsets (no slicing / concatenation copies); the table / qkv-bias gradients of all groups
    accumulate into the same buffers."""

    @staticmethod
    def forward(ctx, qkv, qkv_bias, bias_table, groups, num_heads: int, ws: int, shift: int, scale: float, bias_exp):
        qkv = _chk(qkv, BF16, "qkv")
        qkv_bias = _chk(qkv_bias, F32, "qkv_bias")
        bias_table = _chk(bias_table, F32, "relative_position_bias_table")
        T, C3 = qkv.shape
        C = C3 // 3
        assert sum(B * H * W for B, H, W, _ in groups) == T
        qb = shadow.lookup(qkv_bias)
        if qb is None:
            qb = qkv_bias.detach().to(BF16)
        out = torch.empty(T, C, dtype=BF16, device=qkv.device)
        lse_off, n = [], 0
        for B, H, W, _ in groups:
            lse_off.append(n)
            n += B * (-(-H // ws)) * (-(-W // ws)) * num_heads * ws * ws
        lse = torch.empty(n, dtype=F32, device=qkv.device)
        ready = 1 if (bias_exp is not None and ws == 7) else 0
        bws = bias_exp if ready else torch.empty(num_heads * ATTN_WS_FLOATS, dtype=F32, device=qkv.device)
        for (B, H, W, r0), lo in zip(groups, lse_off):
            _lib.call("esvit_window_attn_fwd", _po(qkv, r0 * C3), _p(qb), _p(bias_table), _p(bws), ready, _po(out, r0 * C),
                      _po(lse, lo), B, H, W, C, num_heads, ws, shift, scale, _stream())
        ctx.save_for_backward(qkv, qb, bias_table, out, lse, bws if ready else None)
        ctx.meta = (tuple(groups), tuple(lse_off), C, num_heads, ws, shift, scale)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        qkv, qb, bias_table, out, lse, bias_exp = ctx.saved_tensors
        groups, lse_off, C, nH, ws, shift, scale = ctx.meta
        g = _chk(g, BF16, "g")
        ready = 1 if bias_exp is not None else 0
        dqkv = torch.empty_like(qkv)
        dtable, first = _acc(("attn_t", bias_table.data_ptr()), tuple(bias_table.shape), qkv.device)
        dqb, _ = _acc(("attn_b", bias_table.data_ptr()), (3 * C,), qkv.device)
        bws = bias_exp if ready else torch.empty(nH * ATTN_WS_FLOATS, dtype=F32, device=qkv.device)
        for (B, H, W, r0), lo in zip(groups, lse_off):
            _lib.call("esvit_window_attn_bwd", _po(qkv, r0 * 3 * C), _p(qb), _p(bias_table), _p(bws), ready, _po(out, r0 * C),
                      _po(g, r0 * C), _po(lse, lo), _po(dqkv, r0 * 3 * C), _p(dtable), _p(dqb), B, H, W, C, nH, ws, shift,
                      scale, _stream())
        return dqkv, (dqb if first else None), (dtable if first else None), None, None, None, None, None, None


class PatchMergeLNGroupsFn(Function):
    """PatchMergeLNFn over resolution groups stored back to back: x fp32 [T, C] -> bf16 [T', 4C] (T' = sum of
    B * ceil(H/2) * ceil(W/2)), one launch per group on pointer offsets."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float, groups):
        x, gamma, beta = _chk(x, F32, "x"), _chk(gamma, F32, "gamma"), _chk(beta, F32, "beta")
        T, C = x.shape
        assert sum(B * H * W for B, H, W, _ in groups) == T
        out_off, n = [], 0
        for B, H, W, _ in groups:
            out_off.append(n)
            n += B * ((H + 1) // 2) * ((W + 1) // 2)
        y = torch.empty(n, 4 * C, dtype=BF16, device=x.device)
        mean = torch.empty(n, dtype=F32, device=x.device)
        rstd = torch.empty(n, dtype=F32, device=x.device)
        for (B, H, W, r0), o0 in zip(groups, out_off):
            _lib.call("esvit_patch_merge_ln_fwd", _po(x, r0 * C), _p(gamma), _p(beta), eps, _po(y, o0 * 4 * C), _po(mean, o0),
                      _po(rstd, o0), B, H, W, C, _stream())
        ctx.save_for_backward(x, mean, rstd, gamma)
        ctx.meta = (tuple(groups), tuple(out_off))
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, mean, rstd, gamma = ctx.saved_tensors
        groups, out_off = ctx.meta
        T, C = x.shape
        g = _chk(g, BF16, "g")
        dx = torch.empty_like(x)
        acc, first = _acc(("merge", gamma.data_ptr()), (2, gamma.numel()), x.device)
        for (B, H, W, r0), o0 in zip(groups, out_off):
            _lib.call("esvit_patch_merge_ln_bwd", _po(g, o0 * 4 * C), _po(x, r0 * C), _po(mean, o0), _po(rstd, o0), _p(gamma),
                      _po(dx, r0 * C), _p(acc[0]), _p(acc[1]), B, H, W, C, _stream())
        return (dx, acc[0], acc[1], None, None) if first else (dx, None, None, None, None)


class TokenMeanGroupsFn(Function):
    """TokenMeanFn over resolution groups stored back to back: region fp32 [T, C] -> pooled [sum B, C]."""

    @staticmethod
    def forward(ctx, region, groups):
        region = _chk(region, F32, "region")
        T, C = region.shape
        nb = sum(B for B, _, _, _ in groups)
        pooled = torch.empty(nb, C, dtype=F32, device=region.device)
        b0 = 0
        for B, H, W, r0 in groups:
            _lib.call("esvit_token_mean_fwd", _po(region, r0 * C), _po(pooled, b0 * C), B, H * W, C, _stream())
            b0 += B
        ctx.meta = (tuple(groups), T, C)
        return pooled

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        groups, T, C = ctx.meta
        g = _chk(g, F32, "g")
        d = torch.empty(T, C, dtype=F32, device=g.device)
        b0 = 0
        for B, H, W, r0 in groups:
            _lib.call("esvit_token_mean_bwd", _po(g, b0 * C), None, _po(d, r0 * C), B, H * W, C, _stream())
            b0 += B
        return d, None


def expand_rel_pos_bias(bias_table: Tensor, num_heads: int, ws: int) -> Optional[Tensor]:
    """ws = 7: the rel-pos bias table expanded ONCE for all attention calls (both crop groups, forward and backward) that
    use it this step -> fp32 [nH*4096] to pass as WindowAttentionFn's bias_exp; ws = 14: None (staged per CTA)."""
    if ws != 7:
        return None
    bias_table = _chk(bias_table.detach(), F32, "relative_position_bias_table")
    bws = torch.empty(num_heads * 4096, dtype=F32, device=bias_table.device)
    _lib.call("esvit_window_attn_expand_bias", _p(bias_table), _p(bws), num_heads, ws, _stream())
    return bws


class GeluFn(Function):
    @staticmethod
    def forward(ctx, x):
        x = _chk(x, BF16, "x")
        y = torch.empty_like(x)
        _lib.call("esvit_gelu_fwd", _p(x), _p(y), x.numel(), _stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _chk(g, BF16, "g")
        dx = torch.empty_like(x)
        _lib.call("esvit_gelu_bwd", _p(x), _p(g), _p(dx), x.numel(), _stream())
        return dx


class BiasGeluFn(Function):
    """y = gelu(x) for x bf16 [..., N] that already holds the producing GEMM's bias; `bias` (fp32 [N] parameter) only
    receives its gradient = column sums of dx, computed inside the GELU backward kernel."""

    @staticmethod
    def forward(ctx, x, bias):
        x = _chk(x, BF16, "x")
        y = torch.empty_like(x)
        _lib.call("esvit_gelu_fwd", _p(x), _p(y), x.numel(), _stream())
        ctx.save_for_backward(x)
        ctx.bias_meta = (bias.shape, bias.device, bias.data_ptr())
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _chk(g, BF16, "g")
        N = x.shape[-1]
        dx = torch.empty_like(x)
        db, first = _acc(("bias", ctx.bias_meta[2]), tuple(ctx.bias_meta[0]), ctx.bias_meta[1])
        _lib.call("esvit_gelu_bwd_dbias", _p(x), _p(g), _p(dx), _p(db), x.numel() // N, N, _stream())
        return dx, (db if first else None)


GEMM_COLSUM_WS_ROWS = 160  # include/esvit_b200.h: esvit_gemm_mul_colsum scratch rows


def gemm_bias_act(a: Tensor, w: Tensor, bias: Optional[Tensor], act: int = 0, want_pre: bool = False):
    """tcgen05/TMA GEMM: act(a @ w^T + bias) -> bf16 [M, N] (and gelu'(a @ w^T + bias) when act != 0 and want_pre)."""
    a, w = _chk(a, BF16, "a"), _chk(w, BF16, "w")
    bias = _chk(bias, F32, "bias")
    K = a.shape[-1]
    M = a.numel() // K
    N = w.shape[0]
    out = torch.empty(*a.shape[:-1], N, dtype=BF16, device=a.device)
    pre = torch.empty_like(out) if (act and want_pre) else None
    _lib.call("esvit_gemm_bias_act", _p(a), _p(w), _p(bias), _p(out), _p(pre), M, N, K, act, _stream())
    return (out, pre) if (act and want_pre) else out


def gemm(a: Tensor, b: Tensor, bias: Optional[Tensor] = None, act: int = 0, want_pre: bool = False, a_mn: bool = False,
         b_mn: bool = False, tile: int = 0):
    """Second-generation tcgen05 GEMM: act(opA(a) @ opB(b) + bias) -> bf16 [M, N].
    a: [..., K] (a_mn: [K, M]);  b: [N, K] (b_mn: [K, N] - a Linear weight read as it lies for the input gradient)."""
    a, b = _chk(a, BF16, "a"), _chk(b, BF16, "b")
    bias = _chk(bias, F32, "bias")
    if a_mn:
        K, M = a.shape
        lead = (M,)
    else:
        K = a.shape[-1]
        M = a.numel() // K
        lead = tuple(a.shape[:-1])
    N = b.shape[1] if b_mn else b.shape[0]
    assert (b.shape[0] if b_mn else b.shape[1]) == K, (a.shape, b.shape)
    out = torch.empty(*lead, N, dtype=BF16, device=a.device)
    pre = torch.empty_like(out) if (act and want_pre) else None
    _lib.call("esvit_gemm_bf16", _p(a), _p(b), _p(bias), _p(out), _p(pre), M, N, K, 1 if a_mn else 0, 1 if b_mn else 0, act,
              tile, _stream())
    return (out, pre) if (act and want_pre) else out


def gemm_mul_colsum(a: Tensor, b: Tensor, mult: Tensor, colsum: Tensor, b_mn: bool = False, tile: int = 0) -> Tensor:
    """out = (a @ opB(b)) * mult (bf16); colsum (fp32 [N]) += column sums of out."""
    a, b, mult = _chk(a, BF16, "a"), _chk(b, BF16, "b"), _chk(mult, BF16, "mult")
    K = a.shape[-1]
    M = a.numel() // K
    N = b.shape[1] if b_mn else b.shape[0]
    out = torch.empty_like(mult)
    ws = torch.empty(GEMM_COLSUM_WS_ROWS * N, dtype=F32, device=a.device)
    _lib.call("esvit_gemm_mul_colsum2", _p(a), _p(b), _p(mult), _p(out), _p(colsum), _p(ws), M, N, K, 1 if b_mn else 0, tile,
              _stream())
    return out


_wgrad_ws = {}


def gemm_wgrad(dy: Tensor, x: Tensor, out: Optional[Tensor] = None, accumulate: bool = False, tile: int = 0) -> Tensor:
    """dw[N, K] (fp32) (+)= dy[T, N]^T @ x[T, K]: the weight gradient of a Linear straight in fp32 (no bf16 round trip,
    no cast kernel, no transposes); deterministic split-K."""
    dy, x = _chk(dy, BF16, "dy"), _chk(x, BF16, "x")
    N, K = dy.shape[-1], x.shape[-1]
    T = dy.numel() // N
    assert x.numel() // K == T
    if out is None:
        out = torch.empty(N, K, dtype=F32, device=dy.device)
        accumulate = False
    key = (dy.device, N, K)
    ws = _wgrad_ws.get(key)
    if ws is None:
        n = _lib.load().esvit_gemm_wgrad_ws_floats(N, K)
        if n <= 0:
            raise ValueError("esvit_gemm_wgrad: weight too large for the split-K workspace")
        ws = _wgrad_ws[key] = torch.empty(n, dtype=F32, device=dy.device)
    _lib.call("esvit_gemm_wgrad", _p(dy), _p(x), _p(out), _p(ws), T, N, K, 1 if accumulate else 0, tile, _stream())
    return out


class LinearGeluFn(Function):
    """gelu(x @ w^T + b) as ONE tcgen05/TMA kernel (esvit_gemm_bias_act): the GEMM epilogue adds the bias, applies the
    exact GELU and also emits gelu'(pre-activation), so the [T, 4C] hidden tensor is written once and never re-read in
    the forward and the backward's dh = dy * gelu' (+ bias gradient) is a pure streaming kernel; dx / dw are library
    GEMMs."""

    @staticmethod
    def forward(ctx, x, w, bias):
        out, pre = gemm_bias_act(x, w, bias, act=1, want_pre=True)
        ctx.save_for_backward(x, w, pre)
        ctx.bias_meta = (bias.shape, bias.device, bias.data_ptr())
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w, pre = ctx.saved_tensors  # pre = gelu'(x @ w^T + b)
        g = _chk(g, BF16, "g")
        N = pre.shape[-1]
        dh = torch.empty_like(pre)
        db, first = _acc(("bias", ctx.bias_meta[2]), tuple(ctx.bias_meta[0]), ctx.bias_meta[1])
        _lib.call("esvit_mul_bwd_dbias", _p(pre), _p(g), _p(dh), _p(db), pre.numel() // N, N, _stream())
        dh2 = dh.reshape(-1, N)
        dx = (dh2 @ w).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = dh2.t() @ x.reshape(-1, x.shape[-1]) if ctx.needs_input_grad[1] else None
        return dx, dw, (db if first else None)




class MlpFn(Function):
    """fc2(gelu(fc1(x))) of the Swin MLP (models/swin_transformer.py:31-35) with both GELU passes inside tcgen05 GEMM
    epilogues.  forward: h, gelu' = esvit_gemm_bias_act(x, w1, b1) (one kernel), y = h @ w2^T + b2 (library GEMM).
    backward: d(pre) = (dy @ w2) * gelu' and the fc1 bias gradient come out of ONE kernel (esvit_gemm_mul_colsum, w2t =
    w2^T bf16 [4C, C]); the hidden-sized dh tensor of the unfused chain (GEMM -> multiply kernel) is never written.
    fc2's bias gradient is produced by the consumer (residual add + LN backward), as for LinearBiasFn."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w2t):
        h, pre = gemm_bias_act(x, w1, b1, act=1, want_pre=True)
        with torch.autocast("cuda", enabled=False):
            y = torch.nn.functional.linear(h, w2, b2)
        ctx.save_for_backward(x, w1, pre, h, w2t)
        ctx.bias_meta = (b1.shape, b1.device, b1.data_ptr())
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w1, pre, h, w2t = ctx.saved_tensors
        g = _chk(g, BF16, "g")
        C, N = g.shape[-1], pre.shape[-1]
        g2 = g.reshape(-1, C)
        M = g2.shape[0]
        dw2 = g2.t() @ h.reshape(-1, N) if ctx.needs_input_grad[3] else None
        dpre = torch.empty_like(pre)
        db1, first = _acc(("bias", ctx.bias_meta[2]), tuple(ctx.bias_meta[0]), ctx.bias_meta[1])
        ws = torch.empty(GEMM_COLSUM_WS_ROWS * N, dtype=F32, device=g.device)
        _lib.call("esvit_gemm_mul_colsum", _p(g2), _p(w2t), _p(pre), _p(dpre), _p(db1), _p(ws), M, N, C, _stream())
        d2 = dpre.reshape(-1, N)
        dx = (d2 @ w1).view(x.shape) if ctx.needs_input_grad[0] else None
        dw1 = d2.t() @ x.reshape(-1, x.shape[-1]) if ctx.needs_input_grad[1] else None
        return dx, dw1, (db1 if first else None), dw2, None, None


class LinearBiasFn(Function):
    """y = x @ w^T + b as ONE library GEMM (bias in the cuBLASLt epilogue).  The backward produces dx and dw with two
    library GEMMs and NO bias gradient: the consumer kernel (window attention / GELU / add+LN backward) column-sums
    it for free, which removes the reference's per-layer `grad.sum(0)` reduction kernels (10 % of the first profile)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        with torch.autocast("cuda", enabled=False):
            return torch.nn.functional.linear(x, w, b)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        dx = (g2 @ w).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = g2.t() @ x.reshape(-1, x.shape[-1]) if ctx.needs_input_grad[1] else None
        return dx, dw, None


class L2NormFn(Function):
    @staticmethod
    def forward(ctx, x, eps: float):
        x = _chk(x, BF16, "x")
        R, Dm = x.numel() // x.shape[-1], x.shape[-1]
        y = torch.empty_like(x)
        inv = torch.empty(R, dtype=F32, device=x.device)
        _lib.call("esvit_l2norm_fwd", _p(x), _p(y), _p(inv), eps, R, Dm, _stream())
        ctx.save_for_backward(x, inv)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, inv = ctx.saved_tensors
        g = _chk(g, BF16, "g")
        R, Dm = x.numel() // x.shape[-1], x.shape[-1]
        dx = torch.empty_like(x)
        _lib.call("esvit_l2norm_bwd", _p(x), _p(g), _p(inv), _p(dx), R, Dm, _stream())
        return dx, None


class WeightNormFn(Function):
    """(v fp32 [K,D], g fp32 [K,1]) -> w bf16 [K,D] = v * g / ||v||_row."""

    @staticmethod
    def forward(ctx, v, g):
        v, g = _chk(v, F32, "weight_v"), _chk(g, F32, "weight_g")
        K, Dm = v.shape
        w = torch.empty(K, Dm, dtype=BF16, device=v.device)
        norm = torch.empty(K, dtype=F32, device=v.device)
        _lib.call("esvit_weight_norm_fwd", _p(v), _p(g), _p(w), _p(norm), K, Dm, _stream())
        ctx.save_for_backward(v, g, norm)
        return w

    @staticmethod
    @once_differentiable
    def backward(ctx, gw):
        v, g, norm = ctx.saved_tensors
        f32 = gw.dtype == F32  # fp32 from esvit_gemm_wgrad, bf16 from a library GEMM
        gw = _chk(gw, F32 if f32 else BF16, "gw")
        K, Dm = v.shape
        dv = torch.empty_like(v)
        dg = torch.empty_like(g) if ctx.needs_input_grad[1] else None
        _lib.call("esvit_weight_norm_bwd", _p(v), _p(g), _p(norm), _p(gw), 1 if f32 else 0, _p(dv), _p(dg), K, Dm, _stream())
        return dv, dg


# ------------------------------------------------------------------------------------------------------------
# losses


def row_lse(x: Tensor, center: Optional[Tensor], inv_temp: float) -> Tensor:
    x = _chk(x, BF16, "logits")
    R, K = x.shape
    lse = torch.empty(R, dtype=F32, device=x.device)
    _lib.call("esvit_row_lse", _p(x), _p(center), inv_temp, _p(lse), R, K, _stream())
    return lse


def ce_q_enabled(K: int) -> bool:
    """The CE kernels run on teacher probabilities stored once per teacher row (fp16, esvit_row_softmax_q) unless
    ESVIT_CE_Q=0 or a row of K logits does not fit one CTA's shared memory."""
    import os
    return os.environ.get("ESVIT_CE_Q", "1") != "0" and K <= _lib.load().esvit_row_softmax_q_max_k()


def row_softmax_q(x: Tensor, center: Tensor, inv_temp: float) -> Tuple[Tensor, Tensor]:
    """(lse fp32 [R] as row_lse, q fp16 [R, K] = 2^12 * softmax((x - center) * inv_temp)): ONE pass over the teacher rows."""
    x, center = _chk(x, BF16, "teacher logits"), _chk(center, F32, "center")
    R, K = x.shape
    lse = torch.empty(R, dtype=F32, device=x.device)
    q = torch.empty(R, K, dtype=torch.float16, device=x.device)
    _lib.call("esvit_row_softmax_q", _p(x), _p(center), inv_temp, _p(lse), _p(q), R, K, _stream())
    return lse, q


class DinoCEFn(Function):
    """loss = sum_r w[r] * ( n_r * LSE(s_r / tau) - sum_j <softmax((t[trow[r,j]] - center) / temp), s_r / tau> ).

    s bf16 [R,K] (grad), t bf16 [Rt,K], center fp32 [K], trow int32 [R,2], w fp32 [R].
    lse_t = None (the default path): the teacher probabilities are computed ONCE per teacher row and stored in fp16
    (row_softmax_q), the CE kernels stream them; lse_t = row_lse(t, center, inv_temp_t): every pairing recomputes the
    teacher exponentials from the logits (the round-1 kernels)."""

    @staticmethod
    def forward(ctx, s, t, center, lse_t, trow, w, inv_temp_t: float, inv_tau_s: float, order=None):
        s, t = _chk(s, BF16, "student logits"), _chk(t, BF16, "teacher logits")
        center, w = _chk(center, F32, "center"), _chk(w, F32, "w")
        trow = _chk(trow, torch.int32, "trow")
        order = _chk(order, torch.int32, "order")
        R, K = s.shape
        lse_s = torch.empty(R, dtype=F32, device=s.device)  # written by the CE kernel itself (one pass over s)
        row_loss = torch.empty(R, dtype=F32, device=s.device)
        ctx.use_q = lse_t is None
        if ctx.use_q:
            _, q = row_softmax_q(t, center, inv_temp_t)
            _lib.call("esvit_dino_ce_q_fwd", _p(s), _p(q), _p(lse_s), _p(trow), _p(order), inv_tau_s, _p(row_loss), R, K,
                      _stream())
            ctx.save_for_backward(s, q, lse_s, trow, w, order)
        else:
            lse_t = _chk(lse_t, F32, "lse_t")
            _lib.call("esvit_dino_ce_fwd", _p(s), _p(t), _p(center), _p(lse_s), _p(lse_t), _p(trow), _p(order), inv_temp_t,
                      inv_tau_s, _p(row_loss), R, K, _stream())
            ctx.save_for_backward(s, t, center, lse_s, lse_t, trow, w, order)
        loss = torch.empty((), dtype=F32, device=s.device)
        _lib.call("esvit_weighted_sum", _p(row_loss), _p(w), R, _p(loss), _stream())
        ctx.temps = (inv_temp_t, inv_tau_s)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        inv_temp_t, inv_tau_s = ctx.temps
        gs = _chk(g.reshape(1).to(F32), F32, "g")
        if ctx.use_q:
            s, q, lse_s, trow, w, order = ctx.saved_tensors
            R, K = s.shape
            ds = torch.empty_like(s)
            _lib.call("esvit_dino_ce_q_bwd", _p(s), _p(q), _p(lse_s), _p(trow), _p(order), _p(w), _p(gs), inv_tau_s, _p(ds),
                      R, K, _stream())
        else:
            s, t, center, lse_s, lse_t, trow, w, order = ctx.saved_tensors
            R, K = s.shape
            ds = torch.empty_like(s)
            _lib.call("esvit_dino_ce_bwd", _p(s), _p(t), _p(center), _p(lse_s), _p(lse_t), _p(trow), _p(order), _p(w),
                      _p(gs), inv_temp_t, inv_tau_s, _p(ds), R, K, _stream())
        return ds, None, None, None, None, None, None, None, None


_colsum_ws = {}


def colsum(t: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """fp32 column sums of a bf16 [R, K] matrix (deterministic)."""
    t = _chk(t, BF16, "teacher logits")
    R, K = t.shape
    key = (t.device, K)
    ws = _colsum_ws.get(key)
    if ws is None:
        rows = _lib.load().esvit_colsum_workspace_rows()
        ws = _colsum_ws[key] = torch.empty(rows * K, dtype=F32, device=t.device)
    if out is None:
        out = torch.empty(K, dtype=F32, device=t.device)
    _lib.call("esvit_colsum", _p(t), R, K, _p(ws), _p(out), _stream())
    return out


def center_ema(center: Tensor, colsum_total: Tensor, rows_total: int, momentum: float,
               out: Optional[Tensor] = None) -> Tensor:
    """out = center*m + (colsum/rows)*(1-m); out defaults to a new tensor, out=center updates in place."""
    assert center.is_cuda and center.dtype == F32 and center.is_contiguous()
    if out is None:
        out = torch.empty_like(center)
    _lib.call("esvit_center_ema", _p(center), _p(colsum_total), float(rows_total), momentum, _p(out), center.numel(),
              _stream())
    return out


def normalize_rows(x: Tensor, eps: float = 1e-12) -> Tensor:
    x = _chk(x, F32, "features")
    y = torch.empty_like(x)
    _lib.call("esvit_normalize_rows", _p(x), _p(y), x.shape[0], x.shape[1], eps, _stream())
    return y


def region_match(s_fea: Tensor, t_fea: Tensor, B: int, ncrops: int, Tg: int, Tl: int) -> Tuple[Tensor, Tensor]:
    """Cosine arg-max of every student region token against the teacher view's tokens of the same image.

    Returns (idx int64 [2, ncrops, B, Tg] with -1 in unused slots, trow int32 [Rs, 2] teacher region row per iq)."""
    sn, tn = normalize_rows(s_fea), normalize_rows(t_fea)
    P = sn.shape[1]
    Rs = sn.shape[0]
    assert Rs == B * (2 * Tg + (ncrops - 2) * Tl) and tn.shape[0] == 2 * B * Tg
    idx = torch.full((2, ncrops, B, Tg), -1, dtype=torch.int64, device=sn.device)
    trow = torch.empty(Rs, 2, dtype=torch.int32, device=sn.device)
    _lib.call("esvit_region_match", _p(sn), _p(tn), B, ncrops, Tg, Tl, P, _p(idx), _p(trow), _stream())
    return idx, trow


# ------------------------------------------------------------------------------------------------------------
# optimiser-side multi-tensor ops


def _ptr_array(tensors: Sequence[Tensor]):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def _numel_array(tensors: Sequence[Tensor]):
    arr = (ctypes.c_longlong * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.numel()
    return arr


def ema_update_(teacher: Sequence[Tensor], student: Sequence[Tensor], momentum: float) -> None:
    """teacher = teacher * m + (1 - m) * student for every tensor pair; bit-exact with the reference loop."""
    assert len(teacher) == len(student)
    for k, q in zip(teacher, student):
        if not (k.is_cuda and q.is_cuda and k.dtype == F32 and q.dtype == F32 and k.is_contiguous()
                and q.is_contiguous() and k.numel() == q.numel()):
            raise RuntimeError("ema_update_: fp32 contiguous CUDA tensors of equal size required")
    _lib.call("esvit_ema_multi", _ptr_array(teacher), _ptr_array(student), _numel_array(teacher), len(teacher),
              float(momentum), _stream())


def clip_grads_(grads: Sequence[Tensor], clip: float) -> Tensor:
    """Per-tensor L2 clipping in place; returns the pre-clip norms as a device tensor (no host sync)."""
    n = len(grads)
    dev = grads[0].device
    for g in grads:
        if not (g.is_cuda and g.dtype == F32 and g.is_contiguous()):
            raise RuntimeError("clip_grads_: fp32 contiguous CUDA gradients required")
    ws = torch.empty(n, dtype=torch.float64, device=dev)
    norms = torch.empty(n, dtype=F32, device=dev)
    _lib.call("esvit_clip_multi", _ptr_array(grads), _numel_array(grads), n, float(clip), _p(ws), _p(norms), _stream())
    return norms
