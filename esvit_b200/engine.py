"""The multi-crop self-distillation training step (train_one_epoch's loop body, main_esvit.py:507-590) driven
through the esvit_b200 modules.  ``main_esvit.py`` can keep its own loop; this class is the same
sequence packaged for bench.py / smoke / tests:

    lr/wd -> teacher fwd (no grad) -> student fwd -> DINO/DDINO loss (+ packed center all-reduce) -> backward
    (DDP gradient all-reduce when wrapped) -> per-tensor clip -> cancel last-layer grads -> AdamW -> teacher EMA

with NO host synchronisation inside the step (the reference syncs ~190 times per step).
"""
from __future__ import annotations

from functools import partial
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

import torch.distributed as dist

from . import ops, utils
from .losses import DDINOLoss, DINOLoss
from .optim import FusedAdamWEMA
from .swin_transformer import SwinTransformer
from .vision_transformer import DINOHead

SWIN_SPECS = {
    "swin_tiny_w7": dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7, drop_path_rate=0.1),
    "swin_small_w7": dict(embed_dim=96, depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24], window_size=7, drop_path_rate=0.2),
    "swin_small_w14": dict(embed_dim=96, depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24], window_size=14, drop_path_rate=0.2),
    "swin_base_w7": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=7, drop_path_rate=0.2),
    "swin_base_w14": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=14, drop_path_rate=0.2),
}


def build_network(spec: dict, out_dim: int, use_dense_prediction: bool, is_teacher: bool = False,
                  norm_last_layer: bool = True, img_size: int = 224, head_kwargs: Optional[dict] = None) -> nn.Module:
    """What main_esvit.py:235-254 builds: Swin backbone (teacher: drop_path 0) + DINOHead(s) assigned to
    ``.head`` / ``.head_dense``."""
    spec = dict(spec)
    if is_teacher:
        spec["drop_path_rate"] = 0.0
    net = SwinTransformer(img_size=img_size, in_chans=3, num_classes=0, patch_size=4, mlp_ratio=4., qkv_bias=True,
                          norm_layer=partial(nn.LayerNorm, eps=1e-6), use_dense_prediction=use_dense_prediction, **spec)
    hk = head_kwargs or {}
    net.head = DINOHead(net.num_features, out_dim, norm_last_layer=norm_last_layer, **hk)
    if use_dense_prediction:
        net.head_dense = DINOHead(net.num_features, out_dim, norm_last_layer=norm_last_layer, **hk)
    return net


class _GradReducer:
    """Bucketed gradient all-reduce overlapped with backward - what DistributedDataParallel does for the reference
    (main_esvit.py:377), in a form that is captured inside the step's CUDA graph: parameters are bucketed in REVERSE
    registration order (= the order backward completes them: the two 65536-wide last layers, 45 % of all gradient bytes,
    first), a post-accumulate-grad hook counts completed gradients per bucket and, when a bucket is full, forks a side
    stream that flattens the bucket, AVG-all-reduces it over NCCL and scatters it back while the main stream continues
    with the remaining backward kernels.  finish() launches whatever is left and joins the side stream before the
    optimiser sweep.  The result is independent of the bucketing (same AVG per element)."""

    def __init__(self, params, bucket_bytes: int = 48 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self.count = [0] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self.side = torch.cuda.Stream()
        self.enabled = True
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    def _hook(self, p) -> None:
        if not self.enabled:
            return
        bi = self.bucket_of[id(p)]
        self.count[bi] += 1
        if self.count[bi] == len(self.buckets[bi]) and not self.launched[bi]:
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        self.launched[bi] = True
        grads = [p.grad for p in self.buckets[bi] if p.grad is not None]
        if not grads:
            return
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            torch._foreach_copy_(grads, [t.view_as(g) for t, g in zip(flat.split([g.numel() for g in grads]), grads)])

    def finish(self) -> None:
        for bi in range(len(self.buckets)):
            if not self.launched[bi]:
                self._launch(bi)
        torch.cuda.current_stream().wait_stream(self.side)
        self.count = [0] * len(self.buckets)
        self.launched = [False] * len(self.buckets)


class SelfDistillStep:
    """One training step.  optimizer: a torch optimizer (the reference's own sequence: clip kernel, cancel grads,
    optimizer.step(), EMA kernel) or an esvit_b200.optim.FusedAdamWEMA (clip + AdamW + EMA in one sweep).
    With use_cuda_graph=True (FusedAdamWEMA only) the whole step - teacher fwd, student fwd/bwd, loss, center update,
    gradient all-reduce, optimiser sweep - is captured once per (teacher_temp, last-layer-frozen) state and replayed:
    no Python / launch overhead in the steady state.  lr / wd / momentum live in device memory and are refreshed with
    an 32-byte async copy before each replay."""

    def __init__(self, student: nn.Module, teacher: nn.Module, loss: nn.Module, optimizer, clip_grad: float = 3.0,
                 freeze_last_layer: int = 1, student_ddp: Optional[nn.Module] = None, use_cuda_graph: bool = False,
                 grad_allreduce: bool = False):
        self.student, self.teacher, self.loss, self.opt = student, teacher, loss, optimizer
        self.student_call = student_ddp if student_ddp is not None else student
        self.clip_grad, self.freeze_last_layer = clip_grad, freeze_last_layer
        self.fused = isinstance(optimizer, FusedAdamWEMA)
        self.use_cuda_graph = use_cuda_graph and self.fused
        self.grad_allreduce = grad_allreduce  # own flat all-reduce of the gradients (used instead of DDP in graph mode)
        self.last_norms = None
        self._graphs = {}
        self._pool = None
        self._static_in = None
        self._static_loss = None
        self._warm = 0
        for p in self.teacher.parameters():
            p.requires_grad = False
        # arena for the ACCUMULATED small gradients (LN affine, biases, rel-pos tables, patch embed): sized from the model
        small = sum(p.numel() for p in self.student.parameters() if p.dim() == 1 or p.numel() <= (1 << 16))
        self._arena_floats = 4 * small + (1 << 18)
        self._reducer = None
        if self.grad_allreduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._reducer = _GradReducer(list(self.student.parameters()))

    # ---- the step body (eager; also what gets captured) ---------------------------------------------------
    def _body(self, images: List[torch.Tensor], epoch: int) -> torch.Tensor:
        with torch.no_grad():
            teacher_output = self.teacher(images[:2])
        student_output = self.student_call(images)
        loss = self.loss(student_output, teacher_output, epoch, None)
        if self.fused:
            self.opt.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=True)
        ops.begin_step(loss.device, self._arena_floats)  # one zero-filled arena for all small gradient accumulators
        try:
            loss.backward()
        finally:
            ops.end_step()
        self.reduce_gradients()
        if self.fused:
            self.opt.step()  # clip + AdamW + EMA, one sweep
        else:
            if self.clip_grad:
                self.last_norms = utils.clip_gradients(self.student, self.clip_grad)
            utils.cancel_gradients_last_layer(epoch, self.student, self.freeze_last_layer)
            self.opt.step()
        return loss.detach()

    def reduce_gradients(self) -> None:
        """DDP's gradient AVG all-reduce (main_esvit.py:377) for the graph-captured step: buckets whose gradients were
        complete during backward have already been launched on the side stream by the post-accumulate hooks
        (_GradReducer); this launches the rest and joins the side stream.  No-op at world size 1."""
        if self._reducer is not None:
            self._reducer.finish()

    def __call__(self, images: Sequence[torch.Tensor], epoch: int, lr: float, wd: float, momentum: float) -> torch.Tensor:
        images = list(images)
        if self.fused:
            self.opt.set_hyper(lr, wd, momentum)
            self.opt.set_skip_last_layer(epoch < self.freeze_last_layer)
        else:
            for i, g in enumerate(self.opt.param_groups):  # main_esvit.py:507-510
                g["lr"] = lr
                if i == 0:
                    g["weight_decay"] = wd
        if not self.use_cuda_graph:
            loss = self._body(images, epoch)
            if not self.fused:
                utils.ema_update(self.student, self.teacher, momentum)
            return loss
        return self._graphed(images, epoch)

    # ---- CUDA-graph path ------------------------------------------------------------------------------------
    def _graphed(self, images: List[torch.Tensor], epoch: int) -> torch.Tensor:
        """Replay path.  The graph reads PRIVATE static input buffers (allocated on first use); every call copies the
        caller's crops into them on the current stream, so the caller may recycle its own (prefetch) buffers freely.  A
        batch of a different shape (e.g. a shorter last batch) runs the eager body instead."""
        if self._static_in is None:
            # crops of one shape share one buffer, back to back: the multi-crop forward then views a resolution group
            # as one batch instead of concatenating it (ops.cat_adjacent)
            self._static_in, i = [], 0
            while i < len(images):
                j = i
                while j < len(images) and images[j].shape == images[i].shape and images[j].dtype == images[i].dtype:
                    j += 1
                buf = torch.empty((j - i,) + tuple(images[i].shape), dtype=images[i].dtype, device=images[i].device)
                self._static_in += [buf[k] for k in range(j - i)]
                i = j
        if len(images) != len(self._static_in) or any(s.shape != im.shape for s, im in zip(self._static_in, images)):
            return self._body(images, epoch)
        for s, im in zip(self._static_in, images):
            s.copy_(im, non_blocking=True)
        if self._warm < 3:  # eager warm-up (allocator, cuBLAS workspaces, cached tables) before any capture
            self._warm += 1
            return self._body(self._static_in, epoch)
        key = (float(self.loss.teacher_temp_schedule[epoch]), epoch < self.freeze_last_layer)
        ent = self._graphs.get(key)
        if ent is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._pool):
                out = self._body(self._static_in, epoch)
            if self._pool is None:
                self._pool = g.pool()
            ent = self._graphs[key] = (g, out)
        ent[0].replay()
        return ent[1]


def make_step(arch: str = "swin_tiny_w7", out_dim: int = 65536, ncrops: int = 10, dense: bool = True,
              device: str = "cuda", lr: float = 5e-4, weight_decay: float = 0.04, clip_grad: float = 3.0,
              freeze_last_layer: int = 1, drop_path: Optional[float] = None, img_size: int = 224,
              head_kwargs: Optional[dict] = None, spec: Optional[dict] = None, ddp: bool = False,
              teacher_temp: float = 0.04, seed: int = 0, optimizer: str = "fused", cuda_graph: bool = False):
    """Build student/teacher/loss/optimizer the way train_esvit does (main_esvit.py:235-435) and return
    (step, student, teacher, loss)."""
    spec = dict(spec if spec is not None else SWIN_SPECS[arch])
    if drop_path is not None:
        spec["drop_path_rate"] = drop_path
    torch.manual_seed(seed)
    student = build_network(spec, out_dim, dense, False, True, img_size, head_kwargs).to(device)
    teacher = build_network(spec, out_dim, dense, True, True, img_size, head_kwargs).to(device)
    teacher.load_state_dict(student.state_dict())
    Loss = DDINOLoss if dense else DINOLoss
    loss = Loss(out_dim, ncrops, teacher_temp, teacher_temp, 0, 100).to(device)
    student_ddp = None
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if ddp and multi and optimizer != "fused":
        student_ddp = nn.parallel.DistributedDataParallel(student, device_ids=[torch.cuda.current_device()])
    if optimizer == "fused":
        for p in teacher.parameters():
            p.requires_grad = False
        opt = FusedAdamWEMA(student, teacher, clip_grad=clip_grad)
    else:  # the reference's own optimizer object (main_esvit.py:410-415)
        opt = torch.optim.AdamW(utils.get_params_groups(student), fused=True)
    step = SelfDistillStep(student, teacher, loss, opt, clip_grad, freeze_last_layer, student_ddp,
                           use_cuda_graph=cuda_graph, grad_allreduce=(optimizer == "fused" and multi))
    return step, student, teacher, loss
