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

from . import utils
from .losses import DDINOLoss, DINOLoss
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


class SelfDistillStep:
    def __init__(self, student: nn.Module, teacher: nn.Module, loss: nn.Module, optimizer: torch.optim.Optimizer,
                 clip_grad: float = 3.0, freeze_last_layer: int = 1, student_ddp: Optional[nn.Module] = None):
        self.student, self.teacher, self.loss, self.opt = student, teacher, loss, optimizer
        self.student_call = student_ddp if student_ddp is not None else student
        self.clip_grad, self.freeze_last_layer = clip_grad, freeze_last_layer
        self.last_norms = None
        for p in self.teacher.parameters():
            p.requires_grad = False

    def __call__(self, images: Sequence[torch.Tensor], epoch: int, lr: float, wd: float, momentum: float) -> torch.Tensor:
        for i, g in enumerate(self.opt.param_groups):  # main_esvit.py:507-510
            g["lr"] = lr
            if i == 0:
                g["weight_decay"] = wd
        images = list(images)
        with torch.no_grad():
            teacher_output = self.teacher(images[:2])
        student_output = self.student_call(images)
        loss = self.loss(student_output, teacher_output, epoch, None)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.clip_grad:
            self.last_norms = utils.clip_gradients(self.student, self.clip_grad)
        utils.cancel_gradients_last_layer(epoch, self.student, self.freeze_last_layer)
        self.opt.step()
        utils.ema_update(self.student, self.teacher, momentum)
        return loss.detach()


def make_step(arch: str = "swin_tiny_w7", out_dim: int = 65536, ncrops: int = 10, dense: bool = True,
              device: str = "cuda", lr: float = 5e-4, weight_decay: float = 0.04, clip_grad: float = 3.0,
              freeze_last_layer: int = 1, drop_path: Optional[float] = None, img_size: int = 224,
              head_kwargs: Optional[dict] = None, spec: Optional[dict] = None, ddp: bool = False,
              teacher_temp: float = 0.04, seed: int = 0):
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
    if ddp:
        student_ddp = nn.parallel.DistributedDataParallel(student, device_ids=[torch.cuda.current_device()])
    opt = torch.optim.AdamW(utils.get_params_groups(student), fused=True)
    step = SelfDistillStep(student, teacher, loss, opt, clip_grad, freeze_last_layer, student_ddp)
    return step, student, teacher, loss
