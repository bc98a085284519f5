"""Step helpers behind the reference's names (utils.py): MultiCropWrapper, clip_gradients,
cancel_gradients_last_layer, get_params_groups, cosine_scheduler - plus the multi-tensor teacher EMA that
replaces the per-parameter loop of main_esvit.py:587-590.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops


class MultiCropWrapper(nn.Module):
    """utils.MultiCropWrapper (utils.py:560-617): consecutive same-resolution crops are concatenated on the batch
    axis, the backbone runs once per resolution, heads run once on the concatenated features.

    The backbone must expose ``forward_features(x)`` returning ``pooled`` or ``(pooled, region)`` (dense mode), or be
    a plain callable returning ``pooled`` in view mode - the contract of the reference's ResNetWrapper."""

    def __init__(self, backbone, head, head_dense=None, use_dense_prediction=False):
        super().__init__()
        if hasattr(backbone, "fc"):
            backbone.fc = nn.Identity()
        self.backbone = backbone
        self.head = head
        self.use_dense_prediction = use_dense_prediction
        self.head_dense = head_dense

    def forward(self, x):
        if not isinstance(x, list):
            x = [x]
        groups, start = [], 0
        for i in range(1, len(x) + 1):
            if i == len(x) or x[i].shape[-1] != x[start].shape[-1]:
                groups.append((start, i))
                start = i
        if self.use_dense_prediction:
            cls_l, fea_l, npatch = [], [], []
            for s, e in groups:
                out_cls, out_fea = self.backbone.forward_features(torch.cat(x[s:e]))
                B, N, C = out_fea.shape
                cls_l.append(out_cls)
                fea_l.append(out_fea.reshape(B * N, C))
                npatch.append(N)
            output_cls, output_fea = torch.cat(cls_l), torch.cat(fea_l)
            return self.head(output_cls), self.head_dense(output_fea), output_fea, npatch
        outs = [self.backbone(torch.cat(x[s:e])) for s, e in groups]
        return self.head(torch.cat(outs))


def clip_gradients(model: nn.Module, clip: float) -> torch.Tensor:
    """utils.clip_gradients (utils.py:106-115): PER-PARAMETER L2 clipping, as ONE multi-tensor kernel pair and
    no host synchronisation.  Returns the pre-clip norms as a device tensor (the reference returns a python list
    that train_one_epoch drops, main_esvit.py:580)."""
    grads = [p.grad for _, p in model.named_parameters() if p.grad is not None]
    if not grads:
        return torch.empty(0)
    return ops.clip_grads_(grads, clip)


def cancel_gradients_last_layer(epoch: int, model: nn.Module, freeze_last_layer: int) -> None:
    """utils.py:118-123."""
    if epoch >= freeze_last_layer:
        return
    for n, p in model.named_parameters():
        if "last_layer" in n:
            p.grad = None


def get_params_groups(model: nn.Module):
    """utils.py:672-683: biases and 1-D parameters are not weight-decayed."""
    regularized, not_regularized = [], []
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if name.endswith(".bias") or len(param.shape) == 1:
            not_regularized.append(param)
        else:
            regularized.append(param)
    return [{'params': regularized}, {'params': not_regularized, 'weight_decay': 0.}]


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0):
    """utils.py:161-173."""
    warmup_schedule = np.array([])
    warmup_iters = warmup_epochs * niter_per_ep
    if warmup_epochs > 0:
        warmup_schedule = np.linspace(start_warmup_value, base_value, warmup_iters)
    iters = np.arange(epochs * niter_per_ep - warmup_iters)
    schedule = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * iters / len(iters)))
    schedule = np.concatenate((warmup_schedule, schedule))
    assert len(schedule) == epochs * niter_per_ep
    return schedule


@torch.no_grad()
def ema_update(student: nn.Module, teacher: nn.Module, momentum: float) -> None:
    """Teacher EMA over zip(student.parameters(), teacher.parameters()) (parameters only, not buffers),
    main_esvit.py:587-590, as one multi-tensor launch; bit-exact with the reference loop."""
    q = [p.detach() for p in student.parameters()]
    k = [p.detach() for p in teacher.parameters()]
    ops.ema_update_(k, q, momentum)
