"""Fused optimiser pass for the self-distillation step: per-tensor gradient clip (utils.clip_gradients) + AdamW
(torch.optim.AdamW semantics with the reference's two param groups, utils.get_params_groups) + teacher EMA
(main_esvit.py:587-590) as ONE multi-tensor sweep (esvit_adamw_ema_multi), with every step-varying scalar in device
memory so that the whole training step can be captured in a CUDA graph and replayed with fresh lr / wd / momentum.

Semantics preserved from the reference loop:
  * clipping is PER PARAMETER TENSOR (coef = clip / (norm + 1e-6), applied when < 1);
  * biases and 1-D parameters are not weight-decayed;
  * parameters whose gradient the reference sets to None (cancel_gradients_last_layer while epoch <
    freeze_last_layer) are skipped entirely by AdamW (no decay, no moments, no step count) - flag bit 1;
  * frozen parameters (requires_grad=False, e.g. last_layer.weight_g) are never stepped but ARE part of the EMA,
    which runs over zip(student.parameters(), teacher.parameters());
  * the EMA is bit-exact with param_k.mul_(m).add_((1 - m) * param_q).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib, ops, shadow

F32 = torch.float32


class FusedAdamWEMA:
    def __init__(self, student: nn.Module, teacher: Optional[nn.Module], betas=(0.9, 0.999), eps: float = 1e-8,
                 clip_grad: float = 3.0):
        named = list(student.named_parameters())
        self.names = [n for n, _ in named]
        self.params: List[torch.Tensor] = [p for _, p in named]
        tparams = [p for _, p in teacher.named_parameters()] if teacher is not None else None
        if tparams is not None:
            assert len(tparams) == len(self.params)
        dev = self.params[0].device
        for p in self.params:
            if not (p.is_cuda and p.dtype == F32 and p.is_contiguous()):
                raise RuntimeError("FusedAdamWEMA needs fp32 contiguous CUDA parameters (no CPU fallback)")
        self.teacher_params = tparams
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        n = len(self.params)
        self.frozen = [not p.requires_grad for p in self.params]
        decay = [0 if (nm.endswith(".bias") or p.dim() == 1) else 1 for nm, p in named]  # utils.py:672-683
        self._base_flags = torch.tensor([d | (2 if f else 0) for d, f in zip(decay, self.frozen)], dtype=F32)
        self._last_layer = torch.tensor([1.0 if "last_layer" in nm else 0.0 for nm in self.names])
        self.state = torch.zeros(n, 2, dtype=F32, device=dev)
        self.state[:, 1] = self._base_flags.to(dev)
        self._flags_host = torch.empty(n, dtype=F32).pin_memory()
        self.hyper = torch.zeros(8, dtype=F32, device=dev)
        self._hyper_host = torch.zeros(8, dtype=F32).pin_memory()
        self._hyper_host[2], self._hyper_host[3], self._hyper_host[4] = betas[0], betas[1], eps
        self._hyper_host[7] = clip_grad if clip_grad else 0.0
        self.sumsq = torch.zeros(n, dtype=torch.float64, device=dev)
        # bf16 shadows of everything the GEMMs consume (Linear weights and biases); written by the sweep itself
        def wants_shadow(nm: str, p: torch.Tensor) -> bool:
            if "relative_position_bias_table" in nm or "patch_embed.proj" in nm or "last_layer" in nm or "norm" in nm:
                return False
            return nm.endswith(".weight") and p.dim() == 2 or nm.endswith(".bias")

        self.shadow_p = [shadow.register(p) if wants_shadow(nm, p) else None for nm, p in named]
        self.shadow_k = None
        if tparams is not None:
            self.shadow_k = [shadow.register(k) if wants_shadow(nm, k) else None for (nm, _), k in zip(named, tparams)]
        self._numel = ops._numel_array(self.params)
        self._ptr = {k: ops._ptr_array(v) for k, v in (("p", self.params), ("m", self.exp_avg), ("v", self.exp_avg_sq))}
        self._ptr["k"] = ops._ptr_array(self.teacher_params) if tparams is not None else None
        self._ptr["sp"] = self._opt_ptr_array(self.shadow_p)
        self._ptr["sk"] = self._opt_ptr_array(self.shadow_k) if self.shadow_k is not None else None
        self._skip_last = None

    @staticmethod
    def _opt_ptr_array(tensors):
        arr = (ctypes.c_void_p * len(tensors))()
        for i, t in enumerate(tensors):
            arr[i] = t.data_ptr() if t is not None else None
        return arr

    # ---- host-side knobs (tiny async H2D copies; safe between graph replays) ---------------------------------
    def set_hyper(self, lr: float, weight_decay: float, momentum: float) -> None:
        h = self._hyper_host
        h[0], h[1] = lr, weight_decay
        h[5], h[6] = momentum, 1.0 - momentum  # float32(m), float32(1 - m) computed in double like the reference
        self.hyper.copy_(h, non_blocking=True)

    def set_skip_last_layer(self, skip: bool) -> None:
        """epoch < freeze_last_layer  <=>  the reference sets last_layer grads to None (utils.py:118-123)."""
        if skip == self._skip_last:
            return
        self._skip_last = skip
        f = self._base_flags.clone()
        if skip:
            f = torch.where(self._last_layer > 0, torch.tensor(2.0) + (f % 2), f)
        self._flags_host.copy_(f)
        self.state[:, 1].copy_(self._flags_host, non_blocking=True)

    def zero_grad(self) -> None:
        """set_to_none: autograd then writes each gradient straight into a fresh buffer (no += pass).  Under CUDA-graph
        capture those buffers come from the graph's private pool, so their addresses are stable across replays."""
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self) -> None:
        for p in self.params:
            g = p.grad
            if g is not None and not (g.dtype == F32 and g.is_contiguous()):
                p.grad = g.float().contiguous()
        gs = [(p.grad if p.grad is not None else p) for p in self.params]  # frozen / cancelled: dummy pointer, never read
        for i, p in enumerate(self.params):
            if p.grad is None and not self.frozen[i] and not (self._skip_last and self._last_layer[i] > 0):
                raise RuntimeError(f"parameter {self.names[i]} received no gradient this step")
        gptr = ops._ptr_array(gs)
        st = ops._stream()
        n = len(self.params)
        _lib.call("esvit_grad_sumsq_multi", gptr, self._numel, n, ops._p(self.sumsq), st)
        _lib.call("esvit_adamw_ema_multi", self._ptr["p"], gptr, self._ptr["m"], self._ptr["v"], self._ptr["k"],
                  self._ptr["sp"], self._ptr["sk"], self._numel, n, ops._p(self.hyper), ops._p(self.state),
                  ops._p(self.sumsq), st)

    def grad_norms(self) -> torch.Tensor:
        """pre-clip per-tensor gradient norms of the last step (device tensor; what clip_gradients returned)."""
        return self.sumsq.sqrt().float()

    def state_dict(self) -> Dict:
        return {"names": self.names, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "state": self.state}

    def load_state_dict(self, sd: Dict) -> None:
        for a, b in zip(self.exp_avg, sd["exp_avg"]):
            a.copy_(b)
        for a, b in zip(self.exp_avg_sq, sd["exp_avg_sq"]):
            a.copy_(b)
        self.state.copy_(sd["state"])
