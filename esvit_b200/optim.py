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
        # ring of pinned staging buffers: the H2D copy of step N reads its buffer at EXECUTION time, while the host may
        # already be preparing step N+k (no sync inside the step) - a single buffer would be overwritten in flight
        self._hyper_ring = [torch.zeros(8, dtype=F32).pin_memory() for _ in range(8)]
        self._hyper_events = [None] * 8
        self._hyper_slot = 0
        self._hyper_host = torch.zeros(8, dtype=F32)
        self._hyper_host[2], self._hyper_host[3], self._hyper_host[4] = betas[0], betas[1], eps
        self._hyper_host[7] = clip_grad if clip_grad else 0.0
        self.betas, self.eps, self.clip_grad = tuple(betas), eps, clip_grad
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
        i = self._hyper_slot
        self._hyper_slot = (i + 1) % len(self._hyper_ring)
        ev = self._hyper_events[i]
        if ev is not None:
            ev.synchronize()  # the copy that last used this slot has executed (8 steps ago: never waits in practice)
        self._hyper_ring[i].copy_(h)
        self.hyper.copy_(self._hyper_ring[i], non_blocking=True)
        ev = self._hyper_events[i] = ev if ev is not None else torch.cuda.Event()
        ev.record()

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

    # ---- checkpoint interchange with the reference's `optimizer` entry (torch.optim.AdamW over utils.get_params_groups) ----
    def _torch_order(self):
        """indices of this optimiser's parameters in torch's numbering: group 0 = regularised, group 1 = biases / 1-D
        (utils.py:672-683), trainable parameters only, each in named_parameters order"""
        reg = [i for i, (nm, p) in enumerate(zip(self.names, self.params))
               if not self.frozen[i] and not (nm.endswith(".bias") or p.dim() == 1)]
        noreg = [i for i, (nm, p) in enumerate(zip(self.names, self.params))
                 if not self.frozen[i] and (nm.endswith(".bias") or p.dim() == 1)]
        return reg, noreg

    def state_dict(self) -> Dict:
        """torch.optim.AdamW layout: state[idx] = {step, exp_avg, exp_avg_sq} (clones) + the two param_groups, so the
        reference's restart_from_checkpoint / a plain torch AdamW can load it."""
        reg, noreg = self._torch_order()
        h = self._hyper_host
        state = {}
        for k, i in enumerate(reg + noreg):
            state[k] = {"step": self.state[i, 0].detach().clone().cpu(), "exp_avg": self.exp_avg[i].detach().clone(),
                        "exp_avg_sq": self.exp_avg_sq[i].detach().clone()}
        common = dict(lr=float(h[0]), betas=self.betas, eps=self.eps, amsgrad=False, maximize=False, foreach=None,
                      capturable=False, differentiable=False, fused=None, decoupled_weight_decay=True)
        groups = [dict(common, weight_decay=float(h[1]), params=list(range(len(reg)))),
                  dict(common, weight_decay=0.0, params=list(range(len(reg), len(reg) + len(noreg))))]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: Dict, **kwargs) -> None:
        """accepts the torch AdamW layout (reference checkpoints) and this class's round-1 private layout"""
        if "exp_avg" in sd:  # round-1 private format
            for a, b in zip(self.exp_avg, sd["exp_avg"]):
                a.copy_(b)
            for a, b in zip(self.exp_avg_sq, sd["exp_avg_sq"]):
                a.copy_(b)
            self.state.copy_(sd["state"])
            return
        reg, noreg = self._torch_order()
        order = reg + noreg
        for k, st in sd["state"].items():
            i = order[int(k)]
            self.exp_avg[i].copy_(st["exp_avg"])
            self.exp_avg_sq[i].copy_(st["exp_avg_sq"])
            self.state[i, 0] = float(st["step"])
        g = sd.get("param_groups")
        if g:
            self._hyper_host[0] = g[0].get("lr", float(self._hyper_host[0]))
            self._hyper_host[1] = g[0].get("weight_decay", float(self._hyper_host[1]))
