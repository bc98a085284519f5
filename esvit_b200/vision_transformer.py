"""DINOHead behind the reference's signature and state_dict keys (models/vision_transformer.py:384-418).

mlp.{0,2,4} Linear(+exact GELU) -> L2 normalise -> weight-normed Linear(bottleneck, out_dim, bias=False), with
parameters ``mlp.N.{weight,bias}``, ``last_layer.weight_g`` [K,1], ``last_layer.weight_v`` [K,D].
The three MLP GEMMs and the last-layer GEMM are bf16 library GEMMs; GELU, the row normalisation and the
weight-norm reparameterisation (fwd + bwd) are esvit_b200 kernels.  Output logits are bf16 [rows, out_dim]
(what the reference produces under autocast); the losses consume them without an fp32 copy.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import linear, ops, shadow
from .swin_transformer import USE_GEMM2

BF16 = torch.bfloat16


class _WeightNormLinear(nn.Module):
    """Parameter container with the names nn.utils.weight_norm(nn.Linear(..., bias=False)) registers."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        lin = nn.Linear(in_features, out_features, bias=False)  # same default init stream as the reference
        self.weight_g = nn.Parameter(lin.weight.detach().norm(2, dim=1, keepdim=True))
        self.weight_v = nn.Parameter(lin.weight.detach().clone())

    def effective_weight(self) -> torch.Tensor:
        return ops.WeightNormFn.apply(self.weight_v, self.weight_g)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if USE_GEMM2:
            return linear.LastLayerFn.apply(x, self.effective_weight())
        with torch.autocast("cuda", enabled=False):
            return F.linear(x, self.effective_weight())


class DINOHead(nn.Module):
    def __init__(self, in_dim, out_dim, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048,
                 bottleneck_dim=256):
        super().__init__()
        if use_bn:
            raise NotImplementedError("use_bn_in_head is False in every EsViT recipe")
        nlayers = max(nlayers, 1)
        if nlayers == 1:
            self.mlp = nn.Linear(in_dim, bottleneck_dim)
        else:
            layers = [nn.Linear(in_dim, hidden_dim), nn.GELU()]
            for _ in range(nlayers - 2):
                layers += [nn.Linear(hidden_dim, hidden_dim), nn.GELU()]
            layers.append(nn.Linear(hidden_dim, bottleneck_dim))
            self.mlp = nn.Sequential(*layers)
        self.apply(self._init_weights)
        self.last_layer = _WeightNormLinear(bottleneck_dim, out_dim)
        self.last_layer.weight_g.data.fill_(1)
        if norm_last_layer:
            self.last_layer.weight_g.requires_grad = False

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(BF16)
        mods = [self.mlp] if isinstance(self.mlp, nn.Linear) else list(self.mlp)
        lins = [m for m in mods if isinstance(m, nn.Linear)]
        if USE_GEMM2 and len(lins) == 3 and len(mods) == 5:
            args = []
            for m in lins:
                args += [m.weight, shadow.as_bf16(m.weight, track_grad=False), m.bias]
            x = linear.HeadMlpFn.apply(x, *args)
            return self.last_layer(ops.L2NormFn.apply(x, 1e-12))
        with torch.autocast("cuda", enabled=False):
            for i, m in enumerate(mods):
                if not isinstance(m, nn.Linear):
                    continue
                if i + 1 < len(mods) and isinstance(mods[i + 1], nn.GELU):
                    # GEMM with bias epilogue; exact GELU kernel whose backward also yields the bias gradient
                    x = ops.BiasGeluFn.apply(
                        ops.LinearBiasFn.apply(x, shadow.as_bf16(m.weight), shadow.as_bf16(m.bias, False)), m.bias)
                else:
                    x = F.linear(x, shadow.as_bf16(m.weight), shadow.as_bf16(m.bias))
        x = ops.L2NormFn.apply(x, 1e-12)
        return self.last_layer(x)
