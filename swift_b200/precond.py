"""``PassPrecond`` with the reference's constructor and attributes (stockeh/swift ``models/precond.py:101-151``).

The reference's own ``PassPrecond`` can wrap ``swift_b200.swinv2.SwinV2`` unchanged (that is the drop-in path:
only ``model._target_`` changes).  This mirror exists so the path also runs where the reference package is not
installed (the GPU box), and so the samplers can use the fused entry point: channel concat, input scaling and the
sampler update all happen inside the CUDA forward instead of as separate PyTorch ops.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .config import instantiate


def _2d_resolution(x):
    if isinstance(x, int):
        return np.array([x, x], dtype=int)
    x = np.asarray(x, dtype=int)
    assert x.shape[0] == 2
    return x


def process_auxiliary(auxiliary, auxiliary_dim: int, batch_size: int, device) -> Optional[torch.Tensor]:
    """models/precond.py:21-31: None -> zeros [1, aux_dim]; scalar / len-1 -> repeated to the batch; -> [B, aux_dim]."""
    if auxiliary_dim == 0:
        return None
    if auxiliary is None:
        return torch.zeros([1, auxiliary_dim], device=device)
    if not isinstance(auxiliary, torch.Tensor):
        auxiliary = torch.tensor(auxiliary, device=device)
    if auxiliary.dim() == 0 or (auxiliary.dim() == 1 and auxiliary.size(0) == 1):
        auxiliary = auxiliary.repeat(batch_size)
    return auxiliary.reshape(-1, auxiliary_dim)


class PassPrecond(torch.nn.Module):
    def __init__(self, model_config, img_resolution, img_channels: int, condition_channels: int = 0,
                 auxiliary_dim: int = 0, sigma_min: float = 0.0, sigma_max: float = float("inf"),
                 sigma_data: float = 1.0):
        super().__init__()
        self.img_resolution = _2d_resolution(img_resolution)
        self.img_channels = img_channels
        self.condition_channels = condition_channels
        self.auxiliary_dim = auxiliary_dim
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.sigma_data = sigma_data
        self.model_config = model_config
        res = [int(v) for v in self.img_resolution]
        self.model = instantiate(model_config, img_resolution=res, in_channels=img_channels + condition_channels,
                                 out_channels=img_channels, auxiliary_dim=auxiliary_dim, _convert_="object")

    def forward(self, x, t, condition=None, auxiliary=None, **model_kwargs):
        auxiliary = process_auxiliary(auxiliary, self.auxiliary_dim, x.size(0), x.device)
        if auxiliary is not None and auxiliary.shape[0] == 1 and x.size(0) > 1:
            auxiliary = auxiliary.expand(x.size(0), -1)      # the [1, aux_dim] zeros of the `None` case broadcast
        arg = x
        if condition is not None and self.condition_channels > 0:
            arg = torch.cat([arg, condition], dim=1)
        return self.model(arg, t.flatten(), auxiliary=auxiliary, **model_kwargs)

    def round_sigma(self, sigma):
        return torch.as_tensor(sigma)
