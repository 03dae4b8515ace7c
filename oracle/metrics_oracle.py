"""CPU restatement of the reference's ensemble verification scores -- TEST INFRASTRUCTURE ONLY.

Follows stockeh/swift ``src/swift/eval/metrics.py`` (pure torch/numpy arithmetic; the module itself also imports ezpz and
xarray for its CLI, which are not installed here):

  lat_weighted_rmse                :39-66    ensemble-mean latitude-weighted RMSE
  lat_weighted_crps                :69-106   fair CRPS: E|X - y| - sum_ij |X_i - X_j| / (2 N (N - 1))
  lat_weighted_spread_skill_ratio  :109-134  sqrt(mean(w * var_N)) / RMSE(ensemble mean)

Pinned: ``tests/golden/metrics.npz`` holds the outputs of the real reference functions on a seeded fixture
(``tests/golden/make_metrics_golden.py``); ``tests/test_ensemble_cpu.py`` checks this file against it.
Only tests, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU leg may import it.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch


def lat_weights(lat: np.ndarray) -> torch.Tensor:
    """cos(lat) normalised to mean 1 (metrics.py:51-52, :79-80, :122-123)."""
    w = np.cos(np.deg2rad(np.asarray(lat, dtype=np.float64)))
    return torch.from_numpy(w / w.mean())


def rmse(pred: torch.Tensor, y: torch.Tensor, lat) -> torch.Tensor:
    """pred [B, N, V, H, W] (or [B, V, H, W]), y [B, V, H, W] -> [V]  (metrics.py:48-65)."""
    if pred.ndim == 5:
        pred = pred.mean(dim=1)
    w = lat_weights(lat).to(pred.dtype).view(1, 1, -1, 1)
    err = (pred - y) ** 2
    return torch.sqrt((err * w).mean(dim=(-2, -1))).mean(dim=0)


def crps(pred: torch.Tensor, y: torch.Tensor, lat) -> torch.Tensor:
    """pred [B, N, V, H, W], y [B, V, H, W] -> [V]  (metrics.py:82-100)."""
    B, N, V, H, W = pred.shape
    w = lat_weights(lat).to(pred.dtype)
    out = []
    for v in range(V):
        p, t = pred[:, :, v], y[:, v]
        err = (torch.abs(p - t.unsqueeze(1)) * w.view(1, 1, H, 1)).mean()
        spread = torch.abs(p.unsqueeze(2) - p.unsqueeze(1)) * w.view(1, 1, 1, H, 1)
        spread = spread.mean(dim=(-2, -1)).sum(dim=(1, 2)) / (2 * N * (N - 1))
        out.append(err - spread.mean())
    return torch.stack(out)


def spread_skill(pred: torch.Tensor, y: torch.Tensor, lat) -> torch.Tensor:
    """pred [B, N, V, H, W], y [B, V, H, W] -> [V]  (metrics.py:116-132)."""
    H = pred.shape[-2]
    w = lat_weights(lat).to(pred.dtype)
    var = torch.var(pred, dim=1) * w.view(1, 1, H, 1)
    spread = var.mean(dim=(-2, -1)).sqrt().mean(dim=0)
    return spread / rmse(pred.mean(dim=1), y, lat)


def all_scores(pred: torch.Tensor, y: torch.Tensor, vars: Sequence[str], lat, postfix: str) -> Dict[str, float]:
    """The reference's key schema: ``{crps,rmse,ssr}_{var}_{postfix}`` (metrics.py:62, :104, :130)."""
    r, c, s = rmse(pred, y, lat), crps(pred, y, lat), spread_skill(pred, y, lat)
    out = {}
    for i, v in enumerate(vars):
        out[f"crps_{v}_{postfix}"] = float(c[i])
        out[f"rmse_{v}_{postfix}"] = float(r[i])
        out[f"ssr_{v}_{postfix}"] = float(s[i])
    return out
