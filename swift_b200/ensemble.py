"""Ensemble verification statistics for resident trajectories.

The reference scores a finished forecast off-line: ``generate.py`` writes every member to zarr and
``python -m swift.eval.metrics`` (``eval/metrics.py:39-134``) re-reads it to compute the latitude-weighted ensemble-mean
RMSE, the fair CRPS and the spread/skill ratio per variable and lead time.  Here the per-(initial condition, variable)
sufficient statistics of those three scores are accumulated by one CUDA kernel per 6 h step from the physical state the
rollout step has just produced (inside the same replayed CUDA graph), and only those sums -- ``steps x ICs x variables
x 4`` float64 values -- are exchanged between GPUs: one ``all_gather`` (NCCL over NVLink) at the end of the rollout,
the single collective of the forecast path.  All members of an initial condition live on one GPU
(``rollout.shard_trajectories`` is IC-major), so no field ever crosses a link.

    stats = EnsembleStatistics(members, n_ic_local, n_var, (H, W), lat, steps, device)
    rollout.attach_statistics(stats, truth_buffer)      # before the first step
    ... rollout.run(steps) ...
    scores = stats.scores(stats.gather())               # {"rmse": [steps, V], "crps": ..., "ssr": ...}
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def lat_weights(lat) -> torch.Tensor:
    """cos(lat) / mean(cos(lat)) as float64 (eval/metrics.py:51-52)."""
    w = np.cos(np.deg2rad(np.asarray(lat, dtype=np.float64)))
    return torch.from_numpy(w / w.mean())


class EnsembleStatistics:
    def __init__(self, members: int, n_ic: int, n_var: int, resolution: Tuple[int, int], lat, steps: int, device):
        if members < 2:
            raise ValueError("ensemble scores need at least 2 members (CRPS spread term divides by N - 1, "
                             "eval/metrics.py:97)")
        self.members, self.n_ic, self.n_var, self.steps = int(members), int(n_ic), int(n_var), int(steps)
        self.res = (int(resolution[0]), int(resolution[1]))
        if len(lat) != self.res[0]:
            raise ValueError(f"lat has {len(lat)} entries for {self.res[0]} grid rows")
        self.device = torch.device(device)
        self.w_lat = lat_weights(lat).to(device=self.device, dtype=torch.float32).contiguous()
        # [step, ic, var, 4]: sum w (mean - y)^2 | sum_n sum w |p_n - y| | sum_{i<j} sum w |p_i - p_j| | sum w var_n
        self.sums = torch.zeros(self.steps, self.n_ic, self.n_var, 4, dtype=torch.float64, device=self.device)

    # ------------------------------------------------------------------ device side
    def accumulate(self, phys: torch.Tensor, truth: torch.Tensor, step_dev: Optional[torch.Tensor] = None,
                   step: int = 0) -> None:
        """phys [n_ic*members, V, H, W] (IC-major), truth [n_ic, V, H, W]; the row of ``sums`` is ``*step_dev`` (device
        int32, graph-replay friendly) or the host integer ``step``."""
        if self.device.type != "cuda":
            raise RuntimeError("EnsembleStatistics.accumulate runs on CUDA only (no CPU fallback); "
                               "use sums_reference() in tests")
        exp_p = (self.n_ic * self.members, self.n_var, *self.res)
        exp_t = (self.n_ic, self.n_var, *self.res)
        for name, t, exp in (("phys", phys, exp_p), ("truth", truth, exp_t)):
            if tuple(t.shape) != exp or t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                raise RuntimeError(f"{name}: expected contiguous float32 CUDA tensor {exp}, got {tuple(t.shape)} {t.dtype}")
        out = self.sums if step_dev is not None else self.sums[step]
        stride = self.n_ic * self.n_var * 4
        _lib.check(_lib.lib().swb200_ensemble_stats(phys.data_ptr(), truth.data_ptr(), self.w_lat.data_ptr(), self.n_ic,
                                                    self.members, self.n_var, self.res[0], self.res[1],
                                                    _lib.ptr(step_dev), self.steps, stride, out.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream), "ensemble_stats")

    # ------------------------------------------------------------------ exchange
    def gather(self, group=None) -> torch.Tensor:
        """All ranks' sums concatenated along the IC axis -> [steps, n_ic_total, V, 4] (every rank gets the result).
        The only collective of the forecast path: NCCL all_gather of steps*n_ic*V*32 bytes per rank."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self.sums
        world = dist.get_world_size(group)
        counts = [torch.zeros(1, dtype=torch.int64, device=self.sums.device) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([self.n_ic], dtype=torch.int64, device=self.sums.device), group=group)
        counts = [int(c.item()) for c in counts]
        mx = max(counts)
        mine = self.sums
        if self.n_ic < mx:                       # ragged shards: pad to the largest
            pad = torch.zeros(self.steps, mx - self.n_ic, self.n_var, 4, dtype=mine.dtype, device=mine.device)
            mine = torch.cat([mine, pad], dim=1)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.contiguous(), group=group)
        return torch.cat([p[:, :c] for p, c in zip(parts, counts)], dim=1)

    # ------------------------------------------------------------------ scores
    def scores(self, sums: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """eval/metrics.py:39-134 from the sufficient statistics -> {"rmse", "crps", "ssr"} each [steps, V] float64."""
        s = self.sums if sums is None else sums
        return scores_from_sums(s, self.members, self.res[0] * self.res[1])

    def as_reference_dict(self, scores: Dict[str, torch.Tensor], vars: Sequence[str], lead_hours: Sequence[int]) -> Dict[str, float]:
        """The reference's flat key schema ``{metric}_{var}_{lead}h`` (eval/metrics.py:62, :104, :130)."""
        out = {}
        for k, lead in enumerate(lead_hours):
            for i, v in enumerate(vars):
                for m in ("crps", "rmse", "ssr"):
                    out[f"{m}_{v}_{lead}h"] = float(scores[m][k, i])
        return out


def scores_from_sums(sums: torch.Tensor, members: int, hw: int) -> Dict[str, torch.Tensor]:
    s = sums.to(torch.float64)
    n = float(members)
    rmse = torch.sqrt(s[..., 0] / hw).mean(dim=-2)                                   # mean over ICs of per-IC RMSE
    crps = (s[..., 1] / (n * hw) - s[..., 2] / (n * (n - 1) * hw)).mean(dim=-2)
    ssr = torch.sqrt(s[..., 3] / hw).mean(dim=-2) / rmse
    return {"rmse": rmse, "crps": crps, "ssr": ssr}


def sums_reference(pred: torch.Tensor, truth: torch.Tensor, lat) -> torch.Tensor:
    """Plain-torch statement of what the kernel accumulates: pred [B, N, V, H, W], truth [B, V, H, W] -> [B, V, 4].
    Host-side helper for tests of the score algebra (the CUDA kernel is the product path)."""
    w = lat_weights(lat).to(pred.dtype).view(1, 1, -1, 1)
    mean = pred.mean(dim=1)
    s0 = (w * (mean - truth) ** 2).sum(dim=(-2, -1))
    s1 = (w.unsqueeze(1) * (pred - truth.unsqueeze(1)).abs()).sum(dim=(1, -2, -1))
    diff = (pred.unsqueeze(2) - pred.unsqueeze(1)).abs() * w.view(1, 1, 1, 1, -1, 1)
    s2 = diff.sum(dim=(1, 2, -2, -1)) / 2
    s3 = (w * pred.var(dim=1)).sum(dim=(-2, -1))
    return torch.stack([s0, s1, s2, s3], dim=-1)
