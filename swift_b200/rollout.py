"""Autoregressive ensemble rollout: the loop ``generate.rollout_and_save`` runs around the sampler
(stockeh/swift ``src/swift/generate.py:79-136``), restated for resident device state.

Per 6 h step and trajectory (member m, initial condition j) the reference does (residual=True branch):

    X      <- cat([X, standardize_x(forcings(step))], 1)          generate.py:100-117
    Y       = sampler(X, generator)                               :118   (fresh N(0,1) latents, one denoiser call)
    X_phys  = unstandardize_x(X)[:, :n_var] + unstandardize_t(Y)  :120-126  (t_means = 0: era5.py:95-100)
    rollout[:, i+1] = X_phys.cpu()                                :129
    X      <- standardize_x(X_phys)                               :131

Differences, all documented in DESIGN.md:
  * trajectories are sharded by flattened (IC, member) index over ranks (the reference shards by member only and
    leaves 4 of 8 ranks half idle for 12 members); no collective is needed on the forecast path;
  * every trajectory owns its noise stream, ``torch.Generator(device).manual_seed(seed_of(m, j))``, so a
    trajectory's result does not depend on world size or batching (the reference's per-member generator is consumed
    in batch order);
  * forcings are pre-staged on the device as a [steps, n_forcings, H, W] table instead of one HDF5 read per sample
    per step on the main thread.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from .sampler import DiffusionSampler


def shard_trajectories(members: int, n_ic: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """(member, ic) pairs owned by ``rank``: contiguous blocks of the IC-major flattened index j*members + m, so the
    members of one initial condition stay on one GPU whenever n_ic is a multiple of world."""
    total = members * n_ic
    per, rem = divmod(total, world)
    lo = rank * per + min(rank, rem)
    hi = lo + per + (1 if rank < rem else 0)
    return [(idx % members, idx // members) for idx in range(lo, hi)]


def trajectory_seed(member: int, ic: int) -> int:
    """Seed of the (member, ic) noise stream; member m of every IC shares the reference's ``manual_seed(m)`` family."""
    return member + 1_000_003 * ic


@dataclass
class Normalizers:
    """Per-channel affine maps of data/era5.py:80-108 as [1, C, 1, 1] device tensors."""
    x_mean: torch.Tensor      # variables only
    x_std: torch.Tensor
    diff_std: torch.Tensor    # t_stds[interval] (t_means are zero for residual targets)

    @staticmethod
    def synthetic(n_var: int, device, diff: float = 0.1) -> "Normalizers":
        one = torch.ones(1, n_var, 1, 1, device=device)
        return Normalizers(torch.zeros_like(one), one, diff * one)


class EnsembleRollout:
    """Advances a batch of independent trajectories that live on one GPU."""

    def __init__(self, net, norm: Normalizers, forcings_std: torch.Tensor, trajectories: Sequence[Tuple[int, int]],
                 solver: str = "scm", solver_kwargs: Optional[dict] = None):
        self.net = net
        self.norm = norm
        self.forcings = forcings_std              # [steps(+), n_forc, H, W] standardised, on device
        self.traj = list(trajectories)
        self.device = forcings_std.device
        self.diffusion = DiffusionSampler(net)
        kw = dict(num_steps=1, sigma_min=0.02, sigma_max=200.0, auxiliary=0.6)      # generate.py:255-260
        kw.update(solver_kwargs or {})
        self.solver_kwargs = kw
        if solver == "scm":
            self._solve = self.diffusion.scm_solver
        elif solver == "2s":
            self._solve = self.diffusion.dpm_solver_2s
        else:
            raise ValueError(f"Unknown solver mode: {solver}")
        self.generators = [torch.Generator(device=self.device).manual_seed(trajectory_seed(m, j))
                           for m, j in self.traj]
        inner = getattr(net, "module", net)
        self.n_var = inner.img_channels
        self.res = tuple(int(v) for v in inner.img_resolution)
        B = len(self.traj)
        n_forc = forcings_std.shape[1]
        self.cond = torch.empty(B, self.n_var + n_forc, *self.res, device=self.device)
        self.latents = torch.empty(B, self.n_var, *self.res, device=self.device)

    def set_state(self, x_std: torch.Tensor) -> None:
        """x_std: [B, n_var, H, W] standardised initial conditions, one row per trajectory."""
        self.cond[:, : self.n_var].copy_(x_std)

    def draw_latents(self) -> torch.Tensor:
        for b, g in enumerate(self.generators):
            self.latents[b].normal_(generator=g)          # == torch.randn(..., generator=g) (factory.py:52-56)
        return self.latents

    @torch.no_grad()
    def step(self, i: int, forcings_i: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One 6 h advance of every trajectory; returns the physical state [B, n_var, H, W]."""
        f = self.forcings[i] if forcings_i is None else forcings_i
        self.cond[:, self.n_var:].copy_(f.unsqueeze(0).expand(self.cond.shape[0], -1, -1, -1) if f.dim() == 3 else f)
        y = self._solve(latents=self.draw_latents(), condition=self.cond, **self.solver_kwargs)
        x_std = self.cond[:, : self.n_var]
        x_phys = torch.addcmul(x_std * self.norm.x_std + self.norm.x_mean, y, self.norm.diff_std)
        x_std.copy_((x_phys - self.norm.x_mean) / self.norm.x_std)
        return x_phys

    @torch.no_grad()
    def run(self, steps: int, on_step: Optional[Callable[[int, torch.Tensor], None]] = None) -> torch.Tensor:
        x_phys = None
        for i in range(steps):
            x_phys = self.step(i)
            if on_step is not None:
                on_step(i, x_phys)
        return x_phys
