"""Autoregressive ensemble rollout: the loop ``generate.rollout_and_save`` runs around the sampler
(stockeh/swift ``src/swift/generate.py:79-136``), restated for resident device state.

Per 6 h step and trajectory (member m, initial condition j) the reference does (residual=True branch):

    X      <- cat([X, standardize_x(forcings(step))], 1)          generate.py:100-117
    Y       = sampler(X, generator)                               :118   (fresh N(0,1) latents, one denoiser call)
    X_phys  = unstandardize_x(X)[:, :n_var] + unstandardize_t(Y)  :120-126  (t_means = 0: era5.py:95-100)
    rollout[:, i+1] = X_phys.cpu()                                :129
    X      <- standardize_x(X_phys)                               :131

Here the whole step is six kinds of kernels and no PyTorch op: counter-based latents, forcings copy into the condition
buffer, the denoiser (patch gather fuses the concat, the head epilogue fuses the sCM update AND the affine glue above,
updating the state in place), and a device-side step counter -- so one captured CUDA graph is replayed for every step.

Differences from the reference, all documented in DESIGN.md:
  * trajectories are sharded by flattened (IC, member) index over ranks (the reference shards by member only and
    leaves 4 of 8 ranks half idle for 12 members); no collective is needed on the forecast path;
  * every trajectory owns its noise stream: Philox4x32-10 keyed by ``trajectory_seed(member, ic)`` with counter
    (element, step), so a trajectory's result does not depend on world size or batching (the reference's per-member
    ``torch.Generator`` is consumed in batch order);
  * forcings are pre-staged on the device as a TIME-indexed [n_times, n_forcings, H, W] table (row = 6 h file index)
    instead of one HDF5 read per sample per step on the main thread; trajectory b reads row
    ``ic_times[ic_b] + step * (interval // 6)`` -- the reference's ``get_forcings(j + int(i * interval // 6)) for j in
    idx`` (generate.py:105-110), where j is the dataset (time) index of the sample's initial condition.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import _lib
from .engine import RolloutGlue
from .sampler import DiffusionSampler, _fused_target


def shard_trajectories(members: int, n_ic: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """(member, ic) pairs owned by ``rank``: contiguous blocks of the IC-major flattened index j*members + m, so the
    members of one initial condition stay on one GPU whenever n_ic is a multiple of world."""
    total = members * n_ic
    per, rem = divmod(total, world)
    lo = rank * per + min(rank, rem)
    hi = lo + per + (1 if rank < rem else 0)
    return [(idx % members, idx // members) for idx in range(lo, hi)]


def trajectory_seed(member: int, ic: int) -> int:
    """Key of the (member, ic) noise stream (member m of every IC extends the reference's ``manual_seed(m)`` family)."""
    return member + 1_000_003 * ic


class ReferenceNoise:
    """The reference's latent stream, replayed with random access (validation mode).

    generate.py:79-118 seeds ONE ``torch.Generator(device).manual_seed(m)`` per member and consumes it in loop order:
    IC batches of ``batch`` samples (DataLoader order, the last one may be ragged), and inside a batch the lead times,
    each ``sampler(X, generator)`` call drawing ``torch.randn((bs, C, H, W), generator=..., device=...)``
    (generating/factory.py:46-61).  A trajectory's noise therefore depends on the batch size and on every batch before
    it.  All trajectories advance together here, so the call that the reference would make for (member, batch b, lead
    i) is reproduced by positioning a CUDA generator at that call's Philox offset: a ``randn`` of a given shape advances
    the offset by a fixed amount (measured once per shape), hence offset(b, i) = b * steps * inc(batch) + i * inc(bs_b).
    Valid for samplers that draw only the latents (1-step sCM, TrigFlow 2S with S_churn = 0)."""

    def __init__(self, trajectories: Sequence[Tuple[int, int]], n_ic: int, batch: int, steps: int,
                 sample_shape: Tuple[int, ...], device):
        if batch < 1 or steps < 1:
            raise ValueError("ReferenceNoise needs batch >= 1 and the number of lead times the reference was run with")
        self.device, self.steps, self.batch, self.shape = device, int(steps), int(batch), tuple(sample_shape)
        self.groups = {}                                 # (member, batch index) -> [(row here, row in the batch)]
        for row, (m, j) in enumerate(trajectories):
            if not 0 <= j < n_ic:
                raise ValueError(f"trajectory ({m}, {j}) outside the {n_ic} initial conditions of the run")
            self.groups.setdefault((m, j // self.batch), []).append((row, j % self.batch))
        self.sizes = {b: min(self.batch, n_ic - b * self.batch) for _, b in self.groups}
        self._inc = {}
        self._gen = torch.Generator(device=device)

    def _increment(self, bs: int) -> int:
        if bs not in self._inc:
            self._gen.manual_seed(0)
            o0 = self._gen.get_offset()
            torch.randn((bs,) + self.shape, generator=self._gen, device=self.device)
            self._inc[bs] = self._gen.get_offset() - o0
        return self._inc[bs]

    def fill(self, latents: torch.Tensor, step: int) -> None:
        """latents[row] <- what the reference draws for that trajectory at lead ``step``."""
        if not 0 <= step < self.steps:
            raise ValueError(f"lead {step} outside the {self.steps}-step stream being replayed")
        for (m, b), rows in self.groups.items():
            bs = self.sizes[b]
            offset = b * self.steps * self._increment(self.batch) + step * self._increment(bs)   # (measuring reseeds)
            self._gen.manual_seed(m)                                         # generate.py:83
            self._gen.set_offset(offset)
            z = torch.randn((bs,) + self.shape, generator=self._gen, device=self.device)
            here = torch.tensor([r for r, _ in rows], device=self.device)
            latents.index_copy_(0, here, z[[k for _, k in rows]])


@dataclass
class Normalizers:
    """Per-channel affine maps of data/era5.py:80-108 as [1, C, 1, 1] device tensors."""
    x_mean: torch.Tensor      # variables only
    x_std: torch.Tensor
    diff_std: torch.Tensor    # t_stds[interval] (t_means are zero for residual targets)
    zero_channel: int = -1    # index of sea_surface_temperature when delta != 24 (era5.py zero_field), else -1

    @staticmethod
    def synthetic(n_var: int, device, diff: float = 0.1) -> "Normalizers":
        one = torch.ones(1, n_var, 1, 1, device=device)
        return Normalizers(torch.zeros_like(one), one, diff * one)


class EnsembleRollout:
    """Advances a batch of independent trajectories that live on one GPU."""

    def __init__(self, net, norm: Normalizers, forcings_std: torch.Tensor, trajectories: Sequence[Tuple[int, int]],
                 solver: str = "scm", solver_kwargs: Optional[dict] = None, use_graph: bool = True,
                 noise: Optional["ReferenceNoise"] = None, residual: bool = True,
                 ic_times: Union[None, Sequence[int], Dict[int, int]] = None, interval: int = 6):
        """``forcings_std`` [n_times, n_forc, H, W]: standardised forcings indexed by time (one row per 6 h file).
        ``ic_times[j]``: the row that holds the forcings valid at the initial time of IC j (the reference's dataset index
        of that sample, generate.py:108) -- a sequence indexed by IC or a dict; ``interval``: hours per step (6, 12, 24:
        generate.py:31), so trajectory (m, j) reads row ``ic_times[j] + step * interval // 6`` at lead ``step``.
        ICs have different valid times in the reference, so ``ic_times`` is mandatory as soon as the trajectories span
        more than one IC (pass zeros for ICs that share one valid time)."""
        self.net = net
        self.residual = bool(residual)          # data/defaults.yaml:7; False = the network predicts the next state itself
        self.norm = norm
        self.forcings = forcings_std.contiguous()   # [n_times, n_forc, H, W] standardised, on device, indexed by time
        self.traj = list(trajectories)
        self.device = forcings_std.device
        if self.forcings.dim() != 4:
            raise ValueError(f"forcings_std must be [n_times, n_forc, H, W], got {tuple(self.forcings.shape)}")
        if interval % 6 != 0 or interval <= 0:
            raise ValueError(f"interval must be a positive multiple of 6 hours, got {interval}")
        self.stride = int(interval) // 6                                           # generate.py:109
        ics = sorted({j for _, j in self.traj})
        if ic_times is None:
            if len(ics) > 1:
                raise ValueError(f"the trajectories span {len(ics)} initial conditions: pass ic_times (the forcings-table row "
                                 "of every IC's initial time; the reference reads get_forcings(j + i*interval//6) per IC j, "
                                 "generate.py:105-110)")
            ic_times = {j: 0 for j in ics}
        try:
            base = [int(ic_times[j]) for _, j in self.traj]
        except (KeyError, IndexError) as e:
            raise ValueError(f"ic_times has no entry for an initial condition of this rank: {e}") from None
        if base and (min(base) < 0 or max(base) >= self.forcings.shape[0]):
            raise ValueError(f"ic_times rows {min(base)}..{max(base)} outside the {self.forcings.shape[0]}-row forcings table")
        self._base_host = base
        self.base = torch.tensor(base, dtype=torch.int32, device=self.device)
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
        inner = getattr(net, "module", net)
        self.n_var = inner.img_channels
        self.res = tuple(int(v) for v in inner.img_resolution)
        self.sigma_data = float(inner.sigma_data)
        B = len(self.traj)
        n_forc = self.forcings.shape[1]
        self.cond = torch.zeros(B, self.n_var + n_forc, *self.res, device=self.device)
        self.latents = torch.empty(B, self.n_var, *self.res, device=self.device)
        self.phys = torch.empty(B, self.n_var, *self.res, device=self.device)
        self.seeds = torch.tensor([trajectory_seed(m, j) for m, j in self.traj], dtype=torch.int64, device=self.device)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.noise = noise                      # None: the per-trajectory Philox streams (default); else the reference's
        self._step_host = 0                     # host mirror of step_dev (the replayed stream is positioned on the host)
        self.lib = _lib.lib()
        self.model = _fused_target(net)
        # the single-kernel-sequence step: 1-step sCM through the fused CUDA entry point
        self.fused = self.model is not None and solver == "scm" and kw["num_steps"] == 1 and self.residual
        self.use_graph = use_graph and self.fused
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._glue = RolloutGlue(self.cond, norm.x_std, norm.x_mean, norm.diff_std, self.phys, norm.zero_channel)
        self._cond_vecs = None
        self._engine_gen = -1                   # Engine.generation the graph / conditioning vectors were made for
        self._stats = None
        self._truth: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ ensemble statistics (eval/metrics.py on device)
    def attach_statistics(self, stats, truth: torch.Tensor) -> None:
        """Score every step on the device: ``stats`` (swift_b200.ensemble.EnsembleStatistics) receives the physical state of
        each step, compared with ``truth`` [n_ic, n_var, H, W] -- a static device buffer the caller refreshes with the
        verifying analysis of the coming step before calling step().  Needs whole ensembles on this GPU in IC-major order
        (what shard_trajectories produces when the ICs divide evenly over the ranks)."""
        n = stats.members
        ok = len(self.traj) == stats.n_ic * n and all(m == k % n and j == self.traj[0][1] + k // n
                                                       for k, (m, j) in enumerate(self.traj))
        if not ok:
            raise ValueError("attach_statistics needs the trajectories of this rank to be whole ensembles in IC-major order "
                             f"({stats.n_ic} ICs x {n} members); got {len(self.traj)} trajectories starting at {self.traj[:2]}")
        self._stats, self._truth = stats, truth
        self._graph = None                       # the statistics kernel becomes part of the captured step

    # ------------------------------------------------------------------ state
    def set_state(self, x_std: torch.Tensor, step: int = 0) -> None:
        """x_std: [B, n_var, H, W] standardised initial conditions, one row per trajectory."""
        if step < 0:
            raise ValueError(f"step must be >= 0, got {step}")
        self.cond[:, : self.n_var].copy_(x_std)
        self.step_dev.fill_(step)
        self._step_host = int(step)

    def forcing_rows(self, step: Optional[int] = None) -> List[int]:
        """Distinct rows of the forcings table read at lead ``step`` (default: the coming step)."""
        s = self._step_host if step is None else int(step)
        return sorted({b + s * self.stride for b in self._base_host})

    def _check_bounds(self, table_needed: bool) -> None:
        """Raise BEFORE launching when the coming step would leave the forcings table or the statistics buffer (the
        kernels read / write rows addressed by the device step counter)."""
        if table_needed and self.forcings.shape[1] > 0:
            rows = self.forcing_rows()
            if rows and rows[-1] >= self.forcings.shape[0]:
                raise RuntimeError(f"step {self._step_host} needs forcings row {rows[-1]} but the table has "
                                   f"{self.forcings.shape[0]} rows (ic_times up to {max(self._base_host)}, stride "
                                   f"{self.stride})")
        if self._stats is not None and self._step_host >= self._stats.steps:
            raise RuntimeError(f"step {self._step_host} exceeds the {self._stats.steps} steps the attached "
                               "EnsembleStatistics was sized for")

    # ------------------------------------------------------------------ pieces of one step
    def _stream(self) -> int:
        return torch.cuda.current_stream().cuda_stream

    def draw_latents(self) -> torch.Tensor:
        if self.noise is not None:
            self.noise.fill(self.latents, self._step_host)
            return self.latents
        n = self.latents[0].numel()
        _lib.check(self.lib.swb200_rollout_noise(self.latents.data_ptr(), self.seeds.data_ptr(),
                                                 self.step_dev.data_ptr(), len(self.traj), n, self._stream()), "noise")
        return self.latents

    def _load_forcings(self) -> None:
        hw = self.res[0] * self.res[1]
        _lib.check(self.lib.swb200_rollout_forcings(self.cond.data_ptr(), self.cond.shape[1], self.n_var,
                                                    self.forcings.data_ptr(), self.forcings.shape[1],
                                                    self.forcings.shape[0], self.base.data_ptr(), self.stride,
                                                    self.step_dev.data_ptr(), len(self.traj), hw, self._stream()),
                   "forcings")

    def _advance(self) -> None:
        _lib.check(self.lib.swb200_rollout_advance(self.step_dev.data_ptr(), self._stream()), "advance")

    def _fused_step(self) -> None:
        """noise -> forcings -> denoiser (+ sCM update + affine glue in the head epilogue) -> step += 1."""
        eng = self.model.engine()
        t = torch.tensor([torch.pi / 2])                     # diffusion.py:435-436 (fp32 pi/2)
        cos_t, sin_t = float(torch.cos(t)), float(torch.sin(t))
        if self.noise is None:                               # replayed reference noise is drawn outside (not capturable)
            self.draw_latents()
        self._load_forcings()
        gain, bias = self._cond_vecs
        sd = self.sigma_data
        x_in = self.latents
        if sd != 1.0:
            x_in.mul_(sd)                                    # x_t = latents * sigma_d (diffusion.py:452)
        eng.forward(x_in, self.cond, gain, bias, scale0=1.0 / sd, xt=x_in, alpha=cos_t, beta=-sin_t * sd,
                    rollout=self._glue)
        n_extra = 3 if self.noise is None else 2              # own kernels besides the forward: (noise,) forcings, advance
        if self._stats is not None:
            self._stats.accumulate(self.phys, self._truth, step_dev=self.step_dev)
            n_extra += 1
        self._advance()
        eng.launches += n_extra

    @torch.no_grad()
    def step(self, forcings_i: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One 6 h advance of every trajectory; returns the physical state [B, n_var, H, W] (a view of a static
        buffer that the next step overwrites)."""
        self._check_bounds(table_needed=forcings_i is None)
        if self.fused and forcings_i is None:
            gen = self.model.engine().generation
            if gen != self._engine_gen:          # weights re-packed / knobs changed: the captured graph and the cached
                self._graph, self._cond_vecs = None, None      # conditioning vectors point into the old engine's buffers
                self._engine_gen = gen
            if self._cond_vecs is None:
                self._cond_vecs = self.diffusion._conditioning(self.model, float(torch.tensor([torch.pi / 2])),
                                                               self.solver_kwargs["auxiliary"], len(self.traj),
                                                               self.device)
            if self.noise is not None:
                self.draw_latents()
            self._step_host += 1
            if not self.use_graph:
                self._fused_step()
            else:
                if self._graph is None:
                    self.model.engine().workspace(len(self.traj))        # allocate before capture
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._fused_step()
                    self._graph = g          # capture does not execute: the replay below is the first real step
                self._graph.replay()
                self.model.engine().launches += (self.model.engine().launches_per_forward(len(self.traj)) +
                                                 (3 if self.noise is None else 2) +
                                                 (1 if self._stats is not None else 0))
            return self.phys
        # generic path (multi-step sCM, 2S, non-fused nets, externally supplied forcings)
        if forcings_i is None:
            self._load_forcings()
        else:
            f = forcings_i
            self.cond[:, self.n_var:].copy_(f.unsqueeze(0).expand(self.cond.shape[0], -1, -1, -1) if f.dim() == 3 else f)
        kw = {k: v for k, v in self.solver_kwargs.items()}
        y = self._solve(latents=self.draw_latents(), condition=self.cond, **kw)
        self._step_host += 1
        x_std = self.cond[:, : self.n_var]
        if self.residual:                                                        # generate.py:120-131
            x_phys = torch.addcmul(x_std * self.norm.x_std + self.norm.x_mean, y, self.norm.diff_std)
            x_new = (x_phys - self.norm.x_mean) / self.norm.x_std
        else:                                                                    # generate.py:132-136
            x_phys = y * self.norm.x_std + self.norm.x_mean
            x_new = y.clone()
        if self.norm.zero_channel >= 0:
            x_phys[:, self.norm.zero_channel] = 0
            x_new[:, self.norm.zero_channel] = 0
        x_std.copy_(x_new)
        self.phys.copy_(x_phys)
        if self._stats is not None:
            self._stats.accumulate(self.phys, self._truth, step_dev=self.step_dev)
        self._advance()
        return self.phys

    @torch.no_grad()
    def run_to_host(self, steps: int, out_host: torch.Tensor, forcings_host: Optional[torch.Tensor] = None,
                    on_host: Optional[Callable[[int, torch.Tensor], None]] = None) -> None:
        """The loop of generate.py:97-136 with HOST buffers: per step the standardised forcings every IC needs at that
        lead (rows ``ic_times[j] + step * interval // 6`` of ``forcings_host``, a pinned time-indexed table with the
        device table's shape) are copied host -> device -- what the reference fetches with ``get_forcings`` per sample,
        de-duplicated over the members of an IC -- and the new physical state of every trajectory
        lands in ``out_host`` [steps or 2, B, n_var, H, W] (pinned; with 2 slots they are used alternately).
        The reference blocks on ``.cpu()`` every step (generate.py:129); here the device->host copy of step i runs on a
        copy stream while step i+1 computes (the state is first parked in a device staging buffer, because the next
        step overwrites ``phys``).  ``on_host(i, view)`` is called once step i's data is complete in host memory."""
        if out_host.dim() != 5 or tuple(out_host.shape[1:]) != tuple(self.phys.shape) or not out_host.is_pinned():
            raise RuntimeError(f"out_host must be a pinned [slots, {', '.join(map(str, self.phys.shape))}] tensor")
        if forcings_host is not None and tuple(forcings_host.shape) != tuple(self.forcings.shape):
            raise RuntimeError(f"forcings_host must mirror the device table {tuple(self.forcings.shape)}, got "
                               f"{tuple(forcings_host.shape)}")
        slots = out_host.shape[0]
        if slots < min(2, steps):
            raise RuntimeError("out_host needs at least 2 slots (one is written while the other is consumed)")
        main = torch.cuda.current_stream()
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staging = torch.empty_like(self.phys)
        staged = torch.cuda.Event()
        drained = [torch.cuda.Event() for _ in range(2)]
        for i in range(steps):
            if forcings_host is not None:
                self._check_bounds(table_needed=True)
                for r in self.forcing_rows():
                    self.forcings[r].copy_(forcings_host[r], non_blocking=True)
            x_phys = self.step()
            if i > 0:
                main.wait_event(drained[(i - 1) & 1])          # the copy stream has finished reading the staging buffer
            self._staging.copy_(x_phys, non_blocking=True)
            staged.record(main)
            self._copy_stream.wait_event(staged)
            with torch.cuda.stream(self._copy_stream):
                out_host[i % slots].copy_(self._staging, non_blocking=True)
                drained[i & 1].record(self._copy_stream)
            if on_host is not None and i > 0:
                drained[(i - 1) & 1].synchronize()
                on_host(i - 1, out_host[(i - 1) % slots])
        self._copy_stream.synchronize()
        if on_host is not None and steps > 0:
            on_host(steps - 1, out_host[(steps - 1) % slots])

    @torch.no_grad()
    def run(self, steps: int, on_step: Optional[Callable[[int, torch.Tensor], None]] = None) -> torch.Tensor:
        for i in range(steps):
            x_phys = self.step()
            if on_step is not None:
                on_step(i, x_phys)
        return self.phys
