"""``rollout_and_save``: the ensemble rollout written to the reference's output layout (stockeh/swift
``src/swift/generate.py:48-152``), on top of ``EnsembleRollout.run_to_host`` and ``ForecastStore``.

What the reference does per (member, IC batch): allocate ``rollout[bs, steps+1, C, H, W]`` on the host, store the
unstandardised initial state at lead 0 (``generate.py:95-96``), run the step loop with a blocking ``.cpu()`` per step
(``:129``), then assign the block to ``store[var][n:n+bs, m]`` variable by variable (``:139-152``).  Here every trajectory
of this rank advances together, the device->host copy of step i overlaps the compute of step i+1, and the store is filled
by writer threads while the GPU keeps stepping (file writes release the GIL).

Not here (out of scope): the ERA5 HDF5 reader, checkpoint discovery, hydra config loading.  The
command line below drives the same code with synthetic initial conditions and random-init Swift-B weights:

    python -m swift_b200.generate --output /tmp/fc.zarr --members 4 --ics 2 --steps 8 [--dump zarr|zarr-step|numpy]
    torchrun --nproc-per-node 8 -m swift_b200.generate ...        # (member, IC) shards, one store, no collective
"""
from __future__ import annotations

import argparse
import os
import time
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Sequence

import numpy as np
import torch

from .rollout import EnsembleRollout, Normalizers, shard_trajectories
from .store import ForecastStore


def era5_variables(n_levels: int = 13) -> list:
    """The 69 prognostic variables of configs/data/era5-flare-1.4.yaml:8-78 in state-tensor order."""
    levels = [50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000][:n_levels]
    names = ["2m_temperature", "10m_u_component_of_wind", "10m_v_component_of_wind", "mean_sea_level_pressure"]
    for v in ("geopotential", "u_component_of_wind", "v_component_of_wind", "temperature", "specific_humidity"):
        names += [f"{v}_{p}" for p in levels]
    return names


@torch.no_grad()
def rollout_and_save(ro: EnsembleRollout, store: ForecastStore, x0_std: torch.Tensor, steps: int,
                     forcings_host: Optional[torch.Tensor] = None, writers: int = 4) -> dict:
    """Advance every trajectory of ``ro`` by ``steps`` and write leads 0..steps of each into ``store`` at
    [ic, member] = ``ro.traj[b]``.  ``x0_std`` [B, n_var, H, W]: standardised initial conditions (host or device).
    Returns timing counters.

    Host memory: the "trajectory" layout keeps every lead time of a trajectory in one chunk, so all B x (steps+1) fields
    are held on the host until the rollout ends (the reference holds bs x (steps+1) per batch, generate.py:95); for long
    rollouts of many resident trajectories (96 x 61 Swift-B states = 53 GB) use the "step" or "numpy" layout, which
    stream every lead time to disk while the GPU keeps stepping."""
    if steps != store.steps:
        raise ValueError(f"store was created for {store.steps} steps, rollout asked for {steps}")
    B = len(ro.traj)
    if x0_std.shape[0] != B:
        raise ValueError(f"x0_std has {x0_std.shape[0]} rows, the rollout owns {B} trajectories")
    if store.layout == "trajectory" and store.batch != 1:
        raise ValueError("rollout_and_save writes one trajectory per chunk: create the store with batch=1")
    dev = ro.device
    x0_dev = x0_std.to(dev, non_blocking=True)
    ro.set_state(x0_dev)
    lead0 = (x0_dev * ro.norm.x_std + ro.norm.x_mean).cpu().numpy()          # unstandardize_x (generate.py:96)
    whole = store.layout == "trajectory"
    block = np.empty((B, steps + 1) + tuple(lead0.shape[1:]), dtype=np.float32) if whole else None
    pool = ThreadPoolExecutor(max_workers=max(1, writers))
    pending: deque = deque()

    def submit(fn, *a):
        pending.append(pool.submit(fn, *a))
        while len(pending) > 8 * max(1, writers):                             # bound the host copies in flight
            pending.popleft().result()

    def put_lead(k: int, fields: np.ndarray) -> None:
        if whole:
            block[:, k] = fields
        else:
            snap = fields.copy() if k > 0 else fields                          # the pinned slot is reused two steps later
            for b, (m, j) in enumerate(ro.traj):
                submit(store.write_step, j, m, k, snap[b])

    t0 = time.perf_counter()
    put_lead(0, lead0)
    out_host = torch.empty((2,) + tuple(ro.phys.shape), dtype=torch.float32).pin_memory()
    ro.run_to_host(steps, out_host, forcings_host, on_host=lambda i, view: put_lead(i + 1, view.numpy()))
    t_roll = time.perf_counter() - t0
    if whole:
        for b, (m, j) in enumerate(ro.traj):
            submit(store.write_trajectories, j, m, block[b:b + 1])
    while pending:
        pending.popleft().result()
    pool.shutdown()
    store.flush()
    t_all = time.perf_counter() - t0
    return {"trajectories": B, "steps": steps, "rollout_s": t_roll, "total_s": t_all,
            "bytes_written": int(B * (steps + 1) * lead0[0].size * 4)}


def main(argv: Optional[Sequence[str]] = None) -> None:
    ap = argparse.ArgumentParser(description="Swift-B ensemble rollout on synthetic initial conditions (generate.py's loop)")
    ap.add_argument("--output", required=True)
    ap.add_argument("--members", type=int, default=1)           # generate.py:29
    ap.add_argument("--steps", type=int, default=8)             # generate.py:30
    ap.add_argument("--ics", type=int, default=1, help="number of initial conditions (the reference: --samples)")
    ap.add_argument("--interval", type=int, default=6, choices=[6, 12, 24])
    ap.add_argument("--dump", default="zarr", choices=["zarr", "zarr-step", "numpy"])
    ap.add_argument("--solver", default="scm", choices=["scm", "2s"])
    ap.add_argument("--config", default="swift_b", choices=["swift_b", "tiny"])
    args = ap.parse_args(argv)

    import torch.distributed as dist

    from . import synthetic as syn
    from .precond import PassPrecond

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("swift_b200.generate needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = syn.SWIFT_B if args.config == "swift_b" else syn.SWIFT_TINY
    n_var = cfg["out_channels"]
    n_forc = cfg["in_channels"] - 2 * n_var
    model_cfg = dict(_target_="swift_b200.swinv2.SwinV2", window_size=cfg["window_size"], shift_size=cfg["shift_size"],
                     patch_size=cfg["patch_size"], depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"])
    net = PassPrecond(model_cfg, img_resolution=cfg["img_resolution"], img_channels=n_var,
                      condition_channels=cfg["in_channels"] - n_var, auxiliary_dim=1)
    net.load_state_dict(syn.random_state_dict(cfg, seed=1, prefix="model."), strict=True)
    net = net.to(dev).eval()
    net.model.max_chunk = 24                                                  # trajectories per kernel launch sequence (bench.py)
    H, W = cfg["img_resolution"]
    variables = era5_variables() if n_var == 69 else [f"field{c}" for c in range(n_var)]
    layout = {"zarr": "trajectory", "zarr-step": "step", "numpy": "numpy"}[args.dump]
    if rank == 0:                                                             # run_on_rank0 (generate.py:272-282)
        ForecastStore.create(args.output, variables, args.ics, args.members, args.steps,
                             lat=np.linspace(-89.296875, 89.296875, H), lon=np.arange(W) * (360.0 / W),
                             interval_hours=args.interval, layout=layout)
    if world > 1:
        dist.barrier()
    store = ForecastStore.open(args.output) if layout != "numpy" else ForecastStore(
        args.output, "numpy", variables, args.ics, args.members, args.steps, (H, W))
    traj = shard_trajectories(args.members, args.ics, rank, world)
    x0 = torch.stack([syn.synthetic_fields(cfg, 1, seed=j)[1][0, :n_var] for _, j in traj]) if traj else None
    stride = args.interval // 6
    ic_times = {j: j for j in range(args.ics)}                                # IC j = the analysis at 6 h file index j
    forc = syn.synthetic_forcings(cfg, args.ics + args.steps * stride + 1, seed=0, n_forcings=n_forc)
    info = {"trajectories": 0}
    if traj:
        skw = dict(num_steps=20, sigma_min=0.02, sigma_max=200.0, auxiliary=0.6) if args.solver == "2s" else None
        ro = EnsembleRollout(net, Normalizers.synthetic(n_var, dev, diff=0.1), forc.to(dev), traj, solver=args.solver,
                             solver_kwargs=skw, ic_times=ic_times, interval=args.interval)
        info = rollout_and_save(ro, store, x0, args.steps, forc.pin_memory())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    print(f"[rank {rank}] {info}")


if __name__ == "__main__":
    main()
