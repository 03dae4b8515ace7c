#!/usr/bin/env python
"""Benchmark of the Swift forecast hot path (BASELINE.json: forecast member-steps/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference [...]                         the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]; configs[2] when launched on 8 GPUs): Swift-B (era5-swinv2-1.4-scm) sCM 1-step
autoregressive rollout, 12 members x 8 initial conditions = 96 resident trajectories PER GPU (weak scaling: at
N = 8 this is the 12 x 64 workload of configs[2]), synthetic 128x256 ERA5-shaped fields, random-init weights.
One bench "step" = one 6 h advance of every resident trajectory = 96 member-steps per GPU (96 denoiser forwards).
K = 60 steps is the full 15-day rollout.

  value  member-steps/s with initial conditions, forcings and weights already resident in HBM (device-timed with
         CUDA events, max over ranks).
  e2e    the same rollout through the public sampler API with HOST buffers: every step copies that step's forcings
         from pinned host memory and reads the new physical state back to pinned host memory (what
         generate.py:100-131 does with `.to(device)` / `.cpu()`).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_MEMBER_STEP = 2.75177e12     # un-padded 2*M*N*K of one Swift-B denoiser call
MEMBERS, ICS_PER_GPU = 12, 8


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            p = [v.strip() for v in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1]))
                smax = float(p[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_member_steps_per_sec(n_timed: int, warmup: int = 1):
    """The reference algorithm (oracle port: plain PyTorch fp32, all host threads) on one Swift-B sCM step, B = 1."""
    import torch
    from oracle import swinv2_oracle as orc
    from swift_b200 import synthetic as syn

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = syn.SWIFT_B
    sd = syn.random_state_dict(cfg, seed=1)
    ocfg = orc.make_cfg(**cfg)
    lat, cond = syn.synthetic_fields(cfg, 1, seed=0)
    net = lambda x, t, c, a: orc.pass_precond(sd, ocfg, x, t, c, a)
    times = []
    with torch.no_grad():
        for i in range(warmup + n_timed):
            t0 = time.perf_counter()
            orc.scm_solver(net, lat, cond, 0.6, num_steps=1)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    per = sum(times) / len(times)
    return 1.0 / per, per, cores, torch.get_num_threads()


def gpu_eager_best(dev, batches=(2, 8)):
    """The eager-PyTorch baseline at several batch sizes (the product runs 24 trajectories per launch sequence, so the
    baseline gets its best batch too); returns the per-batch results and the best fp32 / bf16 rates."""
    runs = []
    for b in batches:
        try:
            runs.append(gpu_eager_member_steps_per_sec(dev, batch=b))
        except Exception as e:
            runs.append({"batch": b, "unavailable": f"{type(e).__name__}: {e}"[:200]})
    best = {"unit": "member-steps/s", "what": runs[0].get("what") if runs else None, "per_batch": runs}
    for key in ("fp32", "bf16_autocast"):
        vals = [(r[key], r["batch"]) for r in runs if r.get(key)]
        if vals:
            best[key], best[key + "_batch"] = max(vals)
    acc = [r["bf16_autocast_per_field_rel_l2_max"] for r in runs if "bf16_autocast_per_field_rel_l2_max" in r]
    if acc:
        best["bf16_autocast_per_field_rel_l2_max"] = max(acc)
    return best


def gpu_eager_member_steps_per_sec(dev, batch: int = 2, n_timed: int = 3):
    """SURVEY.md section 8d's "fair GPU baseline": the reference ALGORITHM as plain PyTorch ops (oracle port with the
    reference's inference attention branch, F.scaled_dot_product_attention) run eagerly on the same B200, in fp32 with
    TF32 off and under bf16 autocast.  A reported baseline like cpu_baseline, never part of the product path."""
    import torch
    from oracle import swinv2_oracle as orc
    from swift_b200 import synthetic as syn

    cfg = syn.SWIFT_B
    sd = {k: v.to(dev) for k, v in syn.random_state_dict(cfg, seed=1).items()}
    ocfg = dict(orc.make_cfg(**cfg), sdpa=True)
    lat, cond = (v.to(dev) for v in syn.synthetic_fields(cfg, batch, seed=0))
    net = lambda x, t, c, a: orc.pass_precond(sd, ocfg, x, t, c, a)
    out = {"batch": batch, "unit": "member-steps/s", "what": "oracle port of swift.models.swinv2 + scm_solver, eager "
           "PyTorch ops (cuBLAS / SDPA / ATen kernels) on this GPU"}
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    ys = {}
    try:
        for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)),
                          ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            try:
                with torch.no_grad(), ctx:
                    orc.scm_solver(net, lat, cond, 0.6, num_steps=1)
                    torch.cuda.synchronize(dev)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(n_timed):
                        ys[name] = orc.scm_solver(net, lat, cond, 0.6, num_steps=1)
                    e1.record()
                    torch.cuda.synchronize(dev)
                out[name] = batch * n_timed / (e0.elapsed_time(e1) / 1e3)
            except Exception as e:                               # a baseline leg must not cost the headline line
                out[name] = None
                out[name + "_error"] = f"{type(e).__name__}: {e}"[:200]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    if len(ys) == 2:      # accuracy of plain autocast next to its speed: per-field rel-L2 of one sCM step vs the fp32 run
        a, b = ys["bf16_autocast"].float(), ys["fp32"]
        out["bf16_autocast_per_field_rel_l2_max"] = float(((a - b).flatten(2).norm(dim=-1) /
                                                           b.flatten(2).norm(dim=-1)).max())
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = max(1, min(args.steps, 20))          # bounded: one member-step is seconds of CPU work
    v, per, cores, threads = cpu_member_steps_per_sec(n, max(1, min(args.warmup, 1)))
    sample = (f"{n} sCM member-steps of Swift-B at batch 1 (of the 96/step workload), fp32, "
              f"oracle port of swift.models.swinv2 + scm_solver, {threads} torch threads")
    line = {
        "impl": "reference", "metric": "forecast member-steps/sec", "value": v, "unit": "member-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, "cpu"),
        "cpu_baseline": {"value": v, "unit": "member-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "member-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_tflops": v * FLOP_PER_MEMBER_STEP / 1e12,
    }
    print(json.dumps(line))


def workload_config(n_gpus: int, where: str, solver: str = "scm"):
    name = ("Swift-B (era5-swinv2-1.4-scm) sCM 1-step" if solver == "scm" else
            "Swift-B (era5-swinv2-1.4-trigflow) TrigFlow 2S 20-step (39 denoiser calls per member-step)")
    return {"workload": name + " autoregressive rollout, 12 members x 8 ICs per GPU "
                        "x K 6h steps (K=60: 15 days), 128x256 synthetic ERA5 fields, random-init weights",
            "members": MEMBERS, "initial_conditions": ICS_PER_GPU * n_gpus, "trajectories_per_gpu": MEMBERS * ICS_PER_GPU,
            "member_steps_per_bench_step": MEMBERS * ICS_PER_GPU * n_gpus,
            "sharding": f"(member, IC) over {n_gpus} GPU(s), no collective on the forecast path",
            "l2_policy": "inputs larger than L2 (96 trajectories x 18.5 MB inputs, 216 MB workspace per sample)",
            "step": "one CUDA graph per 6h step: Philox latents, forcings, the denoiser kernels per trajectory chunk (LayerNorm + residual update fused into the wo / w2 GEMM epilogues, sCM update and std/unstd glue fused in the head epilogue), ensemble statistics",
            "device": where}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from swift_b200 import synthetic as syn
    from swift_b200.precond import PassPrecond
    from swift_b200.rollout import EnsembleRollout, Normalizers, shard_trajectories

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 prints ONE JSON line on stdout: anything native code writes to fd 1 meanwhile (NCCL's version banner,
        # NCCL_DEBUG output) is sent to stderr; the descriptor is restored just before the line is printed
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n_gpus = world

    cfg = syn.SWIFT_B
    model_cfg = dict(_target_="swift_b200.swinv2.SwinV2", window_size=cfg["window_size"], shift_size=cfg["shift_size"],
                     patch_size=cfg["patch_size"], depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"])
    net = PassPrecond(model_cfg, img_resolution=cfg["img_resolution"], img_channels=syn.IMG_CHANNELS,
                      condition_channels=syn.COND_CHANNELS, auxiliary_dim=1, sigma_min=0.0, sigma_max=float("inf"))
    net.load_state_dict(syn.random_state_dict(cfg, seed=1, prefix="model."), strict=True)
    net = net.to(dev).eval()
    net.model.max_chunk = args.chunk
    if args.fuse_ln >= 0:
        net.model.fuse_ln = args.fuse_ln
    if args.x_single >= 0:
        net.model.x_single = bool(args.x_single)
    if args.split_embed >= 0:
        net.model.split_embed = bool(args.split_embed)
    if args.act_bf16:
        net.model.act_fp16 = False
    eng = net.model.engine()

    n_ic = ICS_PER_GPU * n_gpus
    traj = shard_trajectories(MEMBERS, n_ic, rank, world)
    B = len(traj)
    total_steps = args.warmup + args.steps + 2
    # forcings: a TIME-indexed table; IC j is the analysis at 6 h file index j (consecutive dataset indices, what the
    # reference's DataLoader yields), so trajectory (m, j) reads row j + step at every step (generate.py:105-110)
    ic_times = {j: j for j in range(n_ic)}
    forc_host = syn.synthetic_forcings(cfg, n_ic + total_steps, seed=0).pin_memory()
    forc_dev = forc_host.to(dev)
    norm = Normalizers.synthetic(syn.IMG_CHANNELS, dev, diff=0.1)
    skw = dict(num_steps=20, sigma_min=0.02, sigma_max=200.0, auxiliary=0.6) if args.solver == "2s" else None
    ro = EnsembleRollout(net, norm, forc_dev, traj, solver=args.solver, solver_kwargs=skw, use_graph=not args.no_graph,
                         ic_times=ic_times, interval=6)
    stats = None
    if not args.no_stats and len(traj) % MEMBERS == 0:
        # eval/metrics.py on the device: per-step sufficient statistics inside the step's CUDA graph, one NCCL all_gather
        # of the sums after the rollout (synthetic verification fields: zeros)
        import numpy as np
        from swift_b200.ensemble import EnsembleStatistics
        H_, W_ = cfg["img_resolution"]
        stats = EnsembleStatistics(MEMBERS, len(traj) // MEMBERS, syn.IMG_CHANNELS, (H_, W_), np.linspace(-89.3, 89.3, H_),
                                   total_steps + args.steps + 4, dev)
        ro.attach_statistics(stats, torch.zeros(len(traj) // MEMBERS, syn.IMG_CHANNELS, H_, W_, device=dev))
    ics = {}
    x0 = torch.empty(B, syn.IMG_CHANNELS, *cfg["img_resolution"])
    for b, (m, j) in enumerate(traj):
        if j not in ics:
            ics[j] = syn.synthetic_fields(cfg, 1, seed=j)[1][0, :syn.IMG_CHANNELS]
        x0[b] = ics[j]
    x0 = x0.pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident leg (one captured CUDA graph per 6 h step, replayed)
    ro.set_state(x0.to(dev))
    for i in range(args.warmup):
        ro.step()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        ro.step()
    e1.record()
    barrier()
    clock_info = clocks.stop()
    ms_local = e0.elapsed_time(e1)
    ms = max_over_ranks(ms_local)
    # per-rank step time and SM clock: the N-GPU value is the MAX over ranks, so a single power-capped straggler sets it
    per_rank = None
    if world > 1:
        mine = torch.tensor([ms_local / args.steps, float(clock_info.get("sm_mhz") or 0.0)], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        t_r = [float(a[0]) for a in allr]
        per_rank = {"ms_per_step": t_r, "sm_mhz": [float(a[1]) for a in allr], "min": min(t_r), "median": statistics.median(t_r),
                    "max": max(t_r)}
    launches = eng.launches - launches0
    member_steps = B * world * args.steps
    value = member_steps / (ms / 1e3)

    # ---------------- end-to-end leg: host buffers in, host buffers out, every step
    # (EnsembleRollout.run_to_host: per step the forcings come from pinned host memory and the new physical state of every
    #  trajectory lands in pinned host memory; the D2H copy of step i overlaps the compute of step i+1.  The timed region
    #  ends when the last step's state is complete on the host: wall clock, max over ranks.)
    out_host = torch.empty(2, B, syn.IMG_CHANNELS, *cfg["img_resolution"]).pin_memory()
    e2e_steps = args.steps if args.e2e_steps <= 0 else min(args.e2e_steps, args.steps)
    checksum = [0.0]

    def consume(i, view):                                                          # the host reads every step's result
        checksum[0] += float(view[0, 0, 0, 0]) + float(view[-1, -1, -1, -1])

    ro.set_state(x0.to(dev, non_blocking=True))                                    # H2D: initial conditions
    ro.run_to_host(1, out_host, forc_host, on_host=consume)                        # one untimed step
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e2.record()
    ro.run_to_host(e2e_steps, out_host, forc_host, on_host=consume)
    e3.record()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(max(e2.elapsed_time(e3), wall_ms))
    e2e_value = B * world * e2e_steps / (e2e_ms / 1e3)
    h2d = len(ro.forcing_rows(1)) * forc_host[0].numel() * 4       # one forcings row per distinct IC valid time per step
    d2h = out_host[0].numel() * 4
    if not math.isfinite(checksum[0]):
        raise RuntimeError("end-to-end leg produced non-finite output")

    # ---------------- ensemble scores: the one collective of the forecast path (all_gather of the per-step sums)
    stats_info = None
    if stats is not None:
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e4.record()
        allsums = stats.gather()
        sc = stats.scores(allsums)
        e5.record()
        torch.cuda.synchronize()
        k = args.warmup + args.steps - 1                 # last step of the device-resident leg
        stats_info = {"gather_ms": max_over_ranks(e4.elapsed_time(e5)), "ics_scored": int(allsums.shape[1]),
                      "bytes_per_rank": int(stats.sums.numel() * 8),
                      "collective": "all_gather over NCCL" if world > 1 else "none (1 GPU)",
                      "last_step": {m: float(sc[m][k].mean()) for m in ("rmse", "crps", "ssr")}}

    # ---------------- strong scaling: the FIXED 12 x 64 workload of BASELINE.json configs[2] split over the ranks
    strong = None
    if not args.no_strong and args.solver == "scm":
        del ro
        torch.cuda.empty_cache()
        n_ic_s = 64
        traj_s = shard_trajectories(MEMBERS, n_ic_s, rank, world)
        s_steps = 2
        forc_s = syn.synthetic_forcings(cfg, n_ic_s + s_steps + 2, seed=0).to(dev)
        ro_s = EnsembleRollout(net, norm, forc_s, traj_s, use_graph=not args.no_graph, ic_times={j: j for j in range(n_ic_s)})
        st_s = None
        if len(traj_s) % MEMBERS == 0:
            import numpy as np
            from swift_b200.ensemble import EnsembleStatistics
            H_, W_ = cfg["img_resolution"]
            st_s = EnsembleStatistics(MEMBERS, len(traj_s) // MEMBERS, syn.IMG_CHANNELS, (H_, W_), np.linspace(-89.3, 89.3, H_),
                                      s_steps + 2, dev)
            ro_s.attach_statistics(st_s, torch.zeros(len(traj_s) // MEMBERS, syn.IMG_CHANNELS, H_, W_, device=dev))
        x0_s = torch.stack([syn.synthetic_fields(cfg, 1, seed=j)[1][0, :syn.IMG_CHANNELS] for _, j in traj_s[::MEMBERS]])
        ro_s.set_state(x0_s.repeat_interleave(MEMBERS, 0)[:len(traj_s)].to(dev))
        ro_s.step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(s_steps):
            ro_s.step()
        s1.record()
        barrier()
        ms_s = max_over_ranks(s0.elapsed_time(s1)) / s_steps
        g_ms = 0.0
        if st_s is not None:
            s2, s3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s2.record()
            st_s.scores(st_s.gather())
            s3.record()
            torch.cuda.synchronize()
            g_ms = max_over_ranks(s2.elapsed_time(s3))
        strong = {"workload": "12 members x 64 ICs = 768 trajectories in total (BASELINE.json configs[2]), split over the ranks",
                  "trajectories_per_gpu": len(traj_s), "steps_timed": s_steps, "ms_per_step": ms_s,
                  "member_steps_per_s": MEMBERS * n_ic_s / (ms_s / 1e3), "statistics_gather_ms": g_ms,
                  "time_to_solution_s": {"what": "46 080 member-steps (60 steps) + the all_gather of the ensemble statistics, "
                                                 "extrapolated from the timed steps", "value": 60 * ms_s / 1e3 + g_ms / 1e3}}
        del ro_s
        torch.cuda.empty_cache()

    # ---------------- other configurations of BASELINE.json, short runs recorded beside the headline (1 GPU only)
    extras = None
    if world == 1 and not args.no_extras and args.solver == "scm":
        extras = {}
        try:
            extras["trigflow_2s"] = short_2s(net, norm, cfg, dev)
            extras["training_step"] = short_train(dev)
        except Exception as e:                                   # an extra must not cost the headline line
            extras["error"] = f"{type(e).__name__}: {e}"[:300]

    # ---------------- roofline of the dominant kernel (SwiGLU up-projection GEMM: 42.5 % of the FLOPs), timed alone
    roof = dominant_kernel_roofline(eng, dev, min(args.chunk, 8))
    peaks, which = measured_peaks()
    calls = 39 if args.solver == "2s" else 1
    step_tflops = value / world * FLOP_PER_MEMBER_STEP * calls / 1e12
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = gpu_eager = None
    if world == 1 and not args.no_cpu:
        try:
            gpu_eager = gpu_eager_best(dev)
        except Exception as e:                                   # a baseline leg must not cost the headline line
            gpu_eager = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        v, per, cores, threads = cpu_member_steps_per_sec(2, 1)
        cpu = {"value": v, "unit": "member-steps/s", "cores": cores, "kind": "port",
               "sample": f"2 Swift-B sCM member-steps at batch 1 after 1 warm-up ({per:.2f} s each), fp32 oracle port, "
                         f"{threads} torch threads"}
    line = {
        "metric": "forecast member-steps/sec", "value": value, "unit": "member-steps/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": ("fp16" if eng.act_fp16 else "bf16") + " tensor-core operands (tcgen05 kind::f16), fp32 accumulate / LayerNorm / softmax; "
                 "residual stream " + ("stored as ONE fp16 value per element, updated in fp32 (x_single)"
                                       if (eng.act_fp16 and net.model.x_single) else "stored as a 16-bit [hi | lo] pair (~22 bits), updated in fp32"),
        "numerics_knobs": {"act_fp16": bool(eng.act_fp16), "x_single": bool(eng.act_fp16 and net.model.x_single),
                           "fuse_ln": int(net.model.effective_fuse_ln()), "one_step_rel_l2_max": "2.0e-3 (x_single) / 1.4e-3 (pair), bar 1e-2: "
                                                                                     "profiles/r02b_numerics.txt"},
        "data": "synthetic", "config": workload_config(n_gpus, torch.cuda.get_device_name(dev), args.solver),
        "clocks": clock_info,
        "e2e": {"value": e2e_value, "unit": "member-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps},
        "gpu_launches": launches,
        "per_rank": per_rank,
        "strong_scaling": strong,
        "extras": extras,
        "roofline": roof,
        "step_tflops_per_gpu": step_tflops,
        "step_frac_of_sustained_bf16": step_tflops / peaks["bf16_tflops_sustained"],
        "peaks": which,
        "cpu_baseline": cpu,
        "gpu_eager_baseline": gpu_eager,
        "ensemble_statistics": stats_info,
    }
    if world > 1:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


def dominant_kernel_roofline(eng, dev, chunk: int):
    """w1 (SwiGLU) GEMM at the shapes of the rollout: M = chunk*8192, N = 5632, K = 1056, timed alone with CUDA
    events on the launch stream (inputs + outputs of one launch exceed L2 for chunk >= 4)."""
    import ctypes as C
    import torch
    from swift_b200 import _lib

    g = eng.geom
    M, D, Dff = chunk * g.tokens, g.dim, g.dff
    adt = torch.float16 if eng.act_fp16 else torch.bfloat16
    f16 = int(eng.act_fp16)
    A = (torch.randn(M, D, device=dev) * 0.5).to(adt)
    out = torch.empty(M, Dff, device=dev, dtype=adt)
    W = eng._keep["w_1"][0]
    stream = torch.cuda.current_stream().cuda_stream
    lib = eng.lib
    for _ in range(3):
        _lib.check(lib.swb200_gemm_swiglu(eng.model.gemm_tile, f16, A.data_ptr(), D, W.data_ptr(), out.data_ptr(), M, D, Dff, stream))
    torch.cuda.synchronize()
    n = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        _lib.check(lib.swb200_gemm_swiglu(eng.model.gemm_tile, f16, A.data_ptr(), D, W.data_ptr(), out.data_ptr(), M, D, Dff, stream))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * M * (2 * Dff) * D
    peaks, which = measured_peaks()
    achieved = flops / (ms / 1e3) / 1e12
    return {"kernel": "gemm_tcgen05_kernel<NSUB=%d,CG=2,EPI_SWIGLU> (w1 up-projection, 42.5%% of FLOPs)" % (2 if eng.model.gemm_tile == 3 else 1), "bound": "tensor",
            "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape (chunk 8) from the committed
            # ncu --set full capture profiles/r02b_ncu_w1.txt (158.84 + 321.65 MB); algorithmic bytes: A 138 MB + W 12 MB
            # + h 369 MB = 519 MB
            "traffic": 480.49e6 if chunk == 8 else None, "traffic_unit": "bytes/launch (ncu dram read+write)",
            "traffic_source": "profiles/r02b_ncu_w1.txt (ncu --set full of this kernel at this shape, round 2; not re-measured "
                              "in this run: ncu is never run inside a timed bench)",
            "launch_ms": ms, "flops_per_launch": flops, "peak_source": f"{which} burst bf16"}


def short_2s(net, norm, cfg, dev):
    """BASELINE.json configs[3], short: one 6 h step of the TrigFlow 2S sampler (20 Heun steps = 39 denoiser calls per
    member-step) for one 12-member ensemble."""
    import torch
    from swift_b200 import synthetic as syn
    from swift_b200.rollout import EnsembleRollout
    traj = [(m, 0) for m in range(MEMBERS)]
    forc = syn.synthetic_forcings(cfg, 4, seed=0).to(dev)
    ro = EnsembleRollout(net, norm, forc, traj, solver="2s", solver_kwargs=dict(num_steps=20, sigma_min=0.02, sigma_max=200.0,
                                                                               auxiliary=0.6))
    ro.set_state(syn.synthetic_fields(cfg, 1, seed=0)[1][:, :syn.IMG_CHANNELS].expand(MEMBERS, -1, -1, -1).contiguous().to(dev))
    ro.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ro.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"member_steps_per_s": MEMBERS / (ms / 1e3), "denoiser_calls_per_s": 39 * MEMBERS / (ms / 1e3), "ms_per_step": ms,
            "trajectories": MEMBERS, "what": "TrigFlow 2S, 20 steps (39 calls per member-step), one 12-member ensemble, 1 step timed"}


def short_train(dev):
    """BASELINE.json configs[4], short: three sCM training steps (tangent forward, grad-enabled forward, backward, Muon +
    AuxAdam, EMA) at local batch 1 on this GPU; `python bench.py --mode train` under torchrun is the full line."""
    import torch
    t = _train_setup(dev, 0, 1, 1)
    for _ in range(2):
        t["one_step"](None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = t["one_step"](None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res = {"samples_per_s": 1e3 / ms, "ms_per_step": ms, "loss": float(out["loss"]),
           "what": "Swift-B sCM training step, local batch 1: tangent pass + grad-enabled forward + backward + MuonWithAuxAdam + EMA"}
    del t
    torch.cuda.empty_cache()
    return res


def run_tangent(args):
    """Not the headline metric: the forward-mode tangent forward of the sCM training loss (BASELINE.json configs[4], first
    part), Swift-B, batch 1 per call: (F, dF) = jvp(net, (x, t), (v_x, v_t)) through the C ABI."""
    import torch

    from swift_b200 import synthetic as syn
    from swift_b200.precond import PassPrecond

    dev = torch.device("cuda", 0)
    cfg = syn.SWIFT_B
    model_cfg = dict(_target_="swift_b200.swinv2.SwinV2", window_size=cfg["window_size"], shift_size=cfg["shift_size"],
                     patch_size=cfg["patch_size"], depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"])
    net = PassPrecond(model_cfg, img_resolution=cfg["img_resolution"], img_channels=syn.IMG_CHANNELS,
                      condition_channels=syn.COND_CHANNELS, auxiliary_dim=1, sigma_min=0.0, sigma_max=float("inf"))
    net.load_state_dict(syn.random_state_dict(cfg, seed=1, prefix="model."), strict=True)
    eng = net.to(dev).eval().model.engine()
    x = torch.randn(1, cfg["in_channels"], *cfg["img_resolution"], device=dev)
    dx = torch.randn_like(x)
    dx[:, syn.IMG_CHANNELS:] = 0
    t, dt = torch.tensor([0.9], device=dev), torch.tensor([0.45], device=dev)
    aux = torch.full((1, 1), 0.6, device=dev)
    for _ in range(max(3, args.warmup)):
        eng.forward_jvp(x, t, aux, dx, dt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.forward_jvp(x, t, aux, dx, dt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # the whole no-reverse-mode half of the sCM training step (training/loss.py:196-260): noised inputs, primal + tangent
    # pass, tangent target / loss / output cotangent -- three C-ABI calls (swift_b200/scm_target.py)
    from swift_b200.generate import era5_variables
    from swift_b200.scm_target import latitude_weights, scm_output_cotangent, variable_weights
    xs, cond = x[:, :syn.IMG_CHANNELS].contiguous(), x[:, syn.IMG_CHANNELS:].contiguous()
    zs = torch.randn_like(xs)
    w_lat, w_var = latitude_weights(cfg["img_resolution"][0], dev), variable_weights(era5_variables(), dev)
    kw = dict(condition=cond, auxiliary=0.6, tangent_warmup_kimg=3000, w_lat=w_lat, w_var=w_var)
    for _ in range(3):
        out = scm_output_cotangent(net, xs, t, zs, 1_500_000, **kw)
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        out = scm_output_cotangent(net, xs, t, zs, 1_500_000, **kw)
    e3.record()
    torch.cuda.synchronize()
    ms_target = e2.elapsed_time(e3) / args.steps
    # baseline leg: the same half of the loss as eager PyTorch ops on this GPU (oracle port; ONE torch.func.jvp pass, i.e.
    # without the reference's second, grad-enabled forward), fp32 with TF32 off
    eager = None
    if not args.no_cpu:
        try:
            from oracle import scm_loss_oracle as so, swinv2_oracle as orc
            torch.backends.cuda.matmul.allow_tf32 = False
            sd_gpu = {k: v.to(dev) for k, v in syn.random_state_dict(cfg, seed=1).items()}
            ocfg = orc.make_cfg(**cfg)
            onet = lambda a, b: orc.pass_precond(sd_gpu, ocfg, a, b, cond, 0.6)
            t4 = t.view(1, 1, 1, 1)
            with torch.no_grad():
                ref = so.scm_loss(onet, xs, t4, zs, 1_500_000, 3000, w_lat, w_var)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(2):
                    ref = so.scm_loss(onet, xs, t4, zs, 1_500_000, 3000, w_lat, w_var)
                torch.cuda.synchronize()
            eager = {"ms_per_sample": (time.perf_counter() - t0) / 2 * 1e3, "loss": float(ref["loss"]),
                     "cot_rel_l2_ours_vs_eager_fp32": float((out["cot"] - ref["cot"]).norm() / ref["cot"].norm())}
        except Exception as e:
            eager = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    flops = 2 * 2.72e12 + 0.5 * 5 * 8.86e9 * 12        # stacked GEMMs (2x) + five window products per layer instead of two
    print(json.dumps({"metric": "tangent-forward samples/sec (Swift-B, batch 1)", "value": 1e3 / ms, "unit": "samples/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
                      "higher_is_better": True, "dtype": "fp16 tensor-core operands, fp32 accumulate / dual kernels",
                      "data": "synthetic", "tflops": flops / ms / 1e9,
                      "scm_loss_forward_half": {"ms_per_sample": ms_target, "loss": float(out["loss"]),
                                                "what": "scm_output_cotangent: noised inputs + primal/tangent pass + tangent "
                                                        "target, loss and dL/dF_x (training/loss.py:196-260 without the "
                                                        "grad-enabled forward / backward)",
                                                "eager_pytorch_same_gpu": eager},
                      "config": {"workload": "Swift-B forward-mode tangent forward (jvp of the denoiser w.r.t. x and t)"}}))


def _train_setup(dev, rank: int, world: int, B: int):
    """Model, optimiser, EMA and the step closure of the sCM training benchmark (trainer.py:189-247)."""
    import torch

    from swift_b200 import synthetic as syn
    from swift_b200.generate import era5_variables
    from swift_b200.optim import MuonWithAuxAdam, swinv2_param_groups
    from swift_b200.precond import PassPrecond
    from swift_b200.scm_target import latitude_weights, variable_weights
    from swift_b200.training import scm_train_step

    def _ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    cfg = syn.SWIFT_B
    model_cfg = dict(_target_="swift_b200.swinv2.SwinV2", window_size=cfg["window_size"], shift_size=cfg["shift_size"],
                     patch_size=cfg["patch_size"], depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"])
    net = PassPrecond(model_cfg, img_resolution=cfg["img_resolution"], img_channels=syn.IMG_CHANNELS,
                      condition_channels=syn.COND_CHANNELS, auxiliary_dim=1, sigma_min=0.0, sigma_max=float("inf"))
    net.load_state_dict(syn.random_state_dict(cfg, seed=1, prefix="model."), strict=True)      # same weights on every rank
    net = net.to(dev).train()
    opt = MuonWithAuxAdam(swinv2_param_groups(net))                                            # configs/optimizer/muon.yaml
    ema = [p.detach().clone() for p in net.parameters()]
    H, W = cfg["img_resolution"]
    w_lat, w_var = latitude_weights(H, dev), variable_weights(era5_variables(), dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)                                 # every rank its own batch

    def batch():
        x = torch.randn(B, syn.IMG_CHANNELS, H, W, device=dev, generator=gen)
        cond = torch.randn(B, syn.COND_CHANNELS, H, W, device=dev, generator=gen)
        u = torch.rand(B, device=dev, generator=gen)
        sigma = torch.exp(math.log(0.02) + u * (math.log(200.0) - math.log(0.02)))             # loss/noise: loguniform(0.02, 200)
        return x, cond, torch.atan(sigma), torch.randn(x.shape, device=dev, generator=gen)

    ema_beta = 0.5 ** (B * world / 500e3)
    kimg = [1_500_000]
    phase = {}

    def one_step(reducer, timers=None):
        x, cond, t, z = batch()
        mark = (lambda k: timers.setdefault(k, []).append(_ev())) if timers is not None else (lambda k: None)
        mark("start")
        out = scm_train_step(net, x, t, z, kimg[0], condition=cond, auxiliary=0.6, reducer=reducer, tangent_warmup_kimg=3000,
                             w_lat=w_lat, w_var=w_var, timers=phase if timers is not None else None)
        mark("fwd_bwd")
        for buf in net.model._train_engine.grads.values():                                     # trainer.py:221-230
            torch.nan_to_num_(buf, nan=0.0, posinf=1e5, neginf=-1e5)
        opt.step()
        mark("optimizer")
        torch._foreach_lerp_(ema, [p.detach() for p in net.parameters()], 1.0 - ema_beta)      # trainer.py:240-241
        mark("ema")
        kimg[0] += B * world
        return out

    return {"one_step": one_step, "net": net, "phase": phase, "w_lat": w_lat, "w_var": w_var, "cfg": cfg}


def run_train(args):
    """BASELINE.json configs[4]: the sCM training step -- forward-mode tangent pass (loss, dL/dF), grad-enabled forward,
    backward, data-parallel gradient all-reduce overlapped with the backward, MuonWithAuxAdam, EMA -- Swift-B, local batch
    1 per GPU (configs/experiment/era5-swinv2-1.4-scm.yaml), synthetic ERA5-shaped batches.  Not the headline metric."""
    import torch
    import torch.distributed as dist

    from swift_b200 import synthetic as syn
    from swift_b200.generate import era5_variables
    from swift_b200.optim import MuonWithAuxAdam, swinv2_param_groups
    from swift_b200.precond import PassPrecond
    from swift_b200.scm_target import latitude_weights, variable_weights
    from swift_b200.training import GradientAllReduce, scm_train_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.train_batch
    ts = _train_setup(dev, rank, world, B)
    one_step, net, phase, w_lat, w_var, cfg = (ts[k] for k in ("one_step", "net", "phase", "w_lat", "w_var", "cfg"))

    def _ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def timed(reducer, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = one_step(reducer)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps, out

    reducer = GradientAllReduce(net.model)
    for _ in range(max(3, args.warmup)):
        one_step(reducer)
    clocks = ClockSampler(local)
    clocks.start()
    ms, out = timed(reducer, args.steps)
    clock_info = clocks.stop()
    torch.cuda.synchronize()
    comm_ms = reducer.comm_ms() / max(1, args.steps + max(3, args.warmup))
    bytes_per_step = reducer.bytes // max(1, args.steps + max(3, args.warmup))
    ms_nocomm = None
    if world > 1:                                            # the same steps without the all-reduce: what the overlap hides
        ms_nocomm, _ = timed(GradientAllReduce(net.model, enabled=False), max(3, args.steps // 2))
    timers = {}
    for _ in range(3):                                       # phase breakdown (events between the phases)
        one_step(reducer, timers)
    torch.cuda.synchronize()
    order = ["start", "fwd_bwd", "optimizer", "ema"]
    brk = {b: sum(x0.elapsed_time(x1) for x0, x1 in zip(timers[a], timers[b])) / 3 for a, b in zip(order, order[1:])}
    order = ["t0", "pack_tangent", "tangent_loss", "pack_train", "train_forward", "backward"]
    brk.update({"fwd_bwd." + b: sum(x0.elapsed_time(x1) for x0, x1 in zip(phase[a], phase[b])) / 3 for a, b in zip(order, order[1:])})
    loss = float(out["loss"])
    if not math.isfinite(loss):
        raise RuntimeError("training bench produced a non-finite loss")
    eager = None
    if rank == 0 and world == 1 and not args.no_cpu:
        eager = eager_training_step_ms(dev, cfg, w_lat, w_var)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # FLOPs of one sample-step: tangent pass 2 x forward GEMMs (+ dual products), grad-enabled forward 1 x, backward 2 x
    flops = (2 + 1 + 2) * FLOP_PER_MEMBER_STEP
    line = {"metric": "sCM training samples/sec", "value": B * world / (ms / 1e3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 tensor-core operands, fp32 accumulate / gradients / optimiser state "
            "(tangent pass: fp16 operands)", "data": "synthetic",
            "config": {"workload": "Swift-B sCM training step (tangent forward + grad-enabled forward + backward + all-reduce + "
                                   "MuonWithAuxAdam + EMA), local batch %d per GPU" % B, "global_batch": B * world,
                       "parallelism": f"dp{world}", "device": torch.cuda.get_device_name(dev)},
            "clocks": clock_info, "loss": loss, "breakdown_ms": brk,
            "step_tflops_per_gpu": B * flops / (ms / 1e3) / 1e12,
            "allreduce": {"bytes_per_step": int(bytes_per_step), "comm_ms_per_step": comm_ms, "step_ms_without_allreduce": ms_nocomm,
                          "exposed_ms": None if ms_nocomm is None else max(0.0, ms - ms_nocomm),
                          "hidden_ms": None if ms_nocomm is None else max(0.0, comm_ms - max(0.0, ms - ms_nocomm)),
                          "how": "per-stage NCCL all-reduce on a side stream behind an event (training.GradientAllReduce)"},
            "eager_pytorch_same_gpu": eager}
    if world > 1:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


def eager_training_step_ms(dev, cfg, w_lat, w_var):
    """Baseline leg: the reference ALGORITHM of one training step's forward + backward (oracle port: torch.func.jvp pass,
    grad-enabled forward, backward) as eager PyTorch ops on this GPU, fp32 with TF32 off and under bf16 autocast."""
    import torch
    from oracle import scm_loss_oracle as so, swinv2_oracle as orc
    from swift_b200 import synthetic as syn
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        sd = {k: v.to(dev) for k, v in syn.random_state_dict(cfg, seed=1).items()}
        ocfg = orc.make_cfg(**cfg)
        x = torch.randn(1, syn.IMG_CHANNELS, *cfg["img_resolution"], device=dev)
        cond = torch.randn(1, syn.COND_CHANNELS, *cfg["img_resolution"], device=dev)
        t4, z = torch.full((1, 1, 1, 1), 0.9, device=dev), torch.randn_like(x)
        net_of = lambda p: (lambda a, b: orc.pass_precond(p, ocfg, a, b, cond, 0.6))
        res = {}
        for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            with ctx:
                so.scm_parameter_gradients(net_of, sd, x, t4, z, 1_500_000, 3000, w_lat, w_var)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(2):
                    so.scm_parameter_gradients(net_of, sd, x, t4, z, 1_500_000, 3000, w_lat, w_var)
                torch.cuda.synchronize()
            res[name + "_ms_per_sample"] = (time.perf_counter() - t0) / 2 * 1e3
        res["what"] = "oracle port: jvp pass + grad-enabled forward + backward, no optimiser"
        return res
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunk", type=int, default=24, help="trajectories per kernel launch sequence")
    ap.add_argument("--solver", default="scm", choices=["scm", "2s"],
                    help="scm: Swift 1-step consistency sampler (headline); 2s: TrigFlow diffusion baseline, 20 Heun steps = "
                         "39 denoiser calls per 6 h step (BASELINE.json configs[3])")
    ap.add_argument("--mode", default="rollout", choices=["rollout", "tangent", "train"],
                    help="rollout: the headline forecast benchmark; tangent: the forward-mode tangent forward of the sCM loss; "
                         "train: the whole sCM training step, data parallel under torchrun (BASELINE.json configs[4])")
    ap.add_argument("--train-batch", type=int, default=1, help="--mode train: samples per GPU and step")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (0 = same as --steps)")
    ap.add_argument("--fuse-ln", type=int, default=-1, help="override SwinV2.fuse_ln (bit 0: wo, bit 1: w2; 0 = separate LN kernel)")
    ap.add_argument("--act-bf16", action="store_true", help="bf16 GEMM operands (fp16 attention internals, [hi | lo] residual pair) "
                                                            "instead of the fp16 default: the format north_star names, 8.4e-3 one-step error")
    ap.add_argument("--split-embed", type=int, default=-1, help="override SwinV2.split_embed ([hi | lo] operand of the patch-embed GEMM)")
    ap.add_argument("--x-single", type=int, default=-1, help="override SwinV2.x_single (1: one fp16 value per residual element, 0: [hi | lo] pair)")
    ap.add_argument("--no-stats", action="store_true", help="do not accumulate the on-device ensemble scores")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (fixed 12 x 64 workload)")
    ap.add_argument("--no-extras", action="store_true", help="skip the short TrigFlow-2S / training-step summaries")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "tangent":
        run_tangent(args)
    elif args.mode == "train":
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
