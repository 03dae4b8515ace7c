"""In-situ per-kernel time of the rollout step at the clocks of the real (power-capped) run:
    python tools/trace_step.py [chunk] [steps]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from swift_b200 import _lib, synthetic as syn
from swift_b200.precond import PassPrecond
from swift_b200.rollout import EnsembleRollout, Normalizers, shard_trajectories


def main():
    chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    cfg = syn.SWIFT_B
    mc = dict(_target_="swift_b200.swinv2.SwinV2", window_size=cfg["window_size"], shift_size=cfg["shift_size"],
              patch_size=cfg["patch_size"], depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"])
    net = PassPrecond(mc, img_resolution=cfg["img_resolution"], img_channels=69, condition_channels=72, auxiliary_dim=1)
    net.load_state_dict(syn.random_state_dict(cfg, seed=1, prefix="model."), strict=True)
    net = net.cuda().eval()
    net.model.max_chunk = chunk
    if len(sys.argv) > 3:
        net.model.fuse_ln = int(sys.argv[3])
    traj = shard_trajectories(12, 8, 0, 1)
    forc = syn.synthetic_forcings(cfg, steps + 16, seed=0).cuda()
    ro = EnsembleRollout(net, Normalizers.synthetic(69, "cuda"), forc, traj, use_graph=False, ic_times=list(range(8)))
    ro.set_state(torch.randn(len(traj), 69, 128, 256, device="cuda"))
    for _ in range(3):
        ro.step()
    torch.cuda.synchronize()
    lib = _lib.lib()
    lib.swb200_trace_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ro.step()
    e1.record()
    buf = ctypes.create_string_buffer(4096)
    _lib.check(lib.swb200_trace_report(buf, 4096))
    lib.swb200_trace_enable(0)
    total = e0.elapsed_time(e1)
    rep = json.loads(buf.value.decode())
    ksum = sum(v["ms"] for v in rep.values())
    print(f"{steps} steps x {len(traj)} trajectories, chunk {chunk}: {total:.1f} ms wall (events), kernels {ksum:.1f} ms")
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"  {k:18s} {v['ms']:9.2f} ms {100 * v['ms'] / ksum:5.1f}%  {v['launches']:5d} launches  "
              f"{1e3 * v['ms'] / max(1, v['launches']):8.1f} us avg")


if __name__ == "__main__":
    main()
