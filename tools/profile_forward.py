"""Small workload for ncu: two eager 6 h steps of 8 Swift-B trajectories (one chunk of 8 -> M = 65536 token rows), i.e. the
kernel sequence of the rollout without the CUDA graph.   ncu ... python tools/profile_forward.py [fuse_ln]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from swift_b200 import synthetic as syn
from swift_b200.precond import PassPrecond
from swift_b200.rollout import EnsembleRollout, Normalizers


def main():
    cfg = syn.SWIFT_B
    mc = dict(_target_="swift_b200.swinv2.SwinV2", window_size=cfg["window_size"], shift_size=cfg["shift_size"],
              patch_size=cfg["patch_size"], depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"])
    net = PassPrecond(mc, img_resolution=cfg["img_resolution"], img_channels=69, condition_channels=72, auxiliary_dim=1)
    net.load_state_dict(syn.random_state_dict(cfg, seed=1, prefix="model."), strict=True)
    net = net.cuda().eval()
    net.model.max_chunk = 8
    if len(sys.argv) > 1:
        net.model.fuse_ln = int(sys.argv[1])
    traj = [(m, 0) for m in range(8)]
    forc = syn.synthetic_forcings(cfg, 24, seed=0).cuda()
    ro = EnsembleRollout(net, Normalizers.synthetic(69, "cuda"), forc, traj, use_graph=False, ic_times=list(range(8)))
    ro.set_state(torch.randn(len(traj), 69, 128, 256, device="cuda"))
    for _ in range(2):
        ro.step()
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
