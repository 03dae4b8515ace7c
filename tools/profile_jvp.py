"""One Swift-B tangent forward (batch 1) for an ncu launch list:  ncu --metrics gpu__time_duration.sum ... python tools/profile_jvp.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from swift_b200 import synthetic as syn
from swift_b200.precond import PassPrecond

cfg = syn.SWIFT_B
mc = dict(_target_="swift_b200.swinv2.SwinV2", window_size=cfg["window_size"], shift_size=cfg["shift_size"],
          patch_size=cfg["patch_size"], depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"])
net = PassPrecond(mc, img_resolution=cfg["img_resolution"], img_channels=69, condition_channels=72, auxiliary_dim=1)
net.load_state_dict(syn.random_state_dict(cfg, seed=1, prefix="model."), strict=True)
net = net.cuda().eval()
eng = net.model.engine()
x = torch.randn(1, 141, 128, 256, device="cuda")
dx = torch.randn_like(x)
t = torch.tensor([0.9], device="cuda")
dt = torch.tensor([0.4], device="cuda")
aux = torch.full((1, 1), 0.6, device="cuda")
for _ in range(2):
    eng.forward_jvp(x, t, aux, dx, dt)
torch.cuda.synchronize()
print("done")
