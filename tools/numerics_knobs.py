"""Swift-B, one sCM step vs the fp32 oracle (on the GPU, TF32 off) for a list of knob settings:
    python tools/numerics_knobs.py "x_single=1" "x_single=1,split_embed=0" ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from oracle import swinv2_oracle as orc
from swift_b200 import synthetic as syn
from swift_b200.sampler import DiffusionSampler
from test_gpu_forward import build_net, per_field_rel_l2


def main():
    cfg = syn.SWIFT_B
    net, sd = build_net(cfg, img_channels=syn.IMG_CHANNELS)
    lat, cond = syn.synthetic_fields(cfg, 1, seed=0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ocfg = orc.make_cfg(**cfg)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = orc.scm_solver(lambda x, t, c, a: orc.pass_precond(sd_gpu, ocfg, x, t, c, a), lat.cuda(), cond.cuda(), 0.6, num_steps=1)
    for spec in sys.argv[1:] or ["x_single=1"]:
        for kv in spec.split(","):
            k, v = kv.split("=")
            setattr(net.model, k, type(getattr(net.model, k))(int(v)))
        y = DiffusionSampler(net).scm_solver(latents=lat.cuda(), condition=cond.cuda(), auxiliary=0.6, num_steps=1, sigma_min=0.02,
                                             sigma_max=200.0)
        err = per_field_rel_l2(y, ref)
        print(f"{spec:40s} per-field rel-L2 max {err.max():.4e} mean {err.mean():.4e}")


if __name__ == "__main__":
    main()
