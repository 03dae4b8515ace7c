"""Golden vectors for the DISTILLATION branch of the sCM loss, from the REAL reference ``SCMLoss(distillation=True)`` with a
``net_pretrained`` teacher (loss.py:205-210): the teacher is the same reference architecture with the seed-2 fixture weights.

Run in the build container only:    python tests/golden/make_scm_distill_golden.py    -> tests/golden/scm_distill.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import build_reference, install_shims  # noqa: E402
from make_scm_loss_golden import NOISE, VARIABLES, _Dataset, _DDPLike  # noqa: E402


def main():
    install_shims()
    torch.set_num_threads(os.cpu_count())
    from swift.training.loss import NOISE_SAMPLING_METHODS, SCMLoss
    from swift_b200 import synthetic as syn

    out = {}
    for name, cfg in (("tiny", syn.SWIFT_TINY), ("small", syn.SWIFT_SMALL)):
        n_img = cfg["out_channels"]
        H, W = cfg["img_resolution"]
        net = _DDPLike(build_reference(cfg, syn.random_state_dict(cfg, seed=1), img_channels=n_img).train())
        teacher = build_reference(cfg, syn.random_state_dict(cfg, seed=2), img_channels=n_img).eval()
        x, cond = syn.synthetic_fields(cfg, 2, seed=5)
        seed, step, warm = 31, 2_000_000, 3000
        loss_fn = SCMLoss(_Dataset(VARIABLES[:n_img], (n_img, H, W)), dict(NOISE), sigma_data=1.0, tangent_warmup_kimg=warm,
                          distillation=True)
        torch.manual_seed(seed)
        tau = NOISE_SAMPLING_METHODS["loguniform"](x, NOISE["sigma_min"], NOISE["sigma_max"])
        z = torch.randn_like(x) * 1.0
        torch.manual_seed(seed)
        net.grads.clear(), net.outputs.clear()
        loss = loss_fn(net, x, step, condition=cond, auxiliary=0.6, net_pretrained=teacher)
        loss.backward()
        k = name + "_"
        out[k + "t"] = torch.atan(tau / 1.0).numpy()
        out[k + "z"] = z.numpy()
        out[k + "step_warm"] = np.array([step, warm], dtype=np.int64)
        out[k + "loss"] = np.array(loss.item(), dtype=np.float64)
        out[k + "cot"] = net.grads[0].numpy()
        print(k, "loss", loss.item(), "cot norm", float(net.grads[0].norm()))
    np.savez_compressed(os.path.join(HERE, "scm_distill.npz"), **out)


if __name__ == "__main__":
    main()
