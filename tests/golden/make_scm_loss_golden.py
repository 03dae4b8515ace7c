"""Golden vectors for the sCM training loss, from the REAL reference ``swift.training.loss.SCMLoss`` (loss.py:162-260).

Run in the build container only:    python tests/golden/make_scm_loss_golden.py

``SCMLoss.forward`` returns one scalar, but what the backward pass of the training step consumes is its gradient with
respect to the network output F_x -- the (normalised, detached) tangent target g scaled by the loss weights.  It is
recorded here with a tensor hook on the output of the reference network's grad-enabled call, together with the loss, on
the tiny / small fixtures of ``swift_b200.synthetic``.  For the first case of each fixture the parameter gradients of
``loss.backward()`` are recorded too (norm of every tensor, strided samples of nine of them): the target of the reverse
pass that is not built yet.  The loss draws tau and z from the global RNG; both are reproduced
by re-seeding and repeating the reference's own calls (``loguniform``, then ``randn_like``) and stored, so a restatement
can be checked as a deterministic function of (x, condition, t, z, step).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import build_reference, install_shims  # noqa: E402

VARIABLES = ["2m_temperature", "10m_u_component_of_wind", "mean_sea_level_pressure", "geopotential_500",
             "temperature_850", "specific_humidity_700"]
GRAD_SAMPLES = ["model.head.head.0.weight", "model.transformer.layers.1.1.w1.weight", "model.transformer.layers.0.0.to_qkv.weight",
                "model.transformer.layers.0.0.scale", "model.transformer.layers.1.0.norm.modulation.weight", "model.pos_embed",
                "model.patch_embed.emb.weight", "model.latent_embed.l1.weight", "model.auxiliary_embed.weight"]
GRAD_STRIDE = 31
NOISE = dict(dist="loguniform", sigma_min=0.02, sigma_max=200.0)       # experiment/era5-swinv2-1.4-scm.yaml:13-16


class _DDPLike(torch.nn.Module):
    """What SCMLoss sees in training: a wrapper with ``.module`` whose own call is the grad-enabled forward."""

    def __init__(self, module):
        super().__init__()
        self.module = module
        self.grads = []
        self.outputs = []

    def forward(self, *a, **k):
        out = self.module(*a, **k)
        out.register_hook(lambda g: self.grads.append(g.detach().clone()))
        self.outputs.append(out.detach().clone())
        return out


class _Dataset:
    def __init__(self, variables, shape):
        self.variables, self._shape = variables, shape


def main():
    install_shims()
    torch.set_num_threads(os.cpu_count())
    from swift.training.loss import NOISE_SAMPLING_METHODS, SCMLoss
    from swift_b200 import synthetic as syn

    out = {}
    for name, cfg in (("tiny", syn.SWIFT_TINY), ("small", syn.SWIFT_SMALL)):
        n_img = cfg["out_channels"]
        variables = VARIABLES[:n_img]
        H, W = cfg["img_resolution"]
        sd = syn.random_state_dict(cfg, seed=1)
        net = _DDPLike(build_reference(cfg, sd, img_channels=n_img).train())
        x, cond = syn.synthetic_fields(cfg, 2, seed=5)
        for case, (seed, step, warm) in enumerate(((11, 500_000, 3000), (12, 10_000_000, 3000), (13, 0, 0))):
            loss_fn = SCMLoss(_Dataset(variables, (n_img, H, W)), dict(NOISE), sigma_data=1.0, tangent_warmup_kimg=warm)
            torch.manual_seed(seed)
            tau = NOISE_SAMPLING_METHODS["loguniform"](x, NOISE["sigma_min"], NOISE["sigma_max"])
            z = torch.randn_like(x) * 1.0
            torch.manual_seed(seed)
            net.grads.clear(), net.outputs.clear()
            net.zero_grad()
            loss = loss_fn(net, x, step, condition=cond, auxiliary=0.6)
            loss.backward()
            k = f"{name}_{case}_"
            out[k + "t"] = torch.atan(tau / 1.0).numpy()
            out[k + "z"] = z.numpy()
            out[k + "step_warm"] = np.array([step, warm], dtype=np.int64)
            out[k + "loss"] = np.array(loss.item(), dtype=np.float64)
            out[k + "F"] = net.outputs[0].numpy()
            out[k + "cot"] = net.grads[0].numpy()
            if case == 0:
                # what the reverse pass must deliver: d loss / d parameter (norm of every tensor + strided samples of a few)
                named = dict(net.module.named_parameters())
                names = sorted(n for n, p_ in named.items() if p_.grad is not None)
                out[k + "grad_names"] = np.array(names)
                out[k + "grad_norms"] = np.array([float(named[n].grad.norm()) for n in names], dtype=np.float64)
                for n in GRAD_SAMPLES:
                    out[k + "grad:" + n] = named[n].grad.flatten()[::GRAD_STRIDE].numpy().copy()
            print(k, "loss", loss.item(), "cot norm", float(net.grads[0].norm()))
        out[name + "_w_lat"] = loss_fn.w_lat.numpy()
        out[name + "_w_var"] = loss_fn.w_var.numpy()
    np.savez_compressed(os.path.join(HERE, "scm_loss.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
