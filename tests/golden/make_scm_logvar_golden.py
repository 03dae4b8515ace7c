"""Golden vectors for the sCM training loss WITH A LOGVAR HEAD, from the REAL reference ``swift.training.loss.SCMLoss``
(loss.py:222-232, :252-258) around the real ``swift.models.swinv2.SwinV2(logvar=True)``.

Run in the build container only:    python tests/golden/make_scm_logvar_golden.py    -> tests/golden/scm_logvar.npz

Recorded on the tiny / small fixtures of ``swift_b200.synthetic`` (``random_state_dict(cfg, logvar=True)``): the draws (t, z),
the loss, logvar, the gradients arriving at the two outputs of the grad-enabled call (tensor hooks: dL/dF_x and dL/dlogvar)
and, after ``loss.backward()``, the norm of every parameter gradient plus strided samples of a few (the logvar head's own and
the latent MLP's, which the head's gradient reaches through the conditioning vector).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import build_reference, install_shims  # noqa: E402
from make_scm_loss_golden import NOISE, VARIABLES, _Dataset  # noqa: E402

GRAD_SAMPLES = ["model.logvar_embed.weight", "model.logvar_embed.bias", "model.latent_embed.l1.weight", "model.latent_embed.l2.weight",
                "model.head.head.0.weight", "model.transformer.layers.0.0.to_qkv.weight", "model.auxiliary_embed.weight"]
GRAD_STRIDE = 31


class _DDPLike(torch.nn.Module):
    def __init__(self, module):
        super().__init__()
        self.module = module
        self.seen = {}

    def forward(self, *a, **k):
        out = self.module(*a, **k)
        F_x, lv = out
        F_x.register_hook(lambda g: self.seen.__setitem__("cot", g.detach().clone()))
        lv.register_hook(lambda g: self.seen.__setitem__("dlogvar", g.detach().clone()))
        self.seen["F"], self.seen["logvar"] = F_x.detach().clone(), lv.detach().clone()
        return out


def main():
    install_shims()
    torch.set_num_threads(os.cpu_count())
    from swift.training.loss import NOISE_SAMPLING_METHODS, SCMLoss
    from swift_b200 import synthetic as syn

    out = {}
    for name, cfg in (("tiny", syn.SWIFT_TINY), ("small", syn.SWIFT_SMALL)):
        n_img = cfg["out_channels"]
        H, W = cfg["img_resolution"]
        sd = syn.random_state_dict(cfg, seed=1, logvar=True)
        net = _DDPLike(build_reference(cfg, sd, img_channels=n_img, logvar=True).train())
        x, cond = syn.synthetic_fields(cfg, 2, seed=5)
        seed, step, warm = 21, 500_000, 3000
        loss_fn = SCMLoss(_Dataset(VARIABLES[:n_img], (n_img, H, W)), dict(NOISE), sigma_data=1.0, tangent_warmup_kimg=warm)
        torch.manual_seed(seed)
        tau = NOISE_SAMPLING_METHODS["loguniform"](x, NOISE["sigma_min"], NOISE["sigma_max"])
        z = torch.randn_like(x) * 1.0
        torch.manual_seed(seed)
        net.zero_grad()
        loss = loss_fn(net, x, step, condition=cond, auxiliary=0.6)
        loss.backward()
        k = name + "_"
        out[k + "t"] = torch.atan(tau / 1.0).numpy()
        out[k + "z"] = z.numpy()
        out[k + "step_warm"] = np.array([step, warm], dtype=np.int64)
        out[k + "loss"] = np.array(loss.item(), dtype=np.float64)
        out[k + "logvar"] = net.seen["logvar"].numpy()
        out[k + "dlogvar"] = net.seen["dlogvar"].numpy()
        out[k + "cot"] = net.seen["cot"].numpy()
        named = dict(net.module.named_parameters())
        names = sorted(n for n, p_ in named.items() if p_.grad is not None)
        out[k + "grad_names"] = np.array(names)
        out[k + "grad_norms"] = np.array([float(named[n].grad.norm()) for n in names], dtype=np.float64)
        for n in GRAD_SAMPLES:
            out[k + "grad:" + n] = named[n].grad.flatten()[::GRAD_STRIDE].numpy().copy()
        out[k + "w_lat"] = loss_fn.w_lat.numpy()
        out[k + "w_var"] = loss_fn.w_var.numpy()
        print(k, "loss", loss.item(), "logvar", net.seen["logvar"].tolist(), "dlogvar", net.seen["dlogvar"].flatten().tolist())
    np.savez_compressed(os.path.join(HERE, "scm_logvar.npz"), **out)


if __name__ == "__main__":
    main()
