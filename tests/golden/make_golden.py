"""Generate golden vectors by executing the REAL reference (stockeh/swift) on seeded fixtures.

Run in the build container only (``/root/reference`` is mounted there and nowhere else):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so the pin for
``oracle/swinv2_oracle.py`` is the output of the reference's own modules --
``swift.models.swinv2.SwinV2`` inside ``swift.models.precond.PassPrecond`` driven by
``swift.generating.diffusion.DiffusionSampler`` -- on the fixtures of
``swift_b200.synthetic``.  Three import stubs are needed and nothing in the reference is
edited: ``omegaconf`` (type hints only), ``ezpz.get_logger`` and ``hydra.utils.instantiate``
(a 6-line ``_target_`` importer), none of which contribute arithmetic.

Outputs (committed):
  tiny.npz / small.npz  full tensors: forward taps, scm 1-step / 2-step, 2s (3 steps) outputs
  swift_b.npz           Swift-B, B=1: per-channel L2 norms, a strided sub-sample of the 1-step
                        scm output, and block taps' norms (the full tensors are 9 MB each).
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF_SRC = "/root/reference/src"
TAP_STRIDE = 8


def install_shims():
    om = types.ModuleType("omegaconf")
    om.ListConfig = type("ListConfig", (list,), {})
    om.DictConfig = dict
    sys.modules["omegaconf"] = om
    ez = types.ModuleType("ezpz")
    ez.get_logger = logging.getLogger
    sys.modules["ezpz"] = ez
    hy = types.ModuleType("hydra")
    hu = types.ModuleType("hydra.utils")

    def instantiate(config, **kwargs):
        cfg = dict(config)
        cfg.update(kwargs)
        for k in ("_convert_", "_recursive_", "_partial_"):
            cfg.pop(k, None)
        mod, _, name = cfg.pop("_target_").rpartition(".")
        return getattr(importlib.import_module(mod), name)(**cfg)

    hu.instantiate = instantiate
    hy.utils = hu
    sys.modules["hydra"] = hy
    sys.modules["hydra.utils"] = hu
    sys.path.insert(0, REF_SRC)


def build_reference(cfg: dict, sd: dict, img_channels: int, logvar: bool = False):
    from swift.models.precond import PassPrecond

    model_cfg = dict(_target_="swift.models.swinv2.SwinV2",
                     window_size=cfg["window_size"], shift_size=cfg["shift_size"],
                     patch_size=cfg["patch_size"], depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"],
                     logvar=logvar, timestep_weight=1.0)
    net = PassPrecond(model_cfg, img_resolution=cfg["img_resolution"], img_channels=img_channels,
                      condition_channels=cfg["in_channels"] - img_channels,
                      auxiliary_dim=cfg["auxiliary_dim"], sigma_min=0.0, sigma_max=float("inf"), sigma_data=1.0)
    net.load_state_dict({"model." + k: v for k, v in sd.items()}, strict=True)
    return net.eval()


def block_taps(net, x_in, t, aux):
    """Hook the reference modules to record intermediate token tensors."""
    taps = {}
    m = net.model
    hooks = [m.patch_embed.register_forward_hook(lambda mod, i, o: taps.__setitem__("patch_embed", o.detach())),
             m.latent_embed.register_forward_hook(lambda mod, i, o: taps.__setitem__("cond", o.detach()))]
    # the block loop lives inside SwinTransformer.forward; record each ff output's *input* residual sum via
    # a pre-hook on the next attention and the final transformer output
    for i, (attn, ff) in enumerate(m.transformer.layers):
        hooks.append(ff.register_forward_hook(
            lambda mod, inp, out, i=i: taps.__setitem__(f"block{i}", (inp[0] + out).detach())))
    with torch.no_grad():
        y = net(x_in[0], t, x_in[1], aux)
    for h in hooks:
        h.remove()
    return y, taps


def main():
    install_shims()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    from swift.generating.diffusion import DiffusionSampler
    from swift_b200 import synthetic as syn

    for name, cfg, batch in (("tiny", syn.SWIFT_TINY, 2), ("small", syn.SWIFT_SMALL, 2)):
        sd = syn.random_state_dict(cfg, seed=1)
        lat, cond = syn.synthetic_fields(cfg, batch, seed=3)
        net = build_reference(cfg, sd, img_channels=cfg["out_channels"])
        out = {}
        # plain module forward at a generic (t, aux), per-sample t
        t = torch.tensor([0.3, 1.2])[:batch]
        aux = torch.tensor([[0.6], [1.2]])[:batch]
        y, taps = block_taps(net, (lat, cond), t, aux)
        out["fwd_t"] = t.numpy()
        out["fwd_aux"] = aux.numpy()
        out["fwd_y"] = y.numpy()
        for k, v in taps.items():      # token taps are stored every TAP_STRIDE-th token to keep the files small
            out["tap_" + k] = v.numpy() if v.dim() == 2 else v[:, ::TAP_STRIDE].numpy()
        S = DiffusionSampler(net)
        out["scm1"] = S.scm_solver(latents=lat, condition=cond, auxiliary=0.6, num_steps=1,
                                   sigma_min=0.02, sigma_max=200.0).numpy()
        g = torch.Generator().manual_seed(7)
        z = torch.randn(lat.shape, generator=g)
        out["scm2_noise"] = z.numpy()
        out["scm2"] = S.scm_solver(latents=lat, condition=cond, auxiliary=0.6, num_steps=2,
                                   sigma_min=0.02, sigma_max=200.0, randn_like=lambda x: z).numpy()
        out["scm3"] = S.scm_solver(latents=lat, condition=cond, auxiliary=0.6, num_steps=3,
                                   sigma_min=0.02, sigma_max=200.0, randn_like=lambda x: z).numpy()
        out["dpm2s_3"] = S.dpm_solver_2s(latents=lat, condition=cond, auxiliary=0.6, num_steps=3,
                                         sigma_min=0.02, sigma_max=200.0).numpy()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})

    # Swift-B, B=1: digest only
    cfg = syn.SWIFT_B
    sd = syn.random_state_dict(cfg, seed=1)
    lat, cond = syn.synthetic_fields(cfg, 1, seed=0)
    net = build_reference(cfg, sd, img_channels=syn.IMG_CHANNELS)
    t = torch.tensor([np.pi / 2], dtype=torch.float32)
    y, taps = block_taps(net, (lat, cond), t, 0.6)
    S = DiffusionSampler(net)
    scm1 = S.scm_solver(latents=lat, condition=cond, auxiliary=0.6, num_steps=1, sigma_min=0.02, sigma_max=200.0)
    out = {
        "scm1_channel_l2": scm1.flatten(2).norm(dim=-1).numpy(),
        "scm1_sub": scm1[:, :, ::8, ::8].numpy(),
        "scm1_row": scm1[0, ::17, 37, :].numpy(),
        "fwd_channel_l2": y.flatten(2).norm(dim=-1).numpy(),
        "cond": taps["cond"].numpy(),
    }
    for k, v in taps.items():
        if k.startswith("block") or k == "patch_embed":
            out["tapnorm_" + k] = v.norm(dim=-1).numpy()[:, ::64]
            out["tapsub_" + k] = v[:, ::512, ::33].numpy()
    np.savez_compressed(os.path.join(HERE, "swift_b.npz"), **out)
    print("swift_b", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
