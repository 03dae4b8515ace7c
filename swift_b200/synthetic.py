"""Seeded synthetic weights and inputs for the Swift forecast path.

There is no ERA5 data and no trained checkpoint in this environment, so the
benchmark, smoke test and parity fixtures all run on synthetic tensors of the
exact shapes the reference uses (``configs/experiment/era5-swinv2-1.4-scm.yaml``
+ ``configs/data/era5-flare-1.4.yaml``: 69 variables, 3 forcings, 128x256 grid).

Why not the reference's default init: ``SwinV2._init_weights``
(models/swinv2.py:295-303) zero-initialises every ``modulation`` Linear and the
``head`` Linear, so an untouched random-init network returns exactly 0 and any
parity check on it is vacuous.  Here *every* tensor of the state dict
(including biases and LayerNorm affine terms) is drawn from a seeded generator
at the reference's scales, and the same dict is loaded (strict) into whichever
implementation is under test.

The 2-D weights of the six tensor-core GEMM families (patch-embed, to_qkv, wo,
w1, w2, head) are rounded to bf16-representable fp32 values when
``bf16_weights=True``: the CUDA path stores those weights in bf16, so sharing
representable values isolates activation rounding, which is what the 1e-2
per-field tolerance of the north-star is stated for.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

# Swift-B: configs/experiment/era5-swinv2-1.4-scm.yaml:21-28, data/era5-flare-1.4.yaml
SWIFT_B = dict(img_resolution=[128, 256], in_channels=141, out_channels=69, window_size=[16, 16],
               shift_size=[8, 8], patch_size=[2, 2], depth=12, dim=1056, heads=12, auxiliary_dim=1)
IMG_CHANNELS = 69          # prognostic variables
COND_CHANNELS = 72         # previous state (69) + forcings (3)
N_FORCINGS = 3

# A reduced configuration that keeps every structural property of Swift-B that the kernels
# specialise on (head_dim 88, dim a multiple of 264 so that int(8/3*dim) is a multiple of 176,
# 16x16 windows shifted by 8, patch 2x2, asymmetric patch orders) at a size the CPU oracle
# runs in well under a second: 32x64 image -> 16x32 tokens -> 1x2 windows.
SWIFT_TINY = dict(img_resolution=[32, 64], in_channels=13, out_channels=5, window_size=[16, 16],
                  shift_size=[8, 8], patch_size=[2, 2], depth=2, dim=264, heads=3, auxiliary_dim=1)
# 64x64 image -> 32x32 tokens -> 2x2 windows (exercises wrap-around in both axes), depth 3.
SWIFT_SMALL = dict(img_resolution=[64, 64], in_channels=11, out_channels=4, window_size=[16, 16],
                   shift_size=[8, 8], patch_size=[2, 2], depth=3, dim=528, heads=6, auxiliary_dim=1)

_GEMM_KEYS = ("patch_embed.emb.weight", "to_qkv.weight", "wo.weight", "w1.weight", "w2.weight",
              "head.head.0.weight")


def _pair(v) -> Tuple[int, int]:
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


def state_dict_shapes(cfg: dict, logvar: bool = False) -> Dict[str, Tuple[int, ...]]:
    """Shapes of the reference ``SwinV2.state_dict()`` (models/swinv2.py:278-292)."""
    res, patch = _pair(cfg["img_resolution"]), _pair(cfg["patch_size"])
    gh, gw = res[0] // patch[0], res[1] // patch[1]
    d, h = cfg["dim"], cfg["heads"]
    dff = int(8 / 3.0 * d)
    pp = patch[0] * patch[1]
    s: Dict[str, Tuple[int, ...]] = {
        "pos_embed": (1, gh * gw, d),
        "patch_embed.emb.weight": (d, cfg["in_channels"] * pp),
        "patch_embed.emb.bias": (d,),
        "latent_embed.l1.weight": (d, d), "latent_embed.l1.bias": (d,),
        "latent_embed.l2.weight": (d, d), "latent_embed.l2.bias": (d,),
    }
    if logvar:
        s["logvar_embed.weight"] = (1, d)
        s["logvar_embed.bias"] = (1,)
    if cfg.get("auxiliary_dim", 0):
        s["auxiliary_embed.weight"] = (d, cfg["auxiliary_dim"])
        s["auxiliary_embed.bias"] = (d,)
    for i in range(cfg["depth"]):
        a, f = f"transformer.layers.{i}.0", f"transformer.layers.{i}.1"
        s[a + ".scale"] = (1, h, 1, 1)
        for blk in (a, f):
            s[blk + ".norm.norm.weight"] = (d,)
            s[blk + ".norm.norm.bias"] = (d,)
            s[blk + ".norm.modulation.weight"] = (2 * d, d)
            s[blk + ".norm.modulation.bias"] = (2 * d,)
        s[a + ".to_qkv.weight"] = (3 * d, d)
        s[a + ".wo.weight"] = (d, d)
        s[f + ".w1.weight"] = (2 * dff, d)
        s[f + ".w2.weight"] = (d, dff)
    s["head.head.0.weight"] = (cfg["out_channels"] * pp, d)
    return s


def random_state_dict(cfg: dict, seed: int = 1, bf16_weights: bool = True, logvar: bool = False,
                      prefix: str = "") -> Dict[str, torch.Tensor]:
    """A full, non-degenerate fp32 state dict on CPU, one seeded generator per tensor."""
    out: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(state_dict_shapes(cfg, logvar).items()):
        g = torch.Generator().manual_seed(seed * 1_000_003 + idx)
        z = torch.randn(shape, generator=g, dtype=torch.float32)
        if name.endswith("norm.norm.weight"):
            v = 1.0 + 0.1 * z
        elif name.endswith("norm.norm.bias"):
            v = 0.1 * z
        elif name.endswith(".scale"):
            v = math.log(10.0) + 0.3 * z          # reference init: ln 10 (models/swinv2.py:116)
        elif name.endswith("auxiliary_embed.weight"):
            v = 0.5 * z
        elif name.endswith(".bias"):
            v = 0.02 * z
        else:
            v = 0.02 * z                            # reference init scale (trunc_normal std 0.02)
        if bf16_weights and name.endswith(_GEMM_KEYS):
            v = v.to(torch.bfloat16).to(torch.float32)
        out[prefix + name] = v.contiguous()
    return out


def synthetic_fields(cfg: dict, batch: int, seed: int = 0, img_channels: int | None = None):
    """(latents [B,C_img,H,W], condition [B,C_in-C_img,H,W]) ~ N(0,1); ERA5 fields are standardised
    to about unit variance (data/era5.py:150-153)."""
    res = _pair(cfg["img_resolution"])
    c_img = cfg["out_channels"] if img_channels is None else img_channels
    c_cond = cfg["in_channels"] - c_img
    g = torch.Generator().manual_seed(10_007 * seed + 17)
    lat = torch.randn(batch, c_img, *res, generator=g, dtype=torch.float32)
    cond = torch.randn(batch, c_cond, *res, generator=g, dtype=torch.float32)
    return lat, cond


def synthetic_forcings(cfg: dict, steps: int, seed: int = 0, n_forcings: int = N_FORCINGS) -> torch.Tensor:
    """[steps, n_forcings, H, W] standardised forcings: channel 0 varies with the step (insolation-like),
    the others are static (orography / land-sea mask in the reference data)."""
    res = _pair(cfg["img_resolution"])
    g = torch.Generator().manual_seed(20_011 * seed + 5)
    static = torch.randn(1, n_forcings, *res, generator=g)
    f = static.repeat(steps, 1, 1, 1)
    f[:, 0] = torch.randn(steps, *res, generator=g)
    return f.contiguous()
