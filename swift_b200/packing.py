"""Pack a reference-schema SwinV2 state dict into the device layouts the CUDA kernels consume.

Input keys are the reference's ``SwinV2.state_dict()`` names (models/swinv2.py:278-292).
Everything here is one-time, per-checkpoint host plumbing (PyTorch index ops); the layouts are the contract of
``struct swb200_model`` in include/swift_b200.h.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch

from . import _lib

HEAD_DIM = 88          # Swift-B head dim the qkv / attention kernels are specialised for
TILE_N = 176           # GEMM N tile = 2 heads = [88 gate | 88 up]


def _pair(v) -> Tuple[int, int]:
    if isinstance(v, int):
        return (v, v)
    if isinstance(v, (list, tuple)) and len(v) == 2:
        return (int(v[0]), int(v[1]))
    raise TypeError(f"Invalid type {type(v)}")     # same failure mode as models/abstract.py:54


@dataclass
class Geometry:
    img: Tuple[int, int]
    patch: Tuple[int, int]
    window: Tuple[int, int]
    shift: Tuple[int, int]
    in_channels: int
    out_channels: int
    depth: int
    dim: int
    heads: int
    aux_dim: int
    timestep_weight: float

    @property
    def grid(self) -> Tuple[int, int]:
        return (self.img[0] // self.patch[0], self.img[1] // self.patch[1])

    @property
    def tokens(self) -> int:
        return self.grid[0] * self.grid[1]

    @property
    def dff(self) -> int:
        return int(8 / 3.0 * self.dim)      # models/swinv2.py:160

    @property
    def pp(self) -> int:
        return self.patch[0] * self.patch[1]

    @property
    def k_embed(self) -> int:
        return (self.in_channels * self.pp + 7) // 8 * 8


def check_supported(g: Geometry) -> None:
    """Raise ``NotImplementedError`` (loudly, no fallback) for shapes the sm_100a kernels do not cover."""
    if g.window != (16, 16):
        raise NotImplementedError(f"swift_b200 implements 16x16 windows only (got {g.window})")
    if g.dim != g.heads * HEAD_DIM:
        raise NotImplementedError(f"swift_b200 implements head_dim {HEAD_DIM} only (dim={g.dim}, heads={g.heads})")
    if g.dff % HEAD_DIM:
        raise NotImplementedError(f"mlp dim int(8/3*dim)={g.dff} must be a multiple of {HEAD_DIM} "
                                  f"(dim a multiple of 33)")
    if g.grid[0] % 16 or g.grid[1] % 16 or g.img[0] % g.patch[0] or g.img[1] % g.patch[1]:
        raise NotImplementedError(f"token grid {g.grid} must be a multiple of the 16x16 window")


def pack(sd: Dict[str, torch.Tensor], g: Geometry, device: torch.device, split_embed: bool = True,
         split_head: bool = True, act_fp16: bool = True, gemm_tile: int = 3, attn_impl: int = 0, fuse_ln: int = 2,
         attn_fp16: bool = True):
    """Returns (``_lib.Model`` struct, dict of device tensors that must stay alive as long as the struct is used)."""
    check_supported(g)
    bf = torch.float16 if act_fp16 else torch.bfloat16     # one 16-bit operand format for activations and weights
    D, H, L, Dff, pp = g.dim, g.heads, g.depth, g.dff, g.pp
    f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
    keep: Dict[str, torch.Tensor] = {}
    if gemm_tile == 3 and Dff % (2 * HEAD_DIM):
        gemm_tile = 2                               # 352-wide tiles need whole gate/up slot pairs
    half = HEAD_DIM * (2 if gemm_tile == 3 else 1)

    # patch-embed: reference feature order "(p1 p2 c)" -> ours "(c p1 p2)"; zero pad to k_embed; duplicate for [hi|lo]
    w = f32(sd["patch_embed.emb.weight"])
    C_in = g.in_channels
    assert w.shape == (D, pp * C_in), w.shape
    w = w.reshape(D, pp, C_in).permute(0, 2, 1).reshape(D, C_in * pp)
    w = torch.nn.functional.pad(w, (0, g.k_embed - C_in * pp))
    if split_embed:
        w = torch.cat([w, w], dim=1)
    keep["w_embed"] = w.to(bf).contiguous()
    # the patch-embed bias is folded into the position table: x = A W^T + (pos + bias) costs one operand in the epilogue
    keep["pos_embed"] = (f32(sd["pos_embed"]).reshape(g.tokens, D) + f32(sd["patch_embed.emb.bias"])[None, :]).contiguous()

    if g.aux_dim and "auxiliary_embed.weight" in sd:
        keep["aux_w"] = f32(sd["auxiliary_embed.weight"])
        keep["aux_b"] = f32(sd["auxiliary_embed.bias"])
    for n in ("l1", "l2"):
        keep[f"{n}_w"] = f32(sd[f"latent_embed.{n}.weight"])
        keep[f"{n}_b"] = f32(sd[f"latent_embed.{n}.bias"])

    mod_w, mod_b, gam, bet, qs, wq, wo, w1, w2 = [], [], [], [], [], [], [], [], []
    for l in range(L):
        a, f = f"transformer.layers.{l}.0", f"transformer.layers.{l}.1"
        for blk in (a, f):
            mod_w.append(f32(sd[blk + ".norm.modulation.weight"]))
            mod_b.append(f32(sd[blk + ".norm.modulation.bias"]))
            gam.append(f32(sd[blk + ".norm.norm.weight"]))
            bet.append(f32(sd[blk + ".norm.norm.bias"]))
        # exp(clamp(scale, max=ln 100)) (models/swinv2.py:125-126)
        qs.append(torch.clamp(f32(sd[a + ".scale"]).reshape(H), max=math.log(1.0 / 0.01)).exp())
        # to_qkv rows are h*3hd + part*hd + d (rearrange then chunk, models/swinv2.py:120-121) -> part*D + h*hd + d
        q = f32(sd[a + ".to_qkv.weight"]).reshape(H, 3, HEAD_DIM, D).permute(1, 0, 2, 3).reshape(3 * D, D)
        wq.append(q.to(bf))
        wo.append(f32(sd[a + ".wo.weight"]).to(bf))
        # w1 rows: [gate(Dff) | up(Dff)] (chunk(2), models/swinv2.py:99) -> per GEMM tile of 2*half rows:
        # [half gate rows | half up rows] (half = 88 for the 176-wide tiles, 176 for the 352-wide tile)
        w1_ = f32(sd[f + ".w1.weight"])
        gate, up = w1_[:Dff].reshape(Dff // half, 1, half, D), w1_[Dff:].reshape(Dff // half, 1, half, D)
        w1.append(torch.cat([gate, up], dim=1).reshape(2 * Dff, D).to(bf))
        w2.append(f32(sd[f + ".w2.weight"]).to(bf))
    keep["mod_w"] = torch.cat(mod_w, 0).contiguous()
    keep["mod_b"] = torch.cat(mod_b, 0).contiguous()
    keep["ln_gamma"] = torch.stack(gam, 0).contiguous()
    keep["ln_beta"] = torch.stack(bet, 0).contiguous()
    keep["qscale"] = torch.stack(qs, 0).contiguous()
    keep["w_qkv"] = torch.stack(wq, 0).contiguous()
    keep["w_o"] = torch.stack(wo, 0).contiguous()
    keep["w_1"] = torch.stack(w1, 0).contiguous()
    keep["w_2"] = torch.stack(w2, 0).contiguous()
    wh = f32(sd["head.head.0.weight"])
    assert wh.shape == (g.out_channels * pp, D)
    if split_head:
        wh = torch.cat([wh, wh], dim=1)
    keep["w_head"] = wh.to(bf).contiguous()

    m = _lib.Model()
    m.img_h, m.img_w = g.img
    m.patch_h, m.patch_w = g.patch
    m.win_h, m.win_w = g.window
    m.shift_h, m.shift_w = g.shift
    m.in_channels, m.out_channels, m.depth, m.dim, m.heads = g.in_channels, g.out_channels, L, D, H
    m.dff, m.aux_dim, m.k_embed = Dff, (g.aux_dim if "aux_w" in keep else 0), g.k_embed
    m.split_embed, m.split_head = int(split_embed), int(split_head)
    m.act_fp16 = int(act_fp16)
    m.gemm_tile = int(gemm_tile)
    m.attn_impl = int(attn_impl)
    m.fuse_ln = int(fuse_ln)
    m.attn_fp16 = int(attn_fp16)
    m.timestep_weight = float(g.timestep_weight)
    for name in ("w_embed", "b_embed", "pos_embed", "aux_w", "aux_b", "l1_w", "l1_b", "l2_w", "l2_b", "mod_w",
                 "mod_b", "ln_gamma", "ln_beta", "qscale", "w_qkv", "w_o", "w_1", "w_2", "w_head"):
        setattr(m, name, keep[name].data_ptr() if name in keep else None)
    return m, keep


def pack_train(sd: Dict[str, torch.Tensor], g: Geometry, device: torch.device, gemm_tile: int = 3, attn_impl: int = 0):
    """``struct swb200_train_model`` for the grad-enabled forward / backward (include/swift_b200.h): bf16 operands; the
    four per-layer matrices as plain bf16 copies in the REFERENCE's row order plus their transposes (dgrad operands);
    embed / head packs and the fp32 conditioning parameters as in ``pack``.  Returns (struct, tensors to keep alive)."""
    check_supported(g)
    bf = torch.bfloat16
    D, H, L, Dff, pp = g.dim, g.heads, g.depth, g.dff, g.pp
    f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
    keep: Dict[str, torch.Tensor] = {}
    w = f32(sd["patch_embed.emb.weight"]).reshape(D, pp, g.in_channels).permute(0, 2, 1).reshape(D, g.in_channels * pp)
    w = torch.nn.functional.pad(w, (0, g.k_embed - g.in_channels * pp))
    keep["w_embed"] = torch.cat([w, w], dim=1).to(bf).contiguous()
    keep["pos_embed"] = (f32(sd["pos_embed"]).reshape(g.tokens, D) + f32(sd["patch_embed.emb.bias"])[None, :]).contiguous()
    if g.aux_dim and "auxiliary_embed.weight" in sd:
        keep["aux_w"] = f32(sd["auxiliary_embed.weight"])
        keep["aux_b"] = f32(sd["auxiliary_embed.bias"])
    for n in ("l1", "l2"):
        keep[f"{n}_w"] = f32(sd[f"latent_embed.{n}.weight"])
        keep[f"{n}_b"] = f32(sd[f"latent_embed.{n}.bias"])
    names = {"mod_w": [], "mod_b": [], "ln_gamma": [], "ln_beta": [], "qscale": [], "t_qkv": [], "t_o": [], "t_1": [], "t_2": []}
    for l in range(L):
        a, f = f"transformer.layers.{l}.0", f"transformer.layers.{l}.1"
        for blk in (a, f):
            names["mod_w"].append(f32(sd[blk + ".norm.modulation.weight"]))
            names["mod_b"].append(f32(sd[blk + ".norm.modulation.bias"]))
            names["ln_gamma"].append(f32(sd[blk + ".norm.norm.weight"]))
            names["ln_beta"].append(f32(sd[blk + ".norm.norm.bias"]))
        names["qscale"].append(torch.clamp(f32(sd[a + ".scale"]).reshape(H), max=math.log(1.0 / 0.01)).exp())
        names["t_qkv"].append(f32(sd[a + ".to_qkv.weight"]).to(bf))
        names["t_o"].append(f32(sd[a + ".wo.weight"]).to(bf))
        names["t_1"].append(f32(sd[f + ".w1.weight"]).to(bf))
        names["t_2"].append(f32(sd[f + ".w2.weight"]).to(bf))
    keep["mod_w"] = torch.cat(names["mod_w"], 0).contiguous()
    keep["mod_b"] = torch.cat(names["mod_b"], 0).contiguous()
    for n in ("ln_gamma", "ln_beta", "qscale"):
        keep[n] = torch.stack(names[n], 0).contiguous()
    for n in ("qkv", "o", "1", "2"):
        keep["t_" + n] = torch.stack(names["t_" + n], 0).contiguous()                       # [L, N, K]
        keep["tt_" + n] = keep["t_" + n].transpose(1, 2).contiguous()                       # [L, K, N]
    wh = f32(sd["head.head.0.weight"])
    nh = g.out_channels * pp
    kp = (nh + 7) // 8 * 8
    keep["w_head"] = torch.cat([wh, wh], dim=1).to(bf).contiguous()
    keep["tt_head"] = torch.nn.functional.pad(wh.t(), (0, kp - nh)).to(bf).contiguous()    # [D, kp]

    tm = _lib.TrainModel()
    m = tm.base
    m.img_h, m.img_w = g.img
    m.patch_h, m.patch_w = g.patch
    m.win_h, m.win_w = g.window
    m.shift_h, m.shift_w = g.shift
    m.in_channels, m.out_channels, m.depth, m.dim, m.heads = g.in_channels, g.out_channels, L, D, H
    m.dff, m.aux_dim, m.k_embed = Dff, (g.aux_dim if "aux_w" in keep else 0), g.k_embed
    m.split_embed, m.split_head = 1, 1
    m.act_fp16, m.attn_fp16, m.fuse_ln = 0, 0, 0
    m.gemm_tile = int(gemm_tile)
    m.attn_impl = int(attn_impl)
    m.timestep_weight = float(g.timestep_weight)
    for name in ("w_embed", "pos_embed", "aux_w", "aux_b", "l1_w", "l1_b", "l2_w", "l2_b", "mod_w", "mod_b", "ln_gamma",
                 "ln_beta", "qscale", "w_head"):
        setattr(m, name, keep[name].data_ptr() if name in keep else None)
    for n in ("qkv", "o", "1", "2"):
        setattr(tm, "w_" + n, keep["t_" + n].data_ptr())
        setattr(tm, "wt_" + n, keep["tt_" + n].data_ptr())
    tm.wt_head = keep["tt_head"].data_ptr()
    tm.kp_head = kp
    return tm, keep
