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


def _layer_names(l: int):
    a, f = f"transformer.layers.{l}.0", f"transformer.layers.{l}.1"
    return {"scale": a + ".scale", "attn_ln_w": a + ".norm.norm.weight", "attn_ln_b": a + ".norm.norm.bias",
            "attn_mod_w": a + ".norm.modulation.weight", "attn_mod_b": a + ".norm.modulation.bias",
            "to_qkv": a + ".to_qkv.weight", "wo": a + ".wo.weight", "ff_ln_w": f + ".norm.norm.weight",
            "ff_ln_b": f + ".norm.norm.bias", "ff_mod_w": f + ".norm.modulation.weight",
            "ff_mod_b": f + ".norm.modulation.bias", "w1": f + ".w1.weight", "w2": f + ".w2.weight"}


def ref_params(sd: Dict[str, torch.Tensor], g: Geometry, device: torch.device):
    """``struct swb200_ref_params`` for a reference-schema state dict: (struct, fp32 device tensors + pointer arrays that
    must stay alive until the packing kernels have run)."""
    import ctypes as C
    f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
    keep = []

    def ptr(key):
        t = f32(sd[key])
        keep.append(t)
        return t.data_ptr()

    r = _lib.RefParams()
    r.pos_embed, r.patch_w, r.patch_b = ptr("pos_embed"), ptr("patch_embed.emb.weight"), ptr("patch_embed.emb.bias")
    has_aux = bool(g.aux_dim) and "auxiliary_embed.weight" in sd
    if has_aux:
        r.aux_w, r.aux_b = ptr("auxiliary_embed.weight"), ptr("auxiliary_embed.bias")
    r.l1_w, r.l1_b = ptr("latent_embed.l1.weight"), ptr("latent_embed.l1.bias")
    r.l2_w, r.l2_b = ptr("latent_embed.l2.weight"), ptr("latent_embed.l2.bias")
    r.head_w = ptr("head.head.0.weight")
    names = [_layer_names(l) for l in range(g.depth)]
    for field in names[0]:
        arr = (C.c_void_p * g.depth)(*[ptr(names[l][field]) for l in range(g.depth)])
        keep.append(arr)
        setattr(r, field, arr)
    return r, keep, has_aux


def pack(sd: Dict[str, torch.Tensor], g: Geometry, device: torch.device, split_embed: bool = True,
         split_head: bool = True, act_fp16: bool = True, gemm_tile: int = 3, attn_impl: int = 0, fuse_ln: int = 3,
         attn_fp16: bool = True, x_single: bool = True):
    """Returns (``_lib.Model`` struct, dict of device tensors that must stay alive as long as the struct is used).
    The conversions themselves are ``swb200_pack_weights`` (csrc/pack.cu): this function only collects the parameter
    pointers of the state dict and owns the packed buffer."""
    import ctypes as C
    check_supported(g)
    D, L, Dff = g.dim, g.depth, g.dff
    if gemm_tile == 3 and Dff % (2 * HEAD_DIM):
        gemm_tile = 2                               # 352-wide tiles need whole gate/up slot pairs
    sd = dict(sd)
    assert tuple(sd["patch_embed.emb.weight"].shape) == (D, g.pp * g.in_channels), sd["patch_embed.emb.weight"].shape
    assert tuple(sd["head.head.0.weight"].shape) == (g.out_channels * g.pp, D)
    ref, keep_src, has_aux = ref_params(sd, g, device)
    m = _lib.Model()
    m.img_h, m.img_w = g.img
    m.patch_h, m.patch_w = g.patch
    m.win_h, m.win_w = g.window
    m.shift_h, m.shift_w = g.shift
    m.in_channels, m.out_channels, m.depth, m.dim, m.heads = g.in_channels, g.out_channels, L, D, g.heads
    m.dff, m.aux_dim, m.k_embed = Dff, (g.aux_dim if has_aux else 0), g.k_embed
    m.split_embed, m.split_head = int(split_embed), int(split_head)
    m.act_fp16 = int(act_fp16)
    m.gemm_tile = int(gemm_tile)
    m.attn_impl = int(attn_impl)
    m.fuse_ln = int(fuse_ln)
    m.attn_fp16 = int(attn_fp16)
    m.x_single = int(bool(x_single) and bool(act_fp16))
    m.timestep_weight = float(g.timestep_weight)
    lib = _lib.lib()
    nbytes = lib.swb200_packed_bytes(C.byref(m))
    if nbytes == 0:
        _lib.check(lib.swb200_validate(C.byref(m)), "packed_bytes")
    buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
    base = (buf.data_ptr() + 255) // 256 * 256
    _lib.check(lib.swb200_pack_weights(C.byref(m), C.byref(ref), base, nbytes, torch.cuda.current_stream(device).cuda_stream),
               "pack_weights")
    keep: Dict[str, object] = {"packed": buf, "_sources": keep_src}
    bf = torch.float16 if act_fp16 else torch.bfloat16

    def view(ptr_value, shape, dtype):              # typed views into the packed buffer (tools / benches read them)
        off = ptr_value - buf.data_ptr()
        n = int(torch.tensor(shape).prod().item()) * torch.empty((), dtype=dtype).element_size()
        return buf[off:off + n].view(dtype).reshape(shape)

    keep["w_qkv"] = view(m.w_qkv, (L, 3 * D, D), bf)
    keep["w_o"] = view(m.w_o, (L, D, D), bf)
    keep["w_1"] = view(m.w_1, (L, 2 * Dff, D), bf)
    keep["w_2"] = view(m.w_2, (L, D, Dff), bf)
    keep["qscale"] = view(m.qscale, (L, g.heads), torch.float32)
    keep["pos_embed"] = view(m.pos_embed, (g.tokens, D), torch.float32)
    return m, keep


def pack_train(sd: Dict[str, torch.Tensor], g: Geometry, device: torch.device, gemm_tile: int = 3, attn_impl: int = 0):
    """``struct swb200_train_model`` for the grad-enabled forward / backward (include/swift_b200.h): bf16 operands; the
    four per-layer matrices as plain bf16 copies in the REFERENCE's row order plus their transposes (dgrad operands);
    embed / head packs and the fp32 conditioning parameters as in ``pack``.  The conversions are
    ``swb200_pack_train_weights`` (csrc/pack.cu).  Returns (struct, tensors to keep alive)."""
    import ctypes as C
    check_supported(g)
    ref, keep_src, has_aux = ref_params(dict(sd), g, device)
    tm = _lib.TrainModel()
    m = tm.base
    m.img_h, m.img_w = g.img
    m.patch_h, m.patch_w = g.patch
    m.win_h, m.win_w = g.window
    m.shift_h, m.shift_w = g.shift
    m.in_channels, m.out_channels, m.depth, m.dim, m.heads = g.in_channels, g.out_channels, g.depth, g.dim, g.heads
    m.dff, m.aux_dim, m.k_embed = g.dff, (g.aux_dim if has_aux else 0), g.k_embed
    m.split_embed, m.split_head = 1, 1
    m.act_fp16, m.attn_fp16, m.fuse_ln = 0, 0, 0
    m.gemm_tile = int(gemm_tile)
    m.attn_impl = int(attn_impl)
    m.timestep_weight = float(g.timestep_weight)
    lib = _lib.lib()
    nbytes = lib.swb200_train_packed_bytes(C.byref(tm))
    if nbytes == 0:
        _lib.check(lib.swb200_validate(C.byref(m)), "train_packed_bytes")
    buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
    base = (buf.data_ptr() + 255) // 256 * 256
    _lib.check(lib.swb200_pack_train_weights(C.byref(tm), C.byref(ref), base, nbytes,
                                             torch.cuda.current_stream(device).cuda_stream), "pack_train_weights")
    return tm, {"packed": buf, "_sources": keep_src}
