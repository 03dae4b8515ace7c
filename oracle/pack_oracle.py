"""Layout oracle of ``swb200_pack_weights``: the packed device layouts of ``struct swb200_model`` (include/swift_b200.h,
DESIGN.md section 3) restated with PyTorch index ops.

TEST INFRASTRUCTURE ONLY.  ``tests/test_host_cpu.py`` proves on the CPU that a forward evaluated FROM these packed tensors
reproduces the oracle network (i.e. the layouts mean what the header says), ``tests/test_gpu_kernels.py`` checks that the CUDA
packer writes exactly these bytes.  Reference lines: models/swinv2.py:99 (w1 = [gate | up]), :120-121 (to_qkv row order),
:125-126 (logit scale), :226 / :242 (patch / head feature orders).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

HEAD_DIM = 88


def pack_layouts(sd: Dict[str, torch.Tensor], g, device: torch.device, split_embed: bool = True, split_head: bool = True,
                 act_fp16: bool = True, gemm_tile: int = 3) -> Dict[str, torch.Tensor]:
    """The packed tensors of ``struct swb200_model`` by field name (w_embed, pos_embed, aux_w, ..., w_head), built with
    PyTorch index ops from a reference-schema state dict."""
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

    return keep
