"""Drop-in ``SwinV2`` denoiser: the reference's constructor, ``forward`` signature and state-dict schema
(stockeh/swift ``src/swift/models/swinv2.py:254-330``), with the compute done by hand-written sm_100a kernels.

Selecting it: point hydra's ``model._target_`` at ``swift_b200.swinv2.SwinV2`` (the reference names
``swift.models.swinv2.SwinV2`` in ``configs/model/swinv2.yaml:1``); ``PassPrecond`` then instantiates it with the
same keyword arguments (``models/precond.py:123-131``) and ``load_state_dict(state["ema"])`` succeeds strictly.

The sub-modules below only *hold parameters* under the reference's names; none of them has a ``forward``.
``SwinV2.forward`` packs the parameters once per checkpoint (16-bit GEMM weights, reordered rows -- see
``packing.py``) and calls ``libswift_b200.so``.  There is no CPU path.  Three modes:
  * inference (``torch.no_grad()`` or ``.eval()``): the fused forecast kernels;
  * ``jvp=True``: the forward-mode tangent path of the sCM loss (``torch.func.jvp`` works through it);
  * ``.train()`` with grad enabled: the reverse-mode path of ``training.py`` -- ``F_x.backward(cot)`` /
    ``loss.backward()`` fill the parameters' ``.grad`` as the reference module does (training/loss.py:226-260).
"""
from __future__ import annotations

import math
from typing import Optional, Union

import torch
import torch.nn as nn

from . import packing
from .engine import Engine
from .training import DenoiserTrainFn, TrainEngine

try:
    # When the reference package is importable, BE a ``swift.models.swinv2.SwinV2``: ``train.py:271`` gates the optimiser's
    # parameter grouping on ``isinstance(net.model, SwinV2)`` (imported at train.py:21), so a replacement that is not a
    # subclass silently trains with different weight-decay / Muon groups.  Only the type is inherited: the reference's
    # constructor (which would build its nn.Modules) is bypassed, every method it defines is overridden below.
    from swift.models.abstract import AbstractNetwork as _Abstract  # type: ignore
    from swift.models.swinv2 import SwinV2 as _RefSwinV2  # type: ignore

    class _Base(_RefSwinV2):
        def __init__(self, img_resolution, in_channels: int, out_channels: int):
            _Abstract.__init__(self, img_resolution, in_channels, out_channels)
except Exception:  # pragma: no cover - the normal case on the GPU box
    class _Base(nn.Module):
        """Same attributes as swift.models.abstract.AbstractNetwork (models/abstract.py:12-35)."""

        def __init__(self, img_resolution, in_channels: int, out_channels: int):
            super().__init__()
            self.img_resolution = img_resolution
            self.in_channels = in_channels
            self.out_channels = out_channels


# ---------------------------------------------------------------------------- parameter containers
class _ModulatedNorm(nn.Module):            # models/swinv2.py:77-86
    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.norm = nn.LayerNorm(dim, eps)
        self.modulation = nn.Linear(dim, dim * 2, bias=True)


class _Attention(nn.Module):                # models/swinv2.py:105-116
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.norm = _ModulatedNorm(dim)
        self.to_qkv = nn.Linear(dim, dim * 3, bias=False)
        self.wo = nn.Linear(dim, dim, bias=False)
        self.scale = nn.Parameter(torch.log(10 * torch.ones(1, heads, 1, 1)))


class _FeedForward(nn.Module):              # models/swinv2.py:89-96
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.norm = _ModulatedNorm(dim)
        self.w1 = nn.Linear(dim, 2 * hidden, bias=False)
        self.w2 = nn.Linear(hidden, dim, bias=False)


class _Transformer(nn.Module):              # models/swinv2.py:142-172
    def __init__(self, depth: int, dim: int, heads: int):
        super().__init__()
        hidden = int(8 / 3.0 * dim)
        self.layers = nn.Sequential(*[nn.ModuleList([_Attention(dim, heads), _FeedForward(dim, hidden)])
                                      for _ in range(depth)])


class _PatchEmbedding(nn.Module):           # models/swinv2.py:217-222
    def __init__(self, in_features: int, dim: int):
        super().__init__()
        self.emb = nn.Linear(in_features, dim)


class _LatentEmbedding(nn.Module):          # models/swinv2.py:67-71
    def __init__(self, dim: int):
        super().__init__()
        self.l1 = nn.Linear(dim, dim, bias=True)
        self.l2 = nn.Linear(dim, dim, bias=True)


class _OutputHead(nn.Module):               # models/swinv2.py:233-244 (the Rearrange holds no parameters)
    def __init__(self, dim: int, out_features: int):
        super().__init__()
        self.head = nn.Sequential(nn.Linear(dim, out_features, bias=False))


# ---------------------------------------------------------------------------- forward-mode AD hook
class _TangentEval(torch.autograd.Function):
    """dF for given (x, t, aux, dx, dt).  A node of its own so that, when called from ``_DenoiserFn.jvp`` under a
    ``torch.func`` transform, its ``forward`` receives plain tensors (raw device pointers are needed for the C ABI)."""

    @staticmethod
    def forward(x, t, aux, dx, dt, module):
        return module.engine().forward_jvp(x, t, aux, dx, dt)[1]

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def jvp(ctx, *tangents):
        raise NotImplementedError("second-order forward-mode derivatives of swift_b200.SwinV2 are not implemented")

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError("swift_b200.SwinV2 has no reverse-mode backward (SURVEY.md section 8f-3)")


class _DenoiserFn(torch.autograd.Function):
    """The CUDA denoiser as an autograd node that supports forward-mode AD: ``torch.func.jvp`` through
    ``net(x, t, condition, auxiliary, jvp=True)`` (training/loss.py:216-225) lands in ``jvp`` below, which evaluates primal
    and tangent together with ``swb200_forward_jvp``.  No ``backward``: the tangent is used detached by the sCM loss."""

    @staticmethod
    def forward(x, t, aux, module):
        eng = module.engine()
        cond = eng.conditioning(t, aux)
        return eng.forward(x, None, cond[0], cond[1])

    @staticmethod
    def setup_context(ctx, inputs, output):
        x, t, aux, module = inputs
        ctx.module = module
        ctx.save_for_forward(x, t) if aux is None else ctx.save_for_forward(x, t, aux)
        ctx.has_aux = aux is not None

    @staticmethod
    def jvp(ctx, dx, dt, daux, _):
        saved = ctx.saved_tensors
        x, t = saved[0], saved[1]
        aux = saved[2] if ctx.has_aux else None
        dx = torch.zeros_like(x) if dx is None else dx.to(torch.float32).contiguous()
        dt = torch.zeros_like(t) if dt is None else dt.to(torch.float32).reshape(-1).expand_as(t).contiguous()
        if daux is not None and bool((daux != 0).any()):
            raise NotImplementedError("tangents with respect to `auxiliary` are not implemented")
        return _TangentEval.apply(x, t, aux, dx, dt, ctx.module)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError("swift_b200.SwinV2 has no reverse-mode backward (SURVEY.md section 8f-3)")


# ---------------------------------------------------------------------------- the module
class SwinV2(_Base):
    def __init__(self, img_resolution, in_channels: int, out_channels: int, window_size, shift_size, patch_size,
                 depth: int = 6, dim: int = 512, heads: int = 12, auxiliary_dim: int = 0, flash: bool = True,
                 logvar: bool = False, timestep_weight: float = 1.0):
        super().__init__(img_resolution, in_channels, out_channels)
        img, patch = packing._pair(img_resolution), packing._pair(patch_size)
        self.geometry = packing.Geometry(img=img, patch=patch, window=packing._pair(window_size),
                                         shift=packing._pair(shift_size), in_channels=in_channels,
                                         out_channels=out_channels, depth=depth, dim=dim, heads=heads,
                                         aux_dim=auxiliary_dim, timestep_weight=float(timestep_weight))
        packing.check_supported(self.geometry)      # fail at construction, not at the first forward
        gh, gw = self.geometry.grid
        self.auxiliary_dim = auxiliary_dim
        self.timestep_weight = timestep_weight
        self.flash = flash                           # accepted for signature parity; attention is always fused

        self.pos_embed = nn.Parameter(torch.randn(1, gh * gw, dim) * 0.02)
        self.patch_embed = _PatchEmbedding(in_channels * patch[0] * patch[1], dim)
        self.latent_embed = _LatentEmbedding(dim)
        self.logvar_embed = nn.Linear(dim, 1) if logvar else None
        self.auxiliary_embed = nn.Linear(auxiliary_dim, dim) if auxiliary_dim else None
        self.transformer = _Transformer(depth, dim, heads)
        self.head = _OutputHead(dim, out_channels * patch[0] * patch[1])
        self._init_weights()
        self._engine: Optional[Engine] = None
        self._engine_key = None
        self._train_engine: Optional[TrainEngine] = None
        self._train_engine_key = None
        self._grad_hook = None       # on_stage callback of TrainEngine.backward (training.GradientAllReduce.hook) or None
        self.split_embed = True      # [hi|lo] bf16 operands for the two small end GEMMs (accuracy, <1% of FLOPs)
        self.split_head = True
        self.act_fp16 = True         # tensor-core operands in fp16 (else bf16): 8x smaller rounding error, same tcgen05 rate
        self.x_single = True         # fp16 mode only: the forecast path's residual stream is ONE fp16 value per element (not a [hi | lo]
                                     # pair): 40 % less residual traffic, one extra 2^-11 rounding per update
        self.attn_fp16 = True        # bf16 mode only: q / k / v and P stay fp16 inside the attention kernel (bounded values)
        self.gemm_tile = 3           # 1: 128x176 single CTA, 2: 256x176 CTA pair, 3: 256x352 CTA pair
        self.attn_impl = 0           # 0: tcgen05 attention when the shift is a multiple of 8, 1: mma.sync kernel, 2: tcgen05
        self.fuse_ln = None          # LayerNorm + modulation + residual add in the GEMM epilogue: bit 0 = wo, bit 1 = w2; 0 = separate kernel;
                                     # None = automatic (effective_fuse_ln): 3 with the single-value stream (x travels by TMA), 2 with the
                                     # [hi | lo] pair (the short-K wo GEMM is epilogue-bound when fused: measured)
        self.max_chunk = 8           # samples pushed through the kernels per launch sequence

    def _init_weights(self):
        """Same policy as models/swinv2.py:295-303: trunc_normal(0.02), zeros for modulation/head, zero biases."""
        for name, m in self.named_modules():
            if isinstance(m, nn.Linear):
                if "modulation" in name or "head" in name:
                    nn.init.zeros_(m.weight)
                else:
                    nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    # ------------------------------------------------------------------ engine management
    def effective_fuse_ln(self) -> int:
        if self.fuse_ln is not None:
            return int(self.fuse_ln)
        return 3 if (self.act_fp16 and self.x_single) else 2

    def _params_key(self):
        return tuple((p.data_ptr(), p._version, p.device) for p in self.parameters())

    def engine(self) -> Engine:
        """The packed CUDA engine for the current parameter values (re-packed when parameters change)."""
        key = (self._params_key(), self.split_embed, self.split_head, self.max_chunk, self.act_fp16, self.gemm_tile, self.attn_impl,
               self.effective_fuse_ln(), self.attn_fp16, self.x_single)
        if self._engine is None or self._engine_key != key:
            dev = self.pos_embed.device
            if dev.type != "cuda":
                raise RuntimeError("swift_b200.SwinV2 runs on CUDA only: move the module to a B200 with .cuda(); "
                                   "there is no CPU fallback")
            sd = {k: v for k, v in self.state_dict().items()}
            self._engine = Engine(sd, self.geometry, dev, self.split_embed, self.split_head, self.max_chunk,
                                  self.act_fp16, self.gemm_tile, self.attn_impl, self.effective_fuse_ln(), self.attn_fp16,
                                  self.x_single)
            self._engine_key = key
        return self._engine

    def train_engine(self) -> TrainEngine:
        """The bf16 training engine (grad-enabled forward + backward) for the current parameter values."""
        key = (self._params_key(), self.gemm_tile, self.attn_impl)
        if self._train_engine is None or self._train_engine_key != key:
            dev = self.pos_embed.device
            if dev.type != "cuda":
                raise RuntimeError("swift_b200.SwinV2 runs on CUDA only: move the module to a B200 with .cuda(); "
                                   "there is no CPU fallback")
            old = self._train_engine
            self._train_engine = TrainEngine({k: v for k, v in self.state_dict().items()}, self.geometry, dev,
                                             self.gemm_tile, self.attn_impl)
            if old is not None:                  # keep the activation tape / workspace / gradient buffers across steps
                self._train_engine._tape, self._train_engine._ws = old._tape, old._ws
                self._train_engine._cond_scratch = old._cond_scratch
                self._train_engine.grads, self._train_engine._gstruct, self._train_engine._cstruct = (
                    old.grads, old._gstruct, old._cstruct)
            self._train_engine_key = key
        return self._train_engine

    # ------------------------------------------------------------------ reference-compatible forward
    def forward(self, x: torch.Tensor, t: torch.Tensor, auxiliary: Optional[torch.Tensor] = None, jvp: bool = False,
                return_logvar: bool = False) -> Union[torch.Tensor, tuple]:
        """models/swinv2.py:305-330.  x [B, in_channels, H, W]; t scalar, [1] or [B]; auxiliary [B, aux_dim] or None."""
        B = x.shape[0]
        if self.training and torch.is_grad_enabled() and not jvp:
            # reverse mode: the grad-enabled forward of training/loss.py:227 (bf16 operands, activation tape kept)
            if torch.is_autocast_enabled():
                x = x.float()
            x32 = x.to(torch.float32).contiguous()
            t32 = t.detach().to(device=x.device, dtype=torch.float32)
            if t32.dim() == 0 or (t32.dim() == 1 and t32.shape[0] == 1):
                t32 = t32.reshape(-1).repeat(B)
            aux32 = None
            if self.auxiliary_embed is not None and auxiliary is not None:
                aux32 = auxiliary.detach().to(device=x.device, dtype=torch.float32).reshape(-1, self.auxiliary_dim)
                if aux32.shape[0] == 1 and B > 1:
                    aux32 = aux32.expand(B, -1)
                aux32 = aux32.contiguous()
            names = tuple(n for n, _ in self.named_parameters())
            want_lv = self.logvar_embed is not None and return_logvar
            return DenoiserTrainFn.apply(x32, t32.reshape(B).contiguous(), aux32, self, names, want_lv, *self.parameters())
        if not jvp:                                                # (a detach would drop the forward-mode tangents)
            x, t = x.detach(), t.detach()
        x = x.to(torch.float32).contiguous()
        t = t.to(device=x.device, dtype=torch.float32)
        if t.dim() == 0 or (t.dim() == 1 and t.shape[0] == 1):     # models/swinv2.py:316-317
            t = t.reshape(-1).repeat(B)
        t = t.contiguous()
        aux = None
        if self.auxiliary_embed is not None and auxiliary is not None:
            aux = auxiliary.detach().to(device=x.device, dtype=torch.float32).reshape(-1, self.auxiliary_dim)
            if aux.shape[0] == 1 and B > 1:                        # broadcast like the reference's `t + aux_embed(.)`
                aux = aux.expand(B, -1)
            aux = aux.contiguous()
        if jvp:
            # forward-mode tangent path (the reference switches its attention to explicit softmax here, models/swinv2.py:
            # 129-134, so that torch.func.jvp can differentiate it): an autograd node with a jvp rule
            if return_logvar:
                raise NotImplementedError("return_logvar together with jvp=True is not implemented")
            return _DenoiserFn.apply(x, t, aux, self)        # (the engine is built inside the node, below any transform)
        eng = self.engine()
        want_lv = self.logvar_embed is not None and return_logvar
        cond = eng.conditioning(t, aux)
        lv = eng.logvar_head(self.logvar_embed.weight, self.logvar_embed.bias, B) if want_lv else None   # models/swinv2.py:326-328
        y = eng.forward(x, None, cond[0], cond[1])
        return (y, lv) if want_lv else y
