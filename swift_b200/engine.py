"""Host-side driver of the CUDA denoiser: owns packed weights and workspace, issues the C-ABI calls.

PyTorch is used for device memory and streams only; every FLOP of the forward runs in ``libswift_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import itertools
import math
from typing import Dict, Optional, Tuple

import torch

from . import _lib, packing


def _aligned_buffer(nbytes: int, device: torch.device, align: int = 1024) -> Tuple[torch.Tensor, int]:
    buf = torch.empty(nbytes + align, dtype=torch.uint8, device=device)
    base = (buf.data_ptr() + align - 1) // align * align
    return buf, base


_GENERATION = itertools.count(1)


class Engine:
    """One packed SwinV2 denoiser on one GPU.  ``generation`` is unique per Engine object of the process: whatever caches
    device pointers into an engine (captured CUDA graphs, conditioning vectors) keys on it, never on ``id()``."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], geom: packing.Geometry, device: torch.device,
                 split_embed: bool = True, split_head: bool = True, max_chunk: int = 8, act_fp16: bool = True,
                 gemm_tile: int = 3, attn_impl: int = 0, fuse_ln: int = 3, attn_fp16: bool = True,
                 x_single: bool = True):
        if device.type != "cuda":
            raise RuntimeError("swift_b200 runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.lib()
        self.geom = geom
        self.device = device
        self.generation = next(_GENERATION)
        with torch.cuda.device(device):
            self.model, self._keep = packing.pack(state_dict, geom, device, split_embed, split_head, act_fp16,
                                                   gemm_tile, attn_impl, fuse_ln, attn_fp16, x_single)
        self.fuse_ln = int(fuse_ln)
        self.act_fp16 = act_fp16
        _lib.check(self.lib.swb200_validate(C.byref(self.model)), "validate")
        self.max_chunk = max_chunk
        self._ws: Optional[torch.Tensor] = None
        self._ws_base = 0
        self._ws_bytes = 0
        self._cond_scratch: Optional[torch.Tensor] = None
        self.launches = 0           # kernels enqueued by this engine (for bench.py's gpu_launches)

    # ------------------------------------------------------------------ helpers
    @property
    def out_shape(self) -> Tuple[int, int, int]:
        return (self.geom.out_channels, self.geom.img[0], self.geom.img[1])

    def launches_per_forward(self, batch: int) -> int:
        chunks = math.ceil(batch / max(1, min(batch, self.max_chunk)))
        per_layer = 7 - bin(self.fuse_ln).count("1")     # qkv, attention, wo, w1, w2 + one LN kernel per un-fused projection
        return chunks * (2 + per_layer * self.geom.depth + 1)

    def workspace(self, batch: int) -> Tuple[int, int]:
        chunk = max(1, min(batch, self.max_chunk))
        need = self.lib.swb200_workspace_bytes(C.byref(self.model), chunk)
        if need == 0:
            _lib.check(self.lib.swb200_validate(C.byref(self.model)), "workspace_bytes")
        if self._ws is None or self._ws_bytes < need:
            self._ws, self._ws_base = _aligned_buffer(need, self.device)
            self._ws_bytes = need
        return self._ws_base, self._ws_bytes

    @staticmethod
    def _stream() -> int:
        return torch.cuda.current_stream().cuda_stream

    @staticmethod
    def _check_f32(t: torch.Tensor, name: str):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError(f"{name}: expected a contiguous float32 CUDA tensor, got {t.dtype} on {t.device} "
                               f"(contiguous={t.is_contiguous()}); swift_b200 has no CPU fallback")

    # ------------------------------------------------------------------ fp16 range diagnostics
    SATURATION_SLOTS = ("x_hi", "x_lo", "qkv", "attn", "branch", "h")

    def count_saturation(self, on: bool = True) -> None:
        """Switch the fp16 saturation counters of ``swb200_debug_saturation`` on (zeroed) or off.  Process-global."""
        if on:
            self._sat = torch.zeros(len(self.SATURATION_SLOTS), dtype=torch.int64, device=self.device)
            _lib.check(self.lib.swb200_debug_saturation(self._sat.data_ptr()), "debug_saturation")
        else:
            _lib.check(self.lib.swb200_debug_saturation(None), "debug_saturation")

    def saturation_counts(self) -> Dict[str, int]:
        """Elements found at +-65504 since ``count_saturation(True)``, per tensor class (synchronises)."""
        if getattr(self, "_sat", None) is None:
            raise RuntimeError("call count_saturation(True) first")
        return dict(zip(self.SATURATION_SLOTS, self._sat.tolist()))

    # ------------------------------------------------------------------ conditioning
    def conditioning(self, t: torch.Tensor, aux: Optional[torch.Tensor], want_cond: bool = False):
        """t [B] fp32, aux [B, aux_dim] fp32 or None -> (gain, bias) each [2*depth, B, dim] (+ cond [B, dim])."""
        g = self.geom
        B = t.shape[0]
        self._check_f32(t, "t")
        if aux is not None:
            self._check_f32(aux, "auxiliary")
            if aux.shape != (B, g.aux_dim):
                raise RuntimeError(f"auxiliary must be [{B}, {g.aux_dim}], got {tuple(aux.shape)}")
        gain = torch.empty(2 * g.depth, B, g.dim, device=self.device, dtype=torch.float32)
        bias = torch.empty_like(gain)
        cond = torch.empty(B, g.dim, device=self.device, dtype=torch.float32) if want_cond else None
        need = self.lib.swb200_conditioning_scratch_bytes(C.byref(self.model), B)
        if self._cond_scratch is None or self._cond_scratch.numel() < need:
            self._cond_scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.swb200_conditioning(C.byref(self.model), t.data_ptr(), _lib.ptr(aux), B, gain.data_ptr(),
                                                bias.data_ptr(), _lib.ptr(cond), self._cond_scratch.data_ptr(),
                                                self._cond_scratch.numel(), self._stream()), "conditioning")
        self.launches += 5
        return (gain, bias, cond) if want_cond else (gain, bias)

    def logvar_head(self, weight: torch.Tensor, bias: torch.Tensor, B: int) -> torch.Tensor:
        """logvar [B] = logvar_embed(c) (models/swinv2.py:326-327) for the batch of the last ``conditioning`` call (its scratch
        holds the conditioning vector c); weight [1, dim] / bias [1] fp32 on this device."""
        lv = torch.empty(B, device=self.device, dtype=torch.float32)
        w = weight.detach().to(torch.float32).reshape(-1).contiguous()
        b = bias.detach().to(torch.float32).reshape(-1).contiguous()
        _lib.check(self.lib.swb200_logvar_head(C.byref(self.model), self._cond_scratch.data_ptr(), w.data_ptr(), b.data_ptr(), B,
                                               lv.data_ptr(), self._stream()), "logvar_head")
        self.launches += 1
        return lv

    # ------------------------------------------------------------------ forward
    def forward(self, x0: torch.Tensor, x1: Optional[torch.Tensor], gain: torch.Tensor, bias: torch.Tensor,
                out: Optional[torch.Tensor] = None, scale0: float = 1.0, xt: Optional[torch.Tensor] = None,
                fprev: Optional[torch.Tensor] = None, out_f: Optional[torch.Tensor] = None, alpha: float = 0.0,
                beta: float = 1.0, gamma: float = 0.0, rollout: Optional["RolloutGlue"] = None) -> Optional[torch.Tensor]:
        """F = SwinV2(cat([x0*scale0, x1], 1));  returns  y = alpha*xt + beta*F + gamma*fprev  (NCHW fp32).

        With ``rollout`` the per-step glue of generate.py:120-131 is applied to y in the head epilogue (state updated
        in place, physical state written to ``rollout.phys``) and y itself is not materialised (returns None)."""
        g = self.geom
        self._check_f32(x0, "x0")
        B, c0 = x0.shape[0], x0.shape[1]
        c1 = 0
        if x1 is not None:
            self._check_f32(x1, "x1")
            c1 = x1.shape[1]
            if x1.shape[0] != B or tuple(x1.shape[2:]) != g.img:
                raise RuntimeError(f"condition shape {tuple(x1.shape)} does not match {B} x * x {g.img}")
        if tuple(x0.shape[2:]) != g.img or c0 + c1 != g.in_channels:
            raise RuntimeError(f"input channels {c0}+{c1} / resolution {tuple(x0.shape[2:])} do not match the model "
                               f"({g.in_channels} channels at {g.img})")
        if gain.shape != (2 * g.depth, B, g.dim):
            raise RuntimeError("conditioning vectors were computed for a different batch size")
        for nm, tns in (("xt", xt), ("fprev", fprev), ("out_f", out_f), ("out", out)):
            if tns is not None:
                self._check_f32(tns, nm)
                if tuple(tns.shape) != (B, *self.out_shape):
                    raise RuntimeError(f"{nm} must be {(B, *self.out_shape)}, got {tuple(tns.shape)}")
        upd = _lib.Update(_lib.ptr(xt), _lib.ptr(fprev), _lib.ptr(out_f), alpha, beta, gamma)
        if rollout is not None:
            rollout.check(self, B)
            upd.state, upd.state_channels = rollout.state.data_ptr(), rollout.state.shape[1]
            upd.zero_channel = rollout.zero_channel
            upd.x_std, upd.x_mean, upd.d_std = (rollout.x_std.data_ptr(), rollout.x_mean.data_ptr(),
                                                rollout.d_std.data_ptr())
            upd.phys = _lib.ptr(rollout.phys)
        elif out is None:
            out = torch.empty(B, *self.out_shape, device=self.device, dtype=torch.float32)
        ws, ws_bytes = self.workspace(B)
        _lib.check(self.lib.swb200_forward(C.byref(self.model), x0.data_ptr(), c0, scale0, _lib.ptr(x1), c1, B,
                                           gain.data_ptr(), bias.data_ptr(), C.byref(upd), _lib.ptr(out), ws,
                                           ws_bytes, self._stream()), "forward")
        self.launches += self.launches_per_forward(B)
        return out


    # ------------------------------------------------------------------ forward-mode tangent (sCM training loss)
    def forward_jvp(self, x: torch.Tensor, t: torch.Tensor, aux: Optional[torch.Tensor], dx: torch.Tensor,
                    dt: torch.Tensor, cond: Optional[torch.Tensor] = None, scale0: float = 1.0
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(F, dF) = jvp(SwinV2.forward, (x, t), (dx, dt)): x, dx [B, in_channels, H, W]; t, dt [B].
        With ``cond`` [B, c1, H, W] the network input is cat([x * scale0, cond]) and its tangent cat([dx * scale0, 0])
        (x, dx [B, in_channels - c1, H, W]): the concat of models/precond.py:143-145 happens in the patch gather."""
        g = self.geom
        for nm, v in (("x", x), ("dx", dx), ("t", t), ("dt", dt)) + ((("cond", cond),) if cond is not None else ()):
            self._check_f32(v, nm)
        B = x.shape[0]
        c1 = 0 if cond is None else cond.shape[1]
        c0 = g.in_channels - c1
        if (tuple(x.shape[1:]) != (c0, *g.img) or dx.shape != x.shape or t.shape != (B,) or dt.shape != (B,)
                or (cond is not None and (cond.shape[0] != B or tuple(cond.shape[2:]) != tuple(g.img)))):
            raise RuntimeError(f"forward_jvp: x/dx must be [B, {c0}, {g.img[0]}, {g.img[1]}], cond [B, {c1}, ...] and "
                               f"t/dt [B]")
        if aux is not None:
            self._check_f32(aux, "auxiliary")
        vecs = [torch.empty(2 * g.depth, B, g.dim, device=self.device, dtype=torch.float32) for _ in range(4)]
        need = self.lib.swb200_conditioning_jvp_scratch_bytes(C.byref(self.model), B)
        scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.swb200_conditioning_jvp(C.byref(self.model), t.data_ptr(), dt.data_ptr(), _lib.ptr(aux), B,
                                                    *[v.data_ptr() for v in vecs], scratch.data_ptr(), need,
                                                    self._stream()), "conditioning_jvp")
        need = self.lib.swb200_jvp_workspace_bytes(C.byref(self.model))
        if getattr(self, "_jvp_ws", None) is None or self._jvp_ws[2] < need:
            buf, base = _aligned_buffer(need, self.device)
            self._jvp_ws = (buf, base, need)
        y = torch.empty(B, *self.out_shape, device=self.device, dtype=torch.float32)
        dy = torch.empty_like(y)
        _lib.check(self.lib.swb200_forward_jvp(C.byref(self.model), x.data_ptr(), c0, float(scale0), _lib.ptr(cond), c1,
                                               dx.data_ptr(), B, *[v.data_ptr() for v in vecs], y.data_ptr(),
                                               dy.data_ptr(), self._jvp_ws[1], self._jvp_ws[2], self._stream()),
                   "forward_jvp")
        return y, dy


class RolloutGlue:
    """Buffers of the fused rollout epilogue: ``state`` [B, C_state, H, W] (standardised condition buffer, first
    out_channels channels updated in place), per-channel normalisers [C] and the optional physical output."""

    def __init__(self, state: torch.Tensor, x_std: torch.Tensor, x_mean: torch.Tensor, d_std: torch.Tensor,
                 phys: Optional[torch.Tensor] = None, zero_channel: int = -1):
        self.state, self.phys, self.zero_channel = state, phys, int(zero_channel)
        self.x_std, self.x_mean, self.d_std = (t.reshape(-1).to(torch.float32).contiguous() for t in (x_std, x_mean, d_std))

    def check(self, eng: "Engine", B: int) -> None:
        C_out = eng.geom.out_channels
        Engine._check_f32(self.state, "rollout state")
        if self.state.shape[0] != B or self.state.shape[1] < C_out or tuple(self.state.shape[2:]) != eng.geom.img:
            raise RuntimeError(f"rollout state {tuple(self.state.shape)} does not match batch {B} / model output")
        for nm in ("x_std", "x_mean", "d_std"):
            t = getattr(self, nm)
            Engine._check_f32(t, nm)
            if t.numel() != C_out:
                raise RuntimeError(f"{nm} must have {C_out} entries")
        if self.phys is not None:
            Engine._check_f32(self.phys, "phys")
            if tuple(self.phys.shape) != (B, *eng.out_shape):
                raise RuntimeError(f"phys must be {(B, *eng.out_shape)}")
