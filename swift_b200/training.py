"""Reverse mode of the CUDA denoiser: the grad-enabled forward and ``F_x.backward(cot)`` of the sCM training step
(stockeh/swift ``training/loss.py:226-260``: ``F_x = net(x_t / sigma_d, t, condition, auxiliary)`` under autograd;
``training/trainer.py:199-219``: ``loss.backward()`` then the optimiser).

``TrainEngine`` drives the C ABI of ``include/swift_b200.h`` ("reverse mode" section): one ``swb200_train_forward`` that
writes the activation tape, then ``swb200_train_backward_head`` / ``_layer`` (depth-1 .. 0) / ``_embed`` and
``swb200_conditioning_backward``.  Gradients land in flat fp32 buffers (``TrainEngine.grads``) laid out per parameter
family; ``parameter_gradients`` maps them onto the reference's parameter names.  The per-layer granularity exists so
that a data-parallel caller can all-reduce layer l's gradients while layer l-1 is still being differentiated
(``GradientAllReduce``; the reference wraps the net in DDP, ``trainer.py:76-84``).

The training path computes with bf16 tensor-core operands and fp32 accumulation, residual stream and gradients -- the
reference trains under bf16 autocast.  PyTorch is used for memory, streams and ``torch.distributed`` only.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib, packing
from .engine import _aligned_buffer


class TrainEngine:
    """bf16 training twin of ``engine.Engine`` for one set of parameter values (re-built when they change)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], geom: packing.Geometry, device: torch.device,
                 gemm_tile: int = 3, attn_impl: int = 0):
        if device.type != "cuda":
            raise RuntimeError("swift_b200 training runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.lib()
        self.geom, self.device = geom, device
        if gemm_tile == 3 and geom.dff % (2 * packing.HEAD_DIM):
            gemm_tile = 2
        with torch.cuda.device(device):
            self.model, self._keep = packing.pack_train(state_dict, geom, device, gemm_tile, attn_impl)
        # logvar head (models/swinv2.py:281, :326-327): logvar = logvar_embed(c) on the conditioning vector; fp32 copies
        self.lv_w = self.lv_b = None
        if "logvar_embed.weight" in state_dict:
            self.lv_w = state_dict["logvar_embed.weight"].detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()
            self.lv_b = state_dict["logvar_embed.bias"].detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        self._cond = None            # (B, gain, bias, t, aux) of the last conditioning()
        self.logvar: Optional[torch.Tensor] = None   # [B] of the last conditioning() when the model has the head
        self._tape = None            # (buffer, base, bytes, B)
        self._ws = None
        self._cond_scratch: Optional[torch.Tensor] = None
        self._ctx = None
        self.grads: Optional[Dict[str, torch.Tensor]] = None
        self._gstruct = None
        self._cstruct = None

    # ------------------------------------------------------------------ buffers
    @staticmethod
    def _stream() -> int:
        return torch.cuda.current_stream().cuda_stream

    def _buffers(self, B: int):
        tp = C.byref(self.model)
        need_t = self.lib.swb200_train_tape_bytes(tp, B)
        need_w = self.lib.swb200_train_workspace_bytes(tp, B)
        if need_t == 0 or need_w == 0:
            _lib.check(-2, "train buffers (unsupported configuration)")
        if self._tape is None or self._tape[2] < need_t:
            buf, base = _aligned_buffer(need_t, self.device)
            self._tape = (buf, base, need_t)
        if self._ws is None or self._ws[2] < need_w:
            buf, base = _aligned_buffer(need_w, self.device)
            self._ws = (buf, base, need_w)
        return self._tape, self._ws

    def gradient_buffers(self, B: int) -> Dict[str, torch.Tensor]:
        """Flat fp32 gradient buffers (zero-initialised once; the C calls overwrite or accumulate)."""
        g = self.geom
        D, L, H, Dff = g.dim, g.depth, g.heads, g.dff
        nh = g.out_channels * g.pp
        if self.grads is not None and self.grads["dgain"].shape[1] == B:
            return self.grads
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        gr = {"w_qkv": z(L, 3 * D, D), "w_o": z(L, D, D), "w_1": z(L, 2 * Dff, D), "w_2": z(L, D, Dff), "w_head": z(nh, D),
              "w_embed_t": z(g.k_embed, D), "b_embed": z(D), "pos_embed": z(g.tokens, D), "dscale": z(L, H),
              "dgain": z(2 * L, B, D), "dbias": z(2 * L, B, D),
              "l1_w": z(D, D), "l1_b": z(D), "l2_w": z(D, D), "l2_b": z(D), "mod_w": z(2 * L * 2 * D, D), "mod_b": z(2 * L * 2 * D),
              "ln_gamma": z(2 * L, D), "ln_beta": z(2 * L, D)}
        if self.model.base.aux_dim:
            gr["aux_w"], gr["aux_b"] = z(D, g.aux_dim), z(D)
        if self.lv_w is not None:
            gr["lv_w"], gr["lv_b"] = z(D), z(1)
        self.grads = gr
        gs = _lib.TrainGrads()
        for n in ("w_qkv", "w_o", "w_1", "w_2", "w_head", "w_embed_t", "b_embed", "pos_embed", "dscale", "dgain", "dbias"):
            setattr(gs, n, gr[n].data_ptr())
        cs = _lib.CondGrads()
        for n in ("aux_w", "aux_b", "l1_w", "l1_b", "l2_w", "l2_b", "mod_w", "mod_b", "ln_gamma", "ln_beta"):
            setattr(cs, n, gr[n].data_ptr() if n in gr else None)
        self._gstruct, self._cstruct = gs, cs
        return gr

    # ------------------------------------------------------------------ forward
    def conditioning(self, t: torch.Tensor, aux: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """The conditioning stage of the grad-enabled forward (per-sample gain / bias of every ModulatedNorm; activations kept
        for the backward).  With a logvar head, returns (and keeps in ``self.logvar``) logvar [B] = logvar_embed(c)."""
        g = self.geom
        B = t.shape[0]
        bp = C.byref(self.model.base)
        gain = torch.empty(2 * g.depth, B, g.dim, device=self.device, dtype=torch.float32)
        bias = torch.empty_like(gain)
        need = self.lib.swb200_conditioning_scratch_bytes(bp, B)
        if self._cond_scratch is None or self._cond_scratch.numel() < need:
            self._cond_scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.swb200_conditioning(bp, t.data_ptr(), _lib.ptr(aux), B, gain.data_ptr(), bias.data_ptr(), None,
                                                self._cond_scratch.data_ptr(), self._cond_scratch.numel(), self._stream()),
                   "conditioning")
        self._cond = (B, gain, bias)
        self.logvar = None
        if self.lv_w is not None:
            self.logvar = torch.empty(B, device=self.device, dtype=torch.float32)
            _lib.check(self.lib.swb200_logvar_head(bp, self._cond_scratch.data_ptr(), self.lv_w.data_ptr(), self.lv_b.data_ptr(), B,
                                                   self.logvar.data_ptr(), self._stream()), "logvar_head")
        return self.logvar

    def forward(self, x0: torch.Tensor, x1: Optional[torch.Tensor], t: torch.Tensor, aux: Optional[torch.Tensor],
                scale0: float = 1.0, reuse_conditioning: bool = False) -> torch.Tensor:
        """F = SwinV2(cat([x0 * scale0, x1], 1), t, aux) with the activation tape kept for ``backward``.
        ``reuse_conditioning``: ``conditioning(t, aux)`` has just been called for this batch (logvar models need its output
        before the loss target can be finished)."""
        g = self.geom
        for nm, v in (("x0", x0), ("t", t)) + ((("x1", x1),) if x1 is not None else ()) + ((("aux", aux),) if aux is not None else ()):
            if not (v.is_cuda and v.dtype == torch.float32 and v.is_contiguous()):
                raise RuntimeError(f"{nm}: expected a contiguous float32 CUDA tensor (swift_b200 has no CPU fallback)")
        B, c0 = x0.shape[0], x0.shape[1]
        c1 = 0 if x1 is None else x1.shape[1]
        if c0 + c1 != g.in_channels or tuple(x0.shape[2:]) != g.img or t.shape != (B,):
            raise RuntimeError(f"train forward: inputs {tuple(x0.shape)} (+{c1} channels), t {tuple(t.shape)} do not match the model")
        tp = C.byref(self.model)
        if not (reuse_conditioning and self._cond is not None and self._cond[0] == B):
            self.conditioning(t, aux)
        _, gain, bias = self._cond
        tape, ws = self._buffers(B)
        y = torch.empty(B, g.out_channels, *g.img, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.swb200_train_forward(tp, x0.data_ptr(), c0, float(scale0), _lib.ptr(x1), c1, B, gain.data_ptr(),
                                                 bias.data_ptr(), y.data_ptr(), tape[1], tape[2], ws[1], ws[2], self._stream()),
                   "train_forward")
        self._ctx = (B, gain, aux)
        return y

    # ------------------------------------------------------------------ backward
    def backward(self, cot: torch.Tensor, accumulate: bool = False,
                 on_stage: Optional[Callable[[str, int], None]] = None, cond_exchange=None,
                 dlogvar: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Parameter gradients of ``sum(F * cot)`` for the last ``forward``.  ``on_stage(kind, layer)`` is called after
        the kernels of a stage have been enqueued ("head", "layer" l = depth-1 .. 0, "embed", "cond"): the hook for
        overlapping the gradient all-reduce of finished stages with the rest of the backward."""
        if self._ctx is None:
            raise RuntimeError("backward() without a preceding forward()")
        B, gain, aux = self._ctx
        g = self.geom
        if not (cot.is_cuda and cot.dtype == torch.float32 and cot.is_contiguous()) or \
                tuple(cot.shape) != (B, g.out_channels, *g.img):
            raise RuntimeError(f"cot must be a contiguous float32 CUDA tensor {(B, g.out_channels, *g.img)}")
        self.gradient_buffers(B)
        self._gstruct.accumulate = int(bool(accumulate))
        tp, gp = C.byref(self.model), C.byref(self._gstruct)
        tape, ws = self._buffers(B)
        st = self._stream()
        if not accumulate:
            self.grads["dgain"].zero_()
            self.grads["dbias"].zero_()
        _lib.check(self.lib.swb200_train_backward_head(tp, B, cot.data_ptr(), tape[1], ws[1], ws[2], gp, st), "backward_head")
        if on_stage:
            on_stage("head", -1)
        for l in range(g.depth - 1, -1, -1):
            _lib.check(self.lib.swb200_train_backward_layer(tp, l, B, gain.data_ptr(), tape[1], ws[1], ws[2], gp, st),
                       f"backward_layer {l}")
            if on_stage:
                on_stage("layer", l)
        _lib.check(self.lib.swb200_train_backward_embed(tp, B, tape[1], ws[1], ws[2], gp, st), "backward_embed")
        if on_stage:
            on_stage("embed", -1)
        bp = C.byref(self.model.base)
        dgain, dbias, fwd_scratch, Bc, kind = self.grads["dgain"], self.grads["dbias"], self._cond_scratch, B, "cond"
        if (dlogvar is not None) != (self.lv_w is not None) and dlogvar is not None:
            raise RuntimeError("dlogvar given for a model without a logvar head")
        if dlogvar is not None:
            dlogvar = dlogvar.detach().to(device=self.device, dtype=torch.float32).reshape(B).contiguous()
        if cond_exchange is not None and dlogvar is not None:
            dgain, dbias, fwd_scratch, aux, Bc, dlogvar = cond_exchange(self, dgain, dbias, fwd_scratch, aux, B, dlogvar)
            kind = "cond_replicated"
        elif cond_exchange is not None:
            # data parallel: instead of all-reducing the conditioning gradients (the 24 modulation Linears alone are 214 MB
            # of fp32), every rank receives every rank's per-sample INPUTS of this stage (a few hundred KB) and evaluates the
            # whole global batch -- the gradients are sums of per-sample outer products, so all ranks get the same, already
            # averaged result with no collective on the gradients themselves
            dgain, dbias, fwd_scratch, aux, Bc = cond_exchange(self, dgain, dbias, fwd_scratch, aux, B)
            kind = "cond_replicated"
        need = self.lib.swb200_conditioning_backward_scratch_bytes(bp, Bc)
        scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        if dlogvar is not None:
            # the head's gradient enters the conditioning vector before the latent MLP is differentiated (swinv2.py:326-327)
            _lib.check(self.lib.swb200_conditioning_backward_logvar(
                bp, _lib.ptr(aux), Bc, fwd_scratch.data_ptr(), dgain.data_ptr(), dbias.data_ptr(), C.byref(self._cstruct),
                int(bool(accumulate)), scratch.data_ptr(), need, self.lv_w.data_ptr(), dlogvar.data_ptr(),
                self.grads["lv_w"].data_ptr(), self.grads["lv_b"].data_ptr(), st), "conditioning_backward_logvar")
        else:
            if self.lv_w is not None and not accumulate:
                self.grads["lv_w"].zero_()
                self.grads["lv_b"].zero_()
            _lib.check(self.lib.swb200_conditioning_backward(bp, _lib.ptr(aux), Bc, fwd_scratch.data_ptr(), dgain.data_ptr(),
                                                             dbias.data_ptr(), C.byref(self._cstruct), int(bool(accumulate)),
                                                             scratch.data_ptr(), need, st), "conditioning_backward")
        if on_stage:
            on_stage(kind, -1)
        return self.grads

    # ------------------------------------------------------------------ flat buffers -> reference parameter names
    def parameter_gradients(self, scale_params: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Gradients keyed by the reference's ``SwinV2.state_dict()`` names (views of the flat buffers where the layouts
        agree).  ``scale_params``: the ``transformer.layers.{l}.0.scale`` parameters (needed for d exp(min(scale, ln 100)))."""
        g, gr = self.geom, self.grads
        D, L, pp, C_in = g.dim, g.depth, g.pp, g.in_channels
        out: Dict[str, torch.Tensor] = {}
        for l in range(L):
            a, f = f"transformer.layers.{l}.0", f"transformer.layers.{l}.1"
            out[a + ".to_qkv.weight"] = gr["w_qkv"][l]
            out[a + ".wo.weight"] = gr["w_o"][l]
            out[f + ".w1.weight"] = gr["w_1"][l]
            out[f + ".w2.weight"] = gr["w_2"][l]
            sc = scale_params[a + ".scale"].detach().to(torch.float32)
            lim = math.log(1.0 / 0.01)
            # s = exp(clamp(scale, max=ln 100)) (models/swinv2.py:125-126): ds/dscale = s below the clamp, 0 above it
            s = torch.clamp(sc, max=lim).exp() * (sc <= lim)
            out[a + ".scale"] = gr["dscale"][l].reshape(sc.shape) * s
            for k, blk in ((2 * l, a), (2 * l + 1, f)):
                out[blk + ".norm.norm.weight"] = gr["ln_gamma"][k]
                out[blk + ".norm.norm.bias"] = gr["ln_beta"][k]
                out[blk + ".norm.modulation.weight"] = gr["mod_w"][k * 2 * D:(k + 1) * 2 * D]
                out[blk + ".norm.modulation.bias"] = gr["mod_b"][k * 2 * D:(k + 1) * 2 * D]
        out["head.head.0.weight"] = gr["w_head"]
        we = gr["w_embed_t"][: C_in * pp].t()                                 # [D, (c p1 p2)]
        out["patch_embed.emb.weight"] = we.reshape(D, C_in, pp).permute(0, 2, 1).reshape(D, pp * C_in)   # -> (p1 p2 c)
        out["patch_embed.emb.bias"] = gr["b_embed"]
        out["pos_embed"] = gr["pos_embed"].reshape(1, g.tokens, D)
        for n in ("l1", "l2"):
            out[f"latent_embed.{n}.weight"] = gr[f"{n}_w"]
            out[f"latent_embed.{n}.bias"] = gr[f"{n}_b"]
        if "aux_w" in gr:
            out["auxiliary_embed.weight"] = gr["aux_w"]
            out["auxiliary_embed.bias"] = gr["aux_b"]
        if "lv_w" in gr:
            out["logvar_embed.weight"] = gr["lv_w"].reshape(1, D)
            out["logvar_embed.bias"] = gr["lv_b"]
        return out

    def stage_buffers(self, kind: str, layer: int) -> List[torch.Tensor]:
        """The gradient buffers a backward stage has completed (what a data-parallel caller reduces after it)."""
        gr = self.grads
        if kind == "head":
            return [gr["w_head"]]
        if kind == "layer":
            return [gr["w_2"][layer], gr["w_1"][layer], gr["w_o"][layer], gr["w_qkv"][layer], gr["dscale"][layer]]
        if kind == "embed":
            return [gr["w_embed_t"], gr["b_embed"], gr["pos_embed"]]
        if kind == "cond_replicated":           # evaluated on the gathered global batch: already identical on every rank
            return []
        return [gr[n] for n in ("l1_w", "l1_b", "l2_w", "l2_b", "mod_w", "mod_b", "ln_gamma", "ln_beta", "aux_w", "aux_b", "lv_w",
                                "lv_b") if n in gr]


class GradientAllReduce:
    """Data-parallel gradient averaging overlapped with the backward (the reference: DDP, ``trainer.py:76-84``).

    ``hook`` is passed as ``on_stage`` to ``TrainEngine.backward``: after every stage an event is recorded on the compute
    stream and the stage's gradient buffers are all-reduced (NCCL, SUM then scaled by 1/world) on a side stream, so layer
    l's 75 MB travel over NVLink while layer l-1 is being differentiated.  ``finish()`` makes the compute stream wait for
    the last collective.  With a gloo group (CPU tests) the same sequence runs synchronously."""

    def __init__(self, module, group=None, enabled: bool = True):
        """``module``: the ``swift_b200.SwinV2`` whose (current) training engine owns the gradient buffers."""
        import torch.distributed as dist
        self.dist, self.module, self.group = dist, module, group
        self.world = dist.get_world_size(group) if (dist.is_initialized() and enabled) else 1
        dev = next(module.parameters()).device
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.bytes = 0
        self._events: List[Tuple[torch.cuda.Event, torch.cuda.Event]] = []

    def hook(self, kind: str, layer: int) -> None:
        if self.world == 1:
            return
        bufs = self.module._train_engine.stage_buffers(kind, layer)
        if self.stream is None:                      # host tensors (gloo, CPU tests of the bookkeeping): synchronous
            for b in bufs:
                self.dist.all_reduce(b, op=self.dist.ReduceOp.SUM, group=self.group)
                b.mul_(1.0 / self.world)
                self.bytes += b.numel() * 4
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for b in bufs:                            # NCCL averages inside the collective: no extra kernel per buffer
                self.dist.all_reduce(b, op=self.dist.ReduceOp.AVG, group=self.group)
                self.bytes += b.numel() * 4
            e1.record(self.stream)
            self._events.append((e0, e1))

    def exchange_conditioning(self, engine: "TrainEngine", dgain, dbias, fwd_scratch, aux, B: int, dlogvar=None):
        """``cond_exchange`` of ``TrainEngine.backward``: one ``all_gather`` of every rank's conditioning-stage inputs (the
        per-sample gain / bias gradients, the embedding / MLP activations and modulation vectors ``swb200_conditioning``
        left in its scratch, the auxiliary input) -> the same tensors for the global batch of ``world * B`` samples, the
        gradients pre-scaled by 1 / world so that the stage's output is the data-parallel MEAN."""
        if self.world == 1:
            return (dgain, dbias, fwd_scratch, aux, B) if dlogvar is None else (dgain, dbias, fwd_scratch, aux, B, dlogvar)
        g = engine.geom
        D, L2 = g.dim, 2 * g.depth
        n_sc = B * (3 * D + L2 * 2 * D)
        sc = fwd_scratch.view(torch.float32)[:n_sc]
        parts = [dgain.reshape(-1), dbias.reshape(-1), sc]
        if aux is not None:
            parts.append(aux.reshape(-1))
        if dlogvar is not None:
            parts.append(dlogvar.reshape(-1))
        mine = torch.cat(parts).contiguous()
        flat = torch.empty(self.world * mine.numel(), dtype=torch.float32, device=mine.device)
        self.dist.all_gather_into_tensor(flat, mine, group=self.group)      # (a flat output: what both NCCL and gloo accept)
        allr = flat.view(self.world, mine.numel())
        self.bytes += allr.numel() * 4
        W, o = self.world, 0
        n_g = L2 * B * D

        def take(n):
            nonlocal o
            v = allr[:, o:o + n]
            o += n
            return v

        dg = take(n_g).reshape(W, L2, B, D).permute(1, 0, 2, 3).reshape(L2, W * B, D).mul(1.0 / W).contiguous()
        db = take(n_g).reshape(W, L2, B, D).permute(1, 0, 2, 3).reshape(L2, W * B, D).mul(1.0 / W).contiguous()
        scr = take(n_sc)
        emb = scr[:, :B * D].reshape(W * B, D)
        h1 = scr[:, B * D:2 * B * D].reshape(W * B, D)
        cv = scr[:, 2 * B * D:3 * B * D].reshape(W * B, D)
        mod = scr[:, 3 * B * D:].reshape(W * B, L2 * 2 * D)
        scratch_all = torch.cat([emb.reshape(-1), h1.reshape(-1), cv.reshape(-1), mod.reshape(-1)]).contiguous()
        aux_all = take(aux.numel()).reshape(W * B, -1).contiguous() if aux is not None else None
        if dlogvar is not None:                                     # a gradient like dgain / dbias: mean over the ranks
            return dg, db, scratch_all, aux_all, W * B, take(B).reshape(W * B).mul(1.0 / W).contiguous()
        return dg, db, scratch_all, aux_all, W * B

    def finish(self) -> None:
        if self.world > 1 and self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)

    def comm_ms(self) -> float:
        """Sum of the collectives' durations on the side stream (call after a synchronize)."""
        ms = sum(a.elapsed_time(b) for a, b in self._events)
        self._events.clear()
        return ms


# ---------------------------------------------------------------------------------------------- autograd bridge
class DenoiserTrainFn(torch.autograd.Function):
    """``SwinV2.forward`` under autograd: forward = ``TrainEngine.forward``, backward = ``TrainEngine.backward`` returning
    one gradient per parameter, so ``F_x.backward(cot)`` / ``loss.backward()`` fills ``.grad`` exactly as the reference
    module does (training/loss.py:226-260).  Inputs (x, t, auxiliary) get no gradient: the sCM loss does not need one."""

    @staticmethod
    def forward(ctx, x, t, aux, module, names, want_logvar, *params):
        eng = module.train_engine()
        ctx.module, ctx.names, ctx.eng, ctx.want_logvar = module, names, eng, want_logvar
        if x.requires_grad:
            raise NotImplementedError("swift_b200.SwinV2 computes parameter gradients only (no gradient w.r.t. the input)")
        y = eng.forward(x, None, t, aux)
        if want_logvar:                          # (F_x, logvar) as models/swinv2.py:326-328
            return y, eng.logvar.clone()
        return y

    @staticmethod
    def backward(ctx, cot, dlogvar=None):
        eng = ctx.eng
        hook = ctx.module._grad_hook
        if ctx.want_logvar and dlogvar is None:
            dlogvar = torch.zeros_like(eng.logvar)
        eng.backward(cot.to(torch.float32).contiguous(), accumulate=False, on_stage=hook,
                     dlogvar=dlogvar if ctx.want_logvar else None)
        params = dict(ctx.module.named_parameters())
        by_name = eng.parameter_gradients({n: p for n, p in params.items() if n.endswith(".scale")})
        grads = []
        for n in ctx.names:                      # clones: autograd may keep what it is handed, the flat buffers are re-used
            gname = by_name.get(n)
            grads.append(None if gname is None else gname.reshape(params[n].shape).clone())
        return (None, None, None, None, None, None, *grads)


# ---------------------------------------------------------------------------------------------- the training step
def scm_train_step(net, x: torch.Tensor, t: torch.Tensor, z: torch.Tensor, step: int, condition: Optional[torch.Tensor] = None,
                   auxiliary=None, reducer: Optional[GradientAllReduce] = None, accumulate: bool = False,
                   timers: Optional[dict] = None, **loss_kwargs) -> Dict[str, torch.Tensor]:
    """Forward + tangent + backward of one sCM training step without autograd in the loop (what ``Trainer._forward_step`` +
    ``loss.backward()`` do, trainer.py:189-219): ``scm_target.scm_output_cotangent`` (loss, cot from one stacked primal +
    tangent pass), ``TrainEngine.forward`` (the concat with the condition and the 1/sigma_d scaling fused in the patch
    gather) and ``TrainEngine.backward`` with ``reducer.hook`` after every stage, so the data-parallel all-reduce of a
    finished layer overlaps the backward of the next.  Afterwards every parameter's ``.grad`` is a VIEW of the engine's
    flat gradient buffers (valid until the next backward; run the optimiser before it).  Returns the cotangent dict."""
    from .precond import process_auxiliary
    from .scm_target import scm_output_cotangent
    inner = getattr(net, "module", net)
    model = inner.model
    has_lv = model.logvar_embed is not None

    def mark(name):                       # optional phase timing (bench.py --mode train): CUDA events on the launch stream
        if timers is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            timers.setdefault(name, []).append(e)

    mark("t0")
    model.engine()                        # (re-)pack the 16-bit weights of the tangent path for the current parameters
    mark("pack_tangent")
    B, dev = x.shape[0], x.device
    aux = process_auxiliary(auxiliary, inner.auxiliary_dim, B, dev)
    if aux is not None:
        aux = aux.to(torch.float32).expand(B, -1).contiguous()
    t1 = t.to(device=dev, dtype=torch.float32).reshape(B).contiguous()
    # logvar head: the loss weights sample b by exp(-logvar_b) (loss.py:252-258), and logvar depends on (t, aux) only: the
    # conditioning stage of the grad-enabled forward runs first and hands it to the loss target
    logvar = model.train_engine().conditioning(t1, aux) if has_lv else None
    out = scm_output_cotangent(net, x, t, z, step, condition=condition, auxiliary=auxiliary, logvar=logvar, **loss_kwargs)
    mark("tangent_loss")
    eng = model.train_engine()
    mark("pack_train")
    cond = None
    if condition is not None and inner.condition_channels > 0:
        cond = condition.to(device=dev, dtype=torch.float32).contiguous()
    eng.forward(out["x_t"], cond, t1, aux, scale0=1.0 / float(inner.sigma_data), reuse_conditioning=has_lv)
    mark("train_forward")
    eng.backward(out["cot"], accumulate=accumulate, on_stage=reducer.hook if reducer is not None else None,
                 cond_exchange=reducer.exchange_conditioning if reducer is not None else None,
                 dlogvar=out.get("dlogvar"))
    mark("backward")
    if reducer is not None:
        reducer.finish()
    params = dict(model.named_parameters())
    grads = eng.parameter_gradients({n: p for n, p in params.items() if n.endswith(".scale")})
    for n, p in params.items():
        gview = grads.get(n)
        p.grad = None if gview is None else gview.reshape(p.shape)
    return out
