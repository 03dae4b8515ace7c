"""The forward half of the sCM training step on the CUDA path (BASELINE.json configs[4]).

``SCMLoss.forward`` (stockeh/swift ``training/loss.py:192-260``) needs, per batch:

    x_t, dx_t/dt                                   elementwise                       :199-211
    (F, dF) = jvp(net, (x_t/sigma_d, t), (v_x, v_t))   the network + its tangent     :213-221
    F_x = net(x_t/sigma_d, t, ...)                 a SECOND, grad-enabled forward    :227
    g   = normalise(-cos^2 (sigma_d F - dx_t/dt) - r (cos sin x_t + sigma_d dF))     :233-248
    L   = mean sum w (F_x - sg(F_x) - g)^2                                           :253-260

Because F_x - sg(F_x) == 0, the loss value and its gradient with respect to the network output depend on g only:
``dL/dF_x = -2 w g / (B H W)`` (times exp(-logvar_b) for a model with a logvar head, which also receives dL/dlogvar).  ``scm_output_cotangent`` returns exactly that, from ONE stacked primal + tangent pass of
the CUDA engine (``swb200_forward_jvp``; the reference runs the network twice here).  ``scm_backward`` then runs the
grad-enabled forward and ``F_x.backward(cot)`` through the reverse-mode path of ``training.py``.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import _lib
from .precond import process_auxiliary

# training/loss.py:10-25, :36-49
_PRESSURE_LEVEL_VARS = ("geopotential", "u_component_of_wind", "v_component_of_wind", "vertical_velocity", "wind_speed",
                        "temperature", "relative_humidity", "specific_humidity", "vorticity", "potential_vorticity")
_PRESSURE_LEVELS = (50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000)
_SINGLE_LEVEL = {"2m_temperature": 1.0, "sea_surface_temperature": 0.1, "10m_u_component_of_wind": 0.1,
                 "10m_v_component_of_wind": 0.1, "mean_sea_level_pressure": 0.1}


def latitude_weights(n_lat: int, device=None) -> torch.Tensor:
    """training/loss.py:28-32 -> [1, 1, n_lat, 1]."""
    w = torch.cos(torch.deg2rad(torch.linspace(-90, 90, n_lat)))
    return torch.clamp(w / w.mean(), min=0.1).view(1, 1, -1, 1).to(device)


def variable_weights(variables: Sequence[str], device=None) -> torch.Tensor:
    """training/loss.py:35-57 -> [1, C, 1, 1], sum 1.  Unknown variable names raise KeyError as in the reference."""
    total = float(sum(_PRESSURE_LEVELS))
    w = []
    for v in variables:
        if v in _SINGLE_LEVEL:
            w.append(_SINGLE_LEVEL[v])
        else:
            base, _, lv = v.rpartition("_")
            if base not in _PRESSURE_LEVEL_VARS or not lv.isdigit() or int(lv) not in _PRESSURE_LEVELS:
                raise KeyError(v)
            w.append(int(lv) / total)
    w = torch.tensor(w, dtype=torch.float32).view(1, -1, 1, 1)
    return (w / w.sum()).to(device)


@torch.no_grad()
def scm_output_cotangent(net, x: torch.Tensor, t: torch.Tensor, z: torch.Tensor, step: int,
                         condition: Optional[torch.Tensor] = None, auxiliary=None, tangent_warmup_kimg: int = 0,
                         w_lat: Optional[torch.Tensor] = None, w_var: Optional[torch.Tensor] = None,
                         logvar=None, net_pretrained=None) -> Dict[str, torch.Tensor]:
    """x [B, C, H, W] targets, t = atan(tau / sigma_d) ([B] or [B,1,1,1]), z = sigma_d * N(0,1) like x (the draws of
    loss.py:196-200, made by the caller), ``net`` a PassPrecond around ``swift_b200.swinv2.SwinV2`` on a CUDA device.
    Returns {"loss", "cot" (= dL/dF_x), "g", "F", "dF", "x_t"}; all detached fp32.

    Three C-ABI calls: ``swb200_scm_noised_inputs`` (x_t, dx_t/dt, v_x, v_t), ``swb200_forward_jvp`` (the concat with the
    condition and the 1/sigma_d scaling happen in its patch gather) and ``swb200_scm_tangent_target`` (g, its per-sample
    normalisation, cot and the loss).

    Models with a logvar head (``logvar: true``; loss.py:227-232, :252-258) need ``logvar``: a [B] tensor, or a callable
    ``x_t -> [B] tensor`` evaluated once the noised inputs exist (the grad-enabled forward that produces it takes x_t).  The
    sample's squared term is then weighted by exp(-logvar_b), ``+ logvar_b`` is added, and the dict also carries
    ``"dlogvar"`` = dL/dlogvar [B] (``swb200_scm_tangent_target_logvar``).

    Distillation (loss.py:205-210; ``SCMLoss(distillation=True)`` + ``Trainer(net_pretrained=...)``): ``net_pretrained`` is the
    frozen v-prediction teacher, called once as ``net_pretrained(x_t / sigma_d, t, condition, auxiliary)`` under ``no_grad``
    (any module that returns a CUDA tensor: e.g. a second ``PassPrecond`` around ``swift_b200.SwinV2`` in eval mode, which then
    runs the forecast kernels); ``swb200_scm_distill_direction`` puts sigma_d * F_teacher in the place of cos z - sin x."""
    inner = getattr(net, "module", net)
    model = inner.model
    if not hasattr(model, "engine"):
        raise TypeError("scm_output_cotangent needs swift_b200.swinv2.SwinV2 as net.model (there is no fallback path)")
    has_lv = getattr(model, "logvar_embed", None) is not None
    if has_lv and logvar is None:
        # a silently different objective would be worse than an error
        raise RuntimeError("this model has a logvar head: scm_output_cotangent needs logvar (a [B] tensor or a callable of x_t)")
    if not has_lv and logvar is not None:
        raise RuntimeError("logvar given for a model without a logvar head")
    eng = model.engine()
    lib, dev = eng.lib, x.device
    stream = torch.cuda.current_stream().cuda_stream
    sd = float(inner.sigma_data)
    B, C, H, W = x.shape
    x = x.to(torch.float32).contiguous()
    z = z.to(device=dev, dtype=torch.float32).contiguous()
    t1 = t.to(device=dev, dtype=torch.float32).reshape(B).contiguous()
    if z.shape != x.shape:
        raise RuntimeError(f"z must have the shape of x {tuple(x.shape)}, got {tuple(z.shape)}")
    x_t, dxt, v_x = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    v_t = torch.empty_like(t1)
    _lib.check(lib.swb200_scm_noised_inputs(x.data_ptr(), z.data_ptr(), t1.data_ptr(), B, C, H, W, x_t.data_ptr(),
                                            dxt.data_ptr(), v_x.data_ptr(), v_t.data_ptr(), stream), "scm_noised_inputs")
    if net_pretrained is not None:
        Ft = net_pretrained(x_t / sd, t1, condition, auxiliary)
        Ft = Ft.detach().to(device=dev, dtype=torch.float32).contiguous()
        if Ft.shape != x.shape:
            raise RuntimeError(f"net_pretrained returned {tuple(Ft.shape)}, expected {tuple(x.shape)}")
        _lib.check(lib.swb200_scm_distill_direction(Ft.data_ptr(), t1.data_ptr(), sd, B, C, H, W, dxt.data_ptr(), v_x.data_ptr(),
                                                    stream), "scm_distill_direction")
    cond = None
    if condition is not None and inner.condition_channels > 0:          # precond.py:143-145; no tangent in the condition
        cond = condition.to(device=dev, dtype=torch.float32).contiguous()
    aux = process_auxiliary(auxiliary, inner.auxiliary_dim, B, dev)
    if aux is not None:
        aux = aux.to(torch.float32).expand(B, -1).contiguous()
    F, dF = eng.forward_jvp(x_t, t1, aux, v_x, v_t, cond=cond, scale0=1.0 / sd)
    r = min(1.0, step / (tangent_warmup_kimg * 1000)) if tangent_warmup_kimg > 0 else 1.0             # :233-237
    wv = None if w_var is None else w_var.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
    wl = None if w_lat is None else w_lat.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
    if (wv is not None and wv.numel() != C) or (wl is not None and wl.numel() != H):
        raise RuntimeError(f"w_var must have {C} entries and w_lat {H}")
    g, cot = torch.empty_like(x), torch.empty_like(x)
    loss = torch.empty((), device=dev, dtype=torch.float32)
    need = lib.swb200_scm_target_scratch_bytes(B)
    scratch = torch.empty(need // 8, dtype=torch.float64, device=dev)
    if has_lv:
        lv = logvar(x_t) if callable(logvar) else logvar
        lv = lv.detach().to(device=dev, dtype=torch.float32).reshape(B).contiguous()
        dlv = torch.empty_like(lv)
        _lib.check(lib.swb200_scm_tangent_target_logvar(F.data_ptr(), dF.data_ptr(), x_t.data_ptr(), dxt.data_ptr(), t1.data_ptr(),
                                                        float(r), sd, _lib.ptr(wv), _lib.ptr(wl), B, C, H, W, lv.data_ptr(),
                                                        g.data_ptr(), cot.data_ptr(), loss.data_ptr(), dlv.data_ptr(),
                                                        scratch.data_ptr(), need, stream), "scm_tangent_target_logvar")
        return {"loss": loss, "cot": cot, "g": g, "F": F, "dF": dF, "x_t": x_t, "logvar": lv, "dlogvar": dlv}
    _lib.check(lib.swb200_scm_tangent_target(F.data_ptr(), dF.data_ptr(), x_t.data_ptr(), dxt.data_ptr(), t1.data_ptr(),
                                             float(r), sd, _lib.ptr(wv), _lib.ptr(wl), B, C, H, W, g.data_ptr(),
                                             cot.data_ptr(), loss.data_ptr(), scratch.data_ptr(), need, stream),
               "scm_tangent_target")
    return {"loss": loss, "cot": cot, "g": g, "F": F, "dF": dF, "x_t": x_t}


def scm_backward(net, x: torch.Tensor, t: torch.Tensor, z: torch.Tensor, step: int, condition: Optional[torch.Tensor] = None,
                 auxiliary=None, **loss_kwargs) -> Dict[str, torch.Tensor]:
    """The backward of one sCM training step, all of it on the CUDA path: loss value, tangent target and
    ``cot = dL/dF_x`` from ``scm_output_cotangent`` (one stacked primal + tangent pass), then the one grad-enabled forward of
    ``loss.py:227`` through ``net`` -- a PassPrecond around ``swift_b200.swinv2.SwinV2`` in ``.train()`` mode, i.e. the
    reverse-mode path of ``training.py`` -- and ``F_x.backward(cot)``.  Leaves in ``.grad`` what
    ``SCMLoss(...)(net, x, step, ...).backward()`` leaves (trainer.py:206-214); returns the dict of
    ``scm_output_cotangent`` (``loss`` for logging).  No eager PyTorch module takes part."""
    inner = getattr(net, "module", net)
    if not inner.model.training:
        raise RuntimeError("scm_backward needs net.train(): the reverse-mode path is selected by training mode")
    sd = float(inner.sigma_data)
    if inner.model.logvar_embed is not None:
        # loss.py:222-232: the grad-enabled call returns (F_x, logvar); logvar weights the loss, so that call runs as soon as
        # x_t exists and both outputs receive their cotangents afterwards
        held = {}

        def grad_forward(x_t):
            with torch.enable_grad():
                held["F_x"], held["logvar"] = net(x_t / sd, t.to(x.device).reshape(-1), condition, auxiliary, return_logvar=True)
            return held["logvar"]

        out = scm_output_cotangent(net, x, t, z, step, condition=condition, auxiliary=auxiliary, logvar=grad_forward, **loss_kwargs)
        torch.autograd.backward([held["F_x"], held["logvar"]], [out["cot"], out["dlogvar"]])
        return out
    out = scm_output_cotangent(net, x, t, z, step, condition=condition, auxiliary=auxiliary, **loss_kwargs)
    with torch.enable_grad():
        F_x = net(out["x_t"] / sd, t.to(x.device).reshape(-1), condition, auxiliary)
    F_x.backward(out["cot"])
    return out
