"""The forward half of the sCM training step on the CUDA path (BASELINE.json configs[4]).

``SCMLoss.forward`` (stockeh/swift ``training/loss.py:192-260``) needs, per batch:

    x_t, dx_t/dt                                   elementwise                       :199-211
    (F, dF) = jvp(net, (x_t/sigma_d, t), (v_x, v_t))   the network + its tangent     :213-221
    F_x = net(x_t/sigma_d, t, ...)                 a SECOND, grad-enabled forward    :227
    g   = normalise(-cos^2 (sigma_d F - dx_t/dt) - r (cos sin x_t + sigma_d dF))     :233-248
    L   = mean sum w (F_x - sg(F_x) - g)^2                                           :253-260

Because F_x - sg(F_x) == 0, the loss value and its gradient with respect to the network output depend on g only:
``dL/dF_x = -2 w g / (B H W)``.  ``scm_output_cotangent`` returns exactly that, from ONE stacked primal + tangent pass of
the CUDA engine (``swb200_forward_jvp``; the reference runs the network twice here).  The reverse pass that would consume
``cot`` is not built: with the reference module as ``net`` for the backward, a training step is
``F_x = net(...); F_x.backward(cot)`` -- see INTEGRATION.md.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch

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
                         w_lat: Optional[torch.Tensor] = None, w_var: Optional[torch.Tensor] = None
                         ) -> Dict[str, torch.Tensor]:
    """x [B, C, H, W] targets, t = atan(tau / sigma_d) ([B] or [B,1,1,1]), z = sigma_d * N(0,1) like x (the draws of
    loss.py:196-200, made by the caller), ``net`` a PassPrecond around ``swift_b200.swinv2.SwinV2`` on a CUDA device.
    Returns {"loss", "cot" (= dL/dF_x), "g", "F", "dF", "x_t"}; all detached fp32."""
    inner = getattr(net, "module", net)
    model = inner.model
    if not hasattr(model, "engine"):
        raise TypeError("scm_output_cotangent needs swift_b200.swinv2.SwinV2 as net.model (there is no fallback path)")
    sd = float(inner.sigma_data)
    B = x.shape[0]
    x = x.to(torch.float32)
    t4 = t.to(device=x.device, dtype=torch.float32).reshape(B, 1, 1, 1)
    cos_t, sin_t = torch.cos(t4), torch.sin(t4)
    x_t = cos_t * x + sin_t * z                                         # loss.py:203
    dxt_dt = cos_t * z - sin_t * x                                      # :211
    v_x = cos_t * sin_t * dxt_dt / sd                                   # :216
    v_t = (cos_t * sin_t).reshape(B)                                    # :217
    x_in, dx_in = x_t / sd, v_x
    if condition is not None and inner.condition_channels > 0:          # precond.py:143-145; no tangent in the condition
        x_in = torch.cat([x_in, condition.to(torch.float32)], dim=1)
        dx_in = torch.cat([dx_in, torch.zeros_like(condition, dtype=torch.float32)], dim=1)
    aux = process_auxiliary(auxiliary, inner.auxiliary_dim, B, x.device)
    if aux is not None:
        aux = aux.to(torch.float32).expand(B, -1).contiguous()
    F, dF = model.engine().forward_jvp(x_in.contiguous(), t4.reshape(B).contiguous(), aux, dx_in.contiguous(),
                                       v_t.contiguous())
    r = min(1.0, step / (tangent_warmup_kimg * 1000)) if tangent_warmup_kimg > 0 else 1.0             # :233-237
    g = -(cos_t ** 2) * (sd * F - dxt_dt) - r * ((cos_t * sin_t) * x_t + sd * dF)                     # :241-243
    gn = torch.linalg.vector_norm(g, dim=(1, 2, 3), keepdim=True)
    g = g / (gn * math.sqrt(gn.numel() / g.numel()) + 0.1)                                            # :246-248
    w = 1.0
    if w_var is not None:
        w = w * w_var
    if w_lat is not None:
        w = w * w_lat
    loss = (w * g.square()).sum(dim=1).mean()                                                         # :253-260
    cot = -2.0 * w * g / (B * g.shape[2] * g.shape[3])
    return {"loss": loss, "cot": cot, "g": g, "F": F, "dF": dF, "x_t": x_t}
