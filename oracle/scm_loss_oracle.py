"""CPU restatement of the sCM training loss of stockeh/swift (``src/swift/training/loss.py:28-57,162-260``).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else); the product path never calls it.
Pinned to the reference: ``tests/golden/make_scm_loss_golden.py`` runs the real ``SCMLoss`` and records the loss and its
gradient with respect to the network output, plus the parameter gradients of ``loss.backward()`` for one case per fixture
(``tests/golden/scm_loss.npz``); ``tests/test_oracle_golden.py`` checks this file against them.

The loss value is a function of the detached tangent target ``g`` only,

    L = mean_{b,h,w} sum_c  w_var[c] w_lat[h] (F - sg(F) - g)^2 = mean sum w g^2,        dL/dF = -2 w g / (B H W),

so the reverse pass of the training step starts from ``cot = dL/dF``; everything up to ``cot`` needs the network forward
and its forward-mode tangent only (no reverse mode).  With a logvar head (``logvar: true``; loss.py:227-232, :252-258; off in
``model/swinv2.yaml:8``) the grad-enabled call returns (F_x, logvar [B]) and

    L = mean_{b,h,w} sum_c [exp(-logvar_b) w g^2 + logvar_b],   dL/dF = -2 exp(-logvar_b) w g / (B H W),
    dL/dlogvar_b = (C H W - exp(-logvar_b) sum_{c,h,w} w g^2) / (B H W).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Sequence

import torch

# loss.py:10-25
PRESSURE_LEVEL_VARS = ["geopotential", "u_component_of_wind", "v_component_of_wind", "vertical_velocity", "wind_speed",
                       "temperature", "relative_humidity", "specific_humidity", "vorticity", "potential_vorticity"]
PRESSURE_LEVELS = [50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000]
SINGLE_LEVEL_WEIGHTS = {"2m_temperature": 1.0, "sea_surface_temperature": 0.1, "10m_u_component_of_wind": 0.1,
                        "10m_v_component_of_wind": 0.1, "mean_sea_level_pressure": 0.1}


def latitude_weights(n_lat: int) -> torch.Tensor:
    """loss.py:28-32: cos(lat) on linspace(-90, 90), normalised to mean 1, then clamped at 0.1; [1, 1, n_lat, 1]."""
    w = torch.cos(torch.deg2rad(torch.linspace(-90, 90, n_lat)))
    w = w / w.mean()
    return torch.clamp(w, min=0.1).view(1, 1, -1, 1)


def variable_weights(variables: Sequence[str]) -> torch.Tensor:
    """loss.py:35-57: surface weights from the table, pressure-level weights proportional to the level; normalised to
    sum 1 over the variables present; [1, C, 1, 1]."""
    total = float(sum(PRESSURE_LEVELS))
    table = dict(SINGLE_LEVEL_WEIGHTS)
    for v in PRESSURE_LEVEL_VARS:
        for lv in PRESSURE_LEVELS:
            table[f"{v}_{lv}"] = lv / total
    w = torch.tensor([table[v] for v in variables], dtype=torch.float32).view(1, -1, 1, 1)
    return w / w.sum()


def tangent_warmup(step: int, tangent_warmup_kimg: int) -> float:
    """loss.py:233-237."""
    return min(1.0, step / (tangent_warmup_kimg * 1000)) if tangent_warmup_kimg > 0 else 1.0


def scm_tangent_target(F: torch.Tensor, dF: torch.Tensor, x_t: torch.Tensor, dxt_dt: torch.Tensor, cos_t: torch.Tensor,
                       sin_t: torch.Tensor, r: float, sigma_data: float) -> torch.Tensor:
    """loss.py:239-248: the JVP rearrangement and the tangent normalisation (norm over (C, H, W), made invariant to the
    spatial size, + 0.1)."""
    g = -(cos_t ** 2) * (sigma_data * F - dxt_dt) - r * ((cos_t * sin_t) * x_t + sigma_data * dF)
    gn = torch.linalg.vector_norm(g, dim=(1, 2, 3), keepdim=True)
    gn = gn * math.sqrt(gn.numel() / g.numel())
    return g / (gn + 0.1)


def scm_loss(net: Callable, x: torch.Tensor, t: torch.Tensor, z: torch.Tensor, step: int, tangent_warmup_kimg: int,
             w_lat: torch.Tensor, w_var: torch.Tensor, sigma_data: float = 1.0, logvar=None,
             net_pretrained: Callable = None) -> Dict[str, torch.Tensor]:
    """loss.py:192-260 as a deterministic function of the draws (t = atan(tau / sigma_d) [B,1,1,1], z = sigma_d * N(0,1)).

    ``net(x_in, t_flat) -> F`` is the denoiser with condition / auxiliary bound (the reference's ``wrapper``, :213-214);
    it must be differentiable in forward mode (``torch.func.jvp``).  Returns loss, cot = dL/dF, g, F, dF, x_t."""
    cos_t, sin_t = torch.cos(t), torch.sin(t)
    x_t = cos_t * x + sin_t * z                                         # :203
    if net_pretrained is not None:                                      # :205-209 distillation: a frozen v-prediction teacher
        with torch.no_grad():
            dxt_dt = sigma_data * net_pretrained(x_t / sigma_data, t.flatten())
    else:
        dxt_dt = cos_t * z - sin_t * x                                  # :211
    v_x = cos_t * sin_t * dxt_dt / sigma_data                           # :216
    v_t = cos_t * sin_t                                                 # :217
    F, dF = torch.func.jvp(lambda a, b: net(a, b.flatten()), (x_t / sigma_data, t), (v_x, v_t))      # :218-221
    F, dF = F.detach(), dF.detach()
    g = scm_tangent_target(F, dF, x_t, dxt_dt, cos_t, sin_t, tangent_warmup(step, tangent_warmup_kimg), sigma_data)
    w = w_var * w_lat
    bhw = g.shape[0] * g.shape[2] * g.shape[3]
    if logvar is not None:                                              # :227-232, :252-258 (``logvar`` [B], detached values)
        lv = logvar.detach().reshape(-1, 1, 1, 1).to(g.dtype)
        e = torch.exp(-lv)
        loss = (e * (w * g.square()) + lv).sum(dim=1).mean()
        cot = -2.0 * e * w * g / bhw
        s_b = (w * g.square()).sum(dim=(1, 2, 3))
        dlogvar = (g.shape[1] * g.shape[2] * g.shape[3] - e.flatten() * s_b) / bhw
        return {"loss": loss, "cot": cot, "g": g, "F": F, "dF": dF, "x_t": x_t, "dlogvar": dlogvar}
    loss = (w * g.square()).sum(dim=1).mean()                           # :253-260 with F - sg(F) = 0, logvar = 0
    cot = -2.0 * w * g / bhw
    return {"loss": loss, "cot": cot, "g": g, "F": F, "dF": dF, "x_t": x_t}


def scm_parameter_gradients(net_of: Callable, params: Dict[str, torch.Tensor], x: torch.Tensor, t: torch.Tensor,
                            z: torch.Tensor, step: int, tangent_warmup_kimg: int, w_lat: torch.Tensor, w_var: torch.Tensor,
                            sigma_data: float = 1.0) -> Dict[str, torch.Tensor]:
    """d loss / d parameter of one ``SCMLoss`` evaluation (what ``loss.backward()`` leaves in ``.grad``, trainer.py:206-214):
    the target of the reverse pass.  ``net_of(p)`` returns the denoiser ``(x_in, t_flat) -> F`` bound to the parameter dict
    ``p``.  The loss only sees the parameters through ``F_x`` (g is detached, loss.py:241-248), so the gradients are the
    vector-Jacobian product of the grad-enabled forward (:227) with ``cot = dL/dF_x``."""
    out = scm_loss(net_of(params), x, t, z, step, tangent_warmup_kimg, w_lat, w_var, sigma_data)
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    F = net_of(leaf)(out["x_t"] / sigma_data, t.flatten())
    F.backward(out["cot"])
    return {k: v.grad for k, v in leaf.items() if v.grad is not None}


def scm_parameter_gradients_logvar(net_of: Callable, net_lv_of: Callable, params: Dict[str, torch.Tensor], x: torch.Tensor,
                                   t: torch.Tensor, z: torch.Tensor, step: int, tangent_warmup_kimg: int, w_lat: torch.Tensor,
                                   w_var: torch.Tensor, sigma_data: float = 1.0) -> Dict[str, torch.Tensor]:
    """``scm_parameter_gradients`` for a model with a logvar head: ``net_lv_of(p)`` returns ``(x_in, t_flat) -> (F, logvar)``
    (the grad-enabled call of loss.py:222-232).  Both outputs receive their cotangents: dL/dF_x and dL/dlogvar."""
    with torch.no_grad():
        _, lv = net_lv_of(params)(torch.zeros_like(x), t.flatten())      # logvar depends on (t, aux) only
    out = scm_loss(net_of(params), x, t, z, step, tangent_warmup_kimg, w_lat, w_var, sigma_data, logvar=lv)
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    F, lv2 = net_lv_of(leaf)(out["x_t"] / sigma_data, t.flatten())
    torch.autograd.backward([F, lv2], [out["cot"], out["dlogvar"].to(lv2.dtype)])
    grads = {k: v.grad for k, v in leaf.items() if v.grad is not None}
    grads["__loss__"] = out["loss"].detach()
    grads["__dlogvar__"] = out["dlogvar"].detach()
    grads["__cot__"] = out["cot"].detach()
    return grads
