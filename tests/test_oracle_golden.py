"""The oracle restatement (oracle/swinv2_oracle.py) against outputs of the REAL reference.

The golden files were produced by tests/golden/make_golden.py, which imports stockeh/swift from
/root/reference and runs its own SwinV2 / PassPrecond / DiffusionSampler on the seeded fixtures.
Both sides are fp32 on CPU, so agreement is to rounding (different op order only).
"""
import math

import numpy as np
import pytest
import torch

from oracle import swinv2_oracle as orc
from swift_b200 import synthetic as syn

TAP_STRIDE = 8


def _net(params, cfg):
    return lambda x, t, cond, aux: orc.pass_precond(params, cfg, x, t, cond, aux)


def _close(a, b, tol=2e-5):
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    err = (a - b).norm() / b.norm().clamp_min(1e-30)
    assert err < tol, f"rel L2 {err:.3e}"
    assert torch.allclose(a, b, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("name,cfgname", [("tiny", "SWIFT_TINY"), ("small", "SWIFT_SMALL")])
def test_forward_and_taps(golden, name, cfgname):
    g = golden(name)
    c = getattr(syn, cfgname)
    cfg = orc.make_cfg(**c)
    p = syn.random_state_dict(c, seed=1)
    lat, cond = syn.synthetic_fields(c, 2, seed=3)
    taps = {}
    y = orc.swinv2_forward(p, cfg, torch.cat([lat, cond], 1), torch.from_numpy(g["fwd_t"]),
                           torch.from_numpy(g["fwd_aux"]), taps=taps)
    _close(y, g["fwd_y"])
    # the reference's two attention branches (swinv2.py:128-133: SDPA at inference, explicit softmax under jvp)
    y_sdpa = orc.swinv2_forward(p, dict(cfg, sdpa=True), torch.cat([lat, cond], 1), torch.from_numpy(g["fwd_t"]),
                                torch.from_numpy(g["fwd_aux"]))
    _close(y_sdpa, g["fwd_y"])
    _close(taps["cond"], g["tap_cond"])
    for i in range(c["depth"]):
        _close(taps[f"block{i}"][:, ::TAP_STRIDE], g[f"tap_block{i}"])
    # the reference's patch_embed tap is taken before +pos_embed
    _close(taps["embed"][:, ::TAP_STRIDE] - p["pos_embed"][:, ::TAP_STRIDE], g["tap_patch_embed"], tol=1e-4)


@pytest.mark.parametrize("name,cfgname", [("tiny", "SWIFT_TINY"), ("small", "SWIFT_SMALL")])
def test_samplers(golden, name, cfgname):
    g = golden(name)
    c = getattr(syn, cfgname)
    cfg = orc.make_cfg(**c)
    p = syn.random_state_dict(c, seed=1)
    lat, cond = syn.synthetic_fields(c, 2, seed=3)
    net = _net(p, cfg)
    z = torch.from_numpy(g["scm2_noise"])
    _close(orc.scm_solver(net, lat, cond, 0.6, num_steps=1), g["scm1"])
    _close(orc.scm_solver(net, lat, cond, 0.6, num_steps=2, noise_fn=lambda x: z), g["scm2"])
    _close(orc.scm_solver(net, lat, cond, 0.6, num_steps=3, noise_fn=lambda x: z), g["scm3"])
    _close(orc.dpm_solver_2s(net, lat, cond, 0.6, num_steps=3), g["dpm2s_3"], tol=1e-4)


def test_swift_b_digest(golden):
    """Full Swift-B (226 M parameters), one sCM step, against the reference digest (about 10 s on 8 cores)."""
    g = golden("swift_b")
    c = syn.SWIFT_B
    cfg = orc.make_cfg(**c)
    p = syn.random_state_dict(c, seed=1)
    assert sum(v.numel() for v in p.values()) == 225_980_976       # SURVEY.md section 6 [probe]
    lat, cond = syn.synthetic_fields(c, 1, seed=0)
    with torch.no_grad():
        y = orc.scm_solver(_net(p, cfg), lat, cond, 0.6, num_steps=1)
    _close(y.flatten(2).norm(dim=-1), g["scm1_channel_l2"], tol=1e-5)
    _close(y[:, :, ::8, ::8], g["scm1_sub"], tol=5e-5)
    _close(y[0, ::17, 37, :], g["scm1_row"], tol=5e-5)


def test_time_grid_and_embedding():
    ts = orc.scm_time_grid(1, 0.02, 200.0, 1.0)
    assert ts.tolist() == [pytest.approx(math.pi / 2), 0.0]
    ts = orc.scm_time_grid(2, 0.02, 200.0, 1.0)
    assert ts[1].item() == pytest.approx(1.1) and ts[2].item() == 0.0
    e = orc.timestep_embedding(torch.tensor([0.7]), 8)
    f = torch.exp(-math.log(10000.0) * torch.arange(4) / 4)
    np.testing.assert_allclose(e[0, :4].numpy(), torch.sin(0.7 * f).numpy(), rtol=1e-6)
    np.testing.assert_allclose(e[0, 4:].numpy(), torch.cos(0.7 * f).numpy(), rtol=1e-6)


SCM_LOSS_VARIABLES = ["2m_temperature", "10m_u_component_of_wind", "mean_sea_level_pressure", "geopotential_500",
                      "temperature_850", "specific_humidity_700"]


@pytest.mark.parametrize("name,cfgname", [("tiny", "SWIFT_TINY"), ("small", "SWIFT_SMALL")])
def test_scm_loss_and_output_cotangent(golden, name, cfgname):
    """oracle/scm_loss_oracle.py against the real ``SCMLoss`` (training/loss.py:162-260): loss value and dL/dF_x (what the
    reverse pass starts from), three (noise draw, step, tangent warm-up) cases per fixture."""
    from oracle import scm_loss_oracle as so
    g = golden("scm_loss")
    c = getattr(syn, cfgname)
    n_img, (H, W) = c["out_channels"], c["img_resolution"]
    cfg = orc.make_cfg(**c)
    p = syn.random_state_dict(c, seed=1)
    x, cond = syn.synthetic_fields(c, 2, seed=5)
    w_lat, w_var = so.latitude_weights(H), so.variable_weights(SCM_LOSS_VARIABLES[:n_img])
    np.testing.assert_array_equal(w_lat.numpy(), g[name + "_w_lat"])
    np.testing.assert_array_equal(w_var.numpy(), g[name + "_w_var"])
    net = lambda a, b: orc.pass_precond(p, cfg, a, b, cond, 0.6)
    for case in range(3):
        k = f"{name}_{case}_"
        step, warm = (int(v) for v in g[k + "step_warm"])
        o = so.scm_loss(net, x, torch.from_numpy(g[k + "t"]), torch.from_numpy(g[k + "z"]), step, warm, w_lat, w_var)
        assert abs(float(o["loss"]) - float(g[k + "loss"])) < 2e-6 * float(g[k + "loss"])
        _close(o["F"], g[k + "F"])
        _close(o["cot"], g[k + "cot"], tol=1e-5)
    assert so.tangent_warmup(500_000, 3000) == pytest.approx(1 / 6) and so.tangent_warmup(5, 0) == 1.0


@pytest.mark.parametrize("name,cfgname", [("tiny", "SWIFT_TINY"), ("small", "SWIFT_SMALL")])
def test_scm_parameter_gradients_match_reference_backward(golden, name, cfgname):
    """The target of the reverse pass (not built yet in the CUDA path): d loss / d parameter of the real reference's
    ``SCMLoss(...).backward()`` -- the norm of EVERY parameter gradient and strided samples of nine tensors -- reproduced by
    the oracle as the vector-Jacobian product of its forward with cot = dL/dF_x."""
    from oracle import scm_loss_oracle as so
    g = golden("scm_loss")
    c = getattr(syn, cfgname)
    n_img, (H, W) = c["out_channels"], c["img_resolution"]
    cfg = orc.make_cfg(**c)
    p = syn.random_state_dict(c, seed=1)
    x, cond = syn.synthetic_fields(c, 2, seed=5)
    w_lat, w_var = so.latitude_weights(H), so.variable_weights(SCM_LOSS_VARIABLES[:n_img])
    k = f"{name}_0_"
    step, warm = (int(v) for v in g[k + "step_warm"])
    grads = so.scm_parameter_gradients(lambda q: (lambda a, b: orc.pass_precond(q, cfg, a, b, cond, 0.6)), p, x,
                                       torch.from_numpy(g[k + "t"]), torch.from_numpy(g[k + "z"]), step, warm, w_lat, w_var)
    names = [str(s) for s in g[k + "grad_names"]]
    assert sorted(names) == sorted("model." + n for n in p)                  # every parameter receives a gradient
    for nm, ref in zip(names, g[k + "grad_norms"]):
        got = float(grads[nm[len("model."):]].norm())
        assert abs(got - ref) < 2e-5 * ref, (nm, got, ref)
    sampled = [kk for kk in g if kk.startswith(k + "grad:")]
    assert len(sampled) == 9
    for kk in sampled:
        nm = kk.split("grad:")[1][len("model."):]
        _close(grads[nm].flatten()[::31], g[kk], tol=5e-5)


@pytest.mark.parametrize("name,cfgname", [("tiny", "SWIFT_TINY"), ("small", "SWIFT_SMALL")])
def test_scm_logvar_loss_and_gradients_match_reference(golden, name, cfgname):
    """SCMLoss around a model WITH a logvar head (loss.py:222-232, :252-258; tests/golden/make_scm_logvar_golden.py ran the real
    reference): loss, logvar, dL/dF_x, dL/dlogvar, the norm of every parameter gradient and strided samples -- reproduced by
    the oracle (both outputs of the grad-enabled call receive their cotangents)."""
    from oracle import scm_loss_oracle as so
    g = golden("scm_logvar")
    c = getattr(syn, cfgname)
    n_img, (H, W) = c["out_channels"], c["img_resolution"]
    cfg = orc.make_cfg(**c, logvar=True)
    p = syn.random_state_dict(c, seed=1, logvar=True)
    x, cond = syn.synthetic_fields(c, 2, seed=5)
    w_lat, w_var = so.latitude_weights(H), so.variable_weights(SCM_LOSS_VARIABLES[:n_img])
    k = name + "_"
    step, warm = (int(v) for v in g[k + "step_warm"])
    aux = torch.full((2, 1), 0.6)

    def net_lv_of(q):
        return lambda a, b: orc.swinv2_forward(q, cfg, torch.cat([a, cond], 1), b.flatten(), aux, return_logvar=True)

    grads = so.scm_parameter_gradients_logvar(lambda q: (lambda a, b: orc.pass_precond(q, cfg, a, b, cond, 0.6)), net_lv_of, p, x,
                                              torch.from_numpy(g[k + "t"]), torch.from_numpy(g[k + "z"]), step, warm, w_lat, w_var)
    assert abs(float(grads.pop("__loss__")) - float(g[k + "loss"])) < 2e-5 * abs(float(g[k + "loss"]))
    _close(grads.pop("__dlogvar__"), g[k + "dlogvar"], tol=2e-5)
    _close(grads.pop("__cot__"), g[k + "cot"], tol=2e-5)
    names = [str(s) for s in g[k + "grad_names"]]
    assert sorted(names) == sorted("model." + n for n in p)                  # the head's parameters included
    for nm, ref in zip(names, g[k + "grad_norms"]):
        got = float(grads[nm[len("model."):]].norm())
        assert abs(got - ref) < 5e-5 * ref, (nm, got, ref)
    for kk in g:
        if kk.startswith(k + "grad:"):
            nm = kk.split("grad:")[1][len("model."):]
            _close(grads[nm].flatten()[::31], g[kk], tol=1e-4)


@pytest.mark.parametrize("name,cfgname", [("tiny", "SWIFT_TINY"), ("small", "SWIFT_SMALL")])
def test_scm_distillation_loss_matches_reference(golden, name, cfgname):
    """``SCMLoss(distillation=True)`` with a ``net_pretrained`` teacher (loss.py:205-210; the real reference ran in
    tests/golden/make_scm_distill_golden.py with the seed-2 fixture as teacher): loss and dL/dF_x from the oracle."""
    from oracle import scm_loss_oracle as so
    g = golden("scm_distill")
    c = getattr(syn, cfgname)
    n_img, (H, W) = c["out_channels"], c["img_resolution"]
    cfg = orc.make_cfg(**c)
    p, p_teacher = syn.random_state_dict(c, seed=1), syn.random_state_dict(c, seed=2)
    x, cond = syn.synthetic_fields(c, 2, seed=5)
    k = name + "_"
    step, warm = (int(v) for v in g[k + "step_warm"])
    out = so.scm_loss(lambda a, b: orc.pass_precond(p, cfg, a, b, cond, 0.6), x, torch.from_numpy(g[k + "t"]),
                      torch.from_numpy(g[k + "z"]), step, warm, so.latitude_weights(H), so.variable_weights(SCM_LOSS_VARIABLES[:n_img]),
                      net_pretrained=lambda a, b: orc.pass_precond(p_teacher, cfg, a, b, cond, 0.6))
    assert abs(float(out["loss"]) - float(g[k + "loss"])) < 2e-5 * float(g[k + "loss"])
    _close(out["cot"], g[k + "cot"], tol=5e-5)


def test_muon_oracle_matches_reference_golden(golden):
    """oracle/muon_oracle.py against the REAL reference's muon_update / adam_update (tests/golden/make_muon_golden.py)."""
    from oracle import muon_oracle as mo
    g = golden("muon")
    torch.set_num_threads(1)
    for name in ("wide", "tall", "square", "scale"):
        grad, mom = torch.from_numpy(g[f"{name}_grad"]), torch.from_numpy(g[f"{name}_mom"])
        upd, mom_new = mo.muon_update(grad, mom, beta=0.95)
        assert torch.equal(mom_new, torch.from_numpy(g[f"{name}_mom_new"])), name
        # same bf16 op sequence; the CPU matmul may block the reduction differently from run to run of the build
        ref = torch.from_numpy(g[f"{name}_update"])
        assert ((upd - ref).norm() / ref.norm()).item() < 2e-2, name
    p, gr = torch.from_numpy(g["adam_p"]), torch.from_numpy(g["adam_g"])
    p2, b1, b2 = mo.adam_step(p, gr, torch.from_numpy(g["adam_b1"]), torch.from_numpy(g["adam_b2"]), 3, lr=1e-2, weight_decay=0.1)
    assert torch.allclose(b1, torch.from_numpy(g["adam_b1_new"])) and torch.allclose(b2, torch.from_numpy(g["adam_b2_new"]))
    ref_p = p * (1 - 1e-2 * 0.1) - 1e-2 * torch.from_numpy(g["adam_update"])
    assert torch.allclose(p2, ref_p, rtol=1e-6, atol=1e-7)
