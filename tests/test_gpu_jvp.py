"""Forward-mode tangent path (first part): (F, dF) = jvp(net, (x, t), (v_x, v_t)) as the sCM
training loss evaluates it (reference training/loss.py:216-225), against torch.func.jvp of the fp32 oracle."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _per_field(y, ref):
    num = (y.double() - ref.double()).flatten(2).norm(dim=-1)
    return (num / ref.double().flatten(2).norm(dim=-1).clamp_min(1e-30)).max().item()


@pytest.mark.parametrize("act_fp16", [True, False], ids=["act_fp16", "act_bf16"])
@pytest.mark.parametrize("cfgname", ["SWIFT_TINY", "SWIFT_SMALL"])
def test_engine_forward_jvp_vs_oracle(cfgname, act_fp16):
    from oracle import swinv2_oracle as orc
    from swift_b200 import synthetic as syn
    from test_gpu_forward import build_net
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = getattr(syn, cfgname)
    net, sd = build_net(cfg, act_fp16=act_fp16)
    tol = 1.0 if act_fp16 else 4.0                        # bf16 operands: 8x coarser rounding of every GEMM input
    B = 2
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, cfg["in_channels"], *cfg["img_resolution"], generator=g).cuda()
    dx = torch.randn(x.shape, generator=g).cuda()
    dx[:, cfg["out_channels"]:] = 0                       # the condition channels carry no tangent (loss.py:213-215)
    t = torch.tensor([0.3, 1.2]).cuda()
    dt = torch.tensor([0.25, -0.4]).cuda()
    aux = torch.full((B, 1), 0.6).cuda()
    y, dy = net.model.engine().forward_jvp(x, t, aux, dx, dt)
    ocfg = orc.make_cfg(**cfg)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    f = lambda xx, tt: orc.swinv2_forward(sd_gpu, ocfg, xx, tt, aux)
    ref, dref = torch.func.jvp(f, (x, t), (dx, dt))
    print(f"{cfgname}: F per-field rel-L2 {_per_field(y, ref):.3e}, dF per-field rel-L2 {_per_field(dy, dref):.3e}")
    assert _per_field(y, ref) < 5e-3 * tol
    assert _per_field(dy, dref) < 1e-2 * tol
    # each tangent direction separately (x only / t only): catches a missing term that the sum could hide
    for dxx, dtt in ((dx, torch.zeros_like(dt)), (torch.zeros_like(dx), dt)):
        _, d1 = net.model.engine().forward_jvp(x, t, aux, dxx, dtt)
        _, r1 = torch.func.jvp(f, (x, t), (dxx, dtt))
        assert _per_field(d1, r1) < 1e-2 * tol


def test_torch_func_jvp_through_the_module_as_the_scm_loss_does():
    """training/loss.py:216-225: torch.func.jvp(lambda x, t: net(x, t, condition, auxiliary, jvp=True), ...)."""
    from oracle import swinv2_oracle as orc
    from swift_b200 import synthetic as syn
    from test_gpu_forward import build_net
    cfg = syn.SWIFT_TINY
    n_img = cfg["out_channels"]
    net, sd = build_net(cfg, img_channels=n_img)
    B = 2
    g = torch.Generator().manual_seed(12)
    x_t = torch.randn(B, n_img, *cfg["img_resolution"], generator=g).cuda()
    cond = torch.randn(B, cfg["in_channels"] - n_img, *cfg["img_resolution"], generator=g).cuda()
    t = torch.tensor([0.7, 1.4]).cuda()
    cos_t, sin_t = torch.cos(t).view(-1, 1, 1, 1), torch.sin(t).view(-1, 1, 1, 1)
    v_x = cos_t * sin_t * torch.randn(x_t.shape, generator=g).cuda()
    v_t = (torch.cos(t) * torch.sin(t))

    def wrapper(x, tt):
        return net(x, tt, cond, 0.6, jvp=True)

    F, dF = torch.func.jvp(wrapper, (x_t, t), (v_x, v_t))
    ocfg = orc.make_cfg(**cfg)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    ref, dref = torch.func.jvp(lambda x, tt: orc.pass_precond(sd_gpu, ocfg, x, tt, cond, 0.6), (x_t, t), (v_x, v_t))
    assert _per_field(F, ref) < 5e-3
    assert _per_field(dF, dref) < 1e-2
    with torch.no_grad():                                   # jvp=True without a transform is just the forward
        assert torch.equal(net(x_t, t, cond, 0.6, jvp=True), net(x_t, t, cond, 0.6))


def test_swift_b_tangent_forward_vs_oracle_and_rate():
    """Swift-B (config 5's network), batch 1: tangent forward vs torch.func.jvp of the fp32 oracle on the same GPU, and
    its rate next to the plain forward (reported, not asserted: the dual kernels are plain CUDA-core code)."""
    import time
    from oracle import swinv2_oracle as orc
    from swift_b200 import synthetic as syn
    from test_gpu_forward import build_net
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = syn.SWIFT_B
    net, sd = build_net(cfg, img_channels=syn.IMG_CHANNELS)
    g = torch.Generator().manual_seed(13)
    x = torch.randn(1, cfg["in_channels"], *cfg["img_resolution"], generator=g).cuda()
    dx = torch.randn(x.shape, generator=g).cuda()
    dx[:, syn.IMG_CHANNELS:] = 0
    t = torch.tensor([0.9]).cuda()
    dt = torch.tensor([0.45]).cuda()
    aux = torch.full((1, 1), 0.6).cuda()
    eng = net.model.engine()
    y, dy = eng.forward_jvp(x, t, aux, dx, dt)
    torch.cuda.synchronize()
    ocfg = orc.make_cfg(**cfg)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref, dref = torch.func.jvp(lambda xx, tt: orc.swinv2_forward(sd_gpu, ocfg, xx, tt, aux), (x, t), (dx, dt))
    e_f, e_df = _per_field(y, ref), _per_field(dy, dref)
    t0 = time.perf_counter()
    for _ in range(3):
        eng.forward_jvp(x, t, aux, dx, dt)
    torch.cuda.synchronize()
    ms_jvp = (time.perf_counter() - t0) / 3 * 1e3
    cond = eng.conditioning(t, aux)
    t0 = time.perf_counter()
    for _ in range(3):
        eng.forward(x, None, cond[0], cond[1])
    torch.cuda.synchronize()
    ms_fwd = (time.perf_counter() - t0) / 3 * 1e3
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(2):
            torch.func.jvp(lambda xx, tt: orc.swinv2_forward(sd_gpu, ocfg, xx, tt, aux), (x, t), (dx, dt))
    torch.cuda.synchronize()
    ms_ref = (time.perf_counter() - t0) / 2 * 1e3
    print(f"Swift-B tangent forward: F per-field rel-L2 {e_f:.3e}, dF {e_df:.3e}; {ms_jvp:.1f} ms per sample "
          f"(plain forward at batch 1: {ms_fwd:.1f} ms; torch.func.jvp of the fp32 PyTorch restatement on the same GPU: "
          f"{ms_ref:.1f} ms)")
    assert e_f < 5e-3 and e_df < 1e-2


@pytest.mark.parametrize("name,cfgname", [("tiny", "SWIFT_TINY"), ("small", "SWIFT_SMALL")])
def test_scm_output_cotangent_vs_reference_loss_golden(golden, name, cfgname):
    """swift_b200.scm_target.scm_output_cotangent (one stacked primal + tangent pass of the CUDA engine) against the REAL
    reference's SCMLoss (tests/golden/make_scm_loss_golden.py): loss value and dL/dF_x, three (draw, step, warm-up) cases.
    Tolerance: the 16-bit-operand error of F (5e-3) and dF (1e-2) propagated through the normalised tangent target."""
    from swift_b200 import synthetic as syn
    from swift_b200.scm_target import latitude_weights, scm_output_cotangent, variable_weights
    from test_gpu_forward import build_net
    from test_oracle_golden import SCM_LOSS_VARIABLES
    g = golden("scm_loss")
    cfg = getattr(syn, cfgname)
    n_img, (H, W) = cfg["out_channels"], cfg["img_resolution"]
    net, _ = build_net(cfg, img_channels=n_img)
    x, cond = (v.cuda() for v in syn.synthetic_fields(cfg, 2, seed=5))
    w_lat, w_var = latitude_weights(H, "cuda"), variable_weights(SCM_LOSS_VARIABLES[:n_img], "cuda")
    np.testing.assert_array_equal(w_lat.cpu().numpy(), g[name + "_w_lat"])
    np.testing.assert_array_equal(w_var.cpu().numpy(), g[name + "_w_var"])
    for case in range(3):
        k = f"{name}_{case}_"
        step, warm = (int(v) for v in g[k + "step_warm"])
        o = scm_output_cotangent(net, x, torch.from_numpy(g[k + "t"]).cuda(), torch.from_numpy(g[k + "z"]).cuda(), step,
                                 condition=cond, auxiliary=0.6, tangent_warmup_kimg=warm, w_lat=w_lat, w_var=w_var)
        ref_cot, ref_loss = torch.from_numpy(g[k + "cot"]).cuda(), float(g[k + "loss"])
        e_f, e_c = _per_field(o["F"], torch.from_numpy(g[k + "F"]).cuda()), _rel(o["cot"], ref_cot)
        print(f"{k}: F per-field rel-L2 {e_f:.3e}, cotangent rel-L2 {e_c:.3e}, loss {float(o['loss']):.6f} vs {ref_loss:.6f}")
        assert e_f < 5e-3 and e_c < 2e-2
        assert abs(float(o["loss"]) - ref_loss) < 1e-2 * ref_loss
    with pytest.raises(KeyError):
        variable_weights(["2m_temperature", "not_a_variable_500"])


def test_scm_training_step_gradients_vs_reference_backward(golden):
    """scm_target.scm_backward: cot from the tangent forward, then F_x.backward(cot) through swift_b200.SwinV2's reverse-mode
    path -> the parameter gradients of the REAL reference's SCMLoss(...).backward() (tests/golden/scm_loss.npz): the norm of
    every tensor's gradient and strided samples of nine tensors.  Tolerance 5e-2 (bf16 operands in forward and backward;
    the fp32 reference differs from the fp32 oracle by 1e-5 here)."""
    from swift_b200 import synthetic as syn
    from swift_b200.scm_target import scm_backward, latitude_weights, variable_weights
    from test_gpu_forward import build_net
    from test_oracle_golden import SCM_LOSS_VARIABLES
    g = golden("scm_loss")
    for cfgname, k in (("SWIFT_TINY", "tiny_0_"), ("SWIFT_SMALL", "small_0_")):
        cfg = getattr(syn, cfgname)
        n_img, (H, W) = cfg["out_channels"], cfg["img_resolution"]
        net, sd = build_net(cfg, img_channels=n_img)
        net.train()
        x, cond = (v.cuda() for v in syn.synthetic_fields(cfg, 2, seed=5))
        step, warm = (int(v) for v in g[k + "step_warm"])
        out = scm_backward(net, x, torch.from_numpy(g[k + "t"]).cuda(), torch.from_numpy(g[k + "z"]).cuda(), step,
                           condition=cond, auxiliary=0.6, tangent_warmup_kimg=warm, w_lat=latitude_weights(H, "cuda"),
                           w_var=variable_weights(SCM_LOSS_VARIABLES[:n_img], "cuda"))
        assert abs(float(out["loss"]) - float(g[k + "loss"])) < 1e-2 * float(g[k + "loss"])
        grads = {"model." + name: p.grad for name, p in net.model.named_parameters()}
        worst, worst_nm, worst_mat = 0.0, "", 0.0
        for nm, ref in zip((str(s) for s in g[k + "grad_names"]), g[k + "grad_norms"]):
            got = float(grads[nm].norm())
            e = abs(got - ref) / ref
            if e > worst:
                worst, worst_nm = e, nm
            if not nm.endswith(".scale"):            # (the logit scale: one number per head, a sum of cancelling terms)
                worst_mat = max(worst_mat, e)
        worst_s = 0.0
        for kk in g:
            if kk.startswith(k + "grad:"):
                nm = kk.split("grad:")[1]
                worst_s = max(worst_s, _rel(grads[nm].flatten()[::31].cpu(), torch.from_numpy(g[kk])))
        print(f"sCM step on the CUDA path ({cfgname}): worst gradient-norm error {worst:.3e} ({worst_nm}), without the logit "
              f"scales {worst_mat:.3e}; worst sampled-gradient rel-L2 {worst_s:.3e}")
        assert worst < 5e-2 and worst_mat < 1e-2 and worst_s < 5e-2


@pytest.mark.parametrize("route", ["autograd", "train_step"])
def test_scm_logvar_training_step_vs_reference_backward(golden, route):
    """SCMLoss with a logvar head (loss.py:222-232, :252-258) on the CUDA path, both routes: ``scm_backward`` (the grad-enabled
    call returns (F_x, logvar) and both receive cotangents through autograd) and ``scm_train_step`` (no autograd in the loop),
    against the REAL reference's loss, logvar, dL/dlogvar and parameter gradients (tests/golden/scm_logvar.npz)."""
    from swift_b200 import synthetic as syn
    from swift_b200.precond import PassPrecond
    from swift_b200.scm_target import scm_backward, latitude_weights, variable_weights
    from swift_b200.training import scm_train_step
    from test_oracle_golden import SCM_LOSS_VARIABLES
    g = golden("scm_logvar")
    for cfgname, k in (("SWIFT_TINY", "tiny_"), ("SWIFT_SMALL", "small_")):
        cfg = getattr(syn, cfgname)
        n_img, (H, W) = cfg["out_channels"], cfg["img_resolution"]
        model_cfg = dict(_target_="swift_b200.swinv2.SwinV2", window_size=cfg["window_size"], shift_size=cfg["shift_size"],
                         patch_size=cfg["patch_size"], depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"], logvar=True,
                         timestep_weight=1.0)
        net = PassPrecond(model_cfg, img_resolution=cfg["img_resolution"], img_channels=n_img,
                          condition_channels=cfg["in_channels"] - n_img, auxiliary_dim=cfg["auxiliary_dim"], sigma_data=1.0)
        net.load_state_dict(syn.random_state_dict(cfg, seed=1, prefix="model.", logvar=True), strict=True)
        net = net.cuda().train()
        x, cond = (v.cuda() for v in syn.synthetic_fields(cfg, 2, seed=5))
        step, warm = (int(v) for v in g[k + "step_warm"])
        fn = scm_backward if route == "autograd" else scm_train_step
        out = fn(net, x, torch.from_numpy(g[k + "t"]).cuda(), torch.from_numpy(g[k + "z"]).cuda(), step, condition=cond,
                 auxiliary=0.6, tangent_warmup_kimg=warm, w_lat=latitude_weights(H, "cuda"),
                 w_var=variable_weights(SCM_LOSS_VARIABLES[:n_img], "cuda"))
        assert abs(float(out["loss"]) - float(g[k + "loss"])) < 1e-2 * abs(float(g[k + "loss"]))
        assert _rel(out["logvar"].cpu(), torch.from_numpy(g[k + "logvar"]).flatten()) < 2e-3
        assert _rel(out["dlogvar"].cpu(), torch.from_numpy(g[k + "dlogvar"]).flatten()) < 1e-2
        assert _rel(out["cot"].cpu(), torch.from_numpy(g[k + "cot"])) < 2e-3
        grads = {"model." + name: p.grad for name, p in net.model.named_parameters()}
        worst, worst_nm = 0.0, ""
        for nm, ref in zip((str(s) for s in g[k + "grad_names"]), g[k + "grad_norms"]):
            e = abs(float(grads[nm].norm()) - ref) / ref
            if e > worst and not nm.endswith(".scale"):
                worst, worst_nm = e, nm
        worst_s = 0.0
        for kk in g:
            if kk.startswith(k + "grad:"):
                nm = kk.split("grad:")[1]
                worst_s = max(worst_s, _rel(grads[nm].flatten()[::31].cpu().float(), torch.from_numpy(g[kk])))
        print(f"sCM step with a logvar head ({cfgname}, {route}): worst gradient-norm error {worst:.3e} ({worst_nm}); worst "
              f"sampled-gradient rel-L2 {worst_s:.3e}; head gradient norms {float(grads['model.logvar_embed.weight'].norm()):.4e} "
              f"{float(grads['model.logvar_embed.bias'].norm()):.4e}")
        assert worst < 1e-2 and worst_s < 5e-2


def test_eval_forward_returns_logvar_through_the_c_abi():
    """``net(x, t, ..., return_logvar=True)`` in eval mode (models/swinv2.py:326-328): logvar = logvar_embed(c) from
    ``swb200_logvar_head`` against the oracle."""
    from oracle import swinv2_oracle as orc
    from swift_b200 import synthetic as syn
    from swift_b200.swinv2 import SwinV2
    cfg = syn.SWIFT_TINY
    m = SwinV2(img_resolution=cfg["img_resolution"], in_channels=cfg["in_channels"], out_channels=cfg["out_channels"],
               window_size=cfg["window_size"], shift_size=cfg["shift_size"], patch_size=cfg["patch_size"], depth=cfg["depth"],
               dim=cfg["dim"], heads=cfg["heads"], auxiliary_dim=cfg["auxiliary_dim"], logvar=True)
    sd = syn.random_state_dict(cfg, seed=1, logvar=True)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    x = torch.randn(3, cfg["in_channels"], *cfg["img_resolution"])
    t = torch.tensor([0.3, 0.9, 1.4])
    aux = torch.full((3, 1), 0.6)
    with torch.no_grad():
        y, lv = m(x.cuda(), t.cuda(), aux.cuda(), return_logvar=True)
    yo, lvo = orc.swinv2_forward(sd, orc.make_cfg(**cfg, logvar=True), x, t, aux, return_logvar=True)
    assert lv.shape == (3,) and _rel(lv.cpu(), lvo) < 1e-5
    assert _rel(y.cpu(), yo) < 5e-3


def test_scm_distillation_cotangent_vs_reference(golden):
    """The distillation branch (loss.py:205-210) on the CUDA path: the teacher is a second PassPrecond around swift_b200.SwinV2 in
    eval mode (seed-2 weights: the forecast kernels), ``swb200_scm_distill_direction`` swaps the tangent direction; loss and
    dL/dF_x against the REAL reference's ``SCMLoss(distillation=True)`` (tests/golden/scm_distill.npz)."""
    from swift_b200 import synthetic as syn
    from swift_b200.scm_target import scm_output_cotangent, latitude_weights, variable_weights
    from test_gpu_forward import build_net
    from test_oracle_golden import SCM_LOSS_VARIABLES
    g = golden("scm_distill")
    for cfgname, k in (("SWIFT_TINY", "tiny_"), ("SWIFT_SMALL", "small_")):
        cfg = getattr(syn, cfgname)
        n_img, (H, W) = cfg["out_channels"], cfg["img_resolution"]
        net, _ = build_net(cfg, img_channels=n_img)
        teacher, _ = build_net(cfg, seed=2, img_channels=n_img)
        x, cond = (v.cuda() for v in syn.synthetic_fields(cfg, 2, seed=5))
        step, warm = (int(v) for v in g[k + "step_warm"])
        with torch.no_grad():
            out = scm_output_cotangent(net, x, torch.from_numpy(g[k + "t"]).cuda(), torch.from_numpy(g[k + "z"]).cuda(), step,
                                       condition=cond, auxiliary=0.6, tangent_warmup_kimg=warm, w_lat=latitude_weights(H, "cuda"),
                                       w_var=variable_weights(SCM_LOSS_VARIABLES[:n_img], "cuda"), net_pretrained=teacher)
        e_cot = _rel(out["cot"].cpu(), torch.from_numpy(g[k + "cot"]))
        print(f"{k}: distillation cotangent rel-L2 {e_cot:.3e}, loss {float(out['loss']):.6f} vs {float(g[k + 'loss']):.6f}")
        assert e_cot < 3e-3
        assert abs(float(out["loss"]) - float(g[k + "loss"])) < 1e-3 * float(g[k + "loss"])
