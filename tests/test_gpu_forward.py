"""End-to-end parity of the CUDA denoiser / samplers with the reference (golden files) and the oracle.

Tolerance (BASELINE.json north_star): per-field relative L2 <= 1e-2 after one step, fields = output channels,
norms over (H, W).  Weights of the GEMM families are bf16-representable (swift_b200.synthetic), activations are
rounded to bf16 at GEMM inputs, everything else is fp32.
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-2


def per_field_rel_l2(y, ref):
    y, ref = torch.as_tensor(y).double().cpu(), torch.as_tensor(ref).double().cpu()
    num = (y - ref).flatten(2).norm(dim=-1)
    den = ref.flatten(2).norm(dim=-1).clamp_min(1e-30)
    return num / den          # [B, C]


def build_net(cfg, seed=1, img_channels=None, act_fp16=True, fuse_ln=None, x_single=None):
    from swift_b200 import synthetic as syn
    from swift_b200.precond import PassPrecond
    img_channels = cfg["out_channels"] if img_channels is None else img_channels
    model_cfg = dict(_target_="swift_b200.swinv2.SwinV2", window_size=cfg["window_size"],
                     shift_size=cfg["shift_size"], patch_size=cfg["patch_size"], depth=cfg["depth"], dim=cfg["dim"],
                     heads=cfg["heads"], logvar=False, timestep_weight=1.0)
    net = PassPrecond(model_cfg, img_resolution=cfg["img_resolution"], img_channels=img_channels,
                      condition_channels=cfg["in_channels"] - img_channels, auxiliary_dim=cfg["auxiliary_dim"],
                      sigma_min=0.0, sigma_max=float("inf"), sigma_data=1.0)
    sd = syn.random_state_dict(cfg, seed=seed, prefix="model.")
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    net.model.act_fp16 = act_fp16
    if fuse_ln is not None:
        net.model.fuse_ln = fuse_ln
    if x_single is not None:
        net.model.x_single = x_single
    return net.cuda().eval(), {k[len("model."):]: v for k, v in sd.items()}


@pytest.mark.parametrize("fuse_ln", [0, 1, 2, 3], ids=["ln_kernel", "ln_in_wo", "ln_in_w2", "ln_in_wo_w2"])
@pytest.mark.parametrize("act_fp16,x_single", [(True, True), (True, False), (False, False)],
                         ids=["act_fp16_x_single", "act_fp16_x_pair", "act_bf16"])
@pytest.mark.parametrize("name,cfgname", [("tiny", "SWIFT_TINY"), ("small", "SWIFT_SMALL")])
def test_module_forward_vs_reference_golden(golden, name, cfgname, act_fp16, x_single, fuse_ln):
    """fuse_ln: LayerNorm + modulation + residual add as its own kernel or in the wo / w2 GEMM epilogue; x_single: the
    residual stream as one fp16 value per element (default in fp16 mode) or as the [hi | lo] pair."""
    from swift_b200 import synthetic as syn
    g = golden(name)
    cfg = getattr(syn, cfgname)
    net, _ = build_net(cfg, act_fp16=act_fp16, fuse_ln=fuse_ln, x_single=x_single)
    lat, cond = syn.synthetic_fields(cfg, 2, seed=3)
    with torch.no_grad():
        y = net(lat.cuda(), torch.from_numpy(g["fwd_t"]).cuda(), cond.cuda(), torch.from_numpy(g["fwd_aux"]).cuda())
    err = per_field_rel_l2(y, g["fwd_y"])
    print(f"{name} act_fp16={act_fp16}: forward per-field rel-L2 max {err.max():.4e} mean {err.mean():.4e}")
    assert err.max() < (0.25 * TOL if act_fp16 else TOL)


@pytest.mark.parametrize("name,cfgname", [("tiny", "SWIFT_TINY"), ("small", "SWIFT_SMALL")])
def test_samplers_vs_reference_golden(golden, name, cfgname):
    from swift_b200 import synthetic as syn
    from swift_b200.sampler import DiffusionSampler
    g = golden(name)
    cfg = getattr(syn, cfgname)
    net, _ = build_net(cfg)
    lat, cond = syn.synthetic_fields(cfg, 2, seed=3)
    lat, cond = lat.cuda(), cond.cuda()
    z = torch.from_numpy(g["scm2_noise"]).cuda()
    S = DiffusionSampler(net)
    kw = dict(condition=cond, auxiliary=0.6, sigma_min=0.02, sigma_max=200.0)
    res = {
        "scm1": S.scm_solver(latents=lat, num_steps=1, **kw),
        "scm2": S.scm_solver(latents=lat, num_steps=2, randn_like=lambda x: z, **kw),
        "scm3": S.scm_solver(latents=lat, num_steps=3, randn_like=lambda x: z, **kw),
        "dpm2s_3": S.dpm_solver_2s(latents=lat, num_steps=3, **kw),
    }
    for k, v in res.items():
        err = per_field_rel_l2(v, g[k])
        print(f"{name}/{k}: per-field rel-L2 max {err.max():.4e}")
        # multi-step solvers chain 2..5 denoiser calls; the stated bar is for one step
        assert err.max() < (TOL if k == "scm1" else 3 * TOL), (k, err.max())


def test_fused_sampler_equals_generic_path():
    """sampler fast path (fused concat/update) == calling the module through PassPrecond.forward like the reference."""
    from swift_b200 import synthetic as syn
    from swift_b200.sampler import DiffusionSampler
    cfg = syn.SWIFT_TINY
    net, _ = build_net(cfg)
    lat, cond = syn.synthetic_fields(cfg, 2, seed=5)
    lat, cond = lat.cuda(), cond.cuda()
    fused = DiffusionSampler(net).scm_solver(latents=lat, condition=cond, auxiliary=0.6, num_steps=1,
                                              sigma_min=0.02, sigma_max=200.0)
    t = torch.tensor([math.pi / 2], device="cuda")
    with torch.no_grad():
        F = net(lat, t.expand(2), cond, 0.6)
    generic = torch.cos(t) * lat - torch.sin(t) * F
    assert torch.allclose(fused, generic, rtol=1e-5, atol=1e-6)


def test_swift_b_one_step_vs_oracle_and_golden(golden):
    """BASELINE.json configs[0]: Swift-B, single 6 h sCM step, batch 1, 128x256, vs fp32 reference.

    Default numerics (fp16 tensor-core operands, fp32 accumulate) must meet the 1e-2 per-field bar with margin; the all-bf16
    mode is measured too and reported (SURVEY.md section 7.3 predicts ~0.8e-2 mean / ~1e-2 max for it)."""
    from oracle import swinv2_oracle as orc
    from swift_b200 import synthetic as syn
    from swift_b200.sampler import DiffusionSampler
    cfg = syn.SWIFT_B
    net, sd = build_net(cfg, img_channels=syn.IMG_CHANNELS)
    lat, cond = syn.synthetic_fields(cfg, 1, seed=0)
    y = DiffusionSampler(net).scm_solver(latents=lat.cuda(), condition=cond.cuda(), auxiliary=0.6, num_steps=1,
                                         sigma_min=0.02, sigma_max=200.0)
    torch.cuda.synchronize()
    # oracle in fp32 on the GPU (TF32 off) for speed; pinned to the reference by the digest below
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ocfg = orc.make_cfg(**cfg)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = orc.scm_solver(lambda x, t, c, a: orc.pass_precond(sd_gpu, ocfg, x, t, c, a), lat.cuda(), cond.cuda(),
                             0.6, num_steps=1)
    g = golden("swift_b")
    assert np.allclose(ref[:, :, ::8, ::8].cpu().numpy(), g["scm1_sub"], rtol=2e-3, atol=2e-4), \
        "GPU fp32 oracle drifted from the reference digest"
    err = per_field_rel_l2(y, ref)
    print(f"swift_b scm1 (fp16 operands, single-value residual stream = the default): per-field rel-L2 max {err.max():.4e} "
          f"mean {err.mean():.4e}")
    assert err.max() < 0.3 * TOL
    sub = per_field_rel_l2(y[:, :, ::8, ::8], torch.from_numpy(g["scm1_sub"]))
    assert sub.max() < TOL                  # vs the REAL reference's sub-sampled output (512 points per field)
    # the [hi | lo] residual pair (x_single = False): 0.5e-3 closer, 40 % more residual traffic
    net.model.x_single = False
    for fuse_ln in (2, 3):
        net.model.fuse_ln = fuse_ln
        yp = DiffusionSampler(net).scm_solver(latents=lat.cuda(), condition=cond.cuda(), auxiliary=0.6, num_steps=1,
                                              sigma_min=0.02, sigma_max=200.0)
        errp = per_field_rel_l2(yp, ref)
        print(f"swift_b scm1 (fp16 operands, [hi | lo] residual pair, fuse_ln {fuse_ln}): per-field rel-L2 max {errp.max():.4e} "
              f"mean {errp.mean():.4e}")
        assert errp.max() < 0.2 * TOL
    net.model.x_single = True
    # the bf16 operand format north_star names: GEMM operands bf16, attention internals (q, k, v, P) fp16 -- must meet the
    # same bar; with bf16 attention internals as well (attn_fp16 = False) it is borderline, which is reported
    net.model.act_fp16 = False
    for attn_fp16, bar in ((True, TOL), (False, 1.5 * TOL)):
        net.model.attn_fp16 = attn_fp16
        yb = DiffusionSampler(net).scm_solver(latents=lat.cuda(), condition=cond.cuda(), auxiliary=0.6, num_steps=1,
                                              sigma_min=0.02, sigma_max=200.0)
        errb = per_field_rel_l2(yb, ref)
        print(f"swift_b scm1 (bf16 GEMM operands, {'fp16' if attn_fp16 else 'bf16'} attention internals): per-field "
              f"rel-L2 max {errb.max():.4e} mean {errb.mean():.4e}")
        assert errb.max() < bar


def test_swift_b_trigflow_2s_vs_oracle():
    """BASELINE.json configs[3] (TrigFlow 2S diffusion baseline) at Swift-B scale, 3 Heun steps = 5 denoiser calls with
    time-dependent conditioning, vs the fp32 oracle; multi-call drift is reported."""
    from oracle import swinv2_oracle as orc
    from swift_b200 import synthetic as syn
    from swift_b200.sampler import DiffusionSampler
    cfg = syn.SWIFT_B
    net, sd = build_net(cfg, img_channels=syn.IMG_CHANNELS)
    lat, cond = syn.synthetic_fields(cfg, 1, seed=2)
    y = DiffusionSampler(net).dpm_solver_2s(latents=lat.cuda(), condition=cond.cuda(), auxiliary=0.6, num_steps=3,
                                            sigma_min=0.02, sigma_max=200.0)
    torch.backends.cuda.matmul.allow_tf32 = False
    ocfg = orc.make_cfg(**cfg)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = orc.dpm_solver_2s(lambda x, t, c, a: orc.pass_precond(sd_gpu, ocfg, x, t, c, a), lat.cuda(), cond.cuda(),
                                0.6, num_steps=3)
    err = per_field_rel_l2(y, ref)
    print(f"swift_b 2S (5 calls): per-field rel-L2 max {err.max():.4e} mean {err.mean():.4e}")
    assert err.max() < TOL


def test_module_contract():
    from swift_b200 import synthetic as syn
    from swift_b200.swinv2 import SwinV2
    cfg = syn.SWIFT_TINY
    m = SwinV2(**cfg)
    ref_shapes = syn.state_dict_shapes(cfg)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == ref_shapes
    # default init zeroes modulation/head like the reference -> output identically 0
    m = m.cuda().eval()
    x = torch.randn(1, cfg["in_channels"], 32, 64, device="cuda")
    with torch.no_grad():
        y = m(x, torch.tensor(0.5, device="cuda"), torch.tensor([[0.6]], device="cuda"))
    assert y.shape == (1, cfg["out_channels"], 32, 64) and float(y.abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        m.cpu()(x.cpu(), torch.tensor(0.5))
    with torch.no_grad():                                   # jvp=True alone is the plain forward (explicit-softmax flag)
        assert torch.equal(m.cuda()(x, torch.tensor(0.5, device="cuda"), torch.tensor([[0.6]], device="cuda"), jvp=True), y)
    with pytest.raises(NotImplementedError):                # the jvp=True node is forward-mode only (reverse mode: .train())
        xg = x.clone().requires_grad_(True)
        m(xg, torch.tensor(0.5, device="cuda"), torch.tensor([[0.6]], device="cuda"), jvp=True).sum().backward()
    with pytest.raises(NotImplementedError):
        SwinV2(**{**cfg, "window_size": [8, 8]})
    m.train()                                               # grad-enabled training mode = the reverse-mode path
    y_t = m(x, torch.tensor(0.5, device="cuda"), torch.tensor([[0.6]], device="cuda"))
    assert y_t.requires_grad and y_t.shape == y.shape
    with pytest.raises(NotImplementedError):                # parameter gradients only
        m(x.clone().requires_grad_(True), torch.tensor(0.5, device="cuda"))


def test_chunked_batch_matches_single():
    from swift_b200 import synthetic as syn
    cfg = syn.SWIFT_TINY
    net, _ = build_net(cfg)
    lat, cond = syn.synthetic_fields(cfg, 5, seed=9)
    t = torch.full((5,), 0.9, device="cuda")
    with torch.no_grad():
        net.model.max_chunk = 8
        y_all = net(lat.cuda(), t, cond.cuda(), 0.6).clone()
        net.model.max_chunk = 2
        y_chunk = net(lat.cuda(), t, cond.cuda(), 0.6)
    assert torch.equal(y_all, y_chunk)


def test_fp16_range_stress_and_saturation_counters():
    """fp16 operands (the default) have a range of +-65504; every conversion saturates there instead of producing inf
    (DESIGN.md numerics contract).  A checkpoint with outlier channels is where fp16 and bf16 part ways, so the behaviour
    at and beyond the limit is pinned here on fixtures whose wo branch and SwiGLU hidden are scaled towards it
    (LayerNorm follows both, so the fp32 oracle's output hardly moves with the scale):
      * below the limit (|branch|, |h| up to ~1e4): no saturation is counted and the 1e-2 bar holds;
      * beyond it: the output stays finite, ``Engine.saturation_counts()`` reports the clipped elements per tensor class
        (the documented signal to run that checkpoint with ``act_fp16 = False``), and the bf16 operand mode -- same
        weights -- still meets the bar."""
    from oracle import swinv2_oracle as orc
    from swift_b200 import synthetic as syn
    cfg = syn.SWIFT_SMALL
    net, sd = build_net(cfg, fuse_ln=0)                      # un-fused LayerNorm: both branch outputs pass through fp16
    lat, cond = syn.synthetic_fields(cfg, 1, seed=4)
    x = torch.cat([lat, cond], 1).cuda()
    t = torch.tensor([1.1], device="cuda")
    aux = torch.tensor([[0.6]], device="cuda")
    ocfg = orc.make_cfg(**cfg)
    torch.backends.cuda.matmul.allow_tf32 = False

    hd = cfg["dim"] // cfg["heads"]

    def run(k_wo, k_w1, act_fp16):
        """wo branch scaled by 2^k_wo (split between the v rows of to_qkv and wo, so the WEIGHTS stay far inside the fp16
        range), SwiGLU hidden by 4^k_w1 (w1 x 2^k_w1); powers of two keep the weights bf16-representable."""
        sd2 = dict(sd)
        for k in sd:
            if k.endswith(".0.wo.weight"):
                sd2[k] = sd[k] * 2.0 ** (k_wo - k_wo // 2)
            elif k.endswith(".0.to_qkv.weight"):
                w = sd[k].clone().reshape(cfg["heads"], 3, hd, -1)       # rows: (head, q|k|v, d) (swinv2.py:120-121)
                w[:, 2] *= 2.0 ** (k_wo // 2)
                sd2[k] = w.reshape(sd[k].shape)
            elif k.endswith(".1.w1.weight"):
                sd2[k] = sd[k] * 2.0 ** k_w1
        net.load_state_dict({"model." + k: v for k, v in sd2.items()}, strict=True)
        net.model.act_fp16 = act_fp16
        eng = net.model.engine()
        eng.count_saturation(True)
        try:
            with torch.no_grad():
                y = net.model(x, t, aux)
            counts = eng.saturation_counts()
        finally:
            eng.count_saturation(False)
        with torch.no_grad():
            ref = orc.swinv2_forward({k: v.cuda() for k, v in sd2.items()}, ocfg, x, t, aux)
        return y, ref, counts

    y, ref, c = run(0, 0, True)
    assert sum(c.values()) == 0, c
    assert per_field_rel_l2(y, ref).max() < TOL
    # find the scales at which the branch / the hidden start to clip
    k_wo = next(k for k in range(6, 30) if run(k, 0, True)[2]["branch"] > 0)
    k_w1 = next(k for k in range(2, 14) if run(0, k, True)[2]["h"] > 0)
    print(f"fp16 range: wo branch clips from x2^{k_wo}, SwiGLU hidden from w1 x2^{k_w1} (x4^{k_w1})")
    # two octaves below the first clipped element: in range, parity holds
    y, ref, c = run(k_wo - 2, k_w1 - 1, True)
    e_near = per_field_rel_l2(y, ref).max().item()
    assert sum(c.values()) == 0, c
    assert e_near < TOL, e_near
    # two octaves beyond: clipped, counted, finite; bf16 operands keep the bar on the same weights
    y, ref, c = run(k_wo + 2, k_w1 + 1, True)
    e_far = per_field_rel_l2(y, ref).max().item()
    assert torch.isfinite(y).all(), "saturating conversions must never produce inf / NaN"
    assert c["branch"] > 0 and c["h"] > 0, c
    yb, refb, cb = run(k_wo + 2, k_w1 + 1, False)
    e_bf16 = per_field_rel_l2(yb, refb).max().item()
    print(f"fp16 range: near-limit err {e_near:.3e} (counts 0); beyond: fp16 err {e_far:.3e} with {c}; bf16 err {e_bf16:.3e}")
    assert cb["qkv"] == 0, cb                                 # bf16 mode: only its fp16 attention internals are scanned
    assert e_bf16 < TOL, e_bf16
    assert torch.isfinite(yb).all()
