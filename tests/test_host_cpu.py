"""CPU-only checks of the host side: the C-ABI library loads and exports what include/swift_b200.h declares,
the packed weight layouts + window-gather index arithmetic reproduce the oracle when emulated in PyTorch, the
sampler / precond mirrors keep the reference semantics, and the config loader composes the experiment files.
No compute call into the CUDA library is made here."""
import ctypes
import math
import os
import re

import pytest
import torch

from oracle import swinv2_oracle as orc
from swift_b200 import _lib, packing, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HD = 88


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    from swift_b200 import build
    build.build()                                      # nvcc cross-compiles without a GPU
    header = open(os.path.join(ROOT, "include", "swift_b200.h")).read()
    declared = set(re.findall(r"SWB200_API\s+[\w\s\*]+?\b(swb200_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.swb200_abi_version.restype = ctypes.c_int
    assert lib.swb200_abi_version() == _lib.ABI_VERSION


def test_validate_rejects_unsupported_geometry():
    lib = _lib.lib()
    m = _lib.Model()
    m.img_h, m.img_w, m.patch_h, m.patch_w = 128, 256, 2, 2
    m.win_h = m.win_w = 8
    assert lib.swb200_validate(ctypes.byref(m)) != 0
    assert b"16x16 windows" in lib.swb200_last_error()
    m.win_h = m.win_w = 16
    m.dim, m.heads = 1024, 16
    assert lib.swb200_validate(ctypes.byref(m)) != 0
    assert b"head_dim 88" in lib.swb200_last_error()
    m.dim, m.heads, m.dff, m.depth = 1056, 12, 2816, 12
    m.in_channels, m.out_channels, m.k_embed = 141, 69, 568
    m.shift_h = m.shift_w = 8
    m.gemm_tile = 3
    assert lib.swb200_validate(ctypes.byref(m)) == 0
    assert 150e6 < lib.swb200_workspace_bytes(ctypes.byref(m), 1) < 250e6      # ~190 MB per Swift-B sample


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libswift_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


# ------------------------------------------------------------------------------------------------ packing
def _geometry(c):
    return packing.Geometry(img=tuple(c["img_resolution"]), patch=tuple(c["patch_size"]),
                            window=tuple(c["window_size"]), shift=tuple(c["shift_size"]),
                            in_channels=c["in_channels"], out_channels=c["out_channels"], depth=c["depth"],
                            dim=c["dim"], heads=c["heads"], aux_dim=c["auxiliary_dim"], timestep_weight=1.0)


def _emulate_packed_forward(keep, g, x, cond_vec_fn, gemm_tile):
    """What the CUDA kernels compute, written with PyTorch ops on the PACKED tensors (fp32 math): validates the
    layouts of packing.py and the gather-instead-of-roll window indexing of attention.cu."""
    B = x.shape[0]
    D, H, L, Dff, T = g.dim, g.heads, g.depth, g.dff, g.tokens
    gh, gw = g.grid
    p1, p2 = g.patch
    pp = p1 * p2
    # patch gather: k = c*pp + py*p2 + px
    A = x.reshape(B, g.in_channels, gh, p1, gw, p2).permute(0, 2, 4, 1, 3, 5).reshape(B * T, g.in_channels * pp)
    A = torch.nn.functional.pad(A, (0, g.k_embed - A.shape[1]))
    w_embed = keep["w_embed"].float()[:, :g.k_embed]
    tok = A @ w_embed.t() + keep["pos_embed"].repeat(B, 1)           # the bias is folded into the position table
    gain, bias = cond_vec_fn()
    for l in range(L):
        qkv = (tok @ keep["w_qkv"][l].float().t()).reshape(B * T, 3, H, HD)
        q = torch.nn.functional.normalize(qkv[:, 0], dim=-1) * keep["qscale"][l][None, :, None]
        k = torch.nn.functional.normalize(qkv[:, 1], dim=-1)
        v = qkv[:, 2]
        s = g.shift if (l % 2 == 1) else (0, 0)
        attn = torch.empty(B * T, H, HD)
        for b in range(B):
            for wy in range(gh // 16):
                for wx in range(gw // 16):
                    iy, ix = torch.meshgrid(torch.arange(16), torch.arange(16), indexing="ij")
                    rows = (b * gh + (wy * 16 + iy + s[0]) % gh) * gw + (wx * 16 + ix + s[1]) % gw
                    rows = rows.reshape(-1)
                    qq, kk, vv = q[rows].transpose(0, 1), k[rows].transpose(0, 1), v[rows].transpose(0, 1)
                    attn[rows] = (torch.softmax(qq @ kk.transpose(-1, -2), -1) @ vv).transpose(0, 1)
        branch = attn.reshape(B * T, D) @ keep["w_o"][l].float().t()
        ln = torch.nn.functional.layer_norm(branch, (D,), eps=1e-6).reshape(B, T, D)
        tok = tok + (ln * gain[2 * l][:, None] + bias[2 * l][:, None]).reshape(B * T, D)
        half = HD * (2 if gemm_tile == 3 else 1)          # w1 is packed per GEMM tile: [half gate | half up]
        h = (tok @ keep["w_1"][l].float().t()).reshape(B * T, Dff // half, 2, half)
        h = (torch.nn.functional.silu(h[:, :, 0]) * h[:, :, 1]).reshape(B * T, Dff)
        branch = h @ keep["w_2"][l].float().t()
        ln = torch.nn.functional.layer_norm(branch, (D,), eps=1e-6).reshape(B, T, D)
        tok = tok + (ln * gain[2 * l + 1][:, None] + bias[2 * l + 1][:, None]).reshape(B * T, D)
    y = tok @ keep["w_head"].float()[:, :D].t()
    return y.reshape(B, gh, gw, g.out_channels, p1, p2).permute(0, 3, 1, 4, 2, 5).reshape(B, g.out_channels, *g.img)


@pytest.mark.parametrize("gemm_tile", [3, 2])
@pytest.mark.parametrize("cfgname", ["SWIFT_TINY", "SWIFT_SMALL"])
def test_packed_layouts_reproduce_oracle(cfgname, gemm_tile):
    c = getattr(syn, cfgname)
    g = _geometry(c)
    sd = syn.random_state_dict(c, seed=1)
    # the layout oracle (oracle/pack_oracle.py) restates what swb200_pack_weights writes; the GPU suite checks the CUDA
    # packer against it byte for byte, this test checks that the layouts MEAN what include/swift_b200.h says
    from oracle import pack_oracle
    keep = pack_oracle.pack_layouts(sd, g, torch.device("cpu"), gemm_tile=gemm_tile)
    lat, cond = syn.synthetic_fields(c, 2, seed=3)
    x = torch.cat([lat, cond], 1)
    t = torch.tensor([0.3, 1.2])
    aux = torch.tensor([[0.6], [1.2]])

    def cond_vecs():
        cvec = orc.conditioning_vector(sd, t, aux, g.dim, 1, 1.0, 2)
        mod = (cvec @ keep["mod_w"].t() + keep["mod_b"]).reshape(2, 2 * g.depth, 2, g.dim)
        gain = keep["ln_gamma"][None] * (1 + mod[:, :, 0])
        bias = keep["ln_beta"][None] * (1 + mod[:, :, 0]) + mod[:, :, 1]
        return gain.transpose(0, 1), bias.transpose(0, 1)

    y = _emulate_packed_forward(keep, g, x, cond_vecs, gemm_tile)
    ref = orc.swinv2_forward(sd, orc.make_cfg(**c), x, t, aux)
    err = (y - ref).norm() / ref.norm()
    assert err < 2e-5, err


def test_pack_rejects_unsupported_shapes():
    c = dict(syn.SWIFT_TINY)
    with pytest.raises(NotImplementedError, match="16x16"):
        packing.check_supported(_geometry({**c, "window_size": [8, 8]}))
    with pytest.raises(NotImplementedError, match="head_dim"):
        packing.check_supported(_geometry({**c, "dim": 256, "heads": 4}))
    with pytest.raises(TypeError):
        packing._pair("16")


# ------------------------------------------------------------------------------------------------ module / sampler mirrors
def test_module_state_dict_schema_and_cpu_refusal():
    from swift_b200.swinv2 import SwinV2
    c = syn.SWIFT_TINY
    m = SwinV2(**c)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == syn.state_dict_shapes(c)
    sd = syn.random_state_dict(c, seed=2)
    assert m.load_state_dict(sd, strict=True).missing_keys == []
    m.eval()
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.zeros(1, c["in_channels"], 32, 64), torch.tensor(0.5))
    # the logvar variant carries the extra reference keys
    m2 = SwinV2(**c, logvar=True)
    assert "logvar_embed.weight" in m2.state_dict()


class _OraclePrecond(torch.nn.Module):
    """A stand-in `net` (not PassPrecond-named) so the samplers take the generic, reference-identical path."""

    def __init__(self, sd, cfg):
        super().__init__()
        self.sd, self.cfg = sd, cfg
        self.sigma_data, self.img_channels, self.img_resolution = 1.0, cfg["out_channels"], cfg["res"]
        self.auxiliary_dim = cfg["auxiliary_dim"]

    def forward(self, x, t, condition=None, auxiliary=None):
        return orc.pass_precond(self.sd, self.cfg, x, t, condition, auxiliary)


def test_sampler_generic_path_matches_reference_golden(golden):
    from swift_b200.sampler import DiffusionSampler, sampler_factory
    g = golden("tiny")
    c = syn.SWIFT_TINY
    sd = syn.random_state_dict(c, seed=1)
    net = _OraclePrecond(sd, orc.make_cfg(**c))
    lat, cond = syn.synthetic_fields(c, 2, seed=3)
    z = torch.from_numpy(g["scm2_noise"])
    S = DiffusionSampler(net)
    kw = dict(condition=cond, auxiliary=0.6, sigma_min=0.02, sigma_max=200.0)
    assert torch.allclose(S.scm_solver(latents=lat, num_steps=1, **kw), torch.from_numpy(g["scm1"]), atol=1e-4)
    assert torch.allclose(S.scm_solver(latents=lat, num_steps=2, randn_like=lambda x: z, **kw),
                          torch.from_numpy(g["scm2"]), atol=1e-4)
    assert torch.allclose(S.dpm_solver_2s(latents=lat, num_steps=3, **kw), torch.from_numpy(g["dpm2s_3"]), atol=2e-4)
    # factory closure: latents drawn with the caller's generator, like generating/factory.py:52-56
    smp = sampler_factory("scm", net, num_steps=1, sigma_min=0.02, sigma_max=200.0, auxiliary=0.6)
    gen = torch.Generator().manual_seed(4)
    y = smp(cond, generator=gen)
    lat4 = torch.randn((2, c["out_channels"], 32, 64), generator=torch.Generator().manual_seed(4))
    assert torch.allclose(y, orc.scm_solver(net, lat4, cond, 0.6, num_steps=1), atol=1e-5)
    with pytest.raises(ValueError, match="Unknown solver mode"):
        sampler_factory("nope", net)


def test_precond_mirror_and_config_loader():
    from swift_b200.config import instantiate, load_experiment
    cfg = load_experiment("era5-swinv2-1.4-scm")
    assert cfg["model"]["dim"] == 1056 and cfg["model"]["_target_"] == "swift_b200.swinv2.SwinV2"
    assert cfg["solver"] == {"num_steps": 1, "sigma_min": 0.02, "sigma_max": 200, "auxiliary": 0.6}
    assert load_experiment("era5-swinv2-1.4-trigflow")["solver"]["num_steps"] == 20
    # instantiate a (small) model through the same precond -> model._target_ route as generate.py:197-206
    small = dict(cfg["model"], depth=2, dim=264, heads=3)
    net = instantiate(cfg["precond"], model_config=small, img_resolution=[32, 64], img_channels=5,
                      condition_channels=8, sigma_max=float("inf"), _recursive_=False, _convert_="object")
    assert type(net).__name__ == "PassPrecond" and type(net.model).__name__ == "SwinV2"
    assert net.model.in_channels == 13 and net.img_channels == 5 and net.sigma_data == 1.0
    keys = net.state_dict().keys()
    assert all(k.startswith("model.") for k in keys) and "model.head.head.0.weight" in keys
    from swift_b200.precond import process_auxiliary
    a = process_auxiliary(0.6, 1, 4, "cpu")
    assert a.shape == (4, 1) and torch.allclose(a, torch.full((4, 1), 0.6))
    assert process_auxiliary(None, 1, 4, "cpu").shape == (1, 1) and process_auxiliary(0.6, 0, 4, "cpu") is None


def test_reference_noise_bookkeeping():
    """rollout.ReferenceNoise: which reference call (member, IC batch) each local trajectory belongs to, ragged last
    batch, argument checks (the Philox-offset replay itself needs a CUDA generator: tests/test_gpu_rollout.py)."""
    import torch
    from swift_b200.rollout import ReferenceNoise
    traj = [(1, 4), (0, 0), (1, 1), (0, 3), (1, 0)]
    rn = ReferenceNoise(traj, n_ic=5, batch=2, steps=3, sample_shape=(2, 4, 4), device=torch.device("cpu"))
    assert rn.groups == {(1, 2): [(0, 0)], (0, 0): [(1, 0)], (1, 0): [(2, 1), (4, 0)], (0, 1): [(3, 1)]}
    assert rn.sizes == {2: 1, 0: 2, 1: 2}                       # the last reference batch holds one IC
    with pytest.raises(ValueError):
        ReferenceNoise([(0, 5)], n_ic=5, batch=2, steps=3, sample_shape=(1,), device=torch.device("cpu"))
    with pytest.raises(ValueError):
        ReferenceNoise(traj, n_ic=5, batch=0, steps=3, sample_shape=(1,), device=torch.device("cpu"))
    with pytest.raises(ValueError):
        rn.fill(torch.empty(5, 2, 4, 4), 3)


def test_scm_loss_weights_match_reference_golden(golden):
    """swift_b200.scm_target weights (training/loss.py:28-57) against the buffers of the real SCMLoss."""
    import numpy as np
    from swift_b200.scm_target import latitude_weights, variable_weights
    from swift_b200.generate import era5_variables
    g = golden("scm_loss")
    names = ["2m_temperature", "10m_u_component_of_wind", "mean_sea_level_pressure", "geopotential_500",
             "temperature_850", "specific_humidity_700"]
    np.testing.assert_array_equal(latitude_weights(32).numpy(), g["tiny_w_lat"])
    np.testing.assert_array_equal(latitude_weights(64).numpy(), g["small_w_lat"])
    np.testing.assert_array_equal(variable_weights(names[:5]).numpy(), g["tiny_w_var"])
    np.testing.assert_array_equal(variable_weights(names[:4]).numpy(), g["small_w_var"])
    w = variable_weights(era5_variables())                      # the 69 Swift-B variables all have a weight
    assert w.shape == (1, 69, 1, 1) and abs(float(w.sum()) - 1.0) < 1e-6 and float(w.min()) > 0
    with pytest.raises(KeyError):
        variable_weights(["geopotential_475"])
    with pytest.raises(TypeError):
        from swift_b200.scm_target import scm_output_cotangent
        import torch

        class _Net(torch.nn.Module):
            sigma_data, model = 1.0, torch.nn.Identity()
        scm_output_cotangent(_Net(), torch.zeros(1, 1, 2, 2), torch.zeros(1), torch.zeros(1, 1, 2, 2), 0)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm): one JSON line on stdout with the contract's keys, timed on
    the host cores with the oracle port; under torchrun only rank 0 prints."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""               # ranks > 0 exit 0 without work
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "forecast member-steps/sec" and d["unit"] == "member-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                                            "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"] and "model" not in d["config"]
