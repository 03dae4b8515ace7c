"""Glue kernels and window attention through the C ABI vs PyTorch / oracle references."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

HD, HDP = 88, 96


@pytest.fixture(scope="module")
def lib():
    from swift_b200 import _lib
    return _lib.lib()


def _check(rc):
    from swift_b200 import _lib
    _lib.check(rc, "test call")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


ACT = pytest.mark.parametrize("f16", [1, 0], ids=["act_fp16", "act_bf16"])


def _adt(f16):
    return torch.float16 if f16 else torch.bfloat16


@ACT
@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("p1,p2,C0,C1,H,W", [(2, 2, 5, 8, 32, 64), (2, 2, 69, 72, 128, 256), (1, 1, 3, 4, 16, 32),
                                             (2, 4, 3, 0, 16, 64)])
def test_patch_gather(lib, split, p1, p2, C0, C1, H, W, f16):
    from swift_b200 import _lib
    B = 2
    x0 = torch.randn(B, C0, H, W, device="cuda")
    x1 = torch.randn(B, C1, H, W, device="cuda") if C1 else None
    m = _lib.Model()
    m.img_h, m.img_w, m.patch_h, m.patch_w, m.in_channels = H, W, p1, p2, C0 + C1
    Cin = C0 + C1
    kp = (Cin * p1 * p2 + 7) // 8 * 8
    m.k_embed, m.split_embed, m.act_fp16 = kp, split, f16
    dt = _adt(f16)
    lda = kp * (1 + split)
    T = (H // p1) * (W // p2)
    A = torch.full((B * T, lda), float("nan"), device="cuda", dtype=dt)
    scale0 = 0.5
    _check(lib.swb200_patch_gather(C.byref(m), x0.data_ptr(), C0, scale0, _lib.ptr(x1), C1, B, A.data_ptr(), lda,
                                   _stream()))
    torch.cuda.synchronize()
    full = torch.cat([x0 * scale0] + ([x1] if C1 else []), 1)
    # ours: k = c*pp + py*p2 + px
    ref = full.reshape(B, Cin, H // p1, p1, W // p2, p2).permute(0, 2, 4, 1, 3, 5).reshape(B * T, Cin * p1 * p2)
    hi = ref.to(dt)
    assert torch.equal(A[:, :Cin * p1 * p2], hi)
    assert (A[:, Cin * p1 * p2:kp] == 0).all()
    if split:
        lo = (ref - hi.float()).to(dt)
        assert torch.equal(A[:, kp:kp + Cin * p1 * p2], lo)
        assert _rel(A[:, :kp].float()[:, :Cin * p1 * p2] + A[:, kp:].float()[:, :Cin * p1 * p2], ref) < 2e-5 * (0.2 if f16 else 1)


@pytest.mark.parametrize("br16", [0, 1], ids=["branch_fp32", "branch_16bit"])
@pytest.mark.parametrize("D", [264, 1056])
def test_ln_mod_residual_single_value_stream(lib, D, br16):
    """Format word 3 (fp16 operands + single-value residual stream): x is the hi half alone, rounded once per update; the lo
    half of the buffer is neither read nor written."""
    B, T = 3, 256
    M = B * T
    branch = torch.randn(M, D, device="cuda") * 3 + 0.7
    if br16:
        branch = branch.half()
    hi = torch.randn(M, D, device="cuda").half()
    poison = torch.full((M, D), float("nan"), device="cuda", dtype=torch.float16)
    xhl = torch.cat([hi, poison], 1).contiguous()
    gain = torch.randn(B, D, device="cuda")
    bias = torch.randn(B, D, device="cuda")
    x_ref = hi.float() + torch.nn.functional.layer_norm(branch.float(), (D,), eps=1e-6).reshape(B, T, D).mul(
        gain[:, None]).add(bias[:, None]).reshape(M, D)
    _check(lib.swb200_ln_mod_residual(branch.data_ptr(), br16, xhl.data_ptr(), gain.data_ptr(), bias.data_ptr(), M, D, T,
                                      3, _stream()))
    torch.cuda.synchronize()
    assert torch.isnan(xhl[:, D:]).all()                               # lo untouched
    got = xhl[:, :D].float()
    assert _rel(got, x_ref) < 6e-4, f"{_rel(got, x_ref):.3e}"
    # one correctly rounded fp16 value per element (up to the summation order of the statistics)
    assert (got - x_ref.half().float()).abs().max() <= 2 * torch.finfo(torch.float16).eps * x_ref.abs().max()
    rc = lib.swb200_ln_mod_residual(branch.data_ptr(), br16, xhl.data_ptr(), gain.data_ptr(), bias.data_ptr(), M, D, T, 2, _stream())
    assert rc != 0 and b"fp16" in lib.swb200_last_error()              # bf16 keeps the pair


@ACT
@pytest.mark.parametrize("br16", [0, 1], ids=["branch_fp32", "branch_16bit"])
@pytest.mark.parametrize("D", [264, 528, 1056])
def test_ln_mod_residual(lib, D, f16, br16):
    """x += LN(branch)*gain + bias on the residual pair [hi | lo]."""
    dt = _adt(f16)
    B, T = 3, 256
    M = B * T
    branch = torch.randn(M, D, device="cuda") * 3 + 0.7
    if br16:
        branch = branch.to(dt)
    x = torch.randn(M, D, device="cuda")
    hi = x.to(dt)
    lo = (x - hi.float()).to(dt)
    x0 = hi.float() + lo.float()                       # what the pair actually represents
    xhl = torch.cat([hi, lo], 1).contiguous()
    gain = torch.randn(B, D, device="cuda")
    bias = torch.randn(B, D, device="cuda")
    x_ref = x0 + torch.nn.functional.layer_norm(branch.float(), (D,), eps=1e-6).reshape(B, T, D).mul(gain[:, None]).add(
        bias[:, None]).reshape(M, D)
    _check(lib.swb200_ln_mod_residual(branch.data_ptr(), br16, xhl.data_ptr(), gain.data_ptr(), bias.data_ptr(), M, D, T,
                                      f16, _stream()))
    torch.cuda.synchronize()
    got = xhl[:, :D].float() + xhl[:, D:].float()
    assert _rel(got, x_ref) < (3e-6 if f16 else 3e-5), f"{_rel(got, x_ref):.3e}"
    assert _rel(xhl[:, :D].float(), x_ref) < (6e-4 if f16 else 5e-3)


def _window_attention_ref(qkv, B, gh, gw, H, shift):
    """Reference on the same packed bf16 q/k/v: roll, partition, softmax(q k^T) v, reverse, roll back."""
    M = B * gh * gw
    q, k, v = [qkv[i, :, :, :HD].float().reshape(H, B, gh, gw, HD) for i in range(3)]

    def win(t):
        t = torch.roll(t, shifts=(-shift[0], -shift[1]), dims=(2, 3))
        t = t.reshape(H, B, gh // 16, 16, gw // 16, 16, HD).permute(0, 1, 2, 4, 3, 5, 6)
        return t.reshape(H, B, (gh // 16) * (gw // 16), 256, HD)

    qw, kw, vw = win(q), win(k), win(v)
    o = torch.softmax(qw @ kw.transpose(-1, -2), dim=-1) @ vw
    o = o.reshape(H, B, gh // 16, gw // 16, 16, 16, HD).permute(0, 1, 2, 4, 3, 5, 6).reshape(H, B, gh, gw, HD)
    o = torch.roll(o, shifts=(shift[0], shift[1]), dims=(2, 3))
    return o.permute(1, 2, 3, 0, 4).reshape(M, H * HD)


@pytest.mark.parametrize("f16,o16", [(1, 1), (0, 0), (1, 0)], ids=["fp16", "bf16", "fp16qkv-bf16out"])
@pytest.mark.parametrize("impl", [2, 1], ids=["tcgen05", "mma_sync"])
@pytest.mark.parametrize("shift", [(0, 0), (8, 8), (8, 0), (3, 5)])
@pytest.mark.parametrize("B,gh,gw,H", [(1, 16, 32, 3), (2, 32, 32, 2), (1, 64, 128, 12)])
def test_window_attention(lib, shift, B, gh, gw, H, f16, o16, impl):
    if impl == 2 and (shift[0] % 8 or shift[1] % 8):
        pytest.skip("the tcgen05 kernel handles shifts that are multiples of 8 (Swift uses 8)")
    M = B * gh * gw
    g = torch.Generator(device="cuda").manual_seed(5)
    raw = torch.randn(3, H, M, HDP, generator=g, device="cuda")
    raw[..., HD:] = 0
    qs = torch.linspace(4.0, 30.0, H, device="cuda")[:, None, None]
    raw[0] = torch.nn.functional.normalize(raw[0], dim=-1) * qs
    raw[1] = torch.nn.functional.normalize(raw[1], dim=-1)
    qkv = raw.to(_adt(f16)).contiguous()
    out = torch.full((M, H * HD), float("nan"), device="cuda", dtype=_adt(o16))
    _check(lib.swb200_window_attention(qkv.data_ptr(), out.data_ptr(), B, gh, gw, H, shift[0], shift[1], f16, o16, impl,
                                       None, _stream()))
    torch.cuda.synchronize()
    ref = _window_attention_ref(qkv, B, gh, gw, H, shift)
    assert torch.isfinite(out.float()).all()
    tol = 1e-3 if o16 else (3.5e-3 if f16 else 8e-3)      # P / output rounding; fp16 P with a bf16 output: the output's only
    assert _rel(out.float(), ref) < tol, f"{_rel(out.float(), ref):.3e}"


def test_conditioning_matches_oracle(lib):
    from oracle import swinv2_oracle as orc
    from swift_b200 import packing, synthetic as syn
    from swift_b200.engine import Engine
    c = syn.SWIFT_SMALL
    sd = syn.random_state_dict(c, seed=1)
    g = packing.Geometry(img=(64, 64), patch=(2, 2), window=(16, 16), shift=(8, 8), in_channels=c["in_channels"],
                         out_channels=c["out_channels"], depth=c["depth"], dim=c["dim"], heads=c["heads"], aux_dim=1,
                         timestep_weight=1.0)
    eng = Engine(sd, g, torch.device("cuda"))
    t = torch.tensor([0.3, math.pi / 2, 1.1], device="cuda")
    aux = torch.tensor([[0.6], [1.2], [2.4]], device="cuda")
    gain, bias, cond = eng.conditioning(t, aux, want_cond=True)
    torch.cuda.synchronize()
    cref = orc.conditioning_vector(sd, t.cpu(), aux.cpu(), c["dim"], 1, 1.0, 3)
    assert _rel(cond.cpu(), cref) < 1e-5
    D = c["dim"]
    for l in range(c["depth"]):
        for j, blk in enumerate(("0", "1")):
            pre = f"transformer.layers.{l}.{blk}.norm"
            mod = torch.nn.functional.linear(cref, sd[pre + ".modulation.weight"], sd[pre + ".modulation.bias"])
            g_ref = sd[pre + ".norm.weight"] * (1 + mod[:, :D])
            b_ref = sd[pre + ".norm.bias"] * (1 + mod[:, :D]) + mod[:, D:]
            assert _rel(gain[2 * l + j].cpu(), g_ref) < 1e-5
            assert _rel(bias[2 * l + j].cpu(), b_ref) < 1e-5


@pytest.mark.parametrize("shape,weights", [((2, 5, 32, 64), True), ((3, 7, 12, 20), False), ((1, 69, 128, 256), True)])
def test_scm_loss_glue_kernels_vs_oracle(lib, shape, weights):
    """swb200_scm_noised_inputs / swb200_scm_tangent_target against oracle/scm_loss_oracle.py (itself pinned to the real
    SCMLoss) with random F / dF, i.e. independent of the network: fp32 elementwise work, fp64 fixed-order reductions."""
    from oracle import scm_loss_oracle as so
    B, Cn, H, W = shape
    g = torch.Generator().manual_seed(B * 100 + Cn)
    x, z, F, dF = (torch.randn(shape, generator=g).cuda() for _ in range(4))
    t = (torch.rand(B, generator=g) * 1.5 + 0.02).cuda()
    x_t, dxt, vx = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    vt = torch.empty_like(t)
    _check(lib.swb200_scm_noised_inputs(x.data_ptr(), z.data_ptr(), t.data_ptr(), B, Cn, H, W, x_t.data_ptr(),
                                        dxt.data_ptr(), vx.data_ptr(), vt.data_ptr(), _stream()))
    t4 = t.view(B, 1, 1, 1)
    c, s = torch.cos(t4), torch.sin(t4)
    assert torch.allclose(x_t, c * x + s * z, atol=1e-6, rtol=1e-5)
    assert torch.allclose(dxt, c * z - s * x, atol=1e-6, rtol=1e-5)
    assert torch.allclose(vx, c * s * (c * z - s * x), atol=1e-6, rtol=1e-5)
    assert torch.allclose(vt, (c * s).view(B), atol=1e-7, rtol=1e-5)
    w_lat = so.latitude_weights(H).cuda() if weights else None
    w_var = (torch.rand(1, Cn, 1, 1, generator=g) + 0.1).cuda() if weights else None
    for r, sd in ((1.0, 1.0), (0.25, 0.5)):
        gbuf, cot = torch.empty_like(x), torch.empty_like(x)
        loss = torch.empty((), device="cuda")
        need = lib.swb200_scm_target_scratch_bytes(B)
        scratch = torch.empty(need // 8, dtype=torch.float64, device="cuda")
        _check(lib.swb200_scm_tangent_target(F.data_ptr(), dF.data_ptr(), x_t.data_ptr(), dxt.data_ptr(), t.data_ptr(), r, sd,
                                             None if w_var is None else w_var.data_ptr(),
                                             None if w_lat is None else w_lat.data_ptr(), B, Cn, H, W, gbuf.data_ptr(),
                                             cot.data_ptr(), loss.data_ptr(), scratch.data_ptr(), need, _stream()))
        ref_g = so.scm_tangent_target(F, dF, x_t, dxt, c, s, r, sd)
        w = (w_var if weights else 1.0) * (w_lat if weights else 1.0)
        ref_loss = (w * ref_g.square()).sum(dim=1).mean()
        ref_cot = -2.0 * w * ref_g / (B * H * W)
        assert _rel(gbuf, ref_g) < 2e-6 and _rel(cot, ref_cot) < 2e-6
        assert abs(float(loss) - float(ref_loss)) < 1e-5 * float(ref_loss)
    assert lib.swb200_scm_tangent_target(F.data_ptr(), dF.data_ptr(), x_t.data_ptr(), dxt.data_ptr(), t.data_ptr(), 1.0, 1.0,
                                         None, None, B, Cn, H, W, gbuf.data_ptr(), cot.data_ptr(), loss.data_ptr(),
                                         scratch.data_ptr(), 8, _stream()) != 0          # scratch too small: refused


@pytest.mark.parametrize("act_fp16", [True, False], ids=["fp16", "bf16"])
@pytest.mark.parametrize("gemm_tile", [3, 2])
@pytest.mark.parametrize("cfgname", ["SWIFT_TINY", "SWIFT_SMALL"])
def test_pack_weights_matches_layout_oracle(lib, cfgname, gemm_tile, act_fp16):
    """swb200_pack_weights (the C-ABI checkpoint packer, driven through ctypes from raw parameter pointers) writes exactly the
    layouts of oracle/pack_oracle.py, whose meaning tests/test_host_cpu.py::test_packed_layouts_reproduce_oracle proves."""
    import ctypes as C
    from oracle import pack_oracle
    from swift_b200 import _lib, packing, synthetic as syn
    c = getattr(syn, cfgname)
    g = packing.Geometry(img=packing._pair(c["img_resolution"]), patch=packing._pair(c["patch_size"]),
                         window=packing._pair(c["window_size"]), shift=packing._pair(c["shift_size"]),
                         in_channels=c["in_channels"], out_channels=c["out_channels"], depth=c["depth"], dim=c["dim"],
                         heads=c["heads"], aux_dim=c["auxiliary_dim"], timestep_weight=1.0)
    sd = syn.random_state_dict(c, seed=3)
    dev = torch.device("cuda")
    want = pack_oracle.pack_layouts(sd, g, dev, act_fp16=act_fp16, gemm_tile=gemm_tile)
    m, keep = packing.pack(sd, g, dev, act_fp16=act_fp16, gemm_tile=gemm_tile)
    torch.cuda.synchronize()
    buf = keep["packed"]
    for name, ref in want.items():
        ptr = getattr(m, name)
        assert ptr, name
        off = ptr - buf.data_ptr()
        got = buf[off:off + ref.numel() * ref.element_size()].view(ref.dtype).reshape(ref.shape)
        if ref.dtype == torch.float32 and name in ("qscale", "pos_embed"):
            assert torch.allclose(got, ref, rtol=1e-6, atol=0), name          # expf / one fp32 add
        else:
            assert torch.equal(got, ref), name
    assert m.b_embed is None
    # error paths of the C entry point
    assert lib.swb200_pack_weights(C.byref(m), None, buf.data_ptr(), buf.numel(), None) != 0
    r, _keep, _ = packing.ref_params(sd, g, dev)
    assert lib.swb200_pack_weights(C.byref(m), C.byref(r), (buf.data_ptr() + 255) // 256 * 256, 16, None) != 0
