"""tcgen05 GEMM + fused epilogues through the C ABI, against plain PyTorch fp32 references built from the SAME
bf16-rounded operands (so the only difference is accumulation order: tolerance 2e-3 relative L2 / bf16 output ulp)."""
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


def _rand_bf16(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


ACT = pytest.mark.parametrize("f16", [1, 0], ids=["act_fp16", "act_bf16"])


def _adt(f16):
    return torch.float16 if f16 else torch.bfloat16


@ACT
@pytest.mark.parametrize("cg", [3, 2, 1], ids=["tile256x352", "tile256x176", "tile128x176"])
@pytest.mark.parametrize("M,N,K", [(256, 176, 64), (256, 176, 128), (512, 352, 1056), (4096, 1056, 2816),
                                   (300, 276, 568), (8192, 1056, 1056)])
def test_gemm_store_f32(lib, cg, M, N, K, f16):
    """A and W in the model's 16-bit operand format (fp16 or bf16)."""
    A = _rand_bf16((M, K), 1, dtype=_adt(f16))
    W = _rand_bf16((N, K), 2, 0.05, dtype=_adt(f16))
    ldo = (N + 7) // 8 * 8
    out = torch.full((M, ldo), float("nan"), device="cuda")
    _check(lib.swb200_gemm(0, cg, f16, A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), ldo, M, N, K, _stream()))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    got = out[:, :N]
    assert torch.isfinite(got).all(), "non-finite output (tile never written?)"
    assert _rel(got, ref) < 2e-3, f"rel L2 {_rel(got, ref):.3e}"


@ACT
@pytest.mark.parametrize("cg", [3, 2, 1], ids=["tile256x352", "tile256x176", "tile128x176"])
def test_gemm_store_act_strided_operands(lib, cg, f16):
    M, N, K = 512, 528, 264
    Abig = _rand_bf16((M, 2 * K), 3, dtype=_adt(f16))          # A is the left half of a wider buffer (row pitch 2K)
    W = _rand_bf16((N, K), 4, 0.05, dtype=_adt(f16))
    out = torch.zeros((M, N), device="cuda", dtype=_adt(f16))
    _check(lib.swb200_gemm(1, cg, f16, Abig.data_ptr(), 2 * K, W.data_ptr(), K, out.data_ptr(), N, M, N, K, _stream()))
    torch.cuda.synchronize()
    ref = Abig[:, :K].float() @ W.float().t()
    assert _rel(out.float(), ref) < 6e-3


@pytest.mark.parametrize("f16,q16", [(1, 1), (0, 0), (0, 1)], ids=["fp16", "bf16", "bf16-fp16qkv"])
@pytest.mark.parametrize("cg", [3, 2, 1], ids=["tile256x352", "tile256x176", "tile128x176"])
def test_gemm_qkv_epilogue(lib, cg, f16, q16):
    """EPI_QKV (odd head count: slots straddle tiles): rows packed part*D + h*88 + d; q,k L2-normalised in fp32 (eps 1e-12), q * qscale[h]; pad to 96.
    q16: the packed q / k / v are fp16 although the GEMM operands are bf16 (the bf16 model's attention internals)."""
    M, H = 512, 5
    D = H * HD
    A = _rand_bf16((M, D), 5, dtype=_adt(f16))
    W = _rand_bf16((3 * D, D), 6, 0.03, dtype=_adt(f16))
    qscale = torch.linspace(5.0, 20.0, H, device="cuda")
    out = torch.full((3, H, M, HDP), float("nan"), device="cuda", dtype=_adt(q16))
    _check(lib.swb200_gemm_qkv(cg, f16, q16, A.data_ptr(), D, W.data_ptr(), qscale.data_ptr(), out.data_ptr(), M, D, H, _stream()))
    torch.cuda.synchronize()
    y = (A.float() @ W.float().t()).reshape(M, 3, H, HD).permute(1, 2, 0, 3)       # [3, H, M, 88]
    q = torch.nn.functional.normalize(y[0], dim=-1) * qscale[:, None, None]
    k = torch.nn.functional.normalize(y[1], dim=-1)
    ref = torch.stack([q, k, y[2]], 0)
    assert (out[..., HD:] == 0).all(), "pad columns 88..95 must be zero"
    for part, name in enumerate("qkv"):
        tol = 6e-4 if q16 else 4e-3                      # the only rounding left is the output's
        assert _rel(out[part, ..., :HD].float(), ref[part]) < tol, f"{name}: {_rel(out[part, ..., :HD].float(), ref[part]):.3e}"
    if f16:
        assert lib.swb200_gemm_qkv(cg, 1, 0, A.data_ptr(), D, W.data_ptr(), qscale.data_ptr(), out.data_ptr(), M, D, H, _stream()) != 0


@ACT
@pytest.mark.parametrize("cg", [3, 2, 1], ids=["tile256x352", "tile256x176", "tile128x176"])
def test_gemm_swiglu_epilogue(lib, cg, f16):
    M, D, Dff = 512, 264, 704
    A = _rand_bf16((M, D), 7, dtype=_adt(f16))
    W1 = _rand_bf16((2 * Dff, D), 8, 0.06, dtype=_adt(f16))        # reference layout: [gate | up]
    half = 2 * HD if cg == 3 else HD              # rows per tile: [half gate | half up]
    gate, up = W1[:Dff].reshape(Dff // half, 1, half, D), W1[Dff:].reshape(Dff // half, 1, half, D)
    Wp = torch.cat([gate, up], 1).reshape(2 * Dff, D).contiguous()
    out = torch.full((M, Dff), float("nan"), device="cuda", dtype=_adt(f16))
    _check(lib.swb200_gemm_swiglu(cg, f16, A.data_ptr(), D, Wp.data_ptr(), out.data_ptr(), M, D, Dff, _stream()))
    torch.cuda.synchronize()
    h = A.float() @ W1.float().t()
    ref = torch.nn.functional.silu(h[:, :Dff]) * h[:, Dff:]
    assert _rel(out.float(), ref) < 4e-3, f"{_rel(out.float(), ref):.3e}"


@ACT
@pytest.mark.parametrize("cg", [3, 2, 1], ids=["tile256x352", "tile256x176", "tile128x176"])
def test_gemm_embed_epilogue(lib, cg, f16):
    B, T, D, K = 2, 512, 264, 56
    M = B * T
    A = _rand_bf16((M, K), 9, dtype=_adt(f16))
    W = _rand_bf16((D, K), 10, 0.1, dtype=_adt(f16))
    bias = torch.randn(D, device="cuda")
    pos = torch.randn(T, D, device="cuda")
    xhl = torch.full((M, 2 * D), float("nan"), device="cuda", dtype=_adt(f16))
    _check(lib.swb200_gemm_embed(cg, f16, A.data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), pos.data_ptr(), T,
                                 xhl.data_ptr(), M, D, _stream()))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias + pos.repeat(B, 1)
    hi, lo = xhl[:, :D], xhl[:, D:]
    assert _rel(hi.float() + lo.float(), ref) < (3e-6 if f16 else 3e-5), f"{_rel(hi.float() + lo.float(), ref):.3e}"
    assert _rel(hi.float(), ref) < (6e-4 if f16 else 5e-3)          # hi alone is the 16-bit rounding of x
    assert torch.equal(lo, ((hi.float() + lo.float()) - hi.float()).to(_adt(f16)))


@pytest.mark.parametrize("cg", [3, 2, 1], ids=["tile256x352", "tile256x176", "tile128x176"])
def test_gemm_embed_epilogue_single_value_stream(lib, cg):
    """Format word 3: only the hi half of the residual buffer is written."""
    B, T, D, K = 2, 512, 264, 56
    M = B * T
    A = _rand_bf16((M, K), 9, dtype=torch.float16)
    W = _rand_bf16((D, K), 10, 0.1, dtype=torch.float16)
    pos = torch.randn(T, D, device="cuda")
    xhl = torch.full((M, 2 * D), float("nan"), device="cuda", dtype=torch.float16)
    _check(lib.swb200_gemm_embed(cg, 3, A.data_ptr(), K, W.data_ptr(), K, None, pos.data_ptr(), T, xhl.data_ptr(), M, D,
                                 _stream()))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + pos.repeat(B, 1)
    assert torch.isnan(xhl[:, D:]).all()
    assert _rel(xhl[:, :D].float(), ref) < 6e-4


@pytest.mark.parametrize("cg", [3, 2], ids=["tile256x352", "tile256x176"])
@pytest.mark.parametrize("B,T,D,K", [(3, 256, 264, 264), (5, 160, 528, 704), (3, 256, 1056, 1056), (2, 512, 1056, 2816),
                                     (6, 8192, 1056, 264)])
def test_gemm_ln_residual_epilogue_single_value_stream(lib, cg, B, T, D, K):
    """EPI_LN_RES with format word 3: x (hi half only) += LayerNorm(A W^T) * gain + bias, one fp16 rounding per update; the lo
    half is neither read nor written."""
    M = B * T
    A = _rand_bf16((M, K), 21, dtype=torch.float16)
    W = _rand_bf16((D, K), 22, 0.05, dtype=torch.float16)
    g = torch.Generator(device="cuda").manual_seed(23)
    hi = torch.randn(M, D, device="cuda", generator=g).half()
    xhl = torch.cat([hi, torch.full_like(hi, float("nan"))], 1).contiguous()
    x_ref = hi.float()
    ws = torch.empty(lib.swb200_ln_workspace_bytes(M, D) + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    branch = (A.float() @ W.float().t()).half().float()
    for gen in range(2):
        gain = torch.randn(B, D, device="cuda", generator=g)
        bias = torch.randn(B, D, device="cuda", generator=g)
        _check(lib.swb200_gemm_ln_residual(cg, 3, A.data_ptr(), K, W.data_ptr(), K, xhl.data_ptr(), gain.data_ptr(),
                                           bias.data_ptr(), M, D, T, ws_ptr, gen, _stream()))
        torch.cuda.synchronize()
        ln = torch.nn.functional.layer_norm(branch, (D,), eps=1e-6).reshape(B, T, D)
        x_ref = x_ref + (ln * gain[:, None] + bias[:, None]).reshape(M, D)
        got = xhl[:, :D].float()
        assert torch.isnan(xhl[:, D:]).all()
        assert _rel(got, x_ref) < 6e-4, f"gen {gen}: {_rel(got, x_ref):.3e}"
        x_ref = got.clone()


@ACT
@pytest.mark.parametrize("cg", [3, 2, 1], ids=["tile256x352", "tile256x176", "tile128x176"])
@pytest.mark.parametrize("mode", ["plain", "scm", "heun"])
def test_gemm_head_epilogue(lib, cg, mode, f16):
    from swift_b200 import _lib
    B, C_out, Himg, Wimg, p1, p2, D = 2, 5, 32, 64, 2, 2, 264
    gh, gw = Himg // p1, Wimg // p2
    M = B * gh * gw
    A = _rand_bf16((M, D), 11, dtype=_adt(f16))
    W = _rand_bf16((C_out * p1 * p2, D), 12, 0.05, dtype=_adt(f16))
    m = _lib.Model()
    m.act_fp16 = f16
    m.img_h, m.img_w, m.patch_h, m.patch_w, m.out_channels, m.dim = Himg, Wimg, p1, p2, C_out, D
    m.w_head = W.data_ptr()
    xt = torch.randn(B, C_out, Himg, Wimg, device="cuda")
    fprev = torch.randn_like(xt)
    y = torch.full_like(xt, float("nan"))
    out_f = torch.full_like(xt, float("nan"))
    F = (A.float() @ W.float().t()).reshape(B, gh, gw, C_out, p1, p2).permute(0, 3, 1, 4, 2, 5).reshape(B, C_out, Himg, Wimg)
    if mode == "plain":
        upd = _lib.Update(None, None, None, 0.0, 1.0, 0.0)
        ref = F
    elif mode == "scm":
        upd = _lib.Update(xt.data_ptr(), None, None, math.cos(1.1), -math.sin(1.1) * 0.5, 0.0)
        ref = math.cos(1.1) * xt - math.sin(1.1) * 0.5 * F
    else:
        upd = _lib.Update(xt.data_ptr(), fprev.data_ptr(), out_f.data_ptr(), 1.0, -0.25, -0.25)
        ref = xt - 0.25 * (F + fprev)
    _check(lib.swb200_gemm_head(cg, C.byref(m), A.data_ptr(), D, D, B, C.byref(upd), y.data_ptr(), _stream()))
    torch.cuda.synchronize()
    assert _rel(y, ref) < 1e-4, f"{_rel(y, ref):.3e}"
    if mode == "heun":
        assert _rel(out_f, F) < 1e-4


def test_gemm_rejects_bad_arguments(lib):
    from swift_b200 import _lib
    A = _rand_bf16((256, 64), 1)
    out = torch.zeros(256, 176, device="cuda")
    rc = lib.swb200_gemm(0, 2, 0, A.data_ptr(), 60, A.data_ptr(), 64, out.data_ptr(), 176, 256, 176, 60, _stream())
    assert rc != 0 and b"multiples of 8" in lib.swb200_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(lib.swb200_gemm(42, 2, 0, A.data_ptr(), 64, A.data_ptr(), 64, out.data_ptr(), 176, 256, 176, 64, _stream()))


@ACT
@pytest.mark.parametrize("cg", [3, 2, 1], ids=["tile256x352", "tile256x176", "tile128x176"])
@pytest.mark.parametrize("B,T,D,K", [(3, 256, 264, 264), (5, 160, 528, 704), (3, 256, 1056, 1056), (2, 512, 1056, 2816),
                                     (6, 8192, 1056, 264)])
def test_gemm_ln_residual_epilogue(lib, cg, B, T, D, K, f16):
    """EPI_LN_RES: x += LayerNorm(A W^T) * gain[b] + bias[b] on the [hi | lo] residual pair, row statistics exchanged
    between the CTAs that own the column tiles of a row; two launches share the exchange workspace (gen 0, 1).
    (6, 8192, ...) spans several waves of the persistent grid."""
    dt = _adt(f16)
    M = B * T
    A = _rand_bf16((M, K), 21, dtype=dt)
    W = _rand_bf16((D, K), 22, 0.05, dtype=dt)
    g = torch.Generator(device="cuda").manual_seed(23)
    x = torch.randn(M, D, device="cuda", generator=g)
    hi = x.to(dt)
    lo = (x - hi.float()).to(dt)
    xhl = torch.cat([hi, lo], 1).contiguous()
    x_ref = hi.float() + lo.float()
    ws = torch.empty(lib.swb200_ln_workspace_bytes(M, D) + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    branch = (A.float() @ W.float().t()).half().float()             # fp16 rounding of the accumulator, as in the kernel
    for gen in range(2):
        gain = torch.randn(B, D, device="cuda", generator=g)
        bias = torch.randn(B, D, device="cuda", generator=g)
        _check(lib.swb200_gemm_ln_residual(cg, f16, A.data_ptr(), K, W.data_ptr(), K, xhl.data_ptr(), gain.data_ptr(),
                                           bias.data_ptr(), M, D, T, ws_ptr, gen, _stream()))
        torch.cuda.synchronize()
        ln = torch.nn.functional.layer_norm(branch, (D,), eps=1e-6).reshape(B, T, D)
        x_ref = x_ref + (ln * gain[:, None] + bias[:, None]).reshape(M, D)
        got = xhl[:, :D].float() + xhl[:, D:].float()
        assert torch.isfinite(got).all()
        # the pair represents x to ~2^-22 (fp16) / 2^-16 (bf16); the fp16 rounding of the branch flips for the few elements
        # whose fp32 accumulation order differs from the reference's (more of them at large K)
        assert _rel(got, x_ref) < (8e-5 if f16 else 2e-4), f"gen {gen}: {_rel(got, x_ref):.3e}"
        assert _rel(xhl[:, :D].float(), x_ref) < (6e-4 if f16 else 5e-3)
        x_ref = got.clone()


@pytest.mark.timeout(180)
def test_gemm_ln_residual_next_to_a_busy_stream(lib):
    """The fused LayerNorm epilogue spins on statistics published by other CTAs of its own grid, so all of its clusters
    must be co-resident.  The kernel is launched cooperatively: with a second stream keeping every SM busy (long GEMMs
    of another library) the launch waits for room instead of starting half a grid that can never finish -- the failure
    mode was a watchdog trap that kills the context.  Same results as the solo run, bit for bit."""
    f16, cg = 1, 3
    B, T, D, K = 4, 8192, 1056, 2816
    M = B * T
    A = _rand_bf16((M, K), 31, dtype=torch.float16)
    W = _rand_bf16((D, K), 32, 0.05, dtype=torch.float16)
    g = torch.Generator(device="cuda").manual_seed(33)
    x = torch.randn(M, D, device="cuda", generator=g)
    hi = x.half()
    xhl0 = torch.cat([hi, (x - hi.float()).half()], 1).contiguous()
    gain = torch.randn(B, D, device="cuda", generator=g)
    bias = torch.randn(B, D, device="cuda", generator=g)
    ws = torch.empty(lib.swb200_ln_workspace_bytes(M, D) + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256

    def run(xhl, gen):
        _check(lib.swb200_gemm_ln_residual(cg, f16, A.data_ptr(), K, W.data_ptr(), K, xhl.data_ptr(), gain.data_ptr(),
                                           bias.data_ptr(), M, D, T, ws_ptr, gen, _stream()))

    solo = xhl0.clone()
    run(solo, 0)
    torch.cuda.synchronize()
    busy = torch.cuda.Stream()
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    together = xhl0.clone()
    torch.cuda.synchronize()
    with torch.cuda.stream(busy):
        for _ in range(40):                        # ~40 x 0.8 ms of kernels that fill every SM
            c = a @ b
    for i in range(6):                             # our launches arrive while the other stream's grid is running
        t = xhl0.clone() if i < 5 else together
        run(t, 1 + i)
    torch.cuda.synchronize()
    assert torch.equal(together, solo)
    assert torch.isfinite(c.float()).all()
