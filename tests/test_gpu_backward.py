"""Reverse mode on the GPU: every backward kernel against PyTorch autograd of the same op, then
``F_x.backward(cot)`` through ``swift_b200.SwinV2`` against the parameter gradients of the fp32 oracle
(``oracle.scm_loss_oracle.scm_parameter_gradients``, itself pinned to the real reference's ``SCMLoss(...).backward()``).

Tolerances: the training path uses bf16 tensor-core operands with fp32 accumulation (the reference trains under bf16
autocast); a gradient tensor is compared by relative L2 over the whole tensor, bar 3e-2 (bf16 rounding of both operands of
every wgrad / dgrad product: ~2^-9 each, accumulated through up to 12 layers), and its norm within 1e-2."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HD, HDP = 88, 96


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def lib():
    from swift_b200 import _lib
    return _lib.lib()


def _check(rc):
    from swift_b200 import _lib
    _lib.check(rc, "test")


# --------------------------------------------------------------------------------------------- kernels
@pytest.mark.parametrize("R,C_,pitch", [(512, 264, 264), (8192, 1056, 2112), (512, 280, 280), (1024, 5632, 5632)])
def test_transpose16(lib, R, C_, pitch):
    x = torch.randn(R, pitch, device="cuda").to(torch.bfloat16)
    out = torch.zeros(C_, R, device="cuda", dtype=torch.bfloat16)
    _check(lib.swb200_transpose16(x.data_ptr(), R, C_, pitch, out.data_ptr(), R, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(out, x[:, :C_].t().contiguous())


@pytest.mark.parametrize("tile", [3, 2, 1])
@pytest.mark.parametrize("M,N,K,S", [(1056, 1056, 2048, 4), (280, 264, 512, 1), (5632, 1056, 1024, 2), (568, 528, 512, 8)])
def test_gemm_splitk(lib, tile, M, N, K, S):
    """Weight-gradient shape: D[M, N] = A[M, K_total] W[N, K_total]^T as S stacked partial products over K_total = S*K."""
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = (torch.randn(M, S * K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, S * K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    part = torch.full((S, M, N), float("nan"), device="cuda")
    _check(lib.swb200_gemm_splitk(tile, A.data_ptr(), S * K, W.data_ptr(), S * K, part.data_ptr(), N, M, N, K, S, _stream()))
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = False
    for s in range(S):
        ref = A[:, s * K:(s + 1) * K].float() @ W[:, s * K:(s + 1) * K].float().t()
        assert _rel(part[s], ref) < 1e-5, (s, _rel(part[s], ref))
    assert _rel(part.sum(0), A.float() @ W.float().t()) < 1e-5


@pytest.mark.parametrize("B,T,D", [(2, 512, 264), (1, 8192, 1056)])
@pytest.mark.parametrize("with_add", [False, True])
def test_ln_backward(lib, B, T, D, with_add):
    M = B * T
    g = torch.Generator(device="cuda").manual_seed(7)
    branch = torch.randn(M, D, device="cuda", generator=g) * 3 + 0.5
    gain = 1 + 0.3 * torch.randn(B, D, device="cuda", generator=g)
    dx0 = torch.randn(M, D, device="cuda", generator=g) * 1e-3
    add = torch.randn(M, D, device="cuda", generator=g) * 1e-3 if with_add else None
    dx = dx0.clone()
    db16 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    dgain = torch.zeros(B, D, device="cuda")
    dbias = torch.zeros(B, D, device="cuda")
    need = lib.swb200_ln_backward_scratch_bytes(M, D, T)
    scratch = torch.empty(need, dtype=torch.uint8, device="cuda")
    _check(lib.swb200_ln_backward(dx.data_ptr(), None if add is None else add.data_ptr(), branch.data_ptr(), gain.data_ptr(),
                                  db16.data_ptr(), dgain.data_ptr(), dbias.data_ptr(), M, D, T, 0, scratch.data_ptr(), need, _stream()))
    torch.cuda.synchronize()
    b = branch.double().requires_grad_(True)
    gn = gain.double().requires_grad_(True)
    bs = torch.zeros(B, D, device="cuda", dtype=torch.float64, requires_grad=True)
    n = torch.nn.functional.layer_norm(b, (D,), eps=1e-6).reshape(B, T, D)
    y = n * gn[:, None] + bs[:, None]
    dy = (dx0 + (add if add is not None else 0)).double().reshape(B, T, D)
    y.backward(dy)
    assert _rel(db16.float(), b.grad) < 4e-3            # bf16 output
    assert _rel(dgain, gn.grad) < 1e-5 and _rel(dbias, bs.grad) < 1e-5
    assert torch.allclose(dx, dx0 + add if with_add else dx0)


def test_swiglu_backward(lib):
    M, Dff = 512, 704
    g = torch.Generator(device="cuda").manual_seed(3)
    gu = (torch.randn(M, 2 * Dff, device="cuda", generator=g) * 2).to(torch.bfloat16)
    dh = torch.randn(M, Dff, device="cuda", generator=g) * 1e-4
    dgu = torch.empty(M, 2 * Dff, device="cuda", dtype=torch.bfloat16)
    _check(lib.swb200_swiglu_backward(dh.data_ptr(), gu.data_ptr(), dgu.data_ptr(), M, Dff, _stream()))
    torch.cuda.synchronize()
    x = gu.double().requires_grad_(True)
    (torch.nn.functional.silu(x[:, :Dff]) * x[:, Dff:]).backward(dh.double())
    assert _rel(dgu.float(), x.grad) < 4e-3


def _attention_ref(raw, qscale, B, gh, gw, H, shift):
    """models/swinv2.py:118-135 + :186-209 on raw to_qkv outputs [M, 3D] (reference column order), float64 autograd."""
    M = B * gh * gw
    x = raw.reshape(B, gh, gw, H, 3, HD)
    x = torch.roll(x, shifts=(-shift[0], -shift[1]), dims=(1, 2))
    x = x.reshape(B, gh // 16, 16, gw // 16, 16, H, 3, HD).permute(0, 1, 3, 5, 6, 2, 4, 7)
    x = x.reshape(B * (gh // 16) * (gw // 16), H, 3, 256, HD)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
    q = torch.nn.functional.normalize(q, dim=-1) * qscale[None, :, None, None]
    k = torch.nn.functional.normalize(k, dim=-1)
    o = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v                      # [BW, H, 256, HD]
    o = o.reshape(B, gh // 16, gw // 16, H, 16, 16, HD).permute(0, 1, 4, 2, 5, 3, 6).reshape(B, gh, gw, H * HD)
    o = torch.roll(o, shifts=(shift[0], shift[1]), dims=(1, 2))
    return o.reshape(M, H * HD)


@pytest.mark.parametrize("impl", [2, 1], ids=["tcgen05", "mma_sync"])
@pytest.mark.parametrize("shift", [(0, 0), (8, 8), (8, 0)])
@pytest.mark.parametrize("B,gh,gw,H", [(1, 16, 32, 3), (2, 32, 32, 2), (1, 64, 128, 12)])
def test_attention_backward(lib, shift, B, gh, gw, H, impl):
    M, D = B * gh * gw, H * HD
    g = torch.Generator(device="cuda").manual_seed(11)
    raw = torch.randn(M, 3 * D, device="cuda", generator=g)
    qscale = torch.linspace(4.0, 12.0, H, device="cuda")
    packed = torch.empty(3, H, M, HDP, device="cuda", dtype=torch.bfloat16)
    invn = torch.empty(2, H, M, device="cuda")
    _check(lib.swb200_qkv_pack_train(raw.data_ptr(), qscale.data_ptr(), packed.data_ptr(), invn.data_ptr(), M, H, _stream()))
    O = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.full((H, M), float("nan"), device="cuda")
    _check(lib.swb200_window_attention(packed.data_ptr(), O.data_ptr(), B, gh, gw, H, shift[0], shift[1], 0, 0, 2, lse.data_ptr(),
                                       _stream()))
    dO = (torch.randn(M, D, device="cuda", generator=g) * 1e-5).to(torch.bfloat16)
    dqkv = torch.full((M, 3 * D), float("nan"), device="cuda", dtype=torch.bfloat16)
    dscale = torch.zeros(H, device="cuda")
    need = lib.swb200_attention_backward_scratch_bytes(B, gh, gw, H)
    scratch = torch.empty(need, dtype=torch.uint8, device="cuda")
    _check(lib.swb200_attention_backward(packed.data_ptr(), O.data_ptr(), dO.data_ptr(), invn.data_ptr(), qscale.data_ptr(),
                                         lse.data_ptr(), impl, dqkv.data_ptr(), dscale.data_ptr(), B, gh, gw, H, shift[0], shift[1], 0,
                                         scratch.data_ptr(), need, _stream()))
    torch.cuda.synchronize()
    # the log-sum-exp rows of the forward against the definition (token order)
    rq = raw.reshape(M, H, 3, HD)
    qn = torch.nn.functional.normalize(rq[:, :, 0].double(), dim=-1) * qscale.double()[None, :, None]
    kn = torch.nn.functional.normalize(rq[:, :, 1].double(), dim=-1)

    def win(t):                                                      # [M, H, HD] -> [BW, H, 256, HD]
        t = t.reshape(B, gh, gw, H, HD)
        t = torch.roll(t, shifts=(-shift[0], -shift[1]), dims=(1, 2))
        return t.reshape(B, gh // 16, 16, gw // 16, 16, H, HD).permute(0, 1, 3, 5, 2, 4, 6).reshape(-1, H, 256, HD)

    lse_ref = torch.logsumexp(win(qn) @ win(kn).transpose(-1, -2), dim=-1)                       # [BW, H, 256]
    lse_ref = lse_ref.reshape(B, gh // 16, gw // 16, H, 16, 16).permute(0, 1, 4, 2, 5, 3).reshape(B, gh, gw, H)
    lse_ref = torch.roll(lse_ref, shifts=(shift[0], shift[1]), dims=(1, 2)).reshape(M, H).t()
    assert torch.allclose(lse.double(), lse_ref, rtol=0, atol=3e-2), (lse.double() - lse_ref).abs().max()
    # the packed operands and the inverse norms against the definition
    r = raw.reshape(M, H, 3, HD).permute(2, 1, 0, 3)
    assert _rel(invn[0], 1 / r[0].norm(dim=-1)) < 1e-5 and _rel(invn[1], 1 / r[1].norm(dim=-1)) < 1e-5
    assert (packed[..., HD:] == 0).all()
    assert _rel(packed[2, ..., :HD].float(), r[2]) < 4e-3
    # gradients against float64 autograd of the reference formulation
    x = raw.double().requires_grad_(True)
    s = qscale.double().requires_grad_(True)
    out = _attention_ref(x, s, B, gh, gw, H, shift)
    assert _rel(O.float(), out) < 1.5e-2
    out.backward(dO.double())
    got = dqkv.float().reshape(M, H, 3, HD)
    want = x.grad.reshape(M, H, 3, HD)
    for part, name in enumerate("qkv"):
        e = _rel(got[:, :, part], want[:, :, part])
        print(f"attention backward d{name}: rel-L2 {e:.3e}")
        assert e < 2e-2, (name, e)
    assert _rel(dscale, s.grad) < 2e-2, (dscale, s.grad)


# --------------------------------------------------------------------------------------------- end to end
def _build_train(cfg, seed=1):
    from test_gpu_forward import build_net
    net, sd = build_net(cfg, seed=seed, act_fp16=True)
    return net, sd


def _oracle_grads(sd, cfg, x, t, cond, aux, cot):
    """VJP of the fp32 oracle forward (on the GPU, TF32 off) with the output cotangent."""
    from oracle import swinv2_oracle as orc
    torch.backends.cuda.matmul.allow_tf32 = False
    ocfg = orc.make_cfg(**cfg)
    p = {k: v.detach().clone().cuda().requires_grad_(True) for k, v in sd.items()}
    F = orc.pass_precond(p, ocfg, x, t, cond, aux)
    F.backward(cot)
    return {k: v.grad for k, v in p.items()}, F.detach()


@pytest.mark.parametrize("cfgname,B", [("SWIFT_TINY", 2), ("SWIFT_SMALL", 1), ("SWIFT_SMALL", 3)])
def test_backward_through_module_matches_oracle(cfgname, B):
    """net.train(); F = net(x, t, cond, aux); F.backward(cot): every parameter's .grad against the oracle's VJP."""
    from swift_b200 import synthetic as syn
    cfg = getattr(syn, cfgname)
    net, sd = _build_train(cfg)
    net.train()
    lat, cond = syn.synthetic_fields(cfg, B, seed=5)
    x, cond = lat.cuda(), cond.cuda()
    t = torch.linspace(0.4, 1.3, B, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(2)
    cot = torch.randn(x.shape, device="cuda", generator=g) * 1e-5            # the size of an sCM loss cotangent
    F = net(x, t, cond, 0.6)
    assert F.requires_grad
    F.backward(cot)
    ref, F_ref = _oracle_grads(sd, cfg, x, t, cond, 0.6, cot)
    from test_gpu_forward import per_field_rel_l2
    assert per_field_rel_l2(F, F_ref).max() < 1.5e-2                          # bf16 operand forward
    worst = 0.0
    for name, p in net.model.named_parameters():
        assert p.grad is not None, name
        e = _rel(p.grad, ref[name])
        nrm = abs(p.grad.norm().item() / ref[name].norm().item() - 1)
        worst = max(worst, e)
        print(f"{cfgname} B={B} {name:55s} rel-L2 {e:.3e} norm dev {nrm:.3e}")
        # the logit-scale gradient is ONE number per head, the sum of 10^5 cancelling terms: looser bar
        bar_e, bar_n = (8e-2, 5e-2) if name.endswith(".scale") else (3e-2, 1e-2)
        assert e < bar_e and nrm < bar_n, (name, e, nrm)
    # a second step re-uses tape / workspace / gradient buffers: same result bit for bit
    first = {n: p.grad.clone() for n, p in net.model.named_parameters()}
    net.zero_grad(set_to_none=True)
    net(x, t, cond, 0.6).backward(cot)
    for n, p in net.model.named_parameters():
        assert torch.equal(p.grad, first[n]), n


def test_swift_b_backward_gradient_norms():
    """Swift-B, batch 1: the norm of every parameter gradient against the fp32 oracle's VJP (eager PyTorch on the GPU)."""
    from swift_b200 import synthetic as syn
    cfg = syn.SWIFT_B
    from test_gpu_forward import build_net
    net, sd = build_net(cfg, img_channels=syn.IMG_CHANNELS)
    net.train()
    lat, cond = syn.synthetic_fields(cfg, 1, seed=0)
    x, cond = lat.cuda(), cond.cuda()
    t = torch.tensor([1.0], device="cuda")
    cot = torch.randn(x.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4)) * 1e-6
    F = net(x, t, cond, 0.6)
    F.backward(cot)
    torch.cuda.synchronize()
    ref, _ = _oracle_grads(sd, cfg, x, t, cond, 0.6, cot)
    worst_n, worst_e = 0.0, 0.0
    for name, p in net.model.named_parameters():
        nrm = abs(p.grad.norm().item() / ref[name].norm().item() - 1)
        e = _rel(p.grad, ref[name])
        if not name.endswith(".scale"):
            worst_n, worst_e = max(worst_n, nrm), max(worst_e, e)
        else:
            print(f"swift_b {name}: norm dev {nrm:.3e} rel-L2 {e:.3e}")
        if name.endswith(".scale"):
            continue
        if nrm > 5e-3 or e > 2e-2:
            print(f"swift_b {name}: norm dev {nrm:.3e} rel-L2 {e:.3e}")
    print(f"swift_b backward: worst gradient-norm deviation {worst_n:.3e}, worst rel-L2 {worst_e:.3e} over "
          f"{len(ref)} tensors")
    assert worst_n < 2e-2 and worst_e < 5e-2
