"""Optimiser step on the GPU: swb200_muon_step / swb200_adam_step (training/optimizers/muon.py) against the oracle (pinned to
the real reference's functions by tests/golden/muon.npz) and the reference's own golden outputs.

Tolerance: Newton-Schulz is a chain of 15 bf16 GEMMs whose outputs are rounded to bf16 at every step; the tensor-core
accumulation order differs from the CPU's, and a 1-ulp flip early in the chain propagates, so the orthogonalised update is
compared by relative L2 (bar 5e-2; two runs of the REFERENCE's own code -- CPU vs cuBLAS matmuls -- differ by the amount
printed as "reference CPU vs GPU") and by its defining property (singular values in the iteration's fixed band)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("shape", [(264, 264), (528, 264), (264, 704), (1056, 1056), (2112, 1056), (1056, 2816)])
def test_muon_step_vs_oracle(shape):
    from oracle import muon_oracle as mo
    from swift_b200.optim import MuonWithAuxAdam
    g = torch.Generator().manual_seed(shape[0] + shape[1])
    p0 = torch.randn(shape, generator=g) * 0.02
    grad = torch.randn(shape, generator=g) * 1e-4
    mom0 = torch.randn(shape, generator=g) * 1e-4
    p = torch.nn.Parameter(p0.clone().cuda())
    p.grad = grad.clone().cuda()
    opt = MuonWithAuxAdam([dict(params=[p], use_muon=True, lr=0.02, weight_decay=0.01)])
    opt.state[p]["momentum_buffer"] = mom0.clone().cuda()
    opt.step()
    torch.cuda.synchronize()
    upd_ref, mom_ref = mo.muon_update(grad.cuda(), mom0.cuda(), beta=0.95)            # oracle on the GPU (bf16 matmuls of torch)
    assert torch.allclose(opt.state[p]["momentum_buffer"], mom_ref, rtol=1e-6, atol=1e-10)
    assert torch.equal(p.grad.cpu(), grad), "the gradient must not be modified"
    upd = (p0.cuda() * (1 - 0.02 * 0.01) - p.detach()) / 0.02
    e = _rel(upd, upd_ref)
    floor = _rel(mo.muon_update(grad, mom0, beta=0.95)[0].cuda(), upd_ref) if shape[0] * shape[1] <= 1056 * 1056 else float("nan")
    sv = torch.linalg.svdvals(upd.double() / max(1, shape[0] / shape[1]) ** 0.5)
    print(f"muon {shape}: update rel-L2 vs oracle {e:.3e} (reference CPU vs GPU: {floor:.3e}); singular values "
          f"{sv.min():.3f} .. {sv.max():.3f}")
    assert e < 5e-2, e
    assert sv.max() < 1.3                              # "S' ~ Uniform(0.5, 1.5)" (muon.py:10-13)
    if shape[0] != shape[1]:                           # (a random SQUARE matrix has singular values near 0 that 5 steps do not lift)
        assert 0.3 < sv.min()


def test_muon_small_matrix_and_golden(golden):
    """The [1, heads, 1, 1] logit scale goes through the host-side vector path; all four golden cases of the real reference."""
    from swift_b200.optim import MuonWithAuxAdam
    g = golden("muon")
    for name in ("wide", "tall", "square", "scale"):
        grad, mom = torch.from_numpy(g[f"{name}_grad"]).cuda(), torch.from_numpy(g[f"{name}_mom"]).cuda()
        p = torch.nn.Parameter(torch.zeros_like(grad))
        p.grad = grad.clone()
        opt = MuonWithAuxAdam([dict(params=[p], use_muon=True, lr=1.0, weight_decay=0.0)])
        opt.state[p]["momentum_buffer"] = mom.clone()
        opt.step()
        assert torch.allclose(opt.state[p]["momentum_buffer"].cpu(), torch.from_numpy(g[f"{name}_mom_new"]), rtol=1e-6, atol=1e-10)
        e = _rel(-p.detach().cpu(), torch.from_numpy(g[f"{name}_update"]))
        print(f"muon golden {name}: rel-L2 {e:.3e}")
        assert e < 5e-2, (name, e)


def test_adam_step_vs_golden(golden):
    from swift_b200.optim import MuonWithAuxAdam
    g = golden("muon")
    p = torch.nn.Parameter(torch.from_numpy(g["adam_p"]).cuda())
    p.grad = torch.from_numpy(g["adam_g"]).cuda()
    opt = MuonWithAuxAdam([dict(params=[p], use_muon=False, lr=1e-2, betas=(0.9, 0.95), eps=1e-10, weight_decay=0.1)])
    opt.state[p].update(exp_avg=torch.from_numpy(g["adam_b1"]).cuda(), exp_avg_sq=torch.from_numpy(g["adam_b2"]).cuda(), step=2)
    opt.step()
    ref = torch.from_numpy(g["adam_p"]) * (1 - 1e-2 * 0.1) - 1e-2 * torch.from_numpy(g["adam_update"])
    assert torch.allclose(p.detach().cpu(), ref, rtol=2e-5, atol=1e-7)
    assert torch.allclose(opt.state[p]["exp_avg"].cpu(), torch.from_numpy(g["adam_b1_new"]), rtol=1e-5, atol=1e-8)
    assert torch.allclose(opt.state[p]["exp_avg_sq"].cpu(), torch.from_numpy(g["adam_b2_new"]), rtol=1e-5, atol=1e-10)


def test_optimizer_refuses_cpu_parameters():
    from swift_b200.optim import MuonWithAuxAdam
    p = torch.nn.Parameter(torch.zeros(16, 16))
    p.grad = torch.zeros(16, 16)
    with pytest.raises(RuntimeError):
        MuonWithAuxAdam([dict(params=[p], use_muon=True)]).step()
