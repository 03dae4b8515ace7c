"""Ensemble statistics kernel through the C ABI against the oracle / the reference's golden scores, and scored rollouts."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def test_kernel_matches_reference_golden():
    from make_metrics_golden import fixture
    from swift_b200 import ensemble as ens
    pred, truth, lat = fixture()
    B, N, V, H, W = pred.shape
    st = ens.EnsembleStatistics(N, B, V, (H, W), lat, steps=3, device="cuda")
    phys = pred.float().reshape(B * N, V, H, W).contiguous().cuda()
    step_dev = torch.tensor([2], dtype=torch.int32, device="cuda")
    st.accumulate(phys, truth.float().cuda(), step_dev=step_dev)
    torch.cuda.synchronize()
    assert (st.sums[:2] == 0).all()                                   # only the row named by the device step counter
    ref = ens.sums_reference(pred.float().double(), truth.float().double(), lat)
    assert torch.allclose(st.sums[2].cpu(), ref, rtol=2e-6), (st.sums[2].cpu() / ref - 1).abs().max()
    flat = st.as_reference_dict({k: v[2:3] for k, v in st.scores().items()}, [f"v{i}" for i in range(V)], [6])
    g = np.load(os.path.join(ROOT, "tests", "golden", "metrics.npz"))
    for k, v in zip(g["keys"], g["values"]):
        assert flat[str(k)] == pytest.approx(float(v), rel=2e-5), k  # fp32 inputs vs the float64 golden run


@pytest.mark.parametrize("N,V,H,W,B", [(12, 3, 32, 64, 2), (2, 1, 8, 8, 1), (20, 2, 16, 24, 3)])
def test_kernel_shapes_vs_oracle(N, V, H, W, B):
    from oracle import metrics_oracle as mo
    from swift_b200 import ensemble as ens
    g = torch.Generator().manual_seed(N * 100 + V)
    truth = torch.randn(B, V, H, W, generator=g)
    pred = truth.unsqueeze(1) + 0.7 * torch.randn(B, N, V, H, W, generator=g)
    lat = np.linspace(-87, 87, H)
    st = ens.EnsembleStatistics(N, B, V, (H, W), lat, steps=1, device="cuda")
    st.accumulate(pred.reshape(B * N, V, H, W).contiguous().cuda(), truth.cuda(), step=0)
    sc = st.scores()
    for k, fn in (("rmse", mo.rmse), ("crps", mo.crps), ("ssr", mo.spread_skill)):
        torch.testing.assert_close(sc[k][0].cpu(), fn(pred.double(), truth.double(), lat), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("use_graph", [True, False], ids=["graph", "eager"])
def test_scored_rollout(use_graph):
    """Statistics accumulated inside the (graph-replayed) 6 h step equal the oracle scores of the states the rollout
    returned, step by step."""
    from oracle import metrics_oracle as mo
    from swift_b200 import synthetic as syn
    from swift_b200.ensemble import EnsembleStatistics
    from swift_b200.rollout import EnsembleRollout, Normalizers, shard_trajectories
    from test_gpu_forward import build_net
    cfg = syn.SWIFT_SMALL
    n_var = cfg["out_channels"]
    net, _ = build_net(cfg)
    members, n_ic, steps = 3, 2, 3
    traj = shard_trajectories(members, n_ic, 0, 1)
    H, W = cfg["img_resolution"]
    n_forc = cfg["in_channels"] - 2 * n_var
    forc = torch.randn(steps + 2, n_forc, H, W, device="cuda")
    ro = EnsembleRollout(net, Normalizers.synthetic(n_var, "cuda"), forc, traj, use_graph=use_graph, ic_times=[0, 1])
    lat = np.linspace(-88, 88, H)
    st = EnsembleStatistics(members, n_ic, n_var, (H, W), lat, steps, "cuda")
    truth = torch.zeros(n_ic, n_var, H, W, device="cuda")
    ro.attach_statistics(st, truth)
    ro.set_state(torch.randn(len(traj), n_var, H, W, device="cuda"))
    states, truths = [], []
    for i in range(steps):
        truth.copy_(torch.randn(n_ic, n_var, H, W, generator=torch.Generator().manual_seed(i)).cuda())
        states.append(ro.step().clone())
        truths.append(truth.clone())
    sc = st.scores(st.gather())
    for i in range(steps):
        pred = states[i].reshape(n_ic, members, n_var, H, W).double().cpu()
        for k, fn in (("rmse", mo.rmse), ("crps", mo.crps), ("ssr", mo.spread_skill)):
            torch.testing.assert_close(sc[k][i].cpu(), fn(pred, truths[i].double().cpu(), lat), rtol=1e-5, atol=1e-7)
    with pytest.raises(ValueError):
        EnsembleRollout(net, Normalizers.synthetic(n_var, "cuda"), forc, traj[1:], ic_times=[0, 1]).attach_statistics(st, truth)
