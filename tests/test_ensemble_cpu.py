"""Ensemble verification scores, host side: the oracle restatement against the REAL reference functions' outputs
(tests/golden/metrics.npz), the score algebra on the sufficient statistics the CUDA kernel accumulates, and the
world-size-2 exchange (gloo on CPU; NCCL on the GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from make_metrics_golden import fixture  # noqa: E402

from oracle import metrics_oracle as mo  # noqa: E402
from swift_b200 import ensemble as ens  # noqa: E402

VARS = ["v0", "v1", "v2", "v3"]


def _golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "metrics.npz"))
    return {str(k): float(v) for k, v in zip(g["keys"], g["values"])}


def test_oracle_matches_reference_golden():
    pred, truth, lat = fixture()
    got = mo.all_scores(pred, truth, VARS, lat, "6h")
    gold = _golden()
    assert set(got) == set(gold)
    for k, v in gold.items():
        assert got[k] == pytest.approx(v, rel=1e-12), k


def test_scores_from_sufficient_statistics_match_reference_golden():
    """rmse / crps / ssr recomputed from the four per-(IC, variable) sums equal the reference's direct formulas."""
    pred, truth, lat = fixture()
    B, N, V, H, W = pred.shape
    sums = ens.sums_reference(pred, truth, lat)                       # [B, V, 4]
    st = ens.EnsembleStatistics(N, B, V, (H, W), lat, steps=1, device="cpu")
    st.sums[0] = sums
    flat = st.as_reference_dict(st.scores(), VARS, [6])
    gold = _golden()
    assert set(flat) == set(gold)
    for k, v in gold.items():
        assert flat[k] == pytest.approx(v, rel=1e-10), k


def test_statistics_reject_bad_arguments():
    with pytest.raises(ValueError):
        ens.EnsembleStatistics(1, 2, 3, (4, 8), np.zeros(4), 1, "cpu")          # N - 1 = 0 in the CRPS spread term
    with pytest.raises(ValueError):
        ens.EnsembleStatistics(4, 2, 3, (4, 8), np.zeros(5), 1, "cpu")          # lat / grid mismatch
    st = ens.EnsembleStatistics(4, 2, 3, (4, 8), np.zeros(4), 1, "cpu")
    with pytest.raises(RuntimeError, match="CUDA only"):
        st.accumulate(torch.zeros(8, 3, 4, 8), torch.zeros(2, 3, 4, 8))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pred, truth, lat = fixture()
    B, N, V, H, W = pred.shape
    lo, hi = (0, 2) if rank == 0 else (2, 3)                          # ragged shards: 2 ICs and 1 IC
    st = ens.EnsembleStatistics(N, hi - lo, V, (H, W), lat, steps=2, device="cpu")
    st.sums[0] = ens.sums_reference(pred[lo:hi], truth[lo:hi], lat)
    st.sums[1] = ens.sums_reference(pred[lo:hi] * 2, truth[lo:hi] * 2, lat)
    allsums = st.gather()
    sc = st.scores(allsums)
    if rank == 1:                                                     # every rank holds the full result
        q.put({k: v.numpy() for k, v in sc.items()})
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pred, truth, lat = fixture()
    for k, fn in (("rmse", mo.rmse), ("crps", mo.crps), ("ssr", mo.spread_skill)):
        np.testing.assert_allclose(got[k][0], fn(pred, truth, lat).numpy(), rtol=1e-10)
        scale = 1.0 if k == "ssr" else 2.0                            # rmse and crps are homogeneous of degree 1
        np.testing.assert_allclose(got[k][1], scale * fn(pred, truth, lat).numpy(), rtol=1e-10)
