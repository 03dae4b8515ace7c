"""Golden vectors for the ensemble scores: outputs of the REAL reference functions
(``/root/reference/src/swift/eval/metrics.py``: lat_weighted_crps / lat_weighted_rmse /
lat_weighted_spread_skill_ratio) on a seeded fixture.  Build container only:

    python tests/golden/make_metrics_golden.py

The module imports ``ezpz`` and ``xarray`` for its command line; both are stubbed (no arithmetic in them).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def fixture(seed=0, B=3, N=5, V=4, H=16, W=32):
    g = torch.Generator().manual_seed(seed)
    truth = torch.randn(B, V, H, W, generator=g, dtype=torch.float64) * 3 + 1
    pred = truth.unsqueeze(1) + torch.randn(B, N, V, H, W, generator=g, dtype=torch.float64) * \
        torch.tensor([0.5, 1.0, 2.0, 0.1], dtype=torch.float64).view(1, 1, V, 1, 1) + 0.2
    lat = np.linspace(-88.59375, 88.59375, H)
    return pred, truth, lat


def main():
    for name in ("ezpz", "xarray"):
        sys.modules[name] = types.ModuleType(name)
    sys.path.insert(0, "/root/reference/src")
    from swift.eval import metrics as ref

    pred, truth, lat = fixture()
    vars = [f"v{i}" for i in range(pred.shape[2])]
    out = {}
    for fn in (ref.lat_weighted_crps, ref.lat_weighted_rmse, ref.lat_weighted_spread_skill_ratio):
        out.update({k: float(v) for k, v in fn(pred, truth, vars, lat, "6h").items()})
    keys = sorted(out)
    np.savez(os.path.join(HERE, "metrics.npz"), keys=np.array(keys), values=np.array([out[k] for k in keys]),
             fixture=np.array([0, *pred.shape]))
    print(f"wrote metrics.npz with {len(keys)} scores")


if __name__ == "__main__":
    main()
