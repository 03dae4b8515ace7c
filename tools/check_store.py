"""Read back a forecast store written by `python -m swift_b200.generate` and report per-(ic, member) completeness."""
import sys

import numpy as np

sys.path.insert(0, ".")
from swift_b200.store import ForecastStore

st = ForecastStore.open(sys.argv[1])
a = np.asarray(st.read_all())                     # [ic, member, lead, channel, H, W]
filled = np.abs(a).reshape(a.shape[0], a.shape[1], a.shape[2], -1).max(-1) > 0
print("layout", st.layout, "shape", a.shape, "filled (ic, member, lead):", int(filled.sum()), "of", filled.size,
      "finite", bool(np.isfinite(a).all()))
print("members of IC 0 differ:", bool(np.abs(a[0, 0, -1] - a[0, 1, -1]).max() > 0) if a.shape[1] > 1 else None,
      "| lead 0 shared by members:", bool(np.array_equal(a[0, 0, 0], a[0, -1, 0])))
sys.exit(0 if filled.all() and np.isfinite(a).all() else 1)
