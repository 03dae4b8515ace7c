"""Golden vectors for the optimiser step: the REAL reference's ``muon_update`` / ``adam_update``
(/root/reference/src/swift/training/optimizers/muon.py, imported as a plain file: it depends on torch only) on seeded
matrices.  Run here (the reference does not travel to the GPU box):  python tests/golden/make_muon_golden.py"""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/src/swift/training/optimizers/muon.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "muon.npz")


def main():
    spec = importlib.util.spec_from_file_location("ref_muon", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(0)
    torch.set_num_threads(1)
    out = {}
    for name, shape in (("wide", (64, 96)), ("tall", (160, 64)), ("square", (88, 88)), ("scale", (1, 12, 1, 1))):
        g = torch.randn(shape) * 1e-3
        mom = torch.randn(shape) * 1e-3
        out[f"{name}_grad"], out[f"{name}_mom"] = g.numpy().copy(), mom.numpy().copy()
        g2, m2 = g.clone(), mom.clone()
        upd = ref.muon_update(g2, m2, beta=0.95)
        out[f"{name}_update"] = upd.float().reshape(shape).numpy().copy()
        out[f"{name}_mom_new"] = m2.numpy().copy()
    p, g = torch.randn(300), torch.randn(300) * 1e-2
    b1, b2 = torch.randn(300) * 1e-3, torch.rand(300) * 1e-4
    out["adam_p"], out["adam_g"], out["adam_b1"], out["adam_b2"] = (t.numpy().copy() for t in (p, g, b1, b2))
    b1n, b2n = b1.clone(), b2.clone()
    upd = ref.adam_update(g, b1n, b2n, 3, (0.9, 0.95), 1e-10)
    out["adam_update"], out["adam_b1_new"], out["adam_b2_new"] = upd.numpy().copy(), b1n.numpy().copy(), b2n.numpy().copy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
