"""The drop-in boundary against the REAL reference package (CPU, only where /root/reference exists -- this container):
``swift.models.precond.PassPrecond`` instantiates ``swift_b200.swinv2.SwinV2`` through its hydra ``_target_``, a real
reference ``state_dict`` loads strictly, and ``isinstance(net.model, swift.models.swinv2.SwinV2)`` holds, so
``train.py:271-313`` builds the same optimiser parameter groups.  Runs in a subprocess: the class hierarchy of
``swift_b200.swinv2`` is fixed when the module is first imported, and the import shims (omegaconf / ezpz / hydra are not in
this image; no arithmetic in them) must be in place before that."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"

SCRIPT = textwrap.dedent('''
    import sys, torch
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, %(root)r + "/tests/golden")
    import make_golden
    make_golden.install_shims()
    from swift.models.precond import PassPrecond
    from swift.models.swinv2 import SwinV2 as RefSwinV2
    import swift_b200.swinv2 as ours
    from swift_b200 import synthetic as syn
    cfg = syn.SWIFT_TINY
    n_img = cfg["out_channels"]
    kw = dict(window_size=cfg["window_size"], shift_size=cfg["shift_size"], patch_size=cfg["patch_size"], depth=cfg["depth"],
              dim=cfg["dim"], heads=cfg["heads"])
    mk = lambda target: PassPrecond(dict(_target_=target, **kw), img_resolution=cfg["img_resolution"], img_channels=n_img,
                                    condition_channels=cfg["in_channels"] - n_img, auxiliary_dim=1)
    ref, net = mk("swift.models.swinv2.SwinV2"), mk("swift_b200.swinv2.SwinV2")
    assert type(net.model) is ours.SwinV2 and isinstance(net.model, RefSwinV2), type(net.model).__mro__
    res = net.load_state_dict(ref.state_dict(), strict=True)            # generate.py:225-226 / trainer.py:522-535
    assert not res.missing_keys and not res.unexpected_keys
    assert [tuple(p.shape) for p in net.parameters()] == [tuple(p.shape) for p in ref.parameters()]
    assert [n for n, _ in net.named_parameters()] == [n for n, _ in ref.named_parameters()]
    # the Muon / AuxAdam grouping of train.py:286-294 comes out identical
    group = lambda m: sorted(n for n, p in m.named_parameters() if p.ndim >= 2 and "transformer" in n)
    assert group(net) == group(ref) and len(group(net)) > 0
    assert net.img_channels == ref.img_channels and list(net.img_resolution) == list(ref.img_resolution)
    try:
        net(torch.zeros(1, n_img, *cfg["img_resolution"]), torch.tensor([0.5]), torch.zeros(1, cfg["in_channels"] - n_img,
            *cfg["img_resolution"]), 0.6)
    except RuntimeError as e:
        assert "CUDA" in str(e), e                                      # no CPU fallback: loud failure
    else:
        raise AssertionError("the CPU call must fail loudly")
    print("BOUNDARY_OK")
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only present in the build container")
def test_real_passprecond_instantiates_and_loads_our_module():
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "BOUNDARY_OK" in r.stdout, r.stdout + r.stderr
