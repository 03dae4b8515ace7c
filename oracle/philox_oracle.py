"""numpy restatement of the rollout noise stream (swift_b200/csrc/rollout.cu)  --  TEST INFRASTRUCTURE ONLY.

The reference draws latents with ``torch.randn(..., generator=torch.Generator(device).manual_seed(member))``
(generate.py:83, generating/factory.py:52-56); any i.i.d. N(0,1) stream is equivalent for the algorithm.  The product
uses a counter-based stream so that a trajectory's noise does not depend on batching or world size:

    (x0, x1, x2, x3) = Philox4x32-10(counter = (i//4 lo, i//4 hi, step, 0), key = (seed lo, seed hi))
    u(x) = ((x >> 8) + 0.5) / 2^24
    z[4q .. 4q+3] = Box-Muller(u(x0), u(x1)), Box-Muller(u(x2), u(x3))

Philox4x32-10 is the published Random123 algorithm (Salmon et al., SC'11): 10 rounds of
(c0,c1,c2,c3) <- (hi(M1*c2)^c1^k0, lo(M1*c2), hi(M0*c0)^c3^k1, lo(M0*c0)), keys bumped by the Weyl constants.
Integer arithmetic is checked bit-exactly through the resulting floats (a single wrong bit decorrelates the output).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    for _ in range(10):
        p0 = M0 * c0.astype(np.uint64)
        p1 = M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        with np.errstate(over="ignore"):
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def _u01(x):
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def normal(seed: int, step: int, n: int) -> np.ndarray:
    """The n (multiple of 4) N(0,1) values of one trajectory at one step, float32."""
    assert n % 4 == 0
    q = np.arange(n // 4, dtype=np.uint64)
    x = philox4x32_10((q & np.uint64(0xFFFFFFFF)).astype(np.uint32), (q >> np.uint64(32)).astype(np.uint32),
                      np.full(q.shape, step, np.uint32), np.zeros(q.shape, np.uint32),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    r0 = np.sqrt(np.float32(-2.0) * np.log(_u01(x[0])))
    r1 = np.sqrt(np.float32(-2.0) * np.log(_u01(x[2])))
    t0 = np.float32(6.283185307179586) * _u01(x[1])
    t1 = np.float32(6.283185307179586) * _u01(x[3])
    z = np.stack([r0 * np.cos(t0), r0 * np.sin(t0), r1 * np.cos(t1), r1 * np.sin(t1)], axis=-1)
    return z.reshape(-1).astype(np.float32)
