"""numpy Philox4x32-10 (TEST INFRASTRUCTURE ONLY).

The reference draws reset positions from numpy's global MT19937
(navigation_graph.py:271-275, :393-395, :491-493; seeded per worker process,
train_mpe.py:31 -> environment.py:192-196).  Bit-matching that stream is not a
goal (BASELINE.json north_star: parity is "from identical states and actions").
The device reset draws from a counter-based Philox4x32-10 stream keyed by
(seed, global env index, episode index, draw index) so that results do not
depend on how envs are sharded over GPUs; this file is the same generator in
numpy so that device resets can be reproduced bit-exactly by the oracle.

Algorithm: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11),
Philox-4x32 with 10 rounds; constants as in Random123 / cuRAND.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable uint32 arrays.  Returns 4 uint32 arrays."""
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(
        *[np.asarray(a, dtype=np.uint32) for a in (c0, c1, c2, c3, k0, k1)])
    c0, c1, c2, c3, k0, k1 = [a.copy() for a in (c0, c1, c2, c3, k0, k1)]
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0.astype(np.uint64)
            p1 = PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = k0 + PHILOX_W0
            k1 = k1 + PHILOX_W1
    return c0, c1, c2, c3


def u01_24(bits):
    """uint32 -> float32 uniform on [0, 1) with 24 random bits (exact in fp32)."""
    return (np.asarray(bits, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
