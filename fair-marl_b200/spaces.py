"""Minimal stand-ins for ``gym.spaces`` (gym is not a dependency).

The runner only looks at ``__class__.__name__`` ('Box' / 'Discrete'), ``.shape`` and ``.n``
(onpolicy/utils/util.py:32-53, graph_mpe_runner.py:420-431), so these carry exactly that.
"""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {np.dtype(self.dtype).name})"


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()

    def __repr__(self):
        return f"Discrete({self.n})"
