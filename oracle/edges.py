"""Edge lists as the reference builds them (TEST INFRASTRUCTURE ONLY).

* ``process_adj`` -- the policy-side edge list the GNN consumes
  (onpolicy/algorithms/utils/gnn_new.py:381-413): mask ``(adj < max_edge_dist) & (adj > 0)``
  on the float32 ``adj`` batch ``[B', E, E]``, ``nonzero`` in (b, i, j) lexicographic order,
  ``edge_index = [b*E + i ; b*E + j]`` (int64 ``[2, nnz]``), ``edge_attr = adj[b, i, j]``
  as ``[nnz, 1]``.
* ``update_graph`` -- the env-side list used only by the renderer
  (navigation_graph.py:1037-1056): ``<=`` instead of ``<``; CSR -> COO order is the same
  row-major order.
"""
from __future__ import annotations

import numpy as np


def process_adj(adj: np.ndarray, max_edge_dist: float, inclusive: bool = False):
    """adj float32 [B', E, E] -> (edge_index int64 [2, nnz], edge_attr float32 [nnz, 1])."""
    adj = np.asarray(adj, dtype=np.float32)
    assert adj.ndim == 3 and adj.shape[-1] == adj.shape[-2]
    thr = np.float32(max_edge_dist)
    near = (adj <= thr) if inclusive else (adj < thr)
    mask = near & (adj > 0)
    masked = adj * mask.astype(np.float32)          # gnn_new.py:392-393
    b, i, j = np.nonzero(masked)                    # lexicographic (b, i, j), gnn_new.py:398
    E = adj.shape[-1]
    edge_index = np.stack([b * E + i, b * E + j]).astype(np.int64)
    edge_attr = masked[b, i, j][:, None]
    return edge_index, edge_attr


def update_graph(dist_mag: np.ndarray, max_edge_dist: float):
    """Single env, float64 ``cached_dist_mag`` [E, E] -> (edge_list [2, nnz], edge_weight [nnz])."""
    d = np.asarray(dist_mag)
    connect = (d <= max_edge_dist) & (d > 0)
    row, col = np.nonzero(connect)
    return np.stack([row, col]), d[row, col]
