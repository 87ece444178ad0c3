"""Policy-side edge list on the GPU: drop-in for ``TransformerConvNet.process_adj``
(onpolicy/algorithms/utils/gnn_new.py:381-413)."""
from __future__ import annotations

from fair_marl_b200 import _lib


def process_adj(adj, max_edge_dist: float, repeat: int = 1, inclusive: bool = False, return_offsets: bool = False):
    """adj: float32 CUDA tensor [num_graphs, E, E] (or [E, E]).

    Returns ``(edge_index int64 [2, nnz], edge_attr float32 [nnz, 1])`` exactly as the reference:
    mask ``(adj < max_edge_dist) & (adj > 0)``, (b, i, j) lexicographic order, node ids offset by
    ``b * E``.  ``repeat=N`` emits each graph N times consecutively, which equals the reference
    applied to the ``[B*N, E, E]`` batch the policy sees (``adj`` is identical for the N agents of an
    env) while reading each env's matrix once.  This call synchronises once to size the result.
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    if not (torch.is_tensor(adj) and adj.is_cuda):
        raise _lib.FairMarlError("process_adj needs a CUDA tensor (no CPU path)")
    single = adj.dim() == 2
    if single:
        adj = adj[None]
    assert adj.dim() == 3 and adj.shape[-1] == adj.shape[-2]         # gnn_new.py:388-389
    adj = adj.to(torch.float32).contiguous()
    G, E = int(adj.shape[0]), int(adj.shape[1])
    dev = adj.device
    cap = G * repeat * E * E                    # a distance matrix has a zero diagonal, a general input may not
    with torch.cuda.device(dev):
        offsets = torch.empty(G * repeat + 1, dtype=torch.int64, device=dev)
        edge_index = torch.empty((2, max(cap, 1)), dtype=torch.int64, device=dev)
        edge_attr = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
        nnz = torch.zeros(1, dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.fm_edge_list(dev.index or 0, adj.data_ptr(), G, E, float(max_edge_dist), int(inclusive),
                                    int(repeat), cap, offsets.data_ptr(), edge_index.data_ptr(),
                                    edge_attr.data_ptr(), nnz.data_ptr(), stream), "fm_edge_list")
        n = int(nnz.item())
    ei = edge_index[:, :n]
    ea = edge_attr[:n]
    if single:
        ei = ei.contiguous()
    else:
        ea = ea.unsqueeze(1)
    if single:
        ea = ea.unsqueeze(1)
    return (ei, ea, offsets) if return_offsets else (ei, ea)
