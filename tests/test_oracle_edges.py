import numpy as np

from oracle.edges import process_adj, update_graph


def test_process_adj_matches_torch_semantics():
    torch = __import__("torch")
    rng = np.random.default_rng(0)
    adj = rng.random((6, 9, 9)).astype(np.float32) * 2
    adj[:, np.arange(9), np.arange(9)] = 0
    adj[0, 1, 2] = 1.0           # boundary: strict < excludes
    ei, ea = process_adj(adj, 1.0)
    # gnn_new.py:381-413 restated with torch ops (torch_geometric is not needed for this function)
    t = torch.from_numpy(adj)
    m = ((t < 1.0) & (t > 0)).float()
    t = t * m
    nz = t.nonzero(as_tuple=False)
    attr = t[nz[:, 0], nz[:, 1], nz[:, 2]]
    b = nz[:, 0] * 9
    ref = torch.stack([b + nz[:, 1], b + nz[:, 2]], dim=0)
    assert (ei == ref.numpy()).all() and ei.dtype == np.int64
    assert (ea[:, 0] == attr.numpy()).all()
    ei2, _ = process_adj(adj, 1.0, inclusive=True)
    assert ei2.shape[1] == ei.shape[1] + 1


def test_update_graph_row_major():
    d = np.array([[0, .5, 2.], [.5, 0, 1.], [2., 1., 0]])
    el, ew = update_graph(d, 1.0)
    assert el.tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]
    assert ew.tolist() == [.5, .5, 1., 1.]
