"""CPU: the packed weight blob of the fused graph network + the kernel's algebra (numpy twin, tests/gnn_emul.py)
reproduce the dense torch modules (which tests/test_policy.py pins to the reference's GR_Actor / GR_Critic)."""
import numpy as np
import pytest
import torch

from gnn_emul import gnn_forward_packed


def _case(E, NF, aggr_name, **kw):
    import fair_marl_b200 as fm
    from fair_marl_b200.policy import DenseGNNBase, pack_gnn_weights
    cfg = fm.PolicyConfig(node_feat_dim=NF, **kw)
    torch.manual_seed(E * 100 + NF)
    graph_aggr = "node" if aggr_name == "node" else "global"
    if aggr_name != "node":
        cfg.global_aggr_type = aggr_name
    gnn = DenseGNNBase(cfg, graph_aggr).eval()
    with torch.no_grad():
        for p in gnn.parameters():                               # away from the init's symmetries (biases 0, LN gains 1)
            p.add_(0.3 * torch.randn_like(p))
    B, rep = 5, 3
    M = B * rep
    g = torch.Generator().manual_seed(7)
    pos = torch.rand(B, E, 2, generator=g) * 2 - 1
    adj_env = torch.cdist(pos, pos).float()
    adj_env[0, 1, :] = 5.0; adj_env[0, :, 1] = 5.0; adj_env[0, 1, 1] = 0.0      # an isolated node: no incoming edges
    node = torch.randn(M, E, NF, generator=g)
    node[..., -1] = torch.randint(0, 3, (M, E), generator=g).float()
    aid = torch.randint(0, E, (M, 1), generator=g)
    with torch.no_grad():
        ref = gnn(node, adj_env.repeat_interleave(rep, dim=0), aid).numpy()
    out = gnn_forward_packed(pack_gnn_weights(gnn).numpy(), node.numpy(), adj_env.numpy(), rep, aid[:, 0].numpy(),
                             embed_layers=cfg.embed_layer_N, conv_layers=1 + cfg.gnn_layer_N,
                             aggr={"node": 0, "mean": 1, "max": 2, "add": 3}[aggr_name], relu=cfg.gnn_use_ReLU,
                             layer_norm=cfg.use_feature_normalization, max_edge_dist=cfg.max_edge_dist)
    err = np.abs(out - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 2e-5, err.max()


@pytest.mark.parametrize("E,NF,aggr,kw", [
    (9, 11, "node", {}), (9, 11, "mean", {}), (9, 13, "node", {}),
    (17, 11, "max", dict(embed_layer_N=2, gnn_layer_N=1, embed_use_ReLU=False, gnn_use_ReLU=False, use_feature_normalization=False)),
    (11, 11, "add", dict(embed_layer_N=0, gnn_layer_N=0)),
])
def test_packed_blob_and_kernel_algebra_match_the_dense_modules(E, NF, aggr, kw):
    _case(E, NF, aggr, **kw)
