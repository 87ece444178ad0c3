"""numpy restatement of what csrc/fm_policy.cu computes from the packed weight blob (include/fairmarl.h FmGnnConfig):
checks the blob layout and the algebra of the fused graph network (lin1 split into node / type / edge parts, rank-1
edge terms of TransformerConv) against the dense torch modules on the CPU; the kernel itself is checked on the GPU
(tests/test_gpu_policy.py)."""
import numpy as np

H, HEADS = 16, 3
HC, QW = HEADS * H, 3 * HEADS * H + H


def _ln(h, g, b, on):
    if not on:
        return h
    m = h.mean(-1, keepdims=True)
    c = h - m
    v = (c * c).mean(-1, keepdims=True)
    return c / np.sqrt(v + np.float32(1e-5)) * g + b


def gnn_forward_packed(blob, node, adj_env, rep, agent_id, *, embed_layers, conv_layers, aggr, relu, layer_norm, max_edge_dist):
    blob = np.asarray(blob, np.float32)
    node, adj_env = np.asarray(node, np.float32), np.asarray(adj_env, np.float32)
    M, E, NF = node.shape
    KF = NF - 1
    act = (lambda x: np.maximum(x, 0)) if relu else np.tanh
    o = 0

    def take(n, shape):
        nonlocal o
        x = blob[o:o + n].reshape(shape)
        o += n
        return x
    Wn, T, wd = take(16 * H, (16, H)), take(4 * H, (4, H)), take(H, (H,))
    g1, b1 = take(H, (H,)), take(H, (H,))
    hidden = [(take(H * H, (H, H)), take(H, (H,)), take(H, (H,)), take(H, (H,))) for _ in range(embed_layers)]
    convs = [(take(H * QW, (H, QW)), take(QW, (QW,)), take(HC, (HEADS, H))) for _ in range(conv_layers)]
    assert o == blob.size
    out = np.zeros((M, H), np.float32)
    for m in range(M):
        d = adj_env[m // rep]                                     # d[r, c]: edge r -> c
        mask = (d < np.float32(max_edge_dist)) & (d > 0)
        f = node[m]
        ty = np.clip(f[:, KF].astype(np.int64), 0, 3)
        hn = f[:, :KF] @ Wn[:KF] + T[ty]                          # [E, H]
        h = _ln(act(hn[:, None, :] + d[:, :, None] * wd), g1, b1, layer_norm)        # [r, c, H]
        for Wh, bh, g, b in hidden:
            h = _ln(act(h @ Wh + bh), g, b, layer_norm)
        x = (h * mask[:, :, None]).sum(0)                         # [c, H]
        for W, b, we in convs:
            z = x @ W + b                                         # [E, 160]
            q, k, v, skip = z[:, :HC].reshape(E, HEADS, H), z[:, HC:2 * HC].reshape(E, HEADS, H), z[:, 2 * HC:3 * HC].reshape(E, HEADS, H), z[:, 3 * HC:]
            qe = (q * we[None]).sum(-1)                           # [t, h]
            sc = (np.einsum("thc,shc->hts", q, k) + qe.T[:, :, None] * d.T[None]) * np.float32(0.25)   # [h, t, s], d[s, t]
            sc = np.where(mask.T[None], sc, -np.inf)
            mx = sc.max(-1, keepdims=True)
            ex = np.where(np.isfinite(sc), np.exp(sc - np.where(np.isfinite(mx), mx, 0)), 0)
            tot = ex.sum(-1, keepdims=True)
            al = np.where(tot > 0, ex / np.where(tot > 0, tot, 1), 0)                  # [h, t, s]
            o_ = np.einsum("hts,shc->thc", al, v) + (al * d.T[None]).sum(-1).T[:, :, None] * we[None]
            x = act(o_.mean(1) + skip).astype(np.float32)
        if aggr == 0:
            out[m] = x[int(agent_id[m])]
        elif aggr == 1:
            out[m] = x.mean(0)
        elif aggr == 2:
            out[m] = x.max(0)
        else:
            out[m] = x.sum(0)
    return out
