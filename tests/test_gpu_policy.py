"""Fused graph-network forward (csrc/fm_policy.cu, fm_gnn_forward) on the GPU, through the C ABI:

* against the frozen outputs of the UNMODIFIED reference GR_Actor / GR_Critic (tests/golden/policy_n3_o3.npz, produced
  by oracle/make_policy_golden.py): gnn features, log-probs, values at 1e-5 * max(|ref|, 1), actions (argmax) exact;
* against the dense torch modules (pinned to the reference by tests/test_policy.py) on random graphs over the shape family:
  node / mean / max / add aggregation, 11- and 13-wide node rows, Tanh, no LayerNorm, 0..2 hidden embed layers, E = 9 / 11 / 17,
  graphs with isolated nodes, adjacency shared by the N ego graphs of an env (what the step kernel writes) or one per graph;
* the rollout collector with the fused path == the collector with the dense path (actions exact on a deterministic policy).
"""
import ast
import os
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5


def _close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1.0)
    return bool((err <= tol).all()), float(err.max())


def _flat(a):
    return torch.as_tensor(a.reshape((-1,) + a.shape[2:])).cuda()


def test_fused_policy_matches_reference_golden():
    from fair_marl_b200.policy import DenseGraphActor, DenseGraphCritic, PolicyConfig, load_reference_state_dict
    from oracle.make_policy_golden import BASE_ARGS
    z = np.load(os.path.join(GOLDEN, "policy_n3_o3.npz"))
    N, O, B, _ = (int(v) for v in z["meta"])
    cfg = PolicyConfig.from_args(Namespace(**{**BASE_ARGS, "num_agents": N, **ast.literal_eval(str(z["overrides"]))}))
    actor, critic = DenseGraphActor(cfg).eval().cuda(), DenseGraphCritic(cfg).eval().cuda()
    load_reference_state_dict(actor, {k[6:]: z[k] for k in z.files if k.startswith("actor/")})
    load_reference_state_dict(critic, {k[7:]: z[k] for k in z.files if k.startswith("critic/")})
    adj_env = torch.as_tensor(z["in_adj_env"]).cuda().contiguous()
    obs, node, aid = _flat(z["in_obs"]), _flat(z["in_node_obs"]), _flat(z["in_agent_id"])
    rnn, masks = _flat(z["in_rnn_states"]), _flat(z["in_masks"])
    assert actor.gnn_base.fused_available(node) is False            # grad mode: the torch path
    with torch.no_grad():
        assert actor.gnn_base.fused_available(node) and critic.gnn_base.fused_available(node)
        act, logp, h = actor(obs, node, None, aid, rnn, masks, deterministic=True, adj_env=(adj_env, N))
        feat = actor.gnn_base(node, None, aid, adj_env=(adj_env, N))
        val, hc = critic(None, node, None, aid, rnn, masks, adj_env=(adj_env, N))
    assert np.array_equal(act.cpu().numpy(), z["out_actions"])
    for name, got in (("gnn_feat", feat), ("action_log_probs", logp), ("rnn_out", h), ("values", val), ("rnn_out_critic", hc)):
        ok, err = _close(got.cpu().numpy(), z["out_" + name])
        assert ok, (name, err)


@pytest.mark.parametrize("E,NF,aggr,rep,kw", [
    (9, 11, "node", 3, {}), (9, 11, "mean", 3, {}), (9, 13, "node", 3, {}), (9, 13, "mean", 1, {}),
    (17, 11, "max", 7, dict(embed_layer_N=2, gnn_layer_N=1, embed_use_ReLU=False, gnn_use_ReLU=False, use_feature_normalization=False)),
    (11, 11, "add", 3, dict(embed_layer_N=0, gnn_layer_N=0)),
    (7, 11, "node", 2, {}), (19, 11, "mean", 7, {}), (13, 13, "node", 5, {}),
])
def test_fused_gnn_matches_dense_modules(E, NF, aggr, rep, kw):
    import fair_marl_b200 as fm
    from fair_marl_b200.policy import DenseGNNBase
    cfg = fm.PolicyConfig(node_feat_dim=NF, **kw)
    if aggr != "node":
        cfg.global_aggr_type = aggr
    torch.manual_seed(E * 100 + NF)
    gnn = DenseGNNBase(cfg, "node" if aggr == "node" else "global").eval().cuda()
    with torch.no_grad():
        for p in gnn.parameters():
            p.add_(0.3 * torch.randn_like(p))
    B = 701                                                           # several CTAs, ragged last one
    M = B * rep
    g = torch.Generator(device="cuda").manual_seed(7)
    pos = torch.rand(B, E, 2, generator=g, device="cuda") * 2 - 1
    adj_env = torch.cdist(pos, pos).float().contiguous()
    adj_env[0, 1, :] = 5.0; adj_env[0, :, 1] = 5.0; adj_env[0, 1, 1] = 0.0     # a node without edges
    adj_env[1] = 5.0                                                             # a graph without edges
    node = torch.randn(M, E, NF, generator=g, device="cuda")
    node[..., -1] = torch.randint(0, 3, (M, E), generator=g, device="cuda").float()
    aid = torch.randint(0, E, (M, 1), generator=g, device="cuda")
    import copy
    with torch.no_grad():
        adj_rep = adj_env.repeat_interleave(rep, dim=0)
        ref32 = gnn(node, adj_rep, aid)
        ref64 = copy.deepcopy(gnn).double()(node.double(), adj_rep.double(), aid)       # the same function in float64
        out = gnn(node, None, aid, adj_env=(adj_env, rep))
        out2 = gnn.forward_fused(node, adj_env, rep, aid)             # deterministic: bit-identical on a second launch
    # The weights are perturbed away from the init, so the LayerNorms see small variances and fp32 round-off is amplified:
    # the fused kernel is held to the float64 truth as tightly as torch's own fp32 evaluation is (x3, floor 1e-5).
    err = lambda x: float((torch.abs(x.double() - ref64) / torch.clamp(ref64.abs(), min=1.0)).max())
    e_fused, e_torch = err(out), err(ref32)
    assert e_fused <= max(3.0 * e_torch, 1e-5), (e_fused, e_torch)
    assert torch.equal(out, out2)


def test_collector_fused_equals_dense():
    import fair_marl_b200 as fm
    from oracle.navgraph import NavConfig
    from parity_util import sim_config_from
    cfg = NavConfig(num_agents=3, num_obstacles=3, episode_length=8)
    outs = []
    for fused in (True, False):
        env = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=160, seed=11)
        pc = fm.PolicyConfig(num_agents=3)
        torch.manual_seed(0)
        actor, critic = fm.DenseGraphActor(pc).to(env.device).eval(), fm.DenseGraphCritic(pc).to(env.device).eval()
        with torch.no_grad():
            actor.action_out.weight.mul_(100.0)
        col = fm.RolloutCollector(env, actor, critic, deterministic=True, fused=fused)
        assert col.fused is fused
        col.warmup(); col.run()
        torch.cuda.synchronize()
        outs.append((col.buffer.actions_env.clone(), col.buffer.value_preds.clone(), col.buffer.obs.clone()))
        env.close()
    assert torch.equal(outs[0][0], outs[1][0])                        # same actions => same trajectories
    assert torch.equal(outs[0][2], outs[1][2])
    ok, err = _close(outs[0][1].cpu().numpy(), outs[1][1].cpu().numpy(), 2e-5)
    assert ok, err
