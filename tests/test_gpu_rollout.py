"""Device-resident rollout (RolloutCollector + DeviceRolloutBuffer + dense policy) on the GPU:
the step kernel writes the rollout buffer's slabs directly; a replay of the recorded actions through a second
simulator (ring slabs) must reproduce every slab bit for bit, and the buffer bookkeeping must follow
GMPERunner.insert (graph_mpe_runner.py:438-488)."""
import numpy as np
import pytest
import torch

from oracle.navgraph import NavConfig
from parity_util import sim_config_from

pytestmark = pytest.mark.gpu


def _make(cfg, B, seed):
    import fair_marl_b200 as fm
    env = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=B, seed=seed)
    pc = fm.PolicyConfig(num_agents=cfg.num_agents)
    torch.manual_seed(0)
    actor, critic = fm.DenseGraphActor(pc).to(env.device).eval(), fm.DenseGraphCritic(pc).to(env.device).eval()
    with torch.no_grad():
        actor.action_out.weight.mul_(100.0)          # gain 0.01 leaves the logits flat; make the policy state dependent
    return fm, env, actor, critic


@pytest.mark.parametrize("N,O,B", [(3, 3, 96), (7, 3, 40)])
def test_rollout_writes_buffer_in_place_and_replays_bitwise(N, O, B):
    cfg = NavConfig(num_agents=N, num_obstacles=O, episode_length=25)
    fm, env, actor, critic = _make(cfg, B, seed=5)
    col = fm.RolloutCollector(env, actor, critic, deterministic=True, max_graphs=64 * N)   # several policy chunks
    buf = col.buffer
    col.warmup(); col.run()
    torch.cuda.synchronize()
    env2 = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=B, seed=5)
    r = env2.reset_tensor()
    assert torch.equal(r["obs"], buf.obs[0]) and torch.equal(r["node_obs"], buf.node_obs[0]) and torch.equal(r["adj_env"], buf.adj_env[0])
    for t in range(25):
        o = env2.step_tensor(buf.actions_env[t])
        assert torch.equal(o["obs"], buf.obs[t + 1]), t
        assert torch.equal(o["node_obs"], buf.node_obs[t + 1]), t
        assert torch.equal(o["adj_env"], buf.adj_env[t + 1]), t
        assert torch.equal(o["reward"], buf.rewards[t, :, :, 0]), t
        assert torch.equal(o["done"], buf.dones[t].bool()), t
    # ---- bookkeeping (GMPERunner.insert): masks = 1 - done, rnn states cleared at done, views
    done = buf.dones.bool()
    assert not done[:24].any() and done[24].all()
    assert torch.equal(buf.masks[1:], (~done).float().unsqueeze(-1))
    assert (buf.rnn_states[25] == 0).all() and (buf.rnn_states[24] != 0).any()
    assert (buf.active_masks == 1).all()
    assert buf.adj.shape == (26, B, N, 2 * N + O, 2 * N + O) and buf.adj.stride(2) == 0
    assert torch.equal(buf.share_obs[3, :, 1], buf.obs[3].reshape(B, -1))
    assert torch.equal(buf.actions[:, :, :, 0], buf.actions_env.float())
    # ---- the recorded actions / log-probs / values are the policy's on the recorded slabs (recomputed in one chunk)
    t = 7
    M = B * N
    E = 2 * N + O
    adj = buf.adj[t].reshape(M, E, E)
    with torch.no_grad():
        a, lp, h = actor(buf.obs[t].view(M, -1), buf.node_obs[t].view(M, E, -1), adj, buf.agent_id[t].reshape(M, 1),
                         buf.rnn_states[t].view(M, 1, -1), buf.masks[t].view(M, 1), deterministic=True)
        v, _ = critic(None, buf.node_obs[t].view(M, E, -1), adj, buf.agent_id[t].reshape(M, 1),
                      buf.rnn_states_critic[t].view(M, 1, -1), buf.masks[t].view(M, 1))
    assert torch.allclose(lp.view(B, N, 1), buf.action_log_probs[t], rtol=1e-4, atol=1e-5)
    assert torch.allclose(v.view(B, N, 1), buf.value_preds[t], rtol=1e-4, atol=1e-5)
    assert (a.view(B, N) == buf.actions_env[t]).float().mean() > 0.99        # argmax ties may flip between batch shapes
    assert torch.allclose(h.view(B, N, 1, -1), buf.rnn_states[t + 1], rtol=1e-4, atol=1e-5)
    col.finish()
    assert torch.isfinite(buf.returns).all()
    env.close(); env2.close()


@pytest.mark.parametrize("N", [3, 7])
def test_captured_graph_episode_equals_eager_episode(N):
    """N = 7 runs the group kernel: the eager loop consumes prefetched next-episode placements, the captured one
    computes every reset inside the step kernel (no side-stream work under capture); same bits either way."""
    cfg = NavConfig(num_agents=N, num_obstacles=3)
    runs = []
    for use_graph in (False, True):
        fm, env, actor, critic = _make(cfg, 128, seed=11)
        col = fm.RolloutCollector(env, actor, critic, deterministic=True)
        col.warmup(); col.run(); col.buffer.after_update()
        if use_graph:
            col.capture()
        col.run(); col.buffer.after_update(); col.run()
        torch.cuda.synchronize()
        b = col.buffer
        runs.append([x.clone() for x in (b.obs, b.node_obs, b.adj_env, b.rewards, b.dones, b.actions_env, b.rnn_states, b.masks,
                                         b.value_preds, b.action_log_probs)])
        env.close()
    for x, y in zip(*runs):
        assert torch.equal(x, y)


def test_gpu_policy_matches_cpu_policy():
    """Same module, same inputs from the simulator: CUDA forward vs CPU forward within 1e-4 (TF32 is off by default)."""
    cfg = NavConfig(num_agents=3, num_obstacles=3)
    fm, env, actor, critic = _make(cfg, 64, seed=2)
    o = env.reset_tensor()
    for _ in range(4):
        o = env.step_tensor(torch.randint(0, 5, (64, 3), dtype=torch.int32, device=env.device))
    M, E = 64 * 3, 9
    args = (o["obs"].view(M, -1), o["node_obs"].view(M, E, -1), o["adj"].reshape(M, E, E), o["agent_id"].reshape(M, 1),
            torch.zeros(M, 1, 64, device=env.device), torch.ones(M, 1, device=env.device))
    with torch.no_grad():
        a, lp, h = actor(*args, deterministic=True)
        import copy
        a_c, lp_c, h_c = copy.deepcopy(actor).cpu()(*[x.cpu() for x in args], deterministic=True)
    assert torch.allclose(lp.cpu(), lp_c, rtol=1e-4, atol=1e-5) and torch.allclose(h.cpu(), h_c, rtol=1e-4, atol=1e-5)
    env.close()


def test_step_tensor_out_validation():
    cfg = NavConfig(num_agents=3, num_obstacles=3)
    fm, env, actor, critic = _make(cfg, 32, seed=1)
    buf = fm.DeviceRolloutBuffer(25, 32, 3, 9, device=env.device)
    env.reset_tensor(out=buf.env_views(0, with_step=False))
    bad = buf.env_views(1)
    bad["adj"] = buf.adj[1]                                   # [B,N,E,E] expanded view: not the per-env array
    with pytest.raises(ValueError):
        env.step_tensor(torch.zeros(32, 3, dtype=torch.int32, device=env.device), out=bad)
    env.close()
