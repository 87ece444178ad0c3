"""Boundary conformance of ``B200GraphVecEnv`` against the GraphSubprocVecEnv contract
(env_wrappers.py:951-1025) and the host-buffer C-ABI path."""
import numpy as np
import pytest

from oracle.navgraph import INFO_KEYS, NavConfig, NavGraphOracle
from parity_util import assert_close, compare_step_outputs, device_state_to_nav, sim_config_from

pytestmark = pytest.mark.gpu


def test_reference_api_shapes_dtypes_and_autoreset():
    import fair_marl_b200 as fm
    from argparse import Namespace
    args = Namespace(num_agents=3, num_landmarks=3, num_obstacles=3, world_size=2, max_speed=2, collision_rew=30,
                     goal_rew=30, min_dist_thresh=0.05, episode_length=25, fair_rew=1, zeroshift=5, max_edge_dist=1,
                     collaborative=False, num_walls=0, graph_feat_type="relative", num_scripted_agents=0,
                     scenario_name="navigation_graph", use_dones=False, fair_wt=1, n_rollout_threads=128, seed=1,
                     env_name="GraphMPE")
    env = fm.make_train_env(args)
    B, N, E = 128, 3, 9
    assert env.num_envs == B
    assert env.observation_space[0].__class__.__name__ == "Box" and env.observation_space[0].shape == (7,)
    assert env.share_observation_space[0].shape == (21,)
    assert env.action_space[0].__class__.__name__ == "Discrete" and env.action_space[0].n == 5
    assert env.node_observation_space[0].shape == (E, 11) and env.adj_observation_space[0].shape == (E, E)
    assert env.edge_observation_space[0].shape == (1,) and env.agent_id_observation_space[0].shape == (1,)
    assert env.share_agent_id_observation_space[0].shape == (N,)
    obs, ag_id, node, adj = env.reset()
    assert obs.shape == (B, N, 7) and ag_id.shape == (B, N, 1) and node.shape == (B, N, E, 11) and adj.shape == (B, N, E, E)
    assert (ag_id[:, :, 0] == np.arange(N)).all()
    assert (adj[:, 0] == adj[:, 1]).all() and (adj[:, 0] == np.swapaxes(adj[:, 0], 1, 2)).all()
    rng = np.random.default_rng(0)
    for t in range(25):
        acts = np.eye(5)[rng.integers(0, 5, (B, N))]                 # graph_mpe_runner.py:429-431
        obs, ag_id, node, adj, rew, done, infos = env.step(acts)
        assert rew.shape == (B, N) and done.shape == (B, N) and done.dtype == bool
        assert done.all() == (t == 24) and done.any() == (t == 24)
        if t == 3:                                                    # info rows are written on terminal steps only:
            with pytest.raises(RuntimeError, match="info_every_step"):    # stale values must not look like this step's
                infos[0]
            assert infos.as_array().shape[:2] == (B, N)
    # terminal step: obs are the NEW episode's (velocity 0, fairness 0), infos are terminal
    assert (obs[..., 0:2] == 0).all() and (obs[..., 6] == 0).all()
    assert len(infos) == B and len(infos[0]) == N and set(infos[0][0]) == set(INFO_KEYS)
    assert abs(infos[5][1]["Time_taken"] - 2.5) < 1e-5
    assert infos[5][1]["individual_reward"] == pytest.approx(float(rew[5, 1]))
    st = env.get_state()
    assert (st["step"].cpu().numpy() == 0).all() and (st["episode"].cpu().numpy() == 2).all()
    env.close()


@pytest.mark.parametrize("N,B,lanes", [(4, 200, "1"), (4, 1000, "3"), (3, 700, "4"), (7, 300, "2")])
def test_host_api_equals_tensor_api_and_oracle(N, B, lanes, monkeypatch):
    """numpy in / numpy out (fm_step_host, or fm_step_host_lane: env-range lanes whose action rows are converted while the
    previous lane's results cross the bus) against the tensor API and the oracle, across an auto-reset, one-hot and index
    actions."""
    import fair_marl_b200 as fm
    monkeypatch.setenv("FM_HOST_LANES", lanes)
    cfg = NavConfig(num_agents=N, num_obstacles=2)
    e_host = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=B, seed=9)
    e_dev = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=B, seed=9)
    orc = NavGraphOracle(cfg, B, seed=9)
    o_h = e_host.reset()
    o_d = e_dev.reset_tensor()
    assert (o_h[0] == o_d["obs"].cpu().numpy()).all() and (o_h[3] == o_d["adj"].cpu().numpy()).all()
    rng = np.random.default_rng(2)
    import torch
    for t in range(28):
        orc.set_state(device_state_to_nav(e_dev.get_state()))
        a = rng.integers(0, 5, (B, N))
        obs, ag, node, adj, rew, done, infos = e_host.step(np.eye(5)[a] if t % 3 else a)
        d = e_dev.step_tensor(torch.as_tensor(a, dtype=torch.int32, device="cuda"))
        assert (obs == d["obs"].cpu().numpy()).all() and (node == d["node_obs"].cpu().numpy()).all()
        assert (adj == d["adj"].cpu().numpy()).all() and (rew == d["reward"].cpu().numpy()).all()
        assert (done == d["done"].cpu().numpy()).all()
        ref = orc.step(actions=a)
        compare_step_outputs(dict(obs=obs, node_obs=node, adj=adj[:, 0], reward=rew, done=done), ref, cfg)
        if done.all():
            rows = infos.as_array()
            for k, key in enumerate(INFO_KEYS):
                if "by" not in key:
                    assert_close(rows[..., k], ref["info"][key], key)
    e_host.close(); e_dev.close()


def test_episode_statistics_match_oracle_sums():
    import fair_marl_b200 as fm
    import torch
    cfg = NavConfig(num_agents=3, num_obstacles=3)
    B = 333
    env = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=B, seed=4)
    orc = NavGraphOracle(cfg, B, seed=4)
    env.reset_tensor()
    rng = np.random.default_rng(3)
    rew_sum = np.zeros(3)
    info_sum = np.zeros((3, 14))
    for t in range(50):
        orc.set_state(device_state_to_nav(env.get_state()))
        a = rng.integers(0, 5, (B, 3))
        env.step_tensor(torch.as_tensor(a, dtype=torch.int32, device="cuda"))
        ref = orc.step(actions=a)
        rew_sum += ref["reward"].sum(0)
        if ref["reset"].all():
            for k, key in enumerate(INFO_KEYS):
                info_sum[:, k] += ref["info"][key].sum(0)
    v = env.read_stats(clear=True).cpu().numpy()
    np.testing.assert_allclose(v[0:3], rew_sum, rtol=1e-5)
    got = v[3:45].reshape(3, 14)
    for k, key in enumerate(INFO_KEYS):
        if "by" in key:
            continue
        np.testing.assert_allclose(got[:, k], info_sum[:, k], rtol=2e-5, atol=1e-3, err_msg=key)
    assert v[45] == 2 * B and v[46] == 50 * B
    assert (env.read_stats().cpu().numpy() == 0).all()
    s = fm.EpisodeStats(3, device=torch.device("cuda", 0))
    s.all_reduce_async(torch.as_tensor(v, device="cuda"))
    summ = s.summary()
    assert summ["episodes"] == 2 * B and len(summ["Dist_to_goal"]) == 3
    env.close()


def test_errors_are_python_exceptions():
    import fair_marl_b200 as fm
    from fair_marl_b200._lib import FairMarlError
    with pytest.raises(FairMarlError):
        fm.B200GraphVecEnv(fm.SimConfig(num_agents=40), num_envs=8)
    env = fm.B200GraphVecEnv(fm.SimConfig(), num_envs=8)
    with pytest.raises(ValueError):
        env.step(np.zeros((8, 3, 4)))
    with pytest.raises(NotImplementedError):
        env.render()
    env.close()


@pytest.mark.parametrize("B", [1, 7, 8, 9, 31, 33])
def test_ragged_batch_sizes(B):
    """Batches that do not fill a warp / are not multiples of 4 take the unaligned store path."""
    import fair_marl_b200 as fm
    import torch
    for N, O in [(3, 3), (7, 3), (9, 2), (16, 3)]:
        cfg = NavConfig(num_agents=N, num_obstacles=O)
        env = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=B, seed=21)
        orc = NavGraphOracle(cfg, B, seed=21)
        env.reset_tensor()
        a = np.random.default_rng(B).integers(0, 5, (B, N))
        orc.set_state(device_state_to_nav(env.get_state()))
        out = env.step_tensor(torch.as_tensor(a, dtype=torch.int32, device="cuda"))
        ref = orc.step(actions=a)
        o = {k: v.cpu().numpy() for k, v in out.items() if hasattr(v, "cpu")}
        o["adj"] = o["adj_env"]
        compare_step_outputs(o, ref, cfg)
        env.close()


def test_global_graph_features_through_the_reference_api():
    """graph_feat_type='global' (navigation_graph.py:1058-1077): node_obs [B,N,E,7] = [vel, pos, goal, type] in world
    coordinates, the same rows for every agent; checked against the oracle through the numpy API."""
    import fair_marl_b200 as fm
    cfg = NavConfig(num_agents=3, num_obstacles=3, graph_feat_type="global")
    B, N, E = 70, 3, 9
    env = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=B, seed=4)
    assert env.node_observation_space[0].shape == (E, 7)
    orc = NavGraphOracle(cfg, B, seed=4)
    obs, ag_id, node, adj = env.reset()
    ref = orc.reset()
    assert node.shape == (B, N, E, 7)
    assert_close(node, ref["node_obs"], "reset node_obs (global)")
    rng = np.random.default_rng(1)
    for t in range(27):
        orc.set_state(device_state_to_nav(env.get_state()))
        a = rng.integers(0, 5, (B, N))
        obs, ag_id, node, adj, rew, done, infos = env.step(a)
        want = orc.step(actions=a)
        assert node.shape == (B, N, E, 7) and (node[:, 0] == node[:, 1]).all()
        assert_close(node, want["node_obs"], f"node_obs (global) step {t}")
        assert (node[..., 6] == np.array([0, 0, 0, 1, 1, 1, 2, 2, 2], dtype=np.float32)).all()
        assert_close(rew, want["reward"], f"reward step {t}")
    env.close()


def test_dummy_vec_env_returns_reset_count():
    """GraphDummyVecEnv.step_wait (env_wrappers.py:911-928) appends ``reset_count``."""
    import fair_marl_b200 as fm
    cfg = NavConfig(num_agents=3, num_obstacles=3, episode_length=4)
    env = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=1, seed=3, dummy_vec_env=True)
    env.reset()
    counts = []
    for t in range(9):
        out = env.step(np.zeros((1, 3), dtype=np.int64))
        assert len(out) == 8
        counts.append(out[7])
    assert counts == [0, 0, 0, 1, 0, 0, 0, 1, 0]
    env.close()


def test_edge_list_tensor_matches_process_adj_without_sync():
    import torch
    import fair_marl_b200 as fm
    from oracle.edges import process_adj as oracle_process_adj
    cfg = NavConfig(num_agents=7, num_obstacles=3)
    B, N, E = 50, 7, 17
    env = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=B, seed=6)
    o = env.reset_tensor()
    o = env.step_tensor(torch.zeros(B, N, dtype=torch.int32, device=env.device))
    r = env.edge_list_tensor(o["adj_env"], repeat=N)
    n = int(r["nnz"].item())
    adj_policy = o["adj"].reshape(B * N, E, E).cpu().numpy()             # what the policy's process_adj would see
    ei, ea = oracle_process_adj(adj_policy, cfg.max_edge_dist)
    assert n == ei.shape[1]
    assert np.array_equal(r["edge_index"][:, :n].cpu().numpy(), ei)
    assert np.array_equal(r["edge_attr"][:n].cpu().numpy(), ea[:, 0].astype(np.float32))
    assert int(r["offsets"][-1].item()) == n
    env.close()


@pytest.mark.parametrize("N,O,B,feat,mapping", [(3, 3, 200, "relative", "aw"), (3, 3, 4097, "relative", "group"), (7, 3, 130, "relative", "group"),
                                                (4, 2, 64, "global", "group"), (16, 3, 33, "relative", "group")])
def test_soa_observation_equals_api_layout_transposed(N, O, B, feat, mapping):
    """fm_observe_soa (one plane per value, envs fastest; csrc/fm_soa.cu) against fm_observe (API layout) after a reset and in
    the middle of an episode: the same bits, transposed."""
    import torch
    import fair_marl_b200 as fm
    cfg = NavConfig(num_agents=N, num_obstacles=O, graph_feat_type=feat, episode_length=6)
    env = fm.B200GraphVecEnv(sim_config_from(cfg, mapping=mapping), num_envs=B, seed=3)
    env.reset_tensor()
    g = torch.Generator(device="cuda").manual_seed(2)
    for t in range(9):                                                # crosses an auto-reset; the fairness channel switches source
        if t in (0, 1, 4, 8):
            api, soa = env.observe_tensor(), env.observe_soa_tensor()
            assert torch.equal(soa["obs"], api["obs"].permute(1, 2, 0)), t
            assert torch.equal(soa["node_obs"], api["node_obs"].permute(1, 2, 3, 0)), t
            assert torch.equal(soa["adj"], api["adj_env"].permute(1, 2, 0)), t
        env.step_tensor(torch.randint(0, 5, (B, N), generator=g, device="cuda", dtype=torch.int32))
    env.close()


def test_finite_guard_flags_poisoned_envs():
    import torch
    import fair_marl_b200 as fm
    cfg = NavConfig(num_agents=3, num_obstacles=3)
    B = 300
    env = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=B, seed=1)
    env.reset_tensor()
    flags, count = env.check_finite()
    assert int(count) == 0 and not bool(flags.any())
    st = env.get_state()
    st["vel"][7, 1, 0] = float("nan")
    st["pos"][123, 2, 1] = float("inf")
    st["p_dist"][299, 0] = float("-inf")
    env.set_state(st)
    flags, count = env.check_finite()
    assert int(count) == 3 and flags.nonzero().flatten().tolist() == [7, 123, 299]
    env.close()
