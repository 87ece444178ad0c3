"""Formation-family oracle pinned against fixtures recorded from the unmodified reference
(oracle/make_formation_golden.py): one step from every recorded pre-state, outputs / info / post-state at 1e-12."""
from dataclasses import fields

import numpy as np
import pytest

from oracle.formation import NODE_FEAT_DIM, OBS_DIM, FormationOracle, FormationState
from oracle.make_formation_golden import CONFIGS, load, set_walls, state_from
from oracle.navgraph import INFO_KEYS

TOL = 1e-12


def _close(a, b, name):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    fin = np.isfinite(b)
    assert (np.isfinite(a) == fin).all(), name
    err = np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1.0)
    assert err.size == 0 or err.max() <= TOL, f"{name}: {err.max():.3e} at {np.argmax(err)}"


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_step_matches_reference(name):
    cfg, g = load(name)
    pre = state_from(g, "pre_")
    T = pre.pos.shape[0]
    orc = FormationOracle(cfg, T)
    orc.set_state(pre)
    set_walls(orc, g, "pre_")
    out = orc.step(g["actions"], autoreset=False)
    assert out["obs"].shape[-1] == OBS_DIM and out["node_obs"].shape[-1] == NODE_FEAT_DIM
    for k in ("obs", "node_obs", "adj", "reward"):
        _close(out[k], g["out_" + k], k)
    assert (out["done"] == g["out_done"]).all()
    for k in INFO_KEYS:
        _close(out["info"][k], g["info_" + k], k)
    post, ref = orc.get_state(), state_from(g, "post_")
    for f in fields(FormationState):
        if f.name == "episode":
            continue
        a, b = getattr(post, f.name), getattr(ref, f.name)
        if f.name in ("goal_match", "step", "status"):
            assert (a == b).all(), f.name
        else:
            _close(a, b, f.name)
    # the fixtures exercise what this family adds
    assert g["post_status"].any() and (g["pre_occupied"] == 1).any() and g["out_done"].any(axis=1).sum() > 10
    if cfg.assignment == "fair":
        assert (g["pre_goal_match"] != g["post_goal_match"]).any()        # per-step re-assignment changed a match
    else:
        assert (g["pre_goal_match"] == g["post_goal_match"]).all()        # fixed for the episode (set at reset)
    rare = ("status_latched", "contact_force_suppressed", "vacated_goal", "info_unlatched")
    if cfg.num_walls:
        rare += ("wall_end_cap", "wall_force_above_1", "wall_box_hit")
    for branch in rare + (("subset_index_quirk",) if cfg.assignment == "fair" else ()):
        assert orc.branch_hits.get(branch, 0) > 0, branch                 # rare paths of the state machine are in the fixture


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_reset_outputs_match_reference(name):
    """env.reset() outputs (obs / node rows / adj) from the recorded post-reset states."""
    cfg, g = load(name)
    st = state_from(g, "reset_")
    orc = FormationOracle(cfg, st.pos.shape[0])
    # the recorded state is the one AFTER env.reset() observed (the observation updates the occupancy table, :866-931);
    # random_scenario leaves it cleared (:471, :230)
    pre = st.copy()
    pre.occupied[:] = 0.0
    pre.goal_history[:] = -1.0
    orc.set_state(pre)
    set_walls(orc, g, "reset_")
    r = orc.reset(mask=np.zeros(st.pos.shape[0], bool))                   # observe only
    _close(orc.get_state().occupied, st.occupied, "occupancy after the reset observation")
    _close(orc.get_state().goal_history, st.goal_history, "goal_history after the reset observation")
    _close(r["obs"], g["reset_obs"], "reset obs")
    _close(r["node_obs"], g["reset_node_obs"], "reset node_obs")
    _close(r["adj"], g["reset_adj"], "reset adj")
    # acceptance rules of random_scenario hold on the reference's own placements
    for b in range(st.pos.shape[0]):
        for i in range(cfg.num_agents):
            assert not orc._obstacle_collision(b, st.pos[b, i]) and not orc._obstacle_collision(b, st.landmark_pos[b, i])
            for j in range(i):
                assert np.linalg.norm(st.pos[b, i] - st.pos[b, j]) >= 1.05 * 0.1
                assert np.linalg.norm(st.landmark_pos[b, i] - st.landmark_pos[b, j]) >= cfg.goal_clearance * 0.1
    if cfg.assignment == "optimal":                                       # min-sum matching of the reset positions
        from scipy.optimize import linear_sum_assignment
        for b in range(st.pos.shape[0]):
            costs = np.linalg.norm(st.pos[b][:, None] - st.landmark_pos[b][None], axis=-1)
            assert (st.goal_match[b] == linear_sum_assignment(costs)[1]).all()
    if cfg.assignment == "random":
        assert (np.sort(st.goal_match, axis=1) == np.arange(cfg.num_agents)).all()
        assert len({tuple(m) for m in st.goal_match}) > 1


def test_own_reset_obeys_the_rules_and_rollout_runs():
    from oracle.formation import FormationConfig
    cfg = FormationConfig(num_agents=4, num_obstacles=3, episode_length=6)
    orc = FormationOracle(cfg, 16, seed=3)
    orc.reset()
    s = orc.get_state()
    for b in range(16):
        for i in range(4):
            assert not orc._obstacle_collision(b, s.pos[b, i]) and not orc._obstacle_collision(b, s.landmark_pos[b, i])
            for j in range(i):
                assert np.linalg.norm(s.pos[b, i] - s.pos[b, j]) >= 1.05 * 0.1
                assert np.linalg.norm(s.landmark_pos[b, i] - s.landmark_pos[b, j]) >= 1.2 * 0.1
    assert (s.episode == 1).all() and not s.status.any()
    rng = np.random.default_rng(0)
    resets = 0
    for t in range(13):
        out = orc.step(rng.integers(0, 5, (16, 4)))
        resets += int(out["reset"].sum())
        assert np.isfinite(out["obs"]).all() and np.isfinite(out["node_obs"]).all()
    assert resets == 32 and (orc.get_state().episode == 3).all()


@pytest.mark.parametrize("mode", ["optimal", "random"])
def test_own_reset_of_the_base_scenarios(mode):
    from oracle.formation import FormationConfig
    cfg = FormationConfig(num_agents=4, num_obstacles=2, fairness_reward=False, assignment=mode, episode_length=5)
    orc = FormationOracle(cfg, 24, seed=9)
    orc.reset()
    s = orc.get_state()
    assert (np.sort(s.goal_match, axis=1) == np.arange(4)).all()
    for b in range(24):
        for i in range(4):
            for j in range(i):
                assert np.linalg.norm(s.landmark_pos[b, i] - s.landmark_pos[b, j]) >= 1.5 * 0.1
    if mode == "random":
        assert len({tuple(m) for m in s.goal_match}) > 6                  # permutations differ across envs
    out = orc.step(np.random.default_rng(1).integers(0, 5, (24, 4)))
    assert np.isfinite(out["reward"]).all()
    with pytest.raises(ValueError):
        FormationOracle(FormationConfig(assignment=mode), 2)              # these scenarios have no fairness term


def test_share_vec_env_tuple_matches_the_reference_tuples():
    """Host glue of B200FormationVecEnv (pure numpy, no device): oracle outputs packed by share_vec_env_tuple equal the
    tuples the reference env returned when the fixture was recorded; one-hot decode follows environment.py:301-311."""
    from fair_marl_b200.formation import decode_onehot_actions, share_vec_env_tuple
    cfg, g = load("formation_n3_o3_fafr")
    pre = state_from(g, "pre_")
    T, N, E = pre.pos.shape[0], cfg.num_agents, cfg.num_entities
    orc = FormationOracle(cfg, T)
    orc.set_state(pre)
    idx = decode_onehot_actions(np.eye(5)[g["actions"]], T, N)
    assert idx.dtype == np.int32 and (idx == g["actions"]).all() and (decode_onehot_actions(g["actions"], T, N) == idx).all()
    out = orc.step(idx, autoreset=False)
    rows = np.stack([out["info"][k] for k in INFO_KEYS], axis=-1)
    tup = share_vec_env_tuple(dict(obs=out["obs"], node_obs=out["node_obs"], adj_env=out["adj"], reward=out["reward"],
                                   done=out["done"].astype(np.uint8), info=rows))
    obs, agent_id, node, adj, rew, done, infos = tup
    assert obs.shape == (T, N, OBS_DIM) and node.shape == (T, N, E, NODE_FEAT_DIM) and adj.shape == (T, N, E, E)
    assert agent_id.shape == (T, N, 1) and (agent_id[5, :, 0] == np.arange(N)).all()          # get_id: global ids (:1017)
    _close(adj[:, 2], g["out_adj"], "adj of agent 2 == the env's matrix")
    _close(rew, g["out_reward"], "rewards")
    assert done.dtype == bool and (done == g["out_done"]).all()
    assert len(infos) == T and len(infos[0]) == N and list(infos[7][1]) == list(INFO_KEYS)
    assert abs(infos[7][1]["Dist_to_goal"] - g["info_Dist_to_goal"][7, 1]) < 1e-12
    assert len(share_vec_env_tuple(dict(obs=out["obs"], node_obs=out["node_obs"], adj_env=out["adj"]))) == 4   # reset()
    with pytest.raises(ValueError):
        decode_onehot_actions(np.full((T, N, 5), 0.2), T, N)
    with pytest.raises(ValueError):
        decode_onehot_actions(np.zeros((T, N + 1)), T, N)


@pytest.mark.reference
@pytest.mark.parametrize("mode,fair", [("fair", True), ("optimal", False), ("random", False)])
def test_rare_branches_against_the_live_reference(mode, fair):
    """Build container only: directed states the recorded fixtures do not reach -- every goal marked occupied with all
    agents far away (the table is cleared and the agent 'goes to itself', :951 / :1266), partial occupancies with an exact
    1.0 -- injected into the LIVE unmodified reference and stepped once; oracle == reference at 1e-12."""
    from oracle.formation import FormationConfig
    from oracle.make_formation_golden import extract_state, inject_state, make_reference_env
    N, O, B = 3, 3, 12
    cfg = FormationConfig(num_agents=N, num_obstacles=O, goal_rew=30.0, collision_rew=30.0, min_obs_dist=0.3,
                          fairness_reward=fair, assignment=mode)
    orc = FormationOracle(cfg, B, seed=11)
    orc.reset()
    st = orc.get_state()
    rng = np.random.default_rng(3)
    st.occupied[: B // 2] = 1.0
    st.goal_history[: B // 2] = rng.integers(0, N, (B // 2, N)).astype(np.float64)
    st.landmark_pos[: B // 4] += 3.0
    st.occupied[B // 2:] = rng.random((B - B // 2, N))
    st.occupied[B // 2:, 0] = 1.0
    orc.set_state(st)
    a = rng.integers(0, 5, (B, N))
    out = orc.step(a, autoreset=False)
    post = orc.get_state()
    assert orc.branch_hits.get("all_occupied_cleared", 0) > 0
    env, sc = make_reference_env(cfg, seed=1)
    for b in range(B):
        env.reset()
        inject_state(env, sc, st, b)
        oh = np.eye(5)[a[b]]
        ob, ag_id, node, adj, rew, done, info = env.step([oh[i] for i in range(N)])
        _close(out["obs"][b], np.array(ob), f"obs[{b}]")
        _close(out["node_obs"][b], np.array(node), f"node_obs[{b}]")
        _close(out["adj"][b], np.array(adj)[0], f"adj[{b}]")
        _close(out["reward"][b], np.array(rew, dtype=np.float64).reshape(N), f"reward[{b}]")
        assert (out["done"][b] == np.array(done)).all()
        for k in INFO_KEYS:
            _close(out["info"][k][b], np.array([info[i][k] for i in range(N)]), f"{k}[{b}]")
        ref_post = extract_state(env, sc)
        for f in ("pos", "vel", "p_dist", "occupied", "goal_history", "goal_reached", "dists_to_goal", "times_required"):
            _close(getattr(post, f)[b], getattr(ref_post, f)[0], f"{f}[{b}]")
        assert (post.status[b] == ref_post.status[0]).all()


def test_own_reset_with_walls():
    from oracle.formation import FormationConfig
    cfg = FormationConfig(num_agents=3, num_obstacles=2, num_walls=2, episode_length=4)
    orc = FormationOracle(cfg, 32, seed=5)
    out = orc.reset()
    assert out["node_obs"].shape == (32, 3, 10, NODE_FEAT_DIM) and (out["node_obs"][:, :, 8:, 12] == 3.0).all()
    s = orc.get_state()
    assert set(np.unique(orc.wall_orient)) == {0, 1} and (orc.wall_axis[:, 0] == -orc.wall_axis[:, 1]).all()
    assert ((orc.wall_len >= 0.1 - 1e-6) & (orc.wall_len <= 0.4 + 1e-6)).all()          # U(0.2, 0.8) * ws / 4
    assert ((np.abs(orc.wall_axis) >= 0.2 - 1e-6) & (np.abs(orc.wall_axis) <= 0.9 + 1e-6)).all()
    hits0 = orc.branch_hits.get("wall_box_hit", 0)
    for b in range(32):
        for i in range(3):
            assert not orc._obstacle_collision(b, s.pos[b, i]) and not orc._obstacle_collision(b, s.landmark_pos[b, i])
    assert orc.branch_hits.get("wall_box_hit", 0) == hits0
    lens = orc.wall_len.copy()
    for t in range(5):
        out = orc.step(np.random.default_rng(t).integers(0, 5, (32, 3)))
    assert out["reset"].all() or (orc.get_state().episode >= 2).all()
    assert (orc.wall_len != lens).any()                                                  # redrawn per reset in this family
    with pytest.raises(ValueError):
        FormationOracle(FormationConfig(num_walls=3), 2)
