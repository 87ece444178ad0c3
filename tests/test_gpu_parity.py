"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden vectors the
UNMODIFIED reference produced.  All tests here need a B200 (``-m gpu``).

Tolerance: |dev - ref| <= 1e-5 * max(|ref|, 1) for positions, velocities, observations, node
features, adjacency and rewards (BASELINE.json north_star).  Bit-exact:
goal assignments, adjacency given the device's own fp32 positions, edge lists, reset placements.
"""
import numpy as np
import pytest

from oracle.edges import process_adj as oracle_process_adj
from oracle.lexifair import lexifair, lexifair_bruteforce_batched, lexifair_descent
from oracle.make_golden import CONFIGS, load, state_from

DEVICE_CONFIGS = sorted(CONFIGS)
from oracle.navgraph import INFO_KEYS, NavConfig, NavGraphOracle
from parity_util import (assert_close, assert_fairness_close, compare_step_outputs, device_state_to_nav,
                         sim_config_from, state_to_device_dict, state_to_fp32)

pytestmark = pytest.mark.gpu


def _env(cfg, B, **kw):
    import fair_marl_b200 as fm
    return fm.B200GraphVecEnv(sim_config_from(cfg, **kw.pop("sim", {})), num_envs=B, **kw)


def _np(out):
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}


def _actions(a):
    import torch
    return torch.as_tensor(np.asarray(a), dtype=torch.int32, device="cuda")


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", DEVICE_CONFIGS)
def test_step_matches_oracle_on_golden_states(name):
    """One step from every recorded reference state (rounded to fp32): device vs float64 oracle."""
    cfg, g = load(name)
    pre = state_to_fp32(state_from(g, "pre_"))
    T = pre.pos.shape[0]
    env = _env(cfg, T, sim=dict(auto_reset=False, info_every_step=True))
    env.set_state(state_to_device_dict(pre))
    out = _np(env.step_tensor(_actions(g["actions"])))
    orc = NavGraphOracle(cfg, T)
    orc.set_state(pre)
    ref = orc.step(actions=g["actions"], autoreset=False)
    out["adj"] = out["adj_env"]
    compare_step_outputs(out, ref, cfg)
    # post-step state
    post = device_state_to_nav(env.get_state())
    rpost = orc.get_state()
    for f in ("pos", "vel", "p_dist", "dists_to_goal", "times_required", "dist_left_to_goal",
              "dist_traveled_mean", "dist_traveled_stddev"):
        assert_close(getattr(post, f), getattr(rpost, f), f)
    for f in ("num_agent_collisions", "num_obstacle_collisions", "step", "goal_match"):
        assert (getattr(post, f) == getattr(rpost, f)).all(), f
    # info rows
    info = out["info"]
    for k, key in enumerate(INFO_KEYS):
        if key in ("Mean_by_variance", "Time_mean_by_stddev"):
            assert_fairness_close(info[..., k], ref["info"][key], key)
        else:
            assert_close(info[..., k], ref["info"][key], key)
    env.close()


@pytest.mark.parametrize("name", ["n3_o3_fafr", "n7_o3_fafr", "n16_o3_fafr", "n3_o3_w2", "n4_o2_w1"])
def test_step_matches_reference_golden_directly(name):
    """Device outputs against the reference's own float64 outputs (inputs rounded to fp32 on the way
    in, so smooth quantities only: obs[0:6], node_obs, adj)."""
    cfg, g = load(name)
    pre = state_to_fp32(state_from(g, "pre_"))
    env = _env(cfg, pre.pos.shape[0], sim=dict(auto_reset=False))
    env.set_state(state_to_device_dict(pre))
    out = _np(env.step_tensor(_actions(g["actions"])))
    assert_close(out["obs"][..., :6], g["out_obs"][..., :6], "obs", rtol=2e-5)
    assert_close(out["node_obs"], g["out_node_obs"], "node_obs", rtol=2e-5)
    assert_close(out["adj_env"], g["out_adj"], "adj", rtol=2e-5)
    assert (out["done"] == g["out_done"]).all()
    env.close()


@pytest.mark.parametrize("name", ["n3_o3_fafr", "n7_o3_fafr", "n5_o0_fafr", "n16_o3_fafr"])
def test_adjacency_and_edges_bit_exact_on_device_positions(name):
    """Stage 2 of the two-stage check: the oracle evaluated on the device's own
    post-step fp32 positions gives bit-identical adj (float64 -> float32) and edge lists."""
    import fair_marl_b200 as fm
    cfg, g = load(name)
    pre = state_to_fp32(state_from(g, "pre_"))
    T = pre.pos.shape[0]
    env = _env(cfg, T, sim=dict(auto_reset=False))
    env.set_state(state_to_device_dict(pre))
    out = env.step_tensor(_actions(g["actions"]))
    adj_dev = out["adj_env"].cpu().numpy()
    orc = NavGraphOracle(cfg, T)
    orc.set_state(device_state_to_nav(env.get_state()))
    adj_ref = orc.distance_matrix().astype(np.float32)
    assert (adj_dev == adj_ref).all(), np.abs(adj_dev - adj_ref).max()
    for repeat in (1, cfg.num_agents):
        ei, ea = fm.process_adj(out["adj_env"], cfg.max_edge_dist, repeat=repeat)
        rei, rea = oracle_process_adj(np.repeat(adj_ref, repeat, axis=0), cfg.max_edge_dist)
        assert ei.dtype.is_floating_point is False and tuple(ei.shape) == rei.shape
        assert (ei.cpu().numpy() == rei).all()
        assert (ea.cpu().numpy() == rea).all()
    ei, ea = fm.process_adj(out["adj_env"], cfg.max_edge_dist, inclusive=True)
    rei, rea = oracle_process_adj(adj_ref, cfg.max_edge_dist, inclusive=True)
    assert (ei.cpu().numpy() == rei).all() and (ea.cpu().numpy() == rea).all()
    env.close()


@pytest.mark.parametrize("N,O,B", [(3, 3, 512), (7, 3, 256), (5, 0, 200), (16, 3, 96), (4, 2, 130), (1, 1, 33), (2, 0, 31)])
def test_reset_bit_exact_vs_oracle(N, O, B):
    """Device reset == oracle reset (same Philox stream, same acceptance rules): positions and
    assignments bit-exact; also the reference's rejection rules hold."""
    cfg = NavConfig(num_agents=N, num_obstacles=O)
    env = _env(cfg, B, seed=1234, env_offset=77)
    out = _np(env.reset_tensor())
    st = device_state_to_nav(env.get_state())
    orc = NavGraphOracle(cfg, B, seed=1234, env_offset=77)
    ref = orc.reset()
    rs = orc.get_state()
    assert (st.pos == rs.pos).all() and (st.landmark_pos == rs.landmark_pos).all()
    assert (st.obstacle_pos == rs.obstacle_pos).all()
    assert (st.goal_match == rs.goal_match).all()
    assert (st.vel == 0).all() and (st.p_dist == 0).all() and (st.step == 0).all() and (st.episode == 1).all()
    assert (st.dists_to_goal == -1).all() and (st.times_required == -1).all()
    assert_close(out["obs"], ref["obs"], "reset obs")
    assert_close(out["node_obs"], ref["node_obs"], "reset node_obs")
    assert (out["adj_env"] == ref["adj"].astype(np.float32)).all()
    # a-14 acceptance rules
    dmin = 1.05 * (0.05 + 0.05)
    for a in range(N):
        for b in range(a + 1, N):
            assert (np.linalg.norm(st.pos[:, a] - st.pos[:, b], axis=-1) >= dmin).all()
            assert (np.linalg.norm(st.landmark_pos[:, a] - st.landmark_pos[:, b], axis=-1) >= dmin).all()
        for k in range(O):
            assert (np.linalg.norm(st.pos[:, a] - st.obstacle_pos[:, k], axis=-1) >= dmin).all()
    assert (np.abs(st.pos) <= 1).all() and (np.abs(st.landmark_pos) <= 0.8).all()
    # a second reset draws a different episode and still matches
    env.reset_tensor()
    orc.reset()
    st2 = device_state_to_nav(env.get_state())
    assert (st2.pos == orc.s.pos).all() and (st2.goal_match == orc.s.goal_match).all()
    assert not (st2.pos == st.pos).all()
    env.close()


@pytest.mark.parametrize("N,O,B,steps", [(3, 3, 256, 60), (7, 3, 64, 30), (16, 3, 16, 27)])
def test_rollout_with_autoreset_matches_oracle(N, O, B, steps):
    """Random-action rollout across auto-resets.  Every step the oracle is re-synchronised to the
    device state (single-step parity from identical states), then both step; on auto-reset steps the
    new episode's placements and assignment must be bit-exact, reward / done stay terminal."""
    cfg = NavConfig(num_agents=N, num_obstacles=O, goal_rew=30.0, collision_rew=30.0)
    env = _env(cfg, B, seed=7)
    orc = NavGraphOracle(cfg, B, seed=7)
    env.reset_tensor()
    rng = np.random.default_rng(0)
    n_resets = 0
    for t in range(steps):
        orc.set_state(device_state_to_nav(env.get_state()))
        a = rng.integers(0, 5, (B, N))
        out = _np(env.step_tensor(_actions(a)))
        ref = orc.step(actions=a, autoreset=True)
        out["adj"] = out["adj_env"]
        compare_step_outputs(out, ref, cfg)
        post, rpost = device_state_to_nav(env.get_state()), orc.get_state()
        if ref["reset"].any():
            n_resets += 1
            assert ref["reset"].all() and out["done"].all()
            assert (post.pos == rpost.pos).all() and (post.goal_match == rpost.goal_match).all()
            assert (post.step == 0).all() and (post.episode == rpost.episode).all()
            assert_close(post.min_time, rpost.min_time, "min_time")
        else:
            assert_close(post.pos, rpost.pos, "pos")
    assert n_resets == steps // cfg.episode_length
    env.close()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 11, 16, 17, 25, 32])
def test_assignment_bit_exact(n):
    """Stand-alone kernel (d) vs the exact CPU solvers, incl. tie-heavy integer costs."""
    import fair_marl_b200 as fm
    rng = np.random.default_rng(n)
    num = 600 if n <= 8 else 60
    costs = rng.random((num, n, n))
    costs[num // 2:] = rng.integers(0, 4, (num - num // 2, n, n))        # ties
    got = fm.lexifair_batched(costs=costs)
    want = lexifair_bruteforce_batched(costs) if n <= 7 else np.stack([lexifair_descent(c) for c in costs])
    assert got.dtype == np.int32 and (got == want).all()
    # from positions (float64 cdist of float32 points)
    ap = rng.uniform(-1, 1, (num, n, 2)).astype(np.float32)
    gp = (0.8 * rng.uniform(-1, 1, (num, n, 2))).astype(np.float32)
    d = ap.astype(np.float64)[:, :, None, :] - gp.astype(np.float64)[:, None, :, :]
    c2 = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
    got2 = fm.lexifair_batched(agent_pos=ap, goal_pos=gp)
    assert (got2 == lexifair(c2)).all()


def test_assignment_reference_fixed_instance():
    import fair_marl_b200 as fm
    goals = np.array([[0., -0.5], [0.45, -0.5], [0.9, -0.5]])
    agents = np.array([[-0.9, -0.9], [-0.9, 0.], [-0.9, 0.9]])
    d = agents[:, None, :] - goals[None, :, :]
    x, objs = fm.solve_fair_assignment(np.sqrt((d ** 2).sum(-1)))      # marl_fair_assign.py:63-70
    assert (np.where(x == 1)[1] == [2, 1, 0]).all()
    np.testing.assert_allclose(objs, [1.843909, 1.664332, 1.439618], atol=1e-6)


def test_pair_dist_bit_exact():
    """The branch-free float64 square root behind every distance (dsqrt_fast, fm_device.cuh) is
    correctly rounded: fm_pair_dist == numpy's float64 sqrt(dx*dx + dy*dy) bit for bit, over random
    points, coincident points, tiny and large separations and exact squares."""
    import fair_marl_b200 as fm
    rng = np.random.default_rng(7)
    n = 1 << 20
    a = rng.uniform(-1.5, 1.5, (n, 2)).astype(np.float32)
    b = rng.uniform(-1.5, 1.5, (n, 2)).astype(np.float32)
    b[:1000] = a[:1000]                                                     # coincident -> 0
    b[1000:3000] = a[1000:3000] + rng.uniform(-1e-6, 1e-6, (2000, 2)).astype(np.float32)
    a[3000:4000] *= np.float32(1e-20); b[3000:4000] *= np.float32(1e-20)    # tiny magnitudes
    a[4000:5000] *= np.float32(1e15); b[4000:5000] *= np.float32(1e15)      # large magnitudes
    a[5000:6000] = 0; b[5000:6000, 0] = rng.integers(1, 2000, 1000); b[5000:6000, 1] = 0   # exact squares
    a[6000:7000] = 0; b[6000:7000, 0] = 3 * rng.integers(1, 500, 1000); b[6000:7000, 1] = 4 * (b[6000:7000, 0] / 3)
    dev = fm.pair_dist(a, b)
    d = a.astype(np.float64) - b.astype(np.float64)
    ref = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    assert dev.dtype == np.float64 and np.array_equal(dev.view(np.uint64), ref.view(np.uint64))


def test_onehot_and_index_actions_agree():
    import torch
    cfg = NavConfig()
    B = 300
    e1, e2 = _env(cfg, B, seed=3), _env(cfg, B, seed=3)
    e1.reset_tensor(); e2.reset_tensor()
    a = np.random.default_rng(1).integers(0, 5, (B, 3))
    o1 = _np(e1.step_tensor(_actions(a)))
    o2 = _np(e2.step_tensor(torch.as_tensor(np.eye(5, dtype=np.float32)[a], device="cuda")))
    for k in ("obs", "node_obs", "adj_env", "reward"):
        assert (o1[k] == o2[k]).all(), k
    e1.close(); e2.close()


def test_shards_equal_one_batch_bitwise():
    """results are independent of the sharding (RNG keyed by global env index)."""
    cfg = NavConfig(num_agents=3, num_obstacles=3)
    B, parts = 1000, [(0, 250), (250, 137), (387, 613)]
    rng = np.random.default_rng(5)
    acts = rng.integers(0, 5, (30, B, 3))
    full = _env(cfg, B, seed=11)
    shards = [_env(cfg, c, seed=11, env_offset=o) for o, c in parts]
    outs_f = [_np(full.reset_tensor())]
    outs_s = [[_np(s.reset_tensor()) for s in shards]]
    for t in range(30):
        outs_f.append(_np(full.step_tensor(_actions(acts[t]))))
        outs_s.append([_np(s.step_tensor(_actions(acts[t, o:o + c]))) for s, (o, c) in zip(shards, parts)])
    for f, ss in zip(outs_f, outs_s):
        for k in ("obs", "node_obs", "adj_env") + (("reward", "done") if "reward" in f else ()):
            assert (f[k] == np.concatenate([s[k] for s in ss])).all(), k
    # statistics: the sum over shards equals the single batch (up to summation order)
    sf = full.read_stats().cpu().numpy()
    ss = sum(s.read_stats().cpu().numpy() for s in shards)
    np.testing.assert_allclose(sf, ss, rtol=1e-9)
    assert sf[15 * 3] == B and sf[15 * 3 + 1] == 30 * B
    full.close()
    [s.close() for s in shards]


@pytest.mark.parametrize("other", ["aw"])
@pytest.mark.parametrize("N,O,B,W", [(3, 3, 1000, 0), (4, 2, 333, 0), (2, 0, 65, 0), (1, 1, 40, 0), (3, 0, 129, 0),
                                     (3, 3, 1000, 2), (3, 3, 333, 1), (4, 2, 200, 1)])       # the agent-warp wall instantiations
def test_mappings_bitwise_equal(N, O, B, W, other):
    """The agent-warp kernels (fm_aw.cu) and the group-per-env kernels (fm_kernels.cu) perform the same
    arithmetic: every output, the state and the statistics agree bit for bit over a rollout with
    auto-resets, goal latches and info rows -- with walls too (wall circle + wall force, box test, wall rows, wall draws)."""
    cfg = NavConfig(num_agents=N, num_obstacles=O, num_walls=W, goal_rew=30.0, collision_rew=30.0, episode_length=7)
    osim = dict(mapping=other)
    e_t = _env(cfg, B, seed=5, env_offset=11, sim=dict(info_every_step=True, **osim))
    e_g = _env(cfg, B, seed=5, env_offset=11, sim=dict(mapping="group", info_every_step=True))
    assert e_t.mapping == osim["mapping"] and e_g.mapping == "group"
    o_t, o_g = _np(e_t.reset_tensor()), _np(e_g.reset_tensor())
    for k in ("obs", "node_obs", "adj_env"):
        assert (o_t[k] == o_g[k]).all(), k
    rng = np.random.default_rng(2)
    for t in range(23):
        a = rng.integers(0, 5, (B, N))
        if t % 3 == 0:
            a[: B // 2] = 0                                     # some agents idle -> equal travelled distances
        if t == 10:                                             # a masked reset in the middle of an episode
            m = torch_mask(B)
            o_t, o_g = _np(e_t.reset_tensor(mask=m)), _np(e_g.reset_tensor(mask=m))
            for k in ("obs", "node_obs", "adj_env"):
                assert np.array_equal(o_t[k], o_g[k], equal_nan=True), ("masked reset", k)
        o_t, o_g = _np(e_t.step_tensor(_actions(a))), _np(e_g.step_tensor(_actions(a)))
        for k in ("obs", "node_obs", "adj_env", "reward", "done", "info"):
            assert np.array_equal(o_t[k], o_g[k], equal_nan=True), (t, k)
    s_t, s_g = e_t.get_state(), e_g.get_state()
    for k in s_t:
        assert np.array_equal(s_t[k].cpu().numpy(), s_g[k].cpu().numpy(), equal_nan=True), k
    st_t, st_g = e_t.read_stats().cpu().numpy(), e_g.read_stats().cpu().numpy()
    np.testing.assert_allclose(st_t, st_g, rtol=1e-12)
    e_t.close(); e_g.close()


def torch_mask(B):
    import torch
    return torch.as_tensor((np.arange(B) % 3 == 1).astype(np.uint8), device="cuda")


@pytest.mark.parametrize("mapping", ["aw"])
def test_mapping_unavailable_raises(mapping):
    import fair_marl_b200 as fm
    with pytest.raises(fm._lib.FairMarlError, match="not compiled"):
        fm.B200GraphVecEnv(fm.SimConfig(num_agents=7, mapping=mapping), num_envs=8)


@pytest.mark.parametrize("variant", ["k1x2", "k3x1", "k3x2"])
def test_group_kernel_staging_variants_are_bitwise_equal(variant, monkeypatch):
    """The node_obs staging of the group kernel has three shapes (two 32-row buffers, one 96-row buffer, two 96-row
    buffers; fm_create picks by what fits, FM_STAGE forces one).  All of them must produce the same bytes."""
    import torch
    cfg = NavConfig(num_agents=7, num_obstacles=3)
    B = 37                                            # ragged: the last warp holds one env
    outs = []
    for force in (None, variant):
        if force:
            monkeypatch.setenv("FM_STAGE", force)
        else:
            monkeypatch.delenv("FM_STAGE", raising=False)
        env = _env(cfg, B, seed=21, sim=dict(mapping="group"))
        o = env.reset_tensor()
        rec = [o["node_obs"].clone(), o["adj_env"].clone()]
        g = torch.Generator(device="cuda").manual_seed(3)
        for t in range(27):
            a = torch.randint(0, 5, (B, 7), generator=g, device="cuda", dtype=torch.int32)
            o = env.step_tensor(a)
            rec += [o["node_obs"].clone(), o["adj_env"].clone(), o["obs"].clone(), o["reward"].clone()]
        outs.append(rec)
        env.close()
    for x, y in zip(*outs):
        assert torch.equal(x, y)


@pytest.mark.parametrize("N,O,B,lanes,T,mapping", [
    (7, 3, 300, 1, 6, "group"), (16, 3, 70, 1, 6, "group"), (7, 3, 1024, 2, 6, "group"), (5, 0, 33, 1, 6, "group"), (7, 3, 64, 1, 1, "group"),
    (3, 3, 300, 1, 6, "aw"), (3, 3, 1100, 2, 6, "aw"), (4, 2, 70, 1, 1, "aw"), (3, 0, 33, 1, 5, "aw")])
def test_next_episode_prefetch_is_bitwise_invisible(N, O, B, lanes, T, mapping, monkeypatch):
    """Both mappings take the placement + assignment of every env's next episode from a block produced ahead of time on a side
    stream (prefetch_kernel); the terminal step copies its entry (and, after the masked reset below, keeps doing so entry by
    entry: a tag that matches the env's episode key is all a kernel needs).  Same Philox stream, same bits: a run with the
    prefetch disabled (FM_PREFETCH=0: every reset is computed inside the step kernel) must be identical, through
    single steps, through fm_step_many with env-range lanes, and across a masked reset that breaks the lockstep."""
    import torch
    cfg = NavConfig(num_agents=N, num_obstacles=O, episode_length=T)    # T = 1: every step is terminal
    monkeypatch.setenv("FM_LANES", str(lanes))
    runs = []
    for pf in ("2", "0"):                                     # 2: also for the agent-warp mapping (off by default there: slower)
        monkeypatch.setenv("FM_PREFETCH", pf)
        env = _env(cfg, B, seed=31, sim=dict(mapping=mapping), num_slots=8)
        g = torch.Generator(device="cuda").manual_seed(5)
        rec = []

        def keep(o):
            rec.extend([o["obs"].clone(), o["node_obs"].clone(), o["adj_env"].clone()])
            if "reward" in o:
                rec.extend([o["reward"].clone(), o["done"].clone()])

        keep(env.reset_tensor())
        for t in range(14):                                   # two auto-resets through single steps
            keep(env.step_tensor(torch.randint(0, 5, (B, N), generator=g, device="cuda", dtype=torch.int32)))
        acts = torch.randint(0, 5, (7, B, N), generator=g, device="cuda", dtype=torch.int32)
        for slot in env.rollout_tensor(acts):                 # one more through fm_step_many (lanes)
            keep(env.slot_outputs(slot))
        mask = (torch.arange(B, device="cuda") % 3 == 0).to(torch.uint8)
        keep(env.reset_tensor(mask=mask))                     # lockstep broken: inline resets from here on
        for t in range(9):
            keep(env.step_tensor(torch.randint(0, 5, (B, N), generator=g, device="cuda", dtype=torch.int32)))
        keep(env.reset_tensor())                              # lockstep again
        for t in range(8):
            keep(env.step_tensor(torch.randint(0, 5, (B, N), generator=g, device="cuda", dtype=torch.int32)))
        # rollouts under stream capture: the prefetch is launched at the start of every fm_step_many call and joined at
        # its end; two replays of (one call that ends at an episode boundary, one that starts behind it)
        acts2 = torch.randint(0, 5, (2 * T + 3, B, N), generator=g, device="cuda", dtype=torch.int32)
        env.reset_tensor()
        env.rollout_tensor(acts2[:2])                         # eager: streams, plans
        env.reset_tensor()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        slot0 = env._slot
        with torch.cuda.graph(graph):
            s1 = env.rollout_tensor(acts2[:T])
            s2 = env.rollout_tensor(acts2[T:])
        for rep in range(2):
            env._slot = slot0
            graph.replay()
            for slot in (s1 + s2)[-5:]:
                keep(env.slot_outputs(slot))
        st = env.get_state()
        rec.extend([st["goal_match"], st["landmark_pos"], st["obstacle_pos"], st["episode"], st["min_time"]])
        runs.append(rec)
        env.close()
    assert len(runs[0]) == len(runs[1])
    for k, (x, y) in enumerate(zip(*runs)):
        assert torch.equal(x, y), k


@pytest.mark.parametrize("form", ["stream", "three", "fused"])
def test_edge_list_corner_cases(form, monkeypatch):
    """All forms of fm_edge_list (FM_EDGE_FORM: the default streamed count / offsets / persistent emission of fm_edges.cu,
    count / scan / emit, and the single-pass look-back kernel;
    E = 35 x 8 lists needs 78 KB of shared memory, above 96 KB the call falls back).  process_adj on shapes the simulator never produces by itself: no edges at all, every edge, a single graph (2-D
    input), E = 35, more graph copies than lanes (repeat = 40), a graph count that is not a multiple of the 8 graphs
    a CTA handles, and 70 000 graphs (more CTAs than one scan segment)."""
    import torch
    import fair_marl_b200 as fm
    monkeypatch.delenv("FM_EDGE_FUSED", raising=False)
    monkeypatch.setenv("FM_EDGE_FORM", form)
    rng = np.random.default_rng(3)
    dev = torch.device("cuda")

    def check(adj, thr, **kw):
        repeat = kw.get("repeat", 1)
        ei, ea = fm.process_adj(torch.as_tensor(adj, device=dev), thr, **kw)
        a3 = adj if adj.ndim == 3 else adj[None]
        rei, rea = oracle_process_adj(np.repeat(a3, repeat, axis=0), thr, inclusive=kw.get("inclusive", False))
        if adj.ndim == 2:
            assert tuple(ei.shape) == (2, rei.shape[1])
        assert (ei.cpu().numpy() == rei).all() and (ea.cpu().numpy() == rea).all()
        return rei.shape[1]

    assert check(np.zeros((5, 9, 9), np.float32), 1.0) == 0                         # nothing within reach
    far = np.full((3, 9, 9), 7.0, np.float32)
    assert check(far, 1.0) == 0
    full = rng.random((11, 9, 9)).astype(np.float32) * 0.9 + 0.01                  # every off- AND on-diagonal entry is an edge
    assert check(full, 1.0) == 11 * 81
    assert check(full[0], 1.0) == 81                                               # 2-D input
    big = rng.random((13, 35, 35)).astype(np.float32) * 2
    check(big, 1.0)
    check(big, 1.0, repeat=40)
    check(big, 1.0, inclusive=True)
    many = (rng.random((70001, 5, 5)) * 2).astype(np.float32)
    check(many, 1.0)
    check(many[:9], 1.0, repeat=3)
    check(many[:2049], 1.0)                                                          # one graph past a 2 048-graph offsets tile
    wide = (rng.random((3000, 17, 17)) * 2).astype(np.float32)                     # more graphs than one wave of emission warps take
    wide[5] = 9.0                                                                  # a graph without edges between graphs with edges
    check(wide, 1.0)
    check(wide[:300], 1.0, repeat=7)
    if form == "stream":                                                           # more than 256 offsets tiles: every thread of the
        check((rng.random((600001, 3, 3)) * 2).astype(np.float32), 1.0)            # look-back sums several tile totals
    # graph_offsets of every copy, and a caller-side capacity smaller than the list (the ABI truncates, nnz stays the full count)
    from fair_marl_b200 import _lib
    lib = _lib.load()
    for a_np, repeat in ((wide[:300], 3), (big, 2)):
        a = torch.as_tensor(a_np, device=dev)
        G, E = a.shape[0], a.shape[1]
        ei, ea, off = fm.process_adj(a, 1.0, repeat=repeat, return_offsets=True)
        per = ((a_np < 1.0) & (a_np > 0)).reshape(G, -1).sum(1)
        want = np.concatenate([[0], np.cumsum(np.repeat(per, repeat))])
        assert (off.cpu().numpy() == want).all()
        n = int(want[-1])
        cap = n // 2 + 1
        off2 = torch.empty(G * repeat + 1, dtype=torch.int64, device=dev)
        ei2 = torch.full((2, cap), -1, dtype=torch.int64, device=dev)
        ea2 = torch.full((cap,), -1.0, dtype=torch.float32, device=dev)
        nnz = torch.zeros(1, dtype=torch.int64, device=dev)
        _lib.check(lib.fm_edge_list(0, a.data_ptr(), G, E, 1.0, 0, repeat, cap, off2.data_ptr(), ei2.data_ptr(), ea2.data_ptr(),
                                    nnz.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "fm_edge_list")
        assert int(nnz.item()) == n and torch.equal(off2, off)
        assert torch.equal(ei2, ei[:, :cap]) and torch.equal(ea2, ea[:cap, 0])


@pytest.mark.parametrize("N,O,B", [(7, 3, 21), (16, 3, 9), (3, 3, 40), (5, 0, 13)])
def test_outputs_at_any_alignment_and_partial_outputs(N, O, B):
    """The emission goes through TMA bulk stores, which need 16-byte aligned global addresses: the group kernel builds
    its images at the 16-byte phase of their destination and stores the <= 3 head / tail words by lanes, the
    agent-warp kernel falls back to vector / scalar stores.  Output arrays that start 4, 8 or 12 bytes past a
    16-byte boundary, and FmOutputs with null members (the ABI allows any subset), must give the same bytes."""
    import torch
    cfg = NavConfig(num_agents=N, num_obstacles=O, episode_length=5)
    E = 2 * N + O
    shapes = {"obs": (B, N, 7), "node_obs": (B, N, E, 11), "adj": (B, E, E), "reward": (B, N)}
    g = torch.Generator(device="cuda").manual_seed(9)
    acts = torch.randint(0, 5, (8, B, N), generator=g, device="cuda", dtype=torch.int32)

    def run(offset, keep):
        env = _env(cfg, B, seed=17)
        env.reset_tensor()
        rec = []
        for t in range(8):
            out = {}
            for name, shp in shapes.items():
                n = int(np.prod(shp))
                flat = torch.full((n + 8,), float("nan"), device="cuda")
                out[name] = flat[offset:offset + n].view(shp)
            out["done"] = torch.zeros((B, N), dtype=torch.uint8, device="cuda")
            if keep is None:
                env.step_tensor(acts[t], out=out)
            else:                                   # raw ABI call with null members
                import ctypes as C
                from fair_marl_b200 import _lib
                o = _lib.FmOutputs()
                for name in keep:
                    setattr(o, name, out[name].data_ptr())
                _lib.check(env.lib.fm_step(env._h, acts[t].data_ptr(), C.byref(o), env._stream()), "fm_step")
            rec.append({k: v.clone() for k, v in out.items()})
        env.close()
        return rec

    ref = run(0, None)
    for offset in (1, 2, 3):
        got = run(offset, None)
        for a, b in zip(ref, got):
            for k in a:
                assert torch.equal(a[k], b[k]), (offset, k)
    for keep in (("reward", "done"), ("node_obs",), ("adj", "obs"), ("obs", "reward")):
        got = run(2, keep)
        for a, b in zip(ref, got):
            for k in keep:
                assert torch.equal(a[k], b[k]), (keep, k)


@pytest.mark.parametrize("N,O,W,B,prefetch", [(3, 3, 2, 192, "1"), (3, 3, 1, 70, "1"), (4, 2, 1, 100, "0"), (7, 3, 2, 64, "1"),
                                              (16, 3, 2, 24, "1"), (12, 0, 1, 20, "0"),       # step_kernel<16, true>
                                              (20, 2, 2, 8, "1")])                            # step_kernel<32, true>
def test_walls_reset_and_rollout_match_oracle(N, O, W, B, prefetch, monkeypatch):
    """num_walls > 0 (group-per-env kernels; agent-warp kernels at N = 3, O = 3): the reset draws wall axis / orientation from the same Philox stream and
    rejects placements inside the wall boxes exactly like the oracle (bit-exact state), and a random-action rollout
    across auto-resets stays within tolerance step by step -- with the next-episode prefetch on and off."""
    monkeypatch.setenv("FM_PREFETCH", prefetch)
    cfg = NavConfig(num_agents=N, num_obstacles=O, num_walls=W, goal_rew=30.0, collision_rew=30.0, episode_length=9)
    env = _env(cfg, B, seed=11, env_offset=5)
    import parity_util
    want = "aw" if (N, O, W) in ((3, 3, 1), (3, 3, 2), (4, 2, 1)) and parity_util.MAPPING == "auto" else "group"
    assert env.mapping == want and env.num_entities == 2 * N + O + W
    orc = NavGraphOracle(cfg, B, seed=11, env_offset=5)
    out = _np(env.reset_tensor())
    ref = orc.reset()
    st, rs = device_state_to_nav(env.get_state()), orc.get_state()
    assert (st.wall_len == rs.wall_len.astype(np.float32)).all() and (st.wall_axis == rs.wall_axis).all()
    assert (st.wall_orient == rs.wall_orient).all() and set(np.unique(st.wall_orient)) == {0, 1}
    assert (st.pos == rs.pos).all() and (st.landmark_pos == rs.landmark_pos).all() and (st.goal_match == rs.goal_match).all()
    assert not orc._in_wall_box(st.pos.reshape(-1, 2), 0.05, sel=np.repeat(np.arange(B), N)).any()
    assert_close(out["node_obs"], ref["node_obs"], "reset node_obs")
    assert (out["adj_env"] == ref["adj"].astype(np.float32)).all()
    rng = np.random.default_rng(1)
    for t in range(21):
        orc.set_state(device_state_to_nav(env.get_state()))
        a = rng.integers(0, 5, (B, N))
        out = _np(env.step_tensor(_actions(a)))
        ref = orc.step(actions=a, autoreset=True)
        out["adj"] = out["adj_env"]
        compare_step_outputs(out, ref, cfg)
        post, rpost = device_state_to_nav(env.get_state()), orc.get_state()
        if ref["reset"].any():
            assert (post.pos == rpost.pos).all() and (post.goal_match == rpost.goal_match).all()
            assert (post.wall_axis == rpost.wall_axis).all() and (post.wall_orient == rpost.wall_orient).all()
        else:
            assert_close(post.pos, rpost.pos, "pos")
            assert (post.num_obstacle_collisions == rpost.num_obstacle_collisions).all()
    env.close()
