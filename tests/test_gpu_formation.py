"""Formation-family kernels (csrc/fm_formation.cu through fm_formation_*) against the float64 oracle
(oracle/formation.py, itself pinned to the unmodified reference by tests/test_oracle_formation.py)."""
from dataclasses import fields

import numpy as np
import pytest

from oracle.formation import FormationConfig, FormationOracle, FormationState
from oracle.make_formation_golden import load, set_walls, state_from
from oracle.navgraph import INFO_KEYS
from parity_util import assert_close, assert_fairness_close

pytestmark = pytest.mark.gpu

INT_FIELDS = ("goal_match", "step", "episode")
RATIO_KEYS = ("Mean_by_variance", "Time_mean_by_stddev")


def _sim(cfg: FormationConfig, **kw):
    from fair_marl_b200 import FormationSimConfig
    return FormationSimConfig(
        num_agents=cfg.num_agents, num_obstacles=cfg.num_obstacles, world_size=cfg.world_size, max_speed=cfg.max_speed,
        collision_rew=cfg.collision_rew, goal_rew=cfg.goal_rew, min_dist_thresh=cfg.min_dist_thresh,
        min_obs_dist=cfg.min_obs_dist, episode_length=cfg.episode_length, fair_rew=cfg.fair_rew, zeroshift=cfg.zeroshift,
        collaborative=cfg.collaborative, fairness_reward=cfg.fairness_reward, assignment=cfg.assignment,
        num_walls=cfg.num_walls, **kw)


def _fp32(st: FormationState) -> FormationState:
    d = {}
    for f in fields(FormationState):
        a = np.asarray(getattr(st, f.name))
        if f.name in INT_FIELDS:
            d[f.name] = a.astype(np.int64)
        elif f.name == "status":
            d[f.name] = a.astype(bool)
        else:
            with np.errstate(over="ignore"):
                d[f.name] = a.astype(np.float32).astype(np.float64)
    return FormationState(**d)


def _to_device(st: FormationState):
    return {f.name: (np.asarray(getattr(st, f.name)).astype(np.uint8) if f.name == "status" else np.asarray(getattr(st, f.name)))
            for f in fields(FormationState)}


def _from_device(dev) -> FormationState:
    d = {}
    for f in fields(FormationState):
        a = dev[f.name].cpu().numpy()
        d[f.name] = a.astype(np.int64) if f.name in INT_FIELDS else (a.astype(bool) if f.name == "status" else a.astype(np.float64))
    return FormationState(**d)


def _walls_to_device(orc: FormationOracle):
    """Wall geometry of the oracle (kept beside FormationState) as device state entries, rounded to fp32 like the rest."""
    if not orc.cfg.num_walls:
        return {}
    return {"wall_axis": orc.wall_axis.astype(np.float32), "wall_orient": orc.wall_orient.astype(np.int32),
            "wall_len": orc.wall_len.astype(np.float32)}


def _walls_from_device(orc: FormationOracle, dev) -> None:
    if orc.cfg.num_walls:
        orc.wall_axis = dev["wall_axis"].cpu().numpy().astype(np.float64)
        orc.wall_orient = dev["wall_orient"].cpu().numpy().astype(np.int64)
        orc.wall_len = dev["wall_len"].cpu().numpy().astype(np.float64)


def _np(out):
    return {k: v.cpu().numpy() for k, v in out.items()}


def _actions(a):
    import torch
    return torch.as_tensor(np.asarray(a), dtype=torch.int32, device="cuda")


def _compare_step(out, ref, post: FormationState, rpost: FormationState):
    assert_close(out["obs"], ref["obs"], "obs")
    assert_close(out["node_obs"], ref["node_obs"], "node_obs")
    assert_close(out["adj_env"], ref["adj"], "adj")
    assert_close(out["reward"], ref["reward"], "reward")
    assert (out["done"].astype(bool) == ref["done"]).all(), "done"
    for k, key in enumerate(INFO_KEYS):
        (assert_fairness_close if key in RATIO_KEYS else assert_close)(out["info"][..., k], ref["info"][key], key)
    for f in ("goal_match", "step", "status"):
        assert (getattr(post, f) == getattr(rpost, f)).all(), f


@pytest.mark.parametrize("name", ["formation_n3_o3_fafr", "formation_n4_o2_fa", "formation_n7_o3_fafr", "formation_n3_o3_oa",
                                  "formation_n3_o3_ra", "formation_n3_o2_w2", "formation_n3_o3_w1_oa"])
def test_step_matches_oracle_on_reference_states(name):
    """One step from every recorded reference state (rounded to fp32): device vs float64 oracle, outputs, info and the
    whole post-step state -- status latches, occupancy table, goal history and nearest-landmark latches included."""
    import fair_marl_b200 as fm
    cfg, g = load(name)
    pre = _fp32(state_from(g, "pre_"))
    T = pre.pos.shape[0]
    env = fm.B200FormationVecEnv(_sim(cfg, auto_reset=False), num_envs=T)
    orc = FormationOracle(cfg, T)
    orc.set_state(pre)
    set_walls(orc, g, "pre_")
    dev_walls = _walls_to_device(orc)
    env.set_state({**_to_device(pre), **dev_walls})
    _walls_from_device(orc, env.get_state())                  # the fp32 geometry the device holds
    out = _np(env.step_tensor(_actions(g["actions"])))
    ref = orc.step(g["actions"], autoreset=False)
    post, rpost = _from_device(env.get_state()), orc.get_state()
    _compare_step(out, ref, post, rpost)
    for f in ("pos", "vel", "p_dist", "dists_to_goal", "times_required", "dist_left_to_goal", "num_agent_collisions",
              "num_obstacle_collisions", "dist_traveled_mean", "dist_traveled_stddev", "goal_reached", "occupied", "goal_history"):
        assert_close(getattr(post, f), getattr(rpost, f), f)
    # smooth quantities straight against the reference's own float64 outputs (inputs were rounded to fp32 on the way in)
    assert_close(out["adj_env"], g["out_adj"], "adj vs reference")
    assert_close(out["obs"][..., :4], g["out_obs"][..., :4], "vel / pos vs reference")
    assert orc.branch_hits.get("status_latched", 0) > 0
    if cfg.assignment == "fair":
        assert orc.branch_hits.get("subset_index_quirk", 0) > 0
    if cfg.num_walls:
        assert orc.branch_hits.get("wall_end_cap", 0) > 0 and orc.branch_hits.get("wall_box_hit", 0) > 0
        assert (out["node_obs"][:, :, -cfg.num_walls:, 12] == 3.0).all()
    env.close()


@pytest.mark.parametrize("N,O,B,collab,fair,assignment,W", [
    (3, 3, 48, False, True, "fair", 0), (4, 2, 32, True, False, "fair", 0), (2, 1, 32, False, True, "fair", 0),
    (3, 3, 70, False, True, "fair", 0),                                 # a ragged last warp: the lane-store emission path
    (5, 2, 32, False, True, "fair", 0), (7, 3, 40, False, True, "fair", 0),   # N > 4: serial lexifair descent every step
    (3, 3, 48, False, False, "optimal", 0), (4, 1, 32, True, False, "optimal", 0), (3, 3, 48, False, False, "random", 0),
    (6, 2, 32, False, False, "random", 0),
    (3, 2, 40, False, True, "fair", 2), (4, 3, 32, False, False, "optimal", 1), (5, 1, 32, False, True, "fair", 2)])   # walls
def test_reset_and_rollout_match_oracle(N, O, B, collab, fair, assignment, W):
    """Device reset == oracle reset bit for bit (same Philox draws, same acceptance rules, same lexifair), then a rollout
    that steers at the goals (so agents latch and envs finish early) across auto-resets, compared step by step."""
    import fair_marl_b200 as fm
    cfg = FormationConfig(num_agents=N, num_obstacles=O, goal_rew=30.0, collision_rew=30.0, episode_length=15,
                          collaborative=collab, fairness_reward=fair, assignment=assignment, num_walls=W)
    env = fm.B200FormationVecEnv(_sim(cfg), num_envs=B, seed=7, env_offset=3)
    orc = FormationOracle(cfg, B, seed=7, env_offset=3)
    out, ref = _np(env.reset_tensor()), orc.reset()
    st, rs = _from_device(env.get_state()), orc.get_state()
    if W:                                         # same draws, same fp32 arithmetic: the wall geometry is bit-identical
        dev = env.get_state()
        assert (dev["wall_axis"].cpu().numpy() == orc.wall_axis.astype(np.float32)).all()
        assert (dev["wall_orient"].cpu().numpy() == orc.wall_orient).all()
        assert (dev["wall_len"].cpu().numpy() == orc.wall_len.astype(np.float32)).all()
    for f in ("pos", "landmark_pos", "obstacle_pos"):
        assert (getattr(st, f) == getattr(rs, f)).all(), f
    assert (st.goal_match == rs.goal_match).all() and (st.episode == 1).all() and not st.status.any()
    assert_close(st.min_time, rs.min_time, "min_time")
    assert_close(st.occupied, rs.occupied, "occupancy after the reset observation")
    assert_close(out["obs"], ref["obs"], "reset obs")
    assert_close(out["node_obs"], ref["node_obs"], "reset node_obs")
    assert_close(out["adj_env"], ref["adj"], "reset adj")
    rng = np.random.default_rng(5)
    resets = early = 0
    for t in range(32):
        dev = env.get_state()
        cur = _from_device(dev)
        orc.set_state(cur)
        _walls_from_device(orc, dev)
        d = np.take_along_axis(cur.landmark_pos, cur.goal_match[..., None], axis=1) - cur.pos
        seek = np.where(np.abs(d[..., 0]) > np.abs(d[..., 1]), np.where(d[..., 0] > 0, 1, 2), np.where(d[..., 1] > 0, 3, 4))
        a = np.where(rng.random((B, N)) < 0.25, rng.integers(0, 5, (B, N)), seek)
        out = _np(env.step_tensor(_actions(a)))
        ref = orc.step(a, autoreset=True)
        post, rpost = _from_device(env.get_state()), orc.get_state()
        _compare_step(out, ref, post, rpost)
        hit = ref["reset"]
        resets += int(hit.sum())
        early += int((hit & (cur.step + 1 < cfg.episode_length)).sum())
        assert (post.episode == rpost.episode).all()
        if hit.any():
            assert (post.pos[hit] == rpost.pos[hit]).all() and (post.landmark_pos[hit] == rpost.landmark_pos[hit]).all()
            assert (post.obstacle_pos[hit] == rpost.obstacle_pos[hit]).all()
        assert_close(post.pos, rpost.pos, "pos")
        assert_close(post.occupied, rpost.occupied, "occupied")
        assert_close(post.goal_history, rpost.goal_history, "goal_history")
    assert resets >= B                            # time-outs ...
    if N <= 4:
        assert early > 0                          # ... and all-agents-latched early resets both happened
    env.close()


@pytest.mark.parametrize("N,O,B,T,slots", [(3, 3, 4096 + 70, 40, 40), (4, 2, 8192, 12, 12), (3, 3, 300, 30, 4), (5, 2, 4500, 9, 9)])
@pytest.mark.parametrize("lanes", ["1", "2"])
def test_rollout_lanes_equal_single_steps(N, O, B, T, slots, lanes, monkeypatch):
    """fm_formation_step_many (one lane, or two env-range lanes on two streams for B >= 4096; pre-generated actions) against T calls of
    fm_formation_step: every output of every step still in the ring, the final state and the info rows, bit for bit.
    Short episodes, so that time-out resets fall inside the rollout; goal seeking, so that early resets do too."""
    import torch
    import fair_marl_b200 as fm
    monkeypatch.setenv("FM_FORM_LANES", lanes)
    cfg = FormationConfig(num_agents=N, num_obstacles=O, goal_rew=30.0, collision_rew=30.0, episode_length=7)
    e_r = fm.B200FormationVecEnv(_sim(cfg, info_every_step=False), num_envs=B, seed=11, num_slots=slots)
    e_s = fm.B200FormationVecEnv(_sim(cfg, info_every_step=False), num_envs=B, seed=11, num_slots=slots)
    e_r.reset_tensor(); e_s.reset_tensor()
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = torch.randint(0, 5, (T, B, N), generator=g, device="cuda", dtype=torch.int32)
    slot_of = e_r.rollout_tensor(acts)
    for t in range(T):
        o_s = e_s.step_tensor(acts[t])
        if slot_of[t] not in slot_of[t + 1:]:
            o_r = e_r.slot_outputs(slot_of[t])
            for k in ("obs", "node_obs", "adj_env", "reward", "done"):
                assert torch.equal(o_r[k], o_s[k]), (t, k)
    s_r, s_s = e_r.get_state(), e_s.get_state()
    for k in s_r:
        assert torch.equal(s_r[k], s_s[k]), k
    assert torch.equal(e_r._slots[0]["info"], e_s._slots[0]["info"])
    assert int(s_r["episode"].max()) >= 1 + T // 7
    # a second rollout continues the same trajectory
    slot_of = e_r.rollout_tensor(acts[:3])
    for t in range(3):
        o_s = e_s.step_tensor(acts[t])
        assert torch.equal(e_r.slot_outputs(slot_of[t])["node_obs"], o_s["node_obs"]), t
    e_r.close(); e_s.close()


def test_masked_reset_and_errors():
    import torch
    import fair_marl_b200 as fm
    cfg = FormationConfig(num_agents=3, num_obstacles=2)
    env = fm.B200FormationVecEnv(_sim(cfg), num_envs=40, seed=1)
    env.reset_tensor()
    before = _from_device(env.get_state())
    mask = np.zeros(40, np.uint8); mask[::3] = 1
    env.reset_tensor(mask)
    after = _from_device(env.get_state())
    keep = mask == 0
    assert (after.pos[keep] == before.pos[keep]).all() and (after.episode[keep] == 1).all()
    assert (after.episode[~keep] == 2).all() and (after.pos[~keep] != before.pos[~keep]).any()
    with pytest.raises(ValueError):
        env.step_tensor(torch.zeros((40, 3), dtype=torch.int64, device="cuda"))
    with pytest.raises(fm._lib.FairMarlError):
        fm.B200FormationVecEnv(fm.FormationSimConfig(num_agents=8), num_envs=8)
    with pytest.raises(fm._lib.FairMarlError):                          # the base scenarios have no fairness term
        fm.B200FormationVecEnv(fm.FormationSimConfig(assignment="optimal", fairness_reward=True), num_envs=8)
    env.close()


def test_info_rows_only_when_every_agent_is_done():
    """info_every_step = False (rollouts): the info rows are written on the steps on which every agent of the env is done
    -- the values the runner reads -- and left alone otherwise; everything else is unchanged."""
    import fair_marl_b200 as fm
    cfg = FormationConfig(num_agents=3, num_obstacles=3, episode_length=6)
    e_all = fm.B200FormationVecEnv(_sim(cfg), num_envs=64, seed=3)
    e_end = fm.B200FormationVecEnv(_sim(cfg, info_every_step=False), num_envs=64, seed=3)
    e_all.reset_tensor(); e_end.reset_tensor()
    rng = np.random.default_rng(0)
    for t in range(13):
        a = _actions(rng.integers(0, 5, (64, 3)))
        o_all, o_end = _np(e_all.step_tensor(a)), _np(e_end.step_tensor(a))
        for k in ("obs", "node_obs", "adj_env", "reward", "done"):
            assert (o_all[k] == o_end[k]).all(), (t, k)
        fin = o_all["done"].all(axis=1)
        assert (o_all["info"][fin] == o_end["info"][fin]).all()
        if t < 5:
            assert (o_end["info"] == 0).all()                           # untouched before the first terminal step
    e_all.close(); e_end.close()


def test_partial_outputs_do_not_leave_ready_flags_behind():
    """A step without node_obs / adj runs the logic kernel alone (no image kernel to consume the per-16-env ready flags of the
    dependent launch); the steps after it must still build their rows from THEIR recipes."""
    import ctypes as C
    import torch
    import fair_marl_b200 as fm
    from fair_marl_b200 import _lib
    cfg = FormationConfig(num_agents=3, num_obstacles=3, episode_length=9)
    B = 4096
    e_a = fm.B200FormationVecEnv(_sim(cfg), num_envs=B, seed=5)
    e_b = fm.B200FormationVecEnv(_sim(cfg), num_envs=B, seed=5)
    e_a.reset_tensor(); e_b.reset_tensor()
    g = torch.Generator(device="cuda").manual_seed(1)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for t in range(12):
        a = torch.randint(0, 5, (B, 3), generator=g, device="cuda", dtype=torch.int32)
        o_b = e_b.step_tensor(a)
        if t % 3 == 1:                                  # logic kernel alone
            s = e_a._slots[0]
            o = _lib.FmOutputs(s["obs"].data_ptr(), None, None, s["reward"].data_ptr(), s["done"].data_ptr(), s["info"].data_ptr())
            _lib.check(e_a.lib.fm_formation_step(e_a._h, a.data_ptr(), C.byref(o), stream), "fm_formation_step")
            assert torch.equal(s["obs"], o_b["obs"]), t
        else:
            o_a = e_a.step_tensor(a)
            for k in ("obs", "node_obs", "adj_env", "reward"):
                assert torch.equal(o_a[k], o_b[k]), (t, k)
    e_a.close(); e_b.close()
