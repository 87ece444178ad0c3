"""The DEVICE FUNCTIONS of csrc/fm_form.cuh compiled for the host (g++, ASan + UBSan, -ffp-contract=off) and run on
the formation fixtures: a sanitiser pass (out-of-bounds / uninitialised-index bugs in the per-thread local arrays show up
here, not as layout-dependent wrong answers on the GPU) plus a CPU-side parity check of the kernel SOURCE against the
oracle.  The translation unit is assembled from the real csrc files (tests/host_emul/prelude.h stands in for the CUDA
built-ins).  This is test infrastructure: not a CPU path of the product."""
import os
import subprocess
import sys
from dataclasses import fields

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
CSRC = os.environ.get("FM_HOST_EMUL_CSRC", os.path.join(ROOT, "fair-marl_b200", "csrc"))   # override: bisecting an older / newer source tree
EMUL = os.path.join(ROOT, "tests", "host_emul")

from oracle.formation import FormationConfig, FormationOracle, FormationState  # noqa: E402
from oracle.make_formation_golden import load, state_from  # noqa: E402
from oracle.navgraph import INFO_KEYS  # noqa: E402
from parity_util import assert_close, assert_fairness_close  # noqa: E402

STATE_ORDER = ("pos", "vel", "p_dist", "landmark_pos", "obstacle_pos", "goal_match", "dists_to_goal", "times_required",
               "dist_left_to_goal", "num_agent_collisions", "num_obstacle_collisions", "dist_traveled_mean",
               "dist_traveled_stddev", "step", "min_time", "episode", "status", "goal_reached", "occupied", "goal_history")
INT_FIELDS = ("goal_match", "step", "episode")


def _between(text, start, end):
    a = text.index(start)
    return text[a:text.index(end, a)]


@pytest.fixture(scope="module")
def host_exe(tmp_path_factory):
    dev = open(os.path.join(CSRC, "fm_device.cuh")).read()
    launch = open(os.path.join(CSRC, "fm_launch.h")).read()
    small = open(os.path.join(CSRC, "fm_small.cuh")).read()
    form = open(os.path.join(CSRC, "fm_form.cuh")).read()
    philox = _between(dev, "__device__ __forceinline__ void philox4x32_10(", "// U(-ws/2, ws/2)^2 draw number")
    log1p_unit = _between(dev, "__device__ __forceinline__ float log1p_unit(float t) {", "// One contact-force term, core.py:389-392")
    params = _between(launch, "struct FormParams {", "cudaError_t launch_formation")
    small_body = _between(small, "// k-th permutation of 0..N-1", "}  // namespace fm")
    form_body = _between(form, "constexpr int F_OBS", "// ---- device only from here")
    assert "u01_24" in philox and "lexifair_small" in small_body and "form_step_env" in form_body and "fmaf" in log1p_unit
    src = "\n".join(['#include "prelude.h"', "namespace fm {", philox, log1p_unit, params, small_body, form_body, "}  // namespace fm",
                     open(os.path.join(EMUL, "harness.inc")).read()])
    d = tmp_path_factory.mktemp("host_emul")
    cpp, exe = d / "formation_host.cpp", d / "formation_host"
    cpp.write_text(src)
    # -ftrivial-auto-var-init=pattern: a read of an uninitialised local gives a wrong answer here instead of a lucky one
    cmd = ["g++", "-std=c++17", os.environ.get("FM_HOST_EMUL_OPT", "-O1"), "-g", *os.environ.get("FM_HOST_EMUL_FP", "-ffp-contract=off").split(), "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-ftrivial-auto-var-init=pattern",
           "-I", EMUL, "-I", os.path.join(ROOT, "include"), str(cpp), "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
    return str(exe)


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


def _dtype(name):
    return np.int32 if name in INT_FIELDS else (np.uint8 if name == "status" else np.float32)


def _run(exe, tmp, cfg: FormationConfig, st: FormationState, actions=None, is_reset=False, mask=None, seed=0, env_offset=0,
         auto_reset=False):
    B, N, O = st.pos.shape[0], cfg.num_agents, cfg.num_obstacles
    E = 2 * N + O
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    with open(fin, "wb") as f:
        np.array([B, N, O, cfg.episode_length, int(cfg.fairness_reward), int(cfg.collaborative), int(auto_reset), int(is_reset),
                  int(mask is not None), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, env_offset,
                  {"fair": 0, "optimal": 1, "random": 2}[cfg.assignment]], dtype=np.int32).tofile(f)
        np.array([cfg.world_size, cfg.max_speed if cfg.max_speed is not None else -1.0, cfg.collision_rew, cfg.goal_rew,
                  cfg.min_dist_thresh, cfg.min_obs_dist, cfg.fair_rew, cfg.zeroshift], dtype=np.float64).tofile(f)
        for name in STATE_ORDER:
            with np.errstate(over="ignore"):
                np.ascontiguousarray(np.asarray(getattr(st, name)).astype(_dtype(name))).tofile(f)
        np.ascontiguousarray(np.zeros((B, N), np.int32) if actions is None else np.asarray(actions, dtype=np.int32)).tofile(f)
        np.ascontiguousarray(np.ones(B, np.uint8) if mask is None else np.asarray(mask, dtype=np.uint8)).tofile(f)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1")
    res = subprocess.run([exe, fin, fout], capture_output=True, text=True, env=env)
    assert res.returncode == 0, f"host run of the kernel source failed (rc {res.returncode}):\n{res.stderr[-6000:]}"
    raw = np.fromfile(fout, dtype=np.uint8)
    off, d = 0, {}
    for name in STATE_ORDER:
        shape = np.asarray(getattr(st, name)).shape
        dt = np.dtype(_dtype(name))
        n = int(np.prod(shape)) * dt.itemsize
        a = raw[off:off + n].view(dt).reshape(shape)
        off += n
        d[name] = a.astype(np.int64) if name in INT_FIELDS else (a.astype(bool) if name == "status" else a.astype(np.float64))
    out = {}
    for name, shape, dt in (("obs", (B, N, 11), np.float32), ("node_obs", (B, N, E, 13), np.float32), ("adj", (B, E, E), np.float32),
                            ("reward", (B, N), np.float32), ("done", (B, N), np.uint8), ("info", (B, N, 14), np.float32)):
        n = int(np.prod(shape)) * np.dtype(dt).itemsize
        out[name] = raw[off:off + n].view(dt).reshape(shape)
        off += n
    assert off == raw.size
    return out, FormationState(**d)


def _compare(out, ref, post, rpost):
    for k in ("obs", "node_obs", "adj", "reward"):
        assert_close(out[k], ref[k], k)
    assert (out["done"].astype(bool) == ref["done"]).all()
    for k, key in enumerate(INFO_KEYS):
        (assert_fairness_close if key in ("Mean_by_variance", "Time_mean_by_stddev") else assert_close)(out["info"][..., k], ref["info"][key], key)
    for f in ("goal_match", "step", "status"):
        assert (getattr(post, f) == getattr(rpost, f)).all(), f
    for f in ("pos", "vel", "p_dist", "dists_to_goal", "times_required", "dist_left_to_goal", "num_agent_collisions",
              "num_obstacle_collisions", "dist_traveled_mean", "dist_traveled_stddev", "goal_reached", "occupied", "goal_history"):
        assert_close(getattr(post, f), getattr(rpost, f), f)


@pytest.mark.parametrize("name", ["formation_n3_o3_fafr", "formation_n4_o2_fa", "formation_n7_o3_fafr", "formation_n3_o3_oa",
                                  "formation_n3_o3_ra"])
def test_step_source_is_sanitizer_clean_and_matches_oracle(host_exe, tmp_path, name):
    cfg, g = load(name)
    pre = _fp32(state_from(g, "pre_"))
    out, post = _run(host_exe, str(tmp_path), cfg, pre, actions=g["actions"])
    orc = FormationOracle(cfg, pre.pos.shape[0])
    orc.set_state(pre)
    ref = orc.step(g["actions"], autoreset=False)
    _compare(out, ref, post, orc.get_state())


@pytest.mark.parametrize("N,O,collab,fair,assignment", [
    (4, 2, True, False, "fair"), (3, 3, False, True, "fair"), (2, 1, False, True, "fair"),
    (5, 2, False, True, "fair"), (6, 1, False, False, "fair"), (7, 3, False, True, "fair"),     # serial lexifair descent (N > 4)
    (3, 3, False, False, "optimal"), (4, 1, False, False, "optimal"), (5, 0, True, False, "optimal"),
    (3, 3, False, False, "random"), (6, 2, False, False, "random")])
def test_reset_and_rollout_source_is_sanitizer_clean_and_matches_oracle(host_exe, tmp_path, N, O, collab, fair, assignment):
    B = 24
    cfg = FormationConfig(num_agents=N, num_obstacles=O, goal_rew=30.0, collision_rew=30.0, episode_length=8,
                          collaborative=collab, fairness_reward=fair, assignment=assignment)
    orc = FormationOracle(cfg, B, seed=7, env_offset=3)
    ref = orc.reset()
    zero = FormationOracle(cfg, B).get_state()
    for name in ("dists_to_goal", "times_required", "dist_left_to_goal", "goal_reached", "goal_history"):
        getattr(zero, name)[:] = 0.0                       # what fm_formation_create's memset leaves
    zero.min_time[:] = 0.0
    out, st = _run(host_exe, str(tmp_path), cfg, zero, is_reset=True, seed=7, env_offset=3)
    rs = orc.get_state()
    for f in ("pos", "landmark_pos", "obstacle_pos", "goal_match", "episode"):
        assert (getattr(st, f) == getattr(rs, f)).all(), f
    for k in ("obs", "node_obs", "adj"):
        assert_close(out[k], ref[k], "reset " + k)
    rng = np.random.default_rng(5)
    resets = 0
    for t in range(12):
        orc.set_state(st)
        d = np.take_along_axis(st.landmark_pos, st.goal_match[..., None], axis=1) - st.pos
        seek = np.where(np.abs(d[..., 0]) > np.abs(d[..., 1]), np.where(d[..., 0] > 0, 1, 2), np.where(d[..., 1] > 0, 3, 4))
        a = np.where(rng.random((B, N)) < 0.25, rng.integers(0, 5, (B, N)), seek)
        out, st = _run(host_exe, str(tmp_path), cfg, st, actions=a, seed=7, env_offset=3, auto_reset=True)
        ref = orc.step(a, autoreset=True)
        _compare(out, ref, st, orc.get_state())
        resets += int(ref["reset"].sum())
    assert resets >= B


@pytest.mark.parametrize("N,O", [(3, 3), (4, 0), (2, 2)])
def test_rare_branches_directed(host_exe, tmp_path, N, O):
    """States built to reach what the recorded fixtures do not: every goal marked occupied with all agents far away (the
    table is cleared and the agent 'goes to itself', :951 / :1266), no obstacles at all, the smallest team."""
    B = 16
    cfg = FormationConfig(num_agents=N, num_obstacles=O, goal_rew=30.0, collision_rew=30.0, min_obs_dist=0.3, episode_length=25)
    orc = FormationOracle(cfg, B, seed=11)
    orc.reset()
    st = orc.get_state()
    rng = np.random.default_rng(N * 10 + O)
    st.occupied[: B // 2] = 1.0                                   # half of the envs: every goal taken
    st.goal_history[: B // 2] = rng.integers(0, N, (B // 2, N)).astype(np.float64)
    st.landmark_pos[: B // 4] += 3.0                              # a quarter: all goals far from every agent
    st.occupied[B // 2:] = rng.random((B - B // 2, N))            # the rest: partial occupancies, one exact 1.0 each
    st.occupied[B // 2:, 0] = 1.0
    st = _fp32(st)
    orc.set_state(st)
    a = rng.integers(0, 5, (B, N))
    out, post = _run(host_exe, str(tmp_path), cfg, st, actions=a)
    ref = orc.step(a, autoreset=False)
    _compare(out, ref, post, orc.get_state())
    assert orc.branch_hits.get("all_occupied_cleared", 0) > 0


@pytest.mark.parametrize("N,O,fair,collab,max_speed", [(4, 8, True, False, 2.0), (3, 0, False, True, 2.0), (2, 8, True, True, 2.0),
                                                      (4, 3, False, False, 2.0), (3, 2, True, False, None), (4, 1, True, False, 0.7)])
def test_fuzzed_states(host_exe, tmp_path, N, O, fair, collab, max_speed):
    """Random (not reachable-by-rollout) states: clustered agents and goals so that thresholds, occupancies equal to 1.0,
    latched agents and collisions all occur; the maximum obstacle count (FM_FORMATION_MAX_OBSTACLES) included."""
    B = 192
    rng = np.random.default_rng(100 * N + O)
    cfg = FormationConfig(num_agents=N, num_obstacles=O, goal_rew=30.0, collision_rew=30.0, min_obs_dist=0.4,
                          episode_length=25, fairness_reward=fair, collaborative=collab, max_speed=max_speed)
    st = FormationOracle(cfg, B).get_state()
    st.landmark_pos[:] = rng.uniform(-0.8, 0.8, (B, N, 2))
    near = rng.random((B, N)) < 0.5                               # half of the agents sit next to some goal
    which = rng.integers(0, N, (B, N))
    at_goal = np.take_along_axis(st.landmark_pos, which[..., None], axis=1) + rng.normal(0, 0.04, (B, N, 2))
    st.pos[:] = np.where(near[..., None], at_goal, rng.uniform(-1, 1, (B, N, 2)))
    st.vel[:] = rng.normal(0, 0.5, (B, N, 2))
    st.p_dist[:] = rng.uniform(0, 2, (B, N))
    st.obstacle_pos[:] = rng.uniform(-0.8, 0.8, (B, O, 2))
    st.goal_match[:] = np.argsort(rng.random((B, N)), axis=1)
    st.status[:] = rng.random((B, N)) < 0.3
    st.occupied[:] = np.where(rng.random((B, N)) < 0.4, 1.0, rng.random((B, N)))
    st.goal_history[:] = rng.integers(-1, N, (B, N))
    st.goal_reached[:] = rng.integers(-1, N, (B, N))
    latched = rng.random((B, N)) < 0.4
    st.times_required[:] = np.where(latched, rng.integers(1, 10, (B, N)) * 0.1, -1.0)
    st.dists_to_goal[:] = np.where(rng.random((B, N)) < 0.2, -1.0, rng.uniform(0, 2, (B, N)))
    st.dist_left_to_goal[:] = rng.uniform(0, 1, (B, N))
    st.num_agent_collisions[:] = rng.integers(0, 5, (B, N))
    st.num_obstacle_collisions[:] = rng.integers(0, 5, (B, N))
    st.dist_traveled_mean[:] = rng.uniform(0, 2, B)
    st.dist_traveled_stddev[:] = rng.uniform(0, 0.5, B)
    st.step[:] = rng.integers(0, 24, B)
    st.min_time[:] = rng.uniform(0, 1, (B, N))
    # the reference raises (argmin of an empty list, :921) when the nearest goal is taken, someone stands on it and no goal
    # is free; keep one goal free per env so that both sides stay defined
    st.occupied[np.arange(B), rng.integers(0, N, B)] = 0.5
    st = _fp32(st)
    orc = FormationOracle(cfg, B)
    orc.set_state(st)
    a = rng.integers(0, 5, (B, N))
    out, post = _run(host_exe, str(tmp_path), cfg, st, actions=a)
    ref = orc.step(a, autoreset=False)
    _compare(out, ref, post, orc.get_state())
    for branch in ("status_latched", "contact_force_suppressed", "vacated_goal", "info_unlatched", "subset_index_quirk"):
        assert orc.branch_hits.get(branch, 0) > 0, branch
