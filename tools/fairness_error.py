"""How far can the obs[6] / fairness tolerance tighten?  Rolls device and oracle side by side (oracle re-seeded from the
device's fp32 state every step) and prints the distribution of |dev - ref| / max(|ref|, 1) of the fairness observation,
overall and binned by |ref| (diagnostic for tests/parity_util.py::assert_fairness_close; writes JSON lines)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import fair_marl_b200 as fm
from oracle.navgraph import NavConfig, NavGraphOracle
from parity_util import device_state_to_nav, sim_config_from


def run(N, O, B, steps, seek):
    cfg = NavConfig(num_agents=N, num_obstacles=O)
    env = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=B, device=0, seed=1)
    orc = NavGraphOracle(cfg, B, seed=1)
    env.reset_tensor(); orc.reset()
    rng = np.random.default_rng(0)
    errs, refs, rew_err = [], [], []
    for t in range(steps):
        st = device_state_to_nav(env.get_state())
        orc.set_state(st)
        a = rng.integers(0, 5, (B, N))
        if seek:                                  # steer at the goals: latches, equal travelled distances, contacts
            g = np.take_along_axis(st.landmark_pos, st.goal_match[..., None], axis=1) - st.pos
            s = np.where(np.abs(g[..., 0]) > np.abs(g[..., 1]), np.where(g[..., 0] > 0, 1, 2), np.where(g[..., 1] > 0, 3, 4))
            a = np.where(rng.random((B, N)) < 0.15, a, s)
        out = env.step_tensor(torch.as_tensor(a, dtype=torch.int32, device="cuda:0"))
        ref = orc.step(actions=a)
        d, r = out["obs"].cpu().numpy()[..., 6].astype(np.float64), ref["obs"][..., 6]
        errs.append(np.abs(d - r) / np.maximum(np.abs(r), 1)); refs.append(np.abs(r))
        rew_err.append(np.abs(out["reward"].cpu().numpy() - ref["reward"]) / np.maximum(np.abs(ref["reward"]), 1))
    e, r = np.concatenate([x.ravel() for x in errs]), np.concatenate([x.ravel() for x in refs])
    line = {"N": N, "O": O, "B": B, "steps": steps, "seek": seek, "samples": int(e.size), "max_err": float(e.max()),
            "p999": float(np.quantile(e, 0.999)), "over_1e-5": int((e > 1e-5).sum()), "over_1e-6": int((e > 1e-6).sum()),
            "max_reward_err": float(np.concatenate([x.ravel() for x in rew_err]).max()), "bins": {}}
    for lo, hi in ((0, 33), (33, 100), (100, 1000), (1000, 1e9)):
        m = (r >= lo) & (r < hi)
        if m.any():
            line["bins"][f"{lo}-{hi}"] = {"n": int(m.sum()), "max_err": float(e[m].max()), "max_err_over_ref": float((e[m] / np.maximum(r[m], 1)).max())}
    env.close()
    return line


if __name__ == "__main__":
    for N, O, B, steps, seek in ((3, 3, 4096, 50, False), (3, 3, 4096, 50, True), (7, 3, 1024, 50, True)):
        print(json.dumps(run(N, O, B, steps, seek)), flush=True)
