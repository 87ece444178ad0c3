"""Where does the fairness observation exceed 1e-5?  (diagnostic; prints ref / dev / step / inputs)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import fair_marl_b200 as fm
from oracle.navgraph import NavConfig, NavGraphOracle
from parity_util import device_state_to_nav, sim_config_from
cfg = NavConfig(num_agents=3, num_obstacles=3, fairness_reward=False)
B, N = 128, 3
env = fm.B200GraphVecEnv(sim_config_from(cfg), num_envs=B, device=0, seed=1)
orc = NavGraphOracle(cfg, B, seed=1)
env.reset_tensor(); orc.reset()
rng = np.random.default_rng(0)
for t in range(26):
    st = device_state_to_nav(env.get_state())
    orc.set_state(st)
    a = rng.integers(0, 5, (B, N))
    out = env.step_tensor(torch.as_tensor(a, dtype=torch.int32, device="cuda:0"))
    ref = orc.step(actions=a)
    d, r = out["obs"].cpu().numpy()[..., 6].astype(np.float64), ref["obs"][..., 6]
    err = np.abs(d - r) / np.maximum(np.abs(r), 1)
    bad = np.argwhere(err > 1e-5)
    for b, i in bad[:4]:
        post = device_state_to_nav(env.get_state())
        print(f"step {t} env {b} agent {i}: ref {r[b,i]:.9g} dev {d[b,i]:.9g} err {err[b,i]:.3e}  p_dist(pre) {st.p_dist[b]}  dtg(pre) {st.dists_to_goal[b]} treq {st.times_required[b]}")
        print("   oracle p_dist", orc.s.p_dist[b], " device p_dist", post.p_dist[b], "mean/std oracle", orc.s.dist_traveled_mean[b], orc.s.dist_traveled_stddev[b], "dev", post.dist_traveled_mean[b], post.dist_traveled_stddev[b])
print("done")
