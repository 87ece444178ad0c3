"""Diagnostic: first mismatch between two kernel mappings over a rollout (run on the GPU box)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fair_marl_b200 as fm
N, O, B = 3, 3, 1000
a_map = sys.argv[1] if len(sys.argv) > 1 else "aw"
halves = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kw = dict(num_agents=N, num_obstacles=O, goal_rew=30.0, collision_rew=30.0, episode_length=7, info_every_step=True)
e1 = fm.B200GraphVecEnv(fm.SimConfig(mapping=a_map, **kw), num_envs=B, seed=5, env_offset=11)
e2 = fm.B200GraphVecEnv(fm.SimConfig(mapping="group", **kw), num_envs=B, seed=5, env_offset=11)
e1.reset_tensor(); e2.reset_tensor()
rng = np.random.default_rng(2)
for t in range(23):
    a = rng.integers(0, 5, (B, N))
    if t % 3 == 0: a[: B // 2] = 0
    at = torch.as_tensor(a, dtype=torch.int32, device="cuda")
    o1, o2 = e1.step_tensor(at), e2.step_tensor(at)
    bad = False
    for k in ("obs", "node_obs", "adj_env", "reward", "done", "info"):
        x, y = o1[k].cpu().numpy(), o2[k].cpu().numpy()
        neq = ~((x == y) | (np.isnan(x) & np.isnan(y)))
        if neq.any():
            idx = np.argwhere(neq)
            print(f"t={t} {k}: {neq.sum()} mismatches; last-axis histogram {np.bincount(idx[:, -1])}; first {idx[0]} {x[tuple(idx[0])]!r} vs {y[tuple(idx[0])]!r}")
            bad = True
    s1, s2 = e1.get_state(), e2.get_state()
    for k in s1:
        x, y = s1[k].cpu().numpy(), s2[k].cpu().numpy()
        neq = ~((x == y) | (np.isnan(x) & np.isnan(y)))
        if neq.any():
            idx = np.argwhere(neq); print(f"t={t} state {k}: {neq.sum()} mismatches first {idx[0]} {x[tuple(idx[0])]!r} vs {y[tuple(idx[0])]!r}"); bad = True
    if bad: break
else:
    print("all equal")
