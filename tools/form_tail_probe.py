"""Which rare path sets the formation logic kernel's duration?  Per-step times (logic kernel alone) with the auto-reset
on / off and with goal latching disabled (min_dist_thresh = 0)."""
import ctypes as C, sys, torch
sys.path.insert(0, "/root/repo")
import fair_marl_b200 as fm
from fair_marl_b200 import _lib
B = 65536
a = torch.randint(0, 5, (25, B, 3), device="cuda", dtype=torch.int32)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(61)]
for label, kw in (("default", {}), ("auto_reset off", {"auto_reset": False}), ("no latching (min_dist_thresh 0)", {"min_dist_thresh": 0.0})):
    cfg = fm.FormationSimConfig(num_agents=3, num_obstacles=3, goal_rew=30.0, collision_rew=30.0, episode_length=25,
                                fairness_reward=True, info_every_step=False, **kw)
    env = fm.B200FormationVecEnv(cfg, num_envs=B, device=0, seed=0, num_slots=8)
    env.reset_tensor()
    outs = [_lib.FmOutputs(b["obs"].data_ptr(), None, None, b["reward"].data_ptr(), b["done"].data_ptr(), b["info"].data_ptr()) for b in env._slots]
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for k in range(40):
        _lib.check(env.lib.fm_formation_step(env._h, a[k % 25].data_ptr(), C.byref(outs[k % 8]), stream), "step")
    torch.cuda.synchronize()
    evs[0].record()
    for k in range(60):
        _lib.check(env.lib.fm_formation_step(env._h, a[k % 25].data_ptr(), C.byref(outs[k % 8]), stream), "step")
        evs[k + 1].record()
    torch.cuda.synchronize()
    st = env.get_state()
    print(label, "| episodes so far (min/max):", int(st["episode"].min()), int(st["episode"].max()), "|",
          " ".join("%.0f" % (evs[k].elapsed_time(evs[k + 1]) * 1e3) for k in range(60)))
    env.close()
