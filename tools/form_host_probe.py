import time, torch, sys
sys.path.insert(0, "/root/repo")
import fair_marl_b200 as fm
cfg = fm.FormationSimConfig(num_agents=3, num_obstacles=3, goal_rew=30.0, collision_rew=30.0, episode_length=25, fairness_reward=True, info_every_step=False)
for B in (256, 65536):
    env = fm.B200FormationVecEnv(cfg, num_envs=B, device=0, seed=0, num_slots=8)
    a = torch.randint(0, 5, (25, B, 3), device="cuda", dtype=torch.int32)
    env.reset_tensor()
    for k in range(30): env.step_tensor(a[k % 25])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(300): env.step_tensor(a[k % 25])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(B, "host issue us/step", (t1 - t0) / 300 * 1e6, "total us/step", (t2 - t0) / 300 * 1e6)
    # graph replay
    g = torch.cuda.CUDAGraph()
    s0 = env._slot
    with torch.cuda.graph(g):
        for k in range(25): env.step_tensor(a[k])
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay(); torch.cuda.synchronize()
    ev0.record()
    for _ in range(12): g.replay()
    ev1.record(); torch.cuda.synchronize()
    print(B, "graph us/step", ev0.elapsed_time(ev1) / 300 * 1e3)
    env.close()

# logic kernel alone in steady state (FmOutputs without node_obs / adj: the image kernel is skipped)
import ctypes as C
from fair_marl_b200 import _lib
B = 65536
env = fm.B200FormationVecEnv(cfg, num_envs=B, device=0, seed=0, num_slots=8)
a = torch.randint(0, 5, (25, B, 3), device="cuda", dtype=torch.int32)
env.reset_tensor()
outs = [_lib.FmOutputs(b["obs"].data_ptr(), None, None, b["reward"].data_ptr(), b["done"].data_ptr(), b["info"].data_ptr()) for b in env._slots]
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def logic_steps(n):
    for k in range(n):
        _lib.check(env.lib.fm_formation_step(env._h, a[k % 25].data_ptr(), C.byref(outs[k % 8]), stream), "step")
logic_steps(30); torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); logic_steps(300); ev1.record(); torch.cuda.synchronize()
print("logic kernel alone, steady state us/step", ev0.elapsed_time(ev1) / 300 * 1e3)

# per-step durations over two episodes (logic kernel alone): where does the steady-state average come from?
evs = [torch.cuda.Event(enable_timing=True) for _ in range(61)]
torch.cuda.synchronize()
st = env.get_state(); print("episode phase (step of env 0):", int(st["step"][0]))
evs[0].record()
for k in range(60):
    _lib.check(env.lib.fm_formation_step(env._h, a[k % 25].data_ptr(), C.byref(outs[k % 8]), stream), "step")
    evs[k + 1].record()
torch.cuda.synchronize()
print("per-step us:", " ".join("%.0f" % (evs[k].elapsed_time(evs[k + 1]) * 1e3) for k in range(60)))

def per_step(label, act_of, n=60):
    torch.cuda.synchronize()
    evs[0].record()
    for k in range(n):
        _lib.check(env.lib.fm_formation_step(env._h, act_of(k).data_ptr(), C.byref(outs[k % 8]), stream), "step")
        evs[k + 1].record()
    torch.cuda.synchronize()
    print(label, "phase", int(st["step"][0]), ":", " ".join("%.0f" % (evs[k].elapsed_time(evs[k + 1]) * 1e3) for k in range(n)))
st = env.get_state()
per_step("same action slice a[0]", lambda k: a[0])
st = env.get_state()
per_step("slices shifted by 7", lambda k: a[(k + 7) % 25])
zero = torch.zeros_like(a[0])
st = env.get_state()
per_step("no-op actions", lambda k: zero)
