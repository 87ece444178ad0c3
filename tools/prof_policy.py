"""torch.profiler breakdown of one RolloutCollector.collect at a given env batch (GPU).  usage: prof_policy.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fair_marl_b200 as fm
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = fm.B200GraphVecEnv(fm.SimConfig(), num_envs=B, seed=0)
pc = fm.PolicyConfig()
actor, critic = fm.DenseGraphActor(pc).cuda().eval(), fm.DenseGraphCritic(pc).cuda().eval()
col = fm.RolloutCollector(env, actor, critic)
col.warmup(); col.run(steps=3)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    col.run(steps=2)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
