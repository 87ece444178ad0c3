#!/bin/bash
# Round 2, GPU session T: full GPU suite (tightened fairness tolerance, SoA mode, finite guard), smoke.
set -u
OUT=gpurun_out/r02_t; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
python - <<'PY' 2>&1 | tail -3
import torch, sys
sys.path.insert(0, "tests")
import fair_marl_b200 as fm
from oracle.navgraph import NavConfig
from parity_util import sim_config_from
for N, B in ((3, 65536), (7, 262144)):
    env = fm.B200GraphVecEnv(sim_config_from(NavConfig(num_agents=N, num_obstacles=3)), num_envs=B, seed=1)
    env.reset_tensor()
    for fn in (env.observe_soa_tensor, env.observe_tensor):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        E = 2 * N + 3
        byts = B * 4 * (7 * N + 11 * N * E + E * E)
        print(N, B, fn.__name__, "us %.1f" % (e0.elapsed_time(e1) / 20 * 1e3), "GB/s of outputs %.0f" % (byts / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9))
    env.close()
PY
