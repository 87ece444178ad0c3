"""Multi-rank host logic on CPU: world_size-2 ``gloo`` process group.

Checks, without a GPU, that (i) env sharding by global index reproduces the unsharded batch
bit for bit (run with the numpy oracle as the per-rank simulator, which keys its reset streams by
global env index exactly like the device kernels), and (ii) ``EpisodeStats`` sum-all-reduces the
statistics vector in the documented layout.  The oracle is used here as the checker only.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N, O, B_TOTAL, STEPS = 3, 3, 22, 27       # 27 steps: one auto-reset inside the run


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_rollout(offset, count, seed=5):
    from oracle.navgraph import INFO_KEYS, NavConfig, NavGraphOracle
    cfg = NavConfig(num_agents=N, num_obstacles=O)
    orc = NavGraphOracle(cfg, count, seed=seed, env_offset=offset)
    out = orc.reset()
    acts = np.random.default_rng(99).integers(0, 5, (STEPS, B_TOTAL, N))
    K = 15 * N + 2
    stats = np.zeros(K)
    for t in range(STEPS):
        out = orc.step(actions=acts[t, offset:offset + count])
        stats[:N] += out["reward"].sum(axis=0)
        term = out["done"].all(axis=1)
        info = np.stack([out["info"][k] for k in INFO_KEYS], axis=-1)          # [B, N, 14]
        stats[N:15 * N] += info[term].sum(axis=0).reshape(-1)
        stats[15 * N] += term.sum()
        stats[15 * N + 1] += count
    return out, stats


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import fair_marl_b200 as fm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        off, cnt = fm.shard_range(B_TOTAL, world, rank)
        out, stats = _run_rollout(off, cnt)
        es = fm.EpisodeStats(N, device=torch.device("cpu"))
        es.all_reduce_async(torch.from_numpy(stats))
        total = es.result().numpy()
        # gather the final observations of every shard on rank 0
        parts = [None] * world
        dist.all_gather_object(parts, (off, cnt, out["obs"], out["node_obs"], out["adj"], out["reward"]))
        if rank == 0:
            q.put((total, parts, es.summary()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_equals_one_batch_and_stats_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    total, parts, summary = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    full_out, full_stats = _run_rollout(0, B_TOTAL)
    # shards tile the batch contiguously and reproduce it bit for bit
    parts = sorted(parts, key=lambda x: x[0])
    assert [p[0] for p in parts] == [0, 11] and sum(p[1] for p in parts) == B_TOTAL
    for k, name in enumerate(("obs", "node_obs", "adj", "reward")):
        got = np.concatenate([p[2 + k] for p in parts], axis=0)
        assert np.array_equal(got, full_out[name]), name
    # all-reduced statistics == statistics of the unsharded batch (sums of the same float64 terms,
    # grouped differently: equal to rounding)
    np.testing.assert_allclose(total, full_stats, rtol=1e-12, atol=1e-12)
    assert summary["episodes"] == B_TOTAL and summary["env_steps"] == B_TOTAL * STEPS
    assert len(summary["Dist_to_goal"]) == N


def test_shard_range_edge_cases():
    import fair_marl_b200 as fm
    assert fm.shard_range(10, 4, 0) == (0, 3) and fm.shard_range(10, 4, 3) == (8, 2)
    assert fm.shard_range(3, 8, 7) == (3, 0)                 # more ranks than envs: empty tail shards
    with pytest.raises(ValueError):
        fm.shard_range(8, 2, 2)
