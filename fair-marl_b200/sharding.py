"""Env sharding over ranks and the per-step episode-statistics all-reduce.

Envs are independent (one ``World`` per worker process in the reference, env_wrappers.py:951-967),
so rank r owns the contiguous global env range ``shard_range(B_total, world, r)`` and the only
exchange is a sum all-reduce of the small statistics vector -- NCCL on GPUs,
gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

from fair_marl_b200 import _lib


def shard_range(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(offset, count) of the contiguous env range owned by ``rank``; sizes differ by at most 1."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} not in [0, {world_size})")
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def stats_layout(num_agents: int):
    """Slices into the statistics vector (include/fairmarl.h, fm_stats_read)."""
    n = num_agents
    return {"reward_sum": slice(0, n), "info_sum": slice(n, 15 * n), "episodes": 15 * n, "env_steps": 15 * n + 1,
            "len": 15 * n + 2}


class EpisodeStats:
    """Sum all-reduce of the statistics vector across ranks, issued on a side stream so the simulator
    stream never waits for it; ``result()`` synchronises and returns the global vector."""

    def __init__(self, num_agents: int, device=None, group=None):
        import torch
        self.torch = torch
        self.n = num_agents
        self.layout = stats_layout(num_agents)
        self.device = device if device is not None else torch.device("cpu")
        self.group = group
        self.buf = torch.zeros(self.layout["len"], dtype=torch.float64, device=self.device)
        self.side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._work = None

    def all_reduce_async(self, local_vec) -> None:
        """local_vec: float64 tensor [K] on ``device`` (already produced on the current stream)."""
        torch = self.torch
        import torch.distributed as dist
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream(self.device))
            local_vec.record_stream(self.side)          # the caller may drop `local_vec` as soon as this returns
            with torch.cuda.stream(self.side):
                self.buf.copy_(local_vec, non_blocking=True)
                if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                    dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self.buf.copy_(local_vec)
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)

    def join(self) -> None:
        """Make the current stream wait for the all-reduce in flight (no host synchronisation)."""
        if self.side is not None:
            self.torch.cuda.current_stream(self.device).wait_stream(self.side)

    def result(self):
        if self.side is not None:
            self.side.synchronize()
        return self.buf.clone()

    def summary(self, vec=None) -> dict:
        """Per-agent means over finished episodes, keyed like base_runner.process_infos aggregates."""
        v = (self.result() if vec is None else vec).cpu().numpy()
        lay, n = self.layout, self.n
        episodes = max(v[lay["episodes"]], 1.0)
        info = v[lay["info_sum"]].reshape(n, 14) / episodes
        out = {"episodes": float(v[lay["episodes"]]), "env_steps": float(v[lay["env_steps"]]),
               "reward_per_env_step": (v[lay["reward_sum"]] / max(v[lay["env_steps"]], 1.0)).tolist()}
        for k, name in enumerate(_lib.INFO_KEYS):
            out[name] = info[:, k].tolist()
        return out
