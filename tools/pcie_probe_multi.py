"""What can the HOST take?  Bare pinned D2H of one C2 step's results (105.6 MB) per rank, all ranks at once, no simulator:
the ceiling of `e2e` at N GPUs.  Prints one JSON line (rank 0): per-rank and aggregate GB/s, the
topology hints NVML gives (CPU affinity / NUMA node of each GPU), and the same with each rank's host thread and pinned
buffer bound to its GPU's CPU set.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe_multi.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 105_578_496


def affinity_of(idx):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        numa = None
        try:
            numa = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            pass
        return cpus, numa
    except Exception as exc:
        return None, str(exc)[:80]


def measure(reps=15, streams=2):
    src = torch.empty(n, dtype=torch.uint8, device=dev).random_()
    dst = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    ss = [torch.cuda.Stream() for _ in range(streams)]
    chunk = (n + streams - 1) // streams

    def once():
        for i, s in enumerate(ss):
            with torch.cuda.stream(s):
                lo, hi = i * chunk, min(n, (i + 1) * chunk)
                dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(3):
        once()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


cpus, numa = affinity_of(local)
dt_free = measure()
bound = None
if cpus:
    try:
        os.sched_setaffinity(0, cpus)             # host thread (and first-touch of the pinned buffer) on the GPU's CPU set
        bound = measure()
    except Exception as exc:
        bound = None
info = [None] * world
if world > 1:
    dist.all_gather_object(info, {"rank": rank, "cpus": (cpus[0], cpus[-1], len(cpus)) if cpus else None, "numa": numa})
else:
    info = [{"rank": 0, "cpus": (cpus[0], cpus[-1], len(cpus)) if cpus else None, "numa": numa}]
if rank == 0:
    line = {"probe": "pinned D2H, 105.6 MB per rank per step, 2 copy streams, all ranks concurrently (slowest rank)", "n_gpus": world,
            "ms_per_step": dt_free * 1e3, "gb_s_per_rank": n / dt_free / 1e9, "gb_s_aggregate": world * n / dt_free / 1e9,
            "e2e_ceiling_agent_steps_per_s": world * 65536 * 3 / dt_free,
            "bound_to_gpu_cpu_set": None if bound is None else {"ms_per_step": bound * 1e3, "gb_s_aggregate": world * n / bound / 1e9},
            "host_cpus": os.cpu_count(), "gpu_affinity": info}
    print(json.dumps(line), flush=True)
if world > 1:
    dist.destroy_process_group()
