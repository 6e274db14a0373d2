"""Multi-GPU plumbing: independent replicas, no data-path collective ("replicas only").

Block dequant is per-block parallel and a denoise step needs no cross-image data, so N GPUs = N processes
(torchrun, one per GPU) each running the whole hot path on its own batch.  torch.distributed is used ONLY to
start the timed region together (barrier) and to reduce the per-rank device times to their maximum; nothing
crosses NVLink on the data path.  Works with backend "nccl" (GPU box) and "gloo" (CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend: str | None = None):
    """Join the process group described by the torchrun environment (no-op for a single process)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float) -> float:
    """Slowest replica decides the job time (device-measured milliseconds in, maximum out)."""
    if not dist.is_initialized():
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    if not dist.is_initialized():
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def replica_seed(base: int, rank: int) -> int:
    """Every replica works on its own synthetic batch."""
    return base * 1000003 + rank


def aggregate_throughput(units_per_rank: float, local_ms: float) -> tuple[float, float]:
    """Weak scaling: whole-job units = sum over ranks, time = max over ranks.  Returns (units_per_s, max_ms)."""
    total_units = sum_over_ranks(units_per_rank)
    t_ms = max_over_ranks(local_ms)
    return total_units / (t_ms * 1e-3), t_ms


def parse_cpulist(text: str) -> list[int]:
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (format of /sys/devices/system/node/node*/cpulist)."""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pin this replica's host threads to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned staging
    buffers (first touch then places them on that node).  With N replicas sharing two sockets, the end-to-end path (pinned
    host -> H2D -> kernel -> D2H) otherwise crosses the socket interconnect for half of the ranks.  Best effort: returns what
    was done; never raises."""
    info = {"bound": False}
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as exc:      # no sysfs, no permission, old torch: stay unbound
        info["error"] = repr(exc)[:120]
    return info


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()
