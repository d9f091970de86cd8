"""Multi-GPU: frames are independent, so the path shards by frame with NO collective on the data path.
`torch.distributed` (NCCL on the GPUs, gloo in CPU tests) is used only to gather benchmark / checksum stats."""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def shard_range(num_frames: int, rank: int, world_size: int) -> range:
    """Contiguous block partition: rank r owns frames [r*n/W, (r+1)*n/W) (remainder spread over the first ranks)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(num_frames, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_frames(frames: Sequence, rank: int, world_size: int) -> List:
    return [frames[i] for i in shard_range(len(frames), rank, world_size)]


def gather_stats(local: Dict[str, float], device: torch.device | None = None) -> Dict[str, List[float]]:
    """all_gather of a small dict of floats (same keys on every rank) -> {key: [value per rank]}."""
    keys = sorted(local)
    t = torch.tensor([float(local[k]) for k in keys], dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {k: [float(v)] for k, v in zip(keys, t.tolist())}
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return {k: [float(o[i].item()) for o in out] for i, k in enumerate(keys)}


def max_over_ranks(value: float, device: torch.device | None = None) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa_node(device_index: int) -> str:
    """Best effort: pin this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers
    it allocates next (first touch) are local to that GPU's PCIe root.  With 8 ranks feeding 8 GPUs this is what keeps
    the host->device copies off the inter-socket link.  Returns a short description of what was done."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(f"{base}/numa_node").read().strip())
        cpulist = open(f"{base}/local_cpulist").read().strip()
    except Exception as e:  # noqa: BLE001 -- sysfs layout / permissions differ between hosts
        return f"numa: unavailable ({type(e).__name__})"
    cpus = set()
    for part in cpulist.split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    allowed = os.sched_getaffinity(0)
    use = sorted(cpus & allowed)
    if node < 0 or not use:
        return f"numa: node {node}, no usable local CPUs ({len(allowed)} allowed)"
    os.sched_setaffinity(0, use)
    return f"numa: node {node}, bound to {len(use)} local CPUs"
