"""Row sharding of a batch across the GPUs of one box: one process per GPU (torchrun), NCCL only
to reassemble results (the reference has no distributed code at all).

The hot path is embarrassingly parallel over rows: every rank evaluates a contiguous block of
rows with the same kernels and the only shared data is the <= 2 KB chain table.  There is no
data-path collective; `gather_rows` is the optional reassembly step (all-gather, or gather to one
rank) and is reported separately from kernel throughput.  With backend "gloo" (CPU tensors) the
same helpers run in the CPU test-suite.
"""
from __future__ import annotations

from typing import List, Tuple


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index: int):
    """NUMA node of the host bridge a GPU hangs off (sysfs), or None when the platform does not say."""
    import os

    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read())
        return node if node >= 0 and os.path.isdir(f"/sys/devices/system/node/node{node}") else None
    except Exception:
        return None


def bind_to_gpu_numa(device_index: int):
    """Restrict this process to the CPUs of the NUMA node next to its GPU.

    One process per GPU moves ~55 GB/s of results into host memory on the end-to-end path; with eight
    ranks on a two-socket box the pinned buffers must sit on the socket the GPU's PCIe root belongs to, or
    half of the traffic crosses the inter-socket link.  Linux places pages on the node of the thread that
    first touches them, so narrowing the CPU affinity BEFORE the pinned buffers are allocated is enough
    (no libnuma needed).  Returns {"node", "cpus"} or None when nothing was changed."""
    import os

    node = gpu_numa_node(device_index)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read()) & set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception:
        return None


def shard_bounds(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) row range of `rank`; the first n_rows % world_size ranks get
    one extra row."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, rem = divmod(int(n_rows), world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sizes(n_rows: int, world_size: int) -> List[int]:
    return [shard_bounds(n_rows, world_size, r)[1] - shard_bounds(n_rows, world_size, r)[0] for r in range(world_size)]


def gather_rows(local, n_rows: int, dst=None, group=None):
    """Reassemble row shards (dim 0) that were cut with `shard_bounds`.

    dst=None : all-gather -- every rank returns the full (n_rows, ...) tensor.
    dst=k    : gather -- rank k returns the full tensor, other ranks return None.
    Works for ragged shards (n_rows not divisible by the world size).
    """
    import torch
    import torch.distributed as td

    ws = td.get_world_size(group)
    rank = td.get_rank(group)
    sizes = shard_sizes(n_rows, ws)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: local shard has {local.shape[0]} rows, expected {sizes[rank]}")
    local = local.contiguous()
    tail = tuple(local.shape[1:])
    mx = max(sizes)
    if local.shape[0] != mx:  # ragged shards: pad to the largest one so every rank contributes equal counts
        pad = torch.zeros((mx - local.shape[0],) + tail, dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    if dst is None:
        out = torch.empty((ws * mx,) + tail, dtype=local.dtype, device=local.device)
        td.all_gather_into_tensor(out, local, group=group)
    else:
        parts = [torch.empty((mx,) + tail, dtype=local.dtype, device=local.device) for _ in range(ws)] if rank == dst else None
        td.gather(local, parts, dst=dst, group=group)
        if rank != dst:
            return None
        out = torch.cat(parts, dim=0)
    if len(set(sizes)) == 1:
        return out
    out = out.reshape((ws, mx) + tail)
    return torch.cat([out[r, :sizes[r]] for r in range(ws)], dim=0)


def sharded_fkine_jacob0(ets, q_local, n_rows: int = None, gather: bool = False, base=None, tool=None, group=None):
    """Evaluate pose + Jacobian for this rank's rows; optionally all-gather the results."""
    T, J = ets.fkine_jacob0(q_local, base=base, tool=tool)
    if gather:
        return gather_rows(T, n_rows, group=group), gather_rows(J, n_rows, group=group)
    return T, J
