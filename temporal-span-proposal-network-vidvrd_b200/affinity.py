"""CPU placement of one rank of the serving loop.

The e2e loop of a rank is a host thread that issues copies and graph launches plus the pinned buffers those
copies read and write.  On a multi-socket box both should sit on the NUMA node of the rank's GPU; on any box the
ranks should not all run on the same cores (``torchrun`` leaves every rank on the full mask, and with
``OMP_NUM_THREADS=1`` the scheduler is free to stack them).  ``bind_to_gpu`` binds the calling process to the
cores sysfs lists as local to the GPU's PCI function and, when that list is the whole machine (a VM that hides
the topology), to this rank's even share of it.  Call it before allocating pinned memory (first touch).
"""
from __future__ import annotations

import os
from typing import List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_local_cpus(cuda_index: int) -> Optional[List[int]]:
    """Cores local to the GPU according to ``/sys/bus/pci/devices/<bdf>/local_cpulist`` (None if unknown)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(cuda_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            cpus = _parse_cpulist(f.read())
        return cpus or None
    except Exception:  # noqa: BLE001
        return None


def bind_to_gpu(local_rank: int, local_world: int) -> Optional[dict]:
    """Bind this process; returns ``{"cpus": [...], "source": ...}`` (None when nothing could be done)."""
    if not hasattr(os, "sched_setaffinity"):
        return None
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except OSError:
        return None
    local = gpu_local_cpus(local_rank)
    source = "sysfs local_cpulist"
    cpus = [c for c in (local or []) if c in allowed]
    if not cpus or len(cpus) == len(allowed):
        # topology hidden (or one node): give every rank its own contiguous share of the allowed cores
        share = max(len(allowed) // max(local_world, 1), 1)
        cpus = allowed[(local_rank * share) % len(allowed):][:share] or allowed
        source = "even share of %d allowed cores" % len(allowed)
    elif local_world > 1:
        # ranks whose GPUs share a node split that node's cores
        peers = [r for r in range(local_world) if gpu_local_cpus(r) == local]
        if len(peers) > 1 and len(cpus) >= len(peers):
            share = len(cpus) // len(peers)
            cpus = cpus[peers.index(local_rank) * share:][:share]
            source += ", split over %d ranks" % len(peers)
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return None
    return {"cpus": "%d-%d (%d)" % (cpus[0], cpus[-1], len(cpus)), "source": source}
