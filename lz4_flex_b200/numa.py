"""Host-side placement for the one-process-per-GPU runs: bind the calling process (and therefore the pinned staging
buffers it allocates afterwards — first touch) to the NUMA node its GPU hangs off.  On an 8-GPU box the GPUs sit
behind two sockets; without this every rank's H2D/D2H traffic crosses the inter-socket link half of the time and
the end-to-end numbers stop scaling (e2e efficiency 0.69 at N=8)."""
from __future__ import annotations

import os


def _parse_cpulist(text: str) -> list[int]:
    cpus: list[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(device_index: int) -> int | None:
    """NUMA node of a CUDA device from its PCI address (sysfs); None when unknown."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_node(device_index: int, local_rank: int = 0, local_world: int = 1) -> dict:
    """sched_setaffinity to the CPUs of the GPU's NUMA node (all of them: the host side of this codec is a few copy
    threads, not a compute pool).  Returns what was done, for the bench record."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        info["original_affinity"] = sorted(os.sched_getaffinity(0))     # restore_affinity() undoes the binding
    except AttributeError:
        return info
    node = gpu_numa_node(device_index)
    if node is None:
        return info
    try:
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update({"numa_node": node, "cpus": len(allowed), "bound": True})
    except Exception:
        pass
    return info


def restore_affinity(info: dict) -> None:
    """Undo bind_to_gpu_node (e.g. before timing an all-cores CPU baseline in the same process)."""
    if info.get("bound") and info.get("original_affinity"):
        try:
            os.sched_setaffinity(0, info["original_affinity"])
        except Exception:
            pass
