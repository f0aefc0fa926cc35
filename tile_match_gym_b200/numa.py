"""Host-side placement for the host-buffer path: run the process that drives a GPU on the CPU cores of that GPU's NUMA
node, so that the pinned arrays it allocates (first touch) and the step kernel's PCIe writes into them stay on the GPU's
side of the socket interconnect.  Pure host plumbing; a no-op wherever the topology cannot be read."""
from __future__ import annotations

import os


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int):
    """NUMA node of a CUDA device from sysfs, or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:  # noqa: BLE001
        return None


def bind_to_gpu_node(device_index: int):
    """Restrict this process to the cores of the GPU's NUMA node.  Returns the node, or None if nothing was changed."""
    node = gpu_numa_node(device_index)
    if node is None or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read()) & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # noqa: BLE001
        return None
