"""Host-side placement for the end-to-end path: a rank's pinned staging buffers and its copy-issuing thread should live on
the NUMA node its GPU's PCIe root hangs off, otherwise the H2D stream of several ranks funnels through one socket's memory
controllers and inter-socket link (round 1 measured 49 GB/s for one rank but only 22 GB/s per rank at 8 ranks).

Best effort and Linux only: when the topology cannot be read nothing is changed.  torch is used for the device properties
only.
"""
from __future__ import annotations

import os
import subprocess


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


def gpu_pci_bus_id(index: int):
    """'0000:1b:00.0' of CUDA device ``index`` (torch device properties, nvidia-smi as a fallback)."""
    try:
        import torch

        p = torch.cuda.get_device_properties(index)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        pass
    try:
        uuid_order = os.environ.get("CUDA_VISIBLE_DEVICES")
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True,
                             text=True, timeout=20).stdout.split()
        ids = [int(v) for v in uuid_order.split(",")] if uuid_order and all(v.isdigit() for v in uuid_order.split(",")) else None
        bus = out[ids[index]] if ids else out[index]
        return bus.lower()[-12:]
    except Exception:
        return None


def gpu_numa_node(index: int):
    bus = gpu_pci_bus_id(index)
    if not bus:
        return None
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
    except Exception:
        return None
    return node if node >= 0 else None


def bind_to_gpu(index: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node (pinned allocations made afterwards land there by first
    touch).  Returns what was found / done, for the bench line."""
    info = {"gpu": index, "pci": gpu_pci_bus_id(index), "numa_node": None, "cpus_bound": None}
    node = gpu_numa_node(index)
    info["numa_node"] = node
    if node is None or not hasattr(os, "sched_setaffinity"):
        return info
    try:
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus_bound"] = len(cpus)
    except Exception as ex:  # pragma: no cover
        info["error"] = repr(ex)[:120]
    return info
