"""Host-side placement for the host-buffer API (``FrontEnd.features_host``): on a multi-socket box a rank whose
pinned staging buffers live on the far NUMA node moves its 400 MB per step across the socket interconnect, and all
ranks of a node then share that link.  ``bind_to_device_numa_node`` pins the calling process to the CPUs that are
local to its GPU (``/sys/bus/pci/devices/<bus id>/local_cpulist``) BEFORE the pinned buffers are allocated, so the
kernel's first-touch policy places them on the GPU's own node.  No counterpart in the reference (its DataLoader is
single-process, ``dataloader.py:172``)."""
from __future__ import annotations

import contextlib
import os
from typing import Optional

import torch


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def device_local_cpus(device: torch.device | int) -> Optional[set[int]]:
    """CPUs local to the PCI device of a CUDA device, or ``None`` when the platform does not say."""
    props = torch.cuda.get_device_properties(device)
    try:
        path = f"/sys/bus/pci/devices/{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0/local_cpulist"
        with open(path) as f:
            cpus = _parse_cpulist(f.read())
    except (OSError, AttributeError, ValueError):
        return None
    return cpus or None


def bind_to_device_numa_node(device: torch.device | int) -> dict:
    """Restrict the calling process to the CPUs local to ``device`` (intersected with the CPUs it may already use).
    Returns ``{"bound": bool, "cpus": n, "why": ...}``; never raises: placement is an optimisation."""
    try:
        local = device_local_cpus(device)
        if not local:
            return {"bound": False, "cpus": len(os.sched_getaffinity(0)), "why": "no local_cpulist for the device"}
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        if not target:
            return {"bound": False, "cpus": len(allowed), "why": "device-local CPUs are outside this process's cpuset"}
        if target != allowed:
            os.sched_setaffinity(0, target)
        return {"bound": True, "cpus": len(target), "why": "device-local CPUs"}
    except Exception as e:  # noqa: BLE001 - e.g. a sandbox without sched_setaffinity
        return {"bound": False, "cpus": -1, "why": f"{type(e).__name__}: {e}"}


@contextlib.contextmanager
def device_local_affinity(device: torch.device | int):
    """``with device_local_affinity(dev) as info: buf = torch.empty(...).pin_memory()`` - binds for the duration of
    the block only (pinned pages are placed when they are allocated), then restores the previous CPU set so that
    host-side work after it (data generation, a CPU baseline) keeps every core."""
    try:
        before = os.sched_getaffinity(0)
    except Exception:  # noqa: BLE001
        before = None
    info = bind_to_device_numa_node(device)
    try:
        yield info
    finally:
        if before is not None and info.get("bound"):
            with contextlib.suppress(Exception):
                os.sched_setaffinity(0, before)
