"""Host-side placement for the frame feed: run the feeding process on the CPU cores next to its GPU
so that the pinned staging buffers it allocates (first touch) and the copies out of them stay on
that GPU's NUMA node.  With eight ranks streaming frames at once (BASELINE configs 3 and 4) the
host memory system is the shared resource — a remote-node buffer costs a hop over the inter-socket
link for every byte.  NVML knows the mapping (nvmlDeviceGetCpuAffinity); nothing here is
mandatory: without NVML or on a single-node host the process is left where it is."""
from __future__ import annotations

import os
from typing import Optional


def gpu_cpu_affinity(device_index: int) -> Optional[set]:
    """CPU ids NVML reports as local to the (physical) GPU ``device_index`` of this process."""
    try:
        import pynvml
    except ImportError:
        return None
    try:
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = device_index
        if visible:
            entry = visible.split(",")[device_index].strip()
            if entry.isdigit():
                index = int(entry)
            else:
                return None  # UUID lists: leave the placement alone
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        return cpus or None
    except Exception:  # noqa: BLE001 - NVML missing, no permission, MIG, ...
        return None


def bind_to_gpu(device_index: int) -> dict:
    """Restrict the calling process to the CPUs local to its GPU (intersected with what it may
    already use).  Returns what was done, for logs and bench records."""
    local = gpu_cpu_affinity(device_index)
    try:
        allowed = os.sched_getaffinity(0)
    except AttributeError:
        return {"bound": False, "reason": "no sched_getaffinity on this platform"}
    if not local:
        return {"bound": False, "reason": "NVML gave no CPU affinity", "cpus_allowed": len(allowed)}
    target = allowed & local
    if not target or target == allowed:
        return {"bound": False, "reason": "already local (or disjoint masks)",
                "cpus_allowed": len(allowed), "cpus_local": len(local)}
    os.sched_setaffinity(0, target)
    return {"bound": True, "cpus": len(target), "first_cpu": min(target), "last_cpu": max(target),
            "cpus_allowed_before": len(allowed)}
