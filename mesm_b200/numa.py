"""Host-side placement for the ingest path: run a rank's host threads - and therefore first-touch its pinned staging buffers -
on the NUMA node its GPU hangs off.

Why (round-1 scaling run, SCALE_r01.json): with one process per GPU and no placement, half of the ranks of a two-socket node end
up with their pinned host buffers on the other socket, every host->device DMA of those ranks crosses the inter-socket link, and
the aggregate ingest rate of 8 ranks collapsed to ~177 GB/s (22 GB/s per GPU against ~55 GB/s for one GPU alone).

``bind_to_gpu_node(device_index)`` asks NVML for the CPUs local to the GPU, intersects them with what the process is allowed to
use (cgroup / cpuset), pins the calling process to them and sets the memory policy to "prefer the local node" before any pinned
allocation is made.  Everything is best effort: on a single-node box, inside a restricted cpuset or without NVML it changes
nothing and says so in the returned record.  No reference counterpart (the reference trains on one GPU); belongs to the
prepare_batch_input drop-in (dataset/base.py:358-383).
"""
import ctypes
import os


def _cpus_of_node(node):
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            txt = f.read().strip()
    except OSError:
        return set()
    cpus = set()
    for part in txt.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def _gpu_numa_node(pci_bus_id):
    """NUMA node of a PCI device from sysfs (-1 when the platform does not say)."""
    bdf = pci_bus_id.lower()
    if len(bdf.split(":")[0]) == 8:           # NVML prints an 8-digit domain, sysfs uses 4
        bdf = bdf[4:]
    try:
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            return int(f.read().strip())
    except (OSError, ValueError):
        return -1


def _physical_index(index):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [v for v in vis.split(",") if v.strip() != ""]
        if index < len(ids) and ids[index].strip().isdigit():
            return int(ids[index])
    return index


def bind_to_gpu_node(device_index=0, set_mempolicy=True):
    """Pin this process to the CPUs local to GPU ``device_index`` and prefer that NUMA node for new pages.
    Returns a record of what was found / done (bench.py prints it)."""
    rec = {"gpu": device_index, "bound": False}
    try:
        allowed = os.sched_getaffinity(0)
        rec["cpus_allowed"] = len(allowed)
    except (AttributeError, OSError):
        rec["why"] = "sched_getaffinity unavailable"
        return rec
    local, node = set(), -1
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(_physical_index(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        node = _gpu_numa_node(bus if isinstance(bus, str) else bus.decode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        for w, mask in enumerate(words):
            for bit in range(64):
                if mask >> bit & 1:
                    local.add(w * 64 + bit)
    except Exception as e:                              # noqa: BLE001 - NVML missing / not permitted
        rec["nvml"] = f"{type(e).__name__}: {e}"
    if node >= 0 and not local:
        local = _cpus_of_node(node)
    rec["gpu_numa_node"] = node
    rec["cpus_local_to_gpu"] = len(local)
    try:
        with open("/sys/devices/system/node/online") as f:
            rec["nodes_online"] = f.read().strip()
    except OSError:
        pass
    use = local & allowed
    if not use or use == allowed:
        rec["why"] = "no placement needed / possible (local CPUs cover or miss the allowed set)"
        return rec
    try:
        os.sched_setaffinity(0, use)
        rec["bound"] = True
        rec["cpus_bound"] = len(use)
    except OSError as e:
        rec["why"] = f"sched_setaffinity: {e}"
        return rec
    if set_mempolicy and node >= 0:
        try:                                            # MPOL_PREFERRED = 1: new pages (incl. pinned staging buffers) come from `node`
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(node + 2))      # __NR_set_mempolicy on x86_64
            rec["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy errno %d" % ctypes.get_errno()
        except Exception as e:                          # noqa: BLE001
            rec["mempolicy"] = f"{type(e).__name__}: {e}"
    return rec
