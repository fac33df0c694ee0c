"""Host placement of a rank next to its GPU.

The host face of the step (solver.HostStaging) moves a few hundred MB per step between pinned host buffers and the
device.  On a two-socket box four of the eight GPUs hang off each socket; a rank whose pinned buffers were first
touched on the other socket sends every transfer across the socket interconnect (measured in round 2 at 8 GPUs:
result read-back at 9 GB/s on four ranks, 17 GB/s on the other four, against 55 GB/s for one rank alone).
bind_near_gpu() pins the calling process to the CPUs of the GPU's NUMA node and makes that node the preferred one
for its memory BEFORE the pinned buffers are allocated, so that first touch puts them there.

Everything here is best effort: it reports what it did and never raises (containers may hide sysfs or confine the
cpuset)."""
import ctypes
import os

MPOL_PREFERRED = 1
_SYS_set_mempolicy = {"x86_64": 238, "aarch64": 237}


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11]"""
    cpus = []
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus += list(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_pci_address(device_index):
    """'0000:1b:00.0' of a CUDA device (as sysfs spells it)"""
    import torch
    p = torch.cuda.get_device_properties(device_index)
    return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"


def gpu_numa_node(device_index):
    addr = gpu_pci_address(device_index)
    node = _read(f"/sys/bus/pci/devices/{addr}/numa_node")
    try:
        return int(node), addr
    except (TypeError, ValueError):
        return -1, addr


def set_preferred_node(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) through the raw syscall (libnuma is not in the image)"""
    nr = _SYS_set_mempolicy.get(os.uname().machine)
    if nr is None:
        return False
    libc = ctypes.CDLL(None, use_errno=True)
    nbits = 1024
    mask = (ctypes.c_ulong * (nbits // (8 * ctypes.sizeof(ctypes.c_ulong))))()
    w = 8 * ctypes.sizeof(ctypes.c_ulong)
    mask[node // w] |= 1 << (node % w)
    rc = libc.syscall(ctypes.c_long(nr), ctypes.c_int(MPOL_PREFERRED), ctypes.byref(mask), ctypes.c_ulong(nbits + 1))
    return rc == 0


def bind_near_gpu(device_index):
    """Pin this process to the CPUs of the NUMA node of CUDA device `device_index` and prefer that node for memory.
    Returns a dict describing what was done (goes into bench.py's e2e record)."""
    info = dict(bound=False)
    try:
        node, addr = gpu_numa_node(device_index)
        info.update(pci=addr, node=node)
        nodes = [d for d in (os.listdir("/sys/devices/system/node") if os.path.isdir("/sys/devices/system/node") else [])
                 if d.startswith("node") and d[4:].isdigit()]
        info["nodes_on_host"] = len(nodes)
        if node < 0 or len(nodes) < 2:
            info["why"] = "single NUMA node or no affinity reported"
            return info
        allowed = os.sched_getaffinity(0)
        local = set(parse_cpulist(_read(f"/sys/devices/system/node/node{node}/cpulist"))) & allowed
        if not local:
            info["why"] = "the cpuset of this process has no CPU on the GPU's node"
            return info
        os.sched_setaffinity(0, local)
        info["cpus"] = len(local)
        info["mempolicy"] = bool(set_preferred_node(node))
        info["bound"] = True
    except Exception as e:  # best effort by contract
        info["why"] = f"{type(e).__name__}: {e}"
    return info
