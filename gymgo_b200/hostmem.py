"""Pinned host buffers close to the GPU (best effort, Linux only).

The host-facing step (BatchedGoEnv.host_stepper) moves up to 127 MB per ply over PCIe; with one process per GPU and
eight GPUs on a two-socket box, every rank allocating its pinned buffers on whatever NUMA node the kernel picks makes
the device->host copies of the far GPUs cross the socket interconnect.  This module asks the kernel to place a rank's
pinned pages on the NUMA node its GPU hangs off (set_mempolicy(MPOL_PREFERRED) around the allocation) and, where the
cpuset allows it, moves the calling thread to that node's cores.  Everything degrades to a plain pinned allocation
when the topology is not visible (containers often hide it); `describe()` says what happened."""
import ctypes
import os

import torch

_SYS_SET_MEMPOLICY = 238          # x86_64
_MPOL_DEFAULT, _MPOL_PREFERRED = 0, 1
_libc = None


def _syscall():
    global _libc
    if _libc is None:
        _libc = ctypes.CDLL(None, use_errno=True)
    return _libc.syscall


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_list(text):
    out = []
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.extend(range(int(lo), int(hi or lo) + 1))
    return out


def gpu_pci_address(device_index):
    p = torch.cuda.get_device_properties(device_index)
    try:
        return "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except AttributeError:
        return None


def gpu_numa_node(device_index):
    """NUMA node the GPU's PCIe root port belongs to, or None when the platform does not say"""
    addr = gpu_pci_address(device_index)
    if addr is None:
        return None
    node = _read("/sys/bus/pci/devices/%s/numa_node" % addr)
    try:
        node = int(node)
    except (TypeError, ValueError):
        return None
    return node if node >= 0 else None


def online_nodes():
    return _parse_list(_read("/sys/devices/system/node/online"))


def node_cpus(node):
    return _parse_list(_read("/sys/devices/system/node/node%d/cpulist" % node))


class prefer_node(object):
    """context manager: new pages of this thread are taken from `node` first (no-op when node is None or refused)"""

    def __init__(self, node):
        self.node, self.active = node, False

    def __enter__(self):
        if self.node is None or self.node >= 1024:
            return self
        mask = (ctypes.c_ulong * 16)()
        mask[self.node // 64] = 1 << (self.node % 64)
        try:
            rc = _syscall()(_SYS_SET_MEMPOLICY, _MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(1024))
            self.active = rc == 0
        except Exception:  # noqa: BLE001
            self.active = False
        return self

    def __exit__(self, *exc):
        if self.active:
            try:
                _syscall()(_SYS_SET_MEMPOLICY, _MPOL_DEFAULT, None, ctypes.c_ulong(0))
            except Exception:  # noqa: BLE001
                pass
        return False


def bind_thread_near(device_index):
    """move the calling thread to the cores of the GPU's NUMA node when the cpuset contains any of them; -> cpu count
    bound to, or 0 when nothing was changed"""
    node = gpu_numa_node(device_index)
    if node is None or not hasattr(os, "sched_setaffinity"):
        return 0
    want = set(node_cpus(node)) & set(os.sched_getaffinity(0))
    if not want:
        return 0
    try:
        os.sched_setaffinity(0, want)
    except OSError:
        return 0
    return len(want)


def usable_cores():
    """host threads this process may really use: affinity mask capped by the cgroup CPU quota"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = (_read("/sys/fs/cgroup/cpu.max") or "max").split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:  # noqa: BLE001
        pass
    return n


def codec_threads():
    """default thread count of the host codec for THIS process: the usable cores shared out over the ranks torchrun
    started on this node (LOCAL_WORLD_SIZE), so that one process per GPU does not oversubscribe the host"""
    try:
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:
        ranks = 1
    return max(1, usable_cores() // ranks)


def pinned_empty(shape, dtype, device_index=None):
    """pinned host tensor, its pages preferably on the NUMA node of CUDA device `device_index`"""
    node = None if device_index is None else gpu_numa_node(device_index)
    with prefer_node(node):
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
    return t


def describe(device_index):
    node = gpu_numa_node(device_index)
    return {"gpu_pci": gpu_pci_address(device_index), "gpu_numa_node": node, "numa_nodes_online": online_nodes(),
            "cpus_allowed": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None,
            "mems_allowed": _read("/sys/fs/cgroup/cpuset.mems.effective")}
