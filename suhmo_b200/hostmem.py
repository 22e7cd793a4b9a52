"""Host-side placement for the pinned staging buffers of a rank.

The reference runs one MPI rank per core set and leaves placement to `mpirun --bind-to`; here one process drives one GPU
and stages every field through pinned host memory (`LevelData.upload_packed` / `download_packed`). Pinned pages are placed
on the NUMA node of the thread that allocates them, so a rank that allocates from the wrong socket pulls its 6 GB per head
solve across the inter-socket link and shares one memory controller with every other rank. `bind_to_gpu_numa` pins the
calling process to the CPUs next to its GPU BEFORE the staging buffers are allocated. No effect (and no error) where the
topology is not visible (containers without /sys NUMA information, single-node hosts).
"""
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def parse_cpulist(s):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11]"""
    out = []
    for part in (s or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def gpu_pci_bus_id(device):
    """PCI address of a CUDA device as sysfs spells it (0000:1b:00.0)"""
    import torch
    p = torch.cuda.get_device_properties(device)
    dom = getattr(p, "pci_domain_id", 0)
    return "%04x:%02x:%02x.0" % (dom, p.pci_bus_id, p.pci_device_id)


def gpu_numa_node(device):
    """NUMA node of the GPU's PCIe root, or None when the platform does not say"""
    s = _read("/sys/bus/pci/devices/%s/numa_node" % gpu_pci_bus_id(device))
    try:
        n = int(s)
    except (TypeError, ValueError):
        return None
    return n if n >= 0 else None


def node_cpus(node):
    return parse_cpulist(_read("/sys/devices/system/node/node%d/cpulist" % node))


def bind_to_gpu_numa(device, ranks_on_node=1, slot=0):
    """Restrict the calling process to the CPUs of the NUMA node its GPU hangs off (first-touch then places pinned pages
    there). With several ranks on the same node the node's CPUs are split evenly (`slot` of `ranks_on_node`) so that the
    ranks' host threads do not sit on top of each other. Returns a dict describing what was done."""
    info = {"device": int(device), "numa_node": None, "cpus": None, "bound": False}
    try:
        node = gpu_numa_node(device)
        info["numa_node"] = node
        if node is None:
            return info
        allowed = set(os.sched_getaffinity(0))
        info["_prev"] = sorted(allowed)
        cpus = [c for c in node_cpus(node) if c in allowed]
        if not cpus:
            return info
        if ranks_on_node > 1:
            per = max(1, len(cpus) // ranks_on_node)
            mine = cpus[slot * per:(slot + 1) * per]
            cpus = mine or cpus
        os.sched_setaffinity(0, cpus)
        info["cpus"] = "%d-%d (%d)" % (cpus[0], cpus[-1], len(cpus))
        info["bound"] = True
    except Exception as e:  # placement is an optimisation: never fail a solve over it
        info["error"] = repr(e)
    return info


def restore(info):
    """give the process back the CPUs it had before `bind_to_gpu_numa` (pages already allocated stay where they are)"""
    prev = info.pop("_prev", None) if isinstance(info, dict) else None
    if prev and info.get("bound"):
        try:
            os.sched_setaffinity(0, prev)
        except OSError:
            pass
