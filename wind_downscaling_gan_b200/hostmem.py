"""Host-memory placement for the host <-> device pipeline (predict_host / engine.downscale_series).

On a multi-socket box every rank's page-locked buffers land on the NUMA node its process happens to run on; with 8 ranks
on one node the eight H2D/D2H streams share that node's memory controllers and the far GPUs copy across the socket link
(round-1 scaling run: per-GPU H2D fell from 52 GB/s at 1 GPU to ~21 GB/s at 8, ~165 GB/s aggregate).  Binding each rank to
the CPUs of ITS GPU's NUMA node before the pinned buffers are allocated (first-touch placement) keeps every copy local.
Nothing here touches the data path; on a single-node box it is a no-op."""
import os


def gpu_numa_node(device_index):
    """NUMA node of CUDA device `device_index` (sysfs numa_node of its PCI function), or None if the box does not say."""
    import torch
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except (OSError, AttributeError, ValueError, RuntimeError):
        return None


def node_cpus(node):
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


def bind_to_gpu_numa_node(device_index):
    """Restricts this process to the CPUs of the GPU's NUMA node (within its current affinity mask).  Call before allocating
    pinned buffers.  Returns a small report dict for logs / bench lines."""
    node = gpu_numa_node(device_index)
    report = {"gpu": int(device_index), "numa_node": node, "bound": False}
    if node is None or not hasattr(os, "sched_setaffinity"):
        return report
    allowed = os.sched_getaffinity(0)
    cpus = node_cpus(node) & allowed
    report["cpus_on_node"] = len(cpus)
    if cpus and cpus != allowed:
        os.sched_setaffinity(0, cpus)
        report["bound"] = True
    return report
