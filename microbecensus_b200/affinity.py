"""Bind a rank's host threads (and, by first touch, its pinned buffers) to the NUMA node its GPU hangs off.
With one process per GPU the host->device copies of eight ranks otherwise share whichever socket the scheduler picked."""
import os


def _cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Returns the NUMA node bound to, or None when the topology cannot be read (no sysfs entry, single node, ...)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:            # nvml pads the PCI domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = _cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
