"""CPU / memory placement for the host-buffer path on multi-socket boxes.

Pinned host buffers are allocated on the NUMA node of the thread that creates them.  A rank whose buffers live on
the other socket pays the inter-socket link on every upload and download, and eight ranks doing that at once share
it.  `bind_to_device_numa(i)` restricts the calling process to the CPUs that are local to CUDA device i (read from
sysfs: /sys/bus/pci/devices/<bus id>/local_cpulist), so that everything it allocates afterwards is local too.
Call it once per process, before allocating pinned memory.  Best effort: it never raises."""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def device_local_cpus(device_index):
    """CPUs local to CUDA device `device_index`, or None when the platform does not say."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus_id = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus_id}/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
        return cpus or None
    except Exception:
        return None


def bind_to_device_numa(device_index):
    """Restrict this process to the device-local CPUs it is already allowed to use.  Returns a small report."""
    report = {"bound": False}
    try:
        allowed = os.sched_getaffinity(0)
        local = device_local_cpus(device_index)
        report["allowed_cpus"] = len(allowed)
        if not local:
            report["why"] = "no local_cpulist for the device"
            return report
        target = allowed & local
        report["local_cpus"] = len(local)
        if not target or target == allowed:
            report["why"] = "already local" if target else "no allowed CPU is local to the device"
            return report
        os.sched_setaffinity(0, target)
        report.update(bound=True, cpus=len(target))
    except Exception as e:  # placement is an optimisation, never a failure
        report["why"] = repr(e)
    return report
