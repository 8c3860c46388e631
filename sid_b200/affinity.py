"""Host placement for one-process-per-GPU runs: pin the rank's threads (and with them the first touch of
its pinned staging buffers) to the CPUs next to its GPU.  On a two-socket box half of the GPUs hang off
each socket; a rank whose pinned text lives on the other socket pays the inter-socket link on every
host<->device copy of the end-to-end path (sidgpu_call_host)."""
import os


def _sysfs(bus_id, name):
    # nvml: 00000000:1B:00.0 -> sysfs: 0000:1b:00.0
    b = bus_id.lower()
    if len(b.split(":")[0]) == 8:
        b = b[4:]
    try:
        with open("/sys/bus/pci/devices/%s/%s" % (b, name)) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s):
    out = set()
    for part in (s or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.update(range(int(a), int(b) + 1))
        else:
            out.add(int(part))
    return out


def describe(index):
    """What the driver and sysfs say about GPU `index`: bus id, NUMA node, neighbouring CPUs."""
    info = {"gpu": index, "bus_id": None, "numa_node": None, "cpus": [], "source": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        info["bus_id"] = bus.decode() if isinstance(bus, bytes) else bus
        try:
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
            if cpus:
                info["cpus"], info["source"] = cpus, "nvml"
        except Exception:
            pass
    except Exception:
        pass
    if info["bus_id"]:
        node = _sysfs(info["bus_id"], "numa_node")
        info["numa_node"] = int(node) if node not in (None, "") else None
        if not info["cpus"]:
            cpus = sorted(_parse_cpulist(_sysfs(info["bus_id"], "local_cpulist")))
            if cpus:
                info["cpus"], info["source"] = cpus, "sysfs"
    return info


def bind_to_gpu(index):
    """Restricts the calling process to the CPUs next to GPU `index` (no-op when the platform reports none or
    all of them, as single-socket hosts and most VMs do).  Returns the description with what was done."""
    info = describe(index)
    allowed = os.sched_getaffinity(0)
    want = set(info["cpus"]) & allowed
    info["bound"] = False
    if want and want != allowed:
        try:
            os.sched_setaffinity(0, want)
            info["bound"] = True
        except OSError:
            pass
    return info
