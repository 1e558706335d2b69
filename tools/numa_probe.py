#!/usr/bin/env python
"""Host topology of the GPU box and the host->device copy rate from pinned memory placed on each
NUMA node (is the multi-GPU end-to-end figure sensitive to where the pinned buffers live?)."""
import glob, os, subprocess, sys, time
import torch


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return "ERR %r" % (e,)


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


print(sh("nvidia-smi topo -m"))
print(sh("lscpu | grep -i -E 'numa|model name|socket|^cpu\\(s\\)'"))
print("affinity:", len(os.sched_getaffinity(0)), "cpus")
nodes = {}
for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    n = int(d.rsplit("node", 1)[1])
    nodes[n] = cpulist(open(d + "/cpulist").read())
    print("node", n, "cpus", open(d + "/cpulist").read().strip())
for dev in glob.glob("/sys/bus/pci/devices/*"):
    try:
        if open(dev + "/vendor").read().strip() == "0x10de" and open(dev + "/class").read().startswith("0x0302"):
            print(os.path.basename(dev), "numa_node", open(dev + "/numa_node").read().strip())
    except OSError:
        pass
print("cuda:0 pci", torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else sh("nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader"))

full = os.sched_getaffinity(0)
dst = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for n, cpus in nodes.items():
    use = set(cpus) & full
    if not use:
        print("node", n, "not in affinity mask")
        continue
    os.sched_setaffinity(0, use)
    t0 = time.time()
    src = torch.empty(512 << 20, dtype=torch.uint8, pin_memory=True)
    src.fill_(1)
    t_alloc = time.time() - t0
    os.sched_setaffinity(0, full)
    where = sh("grep -c . /proc/%d/numa_maps" % os.getpid())
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("pinned on node %d: H2D %.1f GB/s (alloc+touch %.2f s, numa_maps lines %s)" % (n, (512 << 20) / ms / 1e6, t_alloc, where))
    # where did the pages land?
    ptr = src.data_ptr()
    for line in open("/proc/self/numa_maps"):
        if line.startswith("%x " % ptr) or line.startswith("%012x " % ptr):
            print("   ", line.strip()[:200])
    del src
    torch.cuda.empty_cache()
    try:
        torch._C._host_emptyCache()
    except Exception:  # noqa: BLE001
        pass
