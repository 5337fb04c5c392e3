#!/usr/bin/env python
"""Does the NUMA node of the pinned host buffer matter for PCIe copies on this box?  (round 1 probe)"""
import os, glob, time, subprocess
import torch

def cpus_of(node):
    s = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
    out = []
    for part in s.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out

print(subprocess.run("nvidia-smi topo -m", shell=True, capture_output=True, text=True).stdout)
nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes", nodes, "affinity now", len(os.sched_getaffinity(0)))
try:
    bus = torch.cuda.get_device_properties(0).pci_bus_id
except Exception:
    bus = None
print("gpu0 numa:", subprocess.run("cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c", shell=True, capture_output=True, text=True).stdout)
dev = torch.device("cuda", 0)
d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
full = os.sched_getaffinity(0)
for node in nodes + [None]:
    if node is not None:
        cp = set(cpus_of(node)) & full
        if not cp:
            continue
        os.sched_setaffinity(0, cp)
    else:
        os.sched_setaffinity(0, full)
    h = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print("node", node, name, "%.1f GB/s" % (h.numel() / dt / 1e9))
    del h
