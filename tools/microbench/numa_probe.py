"""Host->device bandwidth of pinned memory placed on each NUMA node of the box (tuning aid for the e2e leg: the 21 GB
upload of C2 ran at 20-48 GB/s depending on the box / run)."""
import ctypes
import glob
import os
import time

import torch

libc = ctypes.CDLL("libc.so.6", use_errno=True)
SYS_set_mempolicy = 238   # x86_64
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2


def set_mempolicy(mode, node=None):
    if node is None:
        return libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(mask), 64)


torch.cuda.init()
prop = torch.cuda.get_device_properties(0)
bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
print("gpu pci", bus)
for f in ("numa_node", "local_cpulist"):
    try:
        print(f, open(f"/sys/bus/pci/devices/{bus}/{f}").read().strip())
    except Exception as e:
        print(f, "unreadable", e)
print("nodes online", open("/sys/devices/system/node/online").read().strip())
for nd in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    print(os.path.basename(nd), "cpus", open(nd + "/cpulist").read().strip(),
          [l.strip() for l in open(nd + "/meminfo") if "MemFree" in l or "MemTotal" in l])
print("affinity", sorted(os.sched_getaffinity(0)))
nodes = [int(os.path.basename(p)[4:]) for p in glob.glob("/sys/devices/system/node/node[0-9]*")]
dst = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
for node in [None] + sorted(nodes):
    rc = set_mempolicy(MPOL_BIND, node)
    t0 = time.perf_counter()
    h = torch.empty(2 << 30, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    t_alloc = time.perf_counter() - t0
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(4):
        t0 = time.perf_counter()
        dst.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, (2 << 30) / (time.perf_counter() - t0) / 1e9)
    print("policy node", node, "rc", rc, "alloc+fill s %.2f" % t_alloc, "H2D GB/s %.1f" % best)
    del h
set_mempolicy(MPOL_DEFAULT)
