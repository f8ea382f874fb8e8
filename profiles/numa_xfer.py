"""PCIe rates of page-locked uploads/downloads with and without binding the process to the GPU's NUMA node."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pynvml
import pydensecrf_b200.densecrf as dcrf
from pydensecrf_b200.utils import pinned_empty

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
bus = pynvml.nvmlDeviceGetPciInfo(h).busId
bus = bus.decode() if isinstance(bus, bytes) else bus
sysdir = "/sys/bus/pci/devices/" + bus[-12:].lower()
def rd(p):
    try: return open(os.path.join(sysdir, p)).read().strip()
    except Exception as e: return "n/a (%s)" % type(e).__name__
print("gpu0 pci", bus, "numa_node", rd("numa_node"), "local_cpulist", rd("local_cpulist"))
print("process affinity: %d cpus; nodes: %s" % (len(os.sched_getaffinity(0)), [n for n in os.listdir("/sys/devices/system/node") if n.startswith("node")]))
try:
    print("current cpu", os.sched_getcpu() if hasattr(os, "sched_getcpu") else "?")
except Exception: pass

W, H, L = 1920, 1080, 21
d = dcrf.DenseCRF2D(W, H, L)
def measure(tag):
    U = pinned_empty((L, W * H), np.float32); U[...] = 1.0
    for _ in range(2): d.setUnaryEnergy(U); d.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): d.setUnaryEnergy(U); d.synchronize()
    h2d = (time.perf_counter() - t0) / 5
    Q = d.inference(0); q = np.asarray(Q); del q, Q
    Q = d.inference(0); q = np.asarray(Q); del q, Q
    t0 = time.perf_counter()
    for _ in range(5):
        Q = d.inference(0); q = np.asarray(Q); del q, Q
    d2h = (time.perf_counter() - t0) / 5
    print("%s: H2D %.2f ms (%.1f GB/s), inference(0)+D2H %.2f ms (%.1f GB/s)" % (tag, h2d * 1e3, U.nbytes / h2d / 1e9, d2h * 1e3, U.nbytes / d2h / 1e9))
measure("default placement")
for node in sorted(n for n in os.listdir("/sys/devices/system/node") if n.startswith("node")):
    cl = open("/sys/devices/system/node/%s/cpulist" % node).read().strip()
    cpus = set()
    for part in cl.split(","):
        if not part: continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    cpus &= os.sched_getaffinity(0) if False else cpus
    try:
        os.sched_setaffinity(0, cpus)
    except Exception as e:
        print(node, "cannot bind:", e); continue
    import pydensecrf_b200._core as core
    measure("bound to %s (cpus %s)" % (node, cl))
