"""Warm lattice-construction time at 1080p (third model in the process: memory pools populated)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from pydensecrf_b200.utils import pinned_empty
from tests import workloads as wl
W, H, L = 1920, 1080, 21
img = wl.smooth_image(H, W, seed=0); U = wl.softmax_unary(H, W, L, seed=0)
Up = pinned_empty(U.shape); Up[...] = U
for rep in range(4):
    t0 = time.perf_counter(); d = dcrf.DenseCRF2D(W, H, L); t1 = time.perf_counter()
    d.setUnaryEnergy(Up); t2 = time.perf_counter()
    d.addPairwiseGaussian(3, 3); d.synchronize(); t3 = time.perf_counter()
    d.addPairwiseBilateral(80, 13, img, 10); d.synchronize(); t4 = time.perf_counter()
    m = d.currentMap(d.inference(10)); t5 = time.perf_counter()
    print("model %d: create %.2f ms, unary upload %.2f, gaussian lattice %.2f, bilateral lattice %.2f, inference(10)+MAP %.2f, total %.2f ms"
          % (rep, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t5-t0)*1e3))
d._profile_enable(True); d._profile_read(True)
d2 = None
# per-kernel-class durations of one more bilateral lattice build (events around every launch)
d.addPairwiseBilateral(80, 13, img, 10); d.synchronize()
p = d._profile_read(True); d._profile_enable(False)
tot = sum(v[1] for v in p.values())
print("bilateral lattice build, GPU time by launch class (ms): total %.3f" % tot)
for k, v in sorted(p.items(), key=lambda kv: -kv[1][1]):
    print("   %-22s x%-3d %.3f" % (k, v[0], v[1]))
