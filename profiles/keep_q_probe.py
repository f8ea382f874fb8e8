"""np.asarray(inference(n)) with every result kept alive (1080p, 21 labels): ms per call, first calls page-locked, later ones ordinary memory."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from tests import workloads as wl
W, H, L = 1920, 1080, 21
d = dcrf.DenseCRF2D(W, H, L)
d.setUnaryEnergy(wl.softmax_unary(H, W, L, seed=0))
d.addPairwiseGaussian(3, 3.0)
d.addPairwiseBilateral(80, 13, wl.smooth_image(H, W, seed=0), 10.0)
kept, ts = [], []
for i in range(8):
    t0 = time.perf_counter()
    q = np.asarray(d.inference(5))
    ts.append((time.perf_counter() - t0) * 1e3)
    kept.append(q)
print("ms per inference(5) + np.asarray, all results kept:", " ".join("%.1f" % t for t in ts), "equal:", all(np.array_equal(k, kept[0]) for k in kept))
