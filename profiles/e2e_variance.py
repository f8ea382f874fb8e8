"""Spread of the serial end-to-end step (setUnaryEnergy(pinned) + np.asarray(inference(10))) over repeated loops of 20 steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from pydensecrf_b200.utils import pinned_empty
from tests import workloads as wl
W, H, L = 1920, 1080, 21
U = wl.softmax_unary(H, W, L, seed=0)
d = dcrf.DenseCRF2D(W, H, L)
d.addPairwiseGaussian(3, 3.0)
d.addPairwiseBilateral(80, 13, wl.smooth_image(H, W, seed=0), 10.0)
Up = pinned_empty(U.shape, np.float32); Up[...] = U
out = None
for _ in range(12):
    d.setUnaryEnergy(Up); out = np.asarray(d.inference(10))
for rep in range(8):
    per = []
    for _ in range(20):
        t0 = time.perf_counter()
        d.setUnaryEnergy(Up); t1 = time.perf_counter()
        Q = d.inference(10); t2 = time.perf_counter()
        out = np.asarray(Q); t3 = time.perf_counter()
        per.append((t1 - t0, t2 - t1, t3 - t2))
    a = np.array(per) * 1e3
    print("loop %d: %.2f ms per step (setUnary %.2f, inference call %.2f, asarray %.2f); worst step %.2f" % (
        rep, a.sum(1).mean(), a[:, 0].mean(), a[:, 1].mean(), a[:, 2].mean(), a.sum(1).max()), flush=True)
