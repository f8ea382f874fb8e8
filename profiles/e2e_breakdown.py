"""Where the end-to-end step of the bench workload spends its time (host clock, one GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from pydensecrf_b200.utils import pinned_empty
from tests import workloads as wl

W, H, L = 1920, 1080, 21
d = dcrf.DenseCRF2D(W, H, L)
U = wl.softmax_unary(H, W, L, seed=0)
Up = pinned_empty(U.shape, np.float32); Up[...] = U
d.addPairwiseGaussian(sxy=3, compat=3)
d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=wl.smooth_image(H, W, seed=0), compat=10)
for _ in range(3):
    d.setUnaryEnergy(Up); q = np.asarray(d.inference(10))
T = np.zeros(4)
n = 10
for _ in range(n):
    t0 = time.perf_counter(); d.setUnaryEnergy(Up); d.synchronize()
    t1 = time.perf_counter(); Q = d.inference(10); d.synchronize()
    t2 = time.perf_counter(); q = np.asarray(Q)
    t3 = time.perf_counter(); del q, Q
    t4 = time.perf_counter()
    T += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
print("per step (ms): setUnaryEnergy+sync %.2f, inference+sync %.2f, np.asarray(Q) %.2f, release %.2f" % tuple(T / n * 1e3))
t0 = time.perf_counter()
for _ in range(n):
    d.setUnaryEnergy(Up); q = np.asarray(d.inference(10))
print("unsplit loop: %.2f ms per step" % ((time.perf_counter() - t0) / n * 1e3))
