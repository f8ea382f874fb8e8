"""Where does the time of a fresh model per image go?  Phases of the reference's usage pattern, averaged over images."""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from tests import workloads as wl
W, H, L = (int(sys.argv[1]), int(sys.argv[2]), 21) if len(sys.argv) > 2 else (512, 512, 21)
imgs = [wl.smooth_image(H, W, seed=i) for i in range(4)]
Us = [wl.softmax_unary(H, W, L, seed=i) for i in range(4)]
KEEP = os.environ.get('KEEP_MAPS', '0') == '1'
kept = []
names = ["ctor", "setUnary", "gaussian", "bilateral", "inference(5)", "argmax", "del Q", "del model"]
acc = np.zeros(len(names))
def one(i, rec):
    t = [time.perf_counter()]
    d = dcrf.DenseCRF2D(W, H, L); t.append(time.perf_counter())
    d.setUnaryEnergy(Us[i % 4]); t.append(time.perf_counter())
    d.addPairwiseGaussian(sxy=3, compat=3); t.append(time.perf_counter())
    d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=imgs[i % 4], compat=10); t.append(time.perf_counter())
    Q = d.inference(5); d.synchronize(); t.append(time.perf_counter())
    m = np.argmax(Q, axis=0); t.append(time.perf_counter())
    if KEEP: kept.append(m)
    del Q; t.append(time.perf_counter())
    del d; t.append(time.perf_counter())
    if rec:
        acc[:] += np.diff(t)
for i in range(4): one(i, False)
n = 20
t0 = time.perf_counter()
for i in range(n): one(i, True)
tot = (time.perf_counter() - t0) / n
print("%dx%d: %.2f ms per image; " % (W, H, tot * 1e3) + " ".join("%s=%.2f" % (k, v / n * 1e3) for k, v in zip(names, acc)), "keep_maps=%s" % KEEP, flush=True)
