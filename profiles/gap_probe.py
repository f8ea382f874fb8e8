"""How much of the 1080p step is not kernel time?  Models with only the bilateral / only the Gaussian term next to the full
one: device time per inference(10) against the sum of the per-launch times of the same launches (events around every launch)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pydensecrf_b200.densecrf as dcrf
import pydensecrf_b200.eigen as eigen
from tests import workloads as wl
W, H, L = 1920, 1080, 21
U, img = wl.softmax_unary(H, W, L, seed=0), wl.smooth_image(H, W, seed=0)
for name in ("both", "bilateral", "gaussian"):
    d = dcrf.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(U)
    if name != "bilateral":
        d.addPairwiseGaussian(sxy=3, compat=3)
    if name != "gaussian":
        d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
    Q = eigen.MatrixXf()
    for _ in range(4):
        d.inferenceInto(10, Q)
    d._timer_start()
    for _ in range(10):
        d.inferenceInto(10, Q)
    ms = d._timer_stop() / 10
    d._profile_enable(True); d._profile_read(True)
    for _ in range(3):
        d.inferenceInto(10, Q)
    p = d._profile_read(True); d._profile_enable(False)
    tot = sum(v[1] for v in p.values()) / 3
    print("%-10s %.3f ms per inference(10); sum of its launches %.3f ms (%d launches); %s" % (
        name, ms, tot, sum(v[0] for v in p.values()) // 3, " ".join("%s=%.4f" % (k, v[1] / v[0]) for k, v in sorted(p.items()))), flush=True)
