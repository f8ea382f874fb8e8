"""The reference's own usage pattern (examples/inference.py:69-122) with plain numpy inputs, timed end to end per image:
DenseCRF2D + setUnaryEnergy + addPairwiseGaussian + addPairwiseBilateral + inference(5) + np.argmax(Q, axis=0).
Runs the same script against the reference C++ (oracle/_ref, one host core) for a handful of images."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from oracle import pyoracle as po
from tests import workloads as wl

for (W, H, L, n_img) in ((512, 512, 21, 40), (1920, 1080, 21, 12)):
    imgs = [wl.smooth_image(H, W, seed=i) for i in range(4)]
    Us = [wl.softmax_unary(H, W, L, seed=i) for i in range(4)]

    def ours(i):
        d = dcrf.DenseCRF2D(W, H, L)
        d.setUnaryEnergy(Us[i % 4])
        d.addPairwiseGaussian(sxy=3, compat=3)
        d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=imgs[i % 4], compat=10)
        Q = d.inference(5)
        return np.argmax(Q, axis=0)

    for i in range(4): ours(i)
    t0 = time.perf_counter()
    maps = [ours(i) for i in range(n_img)]
    t_ours = (time.perf_counter() - t0) / n_img
    line = "%dx%d, %d labels, 5 iterations, plain numpy inputs: %.2f ms per image (%.0f images/s)" % (W, H, L, t_ours * 1e3, 1 / t_ours)
    if po.available("ref"):
        t0 = time.perf_counter()
        o = po.CRF((W, H), L, "ref")
        o.setUnaryEnergy(Us[1]); o.addPairwiseGaussian(3, 3); o.addPairwiseBilateral(80, 13, imgs[1], 10)
        Qo = o.inference(5); mo = np.argmax(Qo, axis=0)
        t_ref = time.perf_counter() - t0
        srt = np.sort(Qo, axis=0); sure = (srt[-1] - srt[-2]) > 1e-3
        line += "; reference C++ on one core: %.0f ms (%.0fx), label maps agree: %s" % (t_ref * 1e3, t_ref / t_ours, np.array_equal(maps[1][sure], mo[sure]))
    print(line)
