"""The reference's example flow (examples/inference.py:60-122: unary_from_labels -> setUnaryEnergy -> pairwise terms ->
inference(5) -> argmax) per 1080p image with a fresh model each time: the numpy unary + its upload against
setUnaryFromLabels (the label map crosses the bus, the energies are formed on the device)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from pydensecrf_b200 import utils
from tests import workloads as wl

W, H, L = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (1920, 1080, 21)))
img = wl.smooth_image(H, W, seed=0)
labels = (np.argmax(-wl.softmax_unary(H, W, L, seed=0), axis=0) + 1).reshape(H, W).astype(np.int32)
labels[::9, ::7] = 0


def run(device_side, reps):
    parts = np.zeros(3)
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        d = dcrf.DenseCRF2D(W, H, L)
        if device_side:
            d.setUnaryFromLabels(labels, 0.7, zero_unsure=True)
        else:
            d.setUnaryEnergy(utils.unary_from_labels(labels, L, 0.7, zero_unsure=True))
        t1 = time.perf_counter()
        d.addPairwiseGaussian(sxy=3, compat=3)
        d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
        t2 = time.perf_counter()
        out = np.argmax(d.inference(5), axis=0)
        t3 = time.perf_counter()
        parts += (t1 - t0, t2 - t1, t3 - t2)
    return parts / reps * 1e3, out


run(False, 2), run(True, 2)
for name, dev in (("utils.unary_from_labels + setUnaryEnergy", False), ("setUnaryFromLabels", True)) * 2:
    p, m = run(dev, 5)
    print("%-42s unary %.2f ms, pairwise terms %.2f ms, inference(5) + argmax %.2f ms, total %.2f ms" % (name, *p, p.sum()), flush=True)
a, b = run(False, 1)[1], run(True, 1)[1]
print("label maps identical:", bool(np.array_equal(a, b)))
