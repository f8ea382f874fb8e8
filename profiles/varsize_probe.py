"""A dataset of varying image sizes (PASCAL VOC style: ~500 x 375, every image a different size), the reference's usage
pattern per image; against the same number of images of one fixed size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from tests import workloads as wl
L = 21
rng = np.random.default_rng(0)
def run(sizes, keep):
    kept = []
    t0 = time.perf_counter()
    for i, (W, H) in enumerate(sizes):
        d = dcrf.DenseCRF2D(W, H, L)
        d.setUnaryEnergy(data[(W, H)][0])
        d.addPairwiseGaussian(sxy=3, compat=3)
        d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=data[(W, H)][1], compat=10)
        m = np.argmax(d.inference(5), axis=0)
        if keep: kept.append(m)
    return (time.perf_counter() - t0) / len(sizes) * 1e3
var = [(int(rng.integers(300, 501)), int(rng.integers(250, 501))) for _ in range(60)]
fixed = [(500, 375)] * 60
data = {s: (wl.softmax_unary(s[1], s[0], L, seed=1, grid=8), wl.smooth_image(s[1], s[0], seed=1, grid=8)) for s in set(var + fixed)}
run(fixed[:5], False)
for name, sizes in (("fixed 500x375", fixed), ("60 different sizes", var), ("fixed 500x375", fixed), ("60 different sizes", var)):
    print("%-20s %.2f ms per image (results dropped), %.2f (kept)" % (name, run(sizes, False), run(sizes, True)), flush=True)
