"""Quick device-resident timing of the bench workload (no e2e / profile passes): ms per inference(10)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pydensecrf_b200.densecrf as dcrf
import pydensecrf_b200.eigen as eigen
from tests import workloads as wl
W, H, L = 1920, 1080, 21
d = dcrf.DenseCRF2D(W, H, L)
d.setUnaryEnergy(wl.softmax_unary(H, W, L, seed=0))
d.addPairwiseGaussian(sxy=3, compat=3)
d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=wl.smooth_image(H, W, seed=0), compat=10)
Q = eigen.MatrixXf()
for fast in (False, True):
    d.setSoftmaxMode(fast)
    for _ in range(3):
        d.inferenceInto(10, Q)
    d._timer_start()
    for _ in range(10):
        d.inferenceInto(10, Q)
    ms = d._timer_stop() / 10
    d._profile_enable(True); d._profile_read(True)
    for _ in range(5):
        d.inferenceInto(10, Q)
    p = d._profile_read(True); d._profile_enable(False)
    print(sys.argv[1] if len(sys.argv) > 1 else "", "fast" if fast else "exact", "%.3f ms" % ms,
          {k: round(v[1] / v[0], 4) for k, v in p.items()})
