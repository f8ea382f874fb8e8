"""Device-resident inference(10) at 1080p for several label counts (the reference's examples use 2): ms per call and per-launch
kernel times.  Row layouts: L=1 -> 4-byte rows, L=2 -> 8-byte rows (the reference's seqCompute arithmetic), L>=3 -> float4 chunks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pydensecrf_b200.densecrf as dcrf
import pydensecrf_b200.eigen as eigen
from tests import workloads as wl
W, H = 1920, 1080
img = wl.smooth_image(H, W, seed=0)
for L in [int(a) for a in sys.argv[1:]] or [2, 3, 4, 8, 21, 32, 64]:
    d = dcrf.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(wl.softmax_unary(H, W, L, seed=0))
    d.addPairwiseGaussian(sxy=3, compat=3)
    d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
    Q = eigen.MatrixXf()
    for _ in range(4): d.inferenceInto(10, Q)
    d._timer_start()
    for _ in range(10): d.inferenceInto(10, Q)
    ms = d._timer_stop() / 10
    d._profile_enable(True); d._profile_read(True)
    for _ in range(3): d.inferenceInto(10, Q)
    p = d._profile_read(True); d._profile_enable(False)
    print("L=%3d: %.3f ms per inference(10) = %.0f MPix*iter/s  %s" % (L, ms, W * H * 10 / ms / 1e3, {k: round(v[1] / v[0], 4) for k, v in p.items()}))
    del d, Q
