"""Quick device-resident timing of the bench workload (no e2e pass): ms per inference(10) and per kernel class, for every
kernel-variant mask given on the command line (default 0 and 7).  Prints one line per mask.
    [LD_LIBRARY_PATH=build/variants/<name>] [DCRF_B200_ROW_ORDER=0|1] python profiles/quick_bench.py <tag> [mask ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pydensecrf_b200.densecrf as dcrf
import pydensecrf_b200.eigen as eigen
from tests import workloads as wl
W, H, L = 1920, 1080, 21
cache = "/tmp/dcrf_bench_inputs.npz"
if os.path.exists(cache):
    z = np.load(cache); U, img = z["U"], z["img"]
else:
    U, img = wl.softmax_unary(H, W, L, seed=0), wl.smooth_image(H, W, seed=0)
    np.savez(cache, U=U, img=img)
tag = sys.argv[1] if len(sys.argv) > 1 else ""
masks = [int(a) for a in sys.argv[2:]] or [0, 7]
d = dcrf.DenseCRF2D(W, H, L)
d.setUnaryEnergy(U)
d.addPairwiseGaussian(sxy=3, compat=3)
d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
Q = eigen.MatrixXf()
ref = None
for mask in masks:
    d._set_kernel_variant(mask)
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
    q = np.asarray(Q).copy()
    if ref is None:
        ref = q
    same = bool(np.array_equal(q, ref))
    if not same:
        print("    max |dQ| against the first mask: %.3g" % float(np.abs(q - ref).max()))
    import hashlib
    print("%-28s q=%s mask=%d  %.3f ms  same_bits=%s  %s" % (
        tag, hashlib.sha1(q.tobytes()).hexdigest()[:8], mask, ms, same,
        " ".join("%s=%.4f" % (k, v[1] / v[0]) for k, v in sorted(p.items()))), flush=True)
