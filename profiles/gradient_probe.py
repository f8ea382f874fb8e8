"""DenseCRF::gradient (learning path) at a training-image size: the engine against the unmodified reference
C++ (oracle/_ref, one core) on the same model -- 2-D Gaussian + bilateral terms, DIAG kernels, symmetric normalisation,
Potts compatibilities, LogLikelihood objective, all three gradients."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from oracle import pyoracle as po
from tests import workloads as wl

W, H, L, IT = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (320, 240, 21, 5)))
img = wl.smooth_image(H, W, seed=1, grid=16)
U = wl.softmax_unary(H, W, L, seed=1, grid=8)
gt = np.argmax(-U, axis=0).astype(np.int16)
gt[::7] = -1

m = dcrf.DenseCRF2D(W, H, L)
m.setUnaryEnergy(U)
m.addPairwiseGaussian(sxy=3, compat=3)
m.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
obj = dcrf.LogLikelihood(gt)
m.gradient(IT, obj)
t0 = time.perf_counter()
R = 5
for _ in range(R):
    got = m.gradient(IT, obj)
t_gpu = (time.perf_counter() - t0) / R
m._profile_enable(True)
m.gradient(IT, obj)
prof = sorted(((k, v[0], v[1]) for k, v in m._profile_read().items()), key=lambda e: -e[2])
m._profile_enable(False)
print("per kernel class (name, launches, ms):", ", ".join("%s %d %.3f" % e for e in prof), flush=True)
print("engine: gradient(%d) %.2f ms at %dx%dx%d" % (IT, t_gpu * 1e3, W, H, L), flush=True)

if os.environ.get("PROBE_NO_REF"):
    sys.exit(0)
r = po.CRF((W, H), L, "ref")
r.setUnaryEnergy(U)
r.addPairwiseGaussian(3, 3)
r.addPairwiseBilateral(80, 13, img, 10)
t0 = time.perf_counter()
want = r.gradient(IT, 0, gt)
t_ref = time.perf_counter() - t0
print("reference (1 core): %.2f s -> %.0fx" % (t_ref, t_ref / t_gpu))
print("value %.9g vs %.9g" % (got[0], want[0]))
for name, a, b in zip(("unary", "lbl_cmp", "kernel"), got[1:], want[1:]):
    if b.size:
        print("%-8s max|d| / max|ref| = %.2e   ref %s" % (name, np.abs(a - b).max() / np.abs(b).max(), np.array2string(b, precision=5)))
