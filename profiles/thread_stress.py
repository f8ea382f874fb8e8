"""Thread-safety stress: several host threads, each creating models, uploading, inferring, downloading and destroying
them concurrently on one GPU.  Run with MALLOC_CHECK_=3 to trap host heap corruption early.  Prints OK."""
import os, sys, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from pydensecrf_b200.utils import pinned_empty
from tests import workloads as wl

W, H, L = 640, 480, 21
img = wl.smooth_image(H, W, seed=0)
U = wl.softmax_unary(H, W, L, seed=0)
n_thr, rounds, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 6, 6
ref = None
out = [None] * n_thr
errs = []

def worker(k):
    try:
        for r in range(rounds):
            d = dcrf.DenseCRF2D(W, H, L)
            d.addPairwiseGaussian(sxy=3, compat=3)
            d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
            uh = pinned_empty(U.shape, np.float32)
            uh[...] = U
            for s in range(steps):
                d.setUnaryEnergy(uh)
                q = np.array(d.inference(5))
            out[k] = q
            del d, uh
    except Exception as e:  # noqa: BLE001
        errs.append(repr(e))

ths = [threading.Thread(target=worker, args=(k,)) for k in range(n_thr)]
for t in ths: t.start()
for t in ths: t.join()
assert not errs, errs
for k in range(1, n_thr):
    assert np.array_equal(out[0], out[k])
print("thread stress OK: %d threads x %d models x %d steps, identical results" % (n_thr, rounds, steps))
