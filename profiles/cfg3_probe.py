"""cfg 3 (4K, 150 labels, two MatrixCompatibility terms) stand-alone: per-launch times of a few iterations.
  [LD_LIBRARY_PATH=build/variants/<name>] python profiles/cfg3_probe.py [tag] [iterations]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydensecrf_b200.densecrf as dcrf
import pydensecrf_b200.eigen as eigen
from tests import workloads as wl
tag = sys.argv[1] if len(sys.argv) > 1 else "-"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W, H, L = 3840, 2160, 150
img, U, Cm = wl.smooth_image(H, W, seed=0), wl.softmax_unary(H, W, L, seed=0), wl.matrix_compat(L)
d = dcrf.DenseCRF2D(W, H, L)
d.setUnaryEnergy(U)
d.addPairwiseGaussian(3, Cm)
d.addPairwiseBilateral(80, 13, img, Cm)
Q = eigen.MatrixXf()
d.inferenceInto(1, Q)
d._profile_enable(True)
d._profile_read(True)
d.inferenceInto(n, Q)
p = d._profile_read(True)
d._profile_enable(False)
print(tag, " ".join("%s=%.3f" % (k, v[1] / max(v[0], 1)) for k, v in sorted(p.items())), "checksum %.6f" % float(np.asarray(Q)[:, ::1001].sum()), flush=True)
