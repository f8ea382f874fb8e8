"""cfg 5 (256^3 volume, 2 labels, 4-D bilateral features) stand-alone: per-launch times."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydensecrf_b200.densecrf as dcrf
import pydensecrf_b200.eigen as eigen
from tests import workloads as wl
D, L, n = 256, 2, 5
vol, U = wl.smooth_volume(D, seed=0), wl.volume_unary(D, seed=0)
d = dcrf.DenseCRF(D ** 3, L)
d.setUnaryEnergy(U)
d.addPairwiseBilateralND((3, 3, 3), (13,), vol, compat=4, chdim=-1, kernel=dcrf.FULL_KERNEL, normalization=dcrf.NORMALIZE_SYMMETRIC)
Q = eigen.MatrixXf()
for _ in range(3):
    d.inferenceInto(n, Q)
d._timer_start()
for _ in range(5):
    d.inferenceInto(n, Q)
ms = d._timer_stop() / 5
d._profile_enable(True); d._profile_read(True)
d.inferenceInto(n, Q)
p = d._profile_read(True); d._profile_enable(False)
print("cfg5 %.3f ms per inference(5), lattice %s; per launch: %s" % (ms, d._lattice_info(0), " ".join("%s=%.4f(x%d)" % (k, v[1] / v[0], v[0]) for k, v in sorted(p.items()))), flush=True)
