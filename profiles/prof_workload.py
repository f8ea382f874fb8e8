"""Workload for ncu captures: BASELINE configs[1] (1080p, 21 labels, Gaussian+bilateral), a few
inference(10) calls.  Usage (on the GPU box, one GPU):
  ncu --set full --clock-control none --import-source on -k regex:'k_fused_update|k_splat|k_blur' -s 130 -c 13 \
      -o gpurun_out/prof_r01 python profiles/prof_workload.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pydensecrf_b200.densecrf as dcrf  # noqa: E402
import pydensecrf_b200.eigen as eigen  # noqa: E402
from tests import workloads as wl  # noqa: E402

W, H, L = 1920, 1080, 21
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
d = dcrf.DenseCRF2D(W, H, L)
d.setUnaryEnergy(wl.softmax_unary(H, W, L, seed=0))
d.addPairwiseGaussian(sxy=3, compat=3)
d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=wl.smooth_image(H, W, seed=0), compat=10)
Q = eigen.MatrixXf()
for _ in range(steps):
    d.inferenceInto(10, Q)
d.synchronize()
print("lattices", [d._lattice_info(k) for k in range(2)])
