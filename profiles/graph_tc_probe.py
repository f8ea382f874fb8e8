import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydensecrf_b200.densecrf as dcrf
import pydensecrf_b200.eigen as eigen
from tests import workloads as wl
W,H,L=256,160,150
d=dcrf.DenseCRF2D(W,H,L); d.setMatrixCompatMode(2)
d.setUnaryEnergy(wl.softmax_unary(H,W,L,seed=1,grid=8))
d.addPairwiseGaussian(3, wl.matrix_compat(L)); d.addPairwiseBilateral(30,13,wl.smooth_image(H,W,seed=2,grid=8), wl.matrix_compat(L,seed=1))
Q=eigen.MatrixXf(); ref=None
for i in range(7):
    d.inferenceInto(3,Q); q=np.asarray(Q).copy()
    if ref is None: ref=q
    print(i, bool(np.array_equal(q,ref)), float(q.sum()))
