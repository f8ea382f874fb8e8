"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel family once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from tests import workloads as wl
W, H = 67, 45
img = wl.smooth_image(H, W, seed=1, grid=8)
for L, compat in [(21, 3.0), (2, 3.0), (5, None), (70, "m"), (150, "m")]:
    U = wl.softmax_unary(H, W, L, seed=L, grid=8)
    c = wl.matrix_compat(L) if compat == "m" else (np.linspace(-3, -1, L).astype(np.float32) if compat is None else compat)
    d = dcrf.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(U)
    d.addPairwiseGaussian(3, c)
    d.addPairwiseBilateral(30, 13, img, c)
    d.addPairwiseBilateralND((30, 30), (13, 13, 13), img, compat=c, chdim=2)
    Q = d.inference(2)
    Q2 = d.inference(2); Q2 = d.inference(2); Q2 = d.inference(2)   # third call captures, fourth replays the CUDA graph
    kl = d.klDivergence(Q)
    m = d.currentMap(Q)
    Qs, a, b = d.startInference(); d.stepInference(Qs, a, b)
    d.setSoftmaxMode(True); d.inference(1); d.setSoftmaxMode(False)
    d._set_iteration_mode(1); d.inference(1)
    print(L, float(np.asarray(Q).sum()), kl, int(m.sum()))
# batch entry (worker threads, shared Gaussian lattice), NumPy argmax dispatch (int64 device argmax), staged upload (> 4 MB)
Ub = np.stack([wl.softmax_unary(H, W, 7, seed=b, grid=8) for b in range(5)])
Ib = np.stack([wl.smooth_image(H, W, seed=b, grid=8) for b in range(5)])
lab = dcrf.map_batch_2d(Ub, Ib, n_iter=2, workers=3)
d = dcrf.DenseCRF2D(256, 256, 21)
d.setUnaryEnergy(wl.softmax_unary(256, 256, 21, seed=1, grid=8))
d.addPairwiseGaussian(3, 3.0)
m2 = np.argmax(d.inference(1), axis=0)
assert d._selftest(1, n=100000, seed=3) == 0
# round 2: super-lattice chunks (pixel count divisible by four), exact-order KL, learning-path kernels, the TMA-fed
# tensor-core update with several tiles per CTA and both accumulator-buffer counts
Ub = np.stack([wl.softmax_unary(48, 64, 21, seed=b, grid=8) for b in range(6)])
Ib = np.stack([wl.smooth_image(48, 64, seed=b, grid=8) for b in range(6)])
os.environ["DCRF_B200_BATCH_CHUNK"] = "2"
lab2 = dcrf.map_batch_2d(Ub, Ib, n_iter=2, workers=3)
d.setKLMode(1); kl1 = d.klDivergence(d.inference(1)); d.setKLMode(0)
g = dcrf.DenseCRF(500, 5)
f = (np.random.default_rng(0).standard_normal((4, 500)) * 2).astype(np.float32)
g.addPairwiseEnergy(f, 1.0)
a_ = np.random.default_rng(1).random((500, 5)).astype(np.float32)
df = g._lattice_gradient(0, a_, a_[::-1].copy())
import pydensecrf_b200.eigen as eigen
tr = np.array(g._pairwise_apply(0, eigen.matrixXf(np.ascontiguousarray(a_.T)), transpose=True))
for L, (Wt, Ht) in [(150, (256, 160)), (180, (64, 48)), (20, (64, 48))]:
    t = dcrf.DenseCRF2D(Wt, Ht, L)
    t.setMatrixCompatMode(2)
    t.setUnaryEnergy(wl.softmax_unary(Ht, Wt, L, seed=L, grid=8))
    t.addPairwiseGaussian(3, wl.matrix_compat(L))
    t.addPairwiseBilateral(30, 13, wl.smooth_image(Ht, Wt, seed=2, grid=8), wl.matrix_compat(L, seed=1))
    print(L, float(np.asarray(t.inference(2)).sum()))
print(int(lab2.sum()), kl1, float(df.sum()), float(tr.sum()))
print(int(lab.sum()), int(m2.sum()))
print("sanitize-ok")
