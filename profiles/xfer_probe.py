"""Host<->device transfer rates through the public API (diagnostic for the e2e number)."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from pydensecrf_b200.utils import pinned_empty
from tests import workloads as wl
W, H, L = 1920, 1080, 21
U = wl.softmax_unary(H, W, L, seed=0); img = wl.smooth_image(H, W, seed=0)
Up = pinned_empty(U.shape); Up[...] = U
d = dcrf.DenseCRF2D(W, H, L); d.setUnaryEnergy(Up)
d.addPairwiseGaussian(3, 3); d.addPairwiseBilateral(80, 13, img, 10)
def t(f, n=10):
    f(); f(); d.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    d.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("setUnaryEnergy pinned   %.2f ms (%.1f GB/s)" % ((a := t(lambda: d.setUnaryEnergy(Up))), U.nbytes / a / 1e6))
print("setUnaryEnergy pageable %.2f ms (%.1f GB/s)" % ((a := t(lambda: d.setUnaryEnergy(U))), U.nbytes / a / 1e6))
Q = d.inference(1)
def dl():
    np.asarray(Q)          # D2H (device copy is newer after the step below)
    d.stepInference(Q, Q0, Q0) if False else None
import pydensecrf_b200.eigen as eigen
def infer0_and_fetch():
    return np.asarray(d.inference(0))
a = t(infer0_and_fetch); b = t(lambda: d.inference(0))
print("inference(0)+fetch Q    %.2f ms, inference(0) alone %.2f ms -> D2H %.2f ms (%.1f GB/s)" % (a, b, a - b, U.nbytes / (a - b) / 1e6))
a = t(lambda: d.currentMap(d.inference(0)))
print("inference(0)+currentMap %.2f ms" % a)
# two models in two threads: transfers of one overlap the compute of the other (GIL is released in the C calls)
def worker(dev_model, n, out):
    for _ in range(n):
        dev_model.setUnaryEnergy(Up)
        out.append(np.asarray(dev_model.inference(10))[0, 0])
d2 = dcrf.DenseCRF2D(W, H, L); d2.setUnaryEnergy(Up); d2.addPairwiseGaussian(3, 3); d2.addPairwiseBilateral(80, 13, img, 10)
for m in (d, d2): worker(m, 2, [])
n = 10
t0 = time.perf_counter(); worker(d, n, []); t1 = (time.perf_counter() - t0) / n * 1e3
ths = [threading.Thread(target=worker, args=(m, n, [])) for m in (d, d2)]
t0 = time.perf_counter(); [x.start() for x in ths]; [x.join() for x in ths]; t2 = (time.perf_counter() - t0) / (2 * n) * 1e3
print("e2e per step: 1 thread %.2f ms, 2 threads/2 models %.2f ms" % (t1, t2))
