"""cfg 4 throughput on ONE GPU with several host threads, each driving its own DenseCRF2D objects (every model
owns a CUDA stream and the binding releases the GIL inside the C-ABI calls, so uploads, lattice builds and
inference of different images overlap).  Prints images/s for 1, 2, 4, 8 threads; inputs in page-locked memory."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from pydensecrf_b200.utils import pinned_empty
from tests import workloads as wl

W = H = 512; L = 21
imgs = [wl.smooth_image(H, W, seed=i) for i in range(4)]
Us = []
for i in range(4):
    u = pinned_empty((L, W * H)); u[...] = wl.softmax_unary(H, W, L, seed=i); Us.append(u)

def one(i):
    d = dcrf.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(Us[i % 4])
    d.addPairwiseGaussian(3, 3)
    d.addPairwiseBilateral(80, 13, imgs[i % 4], 10)
    return d.currentMap(d.inference(5))

for _ in range(8): one(0)
ref = one(1)
for nt in (1, 2, 4, 8):
    n_img = 64 * nt
    def work(t):
        for i in range(t, n_img, nt):
            m = one(i)
            if i % 4 == 1: assert np.array_equal(m, ref)
    ths = [threading.Thread(target=work, args=(t,)) for t in range(nt)]
    t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]; dt = time.perf_counter() - t0
    print("cfg4 one GPU, %d host thread(s): %.0f images/s (%.2f ms per image, %.0f MPix*iter/s incl. lattice builds and transfers)"
          % (nt, n_img / dt, dt / n_img * 1e3, n_img * W * H * 5 / dt / 1e6))

# the same through the library's batch entry (worker threads inside the C++ core, one call from Python)
Bn = 256
Ub = pinned_empty((Bn, L, W * H)); Ib = np.empty((Bn, H, W, 3), np.uint8)
for b in range(Bn):
    Ub[b] = Us[b % 4]; Ib[b] = imgs[b % 4]
for nw in (1, 4, 8):
    dcrf.map_batch_2d(Ub[:16], Ib[:16], n_iter=5, workers=nw)
    t0 = time.perf_counter(); lab = dcrf.map_batch_2d(Ub, Ib, n_iter=5, workers=nw); dt = time.perf_counter() - t0
    assert np.array_equal(lab[1].ravel(), ref)
    print("cfg4 one GPU, map_batch_2d(workers=%d): %.0f images/s (%.2f ms per image)" % (nw, Bn / dt, dt / Bn * 1e3))

if dcrf.device_count() > 1:
    devs = list(range(dcrf.device_count()))
    dcrf.map_batch_2d(Ub[:32], Ib[:32], n_iter=5, device=devs)
    t0 = time.perf_counter(); lab = dcrf.map_batch_2d(Ub, Ib, n_iter=5, device=devs); dt = time.perf_counter() - t0
    assert np.array_equal(lab[1].ravel(), ref)
    print("cfg4 %d GPUs, map_batch_2d(device=%s): %.0f images/s (%.2f ms per image)" % (len(devs), devs, Bn / dt, dt / Bn * 1e3))
