"""Timings of the other BASELINE configs (cfg 3, 4, 5, and cfg 2 on the i.i.d.-noise worst-case image) at full size on
one B200; prints one line each."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pydensecrf_b200.densecrf as dcrf
import pydensecrf_b200.eigen as eigen
from pydensecrf_b200 import utils
from tests import workloads as wl

which = sys.argv[1:] or ["cfg2noise", "cfg4", "cfg5", "cfg3"]

def tm(f):
    t = time.perf_counter(); r = f(); return r, (time.perf_counter() - t) * 1e3

if "cfg2noise" in which:
    # the i.i.d.-uniform image is the worst case for the bilateral lattice (every pixel its own colour cell)
    W, H, L = 1920, 1080, 21
    d = dcrf.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(wl.softmax_unary(H, W, L, seed=0))
    _, ta = tm(lambda: (d.addPairwiseGaussian(3, 3), d.addPairwiseBilateral(80, 13, wl.noise_image(H, W, seed=0), 10), d.synchronize()))
    Q = eigen.MatrixXf()
    for _ in range(3): d.inferenceInto(10, Q)
    d._timer_start()
    for _ in range(5): d.inferenceInto(10, Q)
    ms = d._timer_stop() / 5
    lat = [d._lattice_info(k)[::2] for k in range(2)]
    b = wl.stage_bytes_per_iter(W * H, L, lat) * 10
    print("cfg2 on the noise image (worst case): lattices %.0f ms (V=%s), inference(10) %.2f ms -> %.0f MPix*iter/s; stage model %.2f GB -> %.2f TB/s"
          % (ta, [v for _, v in lat], ms, W * H * 10 / ms / 1e3, b / 1e9, b / ms / 1e9))
    del d, Q

if "cfg4" in which:
    W = H = 512; L = 21; n_img = 24
    imgs = [wl.smooth_image(H, W, seed=i) for i in range(4)]
    Us = [wl.softmax_unary(H, W, L, seed=i) for i in range(4)]
    per = []
    for i in range(n_img):
        t0 = time.perf_counter()
        d = dcrf.DenseCRF2D(W, H, L)
        d.setUnaryEnergy(Us[i % 4]); t1 = time.perf_counter()
        d.addPairwiseGaussian(3, 3); d.addPairwiseBilateral(80, 13, imgs[i % 4], 10); d.synchronize(); t2 = time.perf_counter()
        Q = d.inference(5); d.synchronize(); t3 = time.perf_counter()
        m = d.currentMap(Q); t4 = time.perf_counter()
        per.append([(t1 - t0), (t2 - t1), (t3 - t2), (t4 - t3), (t4 - t0)])
    p = np.array(per[4:]).mean(0) * 1e3
    print("cfg4 512x512 L=21 5 iters per image: create+unary %.2f ms, 2 lattices %.2f ms, inference %.2f ms, MAP D2H %.2f ms, total %.2f ms -> %.1f img/s, %.1f MPix*iter/s incl. init"
          % (p[0], p[1], p[2], p[3], p[4], 1e3 / p[4], W * H * 5 / p[4] / 1e3))

if "cfg5" in which:
    D = 256; L = 2
    vol = wl.smooth_volume(D, seed=0); U = wl.volume_unary(D, seed=0)
    f, tf = tm(lambda: utils.create_pairwise_bilateral(sdims=(3, 3, 3), schan=(13,), img=vol, chdim=-1))
    d = dcrf.DenseCRF(D ** 3, L)
    _, tu = tm(lambda: d.setUnaryEnergy(U))
    _, ta = tm(lambda: (d.addPairwiseEnergy(f, compat=4, kernel=dcrf.FULL_KERNEL, normalization=dcrf.NORMALIZE_SYMMETRIC), d.synchronize()))
    Q = eigen.MatrixXf()
    d.inferenceInto(5, Q); d.synchronize()
    d._timer_start()
    for _ in range(5): d.inferenceInto(5, Q)
    ms = d._timer_stop() / 5
    print("cfg5 256^3 L=2 4-D bilateral FULL_KERNEL 5 iters: numpy features %.0f ms, unary upload %.0f ms, lattice %.0f ms (V=%d), inference(5) %.2f ms -> %.0f MVox*iter/s"
          % (tf, tu, ta, d._lattice_info(0)[2], ms, D ** 3 * 5 / ms / 1e3))
    d2 = dcrf.DenseCRF(D ** 3, L)
    d2.setUnaryEnergy(U)
    _, tb = tm(lambda: (d2.addPairwiseBilateralND((3, 3, 3), (13,), vol, compat=4, chdim=-1, kernel=dcrf.FULL_KERNEL), d2.synchronize()))
    same = np.array_equal(np.array(d2.inference(5)), np.array(d.inference(5)))
    print("cfg5 with addPairwiseBilateralND (features built on the device, 16.8 MB uploaded instead of 268 MB): lattice %.0f ms, Q identical: %s" % (tb, same))
    del d, Q

if "cfg3" in which:
    W, H, L = 3840, 2160, 150
    img = wl.smooth_image(H, W, seed=0); U = wl.softmax_unary(H, W, L, seed=0); Cm = wl.matrix_compat(L)
    d = dcrf.DenseCRF2D(W, H, L)
    _, tu = tm(lambda: d.setUnaryEnergy(U))
    _, ta = tm(lambda: (d.addPairwiseGaussian(3, Cm), d.addPairwiseBilateral(80, 13, img, Cm), d.synchronize()))
    Q = eigen.MatrixXf()
    d.inferenceInto(1, Q); d.synchronize()
    d._profile_enable(True); d._profile_read(True)
    d._timer_start()
    d.inferenceInto(2, Q)
    ms = d._timer_stop() / 2
    p = d._profile_read(True)
    print("cfg3 3840x2160 L=150 MatrixCompatibility: unary upload %.0f ms, lattices %.0f ms (V=%s), %.1f ms per iteration -> %.1f MPix*iter/s; per launch: %s"
          % (tu, ta, [d._lattice_info(k)[2] for k in range(2)], ms, W * H / ms / 1e3, {k: round(v[1] / max(v[0], 1), 3) for k, v in p.items()}))
