"""The BASELINE.json configs other than the headline one, measured by bench.py into the `configs` block of its JSON line:

  cfg2_noise  configs[1] on the i.i.d.-noise image (worst case for the bilateral lattice), 10 iterations
  cfg3        3840x2160, 150 labels, MatrixCompatibility (tensor-core update kernel), 10 iterations
  cfg4        batch of 1024 512x512 images, 21 labels, 5 iterations, through map_batch_2d; under torchrun the batch is
              sharded over the ranks (one process per GPU, no collective)
  cfg5        generic DenseCRF, 256^3 volume, 2 labels, 4-D bilateral features, FULL_KERNEL, 5 iterations
  learning    (not a BASELINE config) DenseCRF::gradient(5) with all gradients at 512x512, 21 labels

Each entry carries the lattice sizes V, the device time, the stage-model HBM fraction (SURVEY 8d bytes over the measured
copy bandwidth) and, unless --no-cpu-baseline, a `parity` block against the reference C++ (oracle/_ref, else the C
restatement) AT FULL SIZE on one host core: max |dQ| and MAP mismatches where the reference's top-2 margin exceeds 1e-3.
The CPU side is bounded: one reference iteration for cfg3 (83 s), the full runs for the others (10-15 s each), two images
of the cfg4 batch.  bench.py is the only caller; oracle/ is used here as the checker and CPU baseline only."""
import time

import numpy as np

from tests import workloads as wl


def _checker():
    from oracle import pyoracle as po

    which = "ref" if po.available("ref") else "orc"
    if which == "orc" and not po.available("orc"):
        po.build()
    return po, which


def _parity(Qg, Qc, extra=None):
    s = np.sort(Qc, axis=0)
    sure = (s[-1] - s[-2]) > 1e-3 if Qc.shape[0] > 1 else np.ones(Qc.shape[1], bool)
    out = {"max_abs_dQ": float(np.abs(Qg - Qc).max()),
           "map_mismatch_where_margin_gt_1e-3": int((Qg.argmax(0)[sure] != Qc.argmax(0)[sure]).sum()),
           "pixels_checked": int(sure.sum())}
    if extra:
        out.update(extra)
    return out


def _time_inference(d, eigen, n_iter, reps, warm=2):
    Q = eigen.MatrixXf()
    for _ in range(warm):
        d.inferenceInto(n_iter, Q)
    d._timer_start()
    for _ in range(reps):
        d.inferenceInto(n_iter, Q)
    return d._timer_stop() / reps


def cfg2_noise(dcrf, eigen, device, hbm_peak, cpu):
    W, H, L, n = 1920, 1080, 21, 10
    img, U = wl.noise_image(H, W, seed=0), wl.softmax_unary(H, W, L, seed=0)
    d = dcrf.DenseCRF2D(W, H, L, device=device)
    d.setUnaryEnergy(U)
    d.addPairwiseGaussian(sxy=3, compat=3)
    d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
    ms = _time_inference(d, eigen, n, reps=5)
    lat = [d._lattice_info(k)[::2] for k in range(2)]
    b = wl.stage_bytes_per_iter(W * H, L, lat) * n
    out = {"workload": "DenseCRF2D 1920x1080, 21 labels, Gaussian(3)+bilateral(80,13) Potts, 10 iters, i.i.d.-noise image",
           "lattice_vertices": [v for _, v in lat], "ms_per_inference": ms, "value": W * H * n / ms / 1e3,
           "unit": "MPix*iter/s", "stage_model": {"bytes": b, "gbs": b / ms / 1e6, "frac": b / ms / 1e6 / hbm_peak}}
    if cpu:
        po, which = _checker()
        c = po.CRF((W, H), L, which)
        c.setUnaryEnergy(U)
        c.addPairwiseGaussian(3, 3.0)
        c.addPairwiseBilateral(80, 13, img, 10.0)
        t0 = time.perf_counter()
        Qc = c.inference(n)
        t = time.perf_counter() - t0
        out["parity"] = _parity(np.asarray(d.inference(n)), Qc, {"against": which, "iterations": n})
        out["cpu_baseline"] = {"value": W * H * n / t / 1e6, "unit": "MPix*iter/s", "cores": 1, "kind": "reference" if which == "ref" else "port"}
    return out


def cfg3(dcrf, eigen, device, hbm_peak, cpu):
    W, H, L, n = 3840, 2160, 150, 10
    img, U, Cm = wl.smooth_image(H, W, seed=0), wl.softmax_unary(H, W, L, seed=0), wl.matrix_compat(L)
    d = dcrf.DenseCRF2D(W, H, L, device=device)
    d.setUnaryEnergy(U)
    d.addPairwiseGaussian(3, Cm)
    d.addPairwiseBilateral(80, 13, img, Cm)
    ms = _time_inference(d, eigen, n, reps=2, warm=1)
    lat = [d._lattice_info(k)[::2] for k in range(2)]
    b = wl.stage_bytes_per_iter(W * H, L, lat) * n
    flop = 2.0 * 3 * 2 * W * H * L * L * n   # 3xTF32, two terms
    out = {"workload": "DenseCRF2D 3840x2160, 150 labels, Gaussian(3)+bilateral(80,13) MatrixCompatibility, 10 iters",
           "lattice_vertices": [v for _, v in lat], "ms_per_inference": ms, "ms_per_iteration": ms / n,
           "value": W * H * n / ms / 1e3, "unit": "MPix*iter/s",
           "stage_model": {"bytes": b, "gbs": b / ms / 1e6, "frac": b / ms / 1e6 / hbm_peak},
           "tensor": {"flop": flop, "tflops": flop / ms / 1e9, "note": "2*3*K*N*L^2 per iteration (3xTF32 split), K = 2 terms"}}
    d._profile_enable(True)
    d._profile_read(True)
    Q = eigen.MatrixXf()
    d.inferenceInto(2, Q)
    p = d._profile_read(True)
    d._profile_enable(False)
    out["per_launch_ms"] = {k: round(v[1] / max(v[0], 1), 4) for k, v in p.items()}
    if cpu:
        po, which = _checker()
        c = po.CRF((W, H), L, which)
        c.setUnaryEnergy(U)
        c.addPairwiseGaussian(3, Cm)
        c.addPairwiseBilateral(80, 13, img, Cm)
        t0 = time.perf_counter()
        Qc = c.inference(1)
        t = time.perf_counter() - t0
        out["parity"] = _parity(np.asarray(d.inference(1)), Qc, {"against": which, "iterations": 1,
                                "note": "one reference iteration at full 4K size (%.0f s on one core)" % t})
        out["cpu_baseline"] = {"value": W * H / t / 1e6, "unit": "MPix*iter/s", "cores": 1, "kind": "reference" if which == "ref" else "port",
                               "sample": "inference(1), %.0f s" % t}
    return out


def cfg5(dcrf, eigen, device, hbm_peak, cpu):
    from pydensecrf_b200 import utils

    D, L, n = 256, 2, 5
    vol, U = wl.smooth_volume(D, seed=0), wl.volume_unary(D, seed=0)
    d = dcrf.DenseCRF(D ** 3, L, device=device)
    d.setUnaryEnergy(U)
    d.addPairwiseBilateralND((3, 3, 3), (13,), vol, compat=4, chdim=-1, kernel=dcrf.FULL_KERNEL,
                             normalization=dcrf.NORMALIZE_SYMMETRIC)
    ms = _time_inference(d, eigen, n, reps=5)
    lat = [d._lattice_info(0)[::2]]
    b = wl.stage_bytes_per_iter(D ** 3, L, lat) * n
    out = {"workload": "DenseCRF 256^3 volume, 2 labels, 4-D bilateral features (xyz/3 + intensity/13), FULL_KERNEL, 5 iters",
           "lattice_vertices": [lat[0][1]], "ms_per_inference": ms, "value": D ** 3 * n / ms / 1e3, "unit": "MVox*iter/s",
           "stage_model": {"bytes": b, "gbs": b / ms / 1e6, "frac": b / ms / 1e6 / hbm_peak}}
    if cpu:
        po, which = _checker()
        f = utils.create_pairwise_bilateral(sdims=(3, 3, 3), schan=(13,), img=vol, chdim=-1)
        c = po.CRF(D ** 3, L, which)
        c.setUnaryEnergy(U)
        c.addPairwiseEnergy(f, 4.0, kernel=po.FULL_KERNEL)
        t0 = time.perf_counter()
        Qc = c.inference(n)
        t = time.perf_counter() - t0
        out["parity"] = _parity(np.asarray(d.inference(n)), Qc, {"against": which, "iterations": n})
        out["cpu_baseline"] = {"value": D ** 3 * n / t / 1e6, "unit": "MVox*iter/s", "cores": 1, "kind": "reference" if which == "ref" else "port"}
    return out


def learning(dcrf, eigen, device, hbm_peak, cpu):
    """SURVEY 8f-4 (not a BASELINE config): DenseCRF::gradient with all three gradients at a training-image size."""
    W, H, L, n = 512, 512, 21, 5
    img, U = wl.smooth_image(H, W, seed=1, grid=16), wl.softmax_unary(H, W, L, seed=1, grid=8)
    gt = np.argmax(-U, axis=0).astype(np.int16)
    gt[::7] = -1
    d = dcrf.DenseCRF2D(W, H, L, device=device)
    d.setUnaryEnergy(U)
    d.addPairwiseGaussian(sxy=3, compat=3)
    d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
    obj = dcrf.LogLikelihood(gt)
    got = d.gradient(n, obj)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        got = d.gradient(n, obj)
    ms = (time.perf_counter() - t0) / reps * 1e3
    out = {"workload": "DenseCRF2D 512x512, 21 labels, Gaussian(3)+bilateral(80,13) Potts, DIAG kernels: gradient(5) of the "
                       "LogLikelihood objective, label-compatibility (2) and kernel (2 + 5) gradients",
           "ms_per_gradient": ms, "timing": "host clock around the blocking call (results come back to the host)",
           "value": W * H * n / ms / 1e3, "unit": "MPix*iter/s"}
    po, which = _checker()
    if cpu and which == "ref":
        c = po.CRF((W, H), L, "ref")
        c.setUnaryEnergy(U)
        c.addPairwiseGaussian(3, 3.0)
        c.addPairwiseBilateral(80, 13, img, 10.0)
        t0 = time.perf_counter()
        want = c.gradient(n, po.LOGLIKELIHOOD, gt)
        t = time.perf_counter() - t0
        rel = [float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)) for a, b in zip(got[1:], want[1:]) if b.size]
        out["parity"] = {"against": "ref", "objective_rel": abs(got[0] - want[0]) / abs(want[0]),
                         "gradient_max_rel_to_largest_entry": max(rel), "tolerance": 1e-3,
                         "note": "the reference sums 5.5 M float products per entry in float, the engine in double"}
        out["cpu_baseline"] = {"value": W * H * n / t / 1e6, "unit": "MPix*iter/s", "cores": 1, "kind": "reference",
                               "sample": "the same gradient(5), %.1f s" % t}
    return out


def cfg4(dcrf, device, owned, hbm_peak, cpu, sync):
    """`owned`: the image indices of the 1024-image batch this rank processes (bench.shard_of).  The batch cycles 16
    distinct synthetic images (every image still gets its own model, lattice build, upload and download); it goes through
    map_batch_2d in calls of 128 images.  Returns (entry, seconds of this rank) -- the caller takes the max over ranks."""
    from pydensecrf_b200.utils import pinned_empty

    W = H = 512
    L, n, chunk, distinct = 21, 5, 128, 16
    imgs = np.stack([wl.smooth_image(H, W, seed=100 + i) for i in range(distinct)])
    Us = np.stack([wl.softmax_unary(H, W, L, seed=100 + i) for i in range(distinct)])
    ub = pinned_empty((chunk, L, H * W), np.float32)
    ib = pinned_empty((chunk, H, W, 3), np.uint8)
    for j in range(chunk):
        ub[j], ib[j] = Us[j % distinct], imgs[j % distinct]
    for _ in range(3):   # warm-up with the timed call shape: memory pools grow to their steady size, lattice cache, threads
        dcrf.map_batch_2d(ub[:min(chunk, max(len(owned), 1))], ib[:min(chunk, max(len(owned), 1))], n_iter=n, device=device)
    calls = [min(chunk, len(owned) - o) for o in range(0, len(owned), chunk)]
    sync()
    t0 = time.perf_counter()
    lab, per_call = None, []
    for c in calls:
        t1 = time.perf_counter()
        lab = dcrf.map_batch_2d(ub[:c], ib[:c], n_iter=n, device=device)
        per_call.append(round((time.perf_counter() - t1) * 1e3, 1))
    t = time.perf_counter() - t0
    entry = {"workload": "batch of 1024 512x512 images, 21 labels, Gaussian(3)+bilateral(80,13) Potts, 5 iters, sharded over the ranks",
             "api": "map_batch_2d(unaries, images, n_iter=5) in calls of 128 images from page-locked buffers: per image the unary + "
                    "image upload, both lattices (built per chunk of 16 images as one super-lattice), inference(5), int16 label map back",
             "images": "16 distinct synthetic images cycled", "unit": "images/s", "ms_per_call_of_128": per_call}
    if cpu:
        po, which = _checker()
        worst, mism, checked = 0.0, 0, 0
        labq, Qg = dcrf.map_batch_2d(ub[:2], ib[:2], n_iter=n, device=device, return_q=True)
        for j in range(2):
            c = po.CRF((W, H), L, which)
            c.setUnaryEnergy(Us[j])
            c.addPairwiseGaussian(3, 3.0)
            c.addPairwiseBilateral(80, 13, imgs[j], 10.0)
            Qc = c.inference(n)
            p = _parity(Qg[j], Qc)
            worst = max(worst, p["max_abs_dQ"])
            mism += p["map_mismatch_where_margin_gt_1e-3"]
            checked += p["pixels_checked"]
            s = np.sort(Qc, axis=0)
            sure = (s[-1] - s[-2]) > 1e-3
            mism += int((labq[j].reshape(-1)[sure] != Qc.argmax(0)[sure]).sum())
            assert np.array_equal(lab[j], labq[j])
        entry["parity"] = {"against": which, "images_checked": 2, "iterations": n, "max_abs_dQ": worst,
                           "map_mismatch_where_margin_gt_1e-3": mism, "pixels_checked": checked}
    return entry, t
