"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref.so).

Run in the build container, where /root/reference exists:
    make -C oracle all && python tests/golden/make_golden.py
The GPU box has no /root/reference; the committed fixtures (and the prebuilt _ref, which travels
with the snapshot but is never committed) are what the tests there compare against.

Every array called ref_* below was produced by the reference's own C++ (through oracle/ref_shim.cpp);
nothing here comes from the restatement or from the CUDA path.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tests import workloads as wl  # noqa: E402

REF_EXAMPLES = "/root/reference/examples"


def save(name, **kw):
    p = os.path.join(HERE, name)
    np.savez_compressed(p, **kw)
    print("%-22s %8.1f KB" % (name, os.path.getsize(p) / 1024))


def toy():
    U, img = wl.toy_unary(), wl.toy_image()
    out = {"U": U, "img": img}
    for nt in range(4):
        c = po.CRF((10, 10), 2, "ref")
        c.setUnaryEnergy(U)
        c.addPairwiseBilateral(2, 2, img, 3, normalization=nt)
        out["ref_Q_nt%d" % nt] = np.ascontiguousarray(c.inference(10))
    # values the survey recorded from the reference's Python build
    assert np.allclose(out["ref_Q_nt3"][0, :5], [0.9659939, 0.9665577, 0.9684372, 0.9732122, 0.9716119], atol=2e-7)
    assert np.allclose(out["ref_Q_nt1"][0, :3], [0.96293056, 0.9643855, 0.9684238], atol=2e-7)
    assert np.allclose(out["ref_Q_nt2"][0, :3], [0.9689846, 0.9687308, 0.96852446], atol=2e-7)
    # generic path with utils.create_pairwise_bilateral feature order (row, col, r, g, b)
    yy, xx = np.meshgrid(np.arange(10), np.arange(10), indexing="ij")
    f = np.concatenate([np.stack([yy, xx]).astype(np.float32) / 2, np.rollaxis(img, 2).astype(np.float32) / 2])
    f = np.ascontiguousarray(f.reshape(5, -1))
    c = po.CRF(100, 2, "ref")
    c.setUnaryEnergy(U)
    c.addPairwiseEnergy(f, 3)
    out["generic_feats"] = f
    out["ref_Q_generic"] = np.ascontiguousarray(c.inference(10))
    save("toy.npz", **out)


def lattices():
    rng = np.random.default_rng(7)
    out = {}
    cases = [(2, 96), (5, 101), (3, 7), (4, 64), (5, 1), (1, 33), (7, 50)]
    out["cases"] = np.array(cases, np.int32)
    for ci, (d, N) in enumerate(cases):
        f = (rng.standard_normal((d, N)) * 2.5).astype(np.float32)
        lt = po.Lattice(f, "ref")
        off, bary, rank, nbr = lt.dump()
        out["f%d" % ci] = f
        out["ref_off%d" % ci] = off
        out["ref_bary%d" % ci] = bary
        out["ref_nbr%d" % ci] = nbr
        for L in (1, 2, 3, 21):
            x = rng.random((N, L), dtype=np.float32)
            out["x%d_L%d" % (ci, L)] = x
            out["ref_y%d_L%d" % (ci, L)] = lt.filter(x)
    save("lattice_small.npz", **out)


def softmax():
    rng = np.random.default_rng(11)
    out = {}
    for L in (1, 2, 3, 4, 5, 8, 21, 150):
        x = (rng.standard_normal((64, L)) * 15).astype(np.float32)
        x[0] = -300.0
        x[1, 0] = 200.0
        out["x_L%d" % L] = x
        out["ref_y_L%d" % L] = po.exp_and_normalize(x, "ref")
    save("softmax.npz", **out)


def model_small():
    W, H, L = 40, 30, 21
    rng = np.random.default_rng(3)
    img = wl.smooth_image(H, W, seed=3, grid=8)
    U = wl.softmax_unary(H, W, L, seed=3, grid=8)
    Mx = wl.matrix_compat(L, seed=3)
    Dg = (-rng.random(L) * 4).astype(np.float32)
    out = {"img": img, "U": U, "Mx": Mx, "Dg": Dg}
    compats = {"potts": (3.0, 10.0), "diag": (Dg, 4.0), "matrix": (Mx, Mx)}
    for cname, (cg, cb) in compats.items():
        for nt in range(4):
            c = po.CRF((W, H), L, "ref")
            c.setUnaryEnergy(U)
            c.addPairwiseGaussian(3, cg, normalization=nt)
            c.addPairwiseBilateral(20, 13, img, cb, normalization=nt)
            out["ref_Q_%s_nt%d" % (cname, nt)] = np.ascontiguousarray(c.inference(5))
            Q = np.array(c.startInference())
            kls = []
            for _ in range(3):
                kls.append(c.klDivergence(Q))
                c.stepInference(Q)
            out["ref_KL_%s_nt%d" % (cname, nt)] = np.array(kls)
    # no unary at all (densecrf.cpp:117 / :178-186)
    c = po.CRF((W, H), L, "ref")
    c.addPairwiseGaussian(3, 3.0)
    out["ref_Q_nounary"] = np.ascontiguousarray(c.inference(2))
    out["ref_Q0_nounary"] = np.ascontiguousarray(c.startInference())
    save("model_small.npz", **out)


def cfg1():
    import cv2

    img = cv2.imread(os.path.join(REF_EXAMPLES, "im1.png"))
    anno = cv2.imread(os.path.join(REF_EXAMPLES, "anno1.png")).astype(np.uint32)
    lbl = anno[:, :, 0] + (anno[:, :, 1] << 8) + (anno[:, :, 2] << 16)
    colors, labels = np.unique(lbl, return_inverse=True)
    assert 0 in colors
    n_labels = len(set(labels.flat)) - 1
    labels = labels.reshape(-1)
    # pydensecrf.utils.unary_from_labels(labels, n_labels, gt_prob=0.7, zero_unsure=True)  (utils.py:6-46)
    n_e, p_e = -np.log(0.3 / (n_labels - 1)), -np.log(0.7)
    U = np.full((n_labels, labels.size), n_e, dtype="float32")
    U[labels - 1, np.arange(labels.size)] = p_e
    U[:, labels == 0] = -np.log(1.0 / n_labels)
    H, W = img.shape[:2]
    out = {"img": img, "labels": labels.astype(np.uint8), "n_labels": n_labels, "U": U}
    # use_2d branch (examples/inference.py:72-90)
    c = po.CRF((W, H), n_labels, "ref")
    c.setUnaryEnergy(U)
    c.addPairwiseGaussian(3, 3.0)
    c.addPairwiseBilateral(80, 13, img, 10.0)
    out["ref_Q_2d"] = np.ascontiguousarray(c.inference(5))
    Q = np.array(c.startInference())
    kls = []
    for _ in range(5):
        kls.append(c.klDivergence(Q))
        c.stepInference(Q)
    out["ref_KL_2d"] = np.array(kls)
    # identical to what the reference's own Python extension (built from /root/reference with the
    # recipe in SURVEY.md section 8c) prints for this input in this container
    assert kls == [-543710.7136065669, -890203.4196090547, -919601.7338062561, -921429.4572714955,
                   -922284.9388205523], kls
    if os.path.exists("/tmp/refpy_Q.npy"):  # Q from that Python build, when it is lying around
        assert np.array_equal(np.load("/tmp/refpy_Q.npy"), out["ref_Q_2d"])
    # generic branch (examples/inference.py:91-111): features in (row, col[, b, g, r]) order
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    mesh = np.stack([yy, xx]).astype(np.float32)
    fg = np.ascontiguousarray((mesh / 3).reshape(2, -1))
    fb = np.ascontiguousarray(np.concatenate([mesh / 80, np.rollaxis(img, 2).astype(np.float32) / 13]).reshape(5, -1))
    c = po.CRF(W * H, n_labels, "ref")
    c.setUnaryEnergy(U)
    c.addPairwiseEnergy(fg, 3.0)
    c.addPairwiseEnergy(fb, 10.0)
    out["ref_Q_generic"] = np.ascontiguousarray(c.inference(5))
    save("cfg1_im1.npz", **out)


def mid():
    """160x120, 21 labels, the cfg 2 recipe at 1/12 scale, 10 iterations."""
    W, H, L = 160, 120, 21
    img = wl.smooth_image(H, W, seed=5, grid=16)
    U = wl.softmax_unary(H, W, L, seed=5, grid=8)
    c = po.CRF((W, H), L, "ref")
    c.setUnaryEnergy(U)
    c.addPairwiseGaussian(3, 3.0)
    c.addPairwiseBilateral(80, 13, img, 10.0)
    Q = np.ascontiguousarray(c.inference(10))
    srt = np.sort(Q, axis=0)
    save("mid_160x120.npz", seed=5, ref_MAP=np.argmax(Q, axis=0).astype(np.uint8),
         ref_margin=(srt[-1] - srt[-2]).astype(np.float32), ref_Q_sub=Q[:, ::5].copy())


def gradient():
    """Permutohedral::gradient (permutohedral.cpp:612-695) from the reference on small lattices."""
    rng = np.random.default_rng(17)
    out = {}
    cases = [(2, 60, 3), (5, 101, 21), (3, 9, 1), (4, 64, 2), (7, 40, 5)]
    out["cases"] = np.array(cases, np.int32)
    for ci, (d, N, L) in enumerate(cases):
        f = (rng.standard_normal((d, N)) * 2.5).astype(np.float32)
        a = rng.standard_normal((N, L)).astype(np.float32)
        b = rng.random((N, L)).astype(np.float32)
        out["f%d" % ci], out["a%d" % ci], out["b%d" % ci] = f, a, b
        out["ref_df%d" % ci] = po.Lattice(f, "ref").gradient(a, b)
    save("gradient.npz", **out)


def _utils_cases():
    """Inputs for the utils goldens: (name, function name, args, kwargs)."""
    rng = np.random.default_rng(21)
    lab2 = rng.integers(0, 5, (6, 7))
    sm3 = rng.random((4, 5, 6)).astype(np.float32)
    sm3 /= sm3.sum(0, keepdims=True)
    sm64 = rng.random((3, 11))
    sm64 /= sm64.sum(0, keepdims=True)
    im_hwc = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    im_chw = np.ascontiguousarray(np.moveaxis(im_hwc, 2, 0))
    vol = rng.integers(0, 256, (3, 4, 5), dtype=np.uint8)
    fim = rng.random((4, 6, 2)).astype(np.float32)
    return [
        ("labels_unsure", "unary_from_labels", (lab2, 4, 0.7), {"zero_unsure": True}),
        ("labels_sure", "unary_from_labels", (lab2, 5, 0.55), {"zero_unsure": False}),
        ("compute_unary", "compute_unary", (lab2, 4), {"GT_PROB": 0.8}),
        ("softmax_default", "unary_from_softmax", (sm3,), {}),
        ("softmax_scale", "unary_from_softmax", (sm3,), {"scale": 0.7, "clip": 1e-3}),
        ("softmax_noclip", "unary_from_softmax", (sm3 + 1e-3,), {"scale": 1.0, "clip": None}),
        ("softmax_f64", "unary_from_softmax", (sm64,), {"scale": 0.25}),
        ("softmax_to_unary", "softmax_to_unary", (sm3 + 1e-4,), {"GT_PROB": 0.9}),
        ("gauss_2d", "create_pairwise_gaussian", ((3, 2.5), (5, 7)), {}),
        ("gauss_3d", "create_pairwise_gaussian", ((1, 2, 3), (3, 4, 5)), {}),
        ("bilat_chdim2", "create_pairwise_bilateral", ((80, 60), (13, 11, 9), im_hwc), {"chdim": 2}),
        ("bilat_chdim0", "create_pairwise_bilateral", ((3, 3), 13, im_chw), {"chdim": 0}),
        ("bilat_nochan", "create_pairwise_bilateral", ((3, 3, 3), (13,), vol), {"chdim": -1}),
        ("bilat_float", "create_pairwise_bilateral", ((2, 2), (0.1, 0.2), fim), {"chdim": 2}),
    ]


def utils_golden():
    """Outputs of the reference's own pydensecrf/utils.py (pure numpy, imported by path) on fixed inputs."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_utils", "/root/reference/pydensecrf/utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for name, fn, args, kw in _utils_cases():
        out["ref_" + name] = getattr(ref, fn)(*args, **kw)
    save("utils.npz", **out)


def gradient_model():
    """DenseCRF::gradient and the parameter vectors of the unmodified reference on tests/gradient_cases.py, before and
    after an update of all three parameter groups (learning path)."""
    from tests import gradient_cases as gc
    from tests.test_gradient import KIND, updated_parameters
    out = {}
    for case in gc.CASES:
        c = gc.make(case)
        n = c["name"]
        ref = gc.build(po.CRF(c["N"], c["L"], "ref"), c)
        ob = c["objective"]
        kind = KIND[ob[0]]
        robust = ob[1] if kind == 0 else 0.0
        cw = po.hamming_weights(c["gt"], ob[1], c["L"]) if kind == 1 else np.zeros(0, np.float32)
        out[n + "/class_weight"] = cw
        for g in range(3):
            out[n + "/params%d" % g] = ref.parameters(g)
        r = ref.gradient(c["n_iter"], kind, c["gt"], robust, cw)
        out[n + "/value"], out[n + "/unary"], out[n + "/lbl"], out[n + "/kernel"] = np.float64(r[0]), r[1], r[2], r[3]
        new = updated_parameters(c, ref.parameters(1), ref.parameters(0))
        for g in range(3):
            out[n + "/new%d" % g] = new[g]
            if new[g].size:
                ref.setParameters(g, new[g])
        r = ref.gradient(c["n_iter"], kind, c["gt"], robust, cw)
        out[n + "/value2"], out[n + "/unary2"], out[n + "/lbl2"], out[n + "/kernel2"] = np.float64(r[0]), r[1], r[2], r[3]
        out[n + "/Q2"] = np.ascontiguousarray(ref.inference(c["n_iter"]))
    save("gradient_model.npz", **out)


if __name__ == "__main__":
    po.build()
    assert po.available("ref"), "oracle/_ref/libref.so missing: /root/reference not visible?"
    toy()
    lattices()
    softmax()
    model_small()
    cfg1()
    mid()
    utils_golden()
    gradient()
    gradient_model()
