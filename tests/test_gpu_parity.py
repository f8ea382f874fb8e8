"""Parity of the CUDA path (through the public Python surface -> C-ABI) with the CPU oracle and the
committed reference vectors.  Run on the B200 box: python -m pytest tests -m gpu.

Tolerances (stated here, as BASELINE.json's north_star asks):
  * lattice tables (vertex ids in first-touch order, barycentric weights, neighbour ids, keys): EXACT
  * Permutohedral::compute (splat/blur/slice): EXACT (same float operations in the same order)
  * Q after mean-field iterations: north_star allows max-abs 1e-4; the engine restates Eigen's
    packet exp and summation order, so the tests hold it to Q_ATOL = 2e-6 (observed: 0 to a few ulp,
    from the libm-expf tail lanes and, for MatrixCompatibility, nothing else)
  * MAP: identical wherever the oracle's top-2 margin exceeds 1e-3
  * KL divergence: relative 2e-4 (the reference sums each pairwise term as ONE float over all L*N products,
    densecrf.cpp:232, which alone carries ~1e-4 relative error at these sizes; the engine accumulates in double)
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
Q_ATOL = 2e-6
KL_RTOL = 2e-4


@pytest.fixture(scope="module")
def dcrf():
    import pydensecrf_b200.densecrf as m

    assert m.device_count() >= 1, "GPU tests need a CUDA device: there is no CPU fallback to fall back to"
    return m


@pytest.fixture(scope="module")
def po(oracle_built):
    return oracle_built


def _golden(name):
    return np.load(os.path.join(G, name))


def _check_map(Q, Qref, margin=1e-3):
    s = np.sort(Qref, axis=0)
    sure = (s[-1] - s[-2]) > margin if Qref.shape[0] > 1 else np.ones(Qref.shape[1], bool)
    assert np.array_equal(np.argmax(Q, 0)[sure], np.argmax(Qref, 0)[sure])


def _lattice_of(dcrf, feats, L=1):
    d = dcrf.DenseCRF(feats.shape[1], L)
    d.addPairwiseEnergy(np.ascontiguousarray(feats, np.float32), 1.0)
    return d


# ---- lattice ---------------------------------------------------------------------------------------
def test_lattice_golden_exact(dcrf):
    g = _golden("lattice_small.npz")
    for ci, (d, N) in enumerate(g["cases"]):
        m = _lattice_of(dcrf, g["f%d" % ci])
        off, bary, keys, nbr, _ = m._lattice_dump(0)
        assert np.array_equal(off, g["ref_off%d" % ci]), (d, N)
        assert np.array_equal(bary, g["ref_bary%d" % ci]), (d, N)
        assert np.array_equal(nbr, g["ref_nbr%d" % ci]), (d, N)
        for L in (1, 2, 3, 21):
            y = m._lattice_filter(0, g["x%d_L%d" % (ci, L)])
            assert np.array_equal(y, g["ref_y%d_L%d" % (ci, L)]), (d, N, L)


@pytest.mark.parametrize("d,N,scale", [(2, 5000, 3.0), (5, 20001, 2.0), (5, 4096, 0.3), (3, 2, 1.0), (4, 777, 5.0),
                                       (7, 3000, 1.5), (1, 640, 4.0), (9, 500, 1.0), (5, 1, 1.0)])
def test_lattice_random_exact(dcrf, po, d, N, scale):
    rng = np.random.default_rng(100 * d + N)
    f = (rng.standard_normal((d, N)) * scale).astype(np.float32)
    o = po.Lattice(f, "orc")
    m = _lattice_of(dcrf, f)
    assert m._lattice_info(0) == (d, N, o.V)
    off, bary, keys, nbr, _ = m._lattice_dump(0)
    ooff, obary, _, onbr = o.dump()
    assert np.array_equal(keys, o.keys())      # the key SET and its first-touch numbering
    assert np.array_equal(off, ooff)
    assert np.array_equal(bary, obary)
    assert np.array_equal(nbr, onbr)
    for L in (1, 2, 4, 21, 37):
        x = rng.standard_normal((N, L)).astype(np.float32)
        assert np.array_equal(m._lattice_filter(0, x), o.filter(x)), L
        assert np.array_equal(m._lattice_filter(0, x, reverse=True), o.filter(x, reverse=True)), L


def test_lattice_ties_phantoms_and_duplicates(dcrf, po):
    rng = np.random.default_rng(1)
    # exact lattice positions: rank ties + round-half-even; N % 4 != 0: phantom origin simplex
    f = rng.integers(-6, 7, (3, 1023)).astype(np.float32) * 0.5
    # all pixels identical: one simplex, maximal contention on d+1 vertices
    g = np.full((5, 4099), 1.2345, np.float32)
    for feats in (f, g):
        o = po.Lattice(feats, "orc")
        m = _lattice_of(dcrf, feats)
        off, bary, keys, nbr, _ = m._lattice_dump(0)
        ooff, obary, _, onbr = o.dump()
        assert m._lattice_info(0)[2] == o.V
        assert np.array_equal(off, ooff) and np.array_equal(bary, obary) and np.array_equal(nbr, onbr)
        x = rng.random((feats.shape[1], 3)).astype(np.float32)
        assert np.array_equal(m._lattice_filter(0, x), o.filter(x))


@pytest.mark.parametrize("d,N,scale", [(12, 700, 1.0), (32, 300, 0.7), (2, 4000, 3.0e4), (5, 3000, 9.0e3)])
def test_lattice_edge_dims_and_key_wrap(dcrf, po, d, N, scale):
    """Feature dimensions beyond the templated ones (generic path up to 32) and coordinates large enough that
    the int16 vertex keys wrap (the reference truncates keys to short)."""
    rng = np.random.default_rng(d + N)
    f = (rng.standard_normal((d, N)) * scale).astype(np.float32)
    o = po.Lattice(f, "orc")
    m = _lattice_of(dcrf, f)
    off, bary, keys, nbr, _ = m._lattice_dump(0)
    ooff, obary, _, onbr = o.dump()
    assert m._lattice_info(0)[2] == o.V
    assert np.array_equal(keys, o.keys())
    assert np.array_equal(off, ooff) and np.array_equal(bary, obary) and np.array_equal(nbr, onbr)
    x = rng.standard_normal((N, 3)).astype(np.float32)
    assert np.array_equal(m._lattice_filter(0, x), o.filter(x))


def test_single_label_and_single_variable(dcrf, po):
    rng = np.random.default_rng(0)
    for N, L in [(1, 3), (7, 1), (2, 2)]:
        f = rng.standard_normal((2, N)).astype(np.float32)
        U = rng.random((L, N)).astype(np.float32)
        d = dcrf.DenseCRF(N, L)
        o = po.CRF(N, L, "orc")
        for c in (d, o):
            c.setUnaryEnergy(U)
            c.addPairwiseEnergy(f, 2.0)
        assert np.abs(np.array(d.inference(3)) - o.inference(3)).max() <= Q_ATOL


def test_lattice_2d_builders_exact(dcrf, po):
    from tests import workloads as wl

    W, H = 127, 65   # N % 4 != 0
    img = wl.smooth_image(H, W, seed=2, grid=16)
    d = dcrf.DenseCRF2D(W, H, 3)
    d.addPairwiseGaussian(sxy=(3, 2), compat=1.0)
    d.addPairwiseBilateral(sxy=(40, 30), srgb=(13, 11, 7), rgbim=img, compat=1.0)
    o = po.CRF((W, H), 3, "orc")
    o.addPairwiseGaussian((3, 2), 1.0)
    o.addPairwiseBilateral((40, 30), (13, 11, 7), img, 1.0)
    import ctypes as C
    for k in range(2):
        off, bary, keys, nbr, norm = d._lattice_dump(k)
        lat = po.Lattice.__new__(po.Lattice)
        lat.lib = po.load("orc")
        lat.h = None
        h = lat.lib.dll.orc_crf_term_lattice
        h.restype = C.c_void_p
        h.argtypes = [C.c_void_p, C.c_int]
        lat.N, lat.d = W * H, (2, 5)[k]
        hh = h(o.h, k)
        lat.V = lat.lib.lattice_nvert(hh)
        lat.h = hh
        ooff, obary, _, onbr = lat.dump()
        lat.h = None  # owned by the model
        assert np.array_equal(off, ooff) and np.array_equal(bary, obary) and np.array_equal(nbr, onbr)
        onorm = np.empty(W * H, np.float32)
        lat.lib.crf_term_norm(o.h, k, onorm.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(norm, onorm)


# ---- mean-field ------------------------------------------------------------------------------------
def test_toy_golden(dcrf):
    g = _golden("toy.npz")
    for nt in range(4):
        d = dcrf.DenseCRF2D(10, 10, 2)
        d.setUnaryEnergy(g["U"])
        d.addPairwiseBilateral(sxy=2, srgb=2, rgbim=g["img"], compat=3, normalization=nt)
        Q = np.array(d.inference(10))
        assert np.abs(Q - g["ref_Q_nt%d" % nt]).max() <= Q_ATOL, nt
    d = dcrf.DenseCRF(100, 2)
    d.setUnaryEnergy(g["U"])
    d.addPairwiseEnergy(g["generic_feats"], compat=3)
    Q = np.array(d.inference(10))
    assert np.abs(Q - g["ref_Q_generic"]).max() <= Q_ATOL
    exp = np.zeros((10, 10), np.int64)
    exp[2:8, 2:8] = 1
    assert np.array_equal(np.argmax(Q, 0).reshape(10, 10), exp)


@pytest.mark.parametrize("cname", ["potts", "diag", "matrix"])
@pytest.mark.parametrize("mode", [0, 1])
def test_model_small_golden(dcrf, cname, mode):
    g = _golden("model_small.npz")
    cg, cb = {"potts": (3.0, 10.0), "diag": (g["Dg"], 4.0), "matrix": (g["Mx"], g["Mx"])}[cname]
    worst = 0.0
    for nt in range(4):
        d = dcrf.DenseCRF2D(40, 30, 21)
        d._set_iteration_mode(mode)   # 0: automatic (fused where possible), 1: staged kernels
        d.setUnaryEnergy(g["U"])
        d.addPairwiseGaussian(3, cg, normalization=nt)
        d.addPairwiseBilateral(20, 13, g["img"], cb, normalization=nt)
        ref = g["ref_Q_%s_nt%d" % (cname, nt)]
        Q = np.array(d.inference(5))
        worst = max(worst, np.abs(Q - ref).max())
        assert np.abs(Q - ref).max() <= Q_ATOL, (nt, np.abs(Q - ref).max())
        _check_map(Q, ref)
        Q, t1, t2 = d.startInference()
        kls = []
        for _ in range(3):
            kls.append(d.klDivergence(Q))
            d.stepInference(Q, t1, t2)
        np.testing.assert_allclose(kls, g["ref_KL_%s_nt%d" % (cname, nt)], rtol=KL_RTOL)
    print("worst |dQ|", cname, mode, worst)


def test_no_unary_golden(dcrf):
    g = _golden("model_small.npz")
    d = dcrf.DenseCRF2D(40, 30, 21)
    d.addPairwiseGaussian(3, 3.0)
    assert np.abs(np.array(d.inference(2)) - g["ref_Q_nounary"]).max() <= Q_ATOL
    Q0 = np.array(d.startInference()[0])
    assert Q0.shape == (21, 1200) and not Q0.any()


def test_cfg1_im1_golden(dcrf):
    """BASELINE configs[0]: examples/inference.py on im1.png + anno1.png, 5 iterations, both branches."""
    g = _golden("cfg1_im1.npz")
    img, U, L = g["img"], g["U"], int(g["n_labels"])
    H, W = img.shape[:2]
    d = dcrf.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(U)
    d.addPairwiseGaussian(sxy=(3, 3), compat=3, kernel=dcrf.DIAG_KERNEL, normalization=dcrf.NORMALIZE_SYMMETRIC)
    d.addPairwiseBilateral(sxy=(80, 80), srgb=(13, 13, 13), rgbim=img, compat=10, kernel=dcrf.DIAG_KERNEL,
                           normalization=dcrf.NORMALIZE_SYMMETRIC)
    Q = np.array(d.inference(5))
    assert np.abs(Q - g["ref_Q_2d"]).max() <= Q_ATOL
    assert np.array_equal(np.argmax(Q, 0), np.argmax(g["ref_Q_2d"], 0))
    Q, t1, t2 = d.startInference()
    kls = []
    for _ in range(5):
        kls.append(d.klDivergence(Q))
        d.stepInference(Q, t1, t2)
    np.testing.assert_allclose(kls, g["ref_KL_2d"], rtol=KL_RTOL)

    from pydensecrf_b200.utils import create_pairwise_bilateral, create_pairwise_gaussian

    d = dcrf.DenseCRF(W * H, L)
    d.setUnaryEnergy(U)
    d.addPairwiseEnergy(create_pairwise_gaussian(sdims=(3, 3), shape=img.shape[:2]), compat=3)
    d.addPairwiseEnergy(create_pairwise_bilateral(sdims=(80, 80), schan=(13, 13, 13), img=img, chdim=2), compat=10)
    Q = np.array(d.inference(5))
    assert np.abs(Q - g["ref_Q_generic"]).max() <= Q_ATOL


def test_mid_golden(dcrf):
    from tests import workloads as wl

    g = _golden("mid_160x120.npz")
    W, H, L = 160, 120, 21
    d = dcrf.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(wl.softmax_unary(H, W, L, seed=5, grid=8))
    d.addPairwiseGaussian(3, 3.0)
    d.addPairwiseBilateral(80, 13, wl.smooth_image(H, W, seed=5, grid=16), 10.0)
    Qm = d.inference(10)
    Q = np.array(Qm)
    assert np.abs(Q[:, ::5] - g["ref_Q_sub"]).max() <= Q_ATOL
    sure = g["ref_margin"] > 1e-3
    assert np.array_equal(np.argmax(Q, 0)[sure], g["ref_MAP"][sure])
    assert np.array_equal(d.currentMap(Qm)[sure], g["ref_MAP"][sure])


@pytest.mark.parametrize("L,W,H", [(150, 67, 45), (64, 50, 40), (21, 64, 48), (250, 31, 29)])
def test_matrix_compat_tensor_cores(dcrf, po, L, W, H):
    """MatrixCompatibility on tcgen05 (3xTF32 split operands, fp32 accumulation in TMEM): default from 64 labels
    up, forced here for the smaller ones.  Not bit-faithful (summation order / split precision), so it is held to
    5e-6 on one message and 2e-5 on Q after 4 iterations -- still 5x inside north_star's 1e-4."""
    from tests import workloads as wl
    import pydensecrf_b200.eigen as eigen

    rng = np.random.default_rng(L)
    img = wl.smooth_image(H, W, seed=L, grid=8)
    U = wl.softmax_unary(H, W, L, seed=L, grid=8)
    Mx = wl.matrix_compat(L, seed=L)
    d = dcrf.DenseCRF2D(W, H, L)
    d.setMatrixCompatMode(2)
    e = dcrf.DenseCRF2D(W, H, L)
    e.setMatrixCompatMode(1)
    o = po.CRF((W, H), L, "orc")
    for c in (d, e, o):
        c.setUnaryEnergy(U)
        c.addPairwiseGaussian((3, 2), Mx)
        c.addPairwiseBilateral((30, 25), (13, 11, 9), img, 5.0)
    Qo = o.inference(4)
    Qe = np.array(e.inference(4))
    Qt = np.array(d.inference(4))
    assert np.abs(Qe - Qo).max() <= Q_ATOL          # exact-order kernel: bit-faithful as before
    err = np.abs(Qt - Qo).max()
    print("tensor-core matrix compat L=%d: max|dQ| = %.3e" % (L, err))
    assert err <= 2e-5
    _check_map(Qt, Qo)
    # one application of the pairwise term on its own
    Q0 = eigen.matrixXf(np.ascontiguousarray(Qo))
    mt, me = np.array(d._pairwise_apply(0, Q0)), np.array(e._pairwise_apply(0, Q0))
    assert np.abs(mt - me).max() <= 5e-6 * max(1.0, np.abs(me).max())


@pytest.mark.parametrize("L,W,H,ntype", [(150, 67, 45, "NORMALIZE_SYMMETRIC"), (64, 50, 40, "NORMALIZE_BEFORE"),
                                          (21, 64, 48, "NORMALIZE_AFTER"), (250, 31, 29, "NO_NORMALIZATION"),
                                          (150, 129, 3, "NORMALIZE_SYMMETRIC"),
                                          # the other instantiations of the TMA kernel: rows that are not 32-byte aligned
                                          # with three (20 labels) and two (100 -> Np 112... / 180) accumulator buffers, one
                                          # label short of / past a 16-column chunk, L % 4 = 1, 2, 3 tails
                                          (20, 40, 30, "NORMALIZE_SYMMETRIC"), (100, 33, 31, "NORMALIZE_BEFORE"),
                                          (180, 30, 20, "NORMALIZE_SYMMETRIC"), (33, 41, 37, "NORMALIZE_AFTER"),
                                          (47, 30, 30, "NORMALIZE_SYMMETRIC"), (130, 25, 21, "NORMALIZE_SYMMETRIC"),
                                          # several tiles per CTA: accumulator buffers and operand stages go round many times
                                          (150, 320, 240, "NORMALIZE_SYMMETRIC")])
def test_tensor_core_update_all_matrix_terms(dcrf, po, L, W, H, ntype):
    """When every term carries a MatrixCompatibility the whole update is one tcgen05 kernel (slice gathers feed the
    A operand, all terms share the TMEM accumulator, the softmax runs in the epilogue).  Held to the tensor-core
    tolerance against the oracle, and compared with the staged tensor-core path (slice -> GEMM -> softmax kernels)."""
    from tests import workloads as wl

    img = wl.smooth_image(H, W, seed=L, grid=8)
    U = wl.softmax_unary(H, W, L, seed=L, grid=8)
    M1, M2 = wl.matrix_compat(L, seed=L), wl.matrix_compat(L, seed=L + 1)
    nt = getattr(dcrf, ntype)
    d = dcrf.DenseCRF2D(W, H, L)
    d.setMatrixCompatMode(2)
    st = dcrf.DenseCRF2D(W, H, L)
    st.setMatrixCompatMode(2)
    st._set_iteration_mode(1)
    o = po.CRF((W, H), L, "orc")
    for c in (d, st, o):
        c.setUnaryEnergy(U)
        c.addPairwiseGaussian((3, 2), M1, normalization=(nt if c is not o else int(nt)))
        c.addPairwiseBilateral((30, 25), (13, 11, 9), img, M2, normalization=(nt if c is not o else int(nt)))
    Qo = o.inference(4)
    Qt = np.array(d.inference(4))
    Qs = np.array(st.inference(4))
    err, err_s = np.abs(Qt - Qo).max(), np.abs(Qt - Qs).max()
    print("tensor-core update L=%d %s: max|dQ| vs oracle %.3e, vs staged tensor-core path %.3e" % (L, ntype, err, err_s))
    # Un-normalised kernels with 250 labels give messages of magnitude ~15 and an ill-conditioned softmax argument: there
    # the 3xTF32 products (staged and fused alike) cost 6.6e-5 after 4 iterations -- inside north_star's 1e-4, but the
    # engine's default (mode 0) therefore keeps NO_NORMALIZATION terms on the exact kernel.
    tol = 1e-4 if ntype == "NO_NORMALIZATION" else 2e-5
    assert err <= tol and err_s <= tol
    _check_map(Qt, Qo)
    # stepInference goes through the same kernel with tmp1 = 0 - unary
    Q, t1, t2 = d.startInference()
    for _ in range(4):
        d.stepInference(Q, t1, t2)
    assert np.abs(np.array(Q) - Qo).max() <= tol
    if W * H * L > 5_000_000:
        # the reference sums every pairwise term as ONE float reduction (densecrf.cpp:232); over 10^7 products its own
        # rounding exceeds KL_RTOL, so the large case is compared in the mode that restates that reduction order
        d.setKLMode(1)
    kd, ko = d.klDivergence(Q), o.klDivergence(Qo)
    print("KL %.6f vs oracle %.6f" % (kd, ko))
    assert abs(kd - ko) <= KL_RTOL * abs(ko) + 1e-3


@pytest.mark.parametrize("L,W,H", [(2, 64, 48), (3, 50, 34), (7, 61, 47), (21, 128, 96), (150, 40, 30), (33, 31, 29)])
def test_inference_vs_oracle(dcrf, po, L, W, H):
    from tests import workloads as wl

    rng = np.random.default_rng(L)
    img = wl.smooth_image(H, W, seed=L, grid=8)
    U = wl.softmax_unary(H, W, L, seed=L, grid=8)
    Mx = wl.matrix_compat(L, seed=L)
    Dg = (-rng.random(L) * 3).astype(np.float32)
    for cg, cb in [(3.0, 10.0), (Dg, Mx), (Mx, 5.0)]:
        d = dcrf.DenseCRF2D(W, H, L)
        d.setMatrixCompatMode(1)   # exact-order kernel: this test is about bit-faithful arithmetic
        o = po.CRF((W, H), L, "orc")
        for c in (d, o):
            c.setUnaryEnergy(U)
            c.addPairwiseGaussian((3, 2), cg)
            c.addPairwiseBilateral((30, 25), (13, 11, 9), img, cb)
        Q, Qo = np.array(d.inference(4)), o.inference(4)
        err = np.abs(Q - Qo).max()
        assert err <= Q_ATOL, (L, err)
        _check_map(Q, Qo)
        np.testing.assert_allclose(d.klDivergence(d.inference(2)), o.klDivergence(np.array(o.inference(2))), rtol=KL_RTOL)


def test_logistic_unary(dcrf, po):
    rng = np.random.default_rng(9)
    N, L, F = 300, 4, 6
    f = rng.standard_normal((3, N)).astype(np.float32)
    Lm = rng.standard_normal((L, F)).astype(np.float32)
    ff = rng.standard_normal((F, N)).astype(np.float32)
    d = dcrf.DenseCRF(N, L)
    d.setUnary(dcrf.LogisticUnary(Lm, ff))
    d.addPairwiseEnergy(f, 2.0)
    o = po.CRF(N, L, "orc")
    o.setUnaryLogistic(Lm, ff)
    o.addPairwiseEnergy(f, 2.0)
    assert np.abs(np.array(d.inference(3)) - o.inference(3)).max() <= 1e-5


def test_step_equals_inference_and_views(dcrf):
    from tests import workloads as wl

    W, H, L = 64, 48, 5
    d = dcrf.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(wl.softmax_unary(H, W, L, seed=2, grid=8))
    d.addPairwiseGaussian(3, 3.0)
    d.addPairwiseBilateral(30, 13, wl.smooth_image(H, W, seed=2, grid=8), 10.0)
    Qi = np.array(d.inference(4))
    Q, t1, t2 = d.startInference()
    for _ in range(4):
        d.stepInference(Q, t1, t2)
    assert np.array_equal(np.array(Q), Qi)   # stepwise == inference(n), bit for bit
    # a live zero-copy view observes the next step, and writes through it are used by the next step
    Q, t1, t2 = d.startInference()
    view = np.asarray(Q)
    assert view.shape == (L, W * H) and view.strides == (4, 4 * L) and view.flags.writeable
    before = view.copy()
    d.stepInference(Q, t1, t2)
    assert not np.array_equal(view, before)
    view[...] = before          # write the old Q back through the view
    d.stepInference(Q, t1, t2)  # must redo exactly the first step
    Q2, a, b = d.startInference()
    d.stepInference(Q2, a, b)
    assert np.array_equal(np.asarray(Q), np.array(Q2))


def test_fused_equals_staged_and_deterministic(dcrf):
    from tests import workloads as wl

    W, H, L = 320, 240, 21
    U = wl.softmax_unary(H, W, L, seed=4)
    img = wl.smooth_image(H, W, seed=4)
    res = []
    for mode in (2, 1, 2):
        d = dcrf.DenseCRF2D(W, H, L)
        d._set_iteration_mode(mode)
        d.setUnaryEnergy(U)
        d.addPairwiseGaussian(3, 3.0)
        d.addPairwiseBilateral(80, 13, img, 10.0)
        res.append(np.array(d.inference(10)))
    assert np.array_equal(res[0], res[1])   # fused kernel == staged kernels, bit for bit
    assert np.array_equal(res[0], res[2])   # run-to-run deterministic (no float atomics anywhere)
    assert np.allclose(res[0].sum(0), 1.0, atol=1e-5)


def test_single_axis_blur_path(po):
    """Blur runs two axes per launch by default; with DCRF_B200_BLUR2=0 every axis is its own launch.  Both must
    reproduce the oracle filter bit for bit (fresh process: the switch is read once)."""
    import subprocess
    import sys

    code = r"""
import os, sys, numpy as np
sys.path.insert(0, %r)
os.environ["DCRF_B200_BLUR2"] = "0"
import pydensecrf_b200.densecrf as dcrf
from oracle import pyoracle as po
rng = np.random.default_rng(4)
for d, N in [(5, 9001), (2, 3000), (4, 777)]:
    f = (rng.standard_normal((d, N)) * 2.0).astype(np.float32)
    o = po.Lattice(f, "orc")
    m = dcrf.DenseCRF(N, 1)
    m.addPairwiseEnergy(f, 1.0)
    for L in (1, 2, 21):
        x = rng.standard_normal((N, L)).astype(np.float32)
        assert np.array_equal(m._lattice_filter(0, x), o.filter(x)), (d, N, L)
print("single-ok")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "single-ok" in r.stdout, r.stdout + r.stderr


def test_chain_blur_forced_small(po):
    """The in-place chain blur is normally used only for lattices with many short paths; force it on small
    inputs (fresh process: the threshold is read once) and demand bit-equality with the oracle filter."""
    import subprocess
    import sys

    code = r"""
import os, sys, numpy as np
sys.path.insert(0, %r)
os.environ["DCRF_B200_CHAIN_MIN"] = "1"
import pydensecrf_b200.densecrf as dcrf
from oracle import pyoracle as po
rng = np.random.default_rng(3)
for d, N, scale in [(5, 20001, 2.0), (2, 5000, 3.0), (3, 777, 1.0), (5, 4096, 0.3), (1, 640, 4.0)]:
    f = (rng.standard_normal((d, N)) * scale).astype(np.float32)
    o = po.Lattice(f, "orc")
    m = dcrf.DenseCRF(N, 1)
    m.addPairwiseEnergy(f, 1.0)
    for L in (1, 2, 3, 21, 37):
        x = rng.standard_normal((N, L)).astype(np.float32)
        assert np.array_equal(m._lattice_filter(0, x), o.filter(x)), (d, N, L)
print("chain-ok")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "chain-ok" in r.stdout, r.stdout + r.stderr


def test_nd_feature_builders_match_numpy_features(dcrf):
    """addPairwiseBilateralND / addPairwiseGaussianND build utils.create_pairwise_* features on the device; the
    lattices must be bit-identical to the ones built from the numpy features."""
    from pydensecrf_b200 import utils

    rng = np.random.default_rng(12)
    cases = [
        (rng.integers(0, 255, (9, 11, 13), dtype=np.uint8), -1, (3, 2.5, 2), (13,)),            # volume, no channel axis
        (rng.integers(0, 255, (24, 31, 3), dtype=np.uint8), 2, (5, 4), (13, 11, 9)),            # RGB, channel last
        (rng.random((2, 19, 23)).astype(np.float32) * 40, 0, (3, 3), 7.0),                      # float, channel first, scalar schan
    ]
    for img, chdim, sdims, schan in cases:
        f = utils.create_pairwise_bilateral(sdims=sdims, schan=schan, img=img, chdim=chdim)
        N = f.shape[1]
        a, b = dcrf.DenseCRF(N, 3), dcrf.DenseCRF(N, 3)
        a.addPairwiseEnergy(f, compat=2.0)
        b.addPairwiseBilateralND(sdims, schan, img, compat=2.0, chdim=chdim)
        for x, y in zip(a._lattice_dump(0), b._lattice_dump(0)):
            assert np.array_equal(x, y)
        shape = tuple(np.delete(img.shape, chdim)) if chdim != -1 else img.shape
        g = utils.create_pairwise_gaussian(sdims=sdims, shape=shape)
        a, b = dcrf.DenseCRF(N, 3), dcrf.DenseCRF(N, 3)
        a.addPairwiseEnergy(g, compat=2.0)
        b.addPairwiseGaussianND(sdims, shape, compat=2.0)
        for x, y in zip(a._lattice_dump(0), b._lattice_dump(0)):
            assert np.array_equal(x, y)
    with pytest.raises(ValueError):
        dcrf.DenseCRF(100, 2).addPairwiseGaussianND((3, 3), (9, 9), compat=1.0)


def test_unary_from_softmax_on_device(dcrf):
    """setUnaryFromSoftmax == utils.unary_from_softmax + setUnaryEnergy, up to the last ulp of log."""
    from pydensecrf_b200 import utils

    rng = np.random.default_rng(13)
    W, H, L = 40, 30, 5
    sm = np.ascontiguousarray(rng.dirichlet(np.ones(L) * 0.3, W * H).T, dtype=np.float32).reshape(L, H, W)
    sm[0, 0, :5] = 0.0
    for scale, clip in [(None, 1e-5), (0.7, 1e-5), (0.5, None)]:
        a, b = dcrf.DenseCRF2D(W, H, L), dcrf.DenseCRF2D(W, H, L)
        a.setUnaryEnergy(utils.unary_from_softmax(sm, scale=scale, clip=clip))
        b.setUnaryFromSoftmax(sm, scale=scale, clip=clip)
        for c in (a, b):
            c.addPairwiseGaussian(3, 3.0)
        qa, qb = np.array(a.inference(3)), np.array(b.inference(3))
        assert np.abs(qa - qb).max() <= 1e-6, (scale, clip)


def test_two_devices_in_one_process(dcrf):
    """Handles carry their device: two models on two GPUs driven from one process give identical results
    (replicas only -- SURVEY 8e).  Skipped on single-GPU boxes."""
    if dcrf.device_count() < 2:
        pytest.skip("needs two GPUs")
    from tests import workloads as wl

    W, H, L = 96, 64, 21
    U = wl.softmax_unary(H, W, L, seed=3, grid=8)
    img = wl.smooth_image(H, W, seed=3, grid=8)
    res = []
    for dev in (0, 1):
        d = dcrf.DenseCRF2D(W, H, L, device=dev)
        d.setUnaryEnergy(U)
        d.addPairwiseGaussian(3, 3.0)
        d.addPairwiseBilateral(80, 13, img, 10.0)
        res.append(np.array(d.inference(5)))
    assert np.array_equal(res[0], res[1])
    # the batch entry spread over both GPUs returns what one GPU returns
    Ub = np.stack([wl.softmax_unary(H, W, L, seed=b, grid=8) for b in range(10)])
    Ib = np.stack([wl.smooth_image(H, W, seed=b, grid=8) for b in range(10)])
    one = dcrf.map_batch_2d(Ub, Ib, n_iter=3, device=0)
    two = dcrf.map_batch_2d(Ub, Ib, n_iter=3, device=[0, 1], workers=4)
    assert np.array_equal(one, two)


def test_device_resident_io_with_torch(dcrf):
    """GPU-resident pipeline: unary taken from a torch CUDA tensor, result read back as a torch view through
    __cuda_array_interface__; must equal the host path bit for bit."""
    torch = pytest.importorskip("torch")
    from tests import workloads as wl

    W, H, L = 96, 64, 21
    U = wl.softmax_unary(H, W, L, seed=3, grid=8)
    img = wl.smooth_image(H, W, seed=3, grid=8)
    res = []
    for use_dev in (False, True):
        d = dcrf.DenseCRF2D(W, H, L, device=0)
        if use_dev:
            d.setUnaryEnergyDevice(torch.from_numpy(U).cuda(0))
        else:
            d.setUnaryEnergy(U)
        d.addPairwiseGaussian(3, 3.0)
        d.addPairwiseBilateral(80, 13, img, 10.0)
        Q = d.inference(5)
        if use_dev:
            t = torch.as_tensor(Q, device="cuda:0")          # (N, L) view on the engine's buffer, no copy
            assert tuple(t.shape) == (W * H, L)
            res.append(t.cpu().numpy().T)
        else:
            res.append(np.array(Q))
    assert np.array_equal(res[0], res[1])
    with pytest.raises(ValueError):
        d.setUnaryEnergyDevice(torch.zeros((L, 5), device="cuda:0"))


def test_fast_softmax_mode_within_tolerance(dcrf, po):
    """setSoftmaxMode(True) trades the bit-faithful softmax for an ex2-based one; it must still meet
    north_star's 1e-4 / MAP bar (held to 2e-5 here)."""
    from tests import workloads as wl

    W, H, L = 256, 192, 21
    U = wl.softmax_unary(H, W, L, seed=6)
    img = wl.smooth_image(H, W, seed=6)
    d = dcrf.DenseCRF2D(W, H, L)
    d.setSoftmaxMode(True)
    o = po.CRF((W, H), L, "orc")
    for c in (d, o):
        c.setUnaryEnergy(U)
        c.addPairwiseGaussian(3, 3.0)
        c.addPairwiseBilateral(80, 13, img, 10.0)
    Q, Qo = np.array(d.inference(10)), o.inference(10)
    err = np.abs(Q - Qo).max()
    print("fast softmax max|dQ| after 10 iterations:", err)
    assert err <= 2e-5
    _check_map(Q, Qo)


def test_512_vs_oracle(dcrf, po):
    """cfg 4 unit (one 512x512, 21 labels, 5 iterations) against the oracle."""
    from tests import workloads as wl

    W = H = 512
    L = 21
    U = wl.softmax_unary(H, W, L, seed=8)
    img = wl.smooth_image(H, W, seed=8)
    d = dcrf.DenseCRF2D(W, H, L)
    o = po.CRF((W, H), L, "orc")
    for c in (d, o):
        c.setUnaryEnergy(U)
        c.addPairwiseGaussian(3, 3.0)
        c.addPairwiseBilateral(80, 13, img, 10.0)
    Q, Qo = np.array(d.inference(5)), o.inference(5)
    assert np.abs(Q - Qo).max() <= Q_ATOL
    _check_map(Q, Qo)


def test_1080p_properties(dcrf):
    """BASELINE configs[1] at full size: size-independent properties (the oracle takes ~20 s here, so the
    value check at this size lives in bench.py's cpu_baseline leg, which compares Q on the same input)."""
    from tests import workloads as wl

    W, H, L = 1920, 1080, 21
    U = wl.softmax_unary(H, W, L, seed=0)
    img = wl.smooth_image(H, W, seed=0)
    d = dcrf.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(U)
    d.addPairwiseGaussian(3, 3.0)
    d.addPairwiseBilateral(80, 13, img, 10.0)
    Qm = d.inference(10)
    Q = np.array(Qm)
    assert Q.shape == (L, W * H) and np.isfinite(Q).all() and (Q >= 0).all()
    assert np.abs(Q.sum(0) - 1.0).max() <= 1e-5
    assert np.array_equal(d.currentMap(Qm), np.argmax(Q, 0).astype(np.int16))
    assert np.array_equal(np.array(d.inference(10)), Q)            # deterministic
    Qs, a, b = d.startInference()
    kl = []
    for _ in range(4):
        kl.append(d.klDivergence(Qs))
        d.stepInference(Qs, a, b)
    assert all(x > y for x, y in zip(kl, kl[1:]))                    # mean-field lowers the KL divergence
    for k, (dd, frac_lo, frac_hi) in enumerate([(2, 0.05, 0.3), (5, 0.05, 1.5)]):
        d_, N_, V_ = d._lattice_info(k)
        assert d_ == dd and N_ == W * H and frac_lo * N_ < V_ < frac_hi * N_


def test_1080p_parity_at_full_size(dcrf, po):
    """BASELINE configs[1] at FULL size against the checker (the unmodified reference C++ when oracle/_ref is present, else
    the pinned restatement): three iterations keep the CPU side at a few seconds; bench.py repeats the comparison for the
    10-iteration step and for every other BASELINE config.  north_star: max |dQ| <= 1e-4, MAP identical where the top-2
    margin exceeds 1e-3; held here to the suite's 2e-6."""
    import os
    from tests import workloads as wl

    W, H, L, n = 1920, 1080, 21, 3
    U = wl.softmax_unary(H, W, L, seed=0)
    img = wl.smooth_image(H, W, seed=0)
    which = "ref" if os.path.exists(os.path.join(os.path.dirname(po.__file__), "_ref", "libref.so")) else "orc"
    c = po.CRF((W, H), L, which)
    d = dcrf.DenseCRF2D(W, H, L)
    for m in (c, d):
        m.setUnaryEnergy(U)
        m.addPairwiseGaussian(3, 3.0)
        m.addPairwiseBilateral(80, 13, img, 10.0)
    Qc = c.inference(n)
    Q = np.array(d.inference(n))
    err = np.abs(Q - Qc).max()
    print("1080p, %d iterations, against %s: max|dQ| = %.3e" % (n, which, err))
    assert err <= Q_ATOL
    _check_map(Q, Qc)
    d.setKLMode(1)  # the reference's float reduction order (see test_kl_modes): 4e7 products per pairwise term here
    kd, kc = d.klDivergence(eigen_of(dcrf, Q)), c.klDivergence(Qc)
    print("KL %.3f vs %.3f" % (kd, kc))
    assert abs(kd - kc) <= KL_RTOL * abs(kc) + 1e-3


def eigen_of(dcrf, Q):
    import pydensecrf_b200.eigen as eigen

    return eigen.matrixXf(np.ascontiguousarray(Q))


def test_restated_arithmetic_selftests(dcrf):
    """The packed (FFMA2) restatement of Eigen's pexp equals the scalar restatement for every float in [-100, 0],
    and the shared-reciprocal column division equals IEEE division on 4e9 random operand pairs."""
    d = dcrf.DenseCRF(4, 3)
    assert d._selftest(0) == 0
    assert d._selftest(1, n=4_000_000_000, seed=12345) == 0


def test_map_batch_2d_equals_per_image_calls(dcrf):
    """the batch entry (worker threads inside the library, one model and stream each) returns exactly what
    the per-image calls return, in the caller's order, for more images than workers; errors surface as ValueError."""
    from tests import workloads as wl

    W, H, L, B = 96, 64, 7, 9
    U = np.stack([wl.softmax_unary(H, W, L, seed=b, grid=8) for b in range(B)])
    I = np.stack([wl.smooth_image(H, W, seed=b, grid=8) for b in range(B)])
    labels, Q = dcrf.map_batch_2d(U, I, sxy_gaussian=3, compat_gaussian=3, sxy_bilateral=(40, 30), srgb=(13, 11, 9),
                                  compat_bilateral=10, n_iter=5, workers=4, return_q=True)
    assert labels.shape == (B, H, W) and labels.dtype == np.int16 and Q.shape == (B, L, H * W)
    for b in range(B):
        d = dcrf.DenseCRF2D(W, H, L)
        d.setUnaryEnergy(U[b])
        d.addPairwiseGaussian(sxy=3, compat=3)
        d.addPairwiseBilateral(sxy=(40, 30), srgb=(13, 11, 9), rgbim=I[b], compat=10)
        Qb = d.inference(5)
        assert np.array_equal(Q[b], np.array(Qb))
        assert np.array_equal(labels[b].ravel(), d.currentMap(Qb))
    # unaries above 4 MB take the staged upload (page-locked double buffer + copy pool): several workers at once
    W2, H2, L2 = 256, 256, 21
    U2 = np.stack([wl.softmax_unary(H2, W2, L2, seed=b, grid=8) for b in range(6)])
    I2 = np.stack([wl.smooth_image(H2, W2, seed=b, grid=8) for b in range(6)])
    assert U2[0].nbytes > (4 << 20)
    lab2 = dcrf.map_batch_2d(U2, I2, n_iter=3, workers=4)
    for b in range(6):
        d = dcrf.DenseCRF2D(W2, H2, L2)
        d.setUnaryEnergy(U2[b])
        d.addPairwiseGaussian(sxy=3, compat=3)
        d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=I2[b], compat=10)
        assert np.array_equal(lab2[b].ravel(), d.currentMap(d.inference(3)))
    only = dcrf.map_batch_2d(U[:2], I[:2], n_iter=2, workers=1)
    assert only.shape == (2, H, W)
    assert dcrf.map_batch_2d(U[:0], I[:0]).shape == (0, H, W)
    with pytest.raises(ValueError):
        dcrf.map_batch_2d(U[:, :, :-1].copy(), I)
    with pytest.raises(ValueError):
        dcrf.map_batch_2d(U, I[:3])


def test_gaussian_lattice_is_shared_between_models(dcrf, po):
    """addPairwiseGaussian depends on the geometry only, so models of one image size share one lattice (built once per
    process).  Sharing must not change results (different normalisations, models destroyed in any order, a different
    sxy gets its own lattice) and the shared lattice must still equal the reference's."""
    from tests import workloads as wl

    W, H, L = 53, 37, 5
    U = wl.softmax_unary(H, W, L, seed=3, grid=8)
    img = wl.smooth_image(H, W, seed=3, grid=8)
    outs = []
    models = []
    for nt in (dcrf.NORMALIZE_SYMMETRIC, dcrf.NO_NORMALIZATION, dcrf.NORMALIZE_BEFORE, dcrf.NORMALIZE_SYMMETRIC):
        d = dcrf.DenseCRF2D(W, H, L)
        d.setUnaryEnergy(U)
        d.addPairwiseGaussian(sxy=(3, 2), compat=3, normalization=nt)
        d.addPairwiseBilateral(sxy=30, srgb=13, rgbim=img, compat=5)
        o = po.CRF((W, H), L, "orc")
        o.setUnaryEnergy(U)
        o.addPairwiseGaussian((3, 2), 3, normalization=int(nt))
        o.addPairwiseBilateral(30, 13, img, 5)
        outs.append((np.array(d.inference(3)), o.inference(3)))
        models.append(d)
    for q, qo in outs:
        assert np.abs(q - qo).max() <= Q_ATOL
    assert np.array_equal(outs[0][0], outs[3][0])
    # lattice of the shared term, read through two different models, equals the reference's
    feats = np.zeros((2, H * W), np.float32)
    feats[0] = (np.arange(H * W) % W) / 3.0
    feats[1] = (np.arange(H * W) // W) / 2.0
    ref_off, ref_bary = po.Lattice(feats, "orc").dump()[:2]
    for d in (models[0], models[2]):
        off, bary = d._lattice_dump(0)[:2]
        assert np.array_equal(off, ref_off) and np.array_equal(bary, ref_bary)
    del models[0], outs
    e = dcrf.DenseCRF2D(W, H, L)
    e.setUnaryEnergy(U)
    e.addPairwiseGaussian(sxy=(3, 2), compat=3)
    e.addPairwiseGaussian(sxy=(5, 5), compat=1)   # second geometry: its own lattice
    assert e._lattice_info(0)[2] != e._lattice_info(1)[2]
    assert np.isfinite(np.array(e.inference(2))).all()


def test_numpy_argmax_on_a_result_runs_on_the_device(dcrf):
    """`np.argmax(Q, axis=0)` (the reference's own post-processing line) dispatches to the device argmax while Q lives on
    the GPU and gives exactly NumPy's answer; everything else NumPy does with a MatrixXf is unchanged."""
    import pydensecrf_b200.eigen as eigen
    from tests import workloads as wl

    W, H, L = 61, 43, 21
    d = dcrf.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(wl.softmax_unary(H, W, L, seed=2, grid=8))
    d.addPairwiseGaussian(sxy=3, compat=3)
    d.addPairwiseBilateral(sxy=40, srgb=13, rgbim=wl.smooth_image(H, W, seed=2, grid=8), compat=10)
    Q = d.inference(5)
    assert Q._argmax_device() is not None                 # still on the device
    m = np.argmax(Q, axis=0)
    assert m.dtype == np.intp and m.shape == (H * W,)
    Qh = np.array(Q)
    assert np.array_equal(m, np.argmax(Qh, axis=0))
    assert np.array_equal(m, d.currentMap(Q))
    assert np.array_equal(np.argmax(Q, 0), m)             # positional axis
    # other axes / functions take NumPy's own implementation on the buffer view
    assert np.array_equal(np.argmax(Q, axis=1), np.argmax(Qh, axis=1))
    assert np.argmax(Q) == np.argmax(Qh)
    assert np.isclose(np.sum(Q), Qh.sum()) and np.allclose(np.max(Q, axis=0), Qh.max(0))
    assert np.stack([Q, Q]).shape == (2, L, H * W)
    # a matrix that only exists on the host is left to NumPy
    Mh = eigen.matrixXf(np.ascontiguousarray(Qh))
    assert Mh._argmax_device() is None
    assert np.array_equal(np.argmax(Mh, axis=0), np.argmax(Qh, axis=0))
    # a view that was written through invalidates the device copy: NumPy's path must see the edit
    v = np.asarray(Q)
    v[:, 0] = 0.0
    v[3, 0] = 1.0
    del v
    assert np.argmax(Q, axis=0)[0] == 3
