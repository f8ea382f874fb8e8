"""The C restatement (oracle/dcrf_oracle.c) against vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Bit-exact: the restatement reproduces the reference's float
arithmetic operation for operation, so `==`, not `allclose`."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def test_lattice_tables_and_filter(oracle_built):
    po = oracle_built
    g = _load("lattice_small.npz")
    for ci, (d, N) in enumerate(g["cases"]):
        lt = po.Lattice(g["f%d" % ci], "orc")
        off, bary, _, nbr = lt.dump()
        assert np.array_equal(off, g["ref_off%d" % ci]), (d, N)
        assert np.array_equal(bary, g["ref_bary%d" % ci]), (d, N)
        assert np.array_equal(nbr, g["ref_nbr%d" % ci]), (d, N)
        assert lt.V == g["ref_nbr%d" % ci].shape[1]
        for L in (1, 2, 3, 21):
            y = lt.filter(g["x%d_L%d" % (ci, L)])
            assert np.array_equal(y, g["ref_y%d_L%d" % (ci, L)]), (d, N, L)


def test_phantom_vertices(oracle_built):
    """N % 4 != 0 adds the origin simplex to the lattice: N=1, d=5 -> 12 vertices."""
    po = oracle_built
    lt = po.Lattice(np.full((5, 1), 3.7, np.float32), "orc")
    assert lt.V == 12
    lt = po.Lattice(np.full((5, 4), 3.7, np.float32), "orc")
    assert lt.V == 6


def test_softmax(oracle_built):
    po = oracle_built
    g = _load("softmax.npz")
    for L in (1, 2, 3, 4, 5, 8, 21, 150):
        y = po.exp_and_normalize(g["x_L%d" % L], "orc")
        assert np.array_equal(y, g["ref_y_L%d" % L]), L


def test_toy_all_normalizations(oracle_built):
    po = oracle_built
    g = _load("toy.npz")
    for nt in range(4):
        c = po.CRF((10, 10), 2, "orc")
        c.setUnaryEnergy(g["U"])
        c.addPairwiseBilateral(2, 2, g["img"], 3, normalization=nt)
        assert np.array_equal(c.inference(10), g["ref_Q_nt%d" % nt]), nt
    c = po.CRF(100, 2, "orc")
    c.setUnaryEnergy(g["U"])
    c.addPairwiseEnergy(g["generic_feats"], 3)
    q = c.inference(10)
    assert np.array_equal(q, g["ref_Q_generic"])
    # the reference's only numerical test: MAP of generic == MAP of 2D (pydensecrf/tests/test_dcrf.py:62-83)
    assert np.array_equal(np.argmax(q, 0), np.argmax(g["ref_Q_nt3"], 0))
    m = np.argmax(q, 0).reshape(10, 10)
    exp = np.zeros((10, 10), np.int64)
    exp[2:8, 2:8] = 1
    assert np.array_equal(m, exp)


@pytest.mark.parametrize("cname", ["potts", "diag", "matrix"])
def test_model_small(oracle_built, cname):
    po = oracle_built
    g = _load("model_small.npz")
    cg, cb = {"potts": (3.0, 10.0), "diag": (g["Dg"], 4.0), "matrix": (g["Mx"], g["Mx"])}[cname]
    for nt in range(4):
        c = po.CRF((40, 30), 21, "orc")
        c.setUnaryEnergy(g["U"])
        c.addPairwiseGaussian(3, cg, normalization=nt)
        c.addPairwiseBilateral(20, 13, g["img"], cb, normalization=nt)
        assert np.array_equal(c.inference(5), g["ref_Q_%s_nt%d" % (cname, nt)]), nt
        Q = np.array(c.startInference())
        kls = []
        for _ in range(3):
            kls.append(c.klDivergence(Q))
            c.stepInference(Q)
        assert kls == list(g["ref_KL_%s_nt%d" % (cname, nt)]), nt


def test_no_unary(oracle_built):
    po = oracle_built
    g = _load("model_small.npz")
    c = po.CRF((40, 30), 21, "orc")
    c.addPairwiseGaussian(3, 3.0)
    assert np.array_equal(c.inference(2), g["ref_Q_nounary"])
    assert np.array_equal(c.startInference(), g["ref_Q0_nounary"])
    assert not g["ref_Q0_nounary"].any()  # startInference without a unary leaves Q all-zero


def test_cfg1_im1(oracle_built):
    """BASELINE configs[0]: examples/inference.py on im1.png + anno1.png, both API branches."""
    po = oracle_built
    g = _load("cfg1_im1.npz")
    img, U, L = g["img"], g["U"], int(g["n_labels"])
    H, W = img.shape[:2]
    c = po.CRF((W, H), L, "orc")
    c.setUnaryEnergy(U)
    c.addPairwiseGaussian(3, 3.0)
    c.addPairwiseBilateral(80, 13, img, 10.0)
    assert np.array_equal(c.inference(5), g["ref_Q_2d"])
    Q = np.array(c.startInference())
    kls = []
    for _ in range(5):
        kls.append(c.klDivergence(Q))
        c.stepInference(Q)
    assert kls == list(g["ref_KL_2d"])


def test_mid(oracle_built):
    from tests import workloads as wl

    po = oracle_built
    g = _load("mid_160x120.npz")
    W, H, L = 160, 120, 21
    c = po.CRF((W, H), L, "orc")
    c.setUnaryEnergy(wl.softmax_unary(H, W, L, seed=5, grid=8))
    c.addPairwiseGaussian(3, 3.0)
    c.addPairwiseBilateral(80, 13, wl.smooth_image(H, W, seed=5, grid=16), 10.0)
    Q = c.inference(10)
    assert np.array_equal(Q[:, ::5], g["ref_Q_sub"])
    assert np.array_equal(np.argmax(Q, 0), g["ref_MAP"])


def test_lattice_gradient_matches_reference_golden(oracle_built):
    """Permutohedral::gradient (learning path): the restatement against vectors the reference produced."""
    po = oracle_built
    g = np.load(os.path.join(G, "gradient.npz"))
    for ci, (d, N, L) in enumerate(g["cases"]):
        lt = po.Lattice(g["f%d" % ci], "orc")
        assert np.array_equal(lt.gradient(g["a%d" % ci], g["b%d" % ci]), g["ref_df%d" % ci]), (d, N, L)
