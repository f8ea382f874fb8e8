"""Differential test: the C restatement against the live reference (oracle/_ref/libref.so) on
seeded random inputs.  Skipped where the reference could not be compiled (no /root/reference and no
prebuilt file); tests/test_oracle_golden.py covers that case with committed vectors."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def po(oracle_built):
    if not oracle_built.available("ref"):
        pytest.skip("oracle/_ref/libref.so not present")
    return oracle_built


@pytest.mark.parametrize("d,N", [(2, 1000), (5, 1001), (5, 4096), (3, 2), (4, 777), (8, 300), (1, 64)])
def test_lattice_bit_exact(po, d, N):
    rng = np.random.default_rng(d * 1000 + N)
    f = (rng.standard_normal((d, N)) * rng.uniform(0.5, 6)).astype(np.float32)
    a, b = po.Lattice(f, "orc"), po.Lattice(f, "ref")
    assert a.V == b.V
    for x, y in zip(a.dump(), b.dump()):
        assert np.array_equal(x, y)
    for L in (1, 2, 3, 4, 21, 33):
        x = rng.standard_normal((N, L)).astype(np.float32)
        for rev in (False, True):
            assert np.array_equal(a.filter(x, rev), b.filter(x, rev)), (L, rev)


@pytest.mark.parametrize("d,N,L", [(2, 500, 3), (5, 333, 21), (5, 64, 1), (3, 100, 2), (8, 120, 4), (1, 40, 7)])
def test_lattice_gradient_bit_exact(po, d, N, L):
    """Permutohedral::gradient (permutohedral.cpp:612-695), incl. features with ties (ranks decide the rotation)."""
    rng = np.random.default_rng(d * 100 + N + L)
    f = (rng.standard_normal((d, N)) * 2).astype(np.float32)
    if N % 2 == 0:
        f = np.round(f * 2) / 2
    a = rng.standard_normal((N, L)).astype(np.float32)
    b = rng.random((N, L)).astype(np.float32)
    assert np.array_equal(po.Lattice(f, "orc").gradient(a, b), po.Lattice(f, "ref").gradient(a, b))


def test_lattice_ties_and_integers(po):
    """features on exact lattice positions exercise rank ties and round-half-even."""
    rng = np.random.default_rng(5)
    f = rng.integers(-6, 7, (3, 512)).astype(np.float32) * 0.5
    a, b = po.Lattice(f, "orc"), po.Lattice(f, "ref")
    for x, y in zip(a.dump(), b.dump()):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("L", [2, 3, 7, 21, 150])
def test_inference_bit_exact(po, L):
    from tests import workloads as wl

    W, H = 48, 36
    rng = np.random.default_rng(L)
    img = wl.smooth_image(H, W, seed=L, grid=8)
    U = wl.softmax_unary(H, W, L, seed=L, grid=8)
    Mx = wl.matrix_compat(L, seed=L)
    Dg = (-rng.random(L) * 3).astype(np.float32)
    for cg, cb in [(3.0, 10.0), (Dg, Mx), (Mx, 5.0)]:
        for nt in (3, 0, 1, 2):
            qs = []
            for which in ("orc", "ref"):
                c = po.CRF((W, H), L, which)
                c.setUnaryEnergy(U)
                c.addPairwiseGaussian((3, 2), cg, normalization=nt)
                c.addPairwiseBilateral((30, 25), (13, 11, 9), img, cb, normalization=nt)
                Q = np.array(c.startInference())
                c.stepInference(Q)
                qs.append((np.array(c.inference(3)), Q, c.klDivergence(Q)))
            assert np.array_equal(qs[0][0], qs[1][0]), (nt,)
            assert np.array_equal(qs[0][1], qs[1][1]), (nt,)
            assert qs[0][2] == qs[1][2]
            if L == 150:
                break


def test_logistic_unary(po):
    rng = np.random.default_rng(9)
    N, L, F = 300, 4, 6
    f = rng.standard_normal((3, N)).astype(np.float32)
    Lm = rng.standard_normal((L, F)).astype(np.float32)
    ff = rng.standard_normal((F, N)).astype(np.float32)
    qs = []
    for which in ("orc", "ref"):
        c = po.CRF(N, L, which)
        c.setUnaryLogistic(Lm, ff)
        c.addPairwiseEnergy(f, 2.0)
        qs.append(np.array(c.inference(3)))
    # Eigen's product order for L*f is not restated: tolerance, not equality
    assert np.abs(qs[0] - qs[1]).max() < 1e-5
