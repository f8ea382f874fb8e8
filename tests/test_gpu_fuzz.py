"""Seeded differential fuzz of the CUDA path against the oracle (C restatement of the reference, pinned bit-for-bit
against the reference itself): random image sizes (odd ones included: phantom vertices, ragged strips and tiles), label
counts 1..40 (every row layout and the compile-time 21-label build), Potts / Diagonal / Matrix compatibilities, all
normalisations, one to three terms, 2-D builders and generic features, inference / stepInference / KL / currentMap."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

Q_ATOL = 2e-6


@pytest.fixture(scope="module")
def dcrf():
    import pydensecrf_b200.densecrf as m

    assert m.device_count() >= 1
    return m


@pytest.fixture(scope="module")
def po(oracle_built):
    return oracle_built


def _compat(rng, L, kind):
    if kind == 0:
        return float(rng.uniform(0.5, 8.0))
    if kind == 1:
        return (-rng.uniform(0.5, 6.0, L)).astype(np.float32)
    m = rng.uniform(-1.0, 1.0, (L, L)).astype(np.float32)
    m -= np.diag(rng.uniform(1.0, 5.0, L).astype(np.float32))
    return m


@pytest.mark.parametrize("seed", range(96))
def test_random_models_match_the_oracle(dcrf, po, seed):
    from tests import workloads as wl

    rng = np.random.default_rng(1000 + seed)
    W, H = int(rng.integers(3, 70)), int(rng.integers(3, 50))
    L = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 13, 20, 21, 22, 33, 40]))
    n_it = int(rng.integers(1, 4))
    img = wl.smooth_image(H, W, seed=seed, grid=4) if rng.random() < 0.7 else wl.noise_image(H, W, seed=seed)
    U = wl.softmax_unary(H, W, L, seed=seed, grid=4) if L > 1 else rng.uniform(0, 3, (1, H * W)).astype(np.float32)
    d = dcrf.DenseCRF2D(W, H, L)
    d.setMatrixCompatMode(1)  # bit-faithful arithmetic everywhere: this test is about exactness
    o = po.CRF((W, H), L, "orc")
    if rng.random() < 0.9:
        for c in (d, o):
            c.setUnaryEnergy(U)
    nterms = int(rng.integers(1, 4))
    for _ in range(nterms):
        kind = int(rng.integers(0, 3))
        comp = _compat(rng, L, kind)
        nt = int(rng.integers(0, 4))
        which = int(rng.integers(0, 3))
        if which == 0:
            sxy = (float(rng.uniform(1, 6)), float(rng.uniform(1, 6)))
            d.addPairwiseGaussian(sxy, comp, normalization=dcrf.NormalizationType(nt))
            o.addPairwiseGaussian(sxy, comp, normalization=nt)
        elif which == 1:
            sxy = (float(rng.uniform(5, 60)), float(rng.uniform(5, 60)))
            srgb = tuple(float(x) for x in rng.uniform(3, 30, 3))
            d.addPairwiseBilateral(sxy, srgb, img, comp, normalization=dcrf.NormalizationType(nt))
            o.addPairwiseBilateral(sxy, srgb, img, comp, normalization=nt)
        else:
            dd = int(rng.integers(1, 7))
            f = rng.uniform(-4, 4, (dd, H * W)).astype(np.float32)
            f[0] += np.arange(H * W, dtype=np.float32) % W / 4.0
            d.addPairwiseEnergy(f, comp, normalization=dcrf.NormalizationType(nt))
            o.addPairwiseEnergy(f, comp, normalization=nt)
    Qo = o.inference(n_it)
    Qm = d.inference(n_it)
    Q = np.array(Qm)
    assert np.isfinite(Q).all()
    err = np.abs(Q - Qo).max()
    assert err <= Q_ATOL, (seed, W, H, L, nterms, err)
    # MAP wherever the oracle's decision has a margin
    if L > 1:
        srt = np.sort(Qo, axis=0)
        sure = (srt[-1] - srt[-2]) > 1e-3
        assert np.array_equal(d.currentMap(Qm)[sure], np.argmax(Qo, 0)[sure])
    # startInference / stepInference against the reference's own start/step path (without a unary the reference starts
    # from Q = 0, not from the uniform distribution inference() starts from: densecrf.cpp:178-186 vs :116-120), and KL
    Qs, t1, t2 = d.startInference()
    Qso = np.array(o.startInference())
    assert np.abs(np.array(Qs) - Qso).max() <= Q_ATOL
    for _ in range(n_it):
        d.stepInference(Qs, t1, t2)
        o.stepInference(Qso)
    assert np.abs(np.array(Qs) - Qso).max() <= Q_ATOL
    kl, klo = d.klDivergence(Qs), o.klDivergence(Qso)
    assert abs(kl - klo) <= 2e-4 * abs(klo) + 1e-2, (seed, kl, klo)


@pytest.mark.parametrize("seed", range(32))
def test_random_generic_models_match_the_oracle(dcrf, po, seed):
    """DenseCRF (no image geometry): any number of variables (N % 4 != 0 adds the reference's phantom vertices), feature
    dimensions 1..9 (templated and run-time lattice builds), features with ties and duplicates."""
    rng = np.random.default_rng(5000 + seed)
    N = int(rng.integers(1, 900))
    L = int(rng.choice([1, 2, 3, 5, 8, 21, 24, 37]))
    n_it = int(rng.integers(1, 4))
    U = rng.uniform(0, 4, (L, N)).astype(np.float32)
    d = dcrf.DenseCRF(N, L)
    d.setMatrixCompatMode(1)
    o = po.CRF(N, L, "orc")
    for c in (d, o):
        c.setUnaryEnergy(U)
    for _ in range(int(rng.integers(1, 3))):
        dd = int(rng.integers(1, 10))
        f = rng.uniform(-3, 3, (dd, N)).astype(np.float32)
        if rng.random() < 0.5:
            f = np.round(f * 2) / 2          # ties on the rounding boundaries, duplicate points
        kind = int(rng.integers(0, 3))
        comp = _compat(rng, L, kind)
        nt = int(rng.integers(0, 4))
        d.addPairwiseEnergy(f, comp, normalization=dcrf.NormalizationType(nt))
        o.addPairwiseEnergy(f, comp, normalization=nt)
    Qo = o.inference(n_it)
    Q = np.array(d.inference(n_it))
    err = np.abs(Q - Qo).max()
    assert np.isfinite(Q).all() and err <= Q_ATOL, (seed, N, L, err)


def test_extreme_unaries_match_the_oracle(dcrf, po):
    """Huge, tiny and equal energies: the clamp of Eigen's exp (x >= -88.376), columns whose entries are all equal, and
    energies far outside float's exp range behave as in the reference (no NaN, same Q)."""
    from tests import workloads as wl

    W, H, L = 24, 17, 21
    rng = np.random.default_rng(77)
    U = rng.uniform(0, 5, (L, H * W)).astype(np.float32)
    U[:, 0] = 0.0                    # all labels equal
    U[:, 1] = 1e30                   # all huge and equal
    U[3, 2] = 1e30                   # one label impossible
    U[:, 3] = np.linspace(0, 200, L)  # spans far beyond exp's range: the clamp decides
    U[:, 4] = 88.0 + np.arange(L, dtype=np.float32) * 0.1
    U[5, 5] = -1e4                   # one label overwhelmingly likely
    img = wl.smooth_image(H, W, seed=1, grid=4)
    d = dcrf.DenseCRF2D(W, H, L)
    o = po.CRF((W, H), L, "orc")
    for c in (d, o):
        c.setUnaryEnergy(U)
        c.addPairwiseGaussian((3, 3), 3.0)
        c.addPairwiseBilateral((20, 20), (13, 13, 13), img, 10.0)
    for n_it in (0, 1, 3):
        Q = np.array(d.inference(n_it))
        Qo = o.inference(n_it)
        assert np.isfinite(Q).all() and np.isfinite(Qo).all()
        assert np.abs(Q - Qo).max() <= Q_ATOL, n_it
