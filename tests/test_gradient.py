"""Learning path: DenseCRF::gradient (densecrf.cpp:237-297), the three objectives (objective.cpp) and the
parameter vectors (densecrf.cpp:298-375) -- the numpy restatement against the unmodified reference and against the
committed vectors the reference produced (tests/golden/gradient_model.npz, made by make_golden.py), and the engine
against both.

Tolerance.  Every gradient entry is a sum over the variables that the reference takes in float (Eigen reductions) and
the engine in double; they are compared relative to the largest entry of the vector: 2e-4 (observed <= 6e-5, worst for
NO_NORMALIZATION whose un-normalised sums are the largest), plus 5e-7 absolute for entries that cancel to zero.  The objective value is a double sum on both sides: 1e-6."""
import os

import numpy as np
import pytest

from tests import gradient_cases as gc

KIND = {"ll": 0, "hamming": 1, "iou": 2}
GRAD_RTOL = 2e-4
GRAD_ATOL = 5e-7   # an entry that cancels to ~0 (every summand is O(1/N)) is float noise of the reference's own sum
VALUE_RTOL = 1e-6
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gradient_model.npz")


def objective_args(po_or_none, c, golden=None):
    ob = c["objective"]
    kind = KIND[ob[0]]
    robust = ob[1] if kind == 0 else 0.0
    cw = ()
    if kind == 1:
        cw = golden[c["name"] + "/class_weight"] if golden is not None else po_or_none.hamming_weights(c["gt"], ob[1], c["L"])
    return kind, robust, cw


def close(got, want, what):
    assert abs(got[0] - want[0]) <= VALUE_RTOL * max(abs(want[0]), 1e-12), (what, "value", got[0], want[0])
    for name, a, b in zip(("unary", "lbl_cmp", "kernel"), got[1:], want[1:]):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape, (what, name, a.shape, b.shape)
        if a.size:
            scale = max(float(np.abs(b).max()), 1e-12)
            assert float(np.abs(a - b).max()) <= GRAD_RTOL * scale + GRAD_ATOL, (what, name, float(np.abs(a - b).max()) / scale)


def updated_parameters(c, base_lbl, base_unary):
    rng = np.random.default_rng(7)
    return (base_unary * rng.uniform(0.8, 1.2, base_unary.shape).astype(np.float32),
            base_lbl * rng.uniform(0.8, 1.2, base_lbl.shape).astype(np.float32),
            gc.kernel_param_update(c, rng))


@pytest.fixture(scope="module")
def po(oracle_built):
    if not oracle_built.available("ref"):
        pytest.skip("oracle/_ref/libref.so not present")
    return oracle_built


@pytest.mark.parametrize("case", gc.CASES, ids=[c[0] for c in gc.CASES])
def test_restatement_matches_reference(po, case):
    c = gc.make(case)
    ref = gc.build(po.CRF(c["N"], c["L"], "ref"), c)
    res = gc.build(po.GradientRestated(c["N"], c["L"]), c)
    kind, robust, cw = objective_args(po, c)
    for g in range(3):
        assert np.array_equal(ref.parameters(g), res.parameters(g)), g
    close(res.gradient(c["n_iter"], kind, c["gt"], robust, cw), ref.gradient(c["n_iter"], kind, c["gt"], robust, cw), "initial")
    # setUnaryParameters / setLabelCompatibilityParameters / setKernelParameters (the kernels rebuild their lattices)
    new = updated_parameters(c, ref.parameters(1), ref.parameters(0))
    for g in range(3):
        if new[g].size:
            ref.setParameters(g, new[g])
            res.setParameters(g, new[g])
        assert np.array_equal(ref.parameters(g), res.parameters(g)), g
    close(res.gradient(c["n_iter"], kind, c["gt"], robust, cw), ref.gradient(c["n_iter"], kind, c["gt"], robust, cw), "updated")


@pytest.mark.parametrize("case", gc.CASES, ids=[c[0] for c in gc.CASES])
def test_restatement_matches_golden(oracle_built, case):
    g = np.load(GOLDEN)
    c = gc.make(case)
    res = gc.build(oracle_built.GradientRestated(c["N"], c["L"]), c)
    kind, robust, cw = objective_args(None, c, g)
    n = c["name"]
    close(res.gradient(c["n_iter"], kind, c["gt"], robust, cw), (float(g[n + "/value"]), g[n + "/unary"], g[n + "/lbl"], g[n + "/kernel"]), "initial")
    for grp in range(3):
        assert np.array_equal(res.parameters(grp), g[n + "/params%d" % grp])
        if g[n + "/new%d" % grp].size:
            res.setParameters(grp, g[n + "/new%d" % grp])
    close(res.gradient(c["n_iter"], kind, c["gt"], robust, cw), (float(g[n + "/value2"]), g[n + "/unary2"], g[n + "/lbl2"], g[n + "/kernel2"]), "updated")


FUZZ = list(range(24))


@pytest.mark.parametrize("seed", FUZZ)
def test_restatement_matches_reference_fuzz(po, seed):
    c = gc.make(gc.fuzz_case(seed))
    ref = gc.build(po.CRF(c["N"], c["L"], "ref"), c)
    res = gc.build(po.GradientRestated(c["N"], c["L"]), c)
    kind, robust, cw = objective_args(po, c)
    close(res.gradient(c["n_iter"], kind, c["gt"], robust, cw), ref.gradient(c["n_iter"], kind, c["gt"], robust, cw), c["name"])


# ---- the engine -------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dcrf():
    import pydensecrf_b200.densecrf as d
    if d.device_count() < 1:
        pytest.fail("no CUDA device: pydensecrf_b200 has no CPU path")
    return d


def engine_objective(dcrf, c, cw):
    ob = c["objective"]
    if ob[0] == "ll":
        return dcrf.LogLikelihood(c["gt"], ob[1])
    if ob[0] == "hamming":
        return dcrf.Hamming(c["gt"], cw)
    return dcrf.IntersectionOverUnion(c["gt"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", gc.CASES, ids=[c[0] for c in gc.CASES])
def test_engine_gradient_matches_reference_golden(dcrf, oracle_built, case):
    """dcrf_gradient / dcrf_{get,set}_parameters against vectors of the unmodified reference, before and after a
    parameter update of all three groups, and against the numpy restatement run on the same inputs."""
    g = np.load(GOLDEN)
    c = gc.make(case)
    n = c["name"]
    m = dcrf.DenseCRF(c["N"], c["L"])
    gc.build(m, c, logistic=lambda Lm, f: m.setUnary(dcrf.LogisticUnary(Lm, f)))
    kind, robust, cw = objective_args(None, c, g)
    obj = engine_objective(dcrf, c, cw)
    if c["objective"][0] == "hamming":  # Hamming(gt, class_weight_pow) builds the reference's class weights (objective.cpp:51-63)
        own = dcrf.Hamming(c["gt"], c["objective"][1]).class_weights
        assert np.allclose(own, cw[:own.size], rtol=1e-6, atol=0) and own.size == cw.size
    getters = (m.unaryParameters, m.labelCompatibilityParameters, m.kernelParameters)
    setters = (m.setUnaryParameters, m.setLabelCompatibilityParameters, m.setKernelParameters)
    for grp in range(3):
        assert np.array_equal(getters[grp](), g[n + "/params%d" % grp]), grp
    got = m.gradient(c["n_iter"], obj)
    close(got, (float(g[n + "/value"]), g[n + "/unary"], g[n + "/lbl"], g[n + "/kernel"]), "initial")
    assert all(np.array_equal(a, b) for a, b in zip(got[1:], m.gradient(c["n_iter"], obj)[1:])), "not deterministic"
    res = gc.build(oracle_built.GradientRestated(c["N"], c["L"]), c)
    close(got, res.gradient(c["n_iter"], kind, c["gt"], robust, cw), "restatement")
    for grp in range(3):
        if g[n + "/new%d" % grp].size:
            setters[grp](g[n + "/new%d" % grp])
            assert np.array_equal(getters[grp](), g[n + "/new%d" % grp]), grp
    close(m.gradient(c["n_iter"], obj), (float(g[n + "/value2"]), g[n + "/unary2"], g[n + "/lbl2"], g[n + "/kernel2"]), "updated")
    # inference after the update runs on the rebuilt lattices and the new compatibilities
    Q = np.array(m.inference(c["n_iter"]))
    assert np.abs(Q - g[n + "/Q2"]).max() <= 2e-5


@pytest.mark.gpu
def test_engine_gradient_selected_outputs_and_2d_builders(dcrf, oracle_built):
    """Gradients that are not asked for come back empty; DenseCRF2D's own feature builders keep their features for the
    kernel gradient (checked against the same model built from explicit feature matrices)."""
    W, H, L = 24, 18, 4
    rng = np.random.default_rng(3)
    U = rng.standard_normal((L, W * H)).astype(np.float32)
    img = rng.integers(0, 255, (H, W, 3)).astype(np.uint8)
    gt = rng.integers(0, L, W * H).astype(np.int16)
    a = dcrf.DenseCRF2D(W, H, L)
    a.setUnaryEnergy(U)
    a.addPairwiseGaussian(sxy=(3, 2), compat=2.0)
    a.addPairwiseBilateral(sxy=(20, 25), srgb=(13, 14, 15), rgbim=img, compat=4.0, kernel=dcrf.FULL_KERNEL)
    yy, xx = np.mgrid[0:H, 0:W]
    fg = np.stack([xx.ravel() / np.float32(3), yy.ravel() / np.float32(2)]).astype(np.float32)
    fb = np.stack([xx.ravel() / np.float32(20), yy.ravel() / np.float32(25)] +
                  [img[..., k].ravel() / np.float32(s) for k, s in enumerate((13, 14, 15))]).astype(np.float32)
    b = dcrf.DenseCRF(W * H, L)
    b.setUnaryEnergy(U)
    b.addPairwiseEnergy(fg, 2.0)
    b.addPairwiseEnergy(fb, 4.0, kernel=dcrf.FULL_KERNEL)
    obj = dcrf.LogLikelihood(gt)
    ga, gb = a.gradient(3, obj), b.gradient(3, obj)
    assert ga[0] == gb[0] and all(np.array_equal(x, y) for x, y in zip(ga[1:], gb[1:]))
    assert ga[1].size == 0 and ga[2].size == 2 and ga[3].size == 2 + 25
    only_k = a.gradient(3, obj, unary=False, label_compatibility=False)
    assert only_k[1].size == 0 and only_k[2].size == 0 and np.array_equal(only_k[3], ga[3])
    res = oracle_built.GradientRestated(W * H, L)
    res.setUnaryEnergy(U)
    res.addPairwiseEnergy(fg, 2.0)
    res.addPairwiseEnergy(fb, 4.0, kernel=2)
    close(ga, res.gradient(3, 0, gt), "2d builders")
    with pytest.raises(ValueError):
        a.setKernelParameters(np.ones(3, np.float32))
    with pytest.raises(ValueError):
        a.gradient(1, dcrf.LogLikelihood(gt[:-1]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["potts_sym_ll", "two_terms_hamming", "matrix_full_iou", "logistic_before", "after_matrix"])
def test_engine_gradient_is_the_derivative_of_the_objective(dcrf, name):
    """Size-independent property (what the reference's gradCheck, optimization.cpp:96-118, tests): the unary and the
    label-compatibility gradients are the exact derivatives of the objective after n mean-field iterations, so a central
    difference of gradient()[0] over the parameter with the largest gradient reproduces that entry (2 %; the reference
    itself shows the same agreement).  The kernel gradient is an approximation by construction (the lattice is piecewise
    constant in the features) and is not part of this check."""
    case = [c for c in gc.CASES if c[0] == name][0]
    c = gc.make(case)
    m = dcrf.DenseCRF(c["N"], c["L"])
    gc.build(m, c, logistic=lambda Lm, f: m.setUnary(dcrf.LogisticUnary(Lm, f)))
    cw = dcrf.Hamming(c["gt"], c["objective"][1]).class_weights if c["objective"][0] == "hamming" else ()
    obj = engine_objective(dcrf, c, cw)
    r = m.gradient(c["n_iter"], obj)
    for getter, setter, g in ((m.unaryParameters, m.setUnaryParameters, r[1]),
                              (m.labelCompatibilityParameters, m.setLabelCompatibilityParameters, r[2])):
        if not g.size:
            continue
        p0 = getter()
        j = int(np.argmax(np.abs(g)))
        eps = 1e-2 * max(abs(float(p0[j])), 0.1)
        vals = []
        for sgn in (1, -1):
            p = p0.copy()
            p[j] += sgn * eps
            setter(p)
            vals.append(m.gradient(c["n_iter"], obj, unary=False, label_compatibility=False, kernel=False)[0])
        setter(p0)
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - g[j]) <= 0.02 * abs(g[j]), (name, j, float(g[j]), fd)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", FUZZ)
def test_engine_gradient_fuzz(dcrf, oracle_built, seed):
    """Random small models (tests/gradient_cases.py:fuzz_case) against the unmodified reference where it is present
    (oracle/_ref travels to the GPU box), else against the numpy restatement."""
    c = gc.make(gc.fuzz_case(seed))
    m = dcrf.DenseCRF(c["N"], c["L"])
    gc.build(m, c, logistic=lambda Lm, f: m.setUnary(dcrf.LogisticUnary(Lm, f)))
    kind = KIND[c["objective"][0]]
    robust = c["objective"][1] if kind == 0 else 0.0
    cw = dcrf.Hamming(c["gt"], c["objective"][1]).class_weights if kind == 1 else ()
    if oracle_built.available("ref"):
        chk = gc.build(oracle_built.CRF(c["N"], c["L"], "ref"), c)
    else:
        chk = gc.build(oracle_built.GradientRestated(c["N"], c["L"]), c)
    want = chk.gradient(c["n_iter"], kind, c["gt"], robust, cw)
    close(m.gradient(c["n_iter"], engine_objective(dcrf, c, cw)), want, c["name"])
