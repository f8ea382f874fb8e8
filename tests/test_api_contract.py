"""API / error contract of the drop-in, following the reference's own tests
(pydensecrf/tests/test_dcrf.py, pydensecrf/tests/test_eigen.py, tests/test_utils.py, tests/issue29.py)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- CPU: eigen surface (reference: pydensecrf/tests/test_eigen.py:7-41) ----------------------------
def test_vector_conversion():
    import pydensecrf_b200.eigen as e

    v = np.random.randn(3).astype(np.float32)
    assert np.all(np.array(e.vectorXf(v)) == v)


def test_matrix_conversion():
    import pydensecrf_b200.eigen as e

    m = np.random.randn(3, 4).astype(np.float32)
    c = e.matrixXf(m)
    assert np.all(np.array(c) == m)
    view = np.asarray(c)
    assert view.shape == (3, 4) and view.strides == (4, 12) and view.flags.writeable  # column-major, like Eigen
    mv = memoryview(c)
    assert mv.format == "f" and mv.ndim == 2 and not mv.readonly


def test_wrong_dims():
    import pydensecrf_b200.eigen as e

    with pytest.raises(ValueError):
        e.matrixXf(np.random.randn(3, 3, 3).astype(np.float32))


def test_wrong_type():
    import pydensecrf_b200.eigen as e

    with pytest.raises(ValueError):
        e.matrixXf(np.random.randn(3, 3).astype(np.float64))


def test_none_type():
    import pydensecrf_b200.eigen as e

    with pytest.raises(TypeError):
        e.matrixXf(None)


def test_empty_matrix():
    import pydensecrf_b200.eigen as e

    assert np.array(e.MatrixXf()).shape == (0, 0)


# ---- CPU: utils (reference: tests/test_utils.py:10-23) -----------------------------------------------
def test_utils_unary():
    from pydensecrf_b200 import utils

    M = 3
    U, P, N = 1. / M, 0.8, 0.2 / (M - 1)
    labels = np.array([[0, 1, 2, 3], [3, 2, 1, 0]])
    unary = -np.log(np.array([[U, P, N, N, N, N, P, U], [U, N, P, N, N, P, N, U], [U, N, N, P, P, N, N, U]]))
    np.testing.assert_almost_equal(utils.compute_unary(labels, M, GT_PROB=P), unary)


def test_utils_features_match_2d_builders():
    from pydensecrf_b200 import utils

    img = np.random.default_rng(0).integers(0, 255, (7, 9, 3), dtype=np.uint8)
    f = utils.create_pairwise_bilateral(sdims=(3, 2), schan=(13, 11, 9), img=img, chdim=2)
    assert f.shape == (5, 63) and f.dtype == np.float32
    assert f[0, 9] == np.float32(1) / np.float32(3) and f[1, 1] == np.float32(0.5)
    assert f[2, 10] == np.float32(img[1, 1, 0]) / np.float32(13)
    g = utils.create_pairwise_gaussian(sdims=(3, 2), shape=(7, 9))
    assert np.array_equal(g, f[:2])
    v = utils.create_pairwise_bilateral(sdims=(3, 3, 3), schan=(13,), img=np.zeros((4, 5, 6), np.uint8), chdim=-1)
    assert v.shape == (4, 120)
    sm = np.array([[0.2, 1.0], [0.8, 0.0]])
    u = utils.unary_from_softmax(sm)
    assert u.dtype == np.float32 and np.isclose(u[1, 1], -np.log(1e-5))


# ---- CPU: the C-ABI library loads without a GPU and exports what the header declares -----------------
def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dcrf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(dcrf_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    lib = ctypes.CDLL(os.path.join(ROOT, "pydensecrf_b200", "libdcrf_b200.so"))
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    lib.dcrf_abi_version.restype = ctypes.c_int
    assert lib.dcrf_abi_version() == 1
    lib.dcrf_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.dcrf_last_error(), bytes)


def test_no_cpu_fallback_and_no_oracle_in_product():
    """The product must fail loudly without a device, and must not reference the oracle."""
    import pydensecrf_b200.densecrf as dcrf

    if dcrf.device_count() == 0:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            dcrf.DenseCRF2D(10, 10, 2)
    pkg = os.path.join(ROOT, "pydensecrf_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".pyx", ".cu", ".h", ".cpp")) and f != "_core.cpp":
                src = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in src and "liborc" not in src and "libref" not in src, f


def test_enum_spellings():
    import pydensecrf_b200.densecrf as dcrf

    assert dcrf.DIAG_KERNEL == dcrf.KernelType.DIAG_KERNEL == 1
    assert dcrf.NORMALIZE_SYMMETRIC == dcrf.NormalizationType.NORMALIZE_SYMMETRIC == 3
    assert [int(x) for x in dcrf.KernelType] == [0, 1, 2]
    assert [int(x) for x in dcrf.NormalizationType] == [0, 1, 2, 3]


# ---- GPU: reference's test_dcrf.py, ported --------------------------------------------------------------
def _simple_unary():
    u1 = np.zeros((10, 10), dtype=np.float32)
    u1[:, [0, -1]] = u1[[0, -1], :] = 1
    u2 = np.zeros((10, 10), dtype=np.float32)
    u2[4:7, 4:7] = 1
    u = np.vstack([u1.flat, u2.flat])
    return (u + 1) / (np.sum(u, axis=0) + 2)


def _simple_img():
    img = np.zeros((10, 10, 3), dtype=np.uint8)
    img[2:8, 2:8, :] = 255
    return img


@pytest.mark.gpu
def test_call_dcrf2d():
    import pydensecrf_b200.densecrf as dcrf

    d = dcrf.DenseCRF2D(10, 10, 2)
    d.setUnaryEnergy(-np.log(_simple_unary()))
    d.addPairwiseBilateral(sxy=2, srgb=2, rgbim=_simple_img(), compat=3)
    np.argmax(d.inference(10), axis=0).reshape(10, 10)


@pytest.mark.gpu
def test_call_dcrf_eq_dcrf2d():
    import pydensecrf_b200.densecrf as dcrf
    from pydensecrf_b200 import utils

    d = dcrf.DenseCRF(100, 2)
    d2 = dcrf.DenseCRF2D(10, 10, 2)
    unary, img = _simple_unary(), _simple_img()
    d.setUnaryEnergy(-np.log(unary))
    d2.setUnaryEnergy(-np.log(unary))
    d.addPairwiseEnergy(utils.create_pairwise_bilateral(sdims=(2, 2), schan=2, img=img, chdim=2), compat=3)
    d2.addPairwiseBilateral(sxy=2, srgb=2, rgbim=img, compat=3)
    res1 = np.argmax(d.inference(10), axis=0).reshape(10, 10)
    res2 = np.argmax(d2.inference(10), axis=0).reshape(10, 10)
    assert np.all(res1 == res2)


@pytest.mark.gpu
def test_compact_wrong():
    import pydensecrf_b200.densecrf as dcrf

    d = dcrf.DenseCRF2D(10, 10, 2)
    d.setUnaryEnergy(np.ones((2, 10 * 10), dtype=np.float32))
    with pytest.raises(ValueError):   # float64 compat: buffer dtype mismatch, not a crash
        d.addPairwiseBilateral(sxy=(3, 3), srgb=(3, 3, 3), rgbim=np.zeros((10, 10, 3), np.uint8), compat=np.array([1.0, 2.0]))
        d.inference(2)


@pytest.mark.gpu
def test_error_contract():
    import pydensecrf_b200.densecrf as dcrf
    import pydensecrf_b200.eigen as e

    d = dcrf.DenseCRF2D(12, 10, 3)
    with pytest.raises(ValueError, match="Bad shape for unary energy"):      # tests/issue29.py
        d.setUnaryEnergy(np.ones((3, 100), np.float32))
    with pytest.raises(ValueError):
        d.setUnaryEnergy(np.ones((3, 120), np.float64))
    with pytest.raises(TypeError):
        d.setUnaryEnergy(None)
    with pytest.raises(ValueError, match="Bad shape for pairwise bilateral"):
        d.addPairwiseBilateral(3, 3, np.zeros((12, 10, 3), np.uint8), 1)
    with pytest.raises(ValueError, match="only works for RGB"):
        d.addPairwiseBilateral(3, 3, np.zeros((10, 12, 4), np.uint8), 1)
    with pytest.raises(ValueError, match="Bad shape for pairwise energy"):
        d.addPairwiseEnergy(np.zeros((2, 100), np.float32), 1)
    with pytest.raises(ValueError, match="dimension >2"):
        d.addPairwiseGaussian(3, np.zeros((3, 3, 3), np.float32))
    with pytest.raises(ValueError):
        d.addPairwiseGaussian(3, 1.0, normalization=7)
    d.setUnaryEnergy(np.ones((3, 120), np.float32))
    d.addPairwiseGaussian(3, 1.0, kernel=dcrf.FULL_KERNEL, normalization=dcrf.NORMALIZE_BEFORE)
    with pytest.raises(ValueError):                                          # wrong-shaped Q
        d.klDivergence(e.matrixXf(np.ones((3, 7), np.float32)))
    with pytest.raises(TypeError):
        d.stepInference(np.ones((3, 120), np.float32), e.MatrixXf(), e.MatrixXf())
    Q = e.matrixXf(np.full((3, 120), 1 / 3, np.float32))                     # user-made Q goes in as well
    d.stepInference(Q, e.MatrixXf(), e.MatrixXf())
    assert np.allclose(np.array(Q).sum(0), 1, atol=1e-5)
    u = dcrf.ConstUnary(np.ones((3, 120), np.float32))
    d.setUnary(u)
    with pytest.raises(ValueError):
        d.setUnary(u)   # consumed, like Unary.move() in the reference
