"""Drop-in checks on the GPU box (python -m pytest tests -m gpu):

* the reference's OWN scripts run unmodified against this engine through the `pydensecrf` alias package: its
  examples/inference.py on im1.png / anno1.png (BASELINE configs[0]; the label map it writes is compared with the MAP of
  the reference's Q for that input, tests/golden/cfg1_im1.npz) and its pydensecrf/tests/test_dcrf.py + test_eigen.py.
  The scripts are not part of this repository: they are read from /root/reference where it exists, else from
  oracle/_ref/scripts/, where oracle/Makefile stages them untracked next to libref.so so that they travel to the GPU box;
* the C-ABI driven with ctypes only (no Cython): create -> unary -> two pairwise terms -> inference -> host pointer,
  against the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def _ref_file(*rel_candidates):
    for base, rel in rel_candidates:
        p = os.path.join(base, rel)
        if os.path.exists(p):
            return p
    pytest.fail("reference script not found: neither /root/reference nor oracle/_ref/scripts (run `make -C oracle all` "
                "where /root/reference is visible)")


def _env():
    e = dict(os.environ)
    e["PYTHONPATH"] = ROOT + os.pathsep + e.get("PYTHONPATH", "")
    return e


def test_reference_example_script_runs_unmodified(tmp_path):
    cv2 = pytest.importorskip("cv2")
    staged = os.path.join(ROOT, "oracle", "_ref", "scripts")
    script = _ref_file(("/root/reference/examples", "inference.py"), (staged, "inference.py"))
    im = _ref_file(("/root/reference/examples", "im1.png"), (staged, "im1.png"))
    anno = _ref_file(("/root/reference/examples", "anno1.png"), (staged, "anno1.png"))
    out = str(tmp_path / "out1.png")
    r = subprocess.run([sys.executable, script, im, anno, out], capture_output=True, text=True, env=_env(), timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Using generic 2D functions" in r.stdout
    kls = [float(line.split(":")[1]) for line in r.stdout.splitlines() if line.startswith("KL-divergence at")]
    assert len(kls) == 5 and all(b < a for a, b in zip(kls, kls[1:])), kls   # mean field decreases the KL divergence
    # the label map the script wrote vs the MAP of the reference's own Q (generic branch) for this input
    g = np.load(os.path.join(G, "cfg1_im1.npz"))
    Qref = g["ref_Q_generic"]
    anno_rgb = cv2.imread(anno).astype(np.uint32)
    lbl = anno_rgb[:, :, 0] + (anno_rgb[:, :, 1] << 8) + (anno_rgb[:, :, 2] << 16)
    colors = np.unique(lbl)
    colors = colors[1:] if 0 in colors else colors
    colorize = np.stack([colors & 0xFF, (colors & 0xFF00) >> 8, (colors & 0xFF0000) >> 16], 1).astype(np.uint8)
    want = colorize[np.argmax(Qref, axis=0)].reshape(anno_rgb.shape)
    got = cv2.imread(out)
    srt = np.sort(Qref, axis=0)
    sure = ((srt[-1] - srt[-2]) > 1e-3).reshape(anno_rgb.shape[:2])
    assert got.shape == want.shape
    assert np.array_equal(got[sure], want[sure])
    assert sure.mean() > 0.99


def test_reference_own_tests_pass_unmodified():
    staged = os.path.join(ROOT, "oracle", "_ref", "scripts")
    t1 = _ref_file(("/root/reference/pydensecrf/tests", "test_dcrf.py"), (staged, "test_dcrf.py"))
    t2 = _ref_file(("/root/reference/pydensecrf/tests", "test_eigen.py"), (staged, "test_eigen.py"))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", os.path.dirname(t1), t1, t2],
                       capture_output=True, text=True, env=_env(), timeout=300, cwd=os.path.dirname(t1))
    assert r.returncode == 0, r.stdout + r.stderr
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_c_abi_with_ctypes_only(oracle_built):
    """No Cython in the loop: what a maintainer binding libdcrf_b200.so from another language would call."""
    from tests import workloads as wl

    lib = C.CDLL(os.path.join(ROOT, "pydensecrf_b200", "libdcrf_b200.so"))
    vp, fp = C.c_void_p, C.POINTER(C.c_float)
    lib.dcrf_last_error.restype = C.c_char_p
    lib.dcrf_create_2d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.dcrf_set_unary_const.argtypes = [vp, fp]
    lib.dcrf_add_pairwise_gaussian.argtypes = [vp, C.c_float, C.c_float, C.c_int, fp, C.c_int, C.c_int]
    lib.dcrf_add_pairwise_bilateral.argtypes = [vp] + [C.c_float] * 5 + [C.POINTER(C.c_uint8), C.c_int, fp, C.c_int, C.c_int]
    lib.dcrf_inference.argtypes = [vp, C.c_int, C.POINTER(vp)]
    lib.dcrf_map.argtypes = [vp, C.c_int, C.POINTER(C.c_int16)]
    lib.dcrf_mat_host_ptr.argtypes = [vp, C.POINTER(fp)]
    lib.dcrf_mat_rows.argtypes = lib.dcrf_mat_cols.argtypes = [vp]
    lib.dcrf_mat_release_host.argtypes = lib.dcrf_mat_destroy.argtypes = lib.dcrf_destroy.argtypes = [vp]
    lib.dcrf_mat_release_host.restype = lib.dcrf_mat_destroy.restype = lib.dcrf_destroy.restype = None

    def ok(st):
        assert st == 0, lib.dcrf_last_error().decode()

    W, H, L = 53, 37, 21
    img = wl.smooth_image(H, W, seed=4, grid=8)
    U = wl.softmax_unary(H, W, L, seed=4, grid=8)
    m, Q = vp(), vp()
    ok(lib.dcrf_create_2d(W, H, L, 0, C.byref(m)))
    ok(lib.dcrf_set_unary_const(m, U.ctypes.data_as(fp)))
    w_g, w_b = (C.c_float * 1)(3.0), (C.c_float * 1)(10.0)
    ok(lib.dcrf_add_pairwise_gaussian(m, 3.0, 3.0, 0, w_g, 1, 3))
    ok(lib.dcrf_add_pairwise_bilateral(m, 80.0, 80.0, 13.0, 13.0, 13.0, img.ctypes.data_as(C.POINTER(C.c_uint8)), 0, w_b, 1, 3))
    ok(lib.dcrf_inference(m, 5, C.byref(Q)))
    assert (lib.dcrf_mat_rows(Q), lib.dcrf_mat_cols(Q)) == (L, W * H)
    p = fp()
    ok(lib.dcrf_mat_host_ptr(Q, C.byref(p)))
    q = np.ctypeslib.as_array(p, shape=(W * H, L)).copy().T     # Eigen layout: N columns of L floats
    lib.dcrf_mat_release_host(Q)
    labels = np.empty(W * H, np.int16)
    ok(lib.dcrf_map(m, 5, labels.ctypes.data_as(C.POINTER(C.c_int16))))
    lib.dcrf_mat_destroy(Q)
    lib.dcrf_destroy(m)

    o = oracle_built.CRF((W, H), L, "orc")
    o.setUnaryEnergy(U)
    o.addPairwiseGaussian(3, 3.0)
    o.addPairwiseBilateral(80, 13, img, 10.0)
    Qo = o.inference(5)
    assert np.abs(q - Qo).max() <= 2e-6
    assert np.array_equal(labels, np.argmax(q, axis=0))


def test_c_abi_learning_path_with_ctypes_only(oracle_built):
    """dcrf_gradient, dcrf_{num,get,set}_parameters and dcrf_{unary,pairwise}_energy driven through ctypes alone (the
    dcrf_objective struct as a C structure), against the numpy restatement of the reference's gradient."""
    from tests import workloads as wl

    lib = C.CDLL(os.path.join(ROOT, "pydensecrf_b200", "libdcrf_b200.so"))
    vp, fp, sp = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int16)

    class Objective(C.Structure):   # include/dcrf_b200.h: dcrf_objective
        _fields_ = [("kind", C.c_int), ("gt", sp), ("robust", C.c_float), ("class_weight", fp), ("n_class_weight", C.c_int)]

    lib.dcrf_last_error.restype = C.c_char_p
    lib.dcrf_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.dcrf_set_unary_const.argtypes = [vp, fp]
    lib.dcrf_add_pairwise_energy.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int]
    lib.dcrf_num_parameters.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    lib.dcrf_get_parameters.argtypes = [vp, C.c_int, fp]
    lib.dcrf_set_parameters.argtypes = [vp, C.c_int, fp, C.c_int]
    lib.dcrf_gradient.argtypes = [vp, C.c_int, C.POINTER(Objective), C.POINTER(C.c_double), fp, fp, fp]
    lib.dcrf_unary_energy.argtypes = [vp, sp, fp]
    lib.dcrf_pairwise_energy.argtypes = [vp, sp, C.c_int, fp]
    lib.dcrf_destroy.argtypes = [vp]
    lib.dcrf_destroy.restype = None

    def ok(st):
        assert st == 0, lib.dcrf_last_error().decode()

    N, L, d = 300, 4, 3
    rng = np.random.default_rng(9)
    U = rng.standard_normal((L, N)).astype(np.float32)
    f = (np.cumsum(rng.standard_normal((d, N)) * 0.3, axis=1)).astype(np.float32)
    gt = rng.integers(0, L, N).astype(np.int16)
    m = vp()
    ok(lib.dcrf_create(N, L, 0, C.byref(m)))
    ok(lib.dcrf_set_unary_const(m, U.ctypes.data_as(fp)))
    w = (C.c_float * 1)(2.0)
    ok(lib.dcrf_add_pairwise_energy(m, f.ctypes.data_as(fp), d, 0, w, 1, 3))     # Potts, DIAG_KERNEL, NORMALIZE_SYMMETRIC
    n = [C.c_int() for _ in range(3)]
    for g in range(3):
        ok(lib.dcrf_num_parameters(m, g, C.byref(n[g])))
    assert [x.value for x in n] == [0, 1, d]
    kp = np.zeros(d, np.float32)
    ok(lib.dcrf_get_parameters(m, 2, kp.ctypes.data_as(fp)))
    assert np.array_equal(kp, np.ones(d, np.float32))
    obj = Objective(0, gt.ctypes.data_as(sp), 0.0, None, 0)
    val = C.c_double()
    gl, gk = np.zeros(1, np.float32), np.zeros(d, np.float32)
    ok(lib.dcrf_gradient(m, 3, C.byref(obj), C.byref(val), None, gl.ctypes.data_as(fp), gk.ctypes.data_as(fp)))
    r = oracle_built.GradientRestated(N, L)
    r.setUnaryEnergy(U)
    r.addPairwiseEnergy(f, 2.0)
    want = r.gradient(3, 0, gt)
    assert abs(val.value - want[0]) <= 1e-6 * abs(want[0])
    assert np.abs(gl - want[2]).max() <= 2e-4 * np.abs(want[2]).max() + 5e-7
    assert np.abs(gk - want[3]).max() <= 2e-4 * np.abs(want[3]).max() + 5e-7
    new = np.array([3.5], np.float32)
    ok(lib.dcrf_set_parameters(m, 1, new.ctypes.data_as(fp), 1))
    assert lib.dcrf_set_parameters(m, 1, new.ctypes.data_as(fp), 2) == 1          # DCRF_EINVAL: wrong count
    e_u, e_p = np.zeros(N, np.float32), np.zeros(N, np.float32)
    ok(lib.dcrf_unary_energy(m, gt.ctypes.data_as(sp), e_u.ctypes.data_as(fp)))
    ok(lib.dcrf_pairwise_energy(m, gt.ctypes.data_as(sp), -1, e_p.ctypes.data_as(fp)))
    assert np.array_equal(e_u, U[gt, np.arange(N)])
    onehot = np.eye(L, dtype=np.float32)[gt]                                      # (N, L)
    t = r.terms[0]
    assert np.allclose(e_p, -0.5 * (-3.5 * r._kernel_apply(t, onehot))[np.arange(N), gt], rtol=1e-6, atol=1e-7)
    lib.dcrf_destroy(m)
