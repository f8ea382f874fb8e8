"""ctypes front-end for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``load("orc")`` -> oracle/liborc.so, the C restatement (oracle/dcrf_oracle.c); always buildable.
* ``load("ref")`` -> oracle/_ref/libref.so, the UNMODIFIED reference C++ behind oracle/ref_shim.cpp;
  present only where oracle/Makefile could see /root/reference (or the prebuilt file travelled).

Both export the same flat interface (prefix ``orc_`` / ``ref_``), so every test can run against
either.  Nothing under pydensecrf_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
POTTS, DIAG, MATRIX = 0, 1, 2
NO_NORMALIZATION, NORMALIZE_BEFORE, NORMALIZE_AFTER, NORMALIZE_SYMMETRIC = 0, 1, 2, 3
CONST_KERNEL, DIAG_KERNEL, FULL_KERNEL = 0, 1, 2

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def build(verbose=False):
    """Compile liborc.so (always) and _ref/libref.so (when /root/reference is visible)."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout, r.stderr)
    if r.returncode:
        raise RuntimeError("oracle build failed")


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t=_fp):
    return a.ctypes.data_as(t)


def _compat(compat, L):
    if np.isscalar(compat):
        return POTTS, np.array([compat], np.float32)
    c = _f32(compat)
    if c.ndim == 1:
        assert c.shape[0] == L
        return DIAG, c
    assert c.shape == (L, L)
    return MATRIX, c


def available(which):
    p = os.path.join(HERE, "liborc.so" if which == "orc" else os.path.join("_ref", "libref.so"))
    return os.path.exists(p)


class _Lib:
    def __init__(self, which):
        self.which = which
        path = os.path.join(HERE, "liborc.so" if which == "orc" else os.path.join("_ref", "libref.so"))
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.dll = C.CDLL(path)
        p = which + "_"
        d = self.dll
        self.cores = 1  # both checkers are single-threaded, like the reference

        def fn(name, res, *args):
            f = getattr(d, p + name)
            f.restype = res
            f.argtypes = list(args)
            return f

        vp = C.c_void_p
        self.lattice_build = fn("lattice_build", vp, _fp, C.c_int, C.c_int)
        self.lattice_free = fn("lattice_free", None, vp)
        self.lattice_nvert = fn("lattice_nvert", C.c_int, vp)
        self.lattice_dump = fn("lattice_dump", None, vp, _ip, _fp, _ip, _ip)
        self.lattice_filter = fn("lattice_filter", None, vp, _fp, _fp, C.c_int, C.c_int)
        self.lattice_gradient = fn("lattice_gradient", None, vp, _fp, _fp, _fp, C.c_int)
        self.exp_and_normalize = fn("exp_and_normalize", None, _fp, _fp, C.c_int, C.c_int)
        self.crf_new = fn("crf_new", vp, C.c_int, C.c_int)
        self.crf_new_2d = fn("crf_new_2d", vp, C.c_int, C.c_int, C.c_int)
        self.crf_free = fn("crf_free", None, vp)
        self.crf_set_unary = fn("crf_set_unary", None, vp, _fp)
        self.crf_set_unary_logistic = fn("crf_set_unary_logistic", None, vp, _fp, C.c_int, _fp)
        self.crf_add_pairwise = fn("crf_add_pairwise", None, vp, _fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int)
        self.crf_add_gaussian = fn("crf_add_gaussian", C.c_int, vp, C.c_float, C.c_float, C.c_int, _fp, C.c_int, C.c_int)
        self.crf_add_bilateral = fn("crf_add_bilateral", C.c_int, vp, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.c_float, C.POINTER(C.c_uint8), C.c_int, _fp, C.c_int, C.c_int)
        self.crf_inference = fn("crf_inference", None, vp, C.c_int, _fp)
        self.crf_start = fn("crf_start", None, vp, _fp)
        self.crf_step = fn("crf_step", None, vp, _fp)
        self.crf_kl = fn("crf_kl", C.c_double, vp, _fp)
        self.crf_persist_start = fn("crf_persist_start", None, vp)
        self.crf_persist_step = fn("crf_persist_step", None, vp)
        self.crf_persist_get = fn("crf_persist_get", None, vp, _fp)
        if which == "orc":
            self.lattice_keys = fn("lattice_keys", None, vp, C.POINTER(C.c_short))
            self.crf_term_apply = fn("crf_term_apply", None, vp, C.c_int, _fp, _fp)
            self.crf_term_norm = fn("crf_term_norm", None, vp, C.c_int, _fp)
        else:
            self.pairwise_apply = fn("pairwise_apply", None, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp,
                                     C.c_int, C.c_int, _fp, _fp)


_libs = {}


def load(which="orc"):
    if which not in _libs:
        _libs[which] = _Lib(which)
    return _libs[which]


class Lattice:
    """Permutohedral lattice of a (d, N) float32 feature matrix."""

    def __init__(self, feats, which="orc"):
        self.lib = load(which)
        feats = _f32(feats)
        self.d, self.N = feats.shape
        self.h = self.lib.lattice_build(_ptr(feats), self.d, self.N)
        if not self.h:
            raise ValueError("lattice_build failed")
        self.V = self.lib.lattice_nvert(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.lattice_free(self.h)
            self.h = None

    def dump(self):
        """offset (N,d+1) int32, bary (N,d+1) f32, rank (N,d+1) int32, nbr (d+1,V,2) int32."""
        n, dp = self.N, self.d + 1
        off = np.zeros((n, dp), np.int32)
        bary = np.zeros((n, dp), np.float32)
        rank = np.zeros((n, dp), np.int32)
        nbr = np.zeros((dp, max(self.V, 1), 2), np.int32)
        self.lib.lattice_dump(self.h, _ptr(off, _ip), _ptr(bary), _ptr(rank, _ip), _ptr(nbr, _ip))
        return off, bary, rank, nbr[:, : self.V]

    def keys(self):
        k = np.zeros((max(self.V, 1), self.d), np.int16)
        self.lib.lattice_keys(self.h, _ptr(k, C.POINTER(C.c_short)))
        return k[: self.V]

    def filter(self, x_NxL, reverse=False):
        """x: (N, L) pixel-major -> filtered (N, L)."""
        x = _f32(x_NxL)
        assert x.shape[0] == self.N
        out = np.empty_like(x)
        self.lib.lattice_filter(self.h, _ptr(out), _ptr(x), x.shape[1], int(reverse))
        return out


def _lattice_gradient(self, a_NxL, b_NxL):
    """Permutohedral::gradient: (N, d) = d(a^T K b)/d(features) for (N, L) a, b."""
    a, b = _f32(a_NxL), _f32(b_NxL)
    assert a.shape == b.shape and a.shape[0] == self.N
    df = np.zeros((self.N, self.d), np.float32)
    self.lib.lattice_gradient(self.h, _ptr(df), _ptr(a), _ptr(b), a.shape[1])
    return df


Lattice.gradient = _lattice_gradient


def exp_and_normalize(x_NxL, which="orc"):
    x = _f32(x_NxL)
    out = np.empty_like(x)
    load(which).exp_and_normalize(_ptr(out), _ptr(x), x.shape[1], x.shape[0])
    return out


class CRF:
    """Mirror of pydensecrf's DenseCRF / DenseCRF2D on the CPU checker.  Q is returned as the
    (L, N) Fortran-ordered view pydensecrf's MatrixXf exposes (pixel-major memory)."""

    def __init__(self, n_or_wh, L, which="orc"):
        self.lib = load(which)
        self.L = L
        if isinstance(n_or_wh, (tuple, list)):
            self.W, self.H = n_or_wh
            self.N = self.W * self.H
            self.h = self.lib.crf_new_2d(self.W, self.H, L)
        else:
            self.N = int(n_or_wh)
            self.W = self.H = 0
            self.h = self.lib.crf_new(self.N, L)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.crf_free(self.h)
            self.h = None

    def setUnaryEnergy(self, u):
        u = _f32(u)
        assert u.shape == (self.L, self.N)
        self.lib.crf_set_unary(self.h, _ptr(u))

    def setUnaryLogistic(self, Lm, f):
        Lm, f = _f32(Lm), _f32(f)
        assert Lm.shape[0] == self.L and Lm.shape[1] == f.shape[0] and f.shape[1] == self.N
        self.lib.crf_set_unary_logistic(self.h, _ptr(Lm), Lm.shape[1], _ptr(f))

    def addPairwiseEnergy(self, feats, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        feats = _f32(feats)
        assert feats.shape[1] == self.N
        ck, cw = _compat(compat, self.L)
        self.lib.crf_add_pairwise(self.h, _ptr(feats), feats.shape[0], ck, _ptr(cw), int(kernel), int(normalization))

    def addPairwiseGaussian(self, sxy, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        if np.isscalar(sxy):
            sxy = (sxy, sxy)
        ck, cw = _compat(compat, self.L)
        assert self.lib.crf_add_gaussian(self.h, sxy[0], sxy[1], ck, _ptr(cw), int(kernel), int(normalization)) == 0

    def addPairwiseBilateral(self, sxy, srgb, rgbim, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        if np.isscalar(sxy):
            sxy = (sxy, sxy)
        if np.isscalar(srgb):
            srgb = (srgb, srgb, srgb)
        im = np.ascontiguousarray(rgbim, dtype=np.uint8)
        assert im.shape == (self.H, self.W, 3)
        ck, cw = _compat(compat, self.L)
        assert self.lib.crf_add_bilateral(self.h, sxy[0], sxy[1], srgb[0], srgb[1], srgb[2],
                                          _ptr(im, C.POINTER(C.c_uint8)), ck, _ptr(cw), int(kernel),
                                          int(normalization)) == 0

    def _q(self):
        return np.empty((self.N, self.L), np.float32)

    def inference(self, n):
        q = self._q()
        self.lib.crf_inference(self.h, int(n), _ptr(q))
        return q.T  # (L, N) view with strides (4, 4L), like pydensecrf's MatrixXf

    def startInference(self):
        q = self._q()
        self.lib.crf_start(self.h, _ptr(q))
        return q.T

    def stepInference(self, Q):
        q = np.ascontiguousarray(Q.T, dtype=np.float32)
        self.lib.crf_step(self.h, _ptr(q))
        Q[...] = q.T

    def step_raw(self, q_NxL):
        """stepInference in place on a C-contiguous (N, L) float32 array (no layout copies)."""
        assert q_NxL.flags.c_contiguous and q_NxL.dtype == np.float32 and q_NxL.shape == (self.N, self.L)
        self.lib.crf_step(self.h, _ptr(q_NxL))

    def persist_start(self):
        """Q, tmp1, tmp2 = startInference() kept inside the checker (no copies): the reference's stepwise protocol."""
        self.lib.crf_persist_start(self.h)

    def persist_step(self):
        """stepInference(Q, tmp1, tmp2) on the matrices persist_start created."""
        self.lib.crf_persist_step(self.h)

    def persist_q(self):
        q = self._q()
        self.lib.crf_persist_get(self.h, _ptr(q))
        return q.T

    def klDivergence(self, Q):
        q = np.ascontiguousarray(Q.T, dtype=np.float32)
        return float(self.lib.crf_kl(self.h, _ptr(q)))
