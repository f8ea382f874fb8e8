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
            sp = C.POINTER(C.c_short)
            self.crf_num_params = fn("crf_num_params", C.c_int, vp, C.c_int)
            self.crf_get_params = fn("crf_get_params", None, vp, C.c_int, _fp)
            self.crf_set_params = fn("crf_set_params", None, vp, C.c_int, _fp, C.c_int)
            self.crf_gradient = fn("crf_gradient", C.c_double, vp, C.c_int, C.c_int, sp, C.c_float, _fp, C.c_int, _fp, _fp, _fp)
            self.hamming_weights = fn("hamming_weights", None, sp, C.c_int, C.c_float, C.c_int, _fp)
            self.crf_unary_energy = fn("crf_unary_energy", None, vp, sp, _fp)
            self.crf_pairwise_energy = fn("crf_pairwise_energy", None, vp, sp, C.c_int, _fp)


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

    def setUnaryEnergy(self, u, f=None):
        if f is not None:
            return self.setUnaryLogistic(u, f)
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


# ---- learning path -------------------------------------------------------------------------------------
LOGLIKELIHOOD, HAMMING, IOU = 0, 1, 2
PARAMS_UNARY, PARAMS_LABEL_COMPATIBILITY, PARAMS_KERNEL = 0, 1, 2


def _ref_params(self, group):
    """unaryParameters / labelCompatibilityParameters / kernelParameters of the UNMODIFIED reference (which="ref" only)."""
    out = np.zeros(self.lib.crf_num_params(self.h, group), np.float32)
    if out.size:
        self.lib.crf_get_params(self.h, group, _ptr(out))
    return out


def _ref_set_params(self, group, v):
    v = _f32(v).ravel()
    self.lib.crf_set_params(self.h, group, _ptr(v), v.size)


def _ref_gradient(self, n_iter, kind, gt, robust=0.0, class_weight=(), unary=True, label_compatibility=True, kernel=True):
    """DenseCRF::gradient of the UNMODIFIED reference (which="ref" only): (value, unary_grad, lbl_cmp_grad, kernel_grad)."""
    gt = np.ascontiguousarray(gt, dtype=np.int16)
    cw = _f32(class_weight).ravel()
    outs = [np.zeros(self.lib.crf_num_params(self.h, g) if want else 0, np.float32)
            for g, want in ((0, unary), (1, label_compatibility), (2, kernel))]
    ptrs = [_ptr(o) if o.size else None for o in outs]
    r = self.lib.crf_gradient(self.h, int(n_iter), int(kind), _ptr(gt, C.POINTER(C.c_short)), float(robust),
                              _ptr(cw) if cw.size else None, cw.size, *ptrs)
    return (float(r),) + tuple(outs)


def _ref_unary_energy(self, labels):
    """DenseCRF::unaryEnergy of the UNMODIFIED reference (which="ref" only)."""
    lab = np.ascontiguousarray(labels, dtype=np.int16).ravel()
    out = np.zeros(self.N, np.float32)
    self.lib.crf_unary_energy(self.h, _ptr(lab, C.POINTER(C.c_short)), _ptr(out))
    return out


def _ref_pairwise_energy(self, labels, term=-1):
    """DenseCRF::pairwiseEnergy of the UNMODIFIED reference (which="ref" only)."""
    lab = np.ascontiguousarray(labels, dtype=np.int16).ravel()
    out = np.zeros(self.N, np.float32)
    self.lib.crf_pairwise_energy(self.h, _ptr(lab, C.POINTER(C.c_short)), int(term), _ptr(out))
    return out


CRF.unaryEnergy = _ref_unary_energy
CRF.pairwiseEnergy = _ref_pairwise_energy
CRF.parameters = _ref_params
CRF.setParameters = _ref_set_params
CRF.gradient = _ref_gradient


def hamming_weights(gt, power, L):
    """class weights of the reference's Hamming(gt, class_weight_pow) (objective.cpp:51-63), from libref."""
    gt = np.ascontiguousarray(gt, dtype=np.int16)
    out = np.zeros(L, np.float32)
    load("ref").hamming_weights(_ptr(gt, C.POINTER(C.c_short)), gt.size, float(power), L, _ptr(out))
    return out


class GradientRestated:
    """numpy restatement of DenseCRF::gradient (src/densecrf.cpp:237-297) and of what it calls -- the objectives
    (src/objective.cpp:35-109), sumAndNormalize (densecrf.cpp:107-114), PairwisePotential::gradient / kernelGradient /
    applyTranspose (src/pairwise.cpp:179-207), DenseKernel::featureGradient / gradient / setParameters (:87-116, :139-163),
    the compatibilities' gradients (src/labelcompatibility.cpp:56-108) and LogisticUnaryEnergy::gradient (src/unary.cpp:67-71)
    -- on top of the C restatement's lattice (Lattice(.., "orc"): compute, compute(reverse), gradient).  Matrices are
    (N, L) pixel-major float32 arrays, i.e. the memory of the reference's MatrixXf(L, N).  Sums over the variables are
    numpy float32 reductions (pairwise), the reference's are Eigen's: equal to ~1e-6 relative per sum.
    Pinned against the unmodified reference in tests/test_oracle_vs_ref.py."""

    def __init__(self, N, L):
        self.N, self.L = N, L
        self.unary = None       # ("const", U (N, L)) or ("logistic", Lm (L, F), f (F, N))
        self.terms = []

    def setUnaryEnergy(self, u_LxN, f=None):
        if f is None:
            self.unary = ("const", _f32(np.asarray(u_LxN).T))
        else:
            self.unary = ("logistic", _f32(u_LxN), _f32(f))

    def addPairwiseEnergy(self, feats, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        ck, cw = _compat(compat, self.L)
        if ck == MATRIX:
            cw = (0.5 * (cw + cw.T)).astype(np.float32)
        t = dict(f=_f32(feats), ck=ck, cw=cw, ktype=int(kernel), ntype=int(normalization))
        self._init_lattice(t, t["f"])
        self.terms.append(t)

    def _init_lattice(self, t, f):  # DenseKernel::initLattice (pairwise.cpp:40-62)
        t["lat"] = Lattice(f, "orc")
        nrm = t["lat"].filter(np.ones((self.N, 1), np.float32))[:, 0]
        if t["ntype"] == NO_NORMALIZATION:
            mean = np.float32(0)
            for v in nrm:
                mean = np.float32(mean + v)
            nrm = np.full(self.N, np.float32(self.N) / mean, np.float32)
        elif t["ntype"] == NORMALIZE_SYMMETRIC:
            nrm = (1.0 / np.sqrt(nrm.astype(np.float64) + 1e-20)).astype(np.float32)
        else:
            nrm = (1.0 / (nrm.astype(np.float64) + 1e-20)).astype(np.float32)
        t["norm"] = nrm

    # parameter vectors --------------------------------------------------------------------------------------------
    def parameters(self, group):
        if group == PARAMS_UNARY:
            return self.unary[1].T.ravel().copy() if self.unary and self.unary[0] == "logistic" else np.zeros(0, np.float32)
        out = []
        for t in self.terms:
            if group == PARAMS_LABEL_COMPATIBILITY:
                out.append(t["cw"][np.triu_indices(self.L)] if t["ck"] == MATRIX else t["cw"])
            else:
                d = t["f"].shape[0]
                if t["ktype"] == DIAG_KERNEL:
                    out.append(t.get("kp", np.ones(d, np.float32)))
                elif t["ktype"] == FULL_KERNEL:
                    out.append(t.get("kp", np.eye(d, dtype=np.float32).ravel()))
        return np.concatenate(out).astype(np.float32) if out else np.zeros(0, np.float32)

    def setParameters(self, group, v):
        v = _f32(v).ravel()
        if group == PARAMS_UNARY:
            Lm = self.unary[1]
            self.unary = ("logistic", np.ascontiguousarray(v.reshape(Lm.shape[1], Lm.shape[0]).T), self.unary[2])
            return
        i = 0
        for t in self.terms:
            if group == PARAMS_LABEL_COMPATIBILITY:
                if t["ck"] == MATRIX:
                    n = self.L * (self.L + 1) // 2
                    w = np.zeros((self.L, self.L), np.float32)
                    w[np.triu_indices(self.L)] = v[i:i + n]
                    t["cw"] = np.where(np.tri(self.L, k=-1, dtype=bool), w.T, w)
                else:
                    n = t["cw"].size
                    t["cw"] = v[i:i + n].copy()
                i += n
            else:
                d = t["f"].shape[0]
                if t["ktype"] == DIAG_KERNEL:
                    t["kp"] = v[i:i + d].copy()
                    self._init_lattice(t, t["kp"][:, None] * t["f"])
                    i += d
                elif t["ktype"] == FULL_KERNEL:
                    t["kp"] = v[i:i + d * d].copy()
                    P = t["kp"].reshape(d, d).T          # column-major resize
                    tf = np.zeros_like(t["f"])
                    for j in range(d):                   # the depth loop of the product, ascending, from zero
                        tf = (tf + P[:, j:j + 1] * t["f"][j:j + 1]).astype(np.float32)
                    self._init_lattice(t, tf)
                    i += d * d

    # operators ----------------------------------------------------------------------------------------------------
    def _kernel_apply(self, t, x, transpose=False):  # DenseKernel::filter (pairwise.cpp:63-80)
        nt = t["ntype"]
        if nt == NORMALIZE_SYMMETRIC or nt == (NORMALIZE_AFTER if transpose else NORMALIZE_BEFORE):
            x = x * t["norm"][:, None]
        x = t["lat"].filter(x, reverse=transpose)
        if nt == NORMALIZE_SYMMETRIC or nt == (NORMALIZE_BEFORE if transpose else NORMALIZE_AFTER):
            x = x * t["norm"][:, None]
        return x

    def _compat_apply(self, t, x):  # labelcompatibility.cpp:46-48, :66-69, :82-87 (W symmetric: transpose is the same)
        if t["ck"] == POTTS:
            return (-t["cw"][0]) * x
        if t["ck"] == DIAG:
            return x * t["cw"][None, :]
        return (x @ t["cw"].T).astype(np.float32)

    def _compat_gradient(self, t, b, q):
        if t["ck"] == POTTS:
            return np.array([-(b * q).sum(dtype=np.float32)], np.float32)
        if t["ck"] == DIAG:
            return (b * q).sum(axis=0, dtype=np.float32)
        g = (b.T @ q).astype(np.float32)   # g = b * Q^T, L x L
        iu = np.triu_indices(self.L)
        return (g[iu] + np.where(iu[0] != iu[1], g.T[iu], np.float32(0))).astype(np.float32)

    def _feature_gradient(self, t, a, b):  # pairwise.cpp:87-116; (N, d)
        lat, nt, nrm = t["lat"], t["ntype"], t["norm"][:, None]
        ones = np.ones_like(a)
        if nt == NO_NORMALIZATION:
            return lat.gradient(a, b)
        if nt == NORMALIZE_SYMMETRIC:
            fa = lat.filter(a * nrm, reverse=True)
            fb = lat.filter(b * nrm)
            r = lat.gradient((np.float32(0.5) * (a * fb + fa * b)) * (nrm * nrm * nrm), ones)
            return -r + lat.gradient(a * nrm, b * nrm)
        if nt == NORMALIZE_AFTER:
            fb = lat.filter(b)
            r = lat.gradient((a * fb) * (nrm * nrm), ones)
            return -r + lat.gradient(a * nrm, b)
        fa = lat.filter(a, reverse=True)
        r = lat.gradient((fa * b) * (nrm * nrm), ones)
        return -r + lat.gradient(a, b * nrm)

    def _kernel_gradient(self, t, a, b):  # DenseKernel::gradient (pairwise.cpp:152-163)
        if t["ktype"] == CONST_KERNEL:
            return np.zeros(0, np.float32)
        fg = self._feature_gradient(t, a, b)            # (N, d)
        if t["ktype"] == DIAG_KERNEL:
            return (t["f"].T * fg).sum(axis=0, dtype=np.float32)
        return (fg.T @ t["f"].T).astype(np.float32).T.ravel()   # p = fg * f^T (d x d), column-major

    @staticmethod
    def _sum_and_normalize(x, q):
        return x.sum(axis=1, dtype=np.float32)[:, None] * q - x

    def _unary_gradient(self, b):
        if not self.unary or self.unary[0] != "logistic":
            return np.zeros(0, np.float32)
        return (b.T @ self.unary[2].T).astype(np.float32).T.ravel()   # g = b * f^T (L x F), column-major

    def _objective(self, kind, gt, robust, cw, Q):  # objective.cpp:37-109
        N, L = Q.shape
        d = np.zeros_like(Q)
        ok = (gt >= 0) & (gt < L)
        idx = np.nonzero(ok)[0]
        g = gt[idx].astype(np.int64)
        qg = Q[idx, g]
        if kind == LOGLIKELIHOOD:
            qq = np.maximum(qg + np.float32(robust), np.float32(1e-20))
            d[idx, g] = qg / qq / np.float32(N)
            return float((np.log(qq.astype(np.float64)) / N).sum()), d
        if kind == HAMMING:
            qq = _f32(cw)[g] * qg
            d[idx, g] = qq
            return float(qq.astype(np.float64).sum()), d
        inn = np.zeros(L)
        un = np.full(L, 1e-20)
        Qv = Q[idx].astype(np.float64)
        onehot = np.zeros((idx.size, L), bool)
        onehot[np.arange(idx.size), g] = True
        inn += np.where(onehot, Qv, 0).sum(axis=0)
        un += np.where(onehot, 1.0, Qv).sum(axis=0)
        d[idx] = np.where(onehot, Qv / (un * L), -Qv * inn / (un * un * L)).astype(np.float32)
        return float((inn / un).sum() / L), d

    def gradient(self, n_iter, kind, gt, robust=0.0, class_weight=()):
        """-> (value, unary_grad, lbl_cmp_grad, kernel_grad)"""
        N, L = self.N, self.L
        gt = np.asarray(gt, np.int16)
        if self.unary is None:
            U = np.zeros((N, L), np.float32)
        elif self.unary[0] == "const":
            U = self.unary[1]
        else:
            U = (self.unary[1] @ self.unary[2]).T.astype(np.float32)
        Q = [exp_and_normalize(-U)]
        for _ in range(n_iter):
            t1 = -U
            for t in self.terms:
                t1 = t1 - self._compat_apply(t, self._kernel_apply(t, Q[-1]))
            Q.append(exp_and_normalize(t1))
        r, b = self._objective(kind, gt, robust, class_weight, Q[-1])
        b = self._sum_and_normalize(b, Q[-1])
        ug = self._unary_gradient(b)
        lg = 0 * self.parameters(PARAMS_LABEL_COMPATIBILITY)
        kg = 0 * self.parameters(PARAMS_KERNEL)
        for it in range(n_iter - 1, -1, -1):
            t1 = np.zeros((N, L), np.float32)
            ip = ik = 0
            for t in self.terms:
                pg = self._compat_gradient(t, b, self._kernel_apply(t, Q[it]))
                lg[ip:ip + pg.size] += pg
                ip += pg.size
                pk = self._kernel_gradient(t, b, self._compat_apply(t, Q[it]))
                kg[ik:ik + pk.size] += pk
                ik += pk.size
                t1 = t1 + self._compat_apply(t, self._kernel_apply(t, b, transpose=True))
            b = self._sum_and_normalize(t1 * Q[it], Q[it])
            ug = ug + self._unary_gradient(b)
        return r, ug, lg, kg
