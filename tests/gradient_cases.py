"""Seeded models for the learning-path tests (DenseCRF::gradient): the same case can be built on the
unmodified reference (oracle CRF(..., "ref")), on the numpy restatement (oracle GradientRestated) and on the engine."""
import numpy as np

POTTS, DIAG, MATRIX = "potts", "diag", "matrix"

# name, N, L, unary ("const" | "logistic" | None), terms [(d, compat, ktype, ntype)], objective, n_iter
CASES = [
    ("potts_sym_ll", 400, 5, "const", [(2, POTTS, 1, 3)], ("ll", 0.0), 3),
    ("two_terms_hamming", 360, 4, "const", [(2, POTTS, 1, 3), (5, DIAG, 1, 3)], ("hamming", 0.5), 2),
    ("matrix_full_iou", 300, 6, "const", [(3, MATRIX, 2, 3)], ("iou",), 2),
    ("logistic_before", 256, 3, "logistic", [(2, DIAG, 1, 1)], ("ll", 0.1), 3),
    ("after_matrix", 333, 7, "const", [(4, MATRIX, 1, 2)], ("hamming", 0.0), 2),
    ("nonorm_const_kernel", 280, 4, "logistic", [(3, POTTS, 0, 0), (2, POTTS, 1, 0)], ("iou",), 2),
    ("labels21", 500, 21, "const", [(2, POTTS, 1, 3), (5, POTTS, 2, 3)], ("ll", 0.0), 2),
    ("two_labels", 200, 2, "const", [(2, DIAG, 1, 3)], ("ll", 0.0), 4),
    ("labels70_matrix", 320, 70, "const", [(3, MATRIX, 1, 3)], ("ll", 0.0), 2),   # tensor-core MatrixCompatibility (>= 64 labels), 4 900 sums
    ("no_terms", 128, 3, "logistic", [], ("hamming", 1.0), 2),
    ("zero_iterations", 150, 4, "logistic", [(2, POTTS, 1, 3)], ("iou",), 0),
]


def make(case):
    name, N, L, unary, terms, objective, n_iter = case
    rng = np.random.default_rng(sum(map(ord, name)))
    out = dict(name=name, N=N, L=L, n_iter=n_iter, unary=None, terms=[])
    if unary == "const":
        out["unary"] = ("const", (rng.standard_normal((L, N)) * 1.5).astype(np.float32))
    elif unary == "logistic":
        F = 4
        out["unary"] = ("logistic", rng.standard_normal((L, F)).astype(np.float32), rng.standard_normal((F, N)).astype(np.float32))
    for d, ck, ktype, ntype in terms:
        # smooth, well separated features: a grid walk plus noise, scaled so that vertices are shared by several variables
        f = (np.cumsum(rng.standard_normal((d, N)) * 0.3, axis=1) + rng.standard_normal((d, N)) * 0.2).astype(np.float32)
        if ck == POTTS:
            compat = float(rng.uniform(0.5, 3))
        elif ck == DIAG:
            compat = rng.uniform(0.2, 2, L).astype(np.float32)
        else:
            compat = rng.uniform(-1, 1, (L, L)).astype(np.float32)
        out["terms"].append((f, compat, ktype, ntype))
    gt = rng.integers(0, L, N).astype(np.int16)
    gt[rng.random(N) < 0.1] = -1          # unlabeled variables
    if objective[0] != "hamming":         # (Hamming(gt, pow) sizes its class table by max(gt): absent classes give it NaN weights)
        gt[rng.random(N) < 0.02] = L + 3  # out of range: ignored like unlabeled ones
    out["gt"] = gt
    out["objective"] = objective
    return out


def build(model, c, logistic=None):
    """Apply the case to any object with pydensecrf's setUnaryEnergy / addPairwiseEnergy.  `logistic(L, f)` sets a
    logistic unary where setUnaryEnergy(L, f) cannot (pydensecrf's wrapper insists on an (nlabels, nvar) first argument,
    densecrf.pyx:80-89: there the way in is setUnary(LogisticUnary(L, f)))."""
    u = c["unary"]
    if u is not None:
        if u[0] == "const":
            model.setUnaryEnergy(u[1])
        elif logistic is not None:
            logistic(u[1], u[2])
        else:
            model.setUnaryEnergy(u[1], u[2])
    for f, compat, ktype, ntype in c["terms"]:
        model.addPairwiseEnergy(f, compat, kernel=ktype, normalization=ntype)
    return model


def kernel_param_update(c, rng):
    """A kernel parameter vector near the default (ones / identity) for the case's terms."""
    out = []
    for f, _, ktype, _ in c["terms"]:
        d = f.shape[0]
        if ktype == 1:
            out.append(rng.uniform(0.8, 1.25, d))
        elif ktype == 2:
            out.append((np.eye(d) + rng.uniform(-0.15, 0.15, (d, d))).T.ravel())
    return np.concatenate(out).astype(np.float32) if out else np.zeros(0, np.float32)


def fuzz_case(seed):
    """A random small model for the fuzz tests: any unary kind, 0-3 terms of random feature dimension, compatibility,
    kernel and normalisation type, any objective, 0-3 iterations."""
    rng = np.random.default_rng(1000 + seed)
    N = int(rng.integers(40, 500))
    L = int(rng.choice([2, 3, 4, 5, 8, 13, 21, 24]))
    unary = [None, "const", "logistic"][int(rng.integers(0, 3))]
    terms = []
    for _ in range(int(rng.integers(0, 4))):
        terms.append((int(rng.integers(1, 7)), [POTTS, DIAG, MATRIX][int(rng.integers(0, 3))], int(rng.integers(0, 3)),
                      int(rng.integers(0, 4))))
    kind = ["ll", "hamming", "iou"][int(rng.integers(0, 3))]
    objective = (kind, float(rng.choice([0.0, 0.05]))) if kind == "ll" else (kind, 0.0) if kind == "hamming" else (kind,)
    return ("fuzz%d" % seed, N, L, unary, terms, objective, int(rng.integers(0, 4)))
