"""Parameter learning with the engine's DenseCRF.gradient, the job of the reference's optimization.cpp + liblbfgs (out of
scope here) done by scipy's L-BFGS: a 3-class synthetic segmentation, a logistic unary on noisy features and two Potts
pairwise terms; the unary matrix and the two Potts weights are learnt by maximising the LogLikelihood of the ground truth
after 5 mean-field iterations.  `--oracle` runs the same script on the numpy restatement of the reference (CPU, small size)
instead of the engine."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.optimize import minimize

use_oracle = "--oracle" in sys.argv
W, H = (48, 36) if use_oracle else (320, 240)
L, F, N_ITER = 3, 4, 5
rng = np.random.default_rng(0)
yy, xx = np.mgrid[0:H, 0:W]
gt = np.zeros((H, W), np.int16)
gt[(xx - W * 0.3) ** 2 + (yy - H * 0.5) ** 2 < (H * 0.3) ** 2] = 1
gt[yy > H * 0.75] = 2
proto = rng.standard_normal((L, 3)) * 60 + 128
img = np.clip(proto[gt] + rng.standard_normal((H, W, 3)) * 25, 0, 255).astype(np.uint8)
feat = np.concatenate([np.eye(L, dtype=np.float32)[gt.ravel()].T + rng.standard_normal((L, W * H)).astype(np.float32) * 1.5,
                       np.ones((1, W * H), np.float32)])          # noisy one-hot evidence + bias: (F, N)
Lm0 = np.zeros((L, F), np.float32)                                 # an uninformative unary to start from
fg = np.stack([xx.ravel() / 3.0, yy.ravel() / 3.0]).astype(np.float32)
fb = np.stack([xx.ravel() / 30.0, yy.ravel() / 30.0] + [img[..., k].ravel() / 20.0 for k in range(3)]).astype(np.float32)
gtf = gt.ravel().copy()
gtf[rng.random(W * H) < 0.2] = -1                                  # a fifth of the pixels carries no annotation

if use_oracle:
    from oracle import pyoracle as po
    m = po.GradientRestated(W * H, L)
    m.setUnaryEnergy(Lm0, feat)
    m.addPairwiseEnergy(fg, 1.0)
    m.addPairwiseEnergy(fb, 1.0)
    grad = lambda: m.gradient(N_ITER, po.LOGLIKELIHOOD, gtf)[:3]
    get = lambda: (m.parameters(0), m.parameters(1))
    def put(u, l): m.setParameters(0, u); m.setParameters(1, l)
    def labels():
        g = po.CRF(W * H, L, "orc"); g.setUnaryEnergy(m.unary[1], m.unary[2])
        for t in m.terms: g.addPairwiseEnergy(t["f"], float(t["cw"][0]))
        return np.argmax(g.inference(N_ITER), axis=0)
else:
    import pydensecrf_b200.densecrf as dcrf
    m = dcrf.DenseCRF(W * H, L)
    m.setUnary(dcrf.LogisticUnary(Lm0, feat))
    m.addPairwiseEnergy(fg, 1.0)
    m.addPairwiseEnergy(fb, 1.0)
    obj = dcrf.LogLikelihood(gtf)
    grad = lambda: m.gradient(N_ITER, obj, kernel=False)[:3]
    get = lambda: (m.unaryParameters(), m.labelCompatibilityParameters())
    def put(u, l): m.setUnaryParameters(u); m.setLabelCompatibilityParameters(l)
    labels = lambda: np.argmax(m.inference(N_ITER), axis=0)

u0, l0 = get()
nu = u0.size
calls = [0]


def f(p):
    put(p[:nu].astype(np.float32), p[nu:].astype(np.float32))
    v, gu, gl = grad()
    calls[0] += 1
    return -v, -np.concatenate([gu, gl]).astype(np.float64)       # the objective is maximised


acc0 = float((labels() == gt.ravel()).mean())
v0 = -f(np.concatenate([u0, l0]).astype(np.float64))[0]
t0 = time.perf_counter()
res = minimize(f, np.concatenate([u0, l0]).astype(np.float64), jac=True, method="L-BFGS-B", options={"maxiter": 40})
dt = time.perf_counter() - t0
put(res.x[:nu].astype(np.float32), res.x[nu:].astype(np.float32))
acc1 = float((labels() == gt.ravel()).mean())
print("%s, %dx%d, %d labels: log-likelihood %.4f -> %.4f in %d gradient calls (%.2f s, %.1f ms each); MAP accuracy %.3f -> %.3f; "
      "Potts weights %s -> %s" % ("numpy restatement" if use_oracle else "engine", W, H, L, v0, -res.fun, calls[0], dt,
                                  dt / max(calls[0] - 1, 1) * 1e3, acc0, acc1, np.round(l0, 3), np.round(res.x[nu:], 3)))
assert -res.fun > v0 and acc1 >= acc0
