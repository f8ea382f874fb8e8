"""Deterministic synthetic inputs for the BASELINE.json configs.

numpy only, seeded with numpy.random.default_rng(seed).  Used by tests/ and bench.py; not part of
the product package.
"""
import numpy as np


def _axis_interp(a, n_out, axis, cubic):
    """Resample `a` along `axis` to n_out samples (align-corners), linear or Catmull-Rom cubic."""
    n_in = a.shape[axis]
    x = np.linspace(0.0, n_in - 1.0, n_out)
    i0 = np.clip(np.floor(x).astype(np.int64), 0, n_in - 2)
    t = (x - i0).astype(np.float32)
    a = np.moveaxis(a, axis, 0).astype(np.float32)
    shp = (n_out,) + (1,) * (a.ndim - 1)
    t = t.reshape(shp)
    if not cubic:
        out = a[i0] * (1 - t) + a[i0 + 1] * t
    else:
        im1 = np.clip(i0 - 1, 0, n_in - 1)
        ip2 = np.clip(i0 + 2, 0, n_in - 1)
        p0, p1, p2, p3 = a[im1], a[i0], a[i0 + 1], a[ip2]
        out = p1 + 0.5 * t * (p2 - p0 + t * (2 * p0 - 5 * p1 + 4 * p2 - p3 + t * (3 * (p1 - p2) + p3 - p0)))
    return np.moveaxis(out, 0, axis)


def smooth_image(H, W, seed=0, grid=32, noise_sigma=4.0):
    """'smooth' RGB image: uniform noise on a coarse grid, bicubic-upsampled, + N(0, sigma) noise."""
    rng = np.random.default_rng(seed)
    g = rng.uniform(0, 255, (H // grid + 2, W // grid + 2, 3)).astype(np.float32)
    im = _axis_interp(_axis_interp(g, H, 0, True), W, 1, True)
    im = im + rng.normal(0, noise_sigma, im.shape).astype(np.float32)
    return np.clip(np.rint(im), 0, 255).astype(np.uint8)


def noise_image(H, W, seed=0):
    """i.i.d. uniform RGB: worst case for the bilateral lattice size."""
    return np.random.default_rng(seed).integers(0, 256, (H, W, 3), dtype=np.uint8)


def softmax_unary(H, W, L, seed=0, grid=16, clip=1e-5):
    """-log softmax(logits), logits = N(0,2) on a coarse grid per label, bilinear-upsampled.
    Same result as pydensecrf.utils.unary_from_softmax(softmax) (utils.py:55-79).  (L, H*W) float32."""
    rng = np.random.default_rng(seed + 1000003)
    g = rng.normal(0, 2, (L, H // grid + 2, W // grid + 2)).astype(np.float32)
    lg = _axis_interp(_axis_interp(g, H, 1, False), W, 2, False)
    lg = lg - lg.max(axis=0, keepdims=True)
    e = np.exp(lg)
    sm = e / e.sum(axis=0, keepdims=True)
    sm = np.clip(sm, clip, 1.0)
    return np.ascontiguousarray(-np.log(sm).reshape(L, -1).astype(np.float32))


def matrix_compat(L, seed=0, w=3.0):
    """cfg 3 compatibility: -(w*I) + 0.1*|N(0,1)|, symmetrised, float32."""
    rng = np.random.default_rng(seed + 7)
    c = -w * np.eye(L) + 0.1 * np.abs(rng.normal(0, 1, (L, L)))
    return np.ascontiguousarray(0.5 * (c + c.T), dtype=np.float32)


def smooth_volume(D, seed=0, grid_cells=8, noise_sigma=4.0):
    """cfg 5: D^3 uint8 volume: uniform noise on a (grid_cells+1)^3 grid, trilinear-upsampled + N(0,4)."""
    rng = np.random.default_rng(seed)
    g = rng.uniform(0, 255, (grid_cells + 1,) * 3).astype(np.float32)
    v = g
    for ax in range(3):
        v = _axis_interp(v, D, ax, False)
    v = v + rng.normal(0, noise_sigma, v.shape).astype(np.float32)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def volume_unary(D, seed=0, grid_cells=8):
    """cfg 5: 2-label unary from a sigmoid of smooth noise.  (2, D^3) float32."""
    rng = np.random.default_rng(seed + 99)
    g = rng.normal(0, 2, (grid_cells + 1,) * 3).astype(np.float32)
    v = g
    for ax in range(3):
        v = _axis_interp(v, D, ax, False)
    p = 1.0 / (1.0 + np.exp(-v.reshape(-1)))
    sm = np.clip(np.stack([p, 1 - p]), 1e-5, 1.0)
    return np.ascontiguousarray(-np.log(sm).astype(np.float32))


def toy_unary():
    """The 10x10 two-label toy of the reference's own tests (pydensecrf/tests/test_dcrf.py:8-18)."""
    u1 = np.zeros((10, 10), np.float32)
    u1[:, [0, -1]] = u1[[0, -1], :] = 1
    u2 = np.zeros((10, 10), np.float32)
    u2[4:7, 4:7] = 1
    u = np.vstack([u1.flat, u2.flat])
    u = (u + 1) / (np.sum(u, axis=0) + 2)
    return np.ascontiguousarray(-np.log(u), dtype=np.float32)


def toy_image():
    """pydensecrf/tests/test_dcrf.py:21-26"""
    img = np.zeros((10, 10, 3), np.uint8)
    img[2:8, 2:8, :] = 255
    return img


def stage_bytes_per_iter(N, L, kernels):
    """Algorithmic bytes of one mean-field iteration, SURVEY.md section 8d stage model.
    kernels = [(d_k, V_k), ...]."""
    K = len(kernels)
    b = (K + 2) * 4 * N * L
    for d, V in kernels:
        b += (4 * N * L + 8 * N * (d + 1) + 4 * V * L)          # splat
        b += (d + 1) * (2 * 4 * V * L + 8 * V)                  # blur
        b += (4 * V * L + 8 * N * (d + 1) + 4 * N * L)          # slice
    return b
