"""numpy helpers with the semantics of `pydensecrf.utils` (reference: pydensecrf/utils.py).

Pure host-side input preparation (unary energies and feature matrices); nothing here is on the GPU
path.  Written against the documented behaviour of the reference functions named in each docstring.
"""
import logging
from numbers import Number

import numpy as np


def pinned_empty(shape, dtype="float32"):
    """Addition to the reference's utils: an uninitialised numpy array in page-locked host memory, so
    that setUnaryEnergy / addPairwiseBilateral upload it by DMA at full PCIe rate."""
    from ._core import pinned_empty as _pe

    return _pe(shape, dtype)


def unary_from_labels(labels, n_labels, gt_prob, zero_unsure=True):
    """Unary energies from a hard label map (reference: utils.py:6-46).

    Every variable gets -log(gt_prob) for its annotated label and -log((1-gt_prob)/(n_labels-1)) for
    the others.  With `zero_unsure`, label value 0 means "unknown" (uniform energy) and the real labels
    are 1..n_labels.  Returns float32 (n_labels, labels.size)."""
    assert 0 < gt_prob < 1, "`gt_prob must be in (0,1)."
    flat = np.asarray(labels).flatten()
    e_other = -np.log((1.0 - gt_prob) / (n_labels - 1))
    e_label = -np.log(gt_prob)
    U = np.full((n_labels, flat.size), e_other, dtype="float32")
    rows = flat - 1 if zero_unsure else flat
    U[rows, np.arange(flat.size)] = e_label   # label 0 lands on the last row here and is overwritten next
    if zero_unsure:
        U[:, flat == 0] = -np.log(1.0 / n_labels)
    return U


def compute_unary(labels, M, GT_PROB=0.5):
    """Deprecated alias kept for compatibility (reference: utils.py:49-52)."""
    logging.warning("pydensecrf.compute_unary is deprecated, use unary_from_labels instead.")
    return unary_from_labels(labels, M, GT_PROB)


def unary_from_softmax(sm, scale=None, clip=1e-5):
    """Negative log of class probabilities (reference: utils.py:55-79).  `sm` has the classes on
    axis 0; `scale` blends towards uniform; `clip` bounds probabilities from below.  float32 (C, N)."""
    n_cls = sm.shape[0]
    if scale is not None:
        assert 0 < scale <= 1, "`scale` needs to be in (0,1]"
        sm = scale * sm + (1 - scale) * (np.ones(sm.shape) / n_cls)
    if clip is not None:
        sm = np.clip(sm, clip, 1.0)
    return -np.log(sm).reshape([n_cls, -1]).astype(np.float32)


def softmax_to_unary(sm, GT_PROB=1):
    """Deprecated alias (reference: utils.py:82-86)."""
    logging.warning("pydensecrf.softmax_to_unary is deprecated, use unary_from_softmax instead.")
    return unary_from_softmax(sm, None if GT_PROB == 1 else GT_PROB, clip=None)


def _scaled_mesh(shape, sdims):
    mesh = np.array(np.meshgrid(*[range(s) for s in shape], indexing="ij"), dtype=np.float32)
    for axis, s in enumerate(sdims):
        mesh[axis] /= s
    return mesh


def create_pairwise_gaussian(sdims, shape):
    """Position-only features, one row per spatial axis, coordinates divided by `sdims`
    (reference: utils.py:89-111).  float32 (len(sdims), prod(shape))."""
    return _scaled_mesh(shape, sdims).reshape([len(sdims), -1])


def create_pairwise_bilateral(sdims, schan, img, chdim=-1):
    """Position + appearance features (reference: utils.py:114-162): spatial coordinates / sdims
    followed by the image channels / schan.  `chdim` is the channel axis of `img` (-1: no channel
    axis).  float32 (n_spatial + n_channels, n_variables)."""
    if chdim == -1:
        chans = img[np.newaxis].astype(np.float32)
    else:
        chans = np.rollaxis(img, chdim).astype(np.float32)
    if isinstance(schan, Number):
        chans /= schan
    else:
        for c, s in enumerate(schan):
            chans[c] /= s
    feats = np.concatenate([_scaled_mesh(chans.shape[1:], sdims), chans])
    return feats.reshape([feats.shape[0], -1])
