"""Drop-in for `pydensecrf.densecrf` (reference: pydensecrf/densecrf.pyx, densecrf.pxd)."""
from ._core import (  # noqa: F401
    CONST_KERNEL,
    DIAG_KERNEL,
    FULL_KERNEL,
    NO_NORMALIZATION,
    NORMALIZE_AFTER,
    NORMALIZE_BEFORE,
    NORMALIZE_SYMMETRIC,
    ConstUnary,
    DenseCRF,
    DenseCRF2D,
    KernelType,
    LogisticUnary,
    NormalizationType,
    Unary,
    abi_version,
    device_count,
)
