"""pydensecrf_b200 -- pydensecrf's Python surface on a B200-native (sm_100a) mean-field engine.

    import pydensecrf_b200.densecrf as dcrf      # instead of pydensecrf.densecrf
    from pydensecrf_b200 import utils            # instead of pydensecrf.utils

The compute lives in libdcrf_b200.so (hand-written CUDA behind the C-ABI in include/dcrf_b200.h);
`_core` is the Cython binding.  There is no CPU fallback: importing works anywhere, but creating a
DenseCRF needs a CUDA device.  Build in-tree with `python -c "import __graft_entry__ as g; g.build()"`.
"""
try:
    from . import _core  # noqa: F401
except ImportError as e:  # fail loudly: nothing in this package computes without the extension
    raise ImportError(
        "pydensecrf_b200: the native extension (_core + libdcrf_b200.so) is not built; run "
        "`python -c 'import __graft_entry__ as g; g.build()'` in the repo root") from e

from . import densecrf, eigen, utils  # noqa: E402,F401

__all__ = ["densecrf", "eigen", "utils"]
