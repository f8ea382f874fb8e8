"""Import-name alias: `import pydensecrf.densecrf as dcrf`, `from pydensecrf.utils import ...` and `import pydensecrf.eigen`
resolve to the B200 engine (pydensecrf_b200), so scripts written for lucasb-eyer/pydensecrf -- its examples/inference.py,
its own tests -- run unmodified.  Nothing is implemented here."""
import sys as _sys

import pydensecrf_b200 as _impl
from pydensecrf_b200 import densecrf, eigen, utils  # noqa: F401

# one module object per name, whichever way it is imported
_sys.modules[__name__ + ".densecrf"] = densecrf
_sys.modules[__name__ + ".eigen"] = eigen
_sys.modules[__name__ + ".utils"] = utils
__all__ = _impl.__all__
