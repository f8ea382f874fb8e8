"""Drop-in for `pydensecrf.eigen` (reference: pydensecrf/eigen.pyx)."""
from ._core import MatrixXf, VectorXf, matrixXf, vectorXf  # noqa: F401
