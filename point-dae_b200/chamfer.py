"""Drop-in for the compiled `chamfer` extension module (extensions/chamfer_dist/chamfer_cuda.cpp:36-39):
`chamfer.forward(xyz1, xyz2)` and `chamfer.backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)`.
The reference's own extensions/chamfer_dist/__init__.py runs unchanged on top of this module."""
from .ops import chamfer_backward as backward  # noqa: F401
from .ops import chamfer_forward as forward  # noqa: F401
