"""Drop-in for the compiled `chamfer` module (extensions/chamfer_dist/chamfer_cuda.cpp:36-39):
`forward(xyz1, xyz2)` and `backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)`; plus the fused
mean-loss epilogue (`mean_loss`, `loss_backward`) used by chamfer_dist.ChamferDistanceL1 / L2."""
from .ops import chamfer_backward as backward  # noqa: F401
from .ops import chamfer_forward as forward  # noqa: F401
from .ops import chamfer_loss_backward as loss_backward  # noqa: F401
from .ops import chamfer_mean_loss as mean_loss  # noqa: F401
