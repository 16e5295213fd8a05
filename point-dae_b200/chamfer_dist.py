"""Mirror of `extensions.chamfer_dist` (extensions/chamfer_dist/__init__.py of the reference):
ChamferFunction and the loss modules on the hot path, same names / arguments / return values.
"""
import torch

from . import chamfer


class ChamferFunction(torch.autograd.Function):
    # reference: extensions/chamfer_dist/__init__.py:14-26
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = chamfer.forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2, grad_idx1, grad_idx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        grad_xyz1, grad_xyz2 = chamfer.backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)
        return grad_xyz1, grad_xyz2


def _drop_zero_points(xyz1, xyz2):
    # reference: __init__.py:38-42 (only taken when batch_size == 1 and ignore_zeros)
    non_zeros1 = torch.sum(xyz1, dim=2).ne(0)
    non_zeros2 = torch.sum(xyz2, dim=2).ne(0)
    return xyz1[non_zeros1].unsqueeze(dim=0), xyz2[non_zeros2].unsqueeze(dim=0)


class ChamferDistanceL2(torch.nn.Module):
    """reference: __init__.py:29-44.  mean(dist1) + mean(dist2)."""

    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1, xyz2 = _drop_zero_points(xyz1, xyz2)
        dist1, dist2, idx1, idx2 = ChamferFunction.apply(xyz1, xyz2)
        return torch.mean(dist1) + torch.mean(dist2)


class ChamferDistanceL2_split(torch.nn.Module):
    """reference: __init__.py:379-395.  (mean(dist1), mean(dist2))."""

    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1, xyz2 = _drop_zero_points(xyz1, xyz2)
        dist1, dist2, _, _ = ChamferFunction.apply(xyz1, xyz2)
        return torch.mean(dist1), torch.mean(dist2)


class ChamferDistanceL1(torch.nn.Module):
    """reference: __init__.py:397-417.  (mean(sqrt(dist1)) + mean(sqrt(dist2))) / 2."""

    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1, xyz2 = _drop_zero_points(xyz1, xyz2)
        dist1, dist2, _, _ = ChamferFunction.apply(xyz1, xyz2)
        dist1 = torch.sqrt(dist1)
        dist2 = torch.sqrt(dist2)
        return (torch.mean(dist1) + torch.mean(dist2)) / 2
