"""A CPU stand-in for the compiled `chamfer` module backed by the C oracle (oracle/pdae_oracle.c), with
the reference's raw-storage semantics (chamfer.cu:159-164: data_ptr walked as dense [B][size(1)][3],
gradients zeros_like).  TEST INFRASTRUCTURE ONLY: it lets the reference's Python loss classes
(extensions/chamfer_dist/__init__.py) run in a CPU-only container to produce golden vectors, and lets
the CPU suite check this repo's host-side loss arithmetic against those vectors."""
import numpy as np
import torch

from oracle import cpu as oracle


def _as_points(t):
    """first B*size(1)*3 floats of the tensor's storage order, shaped (B, size(1), 3)"""
    b, n = t.size(0), t.size(1)
    flat = t.detach().contiguous().view(-1).numpy() if t.is_contiguous() else _storage_order(t)
    return np.ascontiguousarray(flat[: b * n * 3].reshape(b, n, 3))


def _storage_order(t):
    order = sorted(range(t.dim()), key=lambda d: -t.stride(d))
    return t.detach().permute(order).contiguous().view(-1).numpy()


def _write_back(like, grad_points):
    """zeros_like(like) with grad_points written into the leading floats of its storage order"""
    out = torch.zeros_like(like)
    if like.is_contiguous():
        out.view(-1)[: grad_points.size] = torch.from_numpy(grad_points.reshape(-1))
    else:
        order = sorted(range(like.dim()), key=lambda d: -like.stride(d))
        out.permute(order).reshape(-1)[: grad_points.size] = torch.from_numpy(grad_points.reshape(-1))
    return out


def forward(xyz1, xyz2):
    d1, d2, i1, i2 = oracle.chamfer_fwd(_as_points(xyz1), _as_points(xyz2))
    return [torch.from_numpy(d1), torch.from_numpy(d2), torch.from_numpy(i1), torch.from_numpy(i2)]


def backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    g1, g2 = oracle.chamfer_bwd(_as_points(xyz1), _as_points(xyz2), idx1.numpy(), idx2.numpy(),
                                grad_dist1.contiguous().numpy(), grad_dist2.contiguous().numpy())
    return [_write_back(xyz1, g1), _write_back(xyz2, g2)]


def mean_loss(dist1, dist2, l1=False):
    """CPU stand-in of the fused loss epilogue: the reference's own torch arithmetic (__init__.py:43, :413-417)"""
    if l1:
        a, b = torch.mean(torch.sqrt(dist1)), torch.mean(torch.sqrt(dist2))
        return torch.stack([(a + b) / 2, a, b])
    a, b = torch.mean(dist1), torch.mean(dist2)
    return torch.stack([a + b, a, b])


def loss_backward(xyz1, xyz2, idx1, idx2, dist1, dist2, grad_loss, w1, w2, l1=False):
    g = grad_loss.reshape(()).float()
    gd1 = torch.full_like(dist1, 1.0) * (g * w1 / dist1.numel())
    gd2 = torch.full_like(dist2, 1.0) * (g * w2 / dist2.numel())
    if l1:
        gd1, gd2 = gd1 / (2 * torch.sqrt(dist1)), gd2 / (2 * torch.sqrt(dist2))
    return backward(xyz1, xyz2, idx1, idx2, gd1, gd2)
