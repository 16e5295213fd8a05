"""CPU stand-ins, backed by the C oracle, for the functions of `pointdae_b200.ops` (and the names `chamfer.py`
re-exports) that the reference's flagship model reaches.  TEST INFRASTRUCTURE ONLY: `apply()` swaps them in so this
repo's HOST layer (knn_cuda.KNN, pointnet2_utils, group.Group, chamfer_dist, corrupt_util_tensor) can run inside the
reference's real model on CPU tensors; the kernels themselves are checked against the same oracle by the GPU tests."""
import numpy as np
import torch

import _oracle_chamfer
import _oracle_ext
from oracle import cpu as oracle


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def furthest_point_sample(xyz, npoint):
    return _oracle_ext.furthest_point_sampling(xyz, npoint)


def fps_gather(data, number):
    src = data.detach().float().contiguous().numpy()
    idx = oracle.fps(np.ascontiguousarray(src[:, :, :3]), int(number))
    return _t(idx), _t(np.take_along_axis(src, idx[:, :, None].astype(np.int64), axis=1))


def knn_points(ref, query, k, out_kq=False, want_dist=True):
    d, i = oracle.knn(ref.numpy(), query.numpy(), int(k))
    d, i = _t(d), _t(i)
    if out_kq:
        d, i = d.transpose(1, 2).contiguous(), i.transpose(1, 2).contiguous()
    return (d if want_dist else None), i


def _neighbours(xyz, center, m):
    x, c = xyz.detach().numpy(), center.detach().numpy()
    _, idx = oracle.knn(x, c, int(m))
    return x, c, idx, np.stack([x[b][idx[b]] for b in range(x.shape[0])])


def group_points_knn(xyz, center, group_size, want_idx=True, subtract_center=True):
    x, c, idx, nb = _neighbours(xyz, center, group_size)
    if subtract_center:
        nb = nb - c[:, :, None, :]
    return _t(nb), (_t(idx) if want_idx else None)


def fps_group(xyz, num_group, group_size, want_idx=False):
    """stand-in of the single-launch patchifier: the same values as fps_gather + group_points_knn (bit for bit on the GPU)"""
    fps_idx, center = fps_gather(xyz, num_group)
    nb, idx = group_points_knn(xyz, center, group_size, want_idx=want_idx)
    return fps_idx, center, nb, idx


def fps_group_affine(xyz, num_group, group_size, mats, want_idx=False):
    fps_idx, center = fps_gather(xyz, num_group)
    nb, tnb, tc, idx = group_affine(xyz, center, group_size, mats, want_idx=want_idx)
    return fps_idx, center, nb, tnb, tc, idx


def affine_points(points, center, mats):
    p, c = oracle.affine_points(points.detach().float().contiguous().numpy(), center.detach().float().contiguous().numpy(),
                                mats.float().numpy())
    return _t(p).view(points.shape), _t(c).view(center.shape)


def group_affine(xyz, center, group_size, mats, want_idx=False):
    x, c, idx, nb = _neighbours(xyz, center, group_size)
    absn = (nb - c[:, :, None, :]) + c[:, :, None, :]
    tp, tc = oracle.affine_points(absn, c, mats.float().numpy())
    return _t(absn - c[:, :, None, :]), _t(tp - tc[:, :, None, :]), _t(tc), (_t(idx) if want_idx else None)


def chamfer_forward(xyz1, xyz2, symmetric=True, scan_done=None):
    return _oracle_chamfer.forward(xyz1, xyz2)


def chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    return _oracle_chamfer.backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)


def chamfer_mean_loss(dist1, dist2, l1=False):
    d1, d2 = dist1.double(), dist2.double()
    m1, m2 = (d1.sqrt().mean(), d2.sqrt().mean()) if l1 else (d1.mean(), d2.mean())
    loss = (m1.float() + m2.float()) * 0.5 if l1 else m1.float() + m2.float()
    return torch.stack([loss, m1.float(), m2.float()])


def chamfer_loss_backward(xyz1, xyz2, idx1, idx2, dist1, dist2, grad_loss, w1, w2, l1=False):
    g = grad_loss.reshape(-1)[0].float()
    gd1 = torch.full_like(dist1, 1.0 / dist1.numel()) * (g * w1)
    gd2 = torch.full_like(dist2, 1.0 / dist2.numel()) * (g * w2)
    if l1:
        gd1, gd2 = gd1 / (2 * dist1.sqrt()), gd2 / (2 * dist2.sqrt())
    return _oracle_chamfer.backward(xyz1, xyz2, idx1, idx2, gd1, gd2)


def feat_knn(x, k):
    idx, _ = oracle.feat_knn(x.detach().float().contiguous().numpy(), int(k))
    return _t(idx)


def _graph_feature_fwd(x, idx):
    return _t(np.ascontiguousarray(oracle.graph_feature(x.detach().numpy(), idx.numpy()).transpose(0, 2, 3, 1)))  # physical (B,N,k,2C)


def _graph_feature_bwd(gout_phys, idx, c, n):
    return _t(oracle.graph_feature_grad(gout_phys.numpy().transpose(0, 3, 1, 2), idx.numpy()))


class GraphFeatureFunction(torch.autograd.Function):  # same contract as ops.GraphFeatureFunction
    @staticmethod
    def forward(ctx, x, idx):
        ctx.save_for_backward(idx)
        ctx.cn = (x.size(1), x.size(2))
        return _graph_feature_fwd(x.contiguous(), idx).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad):
        (idx,) = ctx.saved_tensors
        c, n = ctx.cn
        return _graph_feature_bwd(grad.permute(0, 2, 3, 1).contiguous().float(), idx, c, n), None


def edge_conv_max(x, idx, weight, scale, shift, slope=0.2):
    return _t(oracle.edge_conv_max(x.detach().float().contiguous().numpy(), idx.numpy(), weight.detach().float().contiguous().numpy(),
                                   scale.detach().float().numpy(), shift.detach().float().numpy(), float(slope)))


def edge_conv(x, idx, weight, bn, slope=0.2):
    """CPU stand-in for ops.edge_conv: the layer spelled out with torch's own differentiable ops (the reference's
    expression, models/dgcnn_util.py:24-34 + the block + max), so the HOST wiring of dgcnn_util.edge_conv /
    dgcnn_encoder_forward can run inside the reference's model on CPU tensors."""
    import torch.nn.functional as F
    b, c, n = x.shape
    k = idx.size(2)
    flat = (idx + torch.arange(b).view(-1, 1, 1) * n).view(-1)
    xt = x.transpose(2, 1).contiguous()
    neigh = xt.view(b * n, c)[flat, :].view(b, n, k, c)
    xi = xt.view(b, n, 1, c).expand(-1, -1, k, -1)
    feature = torch.cat((neigh - xi, xi), dim=3).permute(0, 3, 1, 2)
    y = F.conv2d(feature, weight.reshape(weight.size(0), 2 * c, 1, 1))
    return F.leaky_relu(bn(y), slope).max(dim=-1, keepdim=False)[0]


NAMES = ("edge_conv", "edge_conv_max", "feat_knn", "_graph_feature_fwd", "_graph_feature_bwd", "GraphFeatureFunction", "furthest_point_sample", "fps_gather", "knn_points", "group_points_knn", "fps_group", "fps_group_affine", "affine_points", "group_affine",
         "chamfer_forward", "chamfer_backward", "chamfer_mean_loss", "chamfer_loss_backward")
EXT_NAMES = ("gather_points", "gather_points_grad", "ball_query", "group_points", "group_points_grad", "three_nn",
             "three_interpolate", "three_interpolate_grad")


def apply():
    """Swap the stand-ins into pointdae_b200.ops and the names pointdae_b200.chamfer re-exports (process-wide)."""
    from pointdae_b200 import chamfer, ops
    for name in NAMES:
        setattr(ops, name, globals()[name])
    for name in EXT_NAMES:
        setattr(ops, name, getattr(_oracle_ext, name))
    chamfer.forward, chamfer.backward = chamfer_forward, chamfer_backward
    chamfer.mean_loss, chamfer.loss_backward = chamfer_mean_loss, chamfer_loss_backward
