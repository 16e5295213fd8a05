"""Drop-in for the two hot functions of `models/dgcnn_util.py` of the reference: `knn` (:7-12) and
`get_graph_feature` (:15-36), same signatures.  `install()` rebinds them inside the reference's
models.dgcnn_util / models.PointCAE_DGCNN / segmentation.models.dgcnn_util.
"""
import torch

from . import ops


def knn(x, k):
    """x (B,C,N) -> idx (B,N,k) int64, nearest first, self included.  The reference ranks by the
    expanded form through cuBLAS (tie order unspecified); this ranks by the direct-form squared
    distance with ties -> lower index, and never materialises the B x N x N matrix."""
    return ops.feat_knn(x, k)


def get_graph_feature(x, k=20, idx=None, extra_dim=False):
    batch_size, num_dims, num_points = x.size()
    x = x.view(batch_size, -1, num_points)
    caller_idx = idx
    if idx is None:
        idx = knn(x, k=k) if extra_dim is False else knn(x[:, 6:], k=k)
    if idx.dtype != torch.int64 or not idx.is_contiguous():
        idx = idx.long().contiguous()
    k = idx.size(2)
    if idx is caller_idx:
        # the caller's tensor is offset in place below (as the reference does); autograd must keep the per-cloud
        # indices, so the function saves its own copy (found by tests/test_ref_dgcnn.py: backward raised otherwise)
        idx = idx.clone()
    feature = ops.GraphFeatureFunction.apply(x.float(), idx)
    if caller_idx is not None and caller_idx.dtype == torch.int64:
        # the reference offsets a caller-supplied idx in place (dgcnn_util.py:27); keep that visible effect
        idx_base = torch.arange(0, batch_size, device=caller_idx.device).view(-1, 1, 1) * num_points
        caller_idx += idx_base
    return feature  # (batch_size, 2 * num_dims, num_points, k)


# ---- EdgeConv layers without the k-replicated tensors ---------------------------------------------
def fold_batchnorm(bn):
    """BatchNorm (running statistics) as y * scale + shift."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return scale, bn.bias - scale * bn.running_mean


def _fusable(block):
    """Sequential(Conv2d(2C, Co, 1, bias=False), BatchNorm2d, LeakyReLU) -- what models/dgcnn_util.py:96-110 builds."""
    try:
        conv, bn, act = block[0], block[1], block[2]
    except (TypeError, IndexError):
        return False
    return (isinstance(conv, torch.nn.Conv2d) and conv.bias is None and tuple(conv.kernel_size) == (1, 1)
            and tuple(conv.stride) == (1, 1) and conv.groups == 1 and isinstance(bn, torch.nn.BatchNorm2d)
            and isinstance(act, torch.nn.LeakyReLU) and len(block) == 3
            and (bn.training or bn.running_mean is not None)  # eval mode needs running statistics
            and not torch.is_autocast_enabled())


def edge_conv(x, block, k=20, idx=None):
    """`get_graph_feature(x, k)` -> `block` -> max over k (models/dgcnn_util.py:114-116 and the three layers after it):
    x (B,C,N) -> (B,Co,N), differentiable, training or eval BatchNorm (the block's own mode), on the tensor cores
    (ops.edge_conv); the (B,2C,N,k) and (B,Co,N,k) tensors are never formed.  A block of another shape runs as the
    reference wrote it."""
    if not _fusable(block):
        return block(get_graph_feature(x, k=k, idx=idx)).max(dim=-1, keepdim=False)[0]
    if idx is None:
        idx = knn(x, k)
    return ops.edge_conv(x, idx, block[0].weight, block[1], block[2].negative_slope)


def edge_conv_eval(x, block, k=20, idx=None):
    """round-1 name: the same layer without autograd"""
    with torch.no_grad():
        return edge_conv(x, block, k=k, idx=idx)


def dgcnn_encoder_forward(self, x):
    """Drop-in for `dgcnn_encoder.forward` (models/dgcnn_util.py:112-133): the four EdgeConv layers take the fused route
    (training and eval, with or without autograd), decided per block; conv5 and the pooling are the reference's."""
    batch_size = x.size()[0]
    feats = []
    for block in (self.conv1, self.conv2, self.conv3, self.conv4):
        x = edge_conv(x, block, k=20)
        feats.append(x)
    x = self.conv5(torch.cat(feats, dim=1))
    return torch.nn.functional.adaptive_max_pool1d(x, 1).view(batch_size, -1)
