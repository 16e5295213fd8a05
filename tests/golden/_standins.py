"""Oracle-backed stand-ins for the three compiled / third-party names the reference's Python touches on this path
(`knn_cuda.KNN`, `pointnet2_ops.pointnet2_utils.furthest_point_sample / gather_operation`, `utils.misc.fps`), so the
reference's own Python can run unmodified on CPU tensors in the build container.  Used only by the golden-vector
generators in this directory; the oracle behind them is pinned to the reference kernels by tests/golden/fps_gather.npz."""
import os
import sys
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import cpu as oracle  # noqa: E402


class KNN(nn.Module):
    def __init__(self, k, transpose_mode=False):
        super().__init__()
        self.k, self.transpose_mode = k, transpose_mode

    def forward(self, ref, query):
        if not self.transpose_mode:
            ref, query = ref.transpose(1, 2), query.transpose(1, 2)
        d, i = oracle.knn(ref.contiguous().numpy(), query.contiguous().numpy(), self.k)
        d, i = torch.from_numpy(d), torch.from_numpy(i)
        return (d, i) if self.transpose_mode else (d.transpose(1, 2).contiguous(), i.transpose(1, 2).contiguous())


def furthest_point_sample(xyz, npoint):
    return torch.from_numpy(oracle.fps(xyz.numpy(), npoint))


def gather_operation(features, idx):
    return torch.from_numpy(oracle.gather(features.numpy(), idx.numpy()))


def fps(data, number):  # utils/misc.py:13-20
    fps_idx = furthest_point_sample(data[:, :, :3].contiguous(), number)
    fps_data = gather_operation(data.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2).contiguous()
    return fps_idx, fps_data


pointnet2_utils = types.SimpleNamespace(furthest_point_sample=furthest_point_sample, gather_operation=gather_operation)
misc = types.SimpleNamespace(fps=fps)


def install_modules():
    """sys.modules entries for `from knn_cuda import KNN` / `from pointnet2_ops import pointnet2_utils`."""
    knn = types.ModuleType("knn_cuda")
    knn.KNN = KNN
    sys.modules["knn_cuda"] = knn
    p2 = types.ModuleType("pointnet2_ops")
    p2u = types.ModuleType("pointnet2_ops.pointnet2_utils")
    p2u.furthest_point_sample, p2u.gather_operation = furthest_point_sample, gather_operation
    p2.pointnet2_utils = p2u
    sys.modules["pointnet2_ops"] = p2
    sys.modules["pointnet2_ops.pointnet2_utils"] = p2u
