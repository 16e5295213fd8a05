"""Drop-in for the `knn_cuda` package (KNN_CUDA 0.2): `KNN(k, transpose_mode=False)`.

Call sites in the reference: models/PointCAE_transformer.py:59,76 (transpose_mode=True),
models/MaskSurf_v2.py:79,124 (transpose_mode=False), datasets/corrupt_util_tensor.py:591 (module
level construction at import time -- so __init__ must not touch CUDA).  One batched launch
replaces the upstream per-cloud Python loop (~12 launches per cloud).
"""
import torch
import torch.nn as nn

from . import ops

__version__ = "0.2"


class KNN(nn.Module):
    def __init__(self, k, transpose_mode=False):
        super(KNN, self).__init__()
        self.k = k
        self._t = transpose_mode

    def forward(self, ref, query):
        """transpose_mode=True : ref (B,R,D), query (B,Q,D) -> D (B,Q,k) f32, I (B,Q,k) int64
        transpose_mode=False: ref (B,D,R), query (B,D,Q) -> D (B,k,Q),     I (B,k,Q)
        Distances are Euclidean (sqrt), neighbours ascending, ties -> lower index."""
        assert ref.size(0) == query.size(0), "ref.shape={} != query.shape={}".format(ref.shape, query.shape)
        with torch.no_grad():
            r, q = ref.detach().float(), query.detach().float()
            if not self._t:
                r, q = r.transpose(1, 2), q.transpose(1, 2)
            r, q = r.contiguous(), q.contiguous()
            D, I = ops.knn_points(r, q, self.k, out_kq=not self._t, want_dist=True)
        return D, I
