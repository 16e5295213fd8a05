"""numpy front-end of the CPU oracle (oracle/pdae_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (point-dae_b200/) never imports it.

Each wrapper cites the reference file:line the C function restates.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pdae_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libpdae_oracle.so")

_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off (no implicit fusion) -fopenmp -> oracle/_build/libpdae_oracle.so"""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared", "-std=gnu11",
           "-o", LIB, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
            build()
        _lib = ctypes.CDLL(LIB)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fps_block_size(n):
    """cuda_utils.h:15-21 (opt_n_threads)."""
    return int(lib().pdae_oracle_fps_block_size(int(n)))


def fps(xyz, npoint):
    """sampling_gpu.cu:72-176 / sampling.cpp:67-88.  xyz (B,N,3) -> (B,npoint) int32."""
    xyz = _f32(xyz)
    b, n, _ = xyz.shape
    idx = np.zeros((b, npoint), dtype=np.int32)
    lib().pdae_oracle_fps(_p(xyz), b, n, int(npoint), _p(idx))
    return idx


def gather(feat, idx):
    """sampling_gpu.cu:11-23.  feat (B,C,N), idx (B,M) int32 -> (B,C,M)."""
    feat, idx = _f32(feat), _i32(idx)
    b, c, n = feat.shape
    m = idx.shape[1]
    out = np.zeros((b, c, m), dtype=np.float32)
    lib().pdae_oracle_gather(_p(feat), _p(idx), b, c, n, m, _p(out))
    return out


def gather_grad(gout, idx, n):
    """sampling_gpu.cu:37-50.  gout (B,C,M), idx (B,M) -> (B,C,N)."""
    gout, idx = _f32(gout), _i32(idx)
    b, c, m = gout.shape
    out = np.zeros((b, c, n), dtype=np.float32)
    lib().pdae_oracle_gather_grad(_p(gout), _p(idx), b, c, int(n), m, _p(out))
    return out


def knn(ref, query, k):
    """KNN_CUDA 0.2 algorithm, transpose_mode=True layout: ref (B,R,D), query (B,Q,D) ->
    dist (B,Q,k) f32 (Euclidean), idx (B,Q,k) int64."""
    ref, query = _f32(ref), _f32(query)
    b, r, d = ref.shape
    q = query.shape[1]
    dist = np.zeros((b, q, k), dtype=np.float32)
    idx = np.zeros((b, q, k), dtype=np.int64)
    lib().pdae_oracle_knn(_p(ref), _p(query), b, r, q, d, int(k), _p(dist), _p(idx))
    return dist, idx


def group(xyz, num_group, group_size):
    """models/PointCAE_transformer.py:61-86.  -> neighborhood (B,G,M,3), center (B,G,3), idx, fps_idx."""
    xyz = _f32(xyz)
    b, n, _ = xyz.shape
    fps_idx = np.zeros((b, num_group), dtype=np.int32)
    center = np.zeros((b, num_group, 3), dtype=np.float32)
    idx = np.zeros((b, num_group, group_size), dtype=np.int64)
    nb = np.zeros((b, num_group, group_size, 3), dtype=np.float32)
    lib().pdae_oracle_group(_p(xyz), b, n, int(num_group), int(group_size), _p(fps_idx), _p(center), _p(idx), _p(nb))
    return nb, center, idx, fps_idx


def chamfer_fwd(xyz1, xyz2):
    """chamfer.cu:15-171.  -> dist1 (B,N), dist2 (B,M) f32 squared; idx1, idx2 int32."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.zeros((b, n), dtype=np.float32)
    d2 = np.zeros((b, m), dtype=np.float32)
    i1 = np.zeros((b, n), dtype=np.int32)
    i2 = np.zeros((b, m), dtype=np.int32)
    lib().pdae_oracle_chamfer_fwd(_p(xyz1), _p(xyz2), b, n, m, _p(d1), _p(d2), _p(i1), _p(i2))
    return d1, d2, i1, i2


def chamfer_bwd(xyz1, xyz2, idx1, idx2, gd1, gd2):
    """chamfer.cu:173-229.  -> grad_xyz1 (B,N,3), grad_xyz2 (B,M,3)."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    idx1, idx2, gd1, gd2 = _i32(idx1), _i32(idx2), _f32(gd1), _f32(gd2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.zeros((b, n, 3), dtype=np.float32)
    g2 = np.zeros((b, m, 3), dtype=np.float32)
    lib().pdae_oracle_chamfer_bwd(_p(xyz1), _p(xyz2), _p(idx1), _p(idx2), _p(gd1), _p(gd2), b, n, m, _p(g1), _p(g2))
    return g1, g2


def feat_knn(x, k):
    """Canonical direct-form DGCNN knn (models/dgcnn_util.py:7-12 semantics).  x (B,C,N) ->
    idx (B,N,k) int64, squared dist (B,N,k)."""
    x = _f32(x)
    b, c, n = x.shape
    dist = np.zeros((b, n, k), dtype=np.float32)
    idx = np.zeros((b, n, k), dtype=np.int64)
    lib().pdae_oracle_feat_knn(_p(x), b, c, n, int(k), _p(dist), _p(idx))
    return idx, dist


def graph_feature(x, idx):
    """models/dgcnn_util.py:15-36.  x (B,C,N), idx (B,N,k) -> logical (B,2C,N,k) view of a
    physical (B,N,k,2C) array, as the reference returns."""
    x, idx = _f32(x), _i64(idx)
    b, c, n = x.shape
    k = idx.shape[2]
    out = np.zeros((b, n, k, 2 * c), dtype=np.float32)
    lib().pdae_oracle_graph_feature(_p(x), _p(idx), b, c, n, k, _p(out))
    return out.transpose(0, 3, 1, 2)


def graph_feature_grad(gout_logical, idx):
    """gout logical (B,2C,N,k) -> gx (B,C,N)."""
    idx = _i64(idx)
    g = _f32(np.asarray(gout_logical).transpose(0, 2, 3, 1))  # physical (B,N,k,2C)
    b, n, k, c2 = g.shape
    c = c2 // 2
    gx = np.zeros((b, c, n), dtype=np.float32)
    lib().pdae_oracle_graph_feature_grad(_p(g), _p(idx), b, c, n, k, _p(gx))
    return gx


def ball_query(radius, nsample, xyz, new_xyz):
    """ball_query_gpu.cu:12-47.  xyz (B,N,3), new_xyz (B,M,3) -> idx (B,M,nsample) int32."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((b, m, nsample), dtype=np.int32)
    lib().pdae_oracle_ball_query.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    lib().pdae_oracle_ball_query(_p(new_xyz), _p(xyz), b, n, m, float(radius), int(nsample), _p(idx))
    return idx


def group_points(points, idx):
    """group_points_gpu.cu:11-31.  points (B,C,N), idx (B,P,S) int32 -> (B,C,P,S)."""
    points, idx = _f32(points), _i32(idx)
    b, c, n = points.shape
    _, p, s = idx.shape
    out = np.zeros((b, c, p, s), dtype=np.float32)
    lib().pdae_oracle_group_points(_p(points), _p(idx), b, c, n, p, s, _p(out))
    return out


def group_points_grad(gout, idx, n):
    """group_points_gpu.cu:46-67.  gout (B,C,P,S) -> (B,C,N)."""
    gout, idx = _f32(gout), _i32(idx)
    b, c, p, s = gout.shape
    out = np.zeros((b, c, n), dtype=np.float32)
    lib().pdae_oracle_group_points_grad(_p(gout), _p(idx), b, c, int(n), p, s, _p(out))
    return out


def three_nn(unknown, known):
    """interpolate_gpu.cu:12-62.  unknown (B,n,3), known (B,m,3) -> squared dist (B,n,3), idx (B,n,3) int32."""
    unknown, known = _f32(unknown), _f32(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    d = np.zeros((b, n, 3), dtype=np.float32)
    i = np.zeros((b, n, 3), dtype=np.int32)
    lib().pdae_oracle_three_nn(_p(unknown), _p(known), b, n, m, _p(d), _p(i))
    return d, i


def three_interpolate(points, idx, weight):
    """interpolate_gpu.cu:76-104.  points (B,c,m), idx/weight (B,n,3) -> (B,c,n)."""
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out = np.zeros((b, c, n), dtype=np.float32)
    lib().pdae_oracle_three_interpolate(_p(points), _p(idx), _p(weight), b, c, m, n, _p(out))
    return out


def three_interpolate_grad(gout, idx, weight, m):
    """interpolate_gpu.cu:118-144.  gout (B,c,n) -> (B,c,m)."""
    gout, idx, weight = _f32(gout), _i32(idx), _f32(weight)
    b, c, n = gout.shape
    out = np.zeros((b, c, int(m)), dtype=np.float32)
    lib().pdae_oracle_three_interpolate_grad(_p(gout), _p(idx), _p(weight), b, c, n, int(m), _p(out))
    return out


def affine_points(points, center, mats):
    """datasets/corrupt_util_tensor.py:59-343 chained as in `corrupt_data` :706-728.  points (B,...,3), center (B,G,3),
    mats (B,T,3,3) applied in order to row vectors -> (points', center')."""
    points, center, mats = _f32(points), _f32(center), _f32(mats)
    b, t = mats.shape[0], mats.shape[1]
    p = points.size // (3 * b) if b else 0
    g = center.size // (3 * b) if b else 0
    out_p, out_c = np.zeros_like(points), np.zeros_like(center)
    lib().pdae_oracle_affine_points(_p(points), _p(center), _p(mats), b, p, g, t, _p(out_p), _p(out_c))
    return out_p, out_c


def group_affine(xyz, num_group, group_size, mats):
    """models/PointCAE_transformer.py:1010-1017: Group.forward + corrupt_data + re-centring.
    -> neighborhood, center, t_neighborhood, t_center, idx."""
    xyz, mats = _f32(xyz), _f32(mats)
    b, n, _ = xyz.shape
    t = mats.shape[1]
    fps_idx = np.zeros((b, num_group), dtype=np.int32)
    center = np.zeros((b, num_group, 3), dtype=np.float32)
    idx = np.zeros((b, num_group, group_size), dtype=np.int64)
    nb = np.zeros((b, num_group, group_size, 3), dtype=np.float32)
    tnb = np.zeros_like(nb)
    tc = np.zeros_like(center)
    lib().pdae_oracle_group_affine(_p(xyz), _p(mats), b, n, int(num_group), int(group_size), t, _p(fps_idx), _p(center),
                                   _p(idx), _p(nb), _p(tnb), _p(tc))
    return nb, center, tnb, tc, idx


def edge_conv_max(x, idx, weight, scale, shift, slope=0.2):
    """models/dgcnn_util.py:114-126, one eval-mode EdgeConv layer: x (B,C,N), idx (B,N,k), weight (Co,2C),
    BatchNorm folded to scale / shift (Co) -> (B,Co,N) = max_k LeakyReLU(BN(conv([x_j - x_i; x_i])))."""
    x, idx, weight, scale, shift = _f32(x), _i64(idx), _f32(weight), _f32(scale), _f32(shift)
    b, c, n = x.shape
    k, co = idx.shape[2], weight.shape[0]
    out = np.zeros((b, co, n), dtype=np.float32)
    lib().pdae_oracle_edge_conv_max.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_float] + [ctypes.c_int] * 5 + [ctypes.c_void_p]
    lib().pdae_oracle_edge_conv_max(_p(x), _p(idx), _p(weight), _p(scale), _p(shift), float(slope), b, c, n, k, co, _p(out))
    return out


def edge_gather_extremum(p, q, idx, scale, shift, slope=0.2):
    """The gather half of edge_conv_max in the kernel's operation order: p, q (B,N,Co), idx (B,N,k) -> (B,Co,N)."""
    p, q, idx, scale, shift = _f32(p), _f32(q), _i64(idx), _f32(scale), _f32(shift)
    b, n, co = p.shape
    k = idx.shape[2]
    out = np.zeros((b, co, n), dtype=np.float32)
    lib().pdae_oracle_edge_gather_extremum.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_float] + [ctypes.c_int] * 4 + [ctypes.c_void_p]
    lib().pdae_oracle_edge_gather_extremum(_p(p), _p(q), _p(idx), _p(scale), _p(shift), float(slope), b, n, k, co, _p(out))
    return out
