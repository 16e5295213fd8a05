"""A CPU stand-in for the compiled `pointnet2._ext` module (extensions/pointnet2/_ext_src/src/bindings.cpp:9-22),
backed by the C oracle.  TEST INFRASTRUCTURE ONLY: it lets the reference's own `extensions/pointnet2/pointnet2_utils.py`
run in a CPU-only container to produce golden vectors, and lets the CPU suite run this repo's host-side classes
(QueryAndGroup, GroupAll, the autograd glue) over the same functions."""
import torch

from oracle import cpu as oracle


def _t(a):
    return torch.from_numpy(a)


def furthest_point_sampling(xyz, npoint):
    return _t(oracle.fps(xyz.detach().numpy(), int(npoint)))


def gather_points(features, idx):
    return _t(oracle.gather(features.detach().numpy(), idx.numpy()))


def gather_points_grad(grad_out, idx, n):
    return _t(oracle.gather_grad(grad_out.detach().numpy(), idx.numpy(), int(n)))


def ball_query(new_xyz, xyz, radius, nsample):
    return _t(oracle.ball_query(float(radius), int(nsample), xyz.detach().numpy(), new_xyz.detach().numpy()))


def group_points(points, idx):
    return _t(oracle.group_points(points.detach().numpy(), idx.numpy()))


def group_points_grad(grad_out, idx, n):
    return _t(oracle.group_points_grad(grad_out.detach().numpy(), idx.numpy(), int(n)))


def three_nn(unknown, known):
    d, i = oracle.three_nn(unknown.detach().numpy(), known.detach().numpy())
    return _t(d), _t(i)


def three_interpolate(points, idx, weight):
    return _t(oracle.three_interpolate(points.detach().numpy(), idx.numpy(), weight.detach().numpy()))


def three_interpolate_grad(grad_out, idx, weight, m):
    return _t(oracle.three_interpolate_grad(grad_out.detach().numpy(), idx.numpy(), weight.detach().numpy(), int(m)))


# this repo's `ops` names for the same functions (pointdae_b200.pointnet2_utils calls these)
furthest_point_sample = furthest_point_sampling
