"""Drop-in for the compiled `pointnet2._ext` module (extensions/pointnet2/_ext_src/src/bindings.cpp:9-22),
all nine functions it exports, argument order as in the C++ bindings."""
from . import ops


def furthest_point_sampling(points, nsamples):
    return ops.furthest_point_sample(points, nsamples)


def gather_points(points, idx):
    return ops.gather_points(points, idx)


def gather_points_grad(grad_out, idx, n):
    return ops.gather_points_grad(grad_out, idx, n)


def ball_query(new_xyz, xyz, radius, nsample):
    return ops.ball_query(new_xyz, xyz, radius, nsample)


def group_points(points, idx):
    return ops.group_points(points, idx)


def group_points_grad(grad_out, idx, n):
    return ops.group_points_grad(grad_out, idx, n)


def three_nn(unknowns, knows):
    return ops.three_nn(unknowns, knows)


def three_interpolate(points, idx, weight):
    return ops.three_interpolate(points, idx, weight)


def three_interpolate_grad(grad_out, idx, weight, m):
    return ops.three_interpolate_grad(grad_out, idx, weight, m)
