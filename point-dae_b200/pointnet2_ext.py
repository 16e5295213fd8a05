"""Drop-in for the compiled `pointnet2._ext` module (extensions/pointnet2/_ext_src/src/bindings.cpp:9-22),
the functions on the hot path and its "next" rows.  three_nn / three_interpolate are only used by the
part-segmentation FP modules (out of scope) and raise NotImplementedError."""
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


def _out_of_scope(name):
    def f(*a, **k):
        raise NotImplementedError("pointnet2._ext.%s is outside the geometry hot path (FP modules only)" % name)
    return f


three_nn = _out_of_scope("three_nn")
three_interpolate = _out_of_scope("three_interpolate")
three_interpolate_grad = _out_of_scope("three_interpolate_grad")
