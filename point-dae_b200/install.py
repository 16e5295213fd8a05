"""`install()` makes the reference's imports resolve to this package, so Point-DAE's models/ and
main.py run unchanged:

    import pointdae_b200; pointdae_b200.install()
    # now: from pointnet2_ops import pointnet2_utils (and pointnet2_ops.pointnet2_modules) ; from knn_cuda import KNN ; import chamfer ;
    #      import pointnet2._ext  -> all served by the sm_100a kernels

Only the *compiled / third-party* modules are replaced (the reference's own Python, e.g.
extensions/chamfer_dist/__init__.py, keeps running on top).  `patch_models()` additionally rebinds
the pure-torch hot functions that live inside the reference's packages (dgcnn knn / get_graph_feature,
misc.fps, Group) once those packages are importable.
"""
import importlib
import sys
import types


def install(loss_modules=False):
    """loss_modules=True additionally serves `extensions.chamfer_dist` (the reference's Python loss classes,
    which import `ipdb` at module level, extensions/chamfer_dist/__init__.py:12) from this package's mirror."""
    from . import chamfer, knn_cuda, pointnet2_ext, pointnet2_modules, pointnet2_utils

    pkg = types.ModuleType("pointnet2_ops")
    pkg.__path__ = []  # mark as package
    pkg.pointnet2_utils = pointnet2_utils
    pkg.pointnet2_modules = pointnet2_modules  # models/pointnetv2_util.py:317, pulled in by `import models`
    pkg.__version__ = "3.0.0"
    sys.modules["pointnet2_ops"] = pkg
    sys.modules["pointnet2_ops.pointnet2_utils"] = pointnet2_utils
    sys.modules["pointnet2_ops.pointnet2_modules"] = pointnet2_modules

    sys.modules["knn_cuda"] = knn_cuda
    sys.modules["chamfer"] = chamfer

    # `import pointnet2._ext as _ext` (extensions/pointnet2/pointnet2_utils.py:23-24)
    p2 = sys.modules.get("pointnet2")
    if p2 is None:
        p2 = types.ModuleType("pointnet2")
        p2.__path__ = []
        sys.modules["pointnet2"] = p2
    p2._ext = pointnet2_ext
    sys.modules["pointnet2._ext"] = pointnet2_ext

    if loss_modules:
        from . import chamfer_dist
        ext = sys.modules.get("extensions")
        if ext is None:
            try:
                ext = importlib.import_module("extensions")
            except Exception:
                ext = types.ModuleType("extensions")
                ext.__path__ = []
                sys.modules["extensions"] = ext
        ext.chamfer_dist = chamfer_dist
        sys.modules["extensions.chamfer_dist"] = chamfer_dist
    return True


def patch_models(names=("models.dgcnn_util", "models.PointCAE_DGCNN", "segmentation.models.dgcnn_util")):
    """Rebind knn / get_graph_feature in the reference's already-importable modules, misc.fps, and Group."""
    from . import dgcnn_util, group

    patched = []
    for name in names:
        try:
            mod = importlib.import_module(name)
        except Exception:
            continue
        for fn in ("knn", "get_graph_feature"):
            if hasattr(mod, fn):
                setattr(mod, fn, getattr(dgcnn_util, fn))
                patched.append(name + "." + fn)
    try:
        misc = importlib.import_module("utils.misc")
        misc.fps = group.fps
        patched.append("utils.misc.fps")
    except Exception:
        pass
    try:  # the Drop-Patch corruption run inside forward (datasets/corrupt_util_tensor.py:592-616)
        from . import corrupt_util_tensor
        cut = importlib.import_module("datasets.corrupt_util_tensor")
        cut.dropout_patch_random = corrupt_util_tensor.dropout_patch_random
        patched.append("datasets.corrupt_util_tensor.dropout_patch_random")
        # the affine corruptions between the patchifier and the encoder (:59-343, :706-728): one launch per chain
        for fn in ("corrupt_data", "corrupt_scale_nonorm", "corrupt_tranlate", "corrupt_rotate_360",
                   "corrupt_rotate_z_360", "corrupt_reflection", "corrupt_shear"):
            setattr(cut, fn, getattr(corrupt_util_tensor, fn))
        cut.corruptions.update(corrupt_util_tensor.corruptions)
        patched.append("datasets.corrupt_util_tensor.corrupt_data")
        for name in ("models.PointCAE_transformer", "models.Point_M2AE"):  # `from ... import corrupt_data`
            mod = sys.modules.get(name)
            if mod is not None and hasattr(mod, "corrupt_data"):
                mod.corrupt_data = corrupt_util_tensor.corrupt_data
                patched.append(name + ".corrupt_data")
    except Exception:
        pass
    # every flavour of the patchifier class the reference defines, by the module that defines it
    flavours = {"models.PointCAE_transformer": group.Group, "models.Point_MAE": group.Group,
                "models.Point_MlMAE": group.Group, "models.PointCAE_pointnetv2": group.Group,
                "models.Point_M2AE_modules": group.GroupWithIndex, "models.MaskSurf": group.GroupNormal,
                "models.MaskSurf_v2": group.GroupAttribute, "models.MaskFeat_transformer": group.GroupAttribute,
                "models.MaskFeat_DGCNN": group.GroupAttribute}
    for name, cls in flavours.items():
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "Group"):
            mod.Group = cls
            patched.append(name + ".Group")
    return patched
