"""Loads the reference's own CUDA ops, rebuilt unmodified for sm_100a by oracle/build_ref.py
(oracle/_ref/, git-ignored but shipped to the GPU box).  Test infrastructure only."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_cache = {}


def _load(name, path):
    if name in _cache:
        return _cache[name]
    if not os.path.exists(path):
        _cache[name] = None
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def ref_chamfer():
    """module with forward(xyz1, xyz2) / backward(...) -- extensions/chamfer_dist/chamfer_cuda.cpp:36-39"""
    return _load("chamfer", os.path.join(ROOT, "oracle", "_ref", "chamfer", "chamfer.so"))


def ref_pointnet2():
    """module `_ext` -- extensions/pointnet2/_ext_src/src/bindings.cpp:9-22"""
    return _load("_ext", os.path.join(ROOT, "oracle", "_ref", "pointnet2_ext", "_ext.so"))
