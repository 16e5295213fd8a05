"""Stand-ins for the third-party packages the reference imports at module level but that have nothing to do with the
geometry path (easydict, timm, matplotlib, thop, ...), so that `import models` of the reference can be exercised on top
of `pointdae_b200.install()` in the build container.  TEST INFRASTRUCTURE ONLY.

`timm`'s two names the models call while being constructed (DropPath, trunc_normal_) and EasyDict are implemented;
everything else is an attribute-absorbing stub created on demand for exactly the module names that fail to import."""
import importlib
import sys
import types

import torch
import torch.nn as nn

OURS = ("pointnet2_ops", "knn_cuda", "chamfer", "pointnet2", "extensions", "models", "utils", "datasets", "tools",
        "segmentation", "pointdae_b200")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Stub(self.__name__ + "." + name)
        sub.__path__ = []
        setattr(self, name, sub)
        sys.modules[sub.__name__] = sub
        return sub

    def __call__(self, *a, **k):
        return self


def stub(name):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            m = _Stub(n)
            m.__path__ = []
            sys.modules[n] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)


class EasyDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = dict.__setitem__


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


def install_third_party():
    ed = types.ModuleType("easydict")
    ed.EasyDict = EasyDict
    sys.modules.setdefault("easydict", ed)
    if "timm" not in sys.modules:
        try:
            importlib.import_module("timm")
        except Exception:
            timm, tm, tl = types.ModuleType("timm"), types.ModuleType("timm.models"), types.ModuleType("timm.models.layers")
            timm.__path__, tm.__path__ = [], []
            tl.DropPath, tl.trunc_normal_ = DropPath, torch.nn.init.trunc_normal_
            timm.models, tm.layers = tm, tl
            sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})


def import_with_stubs(name, limit=60):
    """import `name`, stubbing every missing third-party module it trips over; returns (module, [stubbed names])."""
    stubbed = []
    for _ in range(limit):
        try:
            return importlib.import_module(name), stubbed
        except ModuleNotFoundError as e:
            missing = e.name or ""
            if not missing or missing.split(".")[0] in OURS:
                raise
            stub(missing)
            stubbed.append(missing)
            for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                if isinstance(sys.modules[k], _Stub):
                    continue
                del sys.modules[k]  # half-initialised modules of the failed attempt
    raise RuntimeError("too many missing modules: %s" % stubbed)
