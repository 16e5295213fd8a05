"""Run as a script (own process: the reference's top-level package names `models`, `utils`, `datasets` must not leak
into the test session).  Imports the reference's whole `models` package on top of pointdae_b200.install(), applies
patch_models(), builds the reference's PointNet++ encoder on CPU, and prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
LOSS_MODULES = (sys.argv[2] == "1") if len(sys.argv) > 2 else False
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, REF)

import pointdae_b200  # noqa: E402
import _ref_stubs  # noqa: E402

pointdae_b200.install(loss_modules=LOSS_MODULES)
_ref_stubs.install_third_party()
models, stubbed = _ref_stubs.import_with_stubs("models")
patched = pointdae_b200.patch_models()

import pointnet2_ops.pointnet2_utils as p2u  # noqa: E402  (what the reference's `from pointnet2_ops import ...` sees)
from pointdae_b200 import chamfer_dist, group, knn_cuda, pointnet2_utils  # noqa: E402
import models.PointCAE_transformer as pct  # noqa: E402
import models.pointnetv2_util as pv2  # noqa: E402
import extensions.chamfer_dist as ref_cd  # noqa: E402
import utils.misc as misc  # noqa: E402
import datasets.corrupt_util_tensor as cut  # noqa: E402

enc = pv2.PointNetv2_encoder()
out = {
    "stubbed": stubbed,
    "patched": patched,
    "pointnet2_utils_is_ours": p2u is pointnet2_utils,
    "knn_is_ours": pct.KNN is knn_cuda.KNN,
    "group_is_ours": pct.Group is group.Group,
    "loss_class_module": ref_cd.ChamferDistanceL2.__module__,
    "loss_is_ours": ref_cd.ChamferDistanceL2 is chamfer_dist.ChamferDistanceL2,
    "misc_fps_is_ours": misc.fps is group.fps,
    "module_level_knn": type(cut.knn).__module__ if hasattr(cut, "knn") else None,
    "encoder_keys": sorted(enc.state_dict().keys())[:6],
    "encoder_params": sum(p.numel() for p in enc.parameters()),
    "n_model_modules": len([m for m in sys.modules if m.startswith("models.")]),
}
print(json.dumps(out))
