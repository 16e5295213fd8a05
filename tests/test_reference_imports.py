"""Drop-in at the import level: the reference's whole `models` package (models/__init__.py imports every model file)
must import on top of `pointdae_b200.install()` with nothing but unrelated third-party packages stubbed, and
`patch_models()` must rebind the pure-torch hot functions inside it.  Needs /root/reference (build container only;
skipped on the GPU box, and never part of the `-m gpu` run)."""
import json
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not present")

NOT_OURS = {"ipdb", "matplotlib", "mpl_toolkits", "termcolor", "h5py", "timm", "thop", "ptflops", "open3d", "cv2",
            "tensorboardX", "sklearn", "tqdm", "scipy", "yaml", "transforms3d", "trimesh", "plyfile", "pandas", "PIL"}


def probe(loss_modules):
    res = subprocess.run([sys.executable, os.path.join(HERE, "_ref_import_probe.py"), REF, "1" if loss_modules else "0"],
                         capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert res.returncode == 0, res.stderr[-3000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("loss_modules", [False, True])
def test_reference_models_package_imports_on_top_of_install(loss_modules):
    out = probe(loss_modules)
    assert {s.split(".")[0] for s in out["stubbed"]} <= NOT_OURS, out["stubbed"]  # only unrelated third parties were missing
    assert out["pointnet2_utils_is_ours"] and out["knn_is_ours"] and out["group_is_ours"] and out["misc_fps_is_ours"]
    assert out["module_level_knn"] == "pointdae_b200.knn_cuda"  # KNN(k=32) built at import time (corrupt_util_tensor.py:591)
    assert out["n_model_modules"] >= 20
    # the reference's own loss file runs on our `chamfer` unless install(loss_modules=True) serves the mirror
    assert out["loss_is_ours"] == loss_modules
    for name in ("models.dgcnn_util.knn", "models.dgcnn_util.get_graph_feature", "models.PointCAE_DGCNN.get_graph_feature",
                 "utils.misc.fps", "datasets.corrupt_util_tensor.corrupt_data", "models.PointCAE_transformer.corrupt_data",
                 "models.PointCAE_transformer.Group", "models.Point_M2AE_modules.Group", "models.MaskSurf.Group",
                 "models.MaskSurf_v2.Group", "models.PointCAE_pointnetv2.Group"):
        assert name in out["patched"], name
    # models/pointnetv2_util.py:323-325 builds on pointnet2_ops.pointnet2_modules: checkpoint-compatible parameter names
    assert out["encoder_keys"][0].startswith("sa1.mlps.0.0.") and out["encoder_params"] == 805184  # 3->64->64->128, 131->128->128->256, 259->256->512->1024 with BatchNorm, no conv bias
