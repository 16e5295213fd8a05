"""Run as a script (own process).  Builds the reference's flagship model `PointCAE_transformer` (the config of
cfgs/pretrain_PointCAE_transformer_dropout_patch_affine_r3_maskpatch.yaml with a narrower / shallower transformer so it
runs in seconds on CPU), UNMODIFIED from /root/reference, and runs one seeded training forward + backward on CPU in one
of four set-ups, printing one JSON object (loss, a gradient checksum, the host RNG positions afterwards):

  reference    the reference's own glue over oracle-backed stand-ins for its compiled / third-party modules
  install      pointdae_b200.install(): this repo's knn_cuda / pointnet2_utils / chamfer host layer (ops -> oracle)
  patched      + patch_models(): fused Group, one-launch corrupt_data, misc.fps
  patched_loss + install(loss_modules=True): the fused mean-loss autograd node instead of the reference's loss file
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODE = sys.argv[1]
REF = sys.argv[2] if len(sys.argv) > 2 else "/root/reference"
MODEL = sys.argv[3] if len(sys.argv) > 3 else "transformer"
for p in (REF, os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import _ref_stubs  # noqa: E402

torch.nn.Module.cuda = lambda self, *a, **k: self  # build_loss_func calls `.cuda()` on the loss module (:1000-1002)
torch.Tensor.cuda = lambda self, *a, **k: self     # ... and some forwards on fresh tensors (models/Point_M2AE.py:116)
if MODE == "reference":
    import types
    import _oracle_chamfer
    import _standins
    _standins.install_modules()
    sys.modules["chamfer"] = _oracle_chamfer
    sys.modules["pointnet2_ops"].__path__ = []
    _ref_stubs.stub("pointnet2_ops.pointnet2_modules")  # only the PointNet++ encoder file needs it
    p2 = types.ModuleType("pointnet2")
    p2.__path__ = []
    sys.modules["pointnet2"] = p2
    _ref_stubs.stub("pointnet2._ext")
else:
    import pointdae_b200
    import _oracle_ops
    pointdae_b200.install(loss_modules=(MODE == "patched_loss"))
    _oracle_ops.apply()
_ref_stubs.install_third_party()
models, _ = _ref_stubs.import_with_stubs("models")
patched = pointdae_b200.patch_models() if MODE in ("patched", "patched_loss") else []

from easydict import EasyDict  # noqa: E402
from pointdae_b200 import synth  # noqa: E402

cfg = EasyDict(NAME="PointCAE_transformer", corrupt_type=["affine_r3", "Drop-Patch"], all_patch="False", group_size=32,
               num_group=64, loss="cdl2",
               transformer_config=EasyDict(rand_ratio="True", mask_ratio=0.6, mask_type="rand", trans_dim=48,
                                           encoder_dims=48, depth=2, drop_path_rate=0.1, num_heads=2, decoder_depth=1,
                                           decoder_num_heads=2))
if MODEL == "dgcnn":
    # models/PointCAE_DGCNN.py:26-143: DGCNN encoder (get_graph_feature k=20 on 3 / 64 / 64 / 128 channels), folding decoder,
    # ChamferL1 on the coarse and the fine cloud, Drop-Patch corruption inside forward
    cfg = EasyDict(NAME="Point_CAE_DGCNN", corrupt_type=["dropout_patch_pointmae"], loss="cdl1")
if MODEL == "m2ae":
    # models/Point_M2AE.py: three-scale tokenizer (Group of models/Point_M2AE_modules.py, which returns the flattened
    # neighbour indices), corrupt_data on LISTS of patches / centres, ChamferL2 on the finest masked patches
    cfg = EasyDict(NAME="Point_M2AE", corrupt_type=["affine_r3", "Drop-Patch"], mask_ratio=0.8, group_sizes=[16, 8, 8],
                   num_groups=[128, 64, 16], encoder_depths=[1, 1, 1], encoder_dims=[24, 48, 96],
                   local_radius=[0.32, 0.64, 1.28], decoder_depths=[1, 1], decoder_dims=[96, 48], decoder_up_blocks=[1, 1],
                   drop_path_rate=0.1, num_heads=2)
if MODEL == "masksurf":
    # models/MaskSurf.py:342-488 (cfgs/pretrain_MaskSurf.yaml): xyz + normal input, the normal-aware `Group`, and
    # ChamferDistanceL2_withnormal, which reuses the Chamfer match indices to compare normals
    cfg.NAME, cfg.corrupt_type, cfg.loss = "MaskSurf", ["clean"], "cdl2normal"
if MODEL == "pointnetv2":
    # models/PointCAE_pointnetv2.py:62-174 (cfgs/pretrain_PointCAE_affine_r3_dropout_local_4xlonger.yaml): the PointNet++
    # encoder of models/pointnetv2_util.py:320-325 over pointnet2_ops.pointnet2_modules (un-vendored: drop-in modes only)
    cfg = EasyDict(NAME="Point_CAE_PointNetv2", corrupt_type=["dropout_patch_pointmae"], num_group=64, loss="cdl2")
if os.environ.get("PDAE_PROBE_NAME"):  # sweep helper: same configuration, another registered class of the same family
    cfg.NAME = os.environ["PDAE_PROBE_NAME"]
random.seed(0), np.random.seed(0), torch.manual_seed(0)
model = models.build_model_from_cfg(cfg)
pts = torch.from_numpy(synth.clouds(2 if MODEL in ("dgcnn", "pointnetv2") else 3, 1024, seed=9))
if MODEL == "masksurf":
    nrm = torch.nn.functional.normalize(torch.randn(pts.shape, generator=torch.Generator().manual_seed(3)), dim=2)
    pts = torch.cat([pts, nrm], dim=2)
if MODEL == "dgcnn" and os.environ.get("PDAE_PROBE_EVAL"):
    # feature extraction as the runners do it for the SVM evaluation: eval mode, no autograd, encoder only
    with torch.no_grad():
        for m in model.modules():  # trained-looking BatchNorm statistics, some negative scales
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.randn_like(m.weight)), m.bias.copy_(0.3 * torch.randn_like(m.bias))
                m.running_mean.copy_(0.2 * torch.randn_like(m.running_mean)), m.running_var.copy_(0.5 + torch.rand_like(m.running_var))
        model.eval()
        feat = model.dgcnn_encoder(pts.transpose(1, 2).contiguous())
    print(json.dumps({"mode": MODE, "feature_shape": list(feat.shape), "feature_abs_sum": float(feat.double().abs().sum()),
                      "feature_probe": [float(v) for v in feat.reshape(-1)[:6]], "feature_max": float(feat.abs().max()),
                      "encoder_forward": type(model.dgcnn_encoder).forward.__module__}))
    sys.exit(0)
random.seed(1), np.random.seed(1), torch.manual_seed(1)
model.train()
if MODEL == "masksurf":
    try:
        out = model(pts)  # MaskSurf.forward(pts, vis=False) -> (Chamfer term, normal term)
    except TypeError:
        out = model(pts, pts)  # the v2 classes take (corrupted_pts, pts)
    loss = sum(o for o in out if torch.is_tensor(o) and o.dim() == 0 and o.requires_grad)
else:
    loss = model(pts, pts)[0]
loss.backward()
grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
first = sorted(grads)[0]
print(json.dumps({
    "mode": MODE, "loss": float(loss), "n_params_with_grad": len(grads),
    "grad_abs_sum": float(sum(g.double().abs().sum() for g in grads.values())),
    "grad_probe": [float(v) for v in grads[first].reshape(-1)[:4]], "grad_probe_name": first,
    "group_class": (type(model.group_divider).__module__ if hasattr(model, "group_divider") else
                    type(model.group_dividers[0]).__module__ if hasattr(model, "group_dividers") else None),
    "loss_class": type(getattr(model, "loss_func", None) or getattr(model, "rec_loss")).__module__,
    "rng_after": [random.random(), float(np.random.rand()), float(torch.rand(1, dtype=torch.float64))],
    "patched": len(patched)}))
