"""Seeded inputs for every loss class of extensions/chamfer_dist/__init__.py and the one routine that runs a
loss module on them; shared by the golden generator (reference classes) and the tests (this repo's)."""
import numpy as np
import torch


def _normals(rng, b, n):
    v = rng.standard_normal(size=(b, n, 3)).astype(np.float32)
    return v * rng.uniform(0.5, 2.0, size=(b, n, 1)).astype(np.float32)  # deliberately not unit length


def cases(synth):
    """name -> (class name, ordered dict of float32 input arrays in call order)"""
    rng = np.random.default_rng(synth.BASE_SEED + 4242)
    gt = synth.clouds(3, 160, seed=41)
    pred = synth.prediction(gt, seed=41)[:, :128]  # ragged: 128 vs 160 points
    n_pred, n_gt = _normals(rng, 3, 128), _normals(rng, 3, 160)
    c_pred, c_gt = rng.uniform(0, 1, (3, 128, 1)).astype(np.float32), rng.uniform(0, 1, (3, 160, 1)).astype(np.float32)
    p_pred, p_gt = rng.standard_normal((3, 128, 3)).astype(np.float32), rng.standard_normal((3, 160, 3)).astype(np.float32)
    plain = dict(xyz1=pred, xyz2=gt)
    withn = dict(xyz1=pred, xyz2=gt, normal_rebuild=n_pred, normal_gt=n_gt)
    six1 = np.concatenate([pred, n_pred], axis=2)
    six2 = np.concatenate([gt, n_gt], axis=2)
    coarse1, coarse2 = synth.clouds(4, 16, seed=42), synth.clouds(4, 16, seed=43)
    fine1 = (coarse1[:, :, None, :] + 0.05 * rng.standard_normal((4, 16, 12, 3))).astype(np.float32)
    fine2 = (coarse2[:, :, None, :] + 0.05 * rng.standard_normal((4, 16, 12, 3))).astype(np.float32)
    one = synth.clouds(1, 96, seed=44)
    one_z = one.copy()
    one_z[0, ::7] = 0.0  # rows dropped by ignore_zeros (batch of one)
    return {
        "l2": ("ChamferDistanceL2", plain),
        "l2_split": ("ChamferDistanceL2_split", plain),
        "l1": ("ChamferDistanceL1", plain),
        "l2_ignore_zeros": ("ChamferDistanceL2", dict(xyz1=one_z, xyz2=one[:, ::-1].copy())),
        "coarse2fine": ("ChamferDistanceL2_corase2fine", dict(xyz1=coarse1, xyz2=coarse2, fine1=fine1, fine2=fine2)),
        "withnormal": ("ChamferDistanceL2_withnormal", withn),
        "withnormal_curve": ("ChamferDistanceL2_withnormal", dict(withn, curve_rebuild=c_pred, curve_gt=c_gt)),
        "withnormal_curve_pos": ("ChamferDistanceL2_withnormal",
                                 dict(withn, curve_rebuild=c_pred, curve_gt=c_gt, position_rebuild=p_pred,
                                      position_gt=p_gt)),
        "withnormal_visual": ("ChamferDistanceL2_withnormal_visual", withn),
        "withnormal_l1": ("ChamferDistanceL2_withnormalL1", withn),
        "withnormal_strict": ("ChamferDistanceL2_withnormal_strict", withn),
        "strict_normalindex": ("ChamferDistanceL2_withnormal_strict_normalindex", dict(xyz1=six1, xyz2=six2)),
        "normalindex": ("ChamferDistanceL2_withnormal_normalindex", withn),
        "onlynormalindex": ("ChamferDistanceL2_withnormal_onlynormalindex", dict(xyz1=six1, xyz2=six2)),
    }


def run(module, arrays, device):
    """-> (list of returned values as numpy, dict input name -> gradient of sum(weights_i * scalar outputs))"""
    if getattr(module, "ignore_zeros", None) is not None and arrays["xyz1"].shape[0] == 1:
        module.ignore_zeros = True
    tensors = {k: torch.from_numpy(v).to(device).requires_grad_(True) for k, v in arrays.items()}
    res = module(*tensors.values())
    res = list(res) if isinstance(res, (tuple, list)) else [res]
    total = None
    for i, r in enumerate(res):
        if r.requires_grad:
            term = (i + 1.0) * r.sum() if r.dim() == 0 or r.numel() == 1 else (i + 1.0) * r.mean()
            total = term if total is None else total + term
    grads = {}
    if total is not None:
        total.backward()
        grads = {k: t.grad.detach().cpu().numpy() for k, t in tensors.items() if t.grad is not None}
    return [r.detach().cpu().numpy() for r in res], grads
