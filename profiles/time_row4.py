"""Row 4 timings on one B200 (CUDA events, median of 5): the four EdgeConv layers of the DGCNN encoder at the C3 shape
(16 clouds x 2048 points per GPU, k = 20) and the patch Encoder at the C2 / headline shape (128 x 64 patches x 32 points),
forward and forward+backward, this repo's tensor-core route against the reference's layer sequence run by torch on the same
GPU (a) with torch's defaults (cuDNN / cuBLAS may use TF32 for the convolutions: 1e-3 accuracy) and (b) in true fp32.
    python profiles/time_row4.py > gpurun_out/r02/time_row4.json"""
import json
import os
import statistics
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pointdae_b200 import dgcnn_util, encoder, ops  # noqa: E402
from test_gpu_edgeconv_tc import reference_layer  # noqa: E402
from test_gpu_encoder_tc import Encoder  # noqa: E402

DEV = "cuda:0"


def ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def with_tf32(flag, fn):
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = flag
    torch.backends.cuda.matmul.allow_tf32 = flag
    try:
        return fn()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


out = {"edgeconv (B=16, N=2048, k=20)": {}, "encoder": {}}
b, n, k = 16, 2048, 20
for c, co in ((3, 64), (64, 64), (64, 128), (128, 256)):
    x = torch.randn(b, c, n, device=DEV)
    idx = dgcnn_util.knn(x, k)
    block = nn.Sequential(nn.Conv2d(2 * c, co, 1, bias=False), nn.BatchNorm2d(co), nn.LeakyReLU(0.2)).to(DEV).train()
    up = torch.randn(b, co, n, device=DEV)

    def ours(bwd):
        xo = x.clone().requires_grad_(bwd)
        y = ops.edge_conv(xo, idx, block[0].weight, block[1], 0.2)
        if bwd:
            (y * up).sum().backward()

    def ref(bwd):
        xr = x.clone().requires_grad_(bwd)
        y = reference_layer(xr, idx, block)
        if bwd:
            (y * up).sum().backward()

    rec = {"ours_fwd_ms": ms(lambda: ours(False)), "ours_fwd_bwd_ms": ms(lambda: ours(True)),
           "reference_default_tf32_fwd_ms": ms(lambda: ref(False)), "reference_default_tf32_fwd_bwd_ms": ms(lambda: ref(True)),
           "reference_true_fp32_fwd_ms": with_tf32(False, lambda: ms(lambda: ref(False))),
           "reference_true_fp32_fwd_bwd_ms": with_tf32(False, lambda: ms(lambda: ref(True)))}
    with torch.no_grad():
        z = ops.conv1x1(x, torch.randn(2 * co, c, device=DEV), out_point_major=True)
        rec["tensor_core_product_ms"] = ms(lambda: ops.conv1x1(x, torch.randn(2 * co, c, device=DEV), out_point_major=True))
        rec["product_tflops_3xtf32"] = 3 * 2.0 * b * n * c * 2 * co / (rec["tensor_core_product_ms"] * 1e-3) / 1e12
    rec["speedup_fwd_bwd_vs_default"] = rec["reference_default_tf32_fwd_bwd_ms"] / rec["ours_fwd_bwd_ms"]
    rec["graph_feature_bytes_never_written"] = b * n * k * (2 * c + co) * 4
    out["edgeconv (B=16, N=2048, k=20)"]["C=%d -> Co=%d" % (c, co)] = rec
    del x, idx, up, z

for name, (bs, g, npts, ch) in {"H / C2: 128 x 64 patches x 32 points -> 384": (128, 64, 32, 384)}.items():
    torch.manual_seed(0)
    enc_r, enc_o = Encoder(ch).to(DEV).train(), Encoder(ch).to(DEV).train()
    pts = torch.randn(bs, g, npts, 3, device=DEV) * 0.3

    def run(mod, fwd_fn, bwd):
        p = pts.clone().requires_grad_(bwd)
        y = fwd_fn(mod, p)
        if bwd:
            y.sum().backward()

    rec = {"ours_fwd_ms": ms(lambda: run(enc_o, encoder.encoder_forward, False)),
           "ours_fwd_bwd_ms": ms(lambda: run(enc_o, encoder.encoder_forward, True)),
           "reference_default_tf32_fwd_ms": ms(lambda: run(enc_r, lambda m, p: m(p), False)),
           "reference_default_tf32_fwd_bwd_ms": ms(lambda: run(enc_r, lambda m, p: m(p), True)),
           "reference_true_fp32_fwd_ms": with_tf32(False, lambda: ms(lambda: run(enc_r, lambda m, p: m(p), False))),
           "reference_true_fp32_fwd_bwd_ms": with_tf32(False, lambda: ms(lambda: run(enc_r, lambda m, p: m(p), True)))}
    w = torch.randn(512, 512, device=DEV)
    f = torch.randn(1, bs * g * npts, 512, device=DEV)
    t = ms(lambda: ops.conv1x1(f, w, True, True))
    rec["conv 512->512 over 262144 points"] = {"ms": t, "tflops_3xtf32_executed": 3 * 2.0 * bs * g * npts * 512 * 512 / (t * 1e-3) / 1e12,
                                                "tflops_fp32_equivalent": 2.0 * bs * g * npts * 512 * 512 / (t * 1e-3) / 1e12}
    out["encoder"][name] = rec
print(json.dumps(out, indent=1))
