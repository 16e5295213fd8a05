"""One pass of the round-2 kernels for ncu (profiles/r02/call_09_ncu.sh): EdgeConv layer 4 forward + backward (training
mode, 16 x 2048, k = 20), the Encoder's 512 -> 512 product over 262 144 points, DGCNN feature kNN C = 128."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import dgcnn_util, ops, synth  # noqa: E402

DEV = "cuda:0"
x = torch.from_numpy(synth.features(16, 128, 2048, seed=1)).to(DEV)
idx = dgcnn_util.knn(x, 20)
block = nn.Sequential(nn.Conv2d(256, 256, 1, bias=False), nn.BatchNorm2d(256), nn.LeakyReLU(0.2)).to(DEV).train()
for _ in range(2):
    xx = x.clone().requires_grad_(True)
    ops.edge_conv(xx, idx, block[0].weight, block[1], 0.2).sum().backward()
f = torch.randn(1, 128 * 64 * 32, 512, device=DEV)
w = torch.randn(512, 512, device=DEV)
for _ in range(2):
    ops.conv1x1(f, w, True, True)
torch.cuda.synchronize()
