import os, sys
sys.path.insert(0, "/root/repo")
import torch
from pointdae_b200 import synth, dgcnn_util
x = torch.from_numpy(synth.features(16, 64, 2048, seed=64)).cuda()
for _ in range(3): dgcnn_util.knn(x, 20)
torch.cuda.synchronize()
