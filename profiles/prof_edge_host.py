import cProfile, pstats, sys, os, torch, torch.nn as nn
sys.path.insert(0, os.getcwd())
from pointdae_b200 import dgcnn_util, ops
x = torch.randn(16, 64, 2048, device="cuda:0"); idx = dgcnn_util.knn(x, 20)
block = nn.Sequential(nn.Conv2d(128, 64, 1, bias=False), nn.BatchNorm2d(64), nn.LeakyReLU(0.2)).cuda().train()
def f():
    with torch.no_grad():
        for _ in range(200): ops.edge_conv(x, idx, block[0].weight, block[1], 0.2)
    torch.cuda.synchronize()
f()
import time; t=time.perf_counter(); f(); print("per call ms", (time.perf_counter()-t)/200*1e3)
cProfile.run("f()", "/tmp/p.out"); pstats.Stats("/tmp/p.out").sort_stats("cumtime").print_stats(18)
