"""get_graph_feature forward (16 x 2048, k = 20): row-per-CTA kernel vs the element-wise one, against the measured HBM peak."""
import json, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointdae_b200 import dgcnn_util, ops, synth
HBM = 6556.5
out = {}
for C in (64, 128):
    x = torch.from_numpy(synth.features(16, C, 2048, seed=C)).to("cuda:0")
    idx = dgcnn_util.knn(x, 20)
    nbytes = 16 * 2048 * 20 * 2 * C * 4 + 16 * C * 2048 * 4 + 16 * 2048 * 20 * 8
    rec = {}
    res = {}
    for name, env in (("row_per_cta", None), ("elementwise_round1", "1")):
        if env: os.environ["PDAE_GRAPHFEAT_ELEMENTWISE"] = env
        else: os.environ.pop("PDAE_GRAPHFEAT_ELEMENTWISE", None)
        for _ in range(3): ops._graph_feature_fwd(x, idx)
        torch.cuda.synchronize(); ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = ops._graph_feature_fwd(x, idx); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ms = statistics.median(ts); res[name] = r
        rec[name] = {"ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "hbm_frac": nbytes / (ms * 1e-3) / 1e9 / HBM}
    rec["same_bits"] = bool(torch.equal(res["row_per_cta"], res["elementwise_round1"]))
    out["C=%d" % C] = rec
os.environ.pop("PDAE_GRAPHFEAT_ELEMENTWISE", None)
print(json.dumps(out))
