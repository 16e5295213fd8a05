"""Design check for SURVEY.md 8f row 4 (next round), CPU only: one EdgeConv layer of the reference's DGCNN encoder in
eval mode -- get_graph_feature -> Conv2d(2C, Co, 1, bias=False) -> BatchNorm2d -> LeakyReLU(0.2) -> max over k
(models/dgcnn_util.py:96-128) -- equals

    P = W1 @ x, Q = (W2 - W1) @ x                      two (Co x C) x (C x N) GEMMs per cloud
    y[o, i] = act(s_o * (ext_j P[o, idx[i, j]] + Q[o, i]) + t_o),  ext = max where s_o >= 0, min where s_o < 0

because W @ [x_j - x_i; x_i] = W1 x_j + (W2 - W1) x_i and BatchNorm (eval) + LeakyReLU is monotone per channel.
The (B, 2C, N, k) tensor is never formed: traffic drops from B*N*k*2C*4 bytes written + read to B*N*k*Co*4 bytes of
gathered (L2-resident) reads.  Prints the largest deviation from the straightforward evaluation."""
import json

import torch
import torch.nn as nn

torch.manual_seed(0)
out = {}
for C, Co, N, k in [(3, 64, 512, 20), (64, 64, 512, 20), (64, 128, 256, 20), (128, 256, 256, 20)]:
    B = 2
    x = torch.randn(B, C, N)
    idx = torch.stack([torch.stack([torch.randperm(N)[:k] for _ in range(N)]) for _ in range(B)])          # (B, N, k)
    conv, bn, act = nn.Conv2d(2 * C, Co, 1, bias=False), nn.BatchNorm2d(Co), nn.LeakyReLU(0.2)
    with torch.no_grad():
        bn.weight.copy_(torch.randn(Co))            # negative scales included
        bn.bias.copy_(torch.randn(Co))
        bn.running_mean.copy_(torch.randn(Co) * 0.3)
        bn.running_var.copy_(torch.rand(Co) + 0.5)
    bn.eval()
    with torch.no_grad():
        xt = x.transpose(1, 2)                                                                               # (B, N, C)
        nbr = torch.gather(xt.unsqueeze(1).expand(-1, N, -1, -1), 2, idx.unsqueeze(-1).expand(-1, -1, -1, C))  # (B, N, k, C)
        feat = torch.cat((nbr - xt.unsqueeze(2), xt.unsqueeze(2).expand(-1, -1, k, -1)), dim=3).permute(0, 3, 1, 2)
        want = act(bn(conv(feat))).max(dim=-1)[0]                                                            # (B, Co, N)
        W = conv.weight.view(Co, 2 * C)
        W1, W2 = W[:, :C], W[:, C:]
        P, Q = torch.matmul(W1, x), torch.matmul(W2 - W1, x)                                                 # (B, Co, N)
        s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        t = bn.bias - s * bn.running_mean
        g = torch.gather(P.unsqueeze(2).expand(-1, -1, N, -1), 3, idx.unsqueeze(1).expand(-1, Co, -1, -1))    # (B, Co, N, k)
        ext = torch.where((s >= 0).view(1, Co, 1), g.max(dim=3)[0], g.min(dim=3)[0])
        got = act(s.view(1, Co, 1) * (ext + Q) + t.view(1, Co, 1))
    scale = float(want.abs().max())
    out["C=%d Co=%d" % (C, Co)] = {"max_abs_diff": float((got - want).abs().max()), "output_scale": scale,
                                  "relative_to_scale": float((got - want).abs().max()) / scale}
print(json.dumps(out, indent=1))

# ---- training mode: the batch statistics of the convolution output, from per-point gather sums ------------------------
# y[o, i, j] = P[o, idx[i, j]] + Q[o, i].  With S1[o, i] = sum_j P[o, idx[i, j]] and S2[o, i] = sum_j P[o, idx[i, j]]^2
#     mean_o = (sum_i S1 + k sum_i Q) / (B N k)
#     E[y^2]_o = (sum_i S2 + 2 sum_i Q S1 + k sum_i Q^2) / (B N k)
# so BatchNorm's batch mean / variance need one more gather pass (sum and sum of squares next to max / min) and a
# (B N)-long reduction, still without the (B, Co, N, k) tensor; the forward then is the eval formula with these statistics.
train = {}
for C, Co, N, k in [(3, 64, 256, 20), (64, 128, 256, 20)]:
    B = 2
    x = torch.randn(B, C, N)
    idx = torch.stack([torch.stack([torch.randperm(N)[:k] for _ in range(N)]) for _ in range(B)])
    conv, bn, act = nn.Conv2d(2 * C, Co, 1, bias=False), nn.BatchNorm2d(Co), nn.LeakyReLU(0.2)
    with torch.no_grad():
        bn.weight.copy_(torch.randn(Co)), bn.bias.copy_(torch.randn(Co))
    bn.train()
    with torch.no_grad():
        xt = x.transpose(1, 2)
        nbr = torch.gather(xt.unsqueeze(1).expand(-1, N, -1, -1), 2, idx.unsqueeze(-1).expand(-1, -1, -1, C))
        feat = torch.cat((nbr - xt.unsqueeze(2), xt.unsqueeze(2).expand(-1, -1, k, -1)), dim=3).permute(0, 3, 1, 2)
        want = act(bn(conv(feat))).max(dim=-1)[0]
        W = conv.weight.view(Co, 2 * C)
        P, Q = torch.matmul(W[:, :C], x).double(), torch.matmul(W[:, C:] - W[:, :C], x).double()
        g = torch.gather(P.unsqueeze(2).expand(-1, -1, N, -1), 3, idx.unsqueeze(1).expand(-1, Co, -1, -1))
        S1, S2 = g.sum(dim=3), (g * g).sum(dim=3)
        cnt = B * N * k
        mean = (S1.sum(dim=(0, 2)) + k * Q.sum(dim=(0, 2))) / cnt
        ey2 = (S2.sum(dim=(0, 2)) + 2 * (Q * S1).sum(dim=(0, 2)) + k * (Q * Q).sum(dim=(0, 2))) / cnt
        var = ey2 - mean * mean  # biased, as BatchNorm normalises with
        s = bn.weight.double() / torch.sqrt(var + bn.eps)
        t = bn.bias.double() - s * mean
        ext = torch.where((s >= 0).view(1, Co, 1), g.max(dim=3)[0], g.min(dim=3)[0])
        got = act((s.view(1, Co, 1) * (ext + Q) + t.view(1, Co, 1)).float())
    scale = float(want.abs().max())
    train["C=%d Co=%d" % (C, Co)] = {"max_abs_diff": float((got - want).abs().max()), "relative_to_scale": float((got - want).abs().max()) / scale}
print(json.dumps({"training_mode_batch_statistics_from_gather_sums": train}, indent=1))
