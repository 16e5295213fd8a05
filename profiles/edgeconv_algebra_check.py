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
