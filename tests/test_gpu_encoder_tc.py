"""Patch `Encoder` (mini-PointNet, models/PointCAE_transformer.py:20-51) with its four 1x1 convolutions on the tensor cores
(pointdae_b200.encoder.encoder_forward) against the reference's forward on the same module (torch, true fp32): output,
BatchNorm buffers, input and parameter gradients; training and eval."""
import pytest
import torch
import torch.nn as nn

from pointdae_b200 import encoder

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Encoder(nn.Module):  # the reference's module: same layers, same forward (restated, not imported: no /root/reference on the box)
    def __init__(self, encoder_channel):
        super().__init__()
        self.encoder_channel = encoder_channel
        self.first_conv = nn.Sequential(nn.Conv1d(3, 128, 1), nn.BatchNorm1d(128), nn.ReLU(inplace=True), nn.Conv1d(128, 256, 1))
        self.second_conv = nn.Sequential(nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
                                         nn.Conv1d(512, self.encoder_channel, 1))

    def forward(self, point_groups):
        bs, g, n, _ = point_groups.shape
        point_groups = point_groups.reshape(bs * g, n, 3)
        feature = self.first_conv(point_groups.transpose(2, 1))
        feature_global = torch.max(feature, dim=2, keepdim=True)[0]
        feature = torch.cat([feature_global.expand(-1, -1, n), feature], dim=1)
        feature = self.second_conv(feature)
        feature_global = torch.max(feature, dim=2, keepdim=False)[0]
        return feature_global.reshape(bs, g, self.encoder_channel)


@pytest.mark.parametrize("bs,g,n,ch,train", [(2, 16, 32, 384, True), (3, 5, 32, 96, True), (2, 7, 17, 40, False)])
def test_encoder_forward_backward(monkeypatch, bs, g, n, ch, train):
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    torch.manual_seed(bs * 10 + g)
    ref, ours = Encoder(ch).to(DEV), Encoder(ch).to(DEV)
    ours.load_state_dict(ref.state_dict())
    ref.train(train)
    ours.train(train)
    x = torch.randn(bs, g, n, 3, device=DEV) * 0.3
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    want = ref(xr)
    got = encoder.encoder_forward(ours, xo)
    scale = float(want.abs().max())
    assert tuple(got.shape) == (bs, g, ch)
    assert torch.allclose(got, want, rtol=1e-5, atol=2e-5 * scale), float((got - want).abs().max())
    upstream = torch.randn_like(want)
    (want * upstream).sum().backward()
    (got * upstream).sum().backward()

    def close(a, w, what):
        # two maxima over n points route gradients to single rows: a near-tie may pick another row than torch's max, which
        # moves a few entries; the gradient as a whole must agree
        rel = float((a - w).norm() / w.norm().clamp_min(1e-30))
        assert rel < 2e-3, (what, rel, float((a - w).abs().max()), float(w.abs().max()))

    close(xo.grad, xr.grad, "input")
    grads_r, grads_o = dict(ref.named_parameters()), dict(ours.named_parameters())
    for name in grads_r:
        if train and name in ("first_conv.0.bias", "first_conv.3.bias", "second_conv.0.bias"):
            # a bias in front of a training-mode BatchNorm has gradient zero (the batch mean removes it; first_conv.3's
            # bias shifts the global and the local half alike and reaches second_conv's BatchNorm as a constant): both
            # sides hold rounding noise only -- it must be negligible against the same layer's weight gradient
            wname = name.replace("bias", "weight")
            noise = 1e-4 * float(grads_r[wname].grad.abs().max())
            assert float(grads_o[name].grad.abs().max()) <= noise and float(grads_r[name].grad.abs().max()) <= noise, name
            continue
        close(grads_o[name].grad, grads_r[name].grad, name)
    if train:
        for (name, br), (_, bo) in zip(ref.named_buffers(), ours.named_buffers()):
            assert torch.allclose(bo.float(), br.float(), rtol=1e-4, atol=1e-6), name
