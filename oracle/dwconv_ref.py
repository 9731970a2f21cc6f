"""Restatement of the reference's DWConv (ConvFFN's depth-wise 3x3 on the packed token sequence). TEST INFRASTRUCTURE ONLY.

Follows DWConv.forward, detection/mmdet_custom/models/backbones/adapter_modules.py:73-87: the [B, 21n, C] sequence is
three maps (16n tokens at 2H x 2W, 4n at H x W, n at H/2 x W/2); each is convolved with the SAME depth-wise 3x3
(stride 1, zero padding 1, groups = C) and the results are concatenated in the original token order.
Pinned against the real reference class by tests/golden/dwconv_tokens.npz (tests/golden/make_golden.py).
"""
import torch
import torch.nn.functional as F


def dwconv_tokens(x, weight, bias, H, W):
    B, N, C = x.shape
    n = N // 21
    assert N == 21 * n and 4 * n == H * W, (N, H, W)
    pieces = []
    start = 0
    for count, h, w in ((16 * n, 2 * H, 2 * W), (4 * n, H, W), (n, H // 2, W // 2)):
        fmap = x[:, start:start + count].permute(0, 2, 1).reshape(B, C, h, w)
        start += count
        out = F.conv2d(fmap, weight, bias, stride=1, padding=1, groups=C)
        pieces.append(out.reshape(B, C, h * w).permute(0, 2, 1))
    return torch.cat(pieces, 1)


def dwconv_tokens_backward(x, weight, bias, H, W, grad_y):
    x = x.detach().clone().requires_grad_(True)
    w = weight.detach().clone().requires_grad_(True)
    b = bias.detach().clone().requires_grad_(True)
    dwconv_tokens(x, w, b, H, W).backward(grad_y)
    return x.grad, w.grad, b.grad
