#!/usr/bin/env python
"""tools/profile_adapter_kernels.py — the smallest program that launches every adapter-side kernel (LayerNorm forward /
backward, DWConv forward / grad_x / grad_w, column sum, residual add) once per step at the ViT-Adapter-B 16 x 512^2 bf16
autocast shapes, for ncu. `--warm W` untimed steps first, then `--steps K`."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from vit_adapter_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--warm', type=int, default=2)
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--batch', type=int, default=16)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    B, n, C, H = args.batch, 5376, 768, 32
    x = torch.randn(B, n, C, device=dev)
    w, b = 1 + 0.1 * torch.randn(C, device=dev), 0.1 * torch.randn(C, device=dev)
    gy = torch.randn(B, n, C, device=dev).bfloat16()
    gres = torch.randn(B, n, C, device=dev)
    hid = C // 4
    hx = torch.randn(B, n, hid, device=dev).bfloat16()
    hg = torch.randn(B, n, hid, device=dev).bfloat16()
    dw, db = torch.randn(hid, 1, 3, 3, device=dev).bfloat16(), torch.randn(hid, device=dev).bfloat16()
    for _ in range(args.warm + args.steps):
        y, stats = _cabi.layernorm_forward(x, w, b, 1e-6, torch.bfloat16)
        _cabi.layernorm_backward(gy, x, w, stats, gres)
        _cabi.dwconv_forward(hx, dw, db, H, H)
        _cabi.dwconv_backward(hx, dw, hg, H, H)
        _cabi.colsum(gy)
        _cabi.residual_add(x, gy)
    torch.cuda.synchronize()
    print('launches', _cabi.launch_count())


if __name__ == '__main__':
    main()
