#!/usr/bin/env python
"""tools/bench_dwconv.py — adapter ConvFFN depth-wise conv: token-layout kernel vs the reference's op sequence
(slice/transpose/conv2d/transpose/cat through torch + cuDNN), forward and forward+backward, with HBM-roofline fractions
on the compulsory bytes (fwd: read x + write y; bwd: grad_x (read gy, write gx) + grad_w (read x, gy))."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from vit_adapter_b200.adapter import DWConv  # noqa: E402


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', default='B:192:32:16,B:192:32:2,S:96:32:16,L:256:56:1')
    args = ap.parse_args()
    peak = 6533.8
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peak = float(json.load(open(pk))['hbm_gbs'])
    for case in args.cases.split(','):
        name, C, H, B = case.split(':')
        C, H, B = int(C), int(H), int(B)
        W = H
        n = (H // 2) * (W // 2)
        for dtype in (torch.float32, torch.bfloat16):
            m = DWConv(C).cuda().to(dtype)
            x = torch.randn(B, 21 * n, C, device='cuda', dtype=dtype, requires_grad=True)
            gy = torch.randn(B, 21 * n, C, device='cuda', dtype=dtype)
            es = 4 if dtype == torch.float32 else 2
            nbytes = B * 21 * n * C * es
            for tk in (False, True):
                m.token_kernel = tk

                def fwd():
                    with torch.no_grad():
                        return m(x, H, W)

                def fwdbwd():
                    m.zero_grad(set_to_none=True)
                    x.grad = None
                    m(x, H, W).backward(gy)
                tf = timeit(fwd)
                tfb = timeit(fwdbwd)
                print(json.dumps({'case': name, 'C': C, 'H': H, 'batch': B, 'dtype': str(dtype).split('.')[-1],
                                  'impl': 'token_kernel' if tk else 'reference_sequence(torch/cuDNN)',
                                  'fwd_us': tf * 1e3, 'fwd_bwd_us': tfb * 1e3,
                                  'fwd_hbm_frac': 2 * nbytes / (tf * 1e-3) / 1e9 / peak,
                                  'fwd_bwd_hbm_frac': 6 * nbytes / (tfb * 1e-3) / 1e9 / peak}), flush=True)


if __name__ == '__main__':
    main()
