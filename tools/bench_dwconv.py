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
    ap.add_argument('--kernels', action='store_true', help='time the three kernels alone through the C ABI')
    args = ap.parse_args()
    peak = 6533.8
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peak = float(json.load(open(pk))['hbm_gbs'])
    if args.kernels:
        return kernels(args.cases, peak)
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


def kernels(cases, peak):
    """Per-kernel device times through the C ABI (no module / autograd overhead): launches rotate through distinct buffer sets (> 512 MB in total) so that no launch re-reads what the previous wrote."""
    from vit_adapter_b200 import _cabi
    lib = _cabi.load()
    st = torch.cuda.current_stream().cuda_stream
    for case in cases.split(','):
        name, C, H, B = case.split(':')
        C, H, B = int(C), int(H), int(B)
        W, n = H, (H // 2) * (H // 2)
        for dtype in (torch.float32, torch.bfloat16):
            code = _cabi._DTYPES[dtype]
            es = 4 if dtype == torch.float32 else 2
            nbytes = B * 21 * n * C * es
            nset = max(2, int((512 << 20) // (2 * nbytes)) + 1)   # rotate through > 512 MB (L2 is 126 MB)
            nset = min(nset, 64)
            xs = [torch.randn(B, 21 * n, C, device='cuda', dtype=dtype) for _ in range(nset)]
            ys = [torch.empty_like(xs[0]) for _ in range(nset)]
            w = torch.randn(C, 1, 3, 3, device='cuda', dtype=dtype)
            bias = torch.randn(C, device='cuda', dtype=dtype)
            gw = torch.empty(C * 9, device='cuda', dtype=torch.float32)
            gb = torch.empty(C, device='cuda', dtype=torch.float32)
            wsb = lib.adapter_dwconv_backward_weight_workspace_bytes(code, B, 21 * n, C, H, W)
            ws = torch.empty(max(wsb, 1), device='cuda', dtype=torch.uint8)
            fns = {
                'forward': (lambda i: lib.adapter_dwconv_forward(code, xs[i].data_ptr(), w.data_ptr(), bias.data_ptr(), ys[i].data_ptr(),
                                                                 B, 21 * n, C, H, W, st), 2 * nbytes),
                'backward_input': (lambda i: lib.adapter_dwconv_backward_input(code, xs[i].data_ptr(), w.data_ptr(), ys[i].data_ptr(),
                                                                               B, 21 * n, C, H, W, st), 2 * nbytes),
                'backward_weight': (lambda i: lib.adapter_dwconv_backward_weight(code, xs[i].data_ptr(), ys[i].data_ptr(), gw.data_ptr(),
                                                                                 gb.data_ptr(), B, 21 * n, C, H, W, ws.data_ptr(), wsb, st), 2 * nbytes),
            }
            for kname, (fn, alg) in fns.items():
                reps = 3 * nset
                for i in range(nset):
                    assert fn(i) == 0
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(reps):
                    fn(i % nset)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / reps
                print(json.dumps({'case': name, 'C': C, 'H': H, 'batch': B, 'dtype': str(dtype).split('.')[-1], 'kernel': kname,
                                  'us': us, 'algorithmic_MB': alg / 1e6, 'GBps': alg / us / 1e3, 'hbm_frac': alg / us / 1e3 / peak,
                                  'buffers_rotated': nset}), flush=True)


if __name__ == '__main__':
    main()
