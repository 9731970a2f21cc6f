#!/usr/bin/env python
"""tools/bench_layernorm.py — adapter LayerNorm row kernels alone (through the binding: kernel + output allocation), and
torch's LayerNorm forward (+ the cast the consuming Linear adds under autocast) for comparison, at the adapter's shapes.
CUDA events around back-to-back calls rotating through > 512 MB of distinct inputs (no L2 reuse between launches).
HBM fraction on the compulsory bytes: fwd rows*C*(e_in+e_out), bwd rows*C*(2*e_in+e_out)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from vit_adapter_b200 import _cabi  # noqa: E402


def timed(fn, nset, reps):
    for i in range(nset):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i % nset)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    peak = 6533.8
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peak = float(json.load(open(pk))['hbm_gbs'])
    for name, rows, C in (('B c-tokens bs16', 16 * 5376, 768), ('B x-tokens bs16', 16 * 1024, 768), ('S c-tokens bs16', 16 * 5376, 384),
                          ('L c-tokens bs1', 16464, 1024), ('B c-tokens bs2', 2 * 5376, 768)):
        w = 1 + 0.1 * torch.randn(C, device='cuda')
        b = 0.1 * torch.randn(C, device='cuda')
        for out_dtype in (torch.float32, torch.bfloat16):
            eo = 2 if out_dtype == torch.bfloat16 else 4
            nbytes_f = rows * C * (4 + eo)
            nbytes_b = rows * C * (8 + eo)
            nset = min(32, max(2, (512 << 20) // nbytes_f + 1))
            xs = [torch.randn(rows, C, device='cuda') for _ in range(nset)]
            gys = [torch.randn(rows, C, device='cuda').to(out_dtype) for _ in range(min(nset, 4))]
            stats = [_cabi.layernorm_forward(x, w, b, 1e-6, out_dtype)[1] for x in xs]
            tf = timed(lambda i: _cabi.layernorm_forward(xs[i], w, b, 1e-6, out_dtype), nset, 3 * nset)
            tb = timed(lambda i: _cabi.layernorm_backward(gys[i % len(gys)], xs[i], w, stats[i]), nset, 3 * nset)

            def torch_fwd(i):
                y = F.layer_norm(xs[i], (C,), w, b, 1e-6)
                return y.to(out_dtype) if out_dtype != torch.float32 else y
            tt = timed(torch_fwd, nset, 3 * nset)
            print(json.dumps({'shape': name, 'rows': rows, 'C': C, 'x': 'f32', 'y': str(out_dtype).split('.')[-1], 'fwd_us': tf,
                              'bwd_us': tb, 'fwd_GBps': nbytes_f / tf / 1e3, 'bwd_GBps': nbytes_b / tb / 1e3,
                              'fwd_hbm_frac': nbytes_f / tf / 1e3 / peak, 'bwd_hbm_frac': nbytes_b / tb / 1e3 / peak,
                              'torch_fwd_us': tt, 'buffers_rotated': nset}), flush=True)


if __name__ == '__main__':
    main()
