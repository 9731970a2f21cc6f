#!/usr/bin/env python
"""tools/minctas_sweep.py — resident CTAs per SM the query-order kernels are compiled for (tuning keys fwd_min_ctas: 3 / 4 / 6,
bwd_min_ctas: 2 / 3 / 4; 0 = the default, 4 / 3) against time at the BASELINE shapes: L2-flushed CUDA-event medians."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import VARIANTS, adapter_inputs, call_shapes  # noqa: E402
from tools.sweep import timeit  # noqa: E402
from vit_adapter_b200 import _cabi  # noqa: E402

DEV = torch.device('cuda', 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variants', default='B,S,L')
    ap.add_argument('--dtypes', default='f32,bf16')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'minctas_sweep.jsonl'))
    args = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    out = open(args.out, 'w')
    for variant in args.variants.split(','):
        for (name, N, M, D, Lq, shapes) in call_shapes(variant, VARIANTS[variant][3]):
            for dn in args.dtypes.split(','):
                dtype = {'f32': torch.float32, 'bf16': torch.bfloat16}[dn]
                g = {k: v.to(DEV) for k, v in adapter_inputs(name, N, M, D, Lq, shapes, 0, dtype).items()}
                fwd = lambda: _cabi.forward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], 64)
                bwd = lambda: _cabi.backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'], 64)
                row = dict(variant=variant, call=name, dtype=dn, fwd_us={}, bwd_us={})
                _cabi.set_tuning(bwd_sorted=1)
                for m in (0, 3, 6):
                    _cabi.set_tuning(fwd_min_ctas=m)
                    row['fwd_us'][str(m)] = round(timeit(fwd, args.iters, 3, flush)['med'] * 1e3, 1)
                for m in (0, 2, 4):
                    _cabi.set_tuning(bwd_min_ctas=m)
                    row['bwd_us'][str(m)] = round(timeit(bwd, args.iters, 3, flush)['med'] * 1e3, 1)
                _cabi.set_tuning(fwd_min_ctas=0, bwd_min_ctas=0, bwd_sorted=0)
                print(json.dumps(row), flush=True)
                out.write(json.dumps(row) + '\n')
    out.close()


if __name__ == '__main__':
    main()
