#!/usr/bin/env python
"""tools/chunk_sweep.py — queries per CTA chunk of the query-order kernels (tuning keys fwd_chunk / bwd_chunk; 0 = the
library's own choice, pick_chunk in msda_abi.cu) against time, at the BASELINE shapes: L2-flushed CUDA-event medians.
    python tools/chunk_sweep.py --variants B,L --dtypes f32,bf16 --chunks 0,32,64,96,128,160,256"""
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
    ap.add_argument('--variants', default='B,L')
    ap.add_argument('--dtypes', default='f32,bf16')
    ap.add_argument('--chunks', default='0,32,64,96,128,160,192,256')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'chunk_sweep.jsonl'))
    args = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    out = open(args.out, 'w')
    for variant in args.variants.split(','):
        batch = args.batch or VARIANTS[variant][3]
        for (name, N, M, D, Lq, shapes) in call_shapes(variant, batch):
            for dn in args.dtypes.split(','):
                dtype = {'f32': torch.float32, 'bf16': torch.bfloat16}[dn]
                g = {k: v.to(DEV) for k, v in adapter_inputs(name, N, M, D, Lq, shapes, 0, dtype).items()}
                fwd = lambda: _cabi.forward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], 64)
                bwd = lambda: _cabi.backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'], 64)
                row = dict(variant=variant, batch=batch, call=name, dtype=dn, fwd_us={}, bwd_us={})
                for ch in args.chunks.split(','):
                    _cabi.set_tuning(fwd_chunk=int(ch), bwd_chunk=int(ch), bwd_sorted=1)   # query-order kernels only
                    row['fwd_us'][ch] = round(timeit(fwd, args.iters, 3, flush)['med'] * 1e3, 1)
                    row['bwd_us'][ch] = round(timeit(bwd, args.iters, 3, flush)['med'] * 1e3, 1)
                _cabi.set_tuning(fwd_chunk=0, bwd_chunk=0, bwd_sorted=0)
                print(json.dumps(row), flush=True)
                out.write(json.dumps(row) + '\n')
    out.close()


if __name__ == '__main__':
    main()
