#!/usr/bin/env python
"""tools/bwd_cell_check.py — the cell-bucketed backward against the query-order backward at the BASELINE shapes:
max relative difference of the three gradients and L2-flushed CUDA-event timings (median), per chunk setting.
    python tools/bwd_cell_check.py --variants B,L --dtypes f32,bf16 --chunks 0,64,128"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import VARIANTS, adapter_inputs, algorithmic_bytes, call_shapes, n_points  # noqa: E402
from tools.sweep import timeit  # noqa: E402
from vit_adapter_b200 import _cabi  # noqa: E402

DEV = torch.device('cuda', 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variants', default='B,S,L,L64')
    ap.add_argument('--dtypes', default='f32,bf16')
    ap.add_argument('--chunks', default='0')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--batch', type=int, default=0, help='images per call (default: the variant\'s BASELINE batch)')
    ap.add_argument('--dist', default='adapter')
    ap.add_argument('--mode', default='cell', choices=['cell', 'packed16', 'sorted'], help='which opt-in backward to compare with the default')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'bwd_cell_check.jsonl'))
    args = ap.parse_args()
    peak = 6533.8
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    out = open(args.out, 'w')
    for variant in args.variants.split(','):
        batch = args.batch or VARIANTS[variant][3]
        for (name, N, M, D, Lq, shapes) in call_shapes(variant, batch):
            for dn in args.dtypes.split(','):
                dtype = {'f32': torch.float32, 'bf16': torch.bfloat16, 'f16': torch.float16}[dn]
                inp = adapter_inputs(name, N, M, D, Lq, shapes, 0, dtype)
                if args.dist == 'uniform':
                    g = torch.Generator().manual_seed(1)
                    inp['loc'] = torch.rand(inp['loc'].shape, generator=g)
                g = {k: v.to(DEV) for k, v in inp.items()}
                call = lambda: _cabi.backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'], 64)
                _cabi.set_tuning(bwd_cell=1, bwd_sorted=1)   # baseline: the query-order kernel
                ref = call()
                t_old = timeit(call, args.iters, 3, flush)
                torch.cuda.synchronize()
                nbytes = algorithmic_bytes(N, M, D, Lq, shapes, inp['value'].element_size())['bwd']
                for ch in args.chunks.split(','):
                    if args.mode == 'packed16':
                        if dn == 'f32':
                            continue
                        _cabi.set_tuning(bwd_cell=0, bwd_packed16=2)
                    elif args.mode == 'sorted':
                        _cabi.set_tuning(bwd_cell=0, bwd_sorted=2)
                    else:
                        _cabi.set_tuning(bwd_cell=2, bwd_cell_chunk=int(ch))
                    got = call()
                    torch.cuda.synchronize()
                    errs = [float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-30)) for a, b in zip(got, ref)]
                    t_new = timeit(call, args.iters, 3, flush)
                    row = dict(mode=args.mode, variant=variant, batch=batch, call=name, dtype=dn, chunk=int(ch), dist=args.dist, old_us=round(t_old['med'] * 1e3, 1),
                               cell_us=round(t_new['med'] * 1e3, 1), cell_min_us=round(t_new['min'] * 1e3, 1),
                               speedup=round(t_old['med'] / t_new['med'], 2), hbm_frac=round(nbytes / (t_new['med'] * 1e-3) / 1e9 / peak, 3),
                               gsamples=round(n_points(N, M, Lq, len(shapes)) / (t_new['med'] * 1e-3) / 1e9, 2),
                               rel_err_gv_gl_ga=['%.2e' % e for e in errs])
                    print(json.dumps(row), flush=True)
                    out.write(json.dumps(row) + '\n')
                _cabi.set_tuning(bwd_cell=0, bwd_cell_chunk=0, bwd_packed16=0, bwd_sorted=0)
    out.close()


if __name__ == '__main__':
    main()
