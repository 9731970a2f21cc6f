#!/usr/bin/env python
"""tools/sweep.py — per-call timings over the SURVEY App. B shapes: ours (fp32, bf16) vs the reference's
CUDA kernels (oracle/_ref, fp32), forward and backward separately, CUDA events, L2 flushed between
iterations (256 MB scratch write) and warm. Writes gpurun_out/sweep.json and a markdown table."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import VARIANTS, adapter_inputs, algorithmic_bytes, call_shapes, n_points  # noqa: E402
from vit_adapter_b200 import _cabi  # noqa: E402
from oracle import refcuda  # noqa: E402

DEV = torch.device('cuda', 0)


def timeit(fn, iters, warm, flush):
    """flushed: a 256 MB write precedes every timed call (it also keeps the GPU busy while the call is enqueued, so
    no launch latency leaks into the events). warm: 10 back-to-back calls per event pair (the first call of an idle GPU
    would otherwise add the host's ~15 us enqueue latency), reported per call."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    reps = 1 if flush is not None else 10
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        else:
            fn()  # keeps the queue non-empty when the first event is recorded
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    ts.sort()
    return {'min': ts[0], 'med': ts[len(ts) // 2]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variants', default='B,S,L,L64,HTC')
    ap.add_argument('--iters', type=int, default=30)
    ap.add_argument('--qc', default='0', help='comma list of query-chunk overrides to try (0 = heuristic)')
    ap.add_argument('--minb', default='0:0', help='comma list of fwd:bwd min-CTAs-per-SM kernel variants (0 = default)')
    ap.add_argument('--smem', default='0', help='comma list of fwd_smem[:threads[:chunks]] settings (0 auto, 1 off, 2 forced)')
    ap.add_argument('--wide', default='0', help='comma list of fwd_wide settings (0 auto, 1 off, 2 on)')
    ap.add_argument('--regimes', default='flushed,warm')
    ap.add_argument('--dtypes', default='f32,bf16')
    ap.add_argument('--no-ref', action='store_true')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'sweep.json'))
    ap.add_argument('--dist', default='adapter')
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    peak = 6533.8
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peak = float(json.load(open(pk))['hbm_gbs'])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    rows = []
    for variant in args.variants.split(','):
        M, D, side, batch = VARIANTS[variant]
        for ci, (name, N, Mh, Dh, Lq, shapes) in enumerate(call_shapes(variant, batch)):
            host = adapter_inputs(name, N, Mh, Dh, Lq, shapes, seed=ci, dtype=torch.float32)
            if args.dist == 'uniform':
                g = torch.Generator().manual_seed(5)
                host['loc'] = torch.rand(host['loc'].shape, generator=g)
            d32 = {k: v.to(DEV) for k, v in host.items()}
            d16 = dict(d32, value=d32['value'].bfloat16(), grad_out=d32['grad_out'].bfloat16())
            pts = n_points(N, Mh, Lq, len(shapes))
            for dtype, d, es in (('f32', d32, 4), ('bf16', d16, 2)):
                if dtype not in args.dtypes.split(','):
                    continue
                ab = algorithmic_bytes(N, Mh, Dh, Lq, shapes, es)
                for qc, mb, sm, wd in [(int(x), y, z, int(w)) for x in args.qc.split(',') for y in args.minb.split(',') for z in args.smem.split(',') for w in args.wide.split(',')]:
                    fmb, bmb = [int(v) for v in mb.split(':')]
                    smv = [int(v) for v in sm.split(':')] + [0, 0]
                    _cabi.set_tuning(fwd_chunk=qc, bwd_chunk=qc, fwd_min_ctas=fmb, bwd_min_ctas=bmb, fwd_smem=smv[0],
                                     fwd_smem_threads=smv[1], fwd_smem_chunks=smv[2], fwd_wide=wd)
                    mb = mb + '/s' + sm + '/w%d' % wd
                    fw = lambda: _cabi.forward(d['value'], d['shapes'], d['lsi'], d['loc'], d['aw'], 64)
                    bw = lambda: _cabi.backward(d['value'], d['shapes'], d['lsi'], d['loc'], d['aw'], d['grad_out'], 64)
                    for regime, fl in (('flushed', flush), ('warm', None)):
                        if regime not in args.regimes.split(','):
                            continue
                        tf = timeit(fw, args.iters, 5, fl)
                        tb = timeit(bw, args.iters, 5, fl)
                        rows.append({'variant': variant, 'call': name, 'dtype': dtype, 'impl': 'ours', 'qc': '%d/%s' % (qc, mb),
                                     'regime': regime, 'fwd_us': tf['med'] * 1e3, 'bwd_us': tb['med'] * 1e3,
                                     'fwd_min_us': tf['min'] * 1e3, 'bwd_min_us': tb['min'] * 1e3, 'pts': pts,
                                     'gsamples_s': pts / ((tf['med'] + tb['med']) * 1e-3) / 1e9,
                                     'fwd_frac': ab['fwd'] / (tf['med'] * 1e-3) / 1e9 / peak,
                                     'bwd_frac': ab['bwd'] / (tb['med'] * 1e-3) / 1e9 / peak,
                                     'frac': (ab['fwd'] + ab['bwd']) / ((tf['med'] + tb['med']) * 1e-3) / 1e9 / peak})
                        print(json.dumps(rows[-1]), flush=True)
                _cabi.set_tuning(fwd_chunk=0, bwd_chunk=0, fwd_min_ctas=0, bwd_min_ctas=0, fwd_smem=0, fwd_smem_threads=0, fwd_smem_chunks=0, fwd_wide=0)
            if refcuda.available() and not args.no_ref:
                ab = algorithmic_bytes(N, Mh, Dh, Lq, shapes, 4)
                fw = lambda: refcuda.forward(d32['value'], d32['shapes'], d32['lsi'], d32['loc'], d32['aw'])
                bw = lambda: refcuda.backward(d32['value'], d32['shapes'], d32['lsi'], d32['loc'], d32['aw'], d32['grad_out'])
                for regime, fl in (('flushed', flush), ('warm', None)):
                    if regime not in args.regimes.split(','):
                        continue
                    tf = timeit(fw, max(5, args.iters // 3), 2, fl)
                    tb = timeit(bw, max(5, args.iters // 3), 2, fl)
                    rows.append({'variant': variant, 'call': name, 'dtype': 'f32', 'impl': 'ref_cuda', 'qc': None,
                                 'regime': regime, 'fwd_us': tf['med'] * 1e3, 'bwd_us': tb['med'] * 1e3,
                                 'fwd_min_us': tf['min'] * 1e3, 'bwd_min_us': tb['min'] * 1e3, 'pts': pts,
                                 'gsamples_s': pts / ((tf['med'] + tb['med']) * 1e-3) / 1e9,
                                 'fwd_frac': ab['fwd'] / (tf['med'] * 1e-3) / 1e9 / peak,
                                 'bwd_frac': ab['bwd'] / (tb['med'] * 1e-3) / 1e9 / peak,
                                 'frac': (ab['fwd'] + ab['bwd']) / ((tf['med'] + tb['med']) * 1e-3) / 1e9 / peak})
                    print(json.dumps(rows[-1]), flush=True)
            del d32, d16
            torch.cuda.empty_cache()
    json.dump({'peak_gbs': peak, 'gpu': torch.cuda.get_device_name(0), 'rows': rows}, open(args.out, 'w'), indent=1)
    md = ['| variant | call | impl | dtype | qc | regime | fwd µs | bwd µs | Gsamples/s | HBM-roofline frac (fwd / bwd / f+b) |',
          '|---|---|---|---|---|---|--:|--:|--:|---|']
    for r in rows:
        md.append('| %s | %s | %s | %s | %s | %s | %.1f | %.1f | %.2f | %.3f / %.3f / %.3f |' % (
            r['variant'], r['call'], r['impl'], r['dtype'], r['qc'], r['regime'], r['fwd_us'], r['bwd_us'],
            r['gsamples_s'], r['fwd_frac'], r['bwd_frac'], r['frac']))
    open(args.out.replace('.json', '.md'), 'w').write('\n'.join(md) + '\n')


if __name__ == '__main__':
    main()
