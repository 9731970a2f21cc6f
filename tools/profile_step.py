#!/usr/bin/env python
"""tools/profile_step.py — the smallest program that launches the four hot kernels of one bench step
(Injector fwd, bwd; Extractor fwd, bwd), for ncu. `--warm W` untimed steps first, then `--steps K`."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import VARIANTS, adapter_inputs, call_shapes  # noqa: E402
from vit_adapter_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variant', default='B')
    ap.add_argument('--dtype', default='f32')
    ap.add_argument('--warm', type=int, default=2)
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--ref', action='store_true', help="launch the reference's CUDA kernels instead")
    ap.add_argument('--smem', type=int, default=0, help='fwd_smem tuning mode (0 auto, 1 off, 2 forced)')
    ap.add_argument('--smem-threads', type=int, default=0)
    ap.add_argument('--fwd-only', action='store_true')
    ap.add_argument('--bwd-sorted', type=int, default=0, help='bwd_sorted tuning mode (0 auto, 1 off, 2 forced)')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    batch = args.batch or VARIANTS[args.variant][3]
    dtype = torch.float32 if args.dtype == 'f32' else torch.bfloat16
    calls = []
    for i, (name, N, M, D, Lq, shapes) in enumerate(call_shapes(args.variant, batch)):
        h = adapter_inputs(name, N, M, D, Lq, shapes, seed=i, dtype=dtype)
        calls.append({k: v.to(dev) for k, v in h.items()})
    if args.ref:
        from oracle import refcuda
    _cabi.set_tuning(fwd_smem=args.smem, fwd_smem_threads=args.smem_threads, bwd_sorted=args.bwd_sorted)
    for _ in range(args.warm + args.steps):
        for d in calls:
            if args.ref:
                refcuda.forward(d['value'], d['shapes'], d['lsi'], d['loc'], d['aw'])
                refcuda.backward(d['value'], d['shapes'], d['lsi'], d['loc'], d['aw'], d['grad_out'])
            else:
                _cabi.forward(d['value'], d['shapes'], d['lsi'], d['loc'], d['aw'], 64)
                if not args.fwd_only:
                    _cabi.backward(d['value'], d['shapes'], d['lsi'], d['loc'], d['aw'], d['grad_out'], 64)
    torch.cuda.synchronize()
    print('ok')


if __name__ == '__main__':
    main()
