#!/usr/bin/env python
"""tools/profile_block.py — where an adapter interaction (Injector + Extractor incl. ConvFFN, no ViT blocks) spends its
GPU time: torch.profiler kernel table, forward+backward, at the ViT-Adapter-B shapes. Informational (decides what is
worth fusing next); prints the top kernels and writes the full table as JSON."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import vit_adapter_b200 as vab  # noqa: E402
from vit_adapter_b200.adapter import InteractionBlock, deform_inputs  # noqa: E402

CFG = {'S': (384, 6, 1.0), 'B': (768, 12, 0.5), 'L': (1024, 16, 0.5)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variant', default='B')
    ap.add_argument('--image', type=int, default=512)
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--amp', type=int, default=1)
    ap.add_argument('--out', default='')
    ap.add_argument('--top', type=int, default=25)
    ap.add_argument('--reference-sequence', action='store_true', help="every adapter-side kernel off: torch LayerNorm, DWConv op sequence, separate linears + softmax, nn.Linear backward, torch adds (the sampling core stays this repo's)")
    ap.add_argument('--no-fold', action='store_true', help='autograd adds the residual gradient (no folding into the LayerNorm backward)')
    ap.add_argument('--no-residual-kernel', action='store_true', help="torch's mixed-dtype add for the residual epilogues")
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    from vit_adapter_b200.adapter import adapter_modules as am
    if args.no_fold:
        am.apply_norm_residual = lambda norm, x, fused=True: (am.apply_norm(norm, x, fused), x)
    if args.no_residual_kernel:
        am.residual_add = lambda res, branch, fused=True: res + branch
    d, heads, ratio = CFG[args.variant]
    blk = InteractionBlock(d, heads, 4, deform_ratio=ratio, cffn_ratio=0.25, init_values=0.0, extra_extractor=False).to(dev)
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(torch.randn_like(p) * 0.02)
    if args.reference_sequence:
        for m in blk.modules():
            for flag in ('fused_norm', 'token_kernel', 'fused', 'merge_query_linears', 'colsum_bias_grad'):
                if hasattr(m, flag):
                    setattr(m, flag, False)
    if args.amp:
        vab.set_amp_value_dtype(torch.bfloat16)
    img = torch.zeros(args.batch, 3, args.image, args.image, device=dev)
    di1, di2 = deform_inputs(img)
    h = args.image // 16
    x = torch.randn(args.batch, h * h, d, device=dev, requires_grad=True)
    c = torch.randn(args.batch, 21 * (h // 2) ** 2, d, device=dev, requires_grad=True)

    def step():
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=bool(args.amp)):
            xo, co = blk(x, c, [], di1, di2, h, h)
        (xo.float().square().sum() + co.float().square().sum()).backward()   # a loss whose gradient is a real tensor, not an expanded scalar
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / 10
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            step()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages():
        t = getattr(ev, 'device_time_total', None)
        if t is None:
            t = getattr(ev, 'cuda_time_total', 0)
        if t > 0:
            rows.append({'name': ev.key[:110], 'us_per_step': t / 5, 'calls_per_step': ev.count / 5})
    rows.sort(key=lambda r: -r['us_per_step'])
    tot = sum(r['us_per_step'] for r in rows)
    print(json.dumps({'variant': args.variant, 'batch': args.batch, 'amp': bool(args.amp), 'step_ms_events': wall,
                      'kernel_us_total': tot, 'kernels_per_step': sum(r['calls_per_step'] for r in rows)}))
    for r in rows[:args.top]:
        print('%8.1f us %5.1f%% x%-4g %s' % (r['us_per_step'], 100 * r['us_per_step'] / tot, r['calls_per_step'], r['name']))
    if args.out:
        json.dump(rows, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
