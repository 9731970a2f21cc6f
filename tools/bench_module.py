#!/usr/bin/env python
"""tools/bench_module.py — MSDeformAttn MODULE forward+backward (4 linears + sampling core), fused vs the reference's
op sequence, at the adapter's Injector / Extractor shapes. CUDA events, median. Prints JSON lines."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import vit_adapter_b200 as vab  # noqa: E402
from vit_adapter_b200.adapter import deform_inputs  # noqa: E402

CFG = {'S': (384, 6, 1.0), 'B': (768, 12, 0.5), 'L': (1024, 16, 0.5)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variant', default='B')
    ap.add_argument('--image', type=int, default=512)
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    d, heads, ratio = CFG[args.variant]
    di1, di2 = deform_inputs(torch.zeros(args.batch, 3, args.image, args.image, device=dev))
    h = args.image // 16
    nx, nc = h * h, (2 * h) ** 2 + h * h + (h // 2) ** 2
    for name, L, di, nq, nf in (('injector', 3, di1, nx, nc), ('extractor', 1, di2, nc, nx)):
        m = vab.MSDeformAttn(d, L, heads, 4, ratio).to(dev)
        with torch.no_grad():
            for p in m.parameters():
                p.add_(torch.randn_like(p) * 0.02)
        q = torch.randn(args.batch, nq, d, device=dev, requires_grad=True)
        f = torch.randn(args.batch, nf, d, device=dev, requires_grad=True)
        for amp in (False, True):
            vab.set_amp_value_dtype(torch.bfloat16 if amp else torch.float32)
            for fused, merge in ((False, False), (True, False), (True, True)):
                m.fused = fused
                m.merge_query_linears = merge

                def step():
                    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
                        out = m(q, di[0], f, di[1], di[2])
                    out.backward(torch.ones_like(out))
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                ts = []
                for _ in range(args.iters):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    step()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ts.sort()
                print(json.dumps({'variant': args.variant, 'call': name, 'batch': args.batch, 'amp_bf16': amp, 'fused': fused, 'merged_gemm': merge,
                                  'module_fwd_bwd_ms': ts[len(ts) // 2]}), flush=True)
    vab.set_amp_value_dtype(torch.float32)


if __name__ == '__main__':
    main()
