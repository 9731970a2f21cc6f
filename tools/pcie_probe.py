#!/usr/bin/env python
"""tools/pcie_probe.py — what the host link gives: pinned H2D alone, D2H alone, both at once, split over 1 / 2 / 4 streams.
Explains bench.py's e2e number (its timed region is bound by these copies, not by the kernels)."""
import json
import time

import torch


def run(nbytes, n_streams, do_h2d, do_d2h, reps=10):
    per = nbytes // n_streams
    hs = [torch.empty(per, dtype=torch.uint8).pin_memory() for _ in range(n_streams)]
    hd = [torch.empty(per, dtype=torch.uint8).pin_memory() for _ in range(n_streams)]
    ds = [torch.empty(per, dtype=torch.uint8, device='cuda') for _ in range(n_streams)]
    dd = [torch.empty(per, dtype=torch.uint8, device='cuda') for _ in range(n_streams)]
    s_in = [torch.cuda.Stream() for _ in range(n_streams)]
    s_out = [torch.cuda.Stream() for _ in range(n_streams)]

    def once():
        for i in range(n_streams):
            if do_h2d:
                with torch.cuda.stream(s_in[i]):
                    ds[i].copy_(hs[i], non_blocking=True)
            if do_d2h:
                with torch.cuda.stream(s_out[i]):
                    hd[i].copy_(dd[i], non_blocking=True)
    once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return nbytes / dt / 1e9


def bench_like(dep):
    """bench.py's per-step traffic: 8 tensors in, 8 out, of the adapter's sizes; `dep`: copy-out of step k waits for copy-in of step k
    (as the compute in between does)."""
    mb = [132, 19, 9, 25, 25, 33, 17, 132]
    hin = [torch.empty(m << 20, dtype=torch.uint8).pin_memory() for m in mb]
    hout = [torch.empty(m << 20, dtype=torch.uint8).pin_memory() for m in mb]
    din = [[torch.empty(m << 20, dtype=torch.uint8, device='cuda') for m in mb] for _ in range(2)]
    dout = [[torch.empty(m << 20, dtype=torch.uint8, device='cuda') for m in mb] for _ in range(2)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev = [torch.cuda.Event() for _ in range(2)]

    def steps(n):
        for k in range(n):
            slot = k & 1
            with torch.cuda.stream(s_in):
                for d, h in zip(din[slot], hin):
                    d.copy_(h, non_blocking=True)
                ev[slot].record(s_in)
            with torch.cuda.stream(s_out):
                if dep:
                    s_out.wait_event(ev[slot])
                for h, d in zip(hout, dout[slot]):
                    h.copy_(d, non_blocking=True)
    steps(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps(20)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    return sum(mb) * (1 << 20) / dt / 1e9, dt * 1e3


def main():
    """Alone:   python tools/pcie_probe.py
    All GPUs of the box at once (what bench.py --gpus N does to the host):
             python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py
    Every rank drives its own GPU; the measurements start together (barrier) and rank 0 prints, per case, the slowest and
    the mean per-GPU rate and the aggregate over the box."""
    import os
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('gloo')

    def together(fn):
        if world > 1:
            dist.barrier()
        v = fn()
        if world == 1:
            return {'per_gpu_min': v, 'per_gpu_mean': v, 'aggregate': v}
        got = [None] * world
        dist.all_gather_object(got, v)
        return {'per_gpu_min': min(got), 'per_gpu_mean': sum(got) / world, 'aggregate': sum(got)}

    nbytes = 384 << 20
    for ns in (1, 2):
        row = {'gpus_at_once': world, 'streams_per_direction': ns, 'MB_per_direction': nbytes >> 20,
               'h2d_only_GBps': together(lambda: run(nbytes, ns, True, False)),
               'd2h_only_GBps': together(lambda: run(nbytes, ns, False, True)),
               'both_GBps_each_direction': together(lambda: run(nbytes, ns, True, True))}
        if rank == 0:
            print(json.dumps(row), flush=True)
    for dep in (False, True):
        r = together(lambda: bench_like(dep)[0])
        if rank == 0:
            print(json.dumps({'gpus_at_once': world, 'bench_like_traffic': 'copy-out waits for copy-in' if dep else 'independent directions',
                              'GBps_each_direction': r}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
