#!/usr/bin/env python
"""tools/profile_ddp_eager.py — what limits the EAGER data-parallel training step (2-GPU eager bf16 scaled 0.75,
the graphed step 0.96). One torch.profiler capture of a few eager steps on rank 0:

  * wall time per step vs the sum of GPU kernel time per step  -> how much of the step the GPU idles (host-bound share),
  * host calls that BLOCK on the device per step (cudaStreamSynchronize / cudaDeviceSynchronize / cudaMemcpy D2H from
    aten::nonzero, aten::item, aten::_local_scalar_dense) and which module issues them,
  * collectives per step (SyncBatchNorm's all_gather in forward / all_reduce in backward, DDP's bucket all_reduce).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/profile_ddp_eager.py --amp
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variant', default='B')
    ap.add_argument('--image', type=int, default=512)
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--amp', action='store_true')
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--bn', default='sync', choices=['sync', 'local'], help="'local' keeps plain BatchNorm under DDP: isolates SyncBN's cost")
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    import vit_adapter_b200 as vab
    from bench_step import Net
    if args.amp:
        vab.set_amp_value_dtype(torch.bfloat16)
    torch.manual_seed(1 + rank)
    net = Net(args.variant, sync_bn=(world > 1 and args.bn == 'sync'), with_cp=False).to(dev)
    model = nn.parallel.DistributedDataParallel(net, device_ids=[local], gradient_as_bucket_view=True) if world > 1 else net
    opt = torch.optim.AdamW(net.parameters(), lr=6e-5, fused=True)
    img = torch.randn(args.batch, 3, args.image, args.image, device=dev)
    lab = torch.randint(0, 150, (args.batch, args.image // 4, args.image // 4), device=dev)

    def step():
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=args.amp):
            loss = F.cross_entropy(model(img).float(), lab)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(12):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    wall_ms = e0.elapsed_time(e1) / args.steps
    if rank == 0:
        ka = prof.key_averages()
        n = args.steps
        gpu_ms = sum(getattr(k, 'self_device_time_total', 0) for k in ka) / 1e3 / n
        def pick(sub):
            return {k.key: {'calls_per_step': k.count / n, 'cpu_ms_per_step': k.cpu_time_total / 1e3 / n,
                            'gpu_ms_per_step': getattr(k, 'device_time_total', 0) / 1e3 / n}
                    for k in ka if any(s in k.key for s in sub)}
        blocking = pick(['cudaStreamSynchronize', 'cudaDeviceSynchronize', 'cudaEventSynchronize', 'aten::nonzero', 'aten::item',
                         'aten::_local_scalar_dense', 'cudaMemcpyAsync'])
        coll = pick(['c10d::', 'nccl:', 'ncclDevKernel', 'SyncBatchNorm', 'batch_norm_gather_stats', 'batch_norm_backward_reduce'])
        launches = sum(k.count for k in ka if k.key in ('cudaLaunchKernel', 'cudaLaunchKernelExC', 'cuLaunchKernel', 'cuLaunchKernelEx')) / n
        print(json.dumps({'n_gpus': world, 'variant': args.variant, 'amp': args.amp, 'batchnorm': args.bn if world > 1 else 'local',
                          'wall_ms_per_step': wall_ms, 'gpu_kernel_ms_per_step': gpu_ms, 'gpu_idle_share': max(0.0, 1 - gpu_ms / wall_ms),
                          'kernel_launches_per_step': launches, 'blocking_host_calls': blocking, 'collectives': coll}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
