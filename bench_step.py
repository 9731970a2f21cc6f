#!/usr/bin/env python
"""bench_step.py — model-level img/s around the MSDeformAttn hot path (BASELINE.json configs[2], [4]).

    python bench_step.py --variant B --mode train --batch 2            # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench_step.py --gpus 8 ...
    python bench_step.py --variant L --mode infer --image 1024 --batch 1 --amp   # HTC++-shape backbone inference

What runs. The adapter side is THIS repo's drop-in code (vit_adapter_b200.adapter: deform_inputs,
SpatialPriorModule, InteractionBlock -> Injector/Extractor -> MSDeformAttn -> sm_100a kernels). The plain ViT
trunk and the segmentation head are NOT part of the hot path and cannot come from the reference here
(mmcv/mmseg/timm are absent), so they are re-stated minimally: a pre-norm ViT with
F.scaled_dot_product_attention and a light multi-scale stand-in head (1x1 lateral convs + fuse + classifier,
cross-entropy on synthetic labels). The adapter call structure is the reference's ViTAdapter.forward
(segmentation/mmseg_custom/models/backbones/vit_adapter.py:93-137): 4 Injector + 6 Extractor calls per forward.
Batch-sharded data parallel: one process per GPU, DDP gradient all-reduce over NCCL, SyncBatchNorm in the
SPM / output norms as in the reference; per-GPU batch fixed (weak scaling).

`--op ours|ref_cuda` swaps only the sampling core (ours vs the reference's own CUDA kernels from oracle/_ref),
everything else identical, for an A/B at model level.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

CFG = {
    # embed, depth, heads, deform_heads, deform_ratio, interaction_indexes, with_cp
    'S': dict(embed=384, depth=12, heads=6, dheads=6, ratio=1.0, idx=[[0, 2], [3, 5], [6, 8], [9, 11]]),
    'B': dict(embed=768, depth=12, heads=12, dheads=12, ratio=0.5, idx=[[0, 2], [3, 5], [6, 8], [9, 11]]),
    'L': dict(embed=1024, depth=24, heads=16, dheads=16, ratio=0.5, idx=[[0, 5], [6, 11], [12, 17], [18, 23]]),
}


class ViTBlock(nn.Module):
    """Plain pre-norm transformer block (stand-in for timm's Block; dense work -> cuBLAS / SDPA)."""

    def __init__(self, dim, heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.fc1 = nn.Linear(dim, int(dim * mlp_ratio))
        self.fc2 = nn.Linear(int(dim * mlp_ratio), dim)
        self.heads = heads

    def forward(self, x, H, W):
        B, N, C = x.shape
        qkv = self.qkv(self.norm1(x)).view(B, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
        x = x + self.proj(a.transpose(1, 2).reshape(B, N, C))
        return x + self.fc2(F.gelu(self.fc1(self.norm2(x))))


class MiniViTAdapter(nn.Module):
    def __init__(self, variant, sync_bn, with_cp=False):
        super().__init__()
        from vit_adapter_b200.adapter import InteractionBlock, SpatialPriorModule
        c = CFG[variant]
        d = c['embed']
        if sync_bn == 'nosync':
            from vit_adapter_b200.adapter import SyncBatchNormNoHostSync as bn
        else:
            bn = nn.SyncBatchNorm if sync_bn else nn.BatchNorm2d
        self.embed = d
        self.idx = c['idx']
        self.patch_embed = nn.Conv2d(3, d, 16, 16)
        self.pos_embed = nn.Parameter(torch.zeros(1, 14 * 14, d))
        self.blocks = nn.ModuleList([ViTBlock(d, c['heads']) for _ in range(c['depth'])])
        self.level_embed = nn.Parameter(torch.randn(3, d) * 0.02)
        self.spm = SpatialPriorModule(inplanes=64, embed_dim=d, norm_layer=bn)
        self.interactions = nn.ModuleList([
            InteractionBlock(dim=d, num_heads=c['dheads'], n_points=4, init_values=0., drop_path=0.,
                             with_cffn=True, cffn_ratio=0.25, deform_ratio=c['ratio'],
                             extra_extractor=(i == len(self.idx) - 1), with_cp=with_cp)
            for i in range(len(self.idx))])
        self.up = nn.ConvTranspose2d(d, d, 2, 2)
        self.norms = nn.ModuleList([bn(d) for _ in range(4)])

    def forward(self, x):
        from vit_adapter_b200.adapter import deform_inputs
        di1, di2 = deform_inputs(x)
        c1, c2, c3, c4 = self.spm(x)
        n2, n3 = c2.size(1), c3.size(1)
        c = torch.cat([c2 + self.level_embed[0], c3 + self.level_embed[1], c4 + self.level_embed[2]], dim=1)
        x = self.patch_embed(x)
        bs, dim, H, W = x.shape
        x = x.flatten(2).transpose(1, 2)
        pos = F.interpolate(self.pos_embed.reshape(1, 14, 14, dim).permute(0, 3, 1, 2), size=(H, W), mode='bicubic',
                            align_corners=False).reshape(1, dim, H * W).permute(0, 2, 1)
        x = x + pos
        outs = []
        for i, layer in enumerate(self.interactions):
            a, b = self.idx[i]
            x, c = layer(x, c, self.blocks[a:b + 1], di1, di2, H, W)
            outs.append(x.transpose(1, 2).reshape(bs, dim, H, W))
        c2, c3, c4 = c[:, :n2], c[:, n2:n2 + n3], c[:, n2 + n3:]
        c2 = c2.transpose(1, 2).reshape(bs, dim, H * 2, W * 2)
        c3 = c3.transpose(1, 2).reshape(bs, dim, H, W)
        c4 = c4.transpose(1, 2).reshape(bs, dim, H // 2, W // 2)
        c1 = self.up(c2) + c1
        x1, x2, x3, x4 = outs
        c1 = c1 + F.interpolate(x1, scale_factor=4, mode='bilinear', align_corners=False)
        c2 = c2 + F.interpolate(x2, scale_factor=2, mode='bilinear', align_corners=False)
        c3 = c3 + x3
        c4 = c4 + F.interpolate(x4, scale_factor=0.5, mode='bilinear', align_corners=False)
        return [n(f) for n, f in zip(self.norms, (c1, c2, c3, c4))]


class StandInHead(nn.Module):
    """NOT UperNet: a light FPN-style fuse so that every pyramid level receives a gradient."""

    def __init__(self, dim, classes=150, width=256):
        super().__init__()
        self.lat = nn.ModuleList([nn.Conv2d(dim, width, 1) for _ in range(4)])
        self.fuse = nn.Conv2d(width, width, 3, padding=1)
        self.cls = nn.Conv2d(width, classes, 1)

    def forward(self, feats):
        size = feats[0].shape[-2:]
        y = 0
        for f, l in zip(feats, self.lat):
            y = y + F.interpolate(l(f), size=size, mode='bilinear', align_corners=False)
        return self.cls(F.relu(self.fuse(y)))


class Net(nn.Module):
    def __init__(self, variant, sync_bn, with_cp):
        super().__init__()
        self.backbone = MiniViTAdapter(variant, sync_bn, with_cp)
        self.head = StandInHead(self.backbone.embed)

    def forward(self, img):
        return self.head(self.backbone(img))


def use_reference_cuda_core():
    """Route MSDeformAttn's sampling core to the reference's own CUDA kernels (oracle/_ref) for the A/B."""
    from oracle import refcuda
    import vit_adapter_b200.modules.ms_deform_attn as mod

    class RefFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, value, shapes, lsi, loc, aw, step):
            value, loc, aw = value.float().contiguous(), loc.float().contiguous(), aw.float().contiguous()
            ctx.save_for_backward(value, shapes, lsi, loc, aw)
            ctx.step = step
            return refcuda.forward(value, shapes, lsi, loc, aw, step)

        @staticmethod
        def backward(ctx, go):
            value, shapes, lsi, loc, aw = ctx.saved_tensors
            gv, gl, ga = refcuda.backward(value, shapes, lsi, loc, aw, go.float().contiguous(), ctx.step)
            return gv, None, None, gl, ga, None

    mod.MSDeformAttnFunction = RefFn


def add_step_args(ap):
    ap.add_argument('--variant', default='B', choices=sorted(CFG))
    ap.add_argument('--mode', default='train', choices=['train', 'infer'])
    ap.add_argument('--image', type=int, default=512)
    ap.add_argument('--batch', type=int, default=2, help='images per GPU (reference: 2 for B/S, 1 for L)')
    ap.add_argument('--amp', action='store_true', help='torch.autocast(bfloat16) + bf16 sampling core')
    ap.add_argument('--with-cp', action='store_true', help='activation checkpointing in Injector/Extractor (L configs)')
    ap.add_argument('--op', default='ours', choices=['ours', 'ref_cuda'])
    ap.add_argument('--reference-sequence', action='store_true',
                    help="the adapter exactly as the reference runs it: reference CUDA core, torch LayerNorm, the DWConv "
                         "slice/transpose/conv2d sequence, separate offset/weight linears + softmax (implies --op ref_cuda)")
    ap.add_argument('--tf32', action='store_true',
                    help='fp32 GEMMs / convolutions on TF32 tensor cores - the default of the torch 1.9.0 the reference pins '
                         '(segmentation/README.md:24); off by default in the torch of this image')
    ap.add_argument('--graph', action='store_true',
                    help='capture the whole step (forward, backward, optimizer, and under torchrun the DDP all-reduce) in ONE CUDA graph and replay it')
    ap.add_argument('--bucket-cap-mb', type=int, default=None, help='DDP bucket size (default: torch, 25 MB)')
    ap.add_argument('--grad-bf16', action='store_true', help="DDP's stock bf16_compress_hook on the gradient all-reduce")
    ap.add_argument('--bn', default='sync', choices=['sync', 'local', 'nosync'],
                    help="'local': plain BatchNorm under DDP (NOT the reference's training recipe) - isolates what SyncBatchNorm's "
                         "forward costs the eager step: three host synchronisations per layer; 'nosync': "
                         'vit_adapter_b200.adapter.SyncBatchNormNoHostSync (same statistics, no host synchronisation)')


def step_bench(variant='B', mode='train', image=512, batch=2, amp=False, with_cp=False, op='ours', reference_sequence=False,
               tf32=False, graph=False, steps=10, warmup=3, measure_comm=True, bn='sync', bucket_cap_mb=None, grad_bf16=False):
    """One model-level measurement on the CURRENT device / process group (the caller owns torch.distributed): returns the
    result dict. Under a process group the net is wrapped in DDP (gradient all-reduce over NCCL) with SyncBatchNorm, as the
    reference trains (segmentation/dist_train.sh:8-9, configs/_base_/default_runtime.py:9).

    measure_comm (eager DDP training only): the same K steps are timed again under `model.no_sync()` - identical compute,
    no gradient all-reduce - and the difference is reported as the all-reduce time that is NOT hidden behind the backward
    (`allreduce_ms_exposed`)."""
    import vit_adapter_b200 as vab
    from vit_adapter_b200 import _cabi
    from vit_adapter_b200.graphs import GraphedStep
    from vit_adapter_b200.sharding import max_over_ranks

    if tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    ddp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    world = dist.get_world_size() if ddp else 1
    rank = dist.get_rank() if ddp else 0
    dev = torch.device('cuda', torch.cuda.current_device())
    if reference_sequence:
        op = 'ref_cuda'
    if op == 'ref_cuda':
        use_reference_cuda_core()
    vab.set_amp_value_dtype(torch.bfloat16 if amp else torch.float32)

    torch.manual_seed(1234 + rank)
    net = Net(variant, sync_bn=(bn if ddp and bn in ('sync', 'nosync') else False), with_cp=with_cp).to(dev)
    if reference_sequence:
        for m in net.modules():
            for flag in ('fused_norm', 'token_kernel', 'fused', 'merge_query_linears', 'colsum_bias_grad'):
                if hasattr(m, flag):
                    setattr(m, flag, False)
    n_params = sum(p.numel() for p in net.parameters())
    n_adapter = sum(p.numel() for n, p in net.named_parameters() if 'interactions' in n or 'spm' in n)
    model = net
    ddp_kw = {} if bucket_cap_mb is None else {'bucket_cap_mb': bucket_cap_mb}
    if ddp and mode == 'train':
        if graph:
            # DDP under whole-step capture (torch docs, "Usage with DistributedDataParallel"): construct DDP on a side
            # stream, and warm up >= 11 eager iterations so that bucket rebuilding is over before the capture
            side0 = torch.cuda.Stream()
            with torch.cuda.stream(side0):
                model = nn.parallel.DistributedDataParallel(net, device_ids=[dev.index], gradient_as_bucket_view=True, **ddp_kw)
            torch.cuda.current_stream().wait_stream(side0)
            warmup = max(warmup, 11)
        else:
            model = nn.parallel.DistributedDataParallel(net, device_ids=[dev.index], gradient_as_bucket_view=True, **ddp_kw)
        if grad_bf16:   # DDP's stock communication hook: gradients cross NVLink in bf16 (half the all-reduce bytes)
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
            model.register_comm_hook(None, default_hooks.bf16_compress_hook)
    opt = torch.optim.AdamW(net.parameters(), lr=6e-5, weight_decay=0.01, fused=True, capturable=graph) if mode == 'train' else None
    img = torch.randn(batch, 3, image, image, device=dev)
    lab = torch.randint(0, 150, (batch, image // 4, image // 4), device=dev)

    def forward_loss():
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
            if mode == 'train':
                return F.cross_entropy(model(img).float(), lab)
            with torch.no_grad():
                return model(img)

    def eager_step():
        out = forward_loss()
        if mode == 'train':
            opt.zero_grad(set_to_none=True)
            out.backward()
            opt.step()
        return out

    def captured_body():   # the eager step minus zero_grad: the captured backward owns the gradient buffers
        out = forward_loss()
        if mode == 'train':
            out.backward()
            opt.step()
        return out

    if mode == 'infer':
        net.eval()

    def barrier():
        if ddp:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        eager_step()
    barrier()
    step = eager_step
    launches_per_replay = 0
    if graph:
        # whole-step capture: at 2 images per GPU the step is ~1 900 kernels of a few microseconds each and the Python /
        # launch path is as long as the GPU work; one graph launch replaces it. Everything on the path is capturable:
        # no host sync (MSDeformAttn's shape check is memoised), no allocation outside torch's graph pool, the library
        # launches on the capturing stream it is handed.
        gs = GraphedStep(captured_body, warmup=(11 if ddp else 3), warm_fn=eager_step,
                         before_capture=(lambda: opt.zero_grad(set_to_none=True)) if opt is not None else None)
        launches_per_replay = gs.launches
        step = gs
        for _ in range(2):
            step()
        barrier()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev), out

    l0 = _cabi.launch_count()
    ms, out = timed(step, steps)
    launches = _cabi.launch_count() - l0 + launches_per_replay * steps
    allreduce_bytes = 4 * n_params if (ddp and mode == 'train') else 0   # fp32 gradient buckets, every parameter once per step
    exposed = None
    if ddp and mode == 'train' and not graph and measure_comm:
        def nosync_step():
            with model.no_sync():
                return eager_step()
        for _ in range(2):
            nosync_step()
        ms_nosync, _ = timed(nosync_step, steps)
        exposed = max(0.0, (ms - ms_nosync) / steps)
    res = {
        'metric': 'vit_adapter_%s_%s_img_per_s' % (variant, mode), 'value': world * batch * steps / (ms * 1e-3),
        'unit': 'img/s', 'n_gpus': world, 'steps': steps, 'warmup': warmup, 'ms_per_step': ms / steps,
        'higher_is_better': True, 'scaling': 'weak', 'dtype': 'bf16-autocast' if amp else 'f32', 'data': 'synthetic',
        'op': op, 'adapter': 'reference op sequence' if reference_sequence else 'this repo (fused norm / dwconv / softmax+locations)',
        'msda_kernel_launches': launches, 'cuda_graph': bool(graph), 'tf32_gemm': bool(tf32),
        'allreduce_bytes': allreduce_bytes // (2 if grad_bf16 else 1), 'ddp_bucket_cap_mb': bucket_cap_mb, 'ddp_grad_bf16': bool(grad_bf16), 'allreduce_ms_exposed': exposed, 'batchnorm': ('SyncBatchNorm' if bn == 'sync' else 'SyncBatchNormNoHostSync' if bn == 'nosync' else 'BatchNorm2d') if ddp else 'BatchNorm2d',
        'config': {'workload': 'ViT-Adapter-%s backbone (this repo\'s adapter modules + MSDeformAttn) + stand-in head, %dx%d, '
                               '%d img/GPU, %s' % (variant, image, image, batch, mode),
                   'params_total': n_params, 'params_adapter': n_adapter, 'with_cp': with_cp,
                   'parallelism': 'dp%d (DDP all-reduce over NCCL, SyncBN)' % world if world > 1 else 'single GPU',
                   'note': 'ViT trunk and head are minimal stand-ins (mmcv/mmseg/timm absent); adapter path is the drop-in code'},
        'final': float(out.detach().float().mean()) if torch.is_tensor(out) else None,
    }
    vab.set_amp_value_dtype(torch.float32)
    del model, net, opt, step
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return res


def main():
    sys.stdout.flush()
    real_stdout = os.dup(1)   # NCCL prints its version banner on fd 1: keep stdout for the one JSON line
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    add_step_args(ap)
    args = ap.parse_args()

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if args.graph:
            os.environ.setdefault('TORCH_NCCL_ASYNC_ERROR_HANDLING', '0')   # NCCL work captured in a graph has no watchdog-visible events
        dist.init_process_group('nccl', device_id=dev)
    res = step_bench(variant=args.variant, mode=args.mode, image=args.image, batch=args.batch, amp=args.amp, with_cp=args.with_cp,
                     op=args.op, reference_sequence=args.reference_sequence, tf32=args.tf32, graph=args.graph, steps=args.steps,
                     warmup=args.warmup, bn=args.bn, bucket_cap_mb=args.bucket_cap_mb, grad_bf16=args.grad_bf16)
    if rank == 0:
        os.write(real_stdout, (json.dumps(res) + '\n').encode())
    if world > 1:
        if args.graph:
            # tearing the NCCL communicator down after its collectives were captured in a graph hung on the test box:
            # the result is out, every rank is synchronised - leave without the destructor
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
