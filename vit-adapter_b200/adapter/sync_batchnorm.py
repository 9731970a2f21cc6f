"""SyncBatchNorm for the adapter's SpatialPriorModule that never synchronises the host.

The reference builds the SPM with `nn.SyncBatchNorm` (*/mm*_custom/models/backbones/adapter_modules.py:197-223). torch's
implementation gathers (mean, invstd, count) from every rank and then drops the ranks whose batch was EMPTY with a boolean
mask - `count_all[mask]` has a data-dependent shape, i.e. a device->host synchronisation, three times per layer and forward
(profiles/r2_ddp_eager_profile_2gpu.jsonl: 30 cudaStreamSynchronize per training step of ViT-Adapter-B). Every one of them
drains the asynchronously queued kernels of a step that is launch-bound at 2 images per GPU, which is what limited the eager
data-parallel step to 0.75 scaling efficiency at 2 GPUs; under CUDA-graph capture torch itself skips that branch.

In data-parallel training every rank holds at least one sample, so the mask is all-true and can go. This module is
nn.SyncBatchNorm with exactly that one step removed: same collectives (one all_gather in forward, one all_reduce in backward),
same ATen kernels (batch_norm_stats / batch_norm_gather_stats_with_counts / batch_norm_elemt / batch_norm_backward_*),
same state-dict keys, bit-identical results. A rank with an empty batch is NOT supported (use nn.SyncBatchNorm there).
"""
import torch
import torch.distributed as dist
from torch import nn
from torch.autograd import Function


class _SyncBatchNormNoMask(Function):
    @staticmethod
    def forward(ctx, input, weight, bias, running_mean, running_var, eps, momentum, process_group, world_size):
        if not (input.is_contiguous(memory_format=torch.channels_last) or input.is_contiguous(memory_format=torch.channels_last_3d)):
            input = input.contiguous()
        if weight is not None:
            weight = weight.contiguous()
        num_channels = input.shape[1]
        mean, invstd = torch.batch_norm_stats(input, eps)
        count = torch.full((1,), input.numel() // input.size(1), dtype=mean.dtype, device=mean.device)
        combined = torch.cat([mean, invstd, count], dim=0)                      # C, C, 1 -> 2C + 1
        combined_flat = torch.empty(1, combined.numel() * world_size, dtype=combined.dtype, device=combined.device)
        dist.all_gather_into_tensor(combined_flat, combined, process_group, async_op=False)
        combined = combined_flat.reshape(world_size, -1)
        mean_all, invstd_all, count_all = torch.split(combined, num_channels, dim=1)
        counts = count_all.view(-1)                                              # (no empty-rank mask: see the module docstring)
        if running_mean is not None and counts.dtype != running_mean.dtype:
            counts = counts.to(running_mean.dtype)
        mean, invstd = torch.batch_norm_gather_stats_with_counts(input, mean_all, invstd_all, running_mean, running_var,
                                                                  momentum, eps, counts)
        ctx.save_for_backward(input, weight, mean, invstd, count_all.to(torch.int32))
        ctx.process_group = process_group
        return torch.batch_norm_elemt(input, weight, bias, mean, invstd, eps)

    @staticmethod
    def backward(ctx, grad_output):
        if not (grad_output.is_contiguous(memory_format=torch.channels_last) or grad_output.is_contiguous(memory_format=torch.channels_last_3d)):
            grad_output = grad_output.contiguous()
        saved_input, weight, mean, invstd, count_tensor = ctx.saved_tensors
        grad_input = None
        sum_dy, sum_dy_xmu, grad_weight, grad_bias = torch.batch_norm_backward_reduce(
            grad_output, saved_input, mean, invstd, weight, ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        if ctx.needs_input_grad[0]:
            num_channels = sum_dy.shape[0]
            combined = torch.cat([sum_dy, sum_dy_xmu], dim=0)
            dist.all_reduce(combined, dist.ReduceOp.SUM, ctx.process_group, async_op=False)
            sum_dy, sum_dy_xmu = torch.split(combined, num_channels)
            if weight is not None and weight.dtype != mean.dtype:
                weight = weight.to(mean.dtype)
            grad_input = torch.batch_norm_backward_elemt(grad_output, saved_input, mean, invstd, weight, sum_dy, sum_dy_xmu, count_tensor)
        if weight is None or not ctx.needs_input_grad[1]:
            grad_weight = None
        if weight is None or not ctx.needs_input_grad[2]:
            grad_bias = None
        return grad_input, grad_weight, grad_bias, None, None, None, None, None, None


class SyncBatchNormNoHostSync(nn.SyncBatchNorm):
    """Drop-in for nn.SyncBatchNorm when every rank holds at least one sample (data-parallel training)."""

    def forward(self, input):
        sync = (self.training and input.is_cuda and input.numel() > 0 and self.momentum is not None
                and dist.is_available() and dist.is_initialized())
        if sync:
            group = self.process_group if self.process_group is not None else dist.group.WORLD
            world = dist.get_world_size(group)
            sync = world > 1
        if not sync:
            return super().forward(input)
        self._check_input_dim(input)
        self._check_non_zero_input_channels(input)
        if self.track_running_stats:
            self.num_batches_tracked.add_(1)
        running_mean = self.running_mean if self.track_running_stats else None
        running_var = self.running_var if self.track_running_stats else None
        return _SyncBatchNormNoMask.apply(input, self.weight, self.bias, running_mean, running_var, self.eps, self.momentum,
                                          group, world)
