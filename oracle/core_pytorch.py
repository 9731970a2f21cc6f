"""Restatement of the reference's CPU path for this op. TEST INFRASTRUCTURE ONLY.

Follows ms_deform_attn_core_pytorch (detection/ops/functions/ms_deform_attn_func.py:49-71): per level,
reshape the value slab to [N*M, D, H, W], sample it with F.grid_sample(bilinear, zeros padding,
align_corners=False) at grid = 2*loc - 1, weight the L*P samples with the attention weights and sum.
Differentiable through autograd (that is the reference's CPU backward too). Pinned against the real
reference function by tests/test_oracle_golden.py (golden vectors from tests/golden/make_golden.py).
"""
import torch
import torch.nn.functional as F


def ms_deform_attn_core(value, spatial_shapes, sampling_locations, attention_weights):
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    hw = [(int(h), int(w)) for h, w in spatial_shapes]
    grids = sampling_locations * 2 - 1                               # [0,1] -> [-1,1]
    per_level = []
    start = 0
    for lvl, (H, W) in enumerate(hw):
        slab = value[:, start:start + H * W]                        # [N, H*W, M, D]
        start += H * W
        img = slab.permute(0, 2, 3, 1).reshape(N * M, D, H, W)       # [N*M, D, H, W]
        grid = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)
        per_level.append(F.grid_sample(img, grid, mode='bilinear', padding_mode='zeros',
                                       align_corners=False))         # [N*M, D, Lq, P]
    sampled = torch.stack(per_level, dim=-2).reshape(N * M, D, Lq, L * P)
    w = attention_weights.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L * P)
    out = (sampled * w).sum(-1)                                      # [N*M, D, Lq]
    return out.view(N, M * D, Lq).transpose(1, 2).contiguous()


def forward_backward(value, spatial_shapes, sampling_locations, attention_weights, grad_out):
    """out and (grad_value, grad_loc, grad_aw) via autograd — the reference CPU path's fwd+bwd."""
    v = value.detach().clone().requires_grad_(True)
    l = sampling_locations.detach().clone().requires_grad_(True)
    a = attention_weights.detach().clone().requires_grad_(True)
    out = ms_deform_attn_core(v, spatial_shapes, l, a)
    out.backward(grad_out)
    return out.detach(), (v.grad, l.grad, a.grad)
