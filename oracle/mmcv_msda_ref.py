"""CPU restatement of mmcv's `MultiScaleDeformableAttention.forward`. TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED against mmcv itself: mmcv-full==1.4.2 (segmentation/README.md:25) is an un-vendored dependency of the
reference, absent from /root/reference and from this image, and the reference holds no test or golden vector for it
(SURVEY.md §8(c)). This file restates the module's published forward
    value = value_proj(value) (masked by key_padding_mask), offsets / softmaxed weights from (query + query_pos),
    loc = ref + offsets / (W_l, H_l)            (2-d reference points)
        = ref[:2] + offsets / P * ref[2:] * 0.5 (4-d)
    out = output_proj(deformable sampling) ; return dropout(out) + identity, (num_query, bs, C) unless batch_first
as it is exercised by the reference's call site (msdeformattn_pixel_decoder.py:231-242: query_pos = level positional
encodings, per-level reference points, key_padding_mask None) on top of the reference's own pure-torch sampling core
(oracle/core_pytorch.py, pinned on the reference's golden vectors).
"""
import torch
import torch.nn.functional as F

from .core_pytorch import ms_deform_attn_core


def forward(params, query, reference_points, spatial_shapes, num_heads, num_levels, num_points, value=None, identity=None,
            query_pos=None, key_padding_mask=None, batch_first=False):
    """params: dict with sampling_offsets.{weight,bias}, attention_weights.*, value_proj.*, output_proj.* (dropout = 0)."""
    if value is None:
        value = query
    if identity is None:
        identity = query
    if query_pos is not None:
        query = query + query_pos
    if not batch_first:
        query, value = query.permute(1, 0, 2), value.permute(1, 0, 2)
    bs, nq, _ = query.shape
    nv = value.shape[1]
    M, L, P = num_heads, num_levels, num_points
    v = F.linear(value, params['value_proj.weight'], params['value_proj.bias'])
    if key_padding_mask is not None:
        v = v.masked_fill(key_padding_mask[..., None], 0.0)
    v = v.view(bs, nv, M, -1)
    off = F.linear(query, params['sampling_offsets.weight'], params['sampling_offsets.bias']).view(bs, nq, M, L, P, 2)
    aw = F.linear(query, params['attention_weights.weight'], params['attention_weights.bias']).view(bs, nq, M, L * P)
    aw = aw.softmax(-1).view(bs, nq, M, L, P)
    if reference_points.shape[-1] == 2:
        norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1).to(off.dtype)
        loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = reference_points[:, :, None, :, None, :2] + off / P * reference_points[:, :, None, :, None, 2:] * 0.5
    out = ms_deform_attn_core(v, spatial_shapes, loc, aw)
    out = F.linear(out, params['output_proj.weight'], params['output_proj.bias'])
    if not batch_first:
        out = out.permute(1, 0, 2)
    return out + identity
