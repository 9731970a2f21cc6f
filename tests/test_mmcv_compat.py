"""mmcv-API `MultiScaleDeformableAttention` (Mask2Former pixel decoder's deformable encoder) on the
B200 kernels, against oracle/mmcv_msda_ref.py (parity unpinned against mmcv itself - see that file's header)."""
import inspect

import pytest
import torch

from oracle import mmcv_msda_ref
from vit_adapter_b200.mmcv_compat import MultiScaleDeformableAttention, MultiScaleDeformableAttnFunction


def _encoder_inputs(bs, shapes, C, dtype, seed=11, batch_first=False):
    """What MSDeformAttnPixelDecoder.forward feeds each encoder layer (msdeformattn_pixel_decoder.py:176-242): the three
    levels flattened and concatenated, per-level reference points = cell centres scaled by the valid ratios (all 1)."""
    g = torch.Generator().manual_seed(seed)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    n = int(shapes_t.prod(1).sum())
    ref = []
    for (H, W) in shapes:
        ys, xs = torch.meshgrid((torch.arange(H, dtype=dtype) + 0.5) / H, (torch.arange(W, dtype=dtype) + 0.5) / W, indexing='ij')
        ref.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    ref = torch.cat(ref, 0)[None, :, None].repeat(bs, 1, len(shapes), 1)       # [bs, n, L, 2]
    q = torch.randn(n, bs, C, generator=g, dtype=dtype)
    pos = 0.5 * torch.randn(n, bs, C, generator=g, dtype=dtype)
    if batch_first:
        q, pos = q.permute(1, 0, 2).contiguous(), pos.permute(1, 0, 2).contiguous()
    return q, pos, ref, shapes_t, lsi


def _randomise(m, seed=5):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        m.sampling_offsets.weight.copy_(0.02 * torch.randn(m.sampling_offsets.weight.shape, generator=g))
        m.attention_weights.weight.copy_(0.2 * torch.randn(m.attention_weights.weight.shape, generator=g))
        m.attention_weights.bias.copy_(0.2 * torch.randn(m.attention_weights.bias.shape, generator=g))


def test_mmcv_surface():
    sig = inspect.signature(MultiScaleDeformableAttention.__init__)
    assert list(sig.parameters)[1:] == ['embed_dims', 'num_heads', 'num_levels', 'num_points', 'im2col_step', 'dropout',
                                        'batch_first', 'norm_cfg', 'init_cfg']
    assert [sig.parameters[k].default for k in list(sig.parameters)[1:8]] == [256, 8, 4, 4, 64, 0.1, False]
    fsig = inspect.signature(MultiScaleDeformableAttention.forward)
    assert list(fsig.parameters)[1:10] == ['query', 'key', 'value', 'identity', 'query_pos', 'key_padding_mask',
                                           'reference_points', 'spatial_shapes', 'level_start_index']
    m = MultiScaleDeformableAttention(embed_dims=64, num_heads=4, num_levels=3, num_points=4, dropout=0.0)
    assert sorted(m.state_dict()) == sorted(
        f'{n}.{k}' for n in ('sampling_offsets', 'attention_weights', 'value_proj', 'output_proj') for k in ('weight', 'bias'))
    assert m.sampling_offsets.weight.abs().max() == 0 and m.attention_weights.weight.abs().max() == 0
    m.init_weights()
    with pytest.raises(ValueError):
        MultiScaleDeformableAttention(embed_dims=30, num_heads=4)
    assert MultiScaleDeformableAttnFunction.apply is not None


def test_cpu_tensors_raise():
    m = MultiScaleDeformableAttention(embed_dims=32, num_heads=2, num_levels=1, num_points=4, dropout=0.0)
    q, pos, ref, shapes, lsi = _encoder_inputs(1, [(4, 4)], 32, torch.float32)
    with pytest.raises(RuntimeError, match='CPU'):
        m(q, query_pos=pos, reference_points=ref, spatial_shapes=shapes, level_start_index=lsi)


@pytest.mark.gpu
@pytest.mark.parametrize('batch_first', [False, True], ids=['seq-first', 'batch-first'])
@pytest.mark.parametrize('fused', [True, False], ids=['fused', 'unfused'])
def test_encoder_layer_attention_fp64_like_oracle(batch_first, fused):
    shapes = [(12, 10), (6, 5), (3, 3)]
    m = MultiScaleDeformableAttention(embed_dims=64, num_heads=4, num_levels=3, num_points=4, dropout=0.0, batch_first=batch_first)
    _randomise(m)
    m.fused = fused
    q, pos, ref, shapes_t, lsi = _encoder_inputs(2, shapes, 64, torch.float32, batch_first=batch_first)
    params = {k: v.detach().double() for k, v in m.state_dict().items()}
    qd = q.double().requires_grad_()
    want = mmcv_msda_ref.forward(params, qd, ref.double(), shapes_t, 4, 3, 4, query_pos=pos.double(), batch_first=batch_first)
    gout = torch.randn(want.shape, generator=torch.Generator().manual_seed(2))
    want.backward(gout.double())
    md = m.cuda()
    qc = q.cuda().requires_grad_()
    got = md(qc, query_pos=pos.cuda(), reference_points=ref.cuda(), spatial_shapes=shapes_t.cuda(), level_start_index=lsi.cuda())
    got.backward(gout.cuda())
    torch.testing.assert_close(got.detach().cpu().double(), want.detach(), rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(qc.grad.cpu().double(), qd.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_mask2former_encoder_shape_bf16_autocast():
    """The config's shape: embed_dims 1024, 32 heads (D = 32), 3 levels of an 896^2 crop / (8, 16, 32), bs 1."""
    shapes = [(112, 112), (56, 56), (28, 28)]
    m = MultiScaleDeformableAttention(embed_dims=1024, num_heads=32, num_levels=3, num_points=4, dropout=0.0)
    _randomise(m)
    q, pos, ref, shapes_t, lsi = _encoder_inputs(1, shapes, 1024, torch.float32, seed=3)
    params = {k: v.detach() for k, v in m.state_dict().items()}
    want = mmcv_msda_ref.forward(params, q, ref, shapes_t, 32, 3, 4, query_pos=pos)
    md = m.cuda()
    from vit_adapter_b200 import _cabi
    n0 = _cabi.launch_count()
    got32 = md(q.cuda(), query_pos=pos.cuda(), reference_points=ref.cuda(), spatial_shapes=shapes_t.cuda(), level_start_index=lsi.cuda())
    assert _cabi.launch_count() - n0 == 1
    # TF32 is off by default for matmul in torch: fp32 GEMMs are exact enough for a 1e-4 comparison
    torch.testing.assert_close(got32.cpu(), want, rtol=1e-4, atol=1e-4)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        got16 = md(q.cuda(), query_pos=pos.cuda(), reference_points=ref.cuda(), spatial_shapes=shapes_t.cuda(), level_start_index=lsi.cuda())
    err = (got16.float().cpu() - want).abs().max() / want.abs().max()
    assert err < 3e-2, err
