"""Fused entry points (softmax + location arithmetic inside the sampling kernel) against the unfused op
sequence the reference module executes (detection/ops/modules/ms_deform_attn.py:108-128)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

import vit_adapter_b200 as vab
from vit_adapter_b200 import _cabi

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _scale(t):
    return float(t.abs().max()) + 1e-30


def _raw_inputs(N, M, D, Lq, shapes, P, seed, ref_batch, ref_levels, dtype):
    g = torch.Generator().manual_seed(seed)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    L = shapes_t.shape[0]
    S = int(shapes_t.prod(1).sum())
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    value = torch.randn(N, S, M, D, generator=g).to(dtype)
    ref = torch.rand(ref_batch, Lq, ref_levels, 2, generator=g) * 1.1 - 0.05      # some reference points off-image
    offsets = torch.randn(N, Lq, M, L, P, 2, generator=g) * 2.5                     # pixels of each level
    logits = torch.randn(N, Lq, M, L * P, generator=g) * 2.0
    grad_out = torch.randn(N, Lq, M * D, generator=g).to(dtype)
    return [t.to(DEV) for t in (value, shapes_t, lsi, ref, offsets, logits, grad_out)]


def _unfused(value, shapes, lsi, ref, offsets, logits):
    """The reference module's op sequence (ms_deform_attn.py:110-128) in torch + the plain Function."""
    N, Lq, M, L, P, _ = offsets.shape
    weights = F.softmax(logits, -1).view(N, Lq, M, L, P)
    wh = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
    loc = ref[:, :, None, :, None, :] + offsets / wh[None, None, None, :, None, :]
    return vab.MSDeformAttnFunction.apply(value, shapes, lsi, loc.contiguous(), weights.contiguous(), 64)


CASES = [
    # name, N, M, D, Lq, shapes, ref_batch, ref_levels
    ('injector-B', 2, 12, 32, 256, [(32, 32), (16, 16), (8, 8)], 1, 1),
    ('extractor-B', 2, 12, 32, 1344, [(16, 16)], 1, 1),
    ('injector-S', 2, 6, 64, 100, [(20, 20), (10, 10), (5, 5)], 2, 3),
    ('extractor-ragged', 3, 5, 32, 37, [(7, 9)], 3, 1),
]


@pytest.mark.parametrize('cfg', CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16], ids=['f32', 'bf16', 'f16'])
def test_fused_matches_unfused(cfg, dtype):
    _, N, M, D, Lq, shapes, rb, rl = cfg
    value, shapes_t, lsi, ref, offsets, logits, go = _raw_inputs(N, M, D, Lq, shapes, 4, 31, rb, rl, dtype)
    assert _cabi.fused_supported(value, len(shapes), 4)
    if dtype == torch.float16:
        vab.set_amp_value_dtype(torch.float16)   # fp16 tensors run natively only when asked (default: fp32 up-cast, as the reference)
    try:
        v1, o1, l1 = value.clone().requires_grad_(), offsets.clone().requires_grad_(), logits.clone().requires_grad_()
        out1 = vab.MSDeformAttnFusedFunction.apply(v1, shapes_t, lsi, ref, o1, l1)
        out1.backward(go)
        v2, o2, l2 = value.clone().requires_grad_(), offsets.clone().requires_grad_(), logits.clone().requires_grad_()
        out2 = _unfused(v2, shapes_t, lsi, ref, o2, l2)
        out2.backward(go)
        torch.cuda.synchronize()
    finally:
        vab.set_amp_value_dtype(torch.float32)
    assert out1.dtype == dtype and v1.grad.dtype == dtype
    ft, gt = {torch.float32: (1e-5, 1e-4), torch.bfloat16: (1e-2, 1e-2), torch.float16: (2e-3, 2e-3)}[dtype]
    torch.testing.assert_close(out1.float(), out2.float(), rtol=ft, atol=ft * max(1.0, _scale(out2.float())))
    torch.testing.assert_close(v1.grad.float(), v2.grad.float(), rtol=gt, atol=gt * _scale(v2.grad.float()))
    torch.testing.assert_close(o1.grad, o2.grad, rtol=gt, atol=gt * _scale(o2.grad))
    torch.testing.assert_close(l1.grad, l2.grad, rtol=gt, atol=gt * _scale(l2.grad))


@pytest.mark.parametrize('cfg', CASES[:3], ids=[c[0] for c in CASES[:3]])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['f32', 'bf16'])
def test_merged_gemm_layout_is_bit_identical_to_separate_tensors(cfg, dtype):
    """Offsets and logits read in place from one [N, Lq, M*L*P*3] buffer (the merged GEMM output) must give exactly
    the results of the two separate tensors, and the merged gradient must be their concatenation."""
    _, N, M, D, Lq, shapes, rb, rl = cfg
    L = len(shapes)
    value, shapes_t, lsi, ref, offsets, logits, go = _raw_inputs(N, M, D, Lq, shapes, 4, 35, rb, rl, dtype)
    merged = torch.cat([offsets.reshape(N, Lq, -1), logits.reshape(N, Lq, -1)], -1).contiguous()
    v1, o1, l1 = value.clone().requires_grad_(), offsets.clone().requires_grad_(), logits.clone().requires_grad_()
    out1 = vab.MSDeformAttnFusedFunction.apply(v1, shapes_t, lsi, ref, o1, l1)
    out1.backward(go)
    v2, m2 = value.clone().requires_grad_(), merged.clone().requires_grad_()
    out2 = vab.MSDeformAttnMergedFunction.apply(v2, shapes_t, lsi, ref, m2, L, 4)
    out2.backward(go)
    torch.cuda.synchronize()
    assert torch.equal(out1, out2)
    want = torch.cat([o1.grad.reshape(N, Lq, -1), l1.grad.reshape(N, Lq, -1)], -1)
    assert torch.equal(m2.grad, want)
    gt = 1e-4 if dtype == torch.float32 else 1e-2  # atomic order (+ one bf16 rounding of the converted result)
    torch.testing.assert_close(v1.grad.float(), v2.grad.float(), rtol=gt, atol=gt * _scale(v2.grad.float()))


def test_fused_locations_are_bit_identical():
    """With one point carrying all the attention (one-hot logits) the softmax is exactly {0,1}, so any difference
    between fused and unfused outputs could only come from the location arithmetic: there must be none."""
    N, M, D, Lq, shapes = 1, 2, 32, 300, [(40, 40)]
    value, shapes_t, lsi, ref, offsets, logits, _ = _raw_inputs(N, M, D, Lq, shapes, 4, 33, 1, 1, torch.float32)
    logits = torch.full_like(logits, -1e4)
    logits[..., 2] = 0.0
    out1 = vab.MSDeformAttnFusedFunction.apply(value, shapes_t, lsi, ref, offsets, logits)
    out2 = _unfused(value, shapes_t, lsi, ref, offsets, logits)
    assert torch.equal(out1, out2)


def test_fused_unsupported_configurations_fall_back():
    value = torch.zeros(1, 4, 2, 32, device=DEV)
    assert not _cabi.fused_supported(value, 2, 4)          # L = 2 has no fused kernel
    assert not _cabi.fused_supported(value.double(), 1, 4)  # fp64 runs the generic kernels
    assert not _cabi.fused_supported(torch.zeros(1, 4, 2, 24, device=DEV), 1, 4)
    shapes = torch.as_tensor([(2, 2)], dtype=torch.long, device=DEV)
    lsi = torch.zeros(1, dtype=torch.long, device=DEV)
    with pytest.raises(RuntimeError, match='no fused kernel'):
        _cabi.forward_fused(torch.zeros(1, 4, 2, 24, device=DEV), shapes, lsi, torch.zeros(1, 3, 1, 2, device=DEV),
                            torch.zeros(1, 3, 2, 1, 4, 2, device=DEV), torch.zeros(1, 3, 2, 4, device=DEV))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16], ids=['f32', 'bf16-autocast', 'f16-autocast'])
def test_module_fused_vs_reference_sequence(dtype):
    """MSDeformAttn(fused=True) vs the same module running the reference's op sequence (fused=False):
    outputs and every parameter / input gradient."""
    torch.manual_seed(0)
    m = vab.MSDeformAttn(d_model=384, n_levels=3, n_heads=6, n_points=4, ratio=1.0).to(DEV)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.02)
    shapes = torch.as_tensor([(16, 16), (8, 8), (4, 4)], dtype=torch.long, device=DEV)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    query = torch.randn(2, 64, 384, device=DEV)
    feat = torch.randn(2, 336, 384, device=DEV)
    ref = torch.rand(1, 64, 1, 2, device=DEV)
    amp = dtype != torch.float32
    if amp:
        vab.set_amp_value_dtype(dtype)
    try:
        res = []
        for fused, merge in ((True, True), (False, False), (True, False)):
            m.fused = fused
            m.merge_query_linears = merge
            m.zero_grad(set_to_none=True)
            q, f = query.clone().requires_grad_(), feat.clone().requires_grad_()
            n0 = _cabi.launch_count()
            with torch.autocast('cuda', dtype=dtype if amp else torch.bfloat16, enabled=amp):
                out = m(q, ref, f, shapes, lsi)
            out.float().square().mean().backward()
            res.append((out.detach().float(), q.grad, f.grad, {k: p.grad.clone() for k, p in m.named_parameters()},
                        _cabi.launch_count() - n0))
    finally:
        vab.set_amp_value_dtype(torch.float32)
        m.fused = True
        m.merge_query_linears = True
    tol = 1e-4 if not amp else 3e-2
    for other in (0, 2):  # merged-GEMM fused path and separate-linears fused path, each vs the reference sequence
        torch.testing.assert_close(res[other][0], res[1][0], rtol=tol, atol=tol * _scale(res[1][0]))
        torch.testing.assert_close(res[other][1], res[1][1], rtol=tol, atol=tol * _scale(res[1][1]))
        torch.testing.assert_close(res[other][2], res[1][2], rtol=tol, atol=tol * _scale(res[1][2]))
        for k in res[other][3]:
            torch.testing.assert_close(res[other][3][k], res[1][3][k], rtol=tol, atol=tol * _scale(res[1][3][k]), msg=k)
    assert all(r[4] >= 2 for r in res)  # every path ran our kernels


def test_module_fused_matches_reference_golden_f32():
    g = load_golden('module_l3')
    d_model, L, M, P = [int(x) for x in g['cfg']]
    m = vab.MSDeformAttn(d_model, L, M, P, float(g['ratio']))
    m.load_state_dict({k[3:]: v.float() for k, v in g.items() if k.startswith('sd.')}, strict=True)
    m = m.to(DEV)
    shapes = g['shapes'].to(DEV)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    # d_model 32 / 4 heads = 8 channels: no fused kernel -> the module must transparently use the plain path
    out = m(g['query'].float().to(DEV), g['ref_pts'].float().to(DEV), g['feat'].float().to(DEV), shapes, lsi)
    torch.testing.assert_close(out.cpu().double(), g['out_nomask'], rtol=1e-4, atol=1e-4)


def test_learned_reference_points_take_the_differentiable_path():
    """the fused kernels return no gradient for reference_points. A caller whose reference points
    require grad must get the reference's op sequence (which differentiates them, ms_deform_attn.py:115-119) - and the same
    gradient as with fused=False - not a silent zero."""
    from vit_adapter_b200.modules import MSDeformAttn
    torch.manual_seed(0)
    m = MSDeformAttn(d_model=96, n_levels=3, n_heads=3, n_points=4).to(DEV)
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.02)
        m.attention_weights.weight.normal_(0, 0.5)
    shapes = torch.as_tensor([(8, 8), (4, 4), (2, 2)], dtype=torch.long, device=DEV)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    q = torch.randn(2, 10, 96, device=DEV)
    feat = torch.randn(2, 84, 96, device=DEV)
    res = {}
    for fused in (True, False):
        m.fused = fused
        ref = torch.rand(2, 10, 1, 2, device=DEV).mul_(0).add_(torch.linspace(0.1, 0.9, 10, device=DEV).view(1, 10, 1, 1)).requires_grad_()
        out = m(q, ref, feat, shapes, lsi)
        out.square().sum().backward()
        assert ref.grad is not None and float(ref.grad.abs().max()) > 0
        res[fused] = (out.detach(), ref.grad.clone())
    torch.testing.assert_close(res[True][0], res[False][0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(res[True][1], res[False][1], rtol=1e-5, atol=1e-6)
    # and with constant reference points the fused path is still the one that runs (no grad asked, none returned)
    m.fused = True
    ref = torch.rand(1, 10, 1, 2, device=DEV)
    n0 = _cabi.launch_count()
    m(q, ref, feat, shapes, lsi).sum().backward()
    assert _cabi.launch_count() - n0 >= 2 and ref.grad is None


def test_fused_entry_rejects_short_level_metadata():
    """the fused kernels index spatial_shapes / level_start_index with the module's level count; metadata with
    fewer rows must raise on the host instead of being read out of bounds on the device."""
    from vit_adapter_b200.modules import MSDeformAttn
    m = MSDeformAttn(d_model=96, n_levels=3, n_heads=3, n_points=4).to(DEV)
    shapes = torch.as_tensor([(8, 8), (4, 4)], dtype=torch.long, device=DEV)       # only two rows
    lsi = torch.as_tensor([0, 64], dtype=torch.long, device=DEV)
    q = torch.randn(1, 10, 96, device=DEV)
    feat = torch.randn(1, 80, 96, device=DEV)
    ref = torch.rand(1, 10, 1, 2, device=DEV)
    with pytest.raises(RuntimeError, match='spatial_shapes must be'):
        m(q, ref, feat, shapes, lsi)


def test_merged_query_weights_are_cached_until_a_parameter_changes():
    from vit_adapter_b200.modules import MSDeformAttn
    torch.manual_seed(1)
    m = MSDeformAttn(d_model=96, n_levels=1, n_heads=3, n_points=4).to(DEV)
    shapes = torch.as_tensor([(6, 6)], dtype=torch.long, device=DEV)
    lsi = shapes.new_zeros((1,))
    q = torch.randn(2, 12, 96, device=DEV)
    feat = torch.randn(2, 36, 96, device=DEV)
    ref = torch.rand(1, 12, 1, 2, device=DEV)
    out1 = m(q, ref, feat, shapes, lsi)
    key1, w1, _ = m._merged_cache
    out2 = m(q, ref, feat, shapes, lsi)
    assert m._merged_cache[0] == key1 and m._merged_cache[1] is w1 and torch.equal(out1, out2)
    out2.sum().backward()
    assert m.sampling_offsets.weight.grad is not None and m.attention_weights.weight.grad is not None
    assert m.sampling_offsets.bias.grad is not None and m.attention_weights.bias.grad is not None
    # gradients equal those of the un-merged path
    g_merged = [p.grad.clone() for p in (m.sampling_offsets.weight, m.attention_weights.weight, m.sampling_offsets.bias, m.attention_weights.bias)]
    m.zero_grad()
    m.merge_query_linears = False
    m(q, ref, feat, shapes, lsi).sum().backward()
    for a, p in zip(g_merged, (m.sampling_offsets.weight, m.attention_weights.weight, m.sampling_offsets.bias, m.attention_weights.bias)):
        torch.testing.assert_close(a, p.grad, rtol=1e-4, atol=1e-5 * float(p.grad.abs().max()) + 1e-12)
    m.merge_query_linears = True
    with torch.no_grad():
        m.attention_weights.bias.add_(torch.randn_like(m.attention_weights.bias))      # an optimizer step
    out3 = m(q, ref, feat, shapes, lsi)
    assert m._merged_cache[0] != key1 and m._merged_cache[1] is w1                     # refreshed in place, same buffer
    assert not torch.equal(out3, out1)
    m.merge_query_linears = False
    torch.testing.assert_close(out3, m(q, ref, feat, shapes, lsi), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('merge', [True, False], ids=['merged-gemm', 'two-linears'])
@pytest.mark.parametrize('name', ['module_l3_grads', 'module_l1_grads'])
def test_fused_module_gradients_match_reference_goldens(name, merge):
    """the fused / merged paths were only checked against this repo's own unfused path. Here the module's
    default (fused) path is checked DIRECTLY against outputs and fp64 autograd gradients of the real reference module
    (tests/golden/make_golden.py::module_grad_case): every parameter, query and feat. fp32 kernels vs fp64 reference:
    forward 1e-5 relative, gradients 1e-4 relative (the north star's fp32 tolerances)."""
    from vit_adapter_b200.modules import MSDeformAttn
    g = load_golden(name)
    d_model, L, M, P = [int(x) for x in g['cfg']]
    m = MSDeformAttn(d_model, L, M, P, float(g['ratio']))
    m.load_state_dict({k[3:]: v.float() for k, v in g.items() if k.startswith('sd.')}, strict=True)
    m = m.to(DEV)
    m.merge_query_linears = merge
    shapes = g['shapes'].to(DEV)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    q = g['query'].float().to(DEV).requires_grad_()
    feat = g['feat'].float().to(DEV).requires_grad_()
    assert _cabi.fused_supported(torch.empty(1, 1, M, int(float(g['ratio']) * d_model) // M, device=DEV), L, P)
    n0 = _cabi.launch_count()
    out = m(q, g['ref_pts'].float().to(DEV), feat, shapes, lsi)
    out.backward(g['grad_out'].float().to(DEV))
    torch.cuda.synchronize()
    assert _cabi.launch_count() - n0 >= 2          # forward + backward kernels of this library ran
    torch.testing.assert_close(out.detach().cpu().double(), g['out'], rtol=1e-5, atol=1e-5 * _scale(g['out']))
    torch.testing.assert_close(q.grad.cpu().double(), g['grad_query'], rtol=1e-4, atol=1e-4 * _scale(g['grad_query']))
    torch.testing.assert_close(feat.grad.cpu().double(), g['grad_feat'], rtol=1e-4, atol=1e-4 * _scale(g['grad_feat']))
    for k, p_ in m.named_parameters():
        want = g['grad.' + k]
        torch.testing.assert_close(p_.grad.cpu().double(), want, rtol=1e-4, atol=1e-4 * _scale(want), msg=lambda s, k=k: k + ': ' + s)
