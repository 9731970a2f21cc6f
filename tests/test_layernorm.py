"""Adapter LayerNorm prologues: oracle pinned on the LayerNorm the reference's Injector builds (CPU),
CUDA row kernel vs golden / oracle (GPU), and the Injector / Extractor wiring."""
import pytest
import torch
from torch import nn

from conftest import load_golden
from oracle import layernorm_ref

GOLD = ['layernorm_c96', 'layernorm_c768']


@pytest.mark.parametrize('name', GOLD)
def test_oracle_matches_reference_layernorm(name):
    g = load_golden(name)
    eps = float(g['eps'][0])
    assert eps == 1e-6
    torch.testing.assert_close(layernorm_ref.layernorm(g['x'], g['weight'], g['bias'], eps), g['y'], rtol=1e-12, atol=1e-12)
    gx, gw, gb = layernorm_ref.layernorm_backward(g['x'], g['weight'], eps, g['grad_y'])
    torch.testing.assert_close(gx, g['grad_x'], rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(gw, g['grad_weight'], rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(gb, g['grad_bias'], rtol=1e-11, atol=1e-11)


def test_apply_norm_on_cpu_is_the_module():
    from vit_adapter_b200.adapter.adapter_modules import apply_norm
    g = load_golden('layernorm_c96')
    m = nn.LayerNorm(96, eps=1e-6).double()
    m.load_state_dict({'weight': g['weight'], 'bias': g['bias']})
    torch.testing.assert_close(apply_norm(m, g['x']), g['y'], rtol=1e-12, atol=1e-12)


def _module(g, dtype=torch.float32):
    C = g['weight'].numel()
    m = nn.LayerNorm(C, eps=float(g['eps'][0]))
    m.load_state_dict({'weight': g['weight'].float(), 'bias': g['bias'].float()})
    return m.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize('name', GOLD)
@pytest.mark.parametrize('mode', ['f32', 'f32->bf16 (autocast)', 'bf16', 'f32->f16 (autocast)', 'f16'])
def test_kernel_matches_reference_golden(name, mode):
    from vit_adapter_b200 import _cabi
    from vit_adapter_b200.adapter.adapter_modules import apply_norm
    g = load_golden(name)
    m = _module(g)
    low = torch.float16 if 'f16' in mode and 'bf16' not in mode else torch.bfloat16
    xdt = low if mode in ('bf16', 'f16') else torch.float32
    xq = g['x'].to(xdt)
    x = xq.cuda().requires_grad_()
    n0 = _cabi.launch_count()
    with torch.autocast('cuda', dtype=low, enabled=(mode != 'f32')):
        y = apply_norm(m, x)
    assert y.dtype == (torch.float32 if mode == 'f32' else low)
    gyq = g['grad_y'].to(y.dtype)
    y.backward(gyq.cuda())
    assert _cabi.launch_count() - n0 == 3   # row forward, row backward, parameter-gradient sum: no torch LayerNorm
    # expected values from the oracle on the SAME (rounded) inputs, in fp64
    eps = float(g['eps'][0])
    want = layernorm_ref.layernorm(xq.double(), g['weight'].float().double(), g['bias'].float().double(), eps)
    wgx, wgw, wgb = layernorm_ref.layernorm_backward(xq.double(), g['weight'].float().double(), eps, gyq.double())
    otol = 1e-5 if mode == 'f32' else (1e-2 if low == torch.bfloat16 else 2e-3)
    itol = 1e-2 if mode == 'bf16' else (2e-3 if mode == 'f16' else 1e-5)
    sc = lambda t: float(t.abs().max())
    torch.testing.assert_close(y.detach().cpu().double(), want, rtol=otol, atol=otol * sc(want))
    torch.testing.assert_close(x.grad.cpu().double(), wgx, rtol=itol, atol=itol * sc(wgx))
    torch.testing.assert_close(m.weight.grad.cpu().double(), wgw, rtol=1e-5, atol=1e-5 * sc(wgw))
    torch.testing.assert_close(m.bias.grad.cpu().double(), wgb, rtol=1e-5, atol=1e-5 * sc(wgb))
    if mode == 'f32':   # and against the golden itself (inputs not rounded)
        torch.testing.assert_close(y.detach().cpu().double(), g['y'], rtol=1e-5, atol=1e-5 * sc(g['y']))
        torch.testing.assert_close(x.grad.cpu().double(), g['grad_x'], rtol=1e-5, atol=1e-5 * sc(g['grad_x']))


@pytest.mark.gpu
@pytest.mark.parametrize('C,rows', [(384, 777), (768, 5376 * 2 + 3), (1024, 300), (4, 9), (132, 64), (1020, 17)])
def test_kernel_vs_oracle_channel_counts(C, rows):
    """every quads-per-lane instantiation incl. ragged last quads; more rows than resident warps; tiny C"""
    from vit_adapter_b200.adapter.adapter_modules import apply_norm
    gen = torch.Generator().manual_seed(C)
    x = 3.0 * torch.randn(rows, C, generator=gen) + 1.5
    gy = torch.randn(rows, C, generator=gen)
    m = nn.LayerNorm(C, eps=1e-6)
    with torch.no_grad():
        m.weight.copy_(1 + 0.3 * torch.randn(C, generator=gen))
        m.bias.copy_(0.2 * torch.randn(C, generator=gen))
    want = layernorm_ref.layernorm(x.double(), m.weight.detach().double(), m.bias.detach().double(), 1e-6)
    wgx, wgw, wgb = layernorm_ref.layernorm_backward(x.double(), m.weight.detach().double(), 1e-6, gy.double())
    md = m.cuda()
    xc = x.cuda().requires_grad_()
    y = apply_norm(md, xc)
    y.backward(gy.cuda())
    sc = lambda t: float(t.abs().max())
    torch.testing.assert_close(y.detach().cpu().double(), want, rtol=1e-5, atol=1e-5 * sc(want))
    torch.testing.assert_close(xc.grad.cpu().double(), wgx, rtol=1e-5, atol=1e-5 * sc(wgx))
    torch.testing.assert_close(md.weight.grad.cpu().double(), wgw, rtol=2e-5, atol=2e-5 * sc(wgw))
    torch.testing.assert_close(md.bias.grad.cpu().double(), wgb, rtol=2e-5, atol=2e-5 * sc(wgb))
    # deterministic parameter gradients (two-stage reduction, no atomics)
    first = md.weight.grad.clone()
    md.zero_grad(set_to_none=True)
    xc.grad = None
    apply_norm(md, xc).backward(gy.cuda())
    assert torch.equal(first, md.weight.grad)


@pytest.mark.gpu
def test_unsupported_shapes_fall_back_to_torch():
    from vit_adapter_b200 import _cabi
    from vit_adapter_b200.adapter.adapter_modules import apply_norm
    for C, dt in ((30, torch.float32), (2048, torch.float32), (64, torch.float64)):
        m = nn.LayerNorm(C, eps=1e-6).to(dt).cuda()
        x = torch.randn(5, C, dtype=dt, device='cuda')
        n0 = _cabi.launch_count()
        torch.testing.assert_close(apply_norm(m, x), m(x))
        assert _cabi.launch_count() == n0


@pytest.mark.gpu
def test_injector_extractor_with_and_without_row_kernel_agree():
    from vit_adapter_b200.adapter import Extractor, Injector, deform_inputs
    torch.manual_seed(0)
    dev = torch.device('cuda')
    dim, heads, side = 64, 4, 64
    di1, di2 = deform_inputs(torch.zeros(2, 3, side, side, device=dev))
    h = side // 16
    x = torch.randn(2, h * h, dim, device=dev)
    c = torch.randn(2, 21 * (h // 2) ** 2, dim, device=dev)
    for mod, args in ((Injector(dim, heads, 4, 3, 1.0, init_values=0.5).to(dev), (x, di1[0], c, di1[1], di1[2])),
                      (Extractor(dim, heads, 4, 1, 1.0).to(dev), (c, di2[0], x, di2[1], di2[2], h, h))):
        with torch.no_grad():
            for p in mod.parameters():
                p.add_(0.05 * torch.randn_like(p))
        outs, grads = [], []
        for fused in (True, False):
            mod.fused_norm = fused
            mod.zero_grad(set_to_none=True)
            q = args[0].clone().requires_grad_()
            out = mod(q, *args[1:])
            out.square().sum().backward()
            outs.append(out.detach())
            grads.append([q.grad] + [p.grad.clone() for p in mod.parameters()])
        torch.testing.assert_close(outs[0], outs[1], rtol=1e-4, atol=1e-5)
        for a, b in zip(*grads):
            torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-4 * float(b.abs().max()) + 1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize('amp', [False, True], ids=['f32', 'bf16-autocast'])
def test_residual_gradient_is_folded_into_layernorm_backward(amp):
    """x + f(LN(x)): the gradient over the residual connection enters the LayerNorm backward kernel (no separate add)."""
    from vit_adapter_b200 import _cabi
    from vit_adapter_b200.adapter.adapter_modules import apply_norm_residual
    g = load_golden('layernorm_c96')
    m = _module(g)
    x = g['x'].float().cuda().requires_grad_()
    a = torch.randn(g['x'].shape, generator=torch.Generator().manual_seed(9))
    n0 = _cabi.launch_count()
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
        y, xr = apply_norm_residual(m, x)
    assert torch.equal(xr, x) and xr.requires_grad
    gy = g['grad_y'].to(y.dtype)
    ((y.float() * gy.float().cuda()).sum() + (xr * a.cuda()).sum()).backward()
    assert _cabi.launch_count() - n0 == 3
    wgx, _, _ = layernorm_ref.layernorm_backward(g['x'].float().double(), g['weight'].float().double(), float(g['eps'][0]), gy.double())
    want = wgx + a.double()
    torch.testing.assert_close(x.grad.cpu().double(), want, rtol=1e-5, atol=1e-5 * float(want.abs().max()))
    # only the residual output used -> gradient passes straight through
    x2 = g['x'].float().cuda().requires_grad_()
    _, xr2 = apply_norm_residual(m, x2)
    (xr2 * a.cuda()).sum().backward()
    torch.testing.assert_close(x2.grad.cpu(), a, rtol=0, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize('bdt', [torch.bfloat16, torch.float16, torch.float32], ids=['bf16-branch', 'f16-branch', 'f32-branch'])
def test_residual_add_kernel_is_bit_identical_to_torch(bdt):
    from vit_adapter_b200 import _cabi
    gen = torch.Generator().manual_seed(4)
    res = torch.randn(3, 1001, 8 * 13, generator=gen).cuda()
    br = torch.randn(3, 1001, 8 * 13, generator=gen).to(bdt).cuda()
    assert _cabi.residual_add_supported(res, br)
    n0 = _cabi.launch_count()
    out = _cabi.residual_add(res, br)
    assert _cabi.launch_count() - n0 == 1
    assert torch.equal(out, res + br) and out.dtype == torch.float32
    assert not _cabi.residual_add_supported(res[..., :7].contiguous(), br[..., :7].contiguous())   # 3*1001*7 not a multiple of 8
    from vit_adapter_b200.adapter.adapter_modules import residual_add
    r = res.clone().requires_grad_()
    b = br.clone().requires_grad_()
    residual_add(r, b).square().sum().backward()
    r2 = res.clone().requires_grad_()
    b2 = br.clone().requires_grad_()
    (r2 + b2).square().sum().backward()
    assert torch.equal(r.grad, r2.grad) and torch.equal(b.grad, b2.grad) and b.grad.dtype == bdt


@pytest.mark.gpu
@pytest.mark.parametrize('low', [torch.bfloat16, torch.float16], ids=['bf16', 'f16'])
def test_extractor_autocast_with_and_without_adapter_kernels(low):
    """bf16 / fp16 autocast: row-kernel LayerNorm + folded residual gradient + residual-add kernel + DWConv kernel +
    column-sum bias gradients vs torch's op sequence (under fp16 the sampling core runs in fp32, as in the reference)."""
    from vit_adapter_b200.adapter import Extractor, deform_inputs
    torch.manual_seed(1)
    dev = torch.device('cuda')
    dim, heads, side = 64, 4, 64
    _, di2 = deform_inputs(torch.zeros(2, 3, side, side, device=dev))
    h = side // 16
    x = torch.randn(2, h * h, dim, device=dev)
    c = torch.randn(2, 21 * (h // 2) ** 2, dim, device=dev)
    mod = Extractor(dim, heads, 4, 1, 1.0).to(dev)
    res = []
    for fused in (True, False):
        for m in mod.modules():
            for flag in ('fused_norm', 'token_kernel', 'colsum_bias_grad'):
                if hasattr(m, flag):
                    setattr(m, flag, fused)
        mod.zero_grad(set_to_none=True)
        q = c.clone().requires_grad_()
        with torch.autocast('cuda', dtype=low):
            out = mod(q, di2[0], x, di2[1], di2[2], h, h)
        assert out.dtype == torch.float32
        out.square().sum().backward()
        res.append((out.detach(), q.grad, mod.query_norm.weight.grad.clone(), mod.ffn.fc2.bias.grad.clone()))
    for a, b in zip(*res):
        torch.testing.assert_close(a, b, rtol=3e-2, atol=3e-2 * float(b.abs().max()))
