"""Adapter ConvFFN depth-wise 3x3 on the token layout: oracle pinned on the real reference class
(CPU), CUDA kernel vs golden / oracle (GPU)."""
import pytest
import torch

from conftest import load_golden
from oracle import dwconv_ref

GOLD = ['dwconv_tokens', 'dwconv_tokens_c6']


@pytest.mark.parametrize('name', GOLD)
def test_oracle_matches_reference_dwconv(name):
    g = load_golden(name)
    C, H, W, B = [int(v) for v in g['cfg']]
    y = dwconv_ref.dwconv_tokens(g['x'], g['weight'], g['bias'], H, W)
    torch.testing.assert_close(y, g['y'], rtol=1e-12, atol=1e-13)
    gx, gw, gb = dwconv_ref.dwconv_tokens_backward(g['x'], g['weight'], g['bias'], H, W, g['grad_y'])
    torch.testing.assert_close(gx, g['grad_x'], rtol=1e-12, atol=1e-13)
    torch.testing.assert_close(gw, g['grad_weight'], rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(gb, g['grad_bias'], rtol=1e-12, atol=1e-12)


def test_module_falls_back_to_reference_sequence_on_cpu():
    from vit_adapter_b200.adapter import DWConv
    g = load_golden('dwconv_tokens')
    C, H, W, B = [int(v) for v in g['cfg']]
    m = DWConv(C).double()
    m.load_state_dict({'dwconv.weight': g['weight'], 'dwconv.bias': g['bias']})
    torch.testing.assert_close(m(g['x'], H, W), g['y'], rtol=1e-12, atol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize('name', GOLD)
@pytest.mark.parametrize('dtype', [torch.float64, torch.float32, torch.bfloat16, torch.float16], ids=['f64', 'f32', 'bf16', 'f16'])
def test_kernel_matches_reference_golden(name, dtype):
    from vit_adapter_b200 import _cabi
    from vit_adapter_b200.adapter import DWConv
    g = load_golden(name)
    C, H, W, B = [int(v) for v in g['cfg']]
    m = DWConv(C).to(dtype)
    m.load_state_dict({'dwconv.weight': g['weight'].to(dtype), 'dwconv.bias': g['bias'].to(dtype)})
    m = m.cuda()
    x = g['x'].to(dtype).cuda().requires_grad_()
    n0 = _cabi.launch_count()
    y = m(x, H, W)
    y.backward(g['grad_y'].to(dtype).cuda())
    # forward, backward-input, backward-weight (+ its partial-row sum on the run path): our kernels, not conv2d
    assert _cabi.launch_count() - n0 == (4 if C % 4 == 0 and dtype != torch.float64 else 3)
    tol = {torch.float64: 1e-11, torch.float32: 1e-5, torch.bfloat16: 2e-2, torch.float16: 3e-3}[dtype]
    sc = lambda t: float(t.abs().max())
    torch.testing.assert_close(y.detach().cpu().double(), g['y'], rtol=tol, atol=tol * sc(g['y']))
    torch.testing.assert_close(x.grad.cpu().double(), g['grad_x'], rtol=tol, atol=tol * sc(g['grad_x']))
    torch.testing.assert_close(m.dwconv.weight.grad.cpu().double(), g['grad_weight'], rtol=max(tol, 1e-4 if dtype != torch.float64 else tol),
                               atol=max(tol, 1e-4 if dtype != torch.float64 else tol) * sc(g['grad_weight']))
    torch.testing.assert_close(m.dwconv.bias.grad.cpu().double(), g['grad_bias'], rtol=max(tol, 1e-4 if dtype != torch.float64 else tol),
                               atol=max(tol, 1e-4 if dtype != torch.float64 else tol) * sc(g['grad_bias']))


@pytest.mark.gpu
@pytest.mark.parametrize('cfg', [(192, 32, 32, 2), (96, 8, 12, 3), (48, 56, 56, 1), (40, 6, 14, 2), (1024, 2, 2, 1),
                                 # shared-memory tile kernels (map widths multiples of 4, C % 32 == 0): 64- and 32-channel blocks,
                                 # bands that do not divide the map height, the ViT-Adapter-L shape, a non-square grid
                                 (96, 16, 8, 2), (256, 56, 56, 1), (64, 8, 16, 1), (192, 24, 40, 1)],
                         ids=['B-512', 'S-small', 'L-896-C48', 'ragged-runs', 'C1024', 'tile-C96', 'tile-L-896', 'tile-C64-rect', 'tile-odd-bands'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16], ids=['f32', 'bf16', 'f16'])
def test_kernel_vs_oracle_adapter_shapes(cfg, dtype):
    from vit_adapter_b200.adapter import DWConv
    C, H, W, B = cfg
    g = torch.Generator().manual_seed(7)
    n = (H // 2) * (W // 2)
    x = torch.randn(B, 21 * n, C, generator=g)
    gy = torch.randn(B, 21 * n, C, generator=g)
    m = DWConv(C)
    xq, gyq = x.to(dtype).float(), gy.to(dtype).float()
    wq, bq = m.dwconv.weight.detach().to(dtype).float(), m.dwconv.bias.detach().to(dtype).float()
    want = dwconv_ref.dwconv_tokens(xq, wq, bq, H, W)
    wgx, wgw, wgb = dwconv_ref.dwconv_tokens_backward(xq, wq, bq, H, W, gyq)
    md = m.cuda()
    xc = x.to(dtype).cuda().requires_grad_()
    with torch.autocast('cuda', dtype=dtype, enabled=(dtype != torch.float32)):
        y = md(xc, H, W)
    y.backward(gy.to(dtype).cuda())
    tol = {torch.float32: 1e-5, torch.bfloat16: 2e-2, torch.float16: 3e-3}[dtype]
    sc = lambda t: float(t.abs().max())
    torch.testing.assert_close(y.detach().float().cpu(), want, rtol=tol, atol=tol * sc(want))
    torch.testing.assert_close(xc.grad.float().cpu(), wgx, rtol=tol, atol=tol * sc(wgx))
    gtol = {torch.float32: 1e-4, torch.bfloat16: 2e-2, torch.float16: 3e-3}[dtype]
    torch.testing.assert_close(md.dwconv.weight.grad.float().cpu(), wgw, rtol=gtol, atol=gtol * sc(wgw))
    torch.testing.assert_close(md.dwconv.bias.grad.float().cpu(), wgb, rtol=gtol, atol=gtol * sc(wgb))
