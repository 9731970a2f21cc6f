"""Bias gradient of the adapter's Linears through csrc/adapter_colsum.cu: the column-sum kernel against
an fp64 sum, and `functions.linear` against nn.Linear's own forward / backward."""
import pytest
import torch
import torch.nn.functional as F

from vit_adapter_b200.functions import linear


def test_linear_on_cpu_is_f_linear():
    x = torch.randn(3, 5, 8, requires_grad=True)
    w = torch.randn(4, 8, requires_grad=True)
    b = torch.randn(4, requires_grad=True)
    y = linear(x, w, b)
    torch.testing.assert_close(y, F.linear(x, w, b), rtol=0, atol=0)
    y.sum().backward()
    assert b.grad is not None


@pytest.mark.gpu
@pytest.mark.parametrize('rows,C', [(86016, 768), (16384, 432), (5000, 144), (777, 8), (3, 2048), (100000, 192), (1, 64)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16], ids=['f32', 'bf16', 'f16'])
def test_colsum_kernel(rows, C, dtype):
    from vit_adapter_b200 import _cabi
    if dtype == torch.float32 and C > 1024:
        pytest.skip('f32 path takes C <= 1024')
    gen = torch.Generator().manual_seed(rows + C)
    x = torch.randn(rows, C, generator=gen).to(dtype).cuda()
    assert _cabi.colsum_supported(x)
    n0 = _cabi.launch_count()
    got = _cabi.colsum(x)
    assert _cabi.launch_count() - n0 == 2
    want = x.double().sum(0)
    torch.testing.assert_close(got.double(), want, rtol=1e-5, atol=1e-5 * float(x.double().abs().sum(0).max()))
    assert torch.equal(got, _cabi.colsum(x))   # deterministic


@pytest.mark.gpu
def test_colsum_unsupported_shapes():
    from vit_adapter_b200 import _cabi
    assert not _cabi.colsum_supported(torch.zeros(4, 6, device='cuda'))                          # C % 4
    assert not _cabi.colsum_supported(torch.zeros(4, 12, device='cuda', dtype=torch.bfloat16))   # C % 8
    assert not _cabi.colsum_supported(torch.zeros(4, 8, device='cuda', dtype=torch.float64))
    assert not _cabi.colsum_supported(torch.zeros(4, 16, device='cuda')[:, :8])                  # not contiguous


@pytest.mark.gpu
@pytest.mark.parametrize('amp', [False, torch.bfloat16, torch.float16], ids=['f32', 'bf16-autocast', 'f16-autocast'])
def test_linear_matches_nn_linear(amp):
    from vit_adapter_b200 import _cabi
    torch.manual_seed(0)
    lin = torch.nn.Linear(96, 72).cuda()
    x = torch.randn(4, 321, 96, device='cuda')
    gy = torch.randn(4, 321, 72, device='cuda')
    res = []
    for ours in (True, False):
        lin.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_()
        n0 = _cabi.launch_count()
        with torch.autocast('cuda', dtype=amp or torch.bfloat16, enabled=bool(amp)):
            y = linear(xi, lin.weight, lin.bias) if ours else lin(xi)
        y.backward(gy.to(y.dtype))
        assert (_cabi.launch_count() - n0 == 2) == ours
        res.append((y.detach().float(), xi.grad, lin.weight.grad.clone(), lin.bias.grad.clone()))
        assert xi.grad.dtype == torch.float32 and lin.weight.grad.dtype == torch.float32 and lin.bias.grad.dtype == torch.float32
    (y0, gx0, gw0, gb0), (y1, gx1, gw1, gb1) = res
    tol = 2e-2 if amp else 1e-5   # torch sums the bias gradient in bf16 under autocast; the kernel accumulates in fp32
    for a, b in ((y0, y1), (gx0, gx1), (gw0, gw1)):   # the same GEMMs (cuBLAS may pick another algorithm for a transposed view)
        torch.testing.assert_close(a, b, rtol=tol, atol=tol * float(b.abs().max()))
    torch.testing.assert_close(gb0, gb1, rtol=tol, atol=tol * float(gb1.abs().max()))
    want = gy.to(amp or torch.float32).double().sum((0, 1))
    torch.testing.assert_close(gb0.double(), want, rtol=1e-5, atol=1e-5 * float(want.abs().max()))


@pytest.mark.gpu
def test_no_grad_and_frozen_bias_use_f_linear():
    from vit_adapter_b200 import _cabi
    lin = torch.nn.Linear(16, 8).cuda()
    x = torch.randn(5, 16, device='cuda')
    n0 = _cabi.launch_count()
    with torch.no_grad():
        torch.testing.assert_close(linear(x, lin.weight, lin.bias), lin(x), rtol=0, atol=0)
    lin.bias.requires_grad_(False)
    linear(x.requires_grad_(), lin.weight, lin.bias).sum().backward()
    assert _cabi.launch_count() == n0
