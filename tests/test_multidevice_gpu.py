"""Tensors on a device that is not the current one: every binding switches to the tensors' device for allocation and launch
(the reference relies on ATen's device guard for the same thing). Needs 2 GPUs; skipped otherwise."""
import pytest
import torch
from torch import nn

from conftest import make_inputs

import vit_adapter_b200 as vab

pytestmark = pytest.mark.gpu


def _need_two():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')


def test_op_on_non_current_device():
    _need_two()
    inp = make_inputs(2, 3, 32, 50, [(8, 8), (4, 4), (2, 2)], 4, seed=5, dist='adapter')
    res = []
    for d in (0, 1):
        dev = torch.device('cuda', d)
        assert torch.cuda.current_device() == 0
        v = inp['value'].to(dev).requires_grad_()
        loc = inp['loc'].to(dev).requires_grad_()
        aw = inp['aw'].to(dev).requires_grad_()
        out = vab.MSDeformAttnFunction.apply(v, inp['shapes'].to(dev), inp['lsi'].to(dev), loc, aw, 64)
        out.backward(inp['grad_out'].to(dev))
        assert out.device == dev and v.grad.device == dev
        res.append((out.detach().cpu(), v.grad.cpu(), loc.grad.cpu(), aw.grad.cpu()))
    torch.testing.assert_close(res[0][0], res[1][0], rtol=0, atol=0)
    for a, b in zip(res[0][1:], res[1][1:]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)   # atomics order
    with pytest.raises(RuntimeError, match='expected'):
        vab.MSDeformAttnFunction.apply(inp['value'].to('cuda:0'), inp['shapes'].to('cuda:1'), inp['lsi'].to('cuda:0'),
                                       inp['loc'].to('cuda:0'), inp['aw'].to('cuda:0'), 64)


def test_adapter_kernels_on_non_current_device():
    _need_two()
    from vit_adapter_b200.adapter import Extractor, deform_inputs
    torch.manual_seed(0)
    outs = []
    for d in (0, 1):
        dev = torch.device('cuda', d)
        torch.manual_seed(3)
        mod = Extractor(64, 4, 4, 1, 1.0).to(dev)
        _, di2 = deform_inputs(torch.zeros(2, 3, 64, 64, device=dev))
        x = torch.randn(2, 16, 64, generator=torch.Generator().manual_seed(1)).to(dev)
        c = torch.randn(2, 84, 64, generator=torch.Generator().manual_seed(2)).to(dev).requires_grad_()
        with torch.autocast('cuda', dtype=torch.bfloat16):
            out = mod(c, di2[0], x, di2[1], di2[2], 4, 4)
        out.square().sum().backward()
        assert out.device == dev and c.grad.device == dev
        outs.append((out.detach().cpu(), c.grad.cpu()))
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=2e-2, atol=2e-2 * float(outs[0][1].abs().max()))
