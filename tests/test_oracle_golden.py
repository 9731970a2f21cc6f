"""Pin the oracles against golden vectors produced by the REAL reference (tests/golden/make_golden.py).

CPU-only. The reference publishes no stored vectors; these fixtures are outputs of its
own ms_deform_attn_core_pytorch (+ autograd) on its own test recipe (detection/ops/test.py:16-37) and on
adapter-shaped miniatures.
"""
import pytest
import torch

from conftest import OP_CASES, load_golden
from oracle import c_oracle, core_pytorch


@pytest.mark.parametrize('case', OP_CASES)
def test_c_oracle_forward_f64(case):
    g = load_golden(case)
    out = c_oracle.forward(g['value'].double(), g['shapes'], g['lsi'], g['loc'].double(), g['aw'].double())
    # fp64: the reference's own tolerance for CUDA-vs-core is allclose defaults (ops/test.py:43)
    torch.testing.assert_close(out, g['out_f64'], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('case', OP_CASES)
def test_c_oracle_forward_f32(case):
    g = load_golden(case)
    out = c_oracle.forward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'])
    # north-star tolerance for the fp32 forward: 1e-5 relative, 1e-6 absolute
    torch.testing.assert_close(out.double(), g['out_f64'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out, g['out_f32'], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('case', OP_CASES)
def test_c_oracle_backward_f64(case):
    g = load_golden(case)
    gv, gl, ga = c_oracle.backward(g['value'].double(), g['shapes'], g['lsi'], g['loc'].double(),
                                   g['aw'].double(), g['grad_out'].double())
    torch.testing.assert_close(gv, g['grad_value_f64'], rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(ga, g['grad_aw_f64'], rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(gl, g['grad_loc_f64'], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize('case', OP_CASES)
def test_core_pytorch_restatement(case):
    g = load_golden(case)
    out, (gv, gl, ga) = core_pytorch.forward_backward(g['value'].double(), g['shapes'], g['loc'].double(),
                                                       g['aw'].double(), g['grad_out'].double())
    torch.testing.assert_close(out, g['out_f64'], rtol=1e-12, atol=1e-14)
    torch.testing.assert_close(gv, g['grad_value_f64'], rtol=1e-12, atol=1e-14)
    torch.testing.assert_close(gl, g['grad_loc_f64'], rtol=1e-12, atol=1e-14)
    torch.testing.assert_close(ga, g['grad_aw_f64'], rtol=1e-12, atol=1e-14)
    out32 = core_pytorch.ms_deform_attn_core(g['value'], g['shapes'], g['loc'], g['aw'])
    torch.testing.assert_close(out32, g['out_f32'], rtol=0, atol=0)


def test_kat_spot_values():
    """spot values of the seed-3 fixture and the reference's fp64 output."""
    g = load_golden('op_kat_seed3')
    assert abs(float(g['value'][0, 0, 0, 0]) - 4.263520168e-05) < 1e-12
    assert abs(float(g['loc'][0, 0, 0, 0, 0, 0]) - 0.6416304111) < 1e-7
    assert abs(float(g['aw'][0, 0, 0, 0, 1]) - 0.3854852319) < 1e-7
    want = torch.tensor([0.0018993784157779181, 0.004602827532968805, 0.004671175247309776, 0.004384399819001662,
                         0.0037950971737622935, 0.002512764199421532, 0.0018444261512603004, 0.003634679248037905],
                        dtype=torch.float64)
    torch.testing.assert_close(g['out_f64'].flatten(), want, rtol=1e-12, atol=0)
    out = c_oracle.forward(g['value'].double(), g['shapes'], g['lsi'], g['loc'].double(), g['aw'].double())
    torch.testing.assert_close(out.flatten(), want, rtol=1e-12, atol=0)


def test_c_oracle_threads_deterministic():
    g = load_golden('op_inj_edges')
    args = (g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'])
    c_oracle.set_threads(1)
    a = c_oracle.forward(*args)
    ga = c_oracle.backward(*args, g['grad_out'])
    c_oracle.set_threads(4)
    b = c_oracle.forward(*args)
    gb = c_oracle.backward(*args, g['grad_out'])
    c_oracle.set_threads(1)
    assert torch.equal(a, b)
    for x, y in zip(ga, gb):
        assert torch.equal(x, y)


def test_point_index_matches_forward_support():
    """The index dump agrees with what the forward actually reads: perturbing a masked-out corner's token
    does not change the output, perturbing a read corner does."""
    g = load_golden('op_d32_edges')
    N, S, M, D = g['value'].shape
    idx = c_oracle.point_index(g['shapes'], g['lsi'], g['loc'], M, D)
    assert idx.shape[0] == g['aw'].numel()
    assert (idx[:, 2] >= 0).all() and (idx[:, 2] <= 15).all()
    skipped = idx[:, 2] == 0
    assert skipped.any() and (~skipped).any()
    # every in-range point has h_low in [-1, H-1]
    L, P = g['shapes'].shape[0], g['loc'].shape[4]
    lvl = (torch.arange(idx.shape[0]) // P) % L
    H = g['shapes'][lvl, 0].int()
    W = g['shapes'][lvl, 1].int()
    ok = ~skipped
    assert (idx[ok, 0] >= -1).all() and (idx[ok, 0] <= H[ok] - 1).all()
    assert (idx[ok, 1] >= -1).all() and (idx[ok, 1] <= W[ok] - 1).all()
