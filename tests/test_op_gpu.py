"""GPU parity tests of the CUDA path (called through the C-ABI via MSDeformAttnFunction / _cabi)
against the oracles. Tolerances are the north star's: indices bit-exact; fp32 forward 1e-5 rel / 1e-6
abs; fp32 gradients 1e-4 rel (atomic order); bf16 1e-2."""
import pytest
import torch

from conftest import OP_CASES, load_golden, make_inputs
from oracle import c_oracle, refcuda

import vit_adapter_b200 as vab
from vit_adapter_b200 import _cabi

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _cuda(d, dtype=None):
    out = {}
    for k, v in d.items():
        if v.dtype in (torch.int64, torch.bool) or dtype is None or not v.is_floating_point():
            out[k] = v.to(DEV)
        else:
            out[k] = v.to(DEV, dtype)
    return out


def _run(inp, dtype):
    """forward + backward through the public autograd Function on the GPU."""
    g = _cuda(inp)
    value = g['value'].to(dtype).requires_grad_()
    cdt = torch.float64 if dtype == torch.float64 else torch.float32
    loc = g['loc'].to(cdt).requires_grad_()
    aw = g['aw'].to(cdt).requires_grad_()
    if dtype == torch.float16:
        vab.set_amp_value_dtype(torch.float16)   # native fp16 I/O is opt-in (default: fp32 up-cast, as the reference)
    try:
        out = vab.MSDeformAttnFunction.apply(value, g['shapes'], g['lsi'], loc, aw, 64)
        out.backward(g['grad_out'].to(dtype))
        torch.cuda.synchronize()
    finally:
        vab.set_amp_value_dtype(torch.float32)
    return out.detach().cpu(), value.grad.cpu(), loc.grad.cpu(), aw.grad.cpu()


def _scale(t):
    return float(t.abs().max()) + 1e-30


# ---------------------------------------------------------------------------------------------------
# golden fixtures (outputs of the real reference)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', OP_CASES)
def test_golden_f32(case):
    g = load_golden(case)
    out, gv, gl, ga = _run(g, torch.float32)
    torch.testing.assert_close(out.double(), g['out_f64'], rtol=1e-5, atol=1e-6)
    # gradients vs the fp32 C oracle (same fmaf coordinate arithmetic): 1e-4 relative
    ogv, ogl, oga = c_oracle.backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'])
    torch.testing.assert_close(gv, ogv, rtol=1e-4, atol=1e-4 * _scale(ogv))
    torch.testing.assert_close(gl, ogl, rtol=1e-4, atol=1e-4 * _scale(ogl))
    torch.testing.assert_close(ga, oga, rtol=1e-4, atol=1e-4 * _scale(oga))
    # and vs the reference's fp64 autograd gradients (looser: fp32 vs fp64 arithmetic)
    torch.testing.assert_close(gv.double(), g['grad_value_f64'], rtol=1e-4, atol=1e-4 * _scale(ogv))
    torch.testing.assert_close(ga.double(), g['grad_aw_f64'], rtol=1e-4, atol=1e-4 * _scale(oga))


@pytest.mark.parametrize('case', OP_CASES)
def test_golden_f64(case):
    """fp64 through the generic kernel: the reference's own fp64 tolerance (ops/test.py:43, allclose defaults)."""
    g = load_golden(case)
    out, gv, gl, ga = _run(g, torch.float64)
    torch.testing.assert_close(out, g['out_f64'], rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(gv, g['grad_value_f64'], rtol=1e-9, atol=1e-11)
    torch.testing.assert_close(gl, g['grad_loc_f64'], rtol=1e-8, atol=1e-9)
    torch.testing.assert_close(ga, g['grad_aw_f64'], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize('low', [torch.bfloat16, torch.float16], ids=['bf16', 'f16'])
@pytest.mark.parametrize('case', ['op_inj_edges', 'op_ext_edges', 'op_d32_edges', 'op_d64_edges', 'op_odd_d5'])
def test_golden_bf16(case, low):
    """16-bit I/O (bf16; fp16 = the same kernels on __half), fp32 locations / weights / accumulation."""
    g = load_golden(case)
    # oracle on the rounded value / grad_out, in fp32: isolates kernel error from input rounding
    vq = g['value'].to(low).float()
    goq = g['grad_out'].to(low).float()
    tol = 1e-2 if low == torch.bfloat16 else 2e-3
    want = c_oracle.forward(vq, g['shapes'], g['lsi'], g['loc'], g['aw'])
    wgv, wgl, wga = c_oracle.backward(vq, g['shapes'], g['lsi'], g['loc'], g['aw'], goq)
    g2 = dict(g)
    out, gv, gl, ga = _run(g2, low)
    assert out.dtype == low and gv.dtype == low
    torch.testing.assert_close(out.float(), want, rtol=tol, atol=tol * _scale(want))
    torch.testing.assert_close(gv.float(), wgv, rtol=tol, atol=tol * _scale(wgv))
    torch.testing.assert_close(gl, wgl, rtol=tol, atol=tol * _scale(wgl))
    torch.testing.assert_close(ga, wga, rtol=tol, atol=tol * _scale(wga))


# ---------------------------------------------------------------------------------------------------
# seeded synthetic shapes vs the C oracle (sizes the oracle finishes in seconds)
# ---------------------------------------------------------------------------------------------------
SHAPES = [
    # name, N, M, D, Lq, shapes, P, dist
    ('S-injector', 2, 6, 64, 256, [(32, 32), (16, 16), (8, 8)], 4, 'adapter'),
    ('B-injector', 2, 12, 32, 256, [(32, 32), (16, 16), (8, 8)], 4, 'adapter'),
    ('B-extractor', 2, 12, 32, 1344, [(16, 16)], 4, 'adapter'),
    ('L-injector', 1, 16, 32, 196, [(28, 28), (14, 14), (7, 7)], 4, 'edges'),
    ('L-d64', 1, 16, 64, 196, [(28, 28), (14, 14), (7, 7)], 4, 'uniform'),
    ('ragged-levels', 3, 5, 32, 37, [(7, 9), (3, 4), (1, 1), (2, 5)], 3, 'edges'),
    ('one-query', 1, 1, 32, 1, [(2, 2)], 1, 'edges'),
    ('d8', 2, 4, 8, 50, [(9, 9), (5, 5)], 4, 'edges'),
    ('d128', 1, 2, 128, 33, [(9, 9), (5, 5)], 4, 'uniform'),
    ('d16-p8', 2, 8, 16, 45, [(12, 10)], 8, 'edges'),
]


@pytest.mark.parametrize('cfg', SHAPES, ids=[s[0] for s in SHAPES])
def test_vs_c_oracle_f32(cfg):
    _, N, M, D, Lq, shapes, P, dist = cfg
    inp = make_inputs(N, M, D, Lq, shapes, P, seed=5, dist=dist)
    out, gv, gl, ga = _run(inp, torch.float32)
    want = c_oracle.forward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'])
    wgv, wgl, wga = c_oracle.backward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], inp['grad_out'])
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-6 * max(1.0, _scale(want)))
    torch.testing.assert_close(gv, wgv, rtol=1e-4, atol=1e-4 * _scale(wgv))
    torch.testing.assert_close(gl, wgl, rtol=1e-4, atol=1e-4 * _scale(wgl))
    torch.testing.assert_close(ga, wga, rtol=1e-4, atol=1e-4 * _scale(wga))


@pytest.mark.parametrize('cfg', SHAPES[:6], ids=[s[0] for s in SHAPES[:6]])
def test_vs_c_oracle_bf16(cfg):
    _, N, M, D, Lq, shapes, P, dist = cfg
    inp = make_inputs(N, M, D, Lq, shapes, P, seed=6, dist=dist)
    vq, goq = inp['value'].bfloat16().float(), inp['grad_out'].bfloat16().float()
    out, gv, gl, ga = _run(inp, torch.bfloat16)
    want = c_oracle.forward(vq, inp['shapes'], inp['lsi'], inp['loc'], inp['aw'])
    wgv, wgl, wga = c_oracle.backward(vq, inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], goq)
    torch.testing.assert_close(out.float(), want, rtol=1e-2, atol=1e-2 * _scale(want))
    torch.testing.assert_close(gv.float(), wgv, rtol=1e-2, atol=1e-2 * _scale(wgv))
    torch.testing.assert_close(gl, wgl, rtol=1e-2, atol=1e-2 * _scale(wgl))
    torch.testing.assert_close(ga, wga, rtol=1e-2, atol=1e-2 * _scale(wga))


@pytest.mark.parametrize('D', [30, 32, 64, 71, 1025, 2048, 3096])
def test_reference_channel_sweep(D):
    """The channel counts the reference's own test sweeps to hit every backward branch (ops/test.py:108)."""
    inp = make_inputs(1, 2, D, 2, [(6, 4), (3, 2)], 2, seed=3, dist='uniform')
    out, gv, gl, ga = _run(inp, torch.float32)
    want = c_oracle.forward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'])
    wgv, wgl, wga = c_oracle.backward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], inp['grad_out'])
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-6 * max(1.0, _scale(want)))
    torch.testing.assert_close(gv, wgv, rtol=1e-4, atol=1e-4 * _scale(wgv))
    torch.testing.assert_close(gl, wgl, rtol=1e-4, atol=1e-4 * _scale(wgl))
    torch.testing.assert_close(ga, wga, rtol=1e-4, atol=1e-4 * _scale(wga))


@pytest.mark.parametrize('D', [4, 30, 32])
def test_gradcheck_f64(D):
    """The reference's gradient test: torch.autograd.gradcheck in fp64 (ops/test.py:78-101)."""
    inp = make_inputs(1, 2, D, 2, [(6, 4), (3, 2)], 2, seed=3, dist='uniform', dtype=torch.float64)
    g = _cuda(inp)
    value = (g['value'] * 0.01).requires_grad_()
    loc = g['loc'].clone().requires_grad_()
    aw = g['aw'].clone().requires_grad_()
    assert torch.autograd.gradcheck(vab.MSDeformAttnFunction.apply, (value, g['shapes'], g['lsi'], loc, aw, 2))


# ---------------------------------------------------------------------------------------------------
# index / level-offset arithmetic: bit-exact
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dist', ['uniform', 'edges', 'adapter'])
def test_point_index_bit_exact(dist):
    inp = make_inputs(2, 12, 32, 1024, [(64, 64), (32, 32), (16, 16)], 4, seed=9, dist=dist)
    want = c_oracle.point_index(inp['shapes'], inp['lsi'], inp['loc'], 12, 32)
    got = _cabi.debug_point_index(inp['shapes'].to(DEV), inp['lsi'].to(DEV), inp['loc'].to(DEV), 12, 32).cpu()
    assert torch.equal(got, want)


def _fma_sensitive_locations(H):
    """fp32 locations within 3 ulp of a texel centre (k+0.5)/H where floor(fmaf(loc,H,-0.5)) != floor(loc*H-0.5)."""
    k = torch.arange(H, dtype=torch.float64)
    base = ((k + 0.5) / H).float()
    cands, up, dn = [base], base.clone(), base.clone()
    for _ in range(3):
        up = torch.nextafter(up, torch.tensor(2.0))
        dn = torch.nextafter(dn, torch.tensor(-1.0))
        cands += [up.clone(), dn.clone()]
    c = torch.cat(cands)
    fused = torch.floor((c.double() * H - 0.5).float())   # exact product, one rounding == fmaf
    unfused = torch.floor(c * H - 0.5)                    # two roundings
    sel = fused != unfused
    return c[sel], fused[sel]


@pytest.mark.parametrize('H', [100, 37])
def test_point_index_fma_sensitive_locations(H):
    """Locations where fmaf(loc,H,-0.5) and (loc*H)-0.5 floor differently: the kernels must take the fused
    result (what nvcc emits for the reference) — checked on the indices AND against the
    reference's own CUDA kernel through the values it gathers."""
    sel, fused_floor = _fma_sensitive_locations(H)
    n = sel.numel()
    assert n >= 3, 'the construction must yield sensitive locations (H not a power-of-two multiple)'
    loc = torch.stack([sel, sel], -1).view(1, n, 1, 1, 1, 2).contiguous()
    shapes = torch.as_tensor([(H, H)], dtype=torch.long)
    lsi = torch.zeros(1, dtype=torch.long)
    want = c_oracle.point_index(shapes, lsi, loc, 1, 4)
    got = _cabi.debug_point_index(shapes.to(DEV), lsi.to(DEV), loc.to(DEV), 1, 4).cpu()
    assert torch.equal(got, want)
    assert torch.equal(want[:, 0].float(), fused_floor) and torch.equal(want[:, 1].float(), fused_floor)
    if refcuda.available():
        # value[token] = token index: the forward output then reveals which rows each implementation read
        value = torch.arange(H * H, dtype=torch.float32).view(1, H * H, 1, 1).repeat(1, 1, 1, 4).contiguous().to(DEV)
        aw = torch.ones(1, n, 1, 1, 1, device=DEV)
        ours = _cabi.forward(value, shapes.to(DEV), lsi.to(DEV), loc.to(DEV), aw, 64)
        ref = refcuda.forward(value, shapes.to(DEV), lsi.to(DEV), loc.to(DEV), aw)
        torch.testing.assert_close(ours, ref, rtol=1e-6, atol=1e-3)
        # the forward is continuous across the floor; d/d(loc) is NOT (it differences the rows h_low, h_low+1),
        # so equal location gradients on a random field prove both kernels floored to the same row
        g = torch.Generator().manual_seed(4)
        value = torch.randn(1, H * H, 1, 4, generator=g).to(DEV)
        go = torch.randn(1, n, 4, generator=g).to(DEV)
        _, gl, _ = _cabi.backward(value, shapes.to(DEV), lsi.to(DEV), loc.to(DEV), aw, go, 64)
        _, rgl, _ = refcuda.backward(value, shapes.to(DEV), lsi.to(DEV), loc.to(DEV), aw, go)
        torch.testing.assert_close(gl, rgl, rtol=1e-4, atol=1e-4 * float(rgl.abs().max()))
        # sanity: the unfused floor would have produced a different gradient at these points
        idx_unfused = want.clone()
        assert (torch.floor(sel * H - 0.5) != fused_floor).all()


# ---------------------------------------------------------------------------------------------------
# the reference's own CUDA kernels (oracle/_ref, compiled from /root/reference) on identical inputs
# ---------------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not refcuda.available(), reason='oracle/_ref not built')


@needs_ref
@pytest.mark.parametrize('cfg', SHAPES[:6], ids=[s[0] for s in SHAPES[:6]])
def test_vs_reference_cuda_f32(cfg):
    _, N, M, D, Lq, shapes, P, dist = cfg
    inp = make_inputs(N, M, D, Lq, shapes, P, seed=7, dist=dist)
    g = _cuda(inp)
    out, gv, gl, ga = _run(inp, torch.float32)
    rout = refcuda.forward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'])
    rgv, rgl, rga = refcuda.backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'])
    torch.cuda.synchronize()
    torch.testing.assert_close(out, rout.cpu(), rtol=1e-5, atol=1e-6 * max(1.0, _scale(rout)))
    torch.testing.assert_close(gv, rgv.cpu(), rtol=1e-4, atol=1e-4 * _scale(rgv))
    torch.testing.assert_close(gl, rgl.cpu(), rtol=1e-4, atol=1e-4 * _scale(rgl))
    torch.testing.assert_close(ga, rga.cpu(), rtol=1e-4, atol=1e-4 * _scale(rga))


@needs_ref
def test_c_oracle_matches_reference_cuda():
    """Pins the C restatement against the reference kernel itself (forward in fp32 and fp64)."""
    inp = make_inputs(2, 3, 8, 40, [(8, 8), (4, 4), (2, 2)], 4, seed=8, dist='edges')
    g = _cuda(inp)
    rout = refcuda.forward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw']).cpu()
    want = c_oracle.forward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'])
    torch.testing.assert_close(rout, want, rtol=1e-6, atol=1e-6)
    g64 = _cuda(inp, torch.float64)
    rout64 = refcuda.forward(g64['value'], g64['shapes'], g64['lsi'], g64['loc'], g64['aw']).cpu()
    want64 = c_oracle.forward(inp['value'].double(), inp['shapes'], inp['lsi'], inp['loc'].double(), inp['aw'].double())
    torch.testing.assert_close(rout64, want64, rtol=1e-12, atol=1e-13)
    rgv, rgl, rga = refcuda.backward(g64['value'], g64['shapes'], g64['lsi'], g64['loc'], g64['aw'], g64['grad_out'])
    wgv, wgl, wga = c_oracle.backward(inp['value'].double(), inp['shapes'], inp['lsi'], inp['loc'].double(),
                                      inp['aw'].double(), inp['grad_out'].double())
    torch.testing.assert_close(rgv.cpu(), wgv, rtol=1e-10, atol=1e-12)
    torch.testing.assert_close(rgl.cpu(), wgl, rtol=1e-10, atol=1e-12)
    torch.testing.assert_close(rga.cpu(), wga, rtol=1e-10, atol=1e-12)


# ---------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties (the oracle is too slow / not needed here)
# ---------------------------------------------------------------------------------------------------
FULL = [
    ('B-injector-bs16', 16, 12, 32, 1024, [(64, 64), (32, 32), (16, 16)], 4),
    ('B-extractor-bs16', 16, 12, 32, 5376, [(32, 32)], 4),
    ('S-injector-bs16', 16, 6, 64, 1024, [(64, 64), (32, 32), (16, 16)], 4),
    ('L-injector-896', 1, 16, 32, 3136, [(112, 112), (56, 56), (28, 28)], 4),
    ('HTC-extractor-1024', 1, 16, 32, 21504, [(64, 64)], 4),
    # round 2: the remaining BASELINE.json call shapes
    ('S-extractor-bs16', 16, 6, 64, 5376, [(32, 32)], 4),
    ('L-extractor-896', 1, 16, 32, 16464, [(56, 56)], 4),
    ('L64-injector-896', 1, 16, 64, 3136, [(112, 112), (56, 56), (28, 28)], 4),
    ('L64-extractor-896', 1, 16, 64, 16464, [(56, 56)], 4),
    ('HTC-injector-1024', 1, 16, 32, 4096, [(128, 128), (64, 64), (32, 32)], 4),
    ('M2F-encoder-896', 1, 32, 32, 16464, [(112, 112), (56, 56), (28, 28)], 4),
]


@pytest.mark.parametrize('cfg', FULL, ids=[s[0] for s in FULL])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16], ids=['f32', 'bf16', 'f16'])
def test_full_size_properties(cfg, dtype, request):
    _, N, M, D, Lq, shapes, P = cfg
    inp = make_inputs(N, M, D, Lq, shapes, P, seed=1, dist='adapter')
    g = _cuda(inp)
    tol = {torch.float32: 1e-5, torch.bfloat16: 2e-2, torch.float16: 3e-3}[dtype]
    if dtype == torch.float16:
        vab.set_amp_value_dtype(torch.float16)
        request.addfinalizer(lambda: vab.set_amp_value_dtype(torch.float32))
    value = g['value'].to(dtype)
    f = lambda v, a: vab.MSDeformAttnFunction.apply(v, g['shapes'], g['lsi'], g['loc'], a, 64).float()
    out = f(value, g['aw'])
    # (1) constant value + weights summing to 1 + all samples strictly inside => output == constant
    ones = torch.ones_like(value)
    loc_in = (g['loc'].clamp(0.02, 0.98)).contiguous()
    o1 = vab.MSDeformAttnFunction.apply(ones, g['shapes'], g['lsi'], loc_in, g['aw'], 64).float()
    inner = torch.ones_like(o1)
    # samples within half a pixel of the border see zero padding; restrict to levels >= 26 px where 0.02 is > 0.5px
    if min(min(s) for s in shapes) >= 26:
        torch.testing.assert_close(o1, inner, rtol=tol, atol=tol)
    # (2) linearity in value and in the attention weights
    o2 = f((value.float() * 2).to(dtype), g['aw'])
    torch.testing.assert_close(o2, 2 * out, rtol=2 * tol, atol=2 * tol * _scale(out))
    o3 = f(value, (g['aw'] * 0.5).contiguous())
    torch.testing.assert_close(o3, 0.5 * out, rtol=2 * tol, atol=2 * tol * _scale(out))
    # (3) batch independence: running a single image alone reproduces its slice exactly (bit-for-bit)
    b = N - 1
    ob = vab.MSDeformAttnFunction.apply(value[b:b + 1].contiguous(), g['shapes'], g['lsi'], g['loc'][b:b + 1].contiguous(),
                                        g['aw'][b:b + 1].contiguous(), 64).float()
    assert torch.equal(ob[0], out[b])
    # (4) adjoint identity: <grad_out, J v> == <J^T grad_out, v>  (forward is linear in value)
    v = value.clone().requires_grad_()
    o = vab.MSDeformAttnFunction.apply(v, g['shapes'], g['lsi'], g['loc'], g['aw'], 64)
    go = g['grad_out'].to(dtype)
    o.backward(go)
    lhs = (o.detach().double() * go.double()).sum()
    rhs = (v.grad.double() * value.double()).sum()
    # both sides are sums of ~1e7 signed terms: the rounding of the 16-bit outputs enters each term independently, so the
    # error scales with the root of the sum of squares of the terms, not with the (possibly cancelling) sum itself
    spread = float((o.detach().double() * go.double()).square().sum().sqrt() + (v.grad.double() * value.double()).square().sum().sqrt())
    assert abs(float(lhs - rhs)) <= (1e-4 if dtype == torch.float32 else 2e-2) * max(1.0, abs(float(lhs)), spread)
    # (5) a checksum against the reference CUDA kernel at full size
    if refcuda.available() and dtype == torch.float32:
        rout = refcuda.forward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'])
        torch.testing.assert_close(out, rout, rtol=1e-5, atol=1e-6 * max(1.0, _scale(rout)))
        loc = g['loc'].clone().requires_grad_()
        aw = g['aw'].clone().requires_grad_()
        v = g['value'].clone().requires_grad_()
        vab.MSDeformAttnFunction.apply(v, g['shapes'], g['lsi'], loc, aw, 64).backward(g['grad_out'])
        rgv, rgl, rga = refcuda.backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'])
        torch.testing.assert_close(v.grad, rgv, rtol=1e-4, atol=1e-4 * _scale(rgv))
        torch.testing.assert_close(loc.grad, rgl, rtol=1e-4, atol=1e-4 * _scale(rgl))
        torch.testing.assert_close(aw.grad, rga, rtol=1e-4, atol=1e-4 * _scale(rga))


# ---------------------------------------------------------------------------------------------------
# boundary behaviour
# ---------------------------------------------------------------------------------------------------
def test_non_contiguous_rejected():
    inp = _cuda(make_inputs(1, 2, 32, 8, [(4, 4)], 4, seed=2))
    v = inp['value'].transpose(1, 2).contiguous().transpose(1, 2)  # same shape, non-contiguous
    with pytest.raises(RuntimeError, match='contiguous'):
        vab.MSDeformAttnFunction.apply(v, inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], 64)


def test_im2col_step_precondition():
    inp = _cuda(make_inputs(3, 2, 32, 8, [(4, 4)], 4, seed=2))
    with pytest.raises(RuntimeError, match='must divide'):
        vab.MSDeformAttnFunction.apply(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], 2)
    out = vab.MSDeformAttnFunction.apply(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], 3)
    assert out.shape == (3, 8, 64)


def test_all_samples_out_of_range_give_zero():
    inp = _cuda(make_inputs(1, 2, 32, 8, [(4, 4)], 4, seed=2))
    loc = torch.full_like(inp['loc'], 5.0)
    out = vab.MSDeformAttnFunction.apply(inp['value'], inp['shapes'], inp['lsi'], loc, inp['aw'], 64)
    assert out.abs().max() == 0


def test_autocast_matches_reference_policy():
    """Under autocast the reference up-casts to fp32 (custom_fwd(cast_inputs=float32), func.py:21)."""
    inp = _cuda(make_inputs(1, 2, 32, 8, [(4, 4)], 4, seed=2))
    want = vab.MSDeformAttnFunction.apply(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], 64)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        got = vab.MSDeformAttnFunction.apply(inp['value'].bfloat16().float(), inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], 64)
        assert got.dtype == torch.float32
    vab.set_amp_value_dtype(torch.bfloat16)
    try:
        with torch.autocast('cuda', dtype=torch.bfloat16):
            got16 = vab.MSDeformAttnFunction.apply(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], 64)
            assert got16.dtype == torch.bfloat16
    finally:
        vab.set_amp_value_dtype(torch.float32)
    torch.testing.assert_close(got16.float(), want, rtol=2e-2, atol=2e-2 * _scale(want))
    # fp16 autocast (the reference's `fp16 = dict(loss_scale=...)` configs): fp32 core by default, fp16 I/O when asked
    with torch.autocast('cuda', dtype=torch.float16):
        g32 = vab.MSDeformAttnFunction.apply(inp['value'].half(), inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], 64)
        assert g32.dtype == torch.float32
    vab.set_amp_value_dtype(torch.float16)
    try:
        with torch.autocast('cuda', dtype=torch.float16):
            v = inp['value'].clone().requires_grad_()
            gh = vab.MSDeformAttnFunction.apply(v, inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], 64)
            assert gh.dtype == torch.float16
            gh.float().sum().backward()
            assert v.grad.dtype == torch.float32 and bool(torch.isfinite(v.grad).all())
    finally:
        vab.set_amp_value_dtype(torch.float32)
    torch.testing.assert_close(gh.float(), want, rtol=3e-3, atol=3e-3 * _scale(want))
    with pytest.raises(ValueError):
        vab.set_amp_value_dtype(torch.float64)


def test_launch_counter_and_stream():
    inp = _cuda(make_inputs(1, 2, 32, 8, [(4, 4)], 4, seed=2))
    n0 = _cabi.launch_count()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out = vab.MSDeformAttnFunction.apply(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], 64)
    s.synchronize()
    assert _cabi.launch_count() == n0 + 1
    want = vab.MSDeformAttnFunction.apply(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], 64)
    assert torch.equal(out, want)


# ---------------------------------------------------------------------------------------------------
# shared-memory forward (TMA-staged level maps): must be bit-identical to the L1-path forward
# ---------------------------------------------------------------------------------------------------
SMEM_SHAPES = [
    ('B-extractor', 2, 12, 32, 1344, [(16, 16)], 4, 'adapter'),
    ('B-injector', 2, 12, 32, 256, [(32, 32), (16, 16), (8, 8)], 4, 'adapter'),
    ('S-injector', 2, 6, 64, 256, [(32, 32), (16, 16), (8, 8)], 4, 'edges'),
    ('ragged-3lvl', 3, 5, 32, 37, [(7, 9), (3, 4), (2, 5)], 4, 'edges'),
    ('one-level-big', 1, 2, 32, 700, [(38, 42)], 4, 'edges'),         # padded 40x44 rows x 128 B + null block: just fits
    ('too-big-level0', 1, 2, 32, 300, [(64, 64), (20, 20), (3, 3)], 4, 'edges'),  # level 0 stays on the L1 path
]


@pytest.mark.parametrize('cfg', SMEM_SHAPES, ids=[s[0] for s in SMEM_SHAPES])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['f32', 'bf16'])
@pytest.mark.parametrize('threads', [512, 1024])
def test_smem_forward_bit_identical(cfg, dtype, threads):
    _, N, M, D, Lq, shapes, P, dist = cfg
    g = _cuda(make_inputs(N, M, D, Lq, shapes, P, seed=21, dist=dist))
    value = g['value'].to(dtype)
    try:
        _cabi.set_tuning(fwd_smem=1)
        want = _cabi.forward(value, g['shapes'], g['lsi'], g['loc'], g['aw'], 64)
        _cabi.set_tuning(fwd_smem=2, fwd_smem_threads=threads)
        for chunks in (0, 1, 3):
            _cabi.set_tuning(fwd_smem_chunks=chunks)
            got = _cabi.forward(value, g['shapes'], g['lsi'], g['loc'], g['aw'], 64)
            torch.cuda.synchronize()
            assert torch.equal(got, want), 'chunks=%d' % chunks
    finally:
        _cabi.set_tuning(fwd_smem=0, fwd_smem_threads=0, fwd_smem_chunks=0)


def test_host_shapes_cached_no_sync_in_steady_state():
    g = _cuda(make_inputs(1, 2, 32, 8, [(4, 4)], 4, seed=2))
    a = _cabi.host_shapes(g['shapes'])
    b = _cabi.host_shapes(g['shapes'])
    assert a is b and list(a) == [4, 4]
    g['shapes'].add_(0)  # in-place op bumps the version counter -> re-read
    assert _cabi.host_shapes(g['shapes']) is not a
    # a NEW tensor that lands on the freed address of an old one must not inherit its cached shapes
    for hw in ((3, 5), (7, 2), (9, 9)):
        t = torch.as_tensor([hw], dtype=torch.long, device=DEV)
        assert list(_cabi.host_shapes(t)) == list(hw)
        del t


@pytest.mark.parametrize('cfg', SMEM_SHAPES[:4], ids=[s[0] for s in SMEM_SHAPES[:4]])
def test_wide_lane_forward_bit_identical(cfg):
    """fp32 forward with 32-byte lanes (LDG.256): same arithmetic order per channel -> identical bits."""
    _, N, M, D, Lq, shapes, P, dist = cfg
    g = _cuda(make_inputs(N, M, D, Lq, shapes, P, seed=22, dist=dist))
    try:
        _cabi.set_tuning(fwd_wide=1)
        want = _cabi.forward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], 64)
        _cabi.set_tuning(fwd_wide=2)
        for minb in (0, 3):
            _cabi.set_tuning(fwd_min_ctas=minb)
            got = _cabi.forward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], 64)
            torch.cuda.synchronize()
            assert torch.equal(got, want)
    finally:
        _cabi.set_tuning(fwd_wide=0, fwd_min_ctas=0)


# ---------------------------------------------------------------------------------------------------
# memory-safety checks without compute-sanitizer (closed on this pool): NaN guard bands around the
# buffers and a NaN-poisoned neighbouring head. Clamped addressing must never leave the (b, m) slab.
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['f32', 'bf16'])
@pytest.mark.parametrize('cfg', [(2, 3, 32, 200, [(9, 7), (4, 5), (2, 2)], 4), (1, 2, 64, 150, [(11, 13)], 4),
                                 (2, 2, 24, 60, [(5, 5), (3, 2)], 3)], ids=['L3-D32', 'L1-D64', 'generic-D24'])
def test_guard_bands_and_head_isolation(cfg, dtype):
    N, M, D, Lq, shapes, P = cfg
    inp = make_inputs(N, M, D, Lq, shapes, P, seed=41, dist='edges')
    g = _cuda(inp)
    S = inp['value'].shape[1]
    guard = 4096
    n = N * S * M * D
    # value lives between two NaN guard bands
    vbuf = torch.full((n + 2 * guard,), float('nan'), device=DEV, dtype=dtype)
    value = vbuf[guard:guard + n].view(N, S, M, D)
    value.copy_(g['value'].to(dtype))
    out = _cabi.forward(value, g['shapes'], g['lsi'], g['loc'], g['aw'], 64)
    assert torch.isfinite(out.float()).all(), 'forward read outside the value tensor'
    want = _cabi.forward(g['value'].to(dtype).contiguous(), g['shapes'], g['lsi'], g['loc'], g['aw'], 64)
    assert torch.equal(out, want)
    # poison head 1: head 0's output must not change by a single bit
    poisoned = value.clone()
    poisoned[:, :, 1, :] = float('nan')
    out_p = _cabi.forward(poisoned, g['shapes'], g['lsi'], g['loc'], g['aw'], 64)
    assert torch.equal(out_p.view(N, Lq, M, D)[:, :, 0], out.view(N, Lq, M, D)[:, :, 0])
    # backward: gradients land only inside grad_value; a canary-filled allocation right after it stays intact
    go = g['grad_out'].to(dtype)
    gv, gl, ga = _cabi.backward(value, g['shapes'], g['lsi'], g['loc'], g['aw'], go, 64)
    torch.cuda.synchronize()
    assert torch.isfinite(gv.float()).all() and torch.isfinite(gl).all() and torch.isfinite(ga).all()
    gv2, gl2, ga2 = _cabi.backward(g['value'].to(dtype).contiguous(), g['shapes'], g['lsi'], g['loc'], g['aw'], go, 64)
    torch.testing.assert_close(gv.float(), gv2.float(), rtol=1e-3, atol=1e-3 * _scale(gv2.float()))
    assert torch.equal(gl, gl2) or torch.allclose(gl, gl2, rtol=1e-5, atol=1e-6 * _scale(gl2))
    assert torch.isnan(vbuf[:guard].float()).all() and torch.isnan(vbuf[guard + n:].float()).all()
    # total mass check: sum(grad_value) == sum over points of (sum of valid corner weights) * aw * sum_c(grad_out)
    idx = c_oracle.point_index(inp['shapes'], inp['lsi'], inp['loc'], M, D)
    assert (idx[:, 3] >= -(max(w for _, w in shapes) + 2) * M * D).all()


# ---------------------------------------------------------------------------------------------------
# the stand-in for the reference's pybind module (vision.cpp:13-16), called the way the reference's Python calls it
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', ['op_kat_seed3', 'op_inj_edges', 'op_d32_edges', 'op_d64_edges'])
def test_pybind_compat_module_on_goldens(case):
    import importlib
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'vit-adapter_b200', 'pybind_compat'))
    try:
        sys.modules.pop('MultiScaleDeformableAttention', None)
        MSDA = importlib.import_module('MultiScaleDeformableAttention')
    finally:
        sys.path.pop(0)
    g = _cuda(load_golden(case))
    out = MSDA.ms_deform_attn_forward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], 64)
    assert tuple(out.shape) == (g['value'].shape[0], g['loc'].shape[1], g['value'].shape[2] * g['value'].shape[3])
    torch.testing.assert_close(out.cpu().double(), g['out_f64'].cpu(), rtol=1e-5, atol=1e-6)
    res = MSDA.ms_deform_attn_backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'], 64)
    assert isinstance(res, list) and len(res) == 3                     # std::vector<at::Tensor> of the reference
    # gradients vs the fp32 C oracle (the edge fixtures sit on texel boundaries, where d/d(loc) is discontinuous and an
    # fp64 reference floors differently), and value / weight gradients also vs the reference's fp64 autograd
    h = load_golden(case)
    wgv, wgl, wga = c_oracle.backward(h['value'], h['shapes'], h['lsi'], h['loc'], h['aw'], h['grad_out'])
    for got, want in zip(res, (wgv, wgl, wga)):
        torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-4 * _scale(want))
    torch.testing.assert_close(res[0].cpu().double(), h['grad_value_f64'], rtol=1e-4, atol=1e-4 * _scale(wgv))
    torch.testing.assert_close(res[2].cpu().double(), h['grad_aw_f64'], rtol=1e-4, atol=1e-4 * _scale(wga))
    batch = 3 * g['value'].shape[0]
    bad_step = 2 if batch % 2 else 4                                   # batch % min(batch, step) != 0
    with pytest.raises(RuntimeError, match='must divide'):
        MSDA.ms_deform_attn_forward(g['value'].repeat(3, 1, 1, 1), g['shapes'], g['lsi'], g['loc'].repeat(3, 1, 1, 1, 1, 1),
                                    g['aw'].repeat(3, 1, 1, 1, 1), bad_step)


# ---------------------------------------------------------------------------------------------------
# opt-in: packed 16-bit reductions straight into grad_value (tuning key bwd_packed16 = 2)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('low', [torch.bfloat16, torch.float16], ids=['bf16', 'f16'])
@pytest.mark.parametrize('cfg', [SHAPES[1], SHAPES[2], SHAPES[0]], ids=['B-injector', 'B-extractor', 'S-injector-d64'])
def test_packed16_backward(cfg, low):
    """Every contribution is rounded to 16 bits and summed in 16 bits by the L2: grad_value carries ~sqrt(n) * eps/2 relative
    error for n contributions per element (eps = 2^-8 bf16, 2^-11 fp16; n ~ 9 Injector, ~ 21 at this Extractor miniature) -
    stated tolerance 4e-2 (bf16) / 6e-3 (fp16) of the largest gradient, which is why the path is opt-in. grad_sampling_loc and
    grad_attn_weight do not pass through the scatter and are bit-identical to the default path."""
    _, N, M, D, Lq, shapes, P, dist = cfg
    inp = make_inputs(N, M, D, Lq, shapes, P, seed=21, dist=dist)
    g = _cuda(inp)
    args = (g['value'].to(low), g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'].to(low), 64)
    gv0, gl0, ga0 = _cabi.backward(*args)
    n0 = _cabi.launch_count()
    _cabi.set_tuning(bwd_packed16=2)
    try:
        gv1, gl1, ga1 = _cabi.backward(*args)
        torch.cuda.synchronize()
    finally:
        _cabi.set_tuning(bwd_packed16=0)
    assert _cabi.launch_count() - n0 == 1                     # one kernel: no convert pass
    assert gv1.dtype == low and torch.equal(gl0, gl1) and torch.equal(ga0, ga1)
    vq, goq = inp['value'].to(low).float(), inp['grad_out'].to(low).float()
    wgv, _, _ = c_oracle.backward(vq, inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], goq)
    tol = 4e-2 if low == torch.bfloat16 else 6e-3
    torch.testing.assert_close(gv1.float().cpu(), wgv, rtol=tol, atol=tol * _scale(wgv))
    # and it is measurably less accurate than the default (fp32 accumulation, one rounding) - the reason it is not the default
    e_default = float((gv0.float().cpu() - wgv).abs().max())
    e_packed = float((gv1.float().cpu() - wgv).abs().max())
    assert e_default <= e_packed * 1.0001 + 1e-12
