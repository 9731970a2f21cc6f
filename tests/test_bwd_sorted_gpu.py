"""GPU parity tests of the slab-sorted backward (csrc/msda_bwd_sorted.cu) against the C oracle, the query-order backward of
the same library and, where built, the reference's own CUDA kernels.

The sorted backward applies to D in {32, 64}; it is chosen automatically where it measured faster (D = 64, long cell runs) and
`set_tuning(bwd_sorted=2)` forces it wherever it applies (`=1` switches it off).
The cases cover ranges that end inside a batch of 32 positions, cell runs that straddle two warps / two CTAs, empty slabs,
border cells whose clamped key aliases several cells, runtime (L, P) and long runs on hot counters.
Tolerances: fp32 gradients 1e-4 relative (summation order), bf16 1e-2, fp16 2e-3."""
import pytest
import torch

from conftest import make_inputs
from oracle import c_oracle, refcuda

import vit_adapter_b200 as vab
from vit_adapter_b200 import _cabi

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _scale(t):
    return float(t.abs().max()) + 1e-30


def _bwd(inp, dtype, **tuning):
    g = {k: v.to(DEV) for k, v in inp.items()}
    _cabi.set_tuning(**tuning)
    try:
        gv, gl, ga = _cabi.backward(g['value'].to(dtype), g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'].to(dtype), 64)
        torch.cuda.synchronize()
    finally:
        _cabi.set_tuning(**{k: 0 for k in tuning})
    return gv.float().cpu(), gl.cpu(), ga.cpu()


CASES = [
    # name, N, M, D, Lq, shapes, P, dist
    ('B-injector', 2, 12, 32, 256, [(32, 32), (16, 16), (8, 8)], 4, 'adapter'),
    ('B-extractor', 2, 12, 32, 1344, [(16, 16)], 4, 'adapter'),
    ('S-injector-d64', 2, 6, 64, 256, [(32, 32), (16, 16), (8, 8)], 4, 'adapter'),
    ('edges-3lvl', 1, 16, 32, 196, [(28, 28), (14, 14), (7, 7)], 4, 'edges'),
    ('uniform-d64', 1, 16, 64, 196, [(28, 28), (14, 14), (7, 7)], 4, 'uniform'),
    ('ragged-runtime-LP', 3, 5, 32, 37, [(7, 9), (3, 4), (1, 1), (2, 5)], 3, 'edges'),
    ('one-query', 1, 1, 32, 1, [(2, 2)], 1, 'edges'),
    ('p8-one-level', 2, 3, 64, 45, [(12, 10)], 8, 'edges'),
    ('tiny-map-many-points', 1, 2, 32, 700, [(2, 3)], 4, 'uniform'),   # ~470 points per cell: long runs, hot ATOMS keys
    ('wide-row', 1, 2, 32, 90, [(1, 40), (40, 1)], 4, 'edges'),
    ('many-parts', 1, 2, 32, 4200, [(6, 7)], 4, 'uniform'),   # 2 slabs: the sort cuts each into 64 parts (8 per prefix group)
    ('parts-not-multiple-of-groups', 1, 3, 64, 1300, [(9, 5), (4, 4)], 2, 'uniform'),   # 20 parts: groups of 3, the last two short
]


@pytest.mark.parametrize('cfg', CASES, ids=[c[0] for c in CASES])
def test_sorted_backward_vs_oracle_f32(cfg):
    _, N, M, D, Lq, shapes, P, dist = cfg
    inp = make_inputs(N, M, D, Lq, shapes, P, seed=11, dist=dist)
    wgv, wgl, wga = c_oracle.backward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], inp['grad_out'])
    gv, gl, ga = _bwd(inp, torch.float32, bwd_sorted=2)
    torch.testing.assert_close(gv, wgv, rtol=1e-4, atol=1e-4 * _scale(wgv))
    torch.testing.assert_close(gl, wgl, rtol=1e-4, atol=1e-4 * _scale(wgl))
    torch.testing.assert_close(ga, wga, rtol=1e-4, atol=1e-4 * _scale(wga))
    # and against the query-order kernel of this library (same tolerance: only the summation order differs)
    ogv, ogl, oga = _bwd(inp, torch.float32, bwd_sorted=1)
    torch.testing.assert_close(gv, ogv, rtol=1e-4, atol=1e-4 * _scale(ogv))
    torch.testing.assert_close(gl, ogl, rtol=1e-4, atol=1e-4 * _scale(ogl))
    torch.testing.assert_close(ga, oga, rtol=1e-4, atol=1e-4 * _scale(oga))


@pytest.mark.parametrize('low', [torch.bfloat16, torch.float16], ids=['bf16', 'f16'])
@pytest.mark.parametrize('cfg', CASES[:6], ids=[c[0] for c in CASES[:6]])
def test_sorted_backward_vs_oracle_16bit(cfg, low):
    _, N, M, D, Lq, shapes, P, dist = cfg
    inp = make_inputs(N, M, D, Lq, shapes, P, seed=12, dist=dist)
    vq, goq = inp['value'].to(low).float(), inp['grad_out'].to(low).float()
    wgv, wgl, wga = c_oracle.backward(vq, inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], goq)
    tol = 1e-2 if low == torch.bfloat16 else 2e-3
    gv, gl, ga = _bwd(inp, low, bwd_sorted=2)
    torch.testing.assert_close(gv, wgv, rtol=tol, atol=tol * _scale(wgv))
    # location / weight gradients are fp32 sums of exactly representable products: fp32-grade agreement
    torch.testing.assert_close(gl, wgl, rtol=1e-4, atol=1e-4 * _scale(wgl))
    torch.testing.assert_close(ga, wga, rtol=1e-4, atol=1e-4 * _scale(wga))


def test_sorted_backward_out_of_range_points_get_zero_gradients():
    inp = make_inputs(1, 2, 32, 40, [(5, 6), (3, 3)], 4, seed=13, dist='uniform')
    inp['loc'][:, ::2] = 7.5          # every other query samples far outside the maps
    inp['loc'][:, 1, :, :, 0] = -3.0
    gv, gl, ga = _bwd(inp, torch.float32, bwd_sorted=2)
    wgv, wgl, wga = c_oracle.backward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], inp['grad_out'])
    assert float(gl[:, ::2].abs().max()) == 0.0 and float(ga[:, ::2].abs().max()) == 0.0
    torch.testing.assert_close(gv, wgv, rtol=1e-4, atol=1e-4 * _scale(wgv))
    torch.testing.assert_close(gl, wgl, rtol=1e-4, atol=1e-4 * _scale(wgl))
    torch.testing.assert_close(ga, wga, rtol=1e-4, atol=1e-4 * _scale(wga))
    # nothing in range at all: grad_value stays zero, the other gradients are written (not left uninitialised)
    inp['loc'][:] = 9.0
    gv, gl, ga = _bwd(inp, torch.float32, bwd_sorted=2)
    assert float(gv.abs().max()) == 0.0 and float(gl.abs().max()) == 0.0 and float(ga.abs().max()) == 0.0


def test_sorted_backward_all_points_in_one_cell():
    """Every sample of every query lands in the same bilinear cell: one run per warp range, hot counters."""
    inp = make_inputs(2, 4, 32, 333, [(9, 9)], 4, seed=14, dist='uniform')
    inp['loc'] = (0.5 + 0.02 * (inp['loc'] - 0.5)).contiguous()
    gv, gl, ga = _bwd(inp, torch.float32, bwd_sorted=2)
    wgv, wgl, wga = c_oracle.backward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], inp['grad_out'])
    torch.testing.assert_close(gv, wgv, rtol=1e-4, atol=1e-4 * _scale(wgv))
    torch.testing.assert_close(gl, wgl, rtol=1e-4, atol=1e-4 * _scale(wgl))
    torch.testing.assert_close(ga, wga, rtol=1e-4, atol=1e-4 * _scale(wga))


@pytest.mark.skipif(not refcuda.available(), reason='oracle/_ref not built')
@pytest.mark.parametrize('cfg', CASES[:5], ids=[c[0] for c in CASES[:5]])
def test_sorted_backward_vs_reference_cuda(cfg):
    _, N, M, D, Lq, shapes, P, dist = cfg
    inp = make_inputs(N, M, D, Lq, shapes, P, seed=15, dist=dist)
    g = {k: v.to(DEV) for k, v in inp.items()}
    rgv, rgl, rga = refcuda.backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'])
    gv, gl, ga = _bwd(inp, torch.float32, bwd_sorted=2)
    torch.testing.assert_close(gv, rgv.cpu(), rtol=1e-4, atol=1e-4 * _scale(rgv))
    torch.testing.assert_close(gl, rgl.cpu(), rtol=1e-4, atol=1e-4 * _scale(rgl))
    torch.testing.assert_close(ga, rga.cpu(), rtol=1e-4, atol=1e-4 * _scale(rga))


def test_sorted_backward_selection():
    """Without tuning the sorted backward (4 kernels: histogram, prefix, scatter, walk) runs where it measured faster - one
    level with >= 16 samples per value token and head, and >= 0.25 M samples at 64 channels per head / >= 1 M at 32 - and the
    one-kernel query-order backward elsewhere; bwd_sorted=2 / 1 force one or the other. They agree to summation order."""
    big64 = make_inputs(4, 6, 64, 5376, [(32, 32)], 4, seed=16, dist='adapter')      # ViT-Adapter-S Extractor, 4 images: 516 k samples
    n0 = _cabi.launch_count()
    gv, gl, ga = _bwd(big64, torch.float32)
    assert _cabi.launch_count() - n0 == 4
    n0 = _cabi.launch_count()
    ogv, ogl, oga = _bwd(big64, torch.float32, bwd_sorted=1)
    assert _cabi.launch_count() - n0 == 1
    torch.testing.assert_close(gv, ogv, rtol=1e-4, atol=1e-4 * _scale(ogv))
    torch.testing.assert_close(gl, ogl, rtol=1e-4, atol=1e-4 * _scale(ogl))
    torch.testing.assert_close(ga, oga, rtol=1e-4, atol=1e-4 * _scale(oga))
    for inp in (make_inputs(2, 6, 64, 1344, [(16, 16)], 4, seed=16, dist='adapter'),                # dense, but a small call
                make_inputs(4, 6, 32, 5376, [(32, 32)], 4, seed=16, dist='adapter'),                # 516 k samples at 32 channels
                make_inputs(2, 6, 64, 256, [(32, 32), (16, 16), (8, 8)], 4, seed=16, dist='adapter')):  # 2 samples per token
        n0 = _cabi.launch_count()
        _bwd(inp, torch.float32)
        assert _cabi.launch_count() - n0 == 1
        n0 = _cabi.launch_count()
        _bwd(inp, torch.float32, bwd_sorted=2)
        assert _cabi.launch_count() - n0 == 4
    # a head dimension outside {32, 64} has no sorted kernel: one launch whatever the knob says
    inp = make_inputs(1, 2, 16, 50, [(6, 6)], 4, seed=17, dist='uniform')
    n0 = _cabi.launch_count()
    _bwd(inp, torch.float32, bwd_sorted=2)
    assert _cabi.launch_count() - n0 == 1


def test_sorted_backward_many_warps_per_slab():
    """A slab long enough for several CTAs of the walker (ranges of 32..256 positions per warp, runs cut at every boundary)."""
    inp = make_inputs(1, 2, 32, 6000, [(12, 12)], 4, seed=18, dist='uniform')
    wgv, wgl, wga = c_oracle.backward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], inp['grad_out'])
    gv, gl, ga = _bwd(inp, torch.float32, bwd_sorted=2)
    torch.testing.assert_close(gv, wgv, rtol=1e-4, atol=1e-4 * _scale(wgv))
    torch.testing.assert_close(gl, wgl, rtol=1e-4, atol=1e-4 * _scale(wgl))
    torch.testing.assert_close(ga, wga, rtol=1e-4, atol=1e-4 * _scale(wga))


@pytest.mark.parametrize('seed', range(12))
def test_sorted_backward_random_shapes(seed):
    """Seeded random shapes (levels 1-4 with ragged maps, points 1-8, heads 1-7, D in {32, 64}, ragged query counts), forced
    through the sorted backward, against the C oracle."""
    import random
    rnd = random.Random(1000 + seed)
    L = rnd.randint(1, 4)
    shapes = [(rnd.randint(1, 12), rnd.randint(1, 12)) for _ in range(L)]
    N, M, D = rnd.randint(1, 3), rnd.randint(1, 7), rnd.choice([32, 64])
    Lq, P = rnd.choice([1, 5, 33, 97, 260]), rnd.randint(1, 8)
    inp = make_inputs(N, M, D, Lq, shapes, P, seed=2000 + seed, dist=rnd.choice(['uniform', 'edges']))
    wgv, wgl, wga = c_oracle.backward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], inp['grad_out'])
    gv, gl, ga = _bwd(inp, torch.float32, bwd_sorted=2)
    torch.testing.assert_close(gv, wgv, rtol=1e-4, atol=1e-4 * _scale(wgv))
    torch.testing.assert_close(gl, wgl, rtol=1e-4, atol=1e-4 * _scale(wgl))
    torch.testing.assert_close(ga, wga, rtol=1e-4, atol=1e-4 * _scale(wga))


def test_sorted_backward_in_a_cuda_graph():
    """The four kernels of the sorted backward and its workspace allocation capture into a CUDA graph (nothing in the call
    synchronises); replays give the eager result to summation order."""
    inp = make_inputs(2, 6, 64, 1344, [(16, 16)], 4, seed=21, dist='adapter')
    g = {k: v.to(DEV) for k, v in inp.items()}
    _cabi.set_tuning(bwd_sorted=2)
    try:
        _sorted_in_a_graph(g)
    finally:
        _cabi.set_tuning(bwd_sorted=0)


def _sorted_in_a_graph(g):
    eager = _cabi.backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'], 64)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        _cabi.backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'], 64)
        torch.cuda.synchronize()
        n0 = _cabi.launch_count()
        with torch.cuda.graph(graph, stream=s):
            out = _cabi.backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'], 64)
        assert _cabi.launch_count() - n0 == 4   # the sorted path was the one captured
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    for a, b in zip(out, eager):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4 * _scale(b))


@pytest.mark.parametrize('cfg', [CASES[0], CASES[1], CASES[5], CASES[8]], ids=[CASES[i][0] for i in (0, 1, 5, 8)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['f32', 'bf16'])
def test_sorted_backward_stays_inside_its_buffers(cfg, dtype):
    """Every buffer the sorted backward writes - grad_value, grad_sampling_loc, grad_attn_weight and the workspace (fp32
    accumulator, histograms, scan totals, sorted indices) - sits between two guard bands filled with a pattern; the call must
    leave the bands untouched (compute-sanitizer is not available on the GPU pool, so the ABI is called with guarded memory)."""
    import ctypes
    _, N, M, D, Lq, shapes, P, dist = cfg
    inp = make_inputs(N, M, D, Lq, shapes, P, seed=31, dist=dist)
    g = {k: v.to(DEV) for k, v in inp.items()}
    value, go = g['value'].to(dtype).contiguous(), g['grad_out'].to(dtype).contiguous()
    lib = _cabi.load()
    dims = _cabi._dims(value, g['shapes'], g['loc'])
    code = _cabi._DTYPES[dtype]
    GUARD = 4096

    def guarded(nbytes):
        buf = torch.full((nbytes + 2 * GUARD,), 0xA5, dtype=torch.uint8, device=DEV)
        return buf, buf.data_ptr() + GUARD

    _cabi.set_tuning(bwd_sorted=2)
    try:
        ws_bytes = lib.msda_backward_workspace_bytes(ctypes.byref(dims), code)
        assert ws_bytes > 0
        bufs = {name: guarded(n) for name, n in (('gv', value.numel() * value.element_size()), ('gl', g['loc'].numel() * 4),
                                                 ('ga', g['aw'].numel() * 4), ('ws', ws_bytes))}
        rc = lib.msda_backward(ctypes.byref(dims), code, value.data_ptr(), g['shapes'].data_ptr(), g['lsi'].data_ptr(),
                               g['loc'].data_ptr(), g['aw'].data_ptr(), go.data_ptr(), bufs['gv'][1], bufs['gl'][1], bufs['ga'][1],
                               bufs['ws'][1], ws_bytes, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.msda_last_error()
        torch.cuda.synchronize()
    finally:
        _cabi.set_tuning(bwd_sorted=0)
    for name, (buf, _) in bufs.items():
        assert bool((buf[:GUARD] == 0xA5).all()) and bool((buf[-GUARD:] == 0xA5).all()), 'guard band of %s was written' % name
    # and the guarded outputs are the right answer
    gv = bufs['gv'][0][GUARD:-GUARD].view(dtype).view_as(value).float().cpu()
    wgv, _, _ = c_oracle.backward(value.float().cpu(), inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], go.float().cpu())
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    torch.testing.assert_close(gv, wgv, rtol=tol, atol=tol * _scale(wgv))
