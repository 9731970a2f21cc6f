"""The C ABI from plain C: tests/cabi_client/kat_client.c includes only include/msda_b200.h (+ the CUDA runtime API for device
memory), is compiled as C11 with gcc, links the library and reproduces the reference's golden vectors on the GPU."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import OP_CASES, ROOT, load_golden

from vit_adapter_b200 import _cabi

SRC = os.path.join(ROOT, 'tests', 'cabi_client', 'kat_client.c')
CUDA = os.environ.get('CUDA_HOME', '/usr/local/cuda')


def _build(outdir):
    exe = os.path.join(outdir, 'kat_client')
    libdir = os.path.dirname(_cabi.LIB_PATH)
    cmd = ['gcc', '-std=c11', '-O1', '-Wall', '-Wextra', '-Werror', '-I' + os.path.join(ROOT, 'include'), '-I' + os.path.join(CUDA, 'include'),
           SRC, '-o', exe, '-L' + libdir, '-lmsda_b200', '-L' + os.path.join(CUDA, 'lib64'), '-lcudart',
           '-Wl,-rpath,' + libdir, '-Wl,-rpath,' + os.path.join(CUDA, 'lib64')]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def test_header_and_library_are_usable_from_c11():
    _cabi.load()
    with tempfile.TemporaryDirectory() as d:
        exe = _build(d)
        assert os.path.exists(exe)
        r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)   # no argument: usage, no GPU touched
        assert r.returncode == 1 and 'usage' in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize('case', [c for c in OP_CASES if c != 'op_odd_d5'] + ['op_odd_d5'])
def test_c_client_reproduces_golden(case):
    g = load_golden(case)
    value, loc, aw = g['value'].float(), g['loc'].float(), g['aw'].float()
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    with tempfile.TemporaryDirectory() as d:
        exe = _build(d)
        open(os.path.join(d, 'meta.txt'), 'w').write('%d %d %d %d %d %d %d\n' % (N, S, M, D, L, Lq, P))
        value.numpy().astype('<f4').tofile(os.path.join(d, 'value.f32'))
        loc.numpy().astype('<f4').tofile(os.path.join(d, 'loc.f32'))
        aw.numpy().astype('<f4').tofile(os.path.join(d, 'aw.f32'))
        g['grad_out'].float().numpy().astype('<f4').tofile(os.path.join(d, 'grad_out.f32'))
        g['shapes'].numpy().astype('<i8').tofile(os.path.join(d, 'shapes.i64'))
        g['lsi'].numpy().astype('<i8').tofile(os.path.join(d, 'lsi.i64'))
        r = subprocess.run([exe, d], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout.startswith('ok launches=')
        out = np.fromfile(os.path.join(d, 'out.f32'), dtype='<f4').reshape(N, Lq, M * D)
        gv = np.fromfile(os.path.join(d, 'grad_value.f32'), dtype='<f4').reshape(N, S, M, D)
        gl = np.fromfile(os.path.join(d, 'grad_loc.f32'), dtype='<f4').reshape(N, Lq, M, L, P, 2)
        ga = np.fromfile(os.path.join(d, 'grad_aw.f32'), dtype='<f4').reshape(N, Lq, M, L, P)
    # same comparisons as tests/test_op_gpu.py::test_golden_f32: forward and the smooth gradients against the reference's fp64
    # goldens, grad_loc (discontinuous at texel centres, which the 'edges' cases hit exactly) against the fp32 C oracle
    from oracle import c_oracle
    _, ogl, _ = c_oracle.backward(g['value'], g['shapes'], g['lsi'], g['loc'], g['aw'], g['grad_out'])
    for got, want, tol in ((out, g['out_f64'], 1e-5), (gv, g['grad_value_f64'], 1e-4), (ga, g['grad_aw_f64'], 1e-4), (gl, ogl, 1e-4)):
        want = want.double().numpy()
        scale = float(np.abs(want).max()) + 1e-30
        assert np.abs(got.astype(np.float64) - want).max() <= tol * scale + 1e-7, (np.abs(got - want).max(), scale)
