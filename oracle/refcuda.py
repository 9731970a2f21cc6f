"""ctypes wrapper of oracle/_ref/libmsda_refcuda.so — the reference's own CUDA kernels. GPU only.

TEST INFRASTRUCTURE ONLY. The .so is compiled (oracle/Makefile `ref`) from the reference sources where
they lie under /root/reference; it is git-ignored and travels to the GPU box with the snapshot.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_ref', 'libmsda_refcuda.so')
REFERENCE = '/root/reference'
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def build(force=False):
    """Compile the reference kernels if /root/reference is present (no-op elsewhere)."""
    if not os.path.isdir(os.path.join(REFERENCE, 'detection', 'ops', 'src', 'cuda')):
        return LIB_PATH if available() else None
    if force or not available():
        r = subprocess.run(['make', '-C', _HERE, 'ref'] + (['-B'] if force else []),
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError('building the reference CUDA kernels failed:\n' + r.stdout[-4000:])
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError('%s missing (built only where /root/reference exists)' % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _dims(value, shapes, loc):
    N, S, M, D = value.shape
    return N, S, M, D, shapes.shape[0], loc.shape[1], loc.shape[4]


def forward(value, shapes, lsi, loc, aw, im2col_step=64):
    assert value.is_cuda and value.dtype in (torch.float32, torch.float64)
    sfx = 'f32' if value.dtype == torch.float32 else 'f64'
    N, S, M, D, L, Lq, P = _dims(value, shapes, loc)
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    rc = getattr(load(), 'refcuda_forward_' + sfx)(
        _p(value), _p(shapes), _p(lsi), _p(loc), _p(aw), N, S, M, D, L, Lq, P, int(im2col_step), _p(out),
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError('refcuda_forward rc=%d' % rc)
    return out


def backward(value, shapes, lsi, loc, aw, grad_out, im2col_step=64):
    assert value.is_cuda and value.dtype in (torch.float32, torch.float64)
    sfx = 'f32' if value.dtype == torch.float32 else 'f64'
    N, S, M, D, L, Lq, P = _dims(value, shapes, loc)
    gv, gl, ga = torch.empty_like(value), torch.empty_like(loc), torch.empty_like(aw)
    rc = getattr(load(), 'refcuda_backward_' + sfx)(
        _p(value), _p(shapes), _p(lsi), _p(loc), _p(aw), _p(grad_out.contiguous()), N, S, M, D, L, Lq, P,
        int(im2col_step), _p(gv), _p(gl), _p(ga),
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError('refcuda_backward rc=%d' % rc)
    return gv, gl, ga
