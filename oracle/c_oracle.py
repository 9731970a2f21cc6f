"""ctypes wrapper of oracle/msda_oracle.c (CPU, f32/f64). TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_build', 'libmsda_oracle.so')
_lib = None


class _Dims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ('batch', 'spatial_size', 'num_heads', 'channels', 'num_levels', 'num_query', 'num_point')]


def build(force=False):
    src = os.path.join(_HERE, 'msda_oracle.c')
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(LIB_PATH):
        r = subprocess.run(['make', '-C', _HERE, 'oracle'] + (['-B'] if force else []),
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError('building the C oracle failed:\n' + r.stdout)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.msda_oracle_get_threads.restype = ctypes.c_int
    return _lib


def set_threads(n):
    load().msda_oracle_set_threads(int(n))


def get_threads():
    return int(load().msda_oracle_get_threads())


def _prep(value, shapes, lsi, loc, aw):
    dt = value.dtype
    assert dt in (torch.float32, torch.float64), dt
    value = value.detach().cpu().contiguous()
    shapes = shapes.detach().cpu().to(torch.int64).contiguous()
    lsi = lsi.detach().cpu().to(torch.int64).contiguous()
    loc = loc.detach().cpu().to(dt).contiguous()
    aw = aw.detach().cpu().to(dt).contiguous()
    N, S, M, D = value.shape
    L = shapes.shape[0]
    Lq, P = loc.shape[1], loc.shape[4]
    assert loc.shape == (N, Lq, M, L, P, 2) and aw.shape == (N, Lq, M, L, P)
    assert int((shapes[:, 0] * shapes[:, 1]).sum()) == S
    return value, shapes, lsi, loc, aw, _Dims(N, S, M, D, L, Lq, P), ('f32' if dt == torch.float32 else 'f64')


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def forward(value, shapes, lsi, loc, aw):
    """out [N, Lq, M*D] (CPU tensor, dtype of value)."""
    value, shapes, lsi, loc, aw, d, sfx = _prep(value, shapes, lsi, loc, aw)
    out = torch.empty((d.batch, d.num_query, d.num_heads * d.channels), dtype=value.dtype)
    getattr(load(), 'msda_oracle_forward_' + sfx)(ctypes.byref(d), _p(value), _p(shapes), _p(lsi), _p(loc),
                                                  _p(aw), _p(out))
    return out


def backward(value, shapes, lsi, loc, aw, grad_out):
    """(grad_value, grad_loc, grad_aw) CPU tensors."""
    value, shapes, lsi, loc, aw, d, sfx = _prep(value, shapes, lsi, loc, aw)
    grad_out = grad_out.detach().cpu().to(value.dtype).contiguous()
    gv = torch.zeros_like(value)
    gl = torch.empty_like(loc)
    ga = torch.empty_like(aw)
    getattr(load(), 'msda_oracle_backward_' + sfx)(ctypes.byref(d), _p(value), _p(shapes), _p(lsi), _p(loc),
                                                   _p(aw), _p(grad_out), _p(gv), _p(gl), _p(ga))
    return gv, gl, ga


def point_index(shapes, lsi, loc, num_heads, channels):
    """[npoints, 4] int32 (h_low, w_low, corner mask, corner-1 element offset); f32 locations."""
    shapes = shapes.detach().cpu().to(torch.int64).contiguous()
    lsi = lsi.detach().cpu().to(torch.int64).contiguous()
    loc = loc.detach().cpu().to(torch.float32).contiguous()
    N, Lq, M, L, P, _ = loc.shape
    assert M == num_heads
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    d = _Dims(N, S, M, channels, L, Lq, P)
    idx = torch.empty((N * Lq * M * L * P, 4), dtype=torch.int32)
    load().msda_oracle_point_index_f32(ctypes.byref(d), _p(shapes), _p(lsi), _p(loc), _p(idx))
    return idx
