"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/msda_b200.h
declares, validates arguments without touching a GPU, and the product path has no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

import vit_adapter_b200 as vab
from vit_adapter_b200 import _cabi


def _declared_functions():
    text = open(os.path.join(ROOT, 'include', 'msda_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b((?:msda|adapter)_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    names = _declared_functions()
    assert set(names) == set(_cabi.EXPORTS), (names, _cabi.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.msda_abi_version() == 1


def test_library_has_no_torch_dependency():
    import subprocess
    out = subprocess.run(['ldd', _cabi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert 'libtorch' not in out and 'libc10' not in out and 'libpython' not in out and 'libcudart' not in out


def test_argument_validation_without_gpu():
    lib = _cabi.load()
    d = _cabi.MsdaDims(2, 84, 3, 8, 3, 16, 4)
    assert lib.msda_forward(None, 0, None, None, None, None, None, None, None) == -1
    assert lib.msda_forward(ctypes.byref(d), 0, None, None, None, None, None, None, None) == -1  # NULL tensors
    assert b'NULL' in lib.msda_last_error()
    assert lib.msda_forward(ctypes.byref(d), 9, None, None, None, None, None, None, None) == -3  # dtype
    bad = _cabi.MsdaDims(2, 84, 3, 0, 3, 16, 4)
    assert lib.msda_forward(ctypes.byref(bad), 0, None, None, None, None, None, None, None) == -2
    many = _cabi.MsdaDims(2, 84, 3, 8, 17, 16, 4)
    assert lib.msda_forward(ctypes.byref(many), 0, None, None, None, None, None, None, None) == -5
    # workspace: only bf16 needs one (fp32 accumulator of grad_value)
    assert lib.msda_backward_workspace_bytes(ctypes.byref(d), _cabi.MSDA_F32) == 0
    assert lib.msda_backward_workspace_bytes(ctypes.byref(d), _cabi.MSDA_BF16) == 2 * 84 * 3 * 8 * 4
    # im2col_step precondition of the reference (ms_deform_attn_cuda.cu:50-52)
    assert lib.msda_check_im2col_step(16, 64) == 0
    assert lib.msda_check_im2col_step(128, 64) == 0
    assert lib.msda_check_im2col_step(96, 64) == -7
    assert b'must divide' in lib.msda_last_error()


def test_cpu_tensors_are_rejected_like_the_reference():
    """ms_deform_attn.h:38 — AT_ERROR("Not implemented on the CPU"); there is no CPU fallback here either."""
    value = torch.rand(1, 30, 2, 4)
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    lsi = torch.as_tensor([0, 24], dtype=torch.long)
    loc = torch.rand(1, 2, 2, 2, 2, 2)
    aw = torch.rand(1, 2, 2, 2, 2)
    with pytest.raises(RuntimeError, match='CPU'):
        vab.MSDeformAttnFunction.apply(value, shapes, lsi, loc, aw, 2)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, '_lib', None)
    monkeypatch.setattr(_cabi, 'LIB_PATH', '/nonexistent/libmsda_b200.so')
    with pytest.raises(RuntimeError, match='no CPU'):
        _cabi.load()


def test_library_override_from_the_environment_fails_loudly_too():
    """MSDA_B200_LIB (A/B timing of kernel variants, tools/walker_variants.sh) names another build of the library; a path that
    does not exist is an error, not a silent return to the default build."""
    import subprocess
    import sys
    env = dict(os.environ, MSDA_B200_LIB='/nonexistent/libmsda_b200.so')
    code = ('import sys; sys.path.insert(0, %r)\nimport vit_adapter_b200 as vab\n'
            'try:\n    vab._cabi.load()\nexcept RuntimeError as e:\n    print("RAISED", e)\n' % ROOT)
    out = subprocess.run([sys.executable, '-c', code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300).stdout
    assert 'RAISED' in out and '/nonexistent/libmsda_b200.so' in out, out


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under vit-adapter_b200/ or include/ may reference it."""
    pkg = os.path.join(ROOT, 'vit-adapter_b200')
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                text = open(os.path.join(base, f), errors='ignore').read()
                if re.search(r'^\s*(from|import)\s+oracle|oracle[./]|grid_sample', text, flags=re.M):
                    offenders.append(os.path.join(base, f))
    assert not offenders, offenders


def test_pybind_compat_module_has_reference_surface():
    """The stand-in for the reference's pybind module exposes exactly vision.cpp:13-16's two functions."""
    import importlib
    import inspect
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'vit-adapter_b200', 'pybind_compat'))
    try:
        sys.modules.pop('MultiScaleDeformableAttention', None)
        m = importlib.import_module('MultiScaleDeformableAttention')
    finally:
        sys.path.pop(0)
    assert list(inspect.signature(m.ms_deform_attn_forward).parameters) == [
        'value', 'spatial_shapes', 'level_start_index', 'sampling_loc', 'attn_weight', 'im2col_step']
    assert list(inspect.signature(m.ms_deform_attn_backward).parameters) == [
        'value', 'spatial_shapes', 'level_start_index', 'sampling_loc', 'attn_weight', 'grad_output', 'im2col_step']
    with pytest.raises(RuntimeError, match='CPU'):
        m.ms_deform_attn_forward(torch.rand(1, 4, 1, 4), torch.as_tensor([(2, 2)]), torch.zeros(1, dtype=torch.long),
                                 torch.rand(1, 1, 1, 1, 1, 2), torch.rand(1, 1, 1, 1, 1), 64)
    sys.modules.pop('MultiScaleDeformableAttention', None)


def test_backward_workspace_sizes_follow_the_kernel_selection():
    """msda_backward_workspace_bytes() needs no GPU: 16-bit dtypes ask for the fp32 accumulator, calls that take the slab-sorted
    backward (one level, >= 16 samples per value token and head, large enough) add its sort buffers, and the tuning key
    changes the answer."""
    import ctypes
    from vit_adapter_b200 import _cabi
    lib = _cabi.load()
    F32, BF16 = 0, 1
    assert (_cabi._DTYPES[__import__('torch').float32], _cabi._DTYPES[__import__('torch').bfloat16]) == (F32, BF16)

    def ws(N, S, M, D, L, Lq, P, dtype):
        d = _cabi.MsdaDims(N, S, M, D, L, Lq, P)
        return lib.msda_backward_workspace_bytes(ctypes.byref(d), dtype)

    extractor_b = (16, 1024, 12, 32, 1, 5376, 4)      # ViT-Adapter-B Extractor, 16 images: 4.1 M samples, 21 per token and head
    injector_b = (16, 5376, 12, 32, 3, 1024, 4)       # 2.3 samples per token and head
    small64 = (1, 256, 6, 64, 1, 1344, 4)             # dense but tiny
    accum = lambda N, S, M, D, *_: N * S * M * D * 4
    try:
        sorted_bytes = ws(*extractor_b, F32)
        assert sorted_bytes >= 16 * 12 * 5376 * 4 * 4                  # at least one 4-byte index per sample
        assert ws(*extractor_b, BF16) == accum(*extractor_b) + sorted_bytes
        assert ws(*injector_b, F32) == 0 and ws(*injector_b, BF16) == accum(*injector_b)
        assert ws(*small64, F32) == 0
        _cabi.set_tuning(bwd_sorted=1)
        assert ws(*extractor_b, F32) == 0 and ws(*extractor_b, BF16) == accum(*extractor_b)
        _cabi.set_tuning(bwd_sorted=2)
        assert ws(*small64, F32) > 0 and ws(*injector_b, F32) > 0
        assert ws(1, 256, 6, 48, 1, 1344, 4, F32) == 0                 # no sorted kernel for this head width
    finally:
        _cabi.set_tuning(bwd_sorted=0)


def test_traffic_file_covers_the_bench_line_and_the_north_star_block():
    """bench.py reads `roofline.traffic` (ncu DRAM bytes per launch) from profiles/traffic.json under
    <variant>_<dtype>_<call>_<fwd|bwd>: the headline workload and both north-star configurations must be there."""
    import json
    t = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
    for variant, dtype in (('B', 'f32'), ('L', 'bf16'), ('L64', 'bf16')):
        for call in ('injector', 'extractor'):
            for d in ('fwd', 'bwd'):
                key = '%s_%s_%s_%s' % (variant, dtype, call, d)
                assert key in t and float(t[key]) > 1e6, key
