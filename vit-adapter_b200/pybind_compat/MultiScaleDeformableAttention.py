"""`MultiScaleDeformableAttention` — stand-in for the reference's compiled pybind11 extension.

The reference's Python (detection/ops/functions/ms_deform_attn_func.py:11,25-30,41-44) does
    import MultiScaleDeformableAttention as MSDA
    MSDA.ms_deform_attn_forward(value, shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    MSDA.ms_deform_attn_backward(value, shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step)
(exported by detection/ops/src/vision.cpp:13-16). Putting THIS directory on sys.path ahead of the
reference's build makes that unmodified Python run on the sm_100a kernels: same two names, same
argument order, same return types (Tensor; list of 3 Tensors), same errors (RuntimeError for CPU or
non-contiguous tensors). Everything goes through the C ABI of include/msda_b200.h.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from vit_adapter_b200 import _cabi  # noqa: E402


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    return _cabi.forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    return list(_cabi.backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                               im2col_step))
