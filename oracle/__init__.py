"""oracle/ — CHECKERS for the multi-scale deformable attention hot path. TEST INFRASTRUCTURE ONLY.

Nothing under vit-adapter_b200/ (the product) imports this package. Allowed importers: tests/,
__graft_entry__.smoke() and bench.py's `cpu_baseline` / `--impl reference` legs — as the checker or
the reported CPU baseline, never as the thing shipped.

Three independent checkers:
  c_oracle      msda_oracle.c  — plain-C scalar restatement of the reference CUDA kernels
                (ms_deform_im2col_cuda.cuh:33-159, 237-510); f32 + f64; forward, backward, indices.
  core_pytorch  core_pytorch.py — restatement of the reference's CPU path ms_deform_attn_core_pytorch
                (ms_deform_attn_func.py:49-71; F.grid_sample). This is what `--impl reference` times.
  refcuda       _ref/libmsda_refcuda.so — the reference's OWN CUDA kernels compiled from
                /root/reference for sm_100a (ref_cuda_wrap.cu); GPU only.

Parity pin: tests/test_oracle_golden.py checks c_oracle and core_pytorch against golden vectors made
by tests/golden/make_golden.py, which imports the real reference Python from /root/reference.
"""
from . import c_oracle, core_pytorch, dwconv_ref, layernorm_ref, mmcv_msda_ref, refcuda  # noqa: F401
