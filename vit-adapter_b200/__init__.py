"""vit-adapter_b200 — B200-native (sm_100a) multi-scale deformable attention behind ViT-Adapter's API.

Layout
  csrc/        hand-written CUDA kernels + the C-ABI (include/msda_b200.h)
  lib/         built libmsda_b200.so (git-ignored; `python build.py`)
  _cabi.py     ctypes binding (raw pointers + stream)
  functions/   MSDeformAttnFunction            (reference: detection/ops/functions)
  modules/     MSDeformAttn                    (reference: detection/ops/modules)
  adapter/     deform_inputs, Injector, Extractor, InteractionBlock[WithCls], ...
               (reference: */mm*_custom/models/backbones/adapter_modules.py)

The directory name is not a Python identifier; import it as `vit_adapter_b200` (alias package at the
repo root) or put this directory on sys.path under the reference's own name `ops`
(`ln -s vit-adapter_b200 ops`), after which `from ops.modules import MSDeformAttn` and
`from ops.functions import MSDeformAttnFunction` work exactly as in the reference.
"""
from . import _cabi  # noqa: F401
from .functions import (MSDeformAttnFunction, MSDeformAttnFusedFunction, MSDeformAttnMergedFunction,  # noqa: F401
                        set_amp_value_dtype)
from .modules import MSDeformAttn  # noqa: F401
from .graphs import GraphedStep  # noqa: F401

__all__ = ['MSDeformAttnFunction', 'MSDeformAttnFusedFunction', 'MSDeformAttnMergedFunction', 'MSDeformAttn', 'set_amp_value_dtype', 'GraphedStep']
