from .ms_deform_attn_func import (MSDeformAttnFunction, MSDeformAttnFusedFunction, MSDeformAttnMergedFunction,  # noqa: F401
                                   set_amp_value_dtype)
from .linear import linear  # noqa: F401
