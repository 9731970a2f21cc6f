from .ms_deform_attn_func import MSDeformAttnFunction, MSDeformAttnFusedFunction, set_amp_value_dtype  # noqa: F401
