from .ms_deform_attn_func import MSDeformAttnFunction, set_amp_value_dtype  # noqa: F401
