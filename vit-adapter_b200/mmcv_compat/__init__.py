"""mmcv-API stand-in for the deformable encoder of the Mask2Former pixel decoder."""
from .multi_scale_deform_attn import MultiScaleDeformableAttention, MultiScaleDeformableAttnFunction

__all__ = ['MultiScaleDeformableAttention', 'MultiScaleDeformableAttnFunction']
