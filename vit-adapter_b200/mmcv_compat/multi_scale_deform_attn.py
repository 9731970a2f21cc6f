"""`MultiScaleDeformableAttention` with mmcv's calling convention, on the B200 kernels.

The reference's Mask2Former pixel decoder builds its 6-layer deformable encoder from mmcv's attention module
(mmcv-full==1.4.2, un-vendored: segmentation/README.md:25; call sites
segmentation/mmseg_custom/models/utils/transformer.py:28-34 and
segmentation/mmseg_custom/models/plugins/msdeformattn_pixel_decoder.py:53,157,231-242; config
segmentation/configs/ade20k/mask2former_beit_adapter_large_896_80k_ade20k_ss.py:50-61: embed_dims=1024, num_heads=32,
num_levels=3, num_points=4, im2col_step=64, dropout=0.0, batch_first=False). mmcv is not vendored in the reference,
so this follows mmcv's published interface: the constructor keywords, the sub-module names (`sampling_offsets`,
`attention_weights`, `value_proj`, `output_proj` - checkpoints load unchanged), the initialisation, and the forward
contract
    forward(query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
            reference_points=None, spatial_shapes=None, level_start_index=None) -> dropout(attn) + identity
with (num_query, bs, embed_dims) tensors unless batch_first. The arithmetic between the linears is the same
Deformable-DETR math as ops/modules/ms_deform_attn.py:102-129, so the module reuses MSDeformAttn's forward (fused
softmax + location arithmetic included). Unlike mmcv there is no pure-PyTorch fallback on CPU tensors: the product
path is the CUDA library or an error (same policy as the rest of the package).

Parity: unpinned against mmcv itself (absent here and untested in the reference); the tests
compare it with a CPU restatement of the published forward built on the reference's own pure-torch sampling core
(tests/test_mmcv_compat.py).
"""
import warnings

from torch import nn

from ..functions import MSDeformAttnFunction
from ..modules import MSDeformAttn


# mmcv's op: apply(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
# im2col_step) - the same signature as the reference's MSDeformAttnFunction, which does the work.
MultiScaleDeformableAttnFunction = MSDeformAttnFunction


class MultiScaleDeformableAttention(MSDeformAttn):
    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64, dropout=0.1,
                 batch_first=False, norm_cfg=None, init_cfg=None):
        if embed_dims % num_heads != 0:
            raise ValueError(f'embed_dims must be divisible by num_heads, but got {embed_dims} and {num_heads}')
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')  # mmcv words the power-of-2 warning itself (below)
            super().__init__(d_model=embed_dims, n_levels=num_levels, n_heads=num_heads, n_points=num_points, ratio=1.0)
        dim_per_head = embed_dims // num_heads
        if dim_per_head & (dim_per_head - 1):
            warnings.warn("You'd better set embed_dims in MultiScaleDeformAttention to make the dimension of each "
                          'attention head a power of 2 which is more efficient in our CUDA implementation.')
        self.norm_cfg = norm_cfg
        self.init_cfg = init_cfg
        self.dropout = nn.Dropout(dropout)
        self.batch_first = batch_first
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self._is_init = True

    def init_weights(self):
        """Default initialisation: the ring of directions for the offset bias, zero attention logits, xavier for the
        two projections - the same scheme as MSDeformAttn._reset_parameters (ops/modules/ms_deform_attn.py:64-81)."""
        self._reset_parameters()
        self._is_init = True

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:  # (num_query, bs, embed_dims) -> (bs, num_query, embed_dims)
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError(f'Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.')
        output = MSDeformAttn.forward(self, query, reference_points, value, spatial_shapes, level_start_index, key_padding_mask)
        if not self.batch_first:
            output = output.permute(1, 0, 2)
        return self.dropout(output) + identity
