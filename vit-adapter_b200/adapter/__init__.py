from .adapter_modules import (ConvFFN, DropPath, DWConv, Extractor, Injector, InteractionBlock,  # noqa: F401
                              InteractionBlockWithCls, InteractionBlockWithText, SpatialPriorModule, deform_inputs, get_reference_points)
from .sync_batchnorm import SyncBatchNormNoHostSync  # noqa: F401
