"""Drop-in modules registered under the reference's `@register_module` names
(projects/mmdet3d_plugin/__init__.py imports the same set by side effect)."""
from .voxel import Voxelization, HardSimpleVFE, DynamicSimpleVFE  # noqa: F401
from .sparse_encoder_hd import (SparseEncoderHD, SparseBasicBlock, SubMConv3d,  # noqa: F401
                                SparseConv3d, make_sparse_convmodule)
from .second_3d import SECOND3D, SECOND3DFPN  # noqa: F401
from .transformer import (Uni3DETRTransformer, Uni3DETRTransformerDecoder,  # noqa: F401
                          UniCrossAtten, BaseTransformerLayer, MultiheadAttention, FFN, MLP,
                          inverse_sigmoid)
from .head import Uni3DETRHead, NMSFreeCoder, denormalize_bbox  # noqa: F401
from .detector import Uni3DETR  # noqa: F401
from ..compat import register_with_openmmlab

register_with_openmmlab()
