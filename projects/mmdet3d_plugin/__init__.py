"""`plugin_dir='projects/mmdet3d_plugin/'` entry point.

The reference's train/test scripts import this dotted path by side effect to populate the
registries (extra_tools/train.py:106-127). Here the import registers the sm_100a drop-in
modules under the same names; with mmcv/mmdet/mmdet3d installed they are also mirrored into
the OpenMMLab registries (uni3detr_b200.compat.register_with_openmmlab).
"""
from uni3detr_b200.plugin import (Uni3DETR, SparseEncoderHD, SECOND3D, SECOND3DFPN,  # noqa: F401
                                  Uni3DETRHead, Uni3DETRTransformer,
                                  Uni3DETRTransformerDecoder, UniCrossAtten, NMSFreeCoder)
