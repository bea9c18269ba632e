"""Parameter / buffer names and shapes of the reference's own modules (checkpoint compatibility, SURVEY.md
Appendix B).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_keys.py
Output (committed): tests/golden/golden_state_dict_keys.json

The reference's first-party classes are instantiated from /root/reference (third-party builders stubbed by
make_golden.py's machinery: mmcv `build_conv_layer` / `build_norm_layer` / `BaseTransformerLayer` wiring are
[restated], so the names THEY contribute - e.g. `attentions.0.attn.in_proj_weight`, `ffns.0.layers.0.0` - follow
mmcv 1.x as restated there) and their `state_dict()` keys and shapes are stored. tests/test_abi.py checks that
the drop-in modules expose exactly the same names and shapes, so reference `.pth` files load with strict=True.
"""
import json
import os
import sys

import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402

LAYER = dict(type="BaseTransformerLayer",
             attn_cfgs=[dict(type="MultiheadAttention", embed_dims=256, num_heads=8, dropout=0.1),
                        dict(type="UniCrossAtten", num_points=1, embed_dims=256, num_sweeps=1)],
             ffn_cfgs=dict(type="FFN", embed_dims=256, feedforward_channels=64, num_fcs=2, ffn_drop=0.1,
                           act_cfg=dict(type="ReLU", inplace=True)),
             norm_cfg=dict(type="LN"), operation_order=("self_attn", "norm", "cross_attn", "norm", "ffn", "norm"))
BCFG = dict(in_channels=[8, 8, 8], out_channels=[4, 8, 16], layer_nums=[2, 2, 2], layer_strides=[1, 2, 4],
            is_cascade=False, norm_cfg=dict(type="BN3d", eps=1e-3, momentum=0.01))
NCFG = dict(in_channels=[4, 8, 16], out_channels=[8, 8, 8], upsample_strides=[1, 2, 4],
            norm_cfg=dict(type="BN3d", eps=1e-3, momentum=0.01), upsample_cfg=dict(type="deconv3d", bias=False),
            extra_conv=dict(type="Conv3d", num_conv=2, bias=False), use_conv_for_no_stride=True)


def keys(m):
    return {k: list(v.shape) for k, v in m.state_dict().items()}


def main():
    MG.install_stubs()
    MG.load_ref("projects/mmdet3d_plugin/core/bbox/util.py", "projects.mmdet3d_plugin.core.bbox.util")
    tr = MG.load_ref("projects/mmdet3d_plugin/models/utils/uni3detr_transformer.py", "ref_transformer_k")
    head = MG.load_ref("projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py", "ref_head_k")
    bb = MG.load_ref("projects/mmdet3d_plugin/models/backbones/second_3d.py", "ref_second3d_k")
    nk = MG.load_ref("projects/mmdet3d_plugin/models/necks/second3d_fpn.py", "ref_fpn_k")
    out = {}
    out["UniCrossAtten"] = keys(tr.UniCrossAtten(embed_dims=256, num_heads=8, num_points=1, dropout=0.1))
    transformer = tr.Uni3DETRTransformer(decoder=dict(type="Uni3DETRTransformerDecoder", num_layers=2,
                                                      return_intermediate=True, transformerlayers=LAYER))
    out["Uni3DETRTransformer"] = keys(transformer)

    class HeadShell(nn.Module):
        pass
    hs = HeadShell()
    hs.num_query, hs.with_box_refine, hs.as_two_stage = 5, True, False
    hs.transformer = transformer
    hs.embed_dims, hs.num_reg_fcs, hs.cls_out_channels, hs.code_size = 256, 2, 3, 8
    head.Uni3DETRHead._init_layers(hs)
    out["Uni3DETRHead"] = keys(hs)
    out["SECOND3D"] = keys(bb.SECOND3D(conv_cfg=MG.AttrDict(type="Conv3d", kernel=(1, 3, 3), bias=False), **BCFG))
    out["SECOND3DFPN"] = keys(nk.SECOND3DFPN(**NCFG))
    with open(os.path.join(HERE, "golden_state_dict_keys.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    for k, v in out.items():
        print(k, len(v), "entries")


if __name__ == "__main__":
    main()
