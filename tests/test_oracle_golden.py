"""Pin oracle/model.py to golden vectors produced by executing the REFERENCE's own first-party
source files (tests/golden/make_golden.py, third-party imports stubbed). CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from gen_weights import gen_state_dict
from oracle import geometry as G
from oracle import model as M

T = torch.from_numpy


def close(a, b, rtol=1e-5, atol=1e-5):
    a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
    np.testing.assert_allclose(a, np.asarray(b), rtol=rtol, atol=atol)


def test_sine_embed(golden):
    # utils/uni3detr_transformer.py:33-65
    close(M.get_sine_pos_embed(T(golden["sine_in"])), golden["sine_out"], 1e-5, 1e-6)


def test_shift_scale_points(golden):
    # detectors/uni3detr.py:18-46 as used at :181/:187
    for b in range(golden["ssp_in"].shape[0]):
        close(G.shift_scale_unit(golden["ssp_in"][b]), golden["ssp_out"][b], 1e-6, 1e-6)


def test_denormalize_bbox(golden):
    # core/bbox/util.py:44-80
    close(M.denormalize_bbox(T(golden["denorm_in"])), golden["denorm_out"], 1e-6, 1e-6)
    close(M.denormalize_bbox(T(golden["denorm10_in"])), golden["denorm10_out"], 1e-6, 1e-6)


def ca_state_dict():
    shapes = {"attention_weights.weight": (1, 256), "attention_weights.bias": (1,),
              "output_proj.weight": (256, 256), "output_proj.bias": (256,),
              "position_encoder.0.weight": (256, 3), "position_encoder.0.bias": (256,),
              "position_encoder.1.weight": (256,), "position_encoder.1.bias": (256,),
              "position_encoder.3.weight": (256, 256), "position_encoder.3.bias": (256,),
              "position_encoder.4.weight": (256,), "position_encoder.4.bias": (256,)}
    return gen_state_dict(shapes, 11)


def test_uni_cross_atten(golden):
    # utils/uni3detr_transformer.py:271-360 executed from the reference file
    sd = ca_state_dict()
    out = M.uni_cross_atten(sd, "", T(golden["ca_query"]), T(golden["ca_value"])[:, 0],
                            T(golden["ca_qpos"]), T(golden["ca_ref"]))
    close(out, golden["ca_out"], 1e-4, 1e-5)


def head_shapes(nq, ncls, code, L, ffn=64, E=256):
    s = {"tgt_embed.weight": (2 * nq, E), "refpoint_embed.weight": (nq, 3)}
    for l in range(L):
        for i, (o, n) in zip((0, 3, 6), ((E, E), (E, E), (ncls, E))):
            s[f"cls_branches.{l}.{i}.weight"], s[f"cls_branches.{l}.{i}.bias"] = (o, n), (o,)
        for i in (1, 4):
            s[f"cls_branches.{l}.{i}.weight"], s[f"cls_branches.{l}.{i}.bias"] = (E,), (E,)
        for name, last in (("reg_branches", code), ("iou_branches", 1)):
            for i, (o, n) in zip((0, 2, 4), ((E, E), (E, E), (last, E))):
                s[f"{name}.{l}.{i}.weight"], s[f"{name}.{l}.{i}.bias"] = (o, n), (o,)
        p = f"transformer.decoder.layers.{l}."
        s[p + "attentions.0.attn.in_proj_weight"], s[p + "attentions.0.attn.in_proj_bias"] = (3 * E, E), (3 * E,)
        s[p + "attentions.0.attn.out_proj.weight"], s[p + "attentions.0.attn.out_proj.bias"] = (E, E), (E,)
        a = p + "attentions.1."
        s[a + "attention_weights.weight"], s[a + "attention_weights.bias"] = (1, E), (1,)
        s[a + "output_proj.weight"], s[a + "output_proj.bias"] = (E, E), (E,)
        s[a + "position_encoder.0.weight"], s[a + "position_encoder.0.bias"] = (E, 3), (E,)
        s[a + "position_encoder.3.weight"], s[a + "position_encoder.3.bias"] = (E, E), (E,)
        for i in (1, 4):
            s[a + f"position_encoder.{i}.weight"], s[a + f"position_encoder.{i}.bias"] = (E,), (E,)
        s[p + "ffns.0.layers.0.0.weight"], s[p + "ffns.0.layers.0.0.bias"] = (ffn, E), (ffn,)
        s[p + "ffns.0.layers.1.weight"], s[p + "ffns.0.layers.1.bias"] = (E, ffn), (E,)
        for i in range(3):
            s[p + f"norms.{i}.weight"], s[p + f"norms.{i}.bias"] = (E,), (E,)
    for mlp, cin in (("query_scale", E), ("ref_point_head", 384)):
        for i in range(3):
            s[f"transformer.decoder.{mlp}.layers.{i}.weight"] = (E, cin if i == 0 else E)
            s[f"transformer.decoder.{mlp}.layers.{i}.bias"] = (E,)
    return s


def golden_head_cfg(golden):
    nq, ncls, code, L = [int(v) for v in golden["head_meta"]]
    cfg = dict(num_query=nq, transformer=dict(decoder=dict(num_layers=L)),
               bbox_coder=dict(pc_range=[-3.2, -0.2, -2., 3.2, 6.2, 0.56]))
    sd = gen_state_dict(head_shapes(nq, ncls, code, L), 21)
    return cfg, sd, (nq, ncls, code, L)


def test_head_forward(golden):
    # dense_heads/uni3detr_head.py:422-508 + uni3detr_transformer.py:95-212 from the reference files
    cfg, sd, _ = golden_head_cfg(golden)
    outs = M.head_forward(sd, cfg, T(golden["head_feats"]), T(golden["head_fps"]),
                          T(golden["head_rand"]), prefix="")
    for k in ("all_cls_scores", "all_bbox_preds", "all_iou_preds"):
        close(outs[k], golden["head_" + k], 2e-4, 2e-4)


def test_nms_free_coder(golden):
    # core/bbox/coders/nms_free_coder.py:42-136
    outs = {k: T(golden["head_" + k]) for k in ("all_cls_scores", "all_bbox_preds", "all_iou_preds")}
    pc = [-3.2, -0.2, -2., 3.2, 6.2, 0.56]
    res = M.nms_free_decode(outs, dict(num_classes=3, alpha=0.2, post_center_range=pc, max_num=12))
    for i, r in enumerate(res):
        for k in ("bboxes", "scores", "ious"):
            close(r[k], golden[f"coder_{i}_{k}"], 1e-5, 1e-6)
        np.testing.assert_array_equal(r["labels"].numpy(), golden[f"coder_{i}_labels"])


def dense_state_dicts():
    bshapes, nshapes = {}, {}
    ins, outs = [8, 8, 8], [4, 8, 16]
    for i in range(3):
        for j in range(3):
            cin = ins[i] if j == 0 else outs[i]
            bshapes[f"blocks.{i}.{3 * j}.weight"] = (outs[i], cin, 1, 3, 3)
            for n in ("weight", "bias", "running_mean", "running_var"):
                bshapes[f"blocks.{i}.{3 * j + 1}.{n}"] = (outs[i],)
            bshapes[f"blocks.{i}.{3 * j + 1}.num_batches_tracked"] = ()
    for i, s in enumerate([1, 2, 4]):
        nshapes[f"deblocks.{i}.0.weight"] = (8, outs[i], 1, 1, 1) if s == 1 else (outs[i], 8, 1, s, s)
        for n in ("weight", "bias", "running_mean", "running_var"):
            nshapes[f"deblocks.{i}.1.{n}"] = (8,)
        nshapes[f"deblocks.{i}.1.num_batches_tracked"] = ()
    for j in range(2):
        nshapes[f"extra_blocks.{3 * j}.weight"] = (8, 8, 3, 3, 3)
        for n in ("weight", "bias", "running_mean", "running_var"):
            nshapes[f"extra_blocks.{3 * j + 1}.{n}"] = (8,)
        nshapes[f"extra_blocks.{3 * j + 1}.num_batches_tracked"] = ()
    return gen_state_dict(bshapes, 31), gen_state_dict(nshapes, 41)


DENSE_BCFG = dict(in_channels=[8, 8, 8], out_channels=[4, 8, 16], layer_nums=[2, 2, 2],
                  layer_strides=[1, 2, 4], is_cascade=False, norm_cfg=dict(type="BN3d", eps=1e-3),
                  conv_cfg=dict(type="Conv3d", kernel=(1, 3, 3), bias=False))
DENSE_NCFG = dict(in_channels=[4, 8, 16], out_channels=[8, 8, 8], upsample_strides=[1, 2, 4],
                  norm_cfg=dict(type="BN3d", eps=1e-3), upsample_cfg=dict(type="deconv3d", bias=False),
                  extra_conv=dict(type="Conv3d", num_conv=2, bias=False), use_conv_for_no_stride=True)


def test_dense_cnn(golden):
    # backbones/second_3d.py:89-114 + necks/second3d_fpn.py:112-143
    bsd, nsd = dense_state_dicts()
    xs = M.second3d(bsd, DENSE_BCFG, T(golden["dense_in"]), prefix="")
    for i, x in enumerate(xs):
        close(x, golden[f"dense_bb{i}"], 1e-4, 1e-5)
    y = M.second3dfpn(nsd, DENSE_NCFG, xs, prefix="")
    close(y, golden["dense_out"], 1e-4, 1e-5)


def test_encoder_layer_construction(model_cfgs):
    # pts_encoder/sparse_encoder_hd.py:71-104,140-214 run from the reference with recording stubs
    with open(os.path.join(GOLDEN, "golden_encoder_layers.json")) as f:
        ref = json.load(f)
    for name, mc in model_cfgs.items():
        ours = M.encoder_layer_list(mc["pts_middle_encoder"])
        calls = ref[name]
        assert len(ours) == len(calls), name
        for o, c in zip(ours, calls):
            if c["kind"] == "block":
                assert o["kind"] == "block" and o["cin"] == c["cin"] and o["cout"] == c["cout"]
                continue
            assert (o["cin"], o["cout"]) == (c["cin"], c["cout"]), (name, o, c)
            if o["kind"] == "down":
                assert c["conv_type"] == "SparseConv3d"
                assert M._t3(o["stride"]) == M._t3(c["stride"]) and M._t3(o["pad"]) == M._t3(c["padding"])
            elif o["kind"] == "point":
                assert c["conv_type"] == "SparseConv3d" and M._t3(c["k"]) == (1, 1, 1)
            else:
                assert c["conv_type"] == "SubMConv3d"


GB_CASES = [dict(type="nms", nms_thr=0.5), dict(type="nms", nms_thr=0.3, score_thr=0.2),
            dict(type="nms", nms_thr=0.5, score_thr=[0.1, 0.3, 0.2, 0.25], num_thr=15),
            dict(type="nms", nms_thr=0.2, num_thr=20)]


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_get_bboxes_nms_control_flow(case):
    """uni3detr_head.py:826-918 executed from the reference's own file (tests/golden/make_golden_getbboxes.py;
    mmcv nms3d stubbed by the oracle's) vs oracle/postproc.py get_bboxes_nms: bottom-centre shift, per-class
    loop, class-major order, score_thr (scalar / per-class list), num_thr."""
    from oracle import postproc as PP
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_get_bboxes.npz")))
    pp = GB_CASES[case]
    pc = [-3.2, -0.2, -2.0, 3.2, 6.2, 0.56]
    outs = {k: T(g[f"c{case}_{k}"]) for k in ("all_cls_scores", "all_bbox_preds", "all_iou_preds")}
    dec = M.nms_free_decode(outs, dict(num_classes=4, alpha=0.2, post_center_range=pc, max_num=40))
    for i, d in enumerate(dec):
        b, s, l = PP.get_bboxes_nms({k: v.numpy() for k, v in d.items()}, 4, pp)
        np.testing.assert_array_equal(l, g[f"c{case}_s{i}_labels"])
        np.testing.assert_allclose(s, g[f"c{case}_s{i}_scores"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(b, g[f"c{case}_s{i}_bboxes"], rtol=0, atol=1e-5)
