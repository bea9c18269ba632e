"""Generate the golden fixtures that pin oracle/ to the REFERENCE's own source files.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Outputs (committed): tests/golden/configs.json, tests/golden/golden_*.npz

What this does: the reference's first-party files (projects/mmdet3d_plugin/**) are executed
unmodified from /root/reference; only their un-installable third-party imports (mmcv, mmdet,
mmdet3d, spconv - SURVEY.md §8c) are replaced by the stub modules below. Stubs marked
[restated] re-implement published third-party behaviour (mmcv BaseTransformerLayer wiring,
mmdet inverse_sigmoid, mmcv builders) and are therefore NOT a pin for that behaviour; all
arithmetic inside the first-party files (sine embedding, MLPs, decoder loop + reference-point
refinement, UniCrossAtten, group loop, head slicing/sigmoid/range scaling, NMSFreeCoder,
denormalize_bbox, SECOND3D, SECOND3DFPN, shift_scale_points, encoder layer construction)
runs from the reference's own code.

Weights are not stored: they are regenerated from a seed by `gen_state_dict` (shared with the
tests), only inputs/outputs are saved, so the fixtures stay small.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_weights import gen_state_dict  # noqa: E402


# ------------------------------------------------------------------ stubs ------
class _Dummy:
    def __init__(self, name="dummy"):
        self._n = name

    def __call__(self, *a, **k):
        return _Dummy(self._n + "()")

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Dummy(self._n + "." + k)


class _StubModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Dummy(self.__name__ + "." + k)


def stub(name, **attrs):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            m = _StubModule(n)
            m.__path__ = []
            sys.modules[n] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)
    m = sys.modules[name]
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


class _Reg:
    def __init__(self):
        self.d = {}

    def register_module(self, *a, **k):
        def deco(cls):
            self.d[cls.__name__] = cls
            return cls
        return deco

    def build(self, cfg):
        cfg = dict(cfg)
        return self.d[cfg.pop("type")](**cfg)


def _identity_decorator(*a, **k):
    if len(a) == 1 and callable(a[0]) and not k:
        return a[0]
    return lambda f: f


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


ATTENTION, TLS, TRANSFORMER = _Reg(), _Reg(), _Reg()
HEADS, DETECTORS, BACKBONES, NECKS, BBOX_CODERS, MIDDLE = _Reg(), _Reg(), _Reg(), _Reg(), _Reg(), _Reg()


class StubMHA(nn.Module):
    """[restated] mmcv MultiheadAttention wrapper: q=k=query+pos, v=query, +identity."""

    def __init__(self, embed_dims, num_heads, dropout=0.0, **kw):
        super().__init__()
        self.embed_dims = embed_dims
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, dropout)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, **kw):
        qk = query + query_pos
        return query + self.attn(qk, qk, value=query)[0]


class StubFFN(nn.Module):
    """[restated] mmcv FFN."""

    def __init__(self, embed_dims, feedforward_channels, **kw):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(0.)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(0.))

    def forward(self, x, identity=None):
        return x + self.layers(x)


class StubLayer(nn.Module):
    """[restated] mmcv BaseTransformerLayer, order (self_attn,norm,cross_attn,norm,ffn,norm)."""

    def __init__(self, attn_cfgs, ffn_cfgs, operation_order, norm_cfg=None, **kw):
        super().__init__()
        assert tuple(operation_order) == ("self_attn", "norm", "cross_attn", "norm", "ffn", "norm")
        a0, a1 = dict(attn_cfgs[0]), dict(attn_cfgs[1])
        a0.pop("type")
        self.attentions = nn.ModuleList([StubMHA(**a0), ATTENTION.build(a1)])
        f = dict(ffn_cfgs)
        f.pop("type", None)
        self.ffns = nn.ModuleList([StubFFN(**f)])
        self.embed_dims = a0["embed_dims"]
        self.norms = nn.ModuleList([nn.LayerNorm(self.embed_dims) for _ in range(3)])

    def forward(self, query, key=None, value=None, query_pos=None, **kwargs):
        query = self.attentions[0](query, query_pos=query_pos)
        query = self.norms[0](query)
        query = self.attentions[1](query, key, value, None, query_pos=query_pos, **kwargs)
        query = self.norms[1](query)
        query = self.ffns[0](query)
        return self.norms[2](query)


class TransformerLayerSequence(BaseModule):
    """[restated] mmcv TransformerLayerSequence: builds `num_layers` layers."""

    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__(init_cfg)
        self.num_layers = num_layers
        self.layers = nn.ModuleList()
        for _ in range(num_layers):
            c = dict(transformerlayers)
            c.pop("type")
            self.layers.append(StubLayer(**c))
        self.embed_dims = self.layers[0].embed_dims


def inverse_sigmoid(x, eps=1e-5):
    """[restated] mmdet.models.utils.transformer.inverse_sigmoid."""
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


def build_conv_layer(cfg, *args, **kwargs):
    """[restated] mmcv build_conv_layer for type Conv3d."""
    cfg = dict(cfg)
    t = cfg.pop("type")
    assert t == "Conv3d"
    return nn.Conv3d(*args, **kwargs, **cfg)


def build_norm_layer(cfg, num_features):
    cfg = dict(cfg)
    cfg.pop("type")
    return "bn", nn.BatchNorm3d(num_features, **cfg)


def build_upsample_layer(cfg, *args, **kwargs):
    cfg = dict(cfg)
    assert cfg.pop("type") == "deconv3d"
    return nn.ConvTranspose3d(*args, **kwargs, **cfg)


class AttrDict(dict):
    __getattr__ = dict.__getitem__

    def pop(self, *a):
        return dict.pop(self, *a)


def install_stubs():
    stub("mmcv")
    stub("mmcv.cnn", xavier_init=lambda m, distribution="uniform", bias=0.: (
        nn.init.xavier_uniform_(m.weight), nn.init.constant_(m.bias, bias)),
        constant_init=lambda m, val, bias=0.: (nn.init.constant_(m.weight, val),
                                               nn.init.constant_(m.bias, bias)),
        Linear=nn.Linear, build_conv_layer=build_conv_layer, build_norm_layer=build_norm_layer,
        build_upsample_layer=build_upsample_layer)
    stub("mmcv.cnn.bricks.registry", ATTENTION=ATTENTION, TRANSFORMER_LAYER_SEQUENCE=TLS)
    stub("mmcv.cnn.bricks.transformer", MultiScaleDeformableAttention=type("MSDA", (), {}),
         TransformerLayerSequence=TransformerLayerSequence,
         build_transformer_layer_sequence=lambda cfg: TLS.build(cfg))
    stub("mmcv.runner", force_fp32=_identity_decorator, auto_fp16=_identity_decorator,
         BaseModule=BaseModule)
    stub("mmcv.runner.base_module", BaseModule=BaseModule)
    stub("mmcv.ops")
    stub("mmdet.models", HEADS=HEADS, DETECTORS=DETECTORS, BACKBONES=BACKBONES, NECKS=NECKS)
    stub("mmdet.models.utils.builder", TRANSFORMER=TRANSFORMER)
    stub("mmdet.models.utils.transformer", inverse_sigmoid=inverse_sigmoid)
    stub("mmdet.models.dense_heads", DETRHead=type("DETRHead", (nn.Module,), {}))
    stub("mmdet.core")
    stub("mmdet.core.bbox", BaseBBoxCoder=object)
    stub("mmdet.core.bbox.builder", BBOX_CODERS=BBOX_CODERS)
    m3d = stub("mmdet3d")
    m3d.__version__ = "1.0.0rc5"
    stub("mmdet3d.core")
    stub("mmdet3d.core.bbox")
    stub("mmdet3d.core.bbox.coders")
    stub("mmdet3d.core.bbox.iou_calculators.iou3d_calculator")
    stub("mmdet3d.models.builder", MIDDLE_ENCODERS=MIDDLE)
    stub("mmdet3d.models.detectors.mvx_two_stage", MVXTwoStageDetector=type("MVX", (nn.Module,), {}))
    stub("mmdet3d.ops")
    stub("mmdet3d.ops.spconv", IS_SPCONV2_AVAILABLE=False)
    stub("symbol")  # dead import in second_3d.py:2 (module removed in Python 3.10)
    for n in ("projects", "projects.mmdet3d_plugin", "projects.mmdet3d_plugin.core",
              "projects.mmdet3d_plugin.core.bbox", "projects.mmdet3d_plugin.core.merge_all_augs"):
        stub(n)


def load_ref(relpath, modname):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


# ---------------------------------------------------------------- fixtures -----
def dump_configs():
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from uni3detr_b200.compat import Config
    from uni3detr_b200.synth import CONFIG_FILES

    def plain(o):
        if isinstance(o, dict):
            return {k: plain(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [plain(v) for v in o]
        return o
    out = {}
    for name, fn in CONFIG_FILES.items():
        cfg = Config.fromfile(os.path.join(REF, "projects/configs/uni3detr", fn))
        out[name] = plain(cfg.model)
    with open(os.path.join(HERE, "configs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def load_sd(module, seed):
    sd = gen_state_dict({k: tuple(v.shape) for k, v in module.state_dict().items()}, seed)
    module.load_state_dict(sd)
    return sd


def main():
    torch.set_num_threads(1)
    torch.manual_seed(0)
    install_stubs()
    util = load_ref("projects/mmdet3d_plugin/core/bbox/util.py", "projects.mmdet3d_plugin.core.bbox.util")
    tr = load_ref("projects/mmdet3d_plugin/models/utils/uni3detr_transformer.py", "ref_transformer")
    coder = load_ref("projects/mmdet3d_plugin/core/bbox/coders/nms_free_coder.py", "ref_coder")
    head = load_ref("projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py", "ref_head")
    det = load_ref("projects/mmdet3d_plugin/models/detectors/uni3detr.py", "ref_detector")
    bb = load_ref("projects/mmdet3d_plugin/models/backbones/second_3d.py", "ref_second3d")
    nk = load_ref("projects/mmdet3d_plugin/models/necks/second3d_fpn.py", "ref_fpn")
    g = torch.Generator().manual_seed(1234)
    R = lambda *s: torch.randn(*s, generator=g)
    out = {}

    # 1. sine embedding + shift_scale_points + denormalize_bbox
    pos = torch.rand(2, 5, 3, generator=g)
    out["sine_in"], out["sine_out"] = pos, tr.get_sine_pos_embed(pos)
    pts = R(2, 9, 3)
    out["ssp_in"] = pts
    out["ssp_out"] = det.shift_scale_points(pts, src_range=[pts.min(dim=1)[0], pts.max(dim=1)[0]])
    nb = R(6, 8)
    out["denorm_in"], out["denorm_out"] = nb, util.denormalize_bbox(nb, None)
    nb10 = R(4, 10)
    out["denorm10_in"], out["denorm10_out"] = nb10, util.denormalize_bbox(nb10, None)

    # 2. UniCrossAtten.forward (reference class, eval)
    ca = tr.UniCrossAtten(embed_dims=256, num_heads=8, num_points=1, dropout=0.1).eval()
    load_sd(ca, 11)
    B, nq, D, H, W = 2, 6, 3, 4, 5
    value = R(B, 1, 256, D, H, W)
    query, qpos = R(nq, B, 256), R(nq, B, 256)
    ref = R(B, nq, 3) * 1.5
    with torch.no_grad():
        out["ca_value"], out["ca_query"], out["ca_qpos"], out["ca_ref"] = value, query, qpos, ref
        out["ca_out"] = ca(query, None, value, query_pos=qpos, reference_points=ref)

    # 3. decoder + transformer + head forward (reference classes; mmcv layer wiring restated)
    layer_cfg = dict(type="BaseTransformerLayer",
                     attn_cfgs=[dict(type="MultiheadAttention", embed_dims=256, num_heads=8, dropout=0.1),
                                dict(type="UniCrossAtten", num_points=1, embed_dims=256, num_sweeps=1)],
                     ffn_cfgs=dict(type="FFN", embed_dims=256, feedforward_channels=64, num_fcs=2,
                                   ffn_drop=0.1, act_cfg=dict(type="ReLU", inplace=True)),
                     norm_cfg=dict(type="LN"),
                     operation_order=("self_attn", "norm", "cross_attn", "norm", "ffn", "norm"))
    tcfg = dict(decoder=dict(type="Uni3DETRTransformerDecoder", num_layers=2, return_intermediate=True,
                             transformerlayers=layer_cfg))
    transformer = tr.Uni3DETRTransformer(**tcfg).eval()

    class HeadShell(nn.Module):
        pass
    hs = HeadShell()
    nq, ncls, code = 5, 3, 8
    hs.num_query, hs.with_box_refine = nq, True
    hs.pc_range = [-3.2, -0.2, -2., 3.2, 6.2, 0.56]
    hs.transformer = transformer
    hs.embed_dims, hs.num_reg_fcs, hs.cls_out_channels, hs.code_size = 256, 2, ncls, code
    hs.as_two_stage = False
    head.Uni3DETRHead._init_layers(hs)
    hs.eval()
    load_sd(hs, 21)
    B = 2
    feats = R(B, 256, D, H, W)
    fps = torch.rand(B, 2 * nq, 3, generator=g)
    torch.manual_seed(77)
    rp = torch.rand(fps.shape)[:, :nq, :]
    torch.manual_seed(77)
    with torch.no_grad():
        outs = head.Uni3DETRHead.forward(hs, feats, None, fps)
    out["head_feats"], out["head_fps"], out["head_rand"] = feats, fps, rp
    for k, v in outs.items():
        out["head_" + k] = v
    out["head_meta"] = np.array([nq, ncls, code, 2])

    # 4. NMSFreeCoder.decode
    cd = coder.NMSFreeCoder(pc_range=hs.pc_range, post_center_range=hs.pc_range, max_num=12,
                            alpha=0.2, num_classes=ncls)
    preds = cd.decode({k: v.clone() for k, v in outs.items()})
    for i, p in enumerate(preds):
        for k, v in p.items():
            out[f"coder_{i}_{k}"] = v

    # 5. SECOND3D + SECOND3DFPN (reference classes, builders restated)
    bcfg = dict(in_channels=[8, 8, 8], out_channels=[4, 8, 16], layer_nums=[2, 2, 2],
                layer_strides=[1, 2, 4], is_cascade=False,
                norm_cfg=dict(type="BN3d", eps=1e-3, momentum=0.01),
                conv_cfg=AttrDict(type="Conv3d", kernel=(1, 3, 3), bias=False))
    backbone = bb.SECOND3D(**bcfg).eval()
    load_sd(backbone, 31)
    ncfg = dict(in_channels=[4, 8, 16], out_channels=[8, 8, 8], upsample_strides=[1, 2, 4],
                norm_cfg=dict(type="BN3d", eps=1e-3, momentum=0.01),
                upsample_cfg=dict(type="deconv3d", bias=False),
                extra_conv=dict(type="Conv3d", num_conv=2, bias=False), use_conv_for_no_stride=True)
    neck = nk.SECOND3DFPN(**ncfg).eval()
    load_sd(neck, 41)
    xin = R(2, 8, 3, 8, 8)
    with torch.no_grad():
        xs = backbone(xin)
        y = neck(xs)
    out["dense_in"], out["dense_out"] = xin, y
    for i, t in enumerate(xs):
        out[f"dense_bb{i}"] = t

    np.savez_compressed(os.path.join(HERE, "golden_firstparty.npz"),
                        **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                           for k, v in out.items()})

    # 6. SparseEncoderHD construction (reference make_encoder_layers with recording stubs)
    calls = []

    def make_sparse_convmodule(in_channels, out_channels, kernel_size, indice_key=None, stride=1,
                               padding=0, conv_type="SubMConv3d", norm_cfg=None,
                               order=("conv", "norm", "act")):
        cin, cout, k = in_channels, out_channels, kernel_size
        calls.append(dict(kind="module", cin=cin, cout=cout, k=k, stride=stride, padding=padding,
                          conv_type=conv_type, indice_key=indice_key))
        return nn.Identity()

    class SparseBasicBlock(nn.Module):
        def __init__(self, cin, cout, norm_cfg=None, conv_cfg=None):
            super().__init__()
            calls.append(dict(kind="block", cin=cin, cout=cout))

    stub("mmdet3d.ops", SparseBasicBlock=SparseBasicBlock, make_sparse_convmodule=make_sparse_convmodule)
    stub("mmcv.ops", SparseConvTensor=object, SparseSequential=nn.Sequential)
    enc = load_ref("projects/mmdet3d_plugin/models/pts_encoder/sparse_encoder_hd.py", "ref_encoder")
    with open(os.path.join(HERE, "configs.json")) as f:
        cfgs = json.load(f)
    layer_lists = {}
    for name, mc in cfgs.items():
        calls.clear()
        c = dict(mc["pts_middle_encoder"])
        c.pop("type")
        c["order"] = tuple(c["order"])
        enc.SparseEncoderHD(**c)
        layer_lists[name] = json.loads(json.dumps(calls))
    with open(os.path.join(HERE, "golden_encoder_layers.json"), "w") as f:
        json.dump(layer_lists, f, indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    dump_configs()
    main()
